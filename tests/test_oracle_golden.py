"""Pin the oracle: oracle/vispec_oracle.py must reproduce what the unmodified reference returned
(tests/golden/*.pt, written by tests/golden/make_golden.py) on the same seeded inputs."""
import glob
import os

import pytest
import torch

from oracle import synth
from oracle import vispec_oracle as vo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN, "*.pt")) if "sampling" not in p and "qwen" not in p)


def load_case(name):
    g = torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)
    m = g["meta"]
    tcfg, dcfg = synth.TINY_T, synth.TINY_D
    if m["kind"] == "random":
        tw = synth.make_target_weights(tcfg, m["seed"])
        dw = synth.make_draft_weights(dcfg, m["seed"] + 1)
    else:
        tw, dw = synth.make_bigram_fixture(tcfg, dcfg, m["seed"], fc_noise=m["fc_noise"])
    ids, emb, mask = synth.make_prompt(tcfg, tw, m["seed"] + 7, m["n_pre"], m["n_img"], m["n_post"])
    return g, tcfg, dcfg, tw, dw, ids, (emb if mask is not None else None), mask


def test_golden_files_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    torch.set_num_threads(8)
    g, tcfg, dcfg, tw, dw, ids, emb, mask = load_case(name)
    target, draft = vo.TargetOracle(tcfg, tw), vo.DraftOracle(dcfg, dw)
    trace = vo.SpecTrace()
    n_cycles = g["meta"]["n_cycles"]
    out, new_token, idx, acc = vo.spec_generate(target, draft, ids, emb, mask, max_new_tokens=10 ** 9,
                                                max_cycles=n_cycles, trace=trace)
    assert acc == [c["accept_length"] for c in g["cycles"]]
    assert torch.equal(out, g["output_ids"])
    assert new_token == g["new_token"]
    # per-cycle tree structures: cycle i verified the tree built at the end of cycle i-1 (first: prefill)
    trees = [g["first"]] + g["cycles"][:-1]
    for tr, ref in zip(trace.cycles, trees):
        assert torch.equal(tr["draft_tokens"], ref["draft_tokens"])
        assert torch.equal(tr["retrieve"], ref["retrieve_indices"])
        assert torch.equal(tr["tree_mask"], ref["tree_mask"])
        assert torch.equal(tr["tree_pos"], ref["tree_position_ids"])
    for tr, ref in zip(trace.cycles, g["cycles"]):
        assert tr["best"] == ref["best_candidate"]
        assert tr["token"] == int(ref["sample_token"])
        assert torch.equal(tr["logits_root"], ref["tree_logits_root"])       # same ATen ops -> exact on CPU
    assert target.cur_len == g["cycles"][-1]["kv_len"]
    assert draft.stable_kv[0].shape[2] == g["cycles"][-1]["draft_kv_len"]
    assert draft.stable_kv[2] == g["cycles"][-1]["draft_real_len"]
    kv_sum = target.kv[..., :target.cur_len, :].float().sum(dim=(1, 2, 3, 4))
    assert torch.equal(kv_sum, g["kv_checksum"])


@pytest.mark.parametrize("name", CASES)
def test_spec_equals_greedy(name):
    """The invariant of SURVEY.md 7.2: speculative output == plain greedy decode of the target."""
    g, tcfg, dcfg, tw, dw, ids, emb, mask = load_case(name)
    assert g["spec_equals_greedy"]
    target = vo.TargetOracle(tcfg, tw)
    n_new = g["output_ids"].shape[1] - ids.shape[1]
    out = vo.greedy_generate(target, ids, emb, n_new)
    assert torch.equal(out, g["greedy_ids"])
    assert torch.equal(out, g["output_ids"])


def test_oracle_sampling_acceptance_matches_reference():
    """T > 0: the oracle's restatement of utils.py:453-493 must reproduce what the unmodified
    reference returned for the same uniforms (tests/golden/make_golden_sampling.py)."""
    g = torch.load(os.path.join(GOLDEN, "posterior_sampling.pt"), weights_only=False)
    assert len(g["cases"]) >= 24
    seen = set()
    for c in g["cases"]:
        tree_logits = c["tree_logits"].float()
        retrieve, draft_tokens = c["retrieve"], c["draft_tokens"]
        logits = tree_logits[retrieve]
        pad = torch.full((1, 1), -1, dtype=torch.long)
        cand = torch.cat((draft_tokens, pad), dim=1)[0, retrieve]
        best, acc, sample_p = vo.evaluate_posterior_sampling(logits, cand, c["temperature"], c["uni"])
        assert (best, acc) == (c["best"], c["accept_length"])
        assert torch.equal(sample_p, c["sample_p"])
        seen.add(acc)
    assert len(seen) >= 4          # the fixture covers rejections at the root up to full-depth accepts


def test_qwen_target_oracle_matches_reference_golden():
    """Qwen2.5-VL target: GQA + q/k/v biases + multimodal RoPE in the prefill + SDPA with the tree
    mask + rope_delta decode offsets, against the unmodified `modeling_qwen2_5_vl_kv.Qwen2_5_VLModel` run on CPU
    (tests/golden/make_golden_qwen.py)."""
    g = torch.load(os.path.join(GOLDEN, "qwen_target.pt"), weights_only=False)
    tcfg = vo.TargetCfg(**g["cfg"])
    w = synth.make_target_weights(tcfg, seed=g["seed"])
    t = vo.TargetOracle(tcfg, w, kv_max=256)
    L = g["pos3"].shape[1]
    _, hid = t.forward(inputs_embeds=g["inputs_embeds"], position_ids=g["pos3"])
    assert torch.equal(hid, g["prefill_hidden"])
    assert torch.equal(t.kv[:, :, :, :L].float().sum(dim=(1, 2, 4)), g["kv_after_prefill_sum"])
    pos = g["tree_depth"] + L + g["rope_delta"]
    logits, hid2 = t.forward(input_ids=g["tree_ids"], position_ids=pos, tree_mask=g["tree_mask"][None, None])
    assert torch.equal(hid2, g["tree_hidden"])
    assert torch.equal(logits.argmax(-1), g["tree_logits_argmax"])
    t.cur_len = L
    _, hid3 = t.forward(input_ids=g["tree_ids"][:, :1], position_ids=torch.tensor([L + g["rope_delta"]]))
    assert torch.equal(hid3, g["one_hidden"])


def test_warp_logits_matches_reference_processor_list():
    """oracle.warp_logits (temperature, nucleus, top-k) against the reference's prepare_logits_processor output on seeded
    logits incl. a run of exact ties (tests/golden/make_golden_warpers.py).  The same fixture records that the reference's
    nucleus filter raises on the 3-D tree logits of evaluate_posterior, which is why only the temperature warper is on the
    product path."""
    import os
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warpers_sampling.pt"), weights_only=False)
    assert len(g["cases"]) >= 7
    for c in g["cases"]:
        kw = c["kw"]
        got = vo.warp_logits(g["x"], kw["temperature"], kw.get("top_p", 0.0), kw.get("top_k", 0))
        assert torch.equal(got, c["out"]), kw
    assert g["observations"]["top_k_3d"] == "ok" and g["observations"]["top_p_3d"].startswith("RuntimeError")
