"""Qwen2.5-VL target shapes through the CUDA path: GQA, q/k/v biases, theta = 1e6, multimodal
RoPE in the prefill, SDPA attention semantics with the tree mask, decode positions offset by rope_deltas -- against the
oracle and the golden vectors of the unmodified reference module (tests/golden/make_golden_qwen.py)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth  # noqa: E402  (checker only)
from oracle import vispec_oracle as vo  # noqa: E402
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import DraftConfig, TargetConfig  # noqa: E402
from vispec_b200.model import utils as vu  # noqa: E402
from vispec_b200.model.kv_cache import initialize_past_key_values  # noqa: E402
from vispec_b200.model.spec_model_ours import SpecModel  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close_bf16(a, b, frac=0.0, ulps=2):
    """Two layers of bf16 GEMMs on different devices (the golden is the reference on the CPU; here cuBLAS / flash SDPA
    for the prefill and our kernels for the decode): a few bf16 ulps of accumulation-order noise per element are
    expected, a wrong position / bias / head mapping shows up as O(1) errors."""
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    assert float(err.max()) <= 0.08 + 0.04 * float(b.abs().max()), float(err.max())
    assert float(err.mean()) <= 0.01, float(err.mean())
    cos = torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0)
    assert float(cos) >= 0.9995, float(cos)


def _model(g, total_tokens=12, depth=3, top_k=4):
    tc = vo.TargetCfg(**g["cfg"])
    tw = synth.make_target_weights(tc, seed=g["seed"])
    dc = vo.DraftCfg(hidden=tc.hidden, inter=tc.inter, heads=tc.heads, vocab=tc.vocab, max_pos=tc.max_pos, eps=tc.eps,
                     rope_theta=tc.rope_theta, total_tokens=total_tokens, depth=depth, top_k=top_k, num_q=2, bias=True)
    dw = synth.make_draft_weights(dc, seed=7)
    tcfg = TargetConfig(hidden=tc.hidden, inter=tc.inter, layers=tc.layers, heads=tc.heads, kv_heads=tc.kv_heads,
                        vocab=tc.vocab, max_pos=tc.max_pos, eps=tc.eps, rope_theta=tc.rope_theta,
                        mrope_section=tuple(tc.mrope_section), attn_impl=tc.attn_impl)
    dcfg = DraftConfig(hidden=dc.hidden, inter=dc.inter, heads=dc.heads, vocab=dc.vocab, max_pos=dc.max_pos, eps=dc.eps,
                       rope_theta=dc.rope_theta, total_tokens=dc.total_tokens, depth=dc.depth, top_k=dc.top_k,
                       num_q=dc.num_q, bias=dc.bias)
    model = SpecModel.from_state_dicts(tcfg, tw, dcfg, dw, device=DEV, kv_max=256, eos_token_id=None)
    return tc, tw, model


def test_qwen_target_prefill_and_tree_verify_match_reference_golden():
    g = torch.load(os.path.join(GOLDEN, "qwen_target.pt"), weights_only=False)
    tc, tw, model = _model(g)
    lm = model.base_model
    pkv, _, cur_len = initialize_past_key_values(lm)
    L0 = g["pos3"].shape[1]
    out = lm(inputs_embeds=g["inputs_embeds"].to(DEV), past_key_values=pkv, position_ids=g["pos3"][:, None, :].to(DEV),
             output_hidden_states=True)
    close_bf16(out.hidden_states[-1], g["prefill_hidden"], frac=0.90)
    assert lm.rope_deltas == g["rope_delta"] and int(cur_len[0]) == L0
    # tree verify (utils.py:389-412 with the Qwen position rule :398-402): positions on three identical axes
    lm.model.tree_mask = g["tree_mask"][None, None].to(DEV)
    pos = (g["tree_depth"] + L0 + g["rope_delta"])[None, None].expand(3, 1, -1).to(DEV)
    out2 = lm(input_ids=g["tree_ids"].to(DEV), past_key_values=pkv, position_ids=pos, output_hidden_states=True)
    close_bf16(out2.hidden_states[-1], g["tree_hidden"], frac=0.90)
    agree = (out2.logits.argmax(-1).cpu() == g["tree_logits_argmax"]).float().mean()
    assert float(agree) >= 0.9
    # plain decode of one token after rolling back: position = cache position + rope_deltas when none is passed
    lm.model.tree_mask = None
    vu.reset_tree_mode(model)
    cur_len.fill_(L0)
    out3 = lm(input_ids=g["tree_ids"][:, :1].to(DEV), past_key_values=pkv, output_hidden_states=True)
    close_bf16(out3.hidden_states[-1], g["one_hidden"], frac=0.90)


def test_qwen_target_specgenerate_equals_greedy_oracle():
    """Whole draft-and-verify loop on the Qwen-shaped target: the speculative output equals the oracle's plain greedy
    decode with multimodal prompt positions (lossless at T = 0), and drafts are accepted."""
    g = torch.load(os.path.join(GOLDEN, "qwen_target.pt"), weights_only=False)
    tc, tw, model = _model(g)
    gen = torch.Generator().manual_seed(11)
    n_pre, grid, n_post = 6, 4, 18
    ids = torch.randint(1, tc.vocab, (1, n_pre + grid * grid + n_post), generator=gen)
    IMG = 999
    ids[0, n_pre:n_pre + grid * grid] = IMG
    pos3, delta = vu.qwen_rope_index(ids, image_grid_thw=[[1, 2 * grid, 2 * grid]], image_token_id=IMG)
    assert torch.equal(pos3[:, 0], g["pos3"]) and delta == g["rope_delta"]
    emb = torch.nn.functional.embedding(ids, tw["model.embed_tokens.weight"])
    emb[0, n_pre:n_pre + grid * grid] = (torch.randn(grid * grid, tc.hidden, generator=gen) * 0.5).to(emb.dtype)
    mask = ids == IMG
    n_new = 24
    # oracle greedy with the same positions
    t = vo.TargetOracle(tc, tw, kv_max=256)
    logits, _ = t.forward(inputs_embeds=emb, position_ids=pos3[:, 0])
    want = []
    tok = int(logits[0, -1].argmax())
    for _ in range(n_new):
        want.append(tok)
        logits, _ = t.forward(input_ids=torch.tensor([[tok]]), position_ids=torch.tensor([t.cur_len + delta]))
        tok = int(logits[0, -1].argmax())
    out, new_token, idx, acc = model.specgenerate(ids.to(DEV), inputs_embeds=emb.to(DEV), image_mask=mask.to(DEV),
                                                  position_ids=pos3.to(DEV), max_new_tokens=n_new, log=True,
                                                  return_acceptance_len=True)
    got = out[0, ids.shape[1]:].tolist()
    n = min(len(got), len(want))
    agree = sum(int(a == b) for a, b in zip(got[:n], want[:n]))
    assert n >= n_new - 8 and agree == n, (got, want)


def test_qwen7b_shapes_spec_equals_greedy():
    """BASELINE config 3 shapes (Qwen2.5-VL-7B: H=3584, 28 q / 4 kv heads, I=18944, V=152064, 28 layers, q/k/v biases,
    theta=1e6) with a 24 x 24 image grid in the prompt: the speculative output equals plain greedy decoding of the same
    target through the same kernels (size-independent property; the CPU oracle would take minutes at this size)."""
    from bench import synth_draft_weights, synth_target_weights
    tcfg = TargetConfig(hidden=3584, inter=18944, layers=28, heads=28, kv_heads=4, vocab=152064, max_pos=8192, eps=1e-6,
                        rope_theta=1e6, mrope_section=(16, 24, 24), attn_impl="sdpa")
    dcfg = DraftConfig(hidden=3584, inter=18944, heads=28, vocab=152064, max_pos=4096, eps=1e-6, rope_theta=1e6)
    tw = synth_target_weights(tcfg, 3, DEV)
    g = torch.Generator(device=DEV).manual_seed(4)
    for li in range(tcfg.layers):
        for n, rows in (("q", 28 * 128), ("k", 4 * 128), ("v", 4 * 128)):
            tw[f"model.layers.{li}.self_attn.{n}_proj.bias"] = \
                (torch.randn(rows, generator=g, device=DEV) * 0.3).to(torch.bfloat16)
    dw = synth_draft_weights(dcfg, 5, DEV)
    model = SpecModel.from_state_dicts(tcfg, tw, dcfg, dw, device=DEV, kv_max=2048, eos_token_id=None)
    del tw, dw
    torch.cuda.empty_cache()
    n_pre, grid, n_post, IMG = 5, 24, 60, 151655
    cg = torch.Generator().manual_seed(9)
    ids = torch.randint(1, 150000, (1, n_pre + grid * grid + n_post), generator=cg)
    ids[0, n_pre:n_pre + grid * grid] = IMG
    pos3, delta = vu.qwen_rope_index(ids, image_grid_thw=[[1, 2 * grid, 2 * grid]], image_token_id=IMG)
    assert delta == n_pre + grid + n_post - ids.shape[1]
    emb = model.base_model_engine.embed_tokens(ids.to(DEV)).clone()
    emb[0, n_pre:n_pre + grid * grid] = (torch.randn(grid * grid, tcfg.hidden, generator=cg) * 0.02).to(torch.bfloat16).to(DEV)
    mask = (ids == IMG).to(DEV)
    n_tok = 16
    out, new_token, idx, acc = model.specgenerate(ids.to(DEV), inputs_embeds=emb, image_mask=mask, position_ids=pos3.to(DEV),
                                                  log=True, return_acceptance_len=True, max_new_tokens=n_tok)
    assert out.shape[1] - ids.shape[1] == new_token > n_tok
    assert int(model.base_model_engine.state[L.ST_ROPE_DELTA]) == delta
    pkv, _, cur_len = initialize_past_key_values(model.base_model)
    cur_len.zero_()
    lm = model.base_model
    o = lm(inputs_embeds=emb, past_key_values=pkv, position_ids=pos3.to(DEV))
    seq = []
    for _ in range(new_token):
        tok = torch.argmax(o.logits[:, -1])[None, None]
        seq.append(int(tok))
        o = lm(input_ids=tok, past_key_values=pkv)           # position = cache position + rope_deltas
    assert out[0, ids.shape[1]:].tolist() == seq
