#!/usr/bin/env python
"""Golden vectors for the Qwen2.5-VL *target*  from the UNMODIFIED reference module
``vispec/model/modeling_qwen2_5_vl_kv.py``: GQA, q/k/v biases, multimodal RoPE over an image grid in the prefill, SDPA
attention with the tree mask, decode positions offset by ``rope_deltas`` (utils.py:398-402).

Run in the build container only:   python tests/golden/make_golden_qwen.py

Caller-side shims (nothing in the reference is edited): the ``SlidingWindowCache`` name (SURVEY 8c shim 1); a
``"default"`` entry for ``ROPE_INIT_FUNCTIONS`` (transformers 5.x dropped it; the function below is the textbook
``1 / theta^(2i/d)``); ``config.rope_theta`` set as an attribute (5.x moved it into ``rope_parameters``); the KV cache
built by hand on the CPU (shim 3)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import transformers.cache_utils as _cu  # noqa: E402

if not hasattr(_cu, "SlidingWindowCache"):
    _cu.SlidingWindowCache = type("SlidingWindowCache", (), {})

from transformers.models.qwen2_5_vl.configuration_qwen2_5_vl import Qwen2_5_VLTextConfig  # noqa: E402
from vispec.model import kv_cache  # noqa: E402
from vispec.model import modeling_qwen2_5_vl_kv as mq  # noqa: E402

from oracle import synth  # noqa: E402
from oracle.vispec_oracle import TargetCfg  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
QWEN_TINY = TargetCfg(hidden=512, inter=1024, layers=2, heads=4, kv_heads=2, vocab=1000, max_pos=512, eps=1e-6,
                      rope_theta=1e6, qkv_bias=True, mrope_section=(16, 24, 24), attn_impl="sdpa")


def _default_rope(config, device=None, seq_len=None, **kw):
    dim = config.hidden_size // config.num_attention_heads
    inv = 1.0 / (config.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float().to(device) / dim))
    return inv, 1.0


def build_ref(tcfg: TargetCfg, weights):
    mq.ROPE_INIT_FUNCTIONS.setdefault("default", _default_rope)
    cfg = Qwen2_5_VLTextConfig(vocab_size=tcfg.vocab, hidden_size=tcfg.hidden, intermediate_size=tcfg.inter,
                               num_hidden_layers=tcfg.layers, num_attention_heads=tcfg.heads,
                               num_key_value_heads=tcfg.kv_heads, max_position_embeddings=tcfg.max_pos,
                               rms_norm_eps=tcfg.eps, use_sliding_window=False,
                               rope_scaling={"type": "mrope", "mrope_section": list(tcfg.mrope_section),
                                             "rope_type": "default"})
    cfg.rope_theta = tcfg.rope_theta
    cfg._attn_implementation = "sdpa"
    m = mq.Qwen2_5_VLModel(cfg).to(torch.bfloat16).eval()
    # `from_pretrained(torch_dtype=...)` leaves the non-persistent fp32 `inv_freq` buffers alone; `.to(bfloat16)` here
    # would round them, which is an artefact of this harness and not of the reference: put the fp32 values back
    for mod in m.modules():
        if isinstance(mod, mq.Qwen2_5_VLRotaryEmbedding):
            mod.inv_freq = _default_rope(cfg)[0]
            mod.original_inv_freq = mod.inv_freq
    sd = {k[len("model."):]: v for k, v in weights.items() if k.startswith("model.")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in k for k in missing), (missing, unexpected)
    return m


def image_positions(n_pre: int, grid: int, n_post: int):
    """Position ids of `get_rope_index` (modeling_qwen2_5_vl_kv.py:1795-1990) for one image of grid x grid merged
    patches between n_pre and n_post text tokens: text = same id on the three axes; image = (t, h, w) offsets from the
    running start; the text after it continues from max + 1.  Returns ([3, L], rope_delta = max + 1 - L)."""
    t = torch.arange(n_pre)
    pre = torch.stack([t, t, t])
    st = n_pre
    hh = torch.arange(grid).view(-1, 1).expand(grid, grid).flatten()
    ww = torch.arange(grid).view(1, -1).expand(grid, grid).flatten()
    img = torch.stack([torch.zeros(grid * grid, dtype=torch.long), hh, ww]) + st
    st2 = int(img.max()) + 1
    t2 = torch.arange(n_post) + st2
    post = torch.stack([t2, t2, t2])
    pos = torch.cat([pre, img, post], dim=1)
    L = pos.shape[1]
    return pos, int(pos.max()) + 1 - L


def main():
    tcfg = QWEN_TINY
    w = synth.make_target_weights(tcfg, seed=21)
    ref = build_ref(tcfg, w)
    g = torch.Generator().manual_seed(5)
    pos3, delta = image_positions(6, 4, 18)
    L = pos3.shape[1]
    emb = (torch.randn(1, L, tcfg.hidden, generator=g) * 0.5).to(torch.bfloat16)
    data = torch.zeros(2 * tcfg.layers, 1, tcfg.kv_heads, 256, 128, dtype=torch.bfloat16)
    lens = torch.zeros(2 * tcfg.layers, dtype=torch.long)
    pkv = [[kv_cache.KVCache(data[2 * i + j], lens[2 * i + j]) for j in (0, 1)] for i in range(tcfg.layers)]
    with torch.no_grad():
        out = ref(inputs_embeds=emb, position_ids=pos3[:, None, :], past_key_values=pkv, use_cache=True)
    prefill_hidden = out.last_hidden_state.clone()
    kv_after_prefill = data[:, :, :, :L].clone()
    # ---- tree verify step: 11 nodes, parents give the ancestor mask and the depths (utils.py:389-412)
    parents = [-1, 0, 0, 0, 1, 1, 2, 4, 4, 6, 7]
    T = len(parents)
    depth = torch.zeros(T, dtype=torch.long)
    tm = torch.zeros(T, T)
    for i, p in enumerate(parents):
        tm[i, i] = 1
        if p >= 0:
            depth[i] = depth[p] + 1
            tm[i] = torch.maximum(tm[i], tm[p])
    tree_ids = torch.randint(1, tcfg.vocab, (1, T), generator=g)
    ref.tree_mask = tm[None, None]
    position_ids = (depth + L)[None] + delta                       # utils.py:399-401
    position_ids = position_ids[None].expand(3, -1, -1)            # utils.py:402
    with torch.no_grad():
        out2 = ref(input_ids=tree_ids, position_ids=position_ids, past_key_values=pkv, use_cache=True)
    tree_hidden = out2.last_hidden_state.clone()
    assert int(lens[0]) == L + T
    # ---- one plain decode token after rolling the lengths back (the baseline loop)
    ref.tree_mask = None
    lens.fill_(L)
    one = tree_ids[:, :1]
    p1 = torch.tensor([[L + delta]])[None].expand(3, -1, -1)
    with torch.no_grad():
        out3 = ref(input_ids=one, position_ids=p1, past_key_values=pkv, use_cache=True)
    lm = w["lm_head.weight"]
    torch.save({
        "cfg": dict(hidden=tcfg.hidden, inter=tcfg.inter, layers=tcfg.layers, heads=tcfg.heads, kv_heads=tcfg.kv_heads,
                    vocab=tcfg.vocab, max_pos=tcfg.max_pos, eps=tcfg.eps, rope_theta=tcfg.rope_theta, qkv_bias=True,
                    mrope_section=tuple(tcfg.mrope_section), attn_impl="sdpa"),
        "seed": 21, "inputs_embeds": emb, "pos3": pos3, "rope_delta": delta, "tree_ids": tree_ids, "tree_mask": tm,
        "tree_depth": depth, "prefill_hidden": prefill_hidden, "kv_after_prefill_sum": kv_after_prefill.float().sum(dim=(1, 2, 4)),
        "tree_hidden": tree_hidden, "tree_logits_argmax": torch.nn.functional.linear(tree_hidden, lm).float().argmax(-1),
        "one_hidden": out3.last_hidden_state.clone(),
    }, os.path.join(OUT, "qwen_target.pt"))
    print("wrote qwen_target.pt: L", L, "delta", delta, "tree", T)


if __name__ == "__main__":
    main()
