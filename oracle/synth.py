"""Seeded synthetic weights and prompts for the oracle, the golden generator, tests and bench.

TEST INFRASTRUCTURE (see oracle/vispec_oracle.py header).  No checkpoints exist offline, so every
fixture is (config, seed) -> weights, regenerated identically wherever torch's CPU generator runs.
State-dict keys equal the reference's (``modeling_llama_kv.LlamaForCausalLM`` / ``cnets_ours.Model``),
so the same dict loads into the unmodified reference modules for golden generation.

Scaling follows SURVEY.md section 7.2: HF's sigma=0.02 init makes every bf16 draft log-prob tie, so
embeddings get sigma~1 and the LM head sigma~4/sqrt(H).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .vispec_oracle import DraftCfg, TargetCfg


def _randn(gen, *shape, std=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, generator=gen, dtype=torch.float32) * std).to(dtype)


def make_target_weights(cfg: TargetCfg, seed: int, dtype=torch.bfloat16, gain: float = 0.5) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    H, I, V, D = cfg.hidden, cfg.inter, cfg.vocab, cfg.head_dim
    w = {"model.embed_tokens.weight": _randn(g, V, H, std=1.0, dtype=dtype)}
    for li in range(cfg.layers):
        p = f"model.layers.{li}."
        w[p + "self_attn.q_proj.weight"] = _randn(g, cfg.heads * D, H, std=1.0 / math.sqrt(H), dtype=dtype)
        w[p + "self_attn.k_proj.weight"] = _randn(g, cfg.kv_heads * D, H, std=1.0 / math.sqrt(H), dtype=dtype)
        w[p + "self_attn.v_proj.weight"] = _randn(g, cfg.kv_heads * D, H, std=1.0 / math.sqrt(H), dtype=dtype)
        w[p + "self_attn.o_proj.weight"] = _randn(g, cfg.hidden, cfg.heads * D, std=gain / math.sqrt(H), dtype=dtype)
        w[p + "mlp.gate_proj.weight"] = _randn(g, I, H, std=1.0 / math.sqrt(H), dtype=dtype)
        w[p + "mlp.up_proj.weight"] = _randn(g, I, H, std=1.0 / math.sqrt(H), dtype=dtype)
        w[p + "mlp.down_proj.weight"] = _randn(g, H, I, std=gain / math.sqrt(I), dtype=dtype)
        if getattr(cfg, "qkv_bias", False):             # Qwen2.5-VL (drawn last per layer: older seeds keep their weights)
            for n, rows in (("q", cfg.heads * D), ("k", cfg.kv_heads * D), ("v", cfg.kv_heads * D)):
                w[p + f"self_attn.{n}_proj.bias"] = _randn(g, rows, std=0.3, dtype=dtype)
        w[p + "input_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype)
        w[p + "post_attention_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype)
    w["model.norm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype)
    w["lm_head.weight"] = _randn(g, V, H, std=4.0 / math.sqrt(H), dtype=dtype)
    return w


def make_draft_weights(cfg: DraftCfg, seed: int, dtype=torch.bfloat16, gain: float = 0.5) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    H, I, V, D = cfg.hidden, cfg.inter, cfg.vocab, cfg.head_dim
    w = {"embed_tokens.weight": _randn(g, V, H, std=1.0, dtype=dtype)}
    p = "layers.0."
    for n in ("q", "k", "v"):
        w[p + f"self_attn.{n}_proj.weight"] = _randn(g, H, H, std=1.0 / math.sqrt(H), dtype=dtype)
    w[p + "self_attn.o_proj.weight"] = _randn(g, H, H, std=gain / math.sqrt(H), dtype=dtype)
    w[p + "mlp.gate_proj.weight"] = _randn(g, I, H, std=1.0 / math.sqrt(H), dtype=dtype)
    w[p + "mlp.up_proj.weight"] = _randn(g, I, H, std=1.0 / math.sqrt(H), dtype=dtype)
    w[p + "mlp.down_proj.weight"] = _randn(g, H, I, std=gain / math.sqrt(I), dtype=dtype)
    w[p + "post_attention_layernorm.weight"] = (1.0 + 0.1 * torch.randn(H, generator=g)).to(dtype)
    w["fc.weight"] = _randn(g, H, 2 * H, std=1.0 / math.sqrt(2 * H), dtype=dtype)
    w["img_fc.weight"] = _randn(g, H, 2 * H, std=1.0 / math.sqrt(2 * H), dtype=dtype)
    if cfg.bias:
        w["fc.bias"] = _randn(g, H, std=0.1, dtype=dtype)
        w["img_fc.bias"] = _randn(g, H, std=0.1, dtype=dtype)
    w["imadpt.q"] = _randn(g, cfg.num_q, cfg.heads, D, std=D ** -0.5, dtype=dtype)
    w["imadpt.k_proj.weight"] = _randn(g, H, H, std=1.0 / math.sqrt(H), dtype=dtype)
    w["imadpt.v_proj.weight"] = _randn(g, H, H, std=1.0 / math.sqrt(H), dtype=dtype)
    w["imadpt.o_proj.weight"] = _randn(g, H, H, std=1.0 / math.sqrt(H), dtype=dtype)
    return w


def make_bigram_fixture(tcfg: TargetCfg, dcfg: DraftCfg, seed: int, fc_noise: float = 0.0,
                        dtype=torch.bfloat16):
    """tau>0 fixture: the target degenerates to a bigram table
    next = argmax(lm_head(norm(embed[tok]))) and the draft predicts it (perfectly when fc_noise=0)."""
    tw = make_target_weights(tcfg, seed, dtype)
    dw = make_draft_weights(dcfg, seed + 1, dtype)
    for li in range(tcfg.layers):
        tw[f"model.layers.{li}.self_attn.o_proj.weight"].zero_()
        tw[f"model.layers.{li}.mlp.down_proj.weight"].zero_()
    tw["model.norm.weight"].fill_(1.0)
    dw["embed_tokens.weight"] = tw["model.embed_tokens.weight"].clone()
    H = dcfg.hidden
    fc = torch.zeros(H, 2 * H, dtype=torch.float32)
    fc[:, :H] = torch.eye(H)
    if fc_noise > 0:
        g = torch.Generator().manual_seed(seed + 2)
        fc += fc_noise * torch.randn(H, 2 * H, generator=g) / math.sqrt(2 * H)
    dw["fc.weight"] = fc.to(dtype)
    if "fc.bias" in dw:
        dw["fc.bias"].zero_()
    dw["layers.0.self_attn.o_proj.weight"].zero_()
    dw["layers.0.mlp.down_proj.weight"].zero_()
    return tw, dw


def make_prompt(tcfg: TargetCfg, tw: Dict[str, torch.Tensor], seed: int, n_pre: int, n_img: int, n_post: int,
                img_std: float = 0.02, dtype=torch.bfloat16):
    """Synthetic (image-embedding + text) prompt: n_pre text ids, one run of n_img image rows
    ~ img_std*N(0,1), n_post text ids (BASELINE config 1/2 shape).  Image positions carry a
    placeholder id (vocab-1), as ``image_token_index`` does in the reference (spec_model_ours.py:357-359).
    Returns (input_ids [1,L] i64, inputs_embeds [1,L,H], image_mask [1,L] bool or None)."""
    g = torch.Generator().manual_seed(seed)
    L = n_pre + n_img + n_post
    ids = torch.randint(1, tcfg.vocab - 64, (1, L), generator=g)
    emb = torch.nn.functional.embedding(ids, tw["model.embed_tokens.weight"]).clone()
    if n_img == 0:
        return ids, emb, None
    mask = torch.zeros(1, L, dtype=torch.bool)
    mask[0, n_pre:n_pre + n_img] = True
    ids[mask] = tcfg.vocab - 1
    emb[0, n_pre:n_pre + n_img] = _randn(g, n_img, tcfg.hidden, std=img_std, dtype=dtype)
    return ids, emb, mask


def to_device(w: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    return {k: v.to(device) for k, v in w.items()}


# small shapes shared by the golden generator and the tests (head_dim 128 like the 7B models)
TINY_T = TargetCfg(hidden=256, inter=640, layers=2, heads=2, kv_heads=2, vocab=1000, max_pos=512)
TINY_D = DraftCfg(hidden=256, inter=640, heads=2, vocab=1000, max_pos=512)
