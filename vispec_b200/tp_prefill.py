"""Tensor-parallel target prefill (BASELINE config 4 / SURVEY.md 8e): the prompt pass of the target LM
(`LlamaModel.forward`, modeling_llama_kv.py:927-1080) sharded Megatron-style over the ranks of one node.

  * column-parallel q/k/v and gate/up: rank r owns heads [r*nh/N, (r+1)*nh/N) (with their kv heads) and a contiguous range
    of the 64-wide gate|up groups; attention runs on the local heads with no exchange;
  * row-parallel o_proj / down_proj: every rank multiplies its slice and the fp32 partial products are summed with ONE
    all-reduce per boundary (2 per layer, NCCL over NVLink/NVSwitch), then rounded to bf16 once -- the same rounding
    point as the single-device `nn.Linear`, so the result differs from it only by the fp32 summation order;
  * the K/V rows each rank produced for its kv heads are all-gathered once at the end, so every rank leaves with the
    complete KV slab: the draft head and the draft-and-verify decode stay replicated (one persistent kernel per GPU, no
    collective on the decode path), and all ranks decode in lockstep.

The reference has no tensor parallelism (its multi-GPU scheme is prompt sharding over Ray replicas); parity is against the
single-device prefill / the oracle.  The functions are device-agnostic torch code over a process group so that the
sharding and the collectives are covered by a world-size-2 `gloo` test on CPU; the product calls them with CUDA tensors and
an NCCL group (`TargetEngine.enable_tensor_parallel_prefill`)."""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of a balanced contiguous partition of n units (the remainder goes to the first ranks)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def check_shardable(heads: int, kv_heads: int, inter: int, world: int) -> None:
    if kv_heads % world or heads % world:
        raise ValueError(f"tensor-parallel prefill needs heads ({heads}) and kv heads ({kv_heads}) divisible by the "
                         f"world size ({world})")
    if inter % 64 or inter // 64 < world:
        raise ValueError("intermediate size must be a multiple of 64 with at least one 64-wide group per rank")


def _matmul_f32(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x [L, K] @ w[N, K]^T with the fp32 accumulator kept (bf16 inputs)."""
    if x.is_cuda:
        return torch.mm(x, w.t(), out_dtype=torch.float32)
    return x.float() @ w.float().t()


def _rms(t: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    tf = t.float()
    tf = tf * torch.rsqrt(tf.pow(2).mean(-1, keepdim=True) + eps)
    return w * tf.to(t.dtype)


def _rot(t: torch.Tensor) -> torch.Tensor:
    return torch.cat((-t[..., 64:], t[..., :64]), dim=-1)


@torch.no_grad()
def prefill_layers(x: torch.Tensor, layers: Sequence[Dict[str, torch.Tensor]], kv: torch.Tensor, *, heads: int,
                   kv_heads: int, inter: int, eps: float, cos: torch.Tensor, sin: torch.Tensor, past: int, sdpa: bool,
                   rank: int, world: int, group=None) -> torch.Tensor:
    """The decoder layers over the prompt rows `x` [L, H] (bf16), sharded as described in the module docstring.

    `layers[i]` holds the engine's packed weights: ln_in, w_qkv [(nh+2nkv)*128, H] (+ b_qkv), w_o [H, nh*128], ln_post,
    w_gateup [2I, H] (64 gate rows then 64 up rows per 128-row group), w_down [H, I].  `kv` is the reference slab
    [2*layers, 1, kv_heads, max_pos, 128]; rows [past, past+L) of ALL kv heads are valid on every rank on return.
    `cos` / `sin`: [1, 1, L, 128] rotary factors of the prompt positions (already combined for a multimodal RoPE).
    Returns the residual stream after the last layer [L, H] (identical on every rank)."""
    check_shardable(heads, kv_heads, inter, world)
    Lq, H = x.shape
    D = 128
    S = past + Lq
    h0, h1 = shard_bounds(heads, rank, world)
    k0, k1 = shard_bounds(kv_heads, rank, world)
    g0, g1 = shard_bounds(inter // 64, rank, world)
    nh_l, nkv_l = h1 - h0, k1 - k0
    rep = heads // kv_heads
    dtype = x.dtype
    mask = None
    if not (sdpa and past == 0):
        mask = torch.full((Lq, S), torch.finfo(torch.float32).min, device=x.device)
        mask = torch.triu(mask, diagonal=past + 1)[None, None]

    def all_reduce(t: torch.Tensor) -> torch.Tensor:
        if world > 1:
            dist.all_reduce(t, group=group)
        return t

    h = x
    for li, t in enumerate(layers):
        res = h
        xn = _rms(h, t["ln_in"], eps)
        w_qkv, b_qkv = t["w_qkv"], t.get("b_qkv")
        rows = [(h0 * D, h1 * D), (heads * D + k0 * D, heads * D + k1 * D),
                ((heads + kv_heads) * D + k0 * D, (heads + kv_heads) * D + k1 * D)]
        q, k, v = (F.linear(xn, w_qkv[a:b], None if b_qkv is None else b_qkv[a:b]) for a, b in rows)
        q = q.view(1, Lq, nh_l, D).transpose(1, 2)
        k = k.view(1, Lq, nkv_l, D).transpose(1, 2)
        v = v.view(1, Lq, nkv_l, D).transpose(1, 2)
        q = (q * cos) + (_rot(q) * sin)
        k = (k * cos) + (_rot(k) * sin)
        kv[2 * li, :, k0:k1, past:S] = k
        kv[2 * li + 1, :, k0:k1, past:S] = v
        K, V = kv[2 * li, :, k0:k1, :S], kv[2 * li + 1, :, k0:k1, :S]
        if rep != 1:
            K = K[:, :, None].expand(1, nkv_l, rep, S, D).reshape(1, nh_l, S, D)
            V = V[:, :, None].expand(1, nkv_l, rep, S, D).reshape(1, nh_l, S, D)
        if sdpa:       # modeling_qwen2_5_vl_kv.py:1153-1160
            ao = F.scaled_dot_product_attention(q, K, V, is_causal=True) if past == 0 else \
                F.scaled_dot_product_attention(q, K, V, attn_mask=mask.to(dtype))
        else:          # eager, modeling_llama_kv.py:602-623
            aw = torch.matmul(q, K.transpose(2, 3)) / math.sqrt(D)
            aw = F.softmax(aw + mask, dim=-1, dtype=torch.float32).to(dtype)
            ao = torch.matmul(aw, V)
        ao = ao.transpose(1, 2).reshape(Lq, nh_l * D)
        # row-parallel o_proj: fp32 partial products, one all-reduce, one rounding
        o = all_reduce(_matmul_f32(ao, t["w_o"][:, h0 * D:h1 * D])).to(dtype)
        h = res + o
        res = h
        xn = _rms(h, t["ln_post"], eps)
        gu = F.linear(xn, t["w_gateup"][g0 * 128:g1 * 128]).view(Lq, g1 - g0, 2, 64)
        act = (F.silu(gu[:, :, 0, :]) * gu[:, :, 1, :]).reshape(Lq, (g1 - g0) * 64)
        d = all_reduce(_matmul_f32(act, t["w_down"][:, g0 * 64:g1 * 64])).to(dtype)
        h = res + d
    if world > 1:
        # every rank leaves with the complete KV slab (decode is replicated): one all-gather of the local kv-head slices
        mine = kv[:, 0, k0:k1, past:S].contiguous()                     # [2*layers, nkv_l, L, 128]
        parts: List[torch.Tensor] = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        for r, part in enumerate(parts):
            a, b = shard_bounds(kv_heads, r, world)
            if r != rank:
                kv[:, 0, a:b, past:S] = part
    return h


# ---------------------------------------------------------------------------------------------------------------------
# the same pass on the library's own kernels (CUDA): sharded weights through the C ABI, bf16 all-reduce
# ---------------------------------------------------------------------------------------------------------------------
def shard_layer_weights(layers: Sequence[Dict[str, torch.Tensor]], *, heads: int, kv_heads: int, inter: int, rank: int,
                        world: int) -> List[Dict[str, torch.Tensor]]:
    """Per-rank contiguous copies of the weight slices (1 / world of the model): q|k|v rows of the local heads packed as
    [q_local | k_local | v_local] (so that the fused RoPE + KV-append epilogue sees whole heads), the o_proj / down_proj
    COLUMN slices made contiguous, the gate|up rows of the local 64-wide groups (already contiguous)."""
    check_shardable(heads, kv_heads, inter, world)
    D = 128
    h0, h1 = shard_bounds(heads, rank, world)
    k0, k1 = shard_bounds(kv_heads, rank, world)
    g0, g1 = shard_bounds(inter // 64, rank, world)
    out = []
    for t in layers:
        rows = [(h0 * D, h1 * D), (heads * D + k0 * D, heads * D + k1 * D),
                ((heads + kv_heads) * D + k0 * D, (heads + kv_heads) * D + k1 * D)]
        w_qkv = torch.cat([t["w_qkv"][a:b] for a, b in rows], dim=0).contiguous()
        b_qkv = None if t.get("b_qkv") is None else torch.cat([t["b_qkv"][a:b] for a, b in rows], dim=0).contiguous()
        out.append(dict(ln_in=t["ln_in"], ln_post=t["ln_post"], w_qkv=w_qkv, b_qkv=b_qkv,
                        w_o=t["w_o"][:, h0 * D:h1 * D].contiguous(), w_gateup=t["w_gateup"][g0 * 128:g1 * 128],
                        w_down=t["w_down"][:, g0 * 64:g1 * 64].contiguous()))
    return out


@torch.no_grad()
def prefill_layers_native(engine, x: torch.Tensor, local: Sequence[Dict[str, torch.Tensor]], scratch, pos: torch.Tensor,
                          rope, past: int, rank: int, world: int, group=None) -> torch.Tensor:
    """The sharded prompt pass on the engine's own kernels (C ABI): per layer
      RMSNorm (vs_rmsnorm, replicated) -> local q|k|v GEMM with RoPE + in-place KV append of the LOCAL kv heads
      (vs_linear, VS_EPI_HEADS) -> causal attention over the local heads (vs_attention) -> row-parallel o_proj partial
      product (vs_linear) -> bf16 all-reduce (NCCL over NVLink / NVSwitch) -> residual add -> RMSNorm -> local gate|up with
      SiLU*up (VS_EPI_SWIGLU) -> row-parallel down_proj partial product -> bf16 all-reduce -> residual add,
    then one all-gather of the local kv-head rows so that every rank leaves with the complete KV slab.
    The partial products round to bf16 before the sum (half the all-reduce bytes of the fp32 variant; the single-device
    pass rounds the complete sum once), so the result differs from it by a few bf16 ulps -- the decode that follows is
    checked to produce the same tokens (tests/tp_prefill_gpu_check.py).
    `scratch`: engine ScratchBuffers for >= L rows; `pos` int32 [L]; `rope` L.Rope.  Returns the residual stream [L, H]."""
    import ctypes as C

    from . import _lib as L
    from .engine import _stream, run_linear
    cfg, lib = engine.cfg, engine.lib
    heads, kv_heads, inter, H, D = cfg.heads, cfg.kv_heads, cfg.inter, cfg.hidden, 128
    Lq = x.shape[0]
    S = past + Lq
    h0, h1 = shard_bounds(heads, rank, world)
    k0, k1 = shard_bounds(kv_heads, rank, world)
    g0, g1 = shard_bounds(inter // 64, rank, world)
    nh_l, nkv_l, I_l = h1 - h0, k1 - k0, (g1 - g0) * 64
    st = _stream()
    sc = scratch
    h = sc.h
    h[:Lq] = x
    part = torch.empty(Lq, H, dtype=x.dtype, device=x.device)
    eager = 0 if cfg.attn_impl == "sdpa" else 1
    kvs = engine.kv_max * 128
    # VS_TP_TIMING=1: CUDA-event time per kind of step, summed over the layers -> engine._tp_timing (ms); diagnostics only
    timing = os.environ.get("VS_TP_TIMING") == "1"
    marks: List = []

    def mark(name):
        if timing:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))
    mark("start")
    for li, t in enumerate(local):
        k_rows, v_rows = engine.kv[2 * li, 0, k0:k1], engine.kv[2 * li + 1, 0, k0:k1]      # local kv heads of the slab
        L.check(lib.vs_rmsnorm(L.ptr(h), H, L.ptr(t["ln_in"]), L.ptr(sc.x), H, Lq, H, cfg.eps, st), "vs_rmsnorm")
        mark("rmsnorm")
        a = L.LinearArgs()
        a.X, a.ldx, a.W, a.ldw, a.bias = L.ptr(sc.x), H, L.ptr(t["w_qkv"]), H, L.ptr(t["b_qkv"])
        a.T, a.N, a.K = Lq, (nh_l + 2 * nkv_l) * D, H
        a.splits = lib.vs_linear_splits(a.N, a.K, Lq)
        while a.splits > 1 and lib.vs_linear_partials_floats(a.N, Lq, a.splits) > sc.partials.numel():
            a.splits -= 1
        a.partials, a.counters, a.epi = L.ptr(sc.partials), L.ptr(sc.counters), L.EPI_HEADS
        a.n_q_tiles, a.n_k_tiles, a.rope = nh_l, nkv_l, 1
        a.cos, a.sin, a.pos_ids = rope.cos, rope.sin, L.ptr(pos)
        a.Qout, a.q_head_stride = L.ptr(sc.q), sc.max_rows * D
        a.Kout, a.Vout, a.kv_head_stride = L.ptr(k_rows), L.ptr(v_rows), kvs
        a.kv_row_base, a.kv_row_off = None, past
        L.check(lib.vs_linear(C.byref(a), st), "tp q|k|v")
        mark("qkv_gemm")
        at = L.AttnArgs()
        at.Q, at.q_head_stride, at.K, at.V, at.kv_head_stride = L.ptr(sc.q), sc.max_rows * D, L.ptr(k_rows), L.ptr(v_rows), kvs
        at.n_heads, at.n_kv_heads, at.Tq = nh_l, nkv_l, Lq
        at.n_ctx, at.n_ctx_host, at.n_new, at.causal, at.bits, at.eager = None, past, Lq, 1, None, eager
        at.scale = 1.0 / math.sqrt(D)
        at.scores, at.score_ld, at.stats, at.max_chunks = L.ptr(sc.scores), sc.score_ld, L.ptr(sc.stats), sc.max_chunks
        at.opart, at.pv_splits, at.counters = L.ptr(sc.opart), sc.pv_splits, L.ptr(sc.attn_counters)
        at.O, at.ldo = L.ptr(sc.attn), nh_l * D
        L.check(lib.vs_attention(C.byref(at), st), "tp attention")
        mark("attention")
        run_linear(lib, sc, sc.attn, nh_l * D, t["w_o"], None, Lq, H, nh_l * D, part, H)
        mark("o_gemm")
        if world > 1:
            dist.all_reduce(part, group=group)
        mark("all_reduce")
        h[:Lq] += part
        mark("residual_add")
        L.check(lib.vs_rmsnorm(L.ptr(h), H, L.ptr(t["ln_post"]), L.ptr(sc.x), H, Lq, H, cfg.eps, st), "vs_rmsnorm")
        mark("rmsnorm")
        run_linear(lib, sc, sc.x, H, t["w_gateup"], None, Lq, 2 * I_l, H, sc.act, I_l, epi=L.EPI_SWIGLU)
        mark("gateup_gemm")
        run_linear(lib, sc, sc.act, I_l, t["w_down"], None, Lq, H, I_l, part, H)
        mark("down_gemm")
        if world > 1:
            dist.all_reduce(part, group=group)
        mark("all_reduce")
        h[:Lq] += part
        mark("residual_add")
    if world > 1:
        n_layers = len(local)
        mine = engine.kv[:2 * n_layers, 0, k0:k1, past:S].contiguous()
        parts: List[torch.Tensor] = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        for r, p_ in enumerate(parts):
            a0, b0 = shard_bounds(kv_heads, r, world)
            if r != rank:
                engine.kv[:2 * n_layers, 0, a0:b0, past:S] = p_
    mark("kv_all_gather")
    if timing:
        torch.cuda.synchronize()
        acc: Dict[str, float] = {}
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e1)
        engine._tp_timing = {k: round(v, 3) for k, v in acc.items()}
    return h[:Lq]


def collective_bytes_per_rank(n_layers: int, rows: int, hidden: int, kv_heads: int, world: int) -> Dict[str, int]:
    """Payload of the collectives of one TP prefill (for reports): 2 fp32 all-reduces per layer + the KV all-gather."""
    return {"all_reduce": 2 * n_layers * rows * hidden * 4,
            "all_gather": 2 * n_layers * (kv_heads // max(1, world)) * rows * 128 * 2 * max(0, world - 1)}
