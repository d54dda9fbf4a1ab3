"""Host-side engine objects that own the device tensors and drive libvispec_b200.so.

PyTorch is plumbing here: it allocates HBM, provides the stream and (for the target prefill, which is
outside the draft-and-verify path) cuBLAS GEMMs.  Every op of the decode
cycle goes through the C ABI (vispec_b200/_lib.py -> include/vispec_b200.h).

Weight layout in HBM (packed once at load):
  * q|k|v rows fused into one [(nh+2nkv)*128, H] matrix  -> one weight stream, RoPE + KV append in the epilogue
  * gate/up interleaved in 64-row groups [2I, H]         -> SiLU*up in the epilogue
  * KV slab in the reference layout [2*layers, 1, n_kv, max_pos, 128] (vispec/model/kv_cache.py:118-126)
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import _lib as L
from .program import Program


def use_tiled() -> bool:
    """Keep a second, tiled copy of every decode weight for the program kernel (VS_TILED=0: stream the row-major copy
    through tensor maps instead)."""
    return os.environ.get("VS_TILED", "1") != "0"


def weight_layout() -> str:
    """HBM layout of the decode weights streamed by the program kernel: "tiled" (default: 16 KB tiles, stream-K over
    (tile, k-chunk) ranges + partial-tile epilogues), "slab" (round-2 experiment, csrc/slab.h: whole output rows per CTA,
    epilogue from TMEM, no cross-CTA reduction -- correct, but every CTA then walks the full K range and the path is bound
    by tcgen05.mma issue (K/16 instructions per CTA and GEMM): measured 130 vs 196 tokens/s) or
    "rowmajor" (tensor maps over the row-major copy).  VS_LAYOUT / VS_TILED=0 select the alternatives for A/B runs."""
    if not use_tiled():
        return "rowmajor"
    v = os.environ.get("VS_LAYOUT", "tiled")
    if v not in ("slab", "tiled", "rowmajor"):
        raise ValueError(f"VS_LAYOUT={v!r}: expected slab, tiled or rowmajor")
    return v


def use_programs() -> bool:
    """The persistent program kernel (one launch per step) is the product path; VS_MULTI_KERNEL=1 selects the
    one-launch-per-op path (same kernels' arithmetic, kept for A/B measurements and as a cross-check)."""
    return os.environ.get("VS_MULTI_KERNEL", "0") != "1"


# ---------------------------------------------------------------------------------------------------------
# configs (field names follow vispec/train/*_config.json and the LlamaConfig the reference reads)
# ---------------------------------------------------------------------------------------------------------
@dataclass
class TargetConfig:
    hidden: int = 4096
    inter: int = 11008
    layers: int = 32
    heads: int = 32
    kv_heads: int = 32
    vocab: int = 32064
    max_pos: int = 8192           # forced by vispec/model/modeling_llava_next_kv.py:10
    eps: float = 1e-5
    rope_theta: float = 10000.0
    # Qwen2.5-VL target (modeling_qwen2_5_vl_kv.py): multimodal RoPE sections over head_dim / 2 (:740-782; prefill only,
    # decode tokens carry one position on all three axes) and SDPA attention semantics (:1065-1170).  q/k/v biases are
    # picked up from the state dict.
    mrope_section: Optional[tuple] = None
    attn_impl: str = "eager"

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclass
class DraftConfig:
    hidden: int = 4096
    inter: int = 11008
    heads: int = 32
    vocab: int = 32064
    max_pos: int = 4096
    eps: float = 1e-5
    rope_theta: float = 10000.0
    total_tokens: int = 30        # cnets_ours.Model stores total_tokens - 1 (cnets_ours.py:733)
    depth: int = 3
    top_k: int = 8
    num_q: int = 2
    bias: bool = True

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def rope_tables(dim: int, max_pos: int, base: float, dtype, device):
    """cos/sin [max_pos, dim]: fp32 math, one cast to the model dtype (cnets_ours.py:129-155; the module is
    `.to(dtype)`-ed at spec_model_ours.py:98).  Computed on the host so the values do not depend on the GPU's
    transcendental units."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2).float() / dim))
    t = torch.arange(max_pos, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).to(device).contiguous(), emb.sin().to(dtype).to(device).contiguous()


def pack_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I,H],[I,H] -> [2I,H] with 64 gate rows followed by the matching 64 up rows (VS_EPI_SWIGLU layout)."""
    inter, hidden = gate.shape
    assert inter % 64 == 0, "intermediate size must be a multiple of 64"
    return torch.stack((gate.view(inter // 64, 64, hidden), up.view(inter // 64, 64, hidden)), dim=1) \
        .reshape(2 * inter, hidden).contiguous()


def pack_weight_tiles(w: torch.Tensor) -> torch.Tensor:
    """Row-major [N,K] bf16 -> the tiled HBM layout of include/vispec_b200.h (vs_program_register_tiled):
    [ceil(N/128), K/64, 128, 64] blocks of 16 KB, each block already in the 128-byte-swizzled shared-memory image
    (16-byte chunk j of row r stored at chunk j ^ (r & 7)); rows beyond N are zero.  1024-byte aligned."""
    n, k = w.shape
    assert k % 64 == 0, "K must be a multiple of 64"
    n_tiles = (n + 127) // 128
    if n_tiles * 128 != n:
        w = torch.cat([w, torch.zeros(n_tiles * 128 - n, k, dtype=w.dtype, device=w.device)], dim=0)
    t = w.view(n_tiles, 128, k // 64, 8, 8).permute(0, 2, 1, 3, 4)                 # [tile, kc, r, j, e]
    r = torch.arange(128, device=w.device)
    j = torch.arange(8, device=w.device)
    idx = (j[None, :] ^ (r[:, None] & 7))                                          # out[r, j'] = in[r, j' ^ (r & 7)]
    idx = idx[None, None, :, :, None].expand(n_tiles, k // 64, 128, 8, 8)
    out = torch.gather(t, 3, idx).contiguous()
    buf = torch.empty(out.numel() + 512, dtype=w.dtype, device=w.device)           # room to align to 1024 bytes
    off = ((-buf.data_ptr()) % 1024) // 2
    packed = buf[off:off + out.numel()]
    packed.copy_(out.view(-1))
    packed._vs_storage = buf
    return packed


def pack_weight_slab(w: torch.Tensor, epi: int) -> torch.Tensor:
    """Row-major [N,K] bf16 -> the slab image of include/vispec_b200.h (vs_slab_pack) for this device's SM count; `epi`
    picks the row pairing (L.EPI_SWIGLU / L.EPI_HEADS keep the two members of a pair in one accumulator lane).
    Returns None for shapes the layout does not cover (the linear then keeps the stream-K path)."""
    lib = L.load()
    n, k = w.shape
    with torch.cuda.device(w.device):
        nbytes = lib.vs_slab_bytes(n, k, epi, 0)
        if nbytes == 0:
            return None
        buf = torch.empty(nbytes // 2 + 64, dtype=w.dtype, device=w.device)
        off = ((-buf.data_ptr()) % 128) // 2
        slab = buf[off:off + nbytes // 2]
        L.pcheck(lib.vs_slab_pack(w.data_ptr(), w.stride(0), n, k, epi, 0, slab.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), "vs_slab_pack")
    slab._vs_storage = buf
    slab._vs_epi = epi
    return slab


def pack_decode_weight(w: torch.Tensor, epi: int):
    """(row-major weight, streamed copy, epi) in the configured layout, or None when the row-major copy is streamed."""
    lay = weight_layout()
    if lay == "slab":
        s = pack_weight_slab(w, epi)
        if s is not None:
            return (w, s, epi)
        lay = "tiled"
    if lay == "tiled":
        return (w, pack_weight_tiles(w), -1)
    return None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(device):
    if torch.device(device).type != "cuda":
        raise RuntimeError("vispec_b200 runs on CUDA devices only (no CPU fallback)")


# ---------------------------------------------------------------------------------------------------------
# scratch
# ---------------------------------------------------------------------------------------------------------
class ScratchBuffers:
    """vs_scratch: caller-owned work buffers for forwards of up to max_rows rows over up to max_kv keys."""

    def __init__(self, hidden: int, inter: int, heads: int, vocab: int, max_rows: int, max_kv: int, device,
                 extra_n: int = 0):
        lib = L.load()
        self.max_rows = _ceil(max_rows, 32)
        tok_blocks = self.max_rows // 32
        bf, f32, i32 = torch.bfloat16, torch.float32, torch.int32
        z = lambda *s, dtype=bf: torch.zeros(*s, dtype=dtype, device=device)
        self.h = z(self.max_rows, hidden)
        self.x = z(self.max_rows, hidden)
        self.q = z(heads, self.max_rows, 128)
        self.attn = z(self.max_rows, heads * 128)
        self.act = z(self.max_rows, inter)
        shapes = [(3 * heads * 128, hidden), (hidden, heads * 128), (2 * inter, hidden), (hidden, inter),
                  (vocab, hidden), (hidden, 2 * hidden)]
        if extra_n:
            shapes.append((extra_n, hidden))
        need = 1
        for n, k in shapes:
            sp = lib.vs_linear_splits(n, k, self.max_rows)
            need = max(need, lib.vs_linear_partials_floats(n, self.max_rows, sp))
        self.partials = z(need, dtype=f32)
        max_n = max(n for n, _ in shapes)
        self.counters = z(tok_blocks * ((max_n + 127) // 128) + 64, dtype=i32)
        self.score_ld = _ceil(max_kv, 64)
        self.max_chunks = self.score_ld // 64
        self.pv_splits = max(1, min(16, self.max_chunks))
        self.scores = z(heads * self.max_rows * self.score_ld, dtype=f32)
        self.stats = z(heads * tok_blocks * self.max_chunks * 32 * 2, dtype=f32)
        self.opart = z(heads * tok_blocks * self.pv_splits * 32 * 128, dtype=f32)
        self.attn_counters = z(heads * tok_blocks + 64, dtype=i32)
        # keys the single-phase ("fused") attention of a program can hold with these buffers
        self.fused_kv_capacity = lib.vs_attention_fused_capacity(heads, self.pv_splits, self.max_chunks)
        self.c = L.Scratch(self.max_rows, L.ptr(self.h), L.ptr(self.x), L.ptr(self.q), L.ptr(self.attn),
                           L.ptr(self.act), L.ptr(self.partials), self.partials.numel(), L.ptr(self.counters),
                           L.ptr(self.scores), self.score_ld, L.ptr(self.stats), self.max_chunks,
                           L.ptr(self.opart), self.pv_splits, L.ptr(self.attn_counters))


def run_linear(lib, sc: ScratchBuffers, x, ldx, w, bias, T, N, K, y, ldy, epi=L.EPI_STORE):
    """y[T,N] = x[T,K] @ w[N,K]^T (+ bias) through vs_linear with the caller's scratch."""
    a = L.LinearArgs()
    a.X, a.ldx, a.W, a.ldw, a.bias = L.ptr(x), ldx, L.ptr(w), K, L.ptr(bias)
    a.T, a.N, a.K = T, N, K
    a.splits = lib.vs_linear_splits(N, K, T)
    while a.splits > 1 and lib.vs_linear_partials_floats(N, T, a.splits) > sc.partials.numel():
        a.splits -= 1
    a.partials, a.counters, a.epi = L.ptr(sc.partials), L.ptr(sc.counters), epi
    a.Y, a.ldy = L.ptr(y), ldy
    L.check(lib.vs_linear(C.byref(a), _stream()), "vs_linear")


def _dims(hidden, inter, heads, kv_heads, vocab, layers, eps) -> L.Dims:
    return L.Dims(hidden, inter, heads, kv_heads, vocab, layers, eps)


# ---------------------------------------------------------------------------------------------------------
# target LM
# ---------------------------------------------------------------------------------------------------------
class TargetEngine:
    """The target LM of vispec/model/modeling_llama_kv.py with the static KV slab of kv_cache.py.

    `weights` uses the reference state-dict keys (`model.layers.N.self_attn.q_proj.weight`, ...)."""

    def __init__(self, cfg: TargetConfig, weights: Dict[str, torch.Tensor], device="cuda", kv_max: Optional[int] = None):
        _require_cuda(device)
        if cfg.head_dim != 128:
            raise NotImplementedError("head_dim must be 128")
        self.lib = L.load()
        self.cfg, self.device = cfg, torch.device(device)
        self.dtype = torch.bfloat16
        dev = self.device
        g = lambda k: weights[k].to(device=dev, dtype=self.dtype)
        self.embed = g("model.embed_tokens.weight").contiguous()
        self.norm = g("model.norm.weight").contiguous()
        self.lm_head = g("lm_head.weight").contiguous()
        self.kv_max = kv_max or cfg.max_pos
        # reference slab layout (kv_cache.py:118-126), one device
        self.kv = torch.zeros(2 * cfg.layers, 1, cfg.kv_heads, self.kv_max, 128, dtype=self.dtype, device=dev)
        self.cos, self.sin = rope_tables(128, cfg.max_pos, cfg.rope_theta, self.dtype, dev)
        self.layers_t: List[dict] = []
        arr = (L.Layer * cfg.layers)()
        for li in range(cfg.layers):
            p = f"model.layers.{li}."
            w_qkv = torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"),
                               g(p + "self_attn.v_proj.weight")], dim=0).contiguous()
            b_qkv = None
            if (p + "self_attn.q_proj.bias") in weights:
                b_qkv = torch.cat([g(p + "self_attn.q_proj.bias"), g(p + "self_attn.k_proj.bias"),
                                   g(p + "self_attn.v_proj.bias")], dim=0).contiguous()
            t = dict(ln_in=g(p + "input_layernorm.weight").contiguous(), w_qkv=w_qkv, b_qkv=b_qkv,
                     w_o=g(p + "self_attn.o_proj.weight").contiguous(),
                     ln_post=g(p + "post_attention_layernorm.weight").contiguous(),
                     w_gateup=pack_gate_up(g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight")),
                     w_down=g(p + "mlp.down_proj.weight").contiguous())
            self.layers_t.append(t)
            arr[li] = L.Layer(L.ptr(t["ln_in"]), L.ptr(w_qkv), L.ptr(b_qkv), L.ptr(t["w_o"]), L.ptr(t["ln_post"]),
                              L.ptr(t["w_gateup"]), L.ptr(t["w_down"]), L.ptr(self.kv[2 * li]), L.ptr(self.kv[2 * li + 1]),
                              self.kv_max * 128)
        self._layer_arr = arr
        # tiled copies for the persistent program kernel (the row-major copies stay for the prefill / per-op path)
        self.tiled = []
        if use_programs():
            epis = (("w_qkv", L.EPI_HEADS), ("w_o", L.EPI_RESID), ("w_gateup", L.EPI_SWIGLU), ("w_down", L.EPI_RESID))
            for t in self.layers_t:
                for k, epi in epis:
                    self.tiled.append(pack_decode_weight(t[k], epi))
            self.tiled.append(pack_decode_weight(self.lm_head, L.EPI_STORE))
            self.tiled = [p for p in self.tiled if p is not None]
        self.dims = _dims(cfg.hidden, cfg.inter, cfg.heads, cfg.kv_heads, cfg.vocab, cfg.layers, cfg.eps)
        if cfg.attn_impl not in ("eager", "sdpa"):
            raise NotImplementedError(f"attention implementation {cfg.attn_impl!r}")
        self.c = L.Target(self.dims, arr, L.ptr(self.embed), L.ptr(self.norm), L.ptr(self.lm_head),
                          L.Rope(L.ptr(self.cos), L.ptr(self.sin)), 1 if cfg.attn_impl == "sdpa" else 0)
        self.rope_delta = 0
        self.scratch = ScratchBuffers(cfg.hidden, cfg.inter, cfg.heads, cfg.vocab, 32, self.kv_max, dev,
                                      extra_n=(cfg.heads + 2 * cfg.kv_heads) * 128)
        self.state = torch.zeros(L.ST_SLOTS, dtype=torch.int32, device=dev)
        self.hidden_out = torch.zeros(32, cfg.hidden, dtype=self.dtype, device=dev)
        self.logits = torch.zeros(32, cfg.vocab, dtype=self.dtype, device=dev)
        self.amax = torch.zeros(32, dtype=torch.int32, device=dev)
        self._causal_pos = torch.zeros(32, dtype=torch.int32, device=dev)
        # fixed staging buffers: programs bake their operand pointers in
        self._v_ids = torch.zeros(32, dtype=torch.int32, device=dev)
        self._v_pos = torch.zeros(32, dtype=torch.int32, device=dev)
        self._v_bits = torch.zeros(32, dtype=torch.int64, device=dev)
        self._verify_progs: Dict[tuple, Program] = {}
        self._tp = None                      # (group, rank, world) once enable_tensor_parallel_prefill() was called

    def enable_tensor_parallel_prefill(self, group=None):
        """Shard the prompt pass over the ranks of `group` (default: the world group; one process per GPU, NCCL):
        column-parallel q/k/v/gate/up, row-parallel o/down with one all-reduce each, KV all-gathered at the end
        (vispec_b200/tp_prefill.py).  Every rank must then call prefill / specgenerate with the same prompt; the
        draft-and-verify decode stays replicated."""
        import torch.distributed as dist
        from . import tp_prefill
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("enable_tensor_parallel_prefill needs an initialised torch.distributed process group")
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        tp_prefill.check_shardable(self.cfg.heads, self.cfg.kv_heads, self.cfg.inter, world)
        self._tp = (group, rank, world)
        # per-rank weight slices for the pass on the library's own kernels (VS_TP_PREFILL=library keeps the torch pass)
        self._tp_local = None
        if world > 1 and os.environ.get("VS_TP_PREFILL", "native") != "library":
            self._tp_local = tp_prefill.shard_layer_weights(self.layers_t, heads=self.cfg.heads, kv_heads=self.cfg.kv_heads,
                                                            inter=self.cfg.inter, rank=rank, world=world)
        return self

    def _verify_program(self, T: int, has_bits: bool, fused: bool) -> Program:
        key = (T, has_bits, fused)
        p = self._verify_progs.get(key)
        if p is None:
            p = Program(self.device, attn_fused=fused).register_tiled(self.tiled)
            p.add("verify_step", C.byref(self.c), C.byref(self.scratch.c), L.ptr(self._v_ids), L.ptr(self._v_pos),
                  L.ptr(self._v_bits) if has_bits else None, T, L.ptr(self.state), L.ptr(self.hidden_out),
                  L.ptr(self.logits), L.ptr(self.amax))
            p.commit()
            self._verify_progs[key] = p
        return p

    # -- lengths -------------------------------------------------------------------------------------
    def set_len(self, n: int):
        self.state[L.ST_CUR_LEN] = n

    def get_len(self) -> int:
        return int(self.state[L.ST_CUR_LEN].item())

    def set_rope_delta(self, d: int):
        """Qwen2.5-VL `rope_deltas` (modeling_qwen2_5_vl_kv.py:2151-2170): decode position = cache position + delta."""
        self.rope_delta = int(d)
        self.state[L.ST_ROPE_DELTA] = int(d)

    def embed_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.embedding(ids.to(self.device), self.embed)

    # -- prefill -------------------------------------------------------------------------------------------
    def prefill(self, inputs_embeds: torch.Tensor, past: int = 0, position_ids: Optional[torch.Tensor] = None):
        """LlamaModel.forward over the prompt (modeling_llama_kv.py:927-1080): writes K/V rows [past, past+L) of the slab,
        returns (last-row logits fp32 [V], hidden [1,L,H] post-norm).  Runs on the library's own kernels
        (`prefill_native`) unless the tensor-parallel prompt pass is enabled or VS_PREFILL=library."""
        tp_native = self._tp is not None and getattr(self, "_tp_local", None) is not None
        if (self._tp is None or tp_native) and os.environ.get("VS_PREFILL", "native") != "library":
            return self.prefill_native(inputs_embeds, past, position_ids)
        return self.prefill_library(inputs_embeds, past, position_ids)

    @torch.no_grad()
    def prefill_library(self, inputs_embeds: torch.Tensor, past: int = 0, position_ids: Optional[torch.Tensor] = None):
        """The prompt pass with plain torch ops / cuBLAS (and, when enabled, sharded tensor-parallel over NCCL).
        Writes K/V rows [past, past+L) of the slab; returns (last-row logits fp32 [V], hidden [1,L,H] post-norm).
        `position_ids`: [L], or [3, L] multimodal positions for a Qwen2.5-VL target (`mrope_section`)."""
        cfg = self.cfg
        x = inputs_embeds.to(device=self.device, dtype=self.dtype)
        if x.dim() == 2:
            x = x[None]
        Lq = x.shape[1]
        if past + Lq > self.kv_max:
            raise ValueError("prompt longer than the KV slab")
        nh, nkv, D = cfg.heads, cfg.kv_heads, 128
        if position_ids is None:
            pos = torch.arange(past, past + Lq, device=self.device)
        else:
            pos = position_ids.to(self.device).long().reshape(-1, Lq)
            pos = pos[0] if cfg.mrope_section is None else pos
        if cfg.mrope_section is not None:
            # apply_multimodal_rotary_pos_emb (modeling_qwen2_5_vl_kv.py:768-775): channel chunk i from axis i % 3
            pos3 = pos if pos.dim() == 2 else pos[None].expand(3, Lq)
            sec = list(cfg.mrope_section) * 2
            cos = torch.cat([m[i % 3] for i, m in enumerate(self.cos[pos3].split(sec, dim=-1))], dim=-1)[None, None]
            sin = torch.cat([m[i % 3] for i, m in enumerate(self.sin[pos3].split(sec, dim=-1))], dim=-1)[None, None]
        else:
            cos, sin = self.cos[pos][None, None], self.sin[pos][None, None]
        sdpa = cfg.attn_impl == "sdpa"
        S = past + Lq
        mask = torch.full((Lq, S), torch.finfo(torch.float32).min, device=self.device)
        mask = torch.triu(mask, diagonal=past + 1)[None, None]
        h = x
        F = torch.nn.functional
        if self._tp is not None and self._tp[2] > 1:
            from . import tp_prefill
            group, rank, world = self._tp
            h = tp_prefill.prefill_layers(x[0], self.layers_t, self.kv, heads=nh, kv_heads=nkv, inter=cfg.inter, eps=cfg.eps,
                                          cos=cos, sin=sin, past=past, sdpa=sdpa, rank=rank, world=world, group=group)[None]
            hidden = tp_prefill._rms(h, self.norm, cfg.eps)
            logits_last = F.linear(hidden[0, -1], self.lm_head).float()
            self.set_len(S)
            return logits_last, hidden

        def rms(t, w):
            tf = t.float()
            tf = tf * torch.rsqrt(tf.pow(2).mean(-1, keepdim=True) + cfg.eps)
            return w * tf.to(self.dtype)

        def rot(t):
            return torch.cat((-t[..., 64:], t[..., :64]), dim=-1)

        for li, t in enumerate(self.layers_t):
            res = h
            xn = rms(h, t["ln_in"])
            qkv = F.linear(xn, t["w_qkv"], t["b_qkv"])
            q, k, v = qkv.split([nh * D, nkv * D, nkv * D], dim=-1)
            q = q.view(1, Lq, nh, D).transpose(1, 2)
            k = k.view(1, Lq, nkv, D).transpose(1, 2)
            v = v.view(1, Lq, nkv, D).transpose(1, 2)
            q = (q * cos) + (rot(q) * sin)
            k = (k * cos) + (rot(k) * sin)
            self.kv[2 * li, :, :, past:S] = k
            self.kv[2 * li + 1, :, :, past:S] = v
            K, V = self.kv[2 * li, :, :, :S], self.kv[2 * li + 1, :, :, :S]
            if nkv != nh:
                rep = nh // nkv
                K = K[:, :, None].expand(1, nkv, rep, S, D).reshape(1, nh, S, D)
                V = V[:, :, None].expand(1, nkv, rep, S, D).reshape(1, nh, S, D)
            if sdpa:       # modeling_qwen2_5_vl_kv.py:1153-1160 (the mask is inferred when there is no past, :1489-1503)
                ao = F.scaled_dot_product_attention(q, K, V, is_causal=True) if past == 0 else \
                    F.scaled_dot_product_attention(q, K, V, attn_mask=mask.to(self.dtype))
                ao = ao.transpose(1, 2).reshape(1, Lq, nh * D)
            else:
                aw = torch.matmul(q, K.transpose(2, 3)) / math.sqrt(D)
                aw = F.softmax(aw + mask, dim=-1, dtype=torch.float32).to(self.dtype)
                ao = torch.matmul(aw, V).transpose(1, 2).reshape(1, Lq, nh * D)
            h = res + F.linear(ao, t["w_o"])
            res = h
            xn = rms(h, t["ln_post"])
            gu = F.linear(xn, t["w_gateup"]).view(1, Lq, cfg.inter // 64, 2, 64)
            act = (F.silu(gu[..., 0, :]) * gu[..., 1, :]).reshape(1, Lq, cfg.inter)
            h = res + F.linear(act, t["w_down"])
        hidden = rms(h, self.norm)
        logits_last = F.linear(hidden[0, -1], self.lm_head).float()
        self.set_len(S)
        return logits_last, hidden

    # -- prefill on the library's own kernels --------------------------------------------------------------------
    @torch.no_grad()
    def prefill_native(self, inputs_embeds: torch.Tensor, past: int = 0, position_ids: Optional[torch.Tensor] = None):
        """The prompt pass through the C ABI, the default (`vs_decoder_layer` per layer over all prompt rows: RMSNorm,
        tcgen05 q|k|v GEMM with 256-token MMA tiles and RoPE + in-place KV append in its epilogue, causal attention,
        o_proj + residual, SwiGLU MLP), then the final norm.  Same results as the library (cuBLAS / torch) prompt pass
        `prefill_library` up to fp32 summation order; VS_PREFILL=library selects that one."""
        cfg = self.cfg
        x = inputs_embeds.to(device=self.device, dtype=self.dtype)
        x = x[0] if x.dim() == 3 else x
        Lq = x.shape[0]
        S = past + Lq
        if S > self.kv_max:
            raise ValueError("prompt longer than the KV slab")
        big = getattr(self, "_prefill_scratch", None)
        if big is None or big.max_rows < Lq or big.score_ld < S:
            big = self._prefill_scratch = ScratchBuffers(cfg.hidden, cfg.inter, cfg.heads, cfg.vocab, _ceil(Lq, 32),
                                                         _ceil(S, 64), self.device,
                                                         extra_n=(cfg.heads + 2 * cfg.kv_heads) * 128)
        big.h[:Lq] = x
        if cfg.mrope_section is not None:
            # Qwen2.5-VL: rotary factors of row t = channel chunk i from position axis i % 3
            # (apply_multimodal_rotary_pos_emb, modeling_qwen2_5_vl_kv.py:768-775), gathered once into a per-prompt table
            # that the q|k|v epilogue indexes by row
            if position_ids is None:
                pos3 = torch.arange(past, S, device=self.device)[None].expand(3, Lq)
            else:
                pos3 = position_ids.to(self.device).long().reshape(-1, Lq)
                pos3 = pos3 if pos3.shape[0] == 3 else pos3[:1].expand(3, Lq)
            sec = list(cfg.mrope_section) * 2
            cos_t = torch.cat([m[i % 3] for i, m in enumerate(self.cos[pos3].split(sec, dim=-1))], dim=-1).contiguous()
            sin_t = torch.cat([m[i % 3] for i, m in enumerate(self.sin[pos3].split(sec, dim=-1))], dim=-1).contiguous()
            pos = torch.arange(0, Lq, dtype=torch.int32, device=self.device)
            rope = L.Rope(L.ptr(cos_t), L.ptr(sin_t))
        else:
            if position_ids is None:
                pos = torch.arange(past, S, dtype=torch.int32, device=self.device)
            else:
                pos = position_ids.to(self.device).reshape(-1, Lq)[0].to(torch.int32).contiguous()
            rope = L.Rope(L.ptr(self.cos), L.ptr(self.sin))
        st = _stream()
        eager = 0 if cfg.attn_impl == "sdpa" else 1
        if self._tp is not None and getattr(self, "_tp_local", None) is not None:
            # tensor-parallel prompt pass on the same kernels: local heads / gate|up groups, bf16 all-reduce on the two
            # row-parallel boundaries, KV all-gather (vispec_b200/tp_prefill.py)
            from . import tp_prefill
            group, rank, world = self._tp
            tp_prefill.prefill_layers_native(self, x, self._tp_local, big, pos, rope, past, rank, world, group)
        else:
            for li in range(cfg.layers):
                L.check(self.lib.vs_decoder_layer(C.byref(self.dims), C.byref(self._layer_arr[li]), C.byref(big.c),
                                                  C.byref(rope), Lq, L.ptr(pos), None, past, None, past, Lq, 1, None, eager,
                                                  st), "prefill layer")
        hidden = torch.empty(Lq, cfg.hidden, dtype=self.dtype, device=self.device)
        L.check(self.lib.vs_rmsnorm(L.ptr(big.h), cfg.hidden, L.ptr(self.norm), L.ptr(hidden), cfg.hidden, Lq, cfg.hidden,
                                    cfg.eps, st), "vs_rmsnorm")
        # LM head on the one row the caller consumes (modeling_llama_kv.py:1196-1197 computes all rows)
        run_linear(self.lib, self.scratch, hidden[Lq - 1:], cfg.hidden, self.lm_head, None, 1, cfg.vocab, cfg.hidden,
                   self.logits, cfg.vocab)
        logits_last = self.logits[0].float()
        self.set_len(S)
        return logits_last, hidden[None]

    # -- decode / verify through the C ABI -----------------------------------------------------------------
    def verify(self, token_ids: torch.Tensor, pos_ids: torch.Tensor, bits: Optional[torch.Tensor], T: int,
               cur_len: Optional[int] = None):
        """vs_verify_step on T <= 32 tokens appended (uncommitted) at CUR_LEN.  token_ids / pos_ids int32 device.
        (`cur_len` is accepted for compatibility; one verify program serves every context length.)"""
        if use_programs():
            self._v_ids[:T].copy_(token_ids[:T])
            self._v_pos[:T].copy_(pos_ids[:T])
            if bits is not None:
                self._v_bits[:T].copy_(bits[:T])
            self._verify_program(T, bits is not None, True).run()
            return
        L.check(self.lib.vs_verify_step(C.byref(self.c), C.byref(self.scratch.c), L.ptr(token_ids), L.ptr(pos_ids),
                                        L.ptr(bits), T, L.ptr(self.state), L.ptr(self.hidden_out), L.ptr(self.logits),
                                        L.ptr(self.amax), _stream()), "vs_verify_step")

    def decode_causal(self, ids: torch.Tensor):
        """Plain causal decode of 1..32 tokens (the autoregressive baseline loop of
        vispec/evaluation/gen_baseline_answer_coco_caption.py:34-133); commits the rows."""
        T = ids.numel()
        cur = self.get_len()
        self._causal_pos[:T] = torch.arange(cur, cur + T, dtype=torch.int32, device=self.device)
        self.verify(ids.to(device=self.device, dtype=torch.int32).contiguous(), self._causal_pos, None, T, cur_len=cur)
        self.set_len(cur + T)
        return self.logits[:T]


# ---------------------------------------------------------------------------------------------------------
# draft model (spec_layer)
# ---------------------------------------------------------------------------------------------------------
class DraftEngine:
    """The ViSpec draft model of vispec/model/cnets_ours.py (`Model`): one decoder layer without input norm,
    fc / img_fc, ImgAdaptor, dynamic draft tree.  `weights` uses the reference state-dict keys."""

    def __init__(self, cfg: DraftConfig, weights: Dict[str, torch.Tensor], lm_head: torch.Tensor, device="cuda",
                 cache_rows: int = 4096, max_positions: int = 8192 + 64):
        _require_cuda(device)
        if cfg.head_dim != 128:
            raise NotImplementedError("head_dim must be 128")
        if cfg.top_k > 8 or cfg.depth + 2 > L.VS_RETRIEVE_LD or cfg.total_tokens > 32:
            raise NotImplementedError("tree shape outside the kernels' limits (top_k<=8, depth<=6, total_token<=32)")
        self.lib = L.load()
        self.cfg, self.device, self.dtype = cfg, torch.device(device), torch.bfloat16
        dev, H, nh = self.device, cfg.hidden, cfg.heads
        g = lambda k: weights[k].to(device=dev, dtype=self.dtype).contiguous() if k in weights else None
        p = "layers.0."
        self.embed = g("embed_tokens.weight")
        self.w_qkv = torch.cat([g(p + "self_attn.q_proj.weight"), g(p + "self_attn.k_proj.weight"),
                                g(p + "self_attn.v_proj.weight")], dim=0).contiguous()
        self.b_qkv = None
        if (p + "self_attn.q_proj.bias") in weights:
            self.b_qkv = torch.cat([g(p + "self_attn.q_proj.bias"), g(p + "self_attn.k_proj.bias"),
                                    g(p + "self_attn.v_proj.bias")], dim=0).contiguous()
        self.w_o = g(p + "self_attn.o_proj.weight")
        self.ln_post = g(p + "post_attention_layernorm.weight")
        self.w_gateup = pack_gate_up(g(p + "mlp.gate_proj.weight"), g(p + "mlp.up_proj.weight"))
        self.w_down = g(p + "mlp.down_proj.weight")
        self.w_fc, self.b_fc = g("fc.weight"), g("fc.bias")
        self.w_img_fc, self.b_img_fc = g("img_fc.weight"), g("img_fc.bias")
        # vision adaptor (cnets_ours.py:603-661): k|v projections fused, learned queries head-major
        self.w_akv = torch.cat([g("imadpt.k_proj.weight"), g("imadpt.v_proj.weight")], dim=0).contiguous()
        self.b_akv = None
        if "imadpt.k_proj.bias" in weights:
            self.b_akv = torch.cat([g("imadpt.k_proj.bias"), g("imadpt.v_proj.bias")], dim=0).contiguous()
        self.w_ao = g("imadpt.o_proj.weight")
        self.adpt_q = torch.zeros(nh, 32, 128, dtype=self.dtype, device=dev)
        self.adpt_q[:, :cfg.num_q] = g("imadpt.q").view(cfg.num_q, nh, 128).transpose(0, 1)
        self.lm_head = lm_head
        self.tiled = []                                   # the target engine's lm_head pair is added by bind_tiled
        if use_programs():
            for w, epi in ((self.w_qkv, L.EPI_HEADS), (self.w_o, L.EPI_RESID), (self.w_gateup, L.EPI_SWIGLU),
                           (self.w_down, L.EPI_RESID), (self.w_fc, L.EPI_STORE), (self.w_img_fc, L.EPI_STORE)):
                self.tiled.append(pack_decode_weight(w, epi))
            self.tiled = [p for p in self.tiled if p is not None]
        # Draft positions follow the UNCOMPRESSED sequence index (image tokens keep their original positions,
        # cnets_ours.py:932-937), so they run up to the target's KV length, not the draft's max_position_embeddings: the
        # tables cover both (rows < cfg.max_pos equal the reference's; it raises an index error beyond them)
        self.rope_rows = max(cfg.max_pos, int(max_positions))
        self.cos, self.sin = rope_tables(128, self.rope_rows, cfg.rope_theta, self.dtype, dev)
        self.cache_rows = cache_rows
        self.k_cache = torch.zeros(nh, cache_rows, 128, dtype=self.dtype, device=dev)
        self.v_cache = torch.zeros(nh, cache_rows, 128, dtype=self.dtype, device=dev)
        self.dims = _dims(H, cfg.inter, nh, nh, cfg.vocab, 1, cfg.eps)
        self.scratch = ScratchBuffers(H, cfg.inter, nh, cfg.vocab, 32, cache_rows, dev)
        self.x1 = torch.zeros(32, 2 * H, dtype=self.dtype, device=dev)
        self.x2 = torch.zeros(32, 2 * H, dtype=self.dtype, device=dev)
        self.logits = torch.zeros(32, cfg.vocab, dtype=self.dtype, device=dev)
        self.c = self._make_struct(self.x1, self.x2, self.logits)
        self.last_img_hidden = torch.zeros(1, H, dtype=self.dtype, device=dev)
        self._alloc_tree()
        self.state: Optional[torch.Tensor] = None       # shared with the target engine (bind_state)
        self.adaptor_c = L.Adaptor(L.ptr(self.w_akv), L.ptr(self.b_akv), L.ptr(self.w_ao), L.ptr(self.adpt_q), cfg.num_q)
        self._ws_cache: Dict[int, tuple] = {}           # prefill work buffers by row count (vs_img_adaptor_prefill)
        self._prog_levels: Dict[bool, Program] = {}     # LM head + draft levels + tree build after the prefill
        self._prog_tree: Dict[bool, Program] = {}       # stable forward of the accepted rows + the above
        # (keyed by: single-phase attention usable for the current draft KV length)

    def _layer_struct(self) -> L.Layer:
        return L.Layer(None, L.ptr(self.w_qkv), L.ptr(self.b_qkv), L.ptr(self.w_o), L.ptr(self.ln_post),
                       L.ptr(self.w_gateup), L.ptr(self.w_down), L.ptr(self.k_cache), L.ptr(self.v_cache),
                       self.cache_rows * 128)

    def _make_struct(self, x1, x2, logits) -> L.Draft:
        return L.Draft(self.dims, self._layer_struct(), L.ptr(self.embed), L.ptr(self.w_fc), L.ptr(self.b_fc),
                       L.ptr(self.w_img_fc), L.ptr(self.b_img_fc), L.ptr(self.lm_head),
                       L.Rope(L.ptr(self.cos), L.ptr(self.sin)), L.ptr(x1), L.ptr(x2), L.ptr(logits))

    def _alloc_tree(self):
        cfg, dev = self.cfg, self.device
        k, depth, total = cfg.top_k, cfg.depth, cfg.total_tokens - 1
        n_sc = k + depth * k * k
        i32, f32, u64 = torch.int32, torch.float32, torch.int64
        z = lambda n, dt: torch.zeros(n, dtype=dt, device=dev)
        t = dict(scores_flat=z(n_sc, f32), tokens_flat=z(n_sc, i32), parents=z(1 + depth * k, i32),
                 frontier_scores=z(k, f32), frontier_ids=z(32, i32), cs_index=z(k, i32), frontier_bits=z(32, u64),
                 top_val=z(k * k, f32), top_idx=z(k * k, i32), gather_idx=z(32, i32), draft_pos_ids=z(32, i32),
                 draft_tokens=z(32, i32), tree_bits=z(32, u64), tree_pos=z(32, i32), verify_pos=z(32, i32),
                 retrieve=z(32 * L.VS_RETRIEVE_LD, i32), sel=z(L.VS_MAX_ACCEPT, i32),
                 stable_ids=z(32, i32), stable_pos=z(32, i32))
        self.tree = t
        self.tb = L.TreeBufs(k, depth, total, *[L.ptr(t[name]) for name, _ in L.TreeBufs._fields_[3:]])
        self.out_rec = z(2 + L.VS_MAX_ACCEPT, i32)

    def bind_state(self, state: torch.Tensor):
        self.state = state

    def bind_tiled(self, pairs):
        """Tiled copies owned by the target engine that the draft also streams (the shared LM head)."""
        self.tiled = self.tiled + [p for p in pairs if p[0].data_ptr() == self.lm_head.data_ptr()]

    def reset_kv(self):
        """cnets_ours.py:1040-1041."""
        self.state[L.ST_D_LEN] = 0
        self.state[L.ST_D_POS] = 0

    # -- first call: image/prefill branch (cnets_ours.py:879-975) -----------------------------------------
    def _prefill_ws(self, rows: int):
        """Work buffers of vs_img_adaptor_prefill for prompts of up to `rows` rows, cached by size (no per-prompt
        allocation once a size has been seen)."""
        rows = _ceil(max(rows, 32), 32)
        ws = self._ws_cache.get(rows)
        if ws is None:
            cfg, dev, H, nh = self.cfg, self.device, self.cfg.hidden, self.cfg.heads
            z = lambda *s, dtype=self.dtype: torch.zeros(*s, dtype=dtype, device=dev)
            t = dict(img_k=z(nh, rows, 128), img_v=z(nh, rows, 128), x1=z(rows, 2 * H), x2=z(rows, 2 * H), fused=z(rows, H),
                     adapted=z(32, H), attn_out=z(32, H), pos=z(rows, dtype=torch.int32), packed=z(rows, H))
            sc = ScratchBuffers(H, cfg.inter, nh, cfg.vocab, rows, rows, dev, extra_n=2 * H)
            c = L.PrefillWs(rows, *[L.ptr(t[k]) for k in ("img_k", "img_v", "x1", "x2", "fused", "adapted", "attn_out", "pos",
                                                           "packed")])
            ws = self._ws_cache[rows] = (t, sc, c)
            while len(self._ws_cache) > 4:                       # keep the few most recent sizes
                self._ws_cache.pop(next(iter(self._ws_cache)))
        return ws

    @torch.no_grad()
    def prefill(self, hidden: torch.Tensor, inputs_embeds: torch.Tensor, image_mask: Optional[torch.Tensor],
                sample_token: int):
        """hidden [1,L,H] target features; inputs_embeds [1,L,H] ALREADY shifted by one (cnets_ours.py:1066-1082);
        image_mask [1,L] bool in the unshifted frame or None.  One C-ABI call (vs_img_adaptor_prefill) builds the packed
        sequence (text rows through fc/img_fc, the image run compressed by the adaptor to num_q-1 tokens at the run's last
        positions) and runs the decoder layer over it; a second one expands the draft tree."""
        if os.environ.get("VS_DRAFT_PREFILL", "native") == "python":
            return self.prefill_python(hidden, inputs_embeds, image_mask, sample_token)
        cfg, lib, dev, H = self.cfg, self.lib, self.device, self.cfg.hidden
        hid = hidden[0].to(dev, self.dtype).contiguous()
        emb = inputs_embeds[0].to(dev, self.dtype).contiguous()
        Lq = hid.shape[0]
        img_start, img_end = -1, -1
        if image_mask is not None:
            im = image_mask[0, 1:].to(dev)                                  # :880 (shifted frame, L-1 entries)
            idx = torch.nonzero(im).flatten()
            if idx.numel() == 0:
                raise NotImplementedError("image_mask without an image token (the reference fails too, cnets_ours.py:938-941)")
            first, last = [int(v) for v in torch.stack((idx[0], idx[-1])).tolist()]     # the one host sync: two ints
            if last - first + 1 != idx.numel():
                raise NotImplementedError("exactly one contiguous image run is supported "
                                          "(the reference fails on more, cnets_ours.py:938-941,1021)")
            img_start, img_end = first, last + 1
        n_c = Lq if img_start < 0 else Lq - (img_end - img_start) + cfg.num_q - 1
        if n_c + 8 * (cfg.depth + 1) + 8 > self.cache_rows:
            raise ValueError("draft KV cache too small for this prompt")
        if Lq + cfg.depth + 2 > self.rope_rows:
            raise ValueError("prompt positions exceed the draft's rotary table")
        t, sc, ws = self._prefill_ws(max(n_c, img_end - img_start, 1))
        g_new = torch.empty(1, H, dtype=self.dtype, device=dev)
        got = lib.vs_img_adaptor_prefill(C.byref(self.c), C.byref(self.adaptor_c), C.byref(sc.c), C.byref(ws), L.ptr(hid),
                                         L.ptr(emb), Lq, img_start, img_end, L.ptr(g_new), L.ptr(self.scratch.h),
                                         L.ptr(self.state), _stream())
        if got < 0:
            L.check(got, "vs_img_adaptor_prefill")
        assert got == n_c
        self.last_img_hidden = g_new
        self.prefill_packed = t["packed"][:n_c]                              # views kept for the parity tests
        self.prefill_out = sc.h[:n_c]
        self.prefill_pos = t["pos"][:n_c]
        self._expand_after_prefill(n_c, sample_token)

    def _expand_after_prefill(self, n_c: int, sample_token: int):
        cfg, lib, dev = self.cfg, self.lib, self.device
        if use_programs():
            self.state[L.ST_NEXT_TOKEN] = int(sample_token)
            fused = True
            if self._prog_levels.get(fused) is None:
                pr = Program(dev, attn_fused=fused).register_tiled(self.tiled)
                pr.add("draft_levels", C.byref(self.c), C.byref(self.scratch.c), C.byref(self.tb),
                       L.ptr(self.state), 1, self.state.data_ptr() + 4 * L.ST_NEXT_TOKEN, 0)
                self._prog_levels[fused] = pr.commit()
            self._prog_levels[fused].run()
            return
        L.check(lib.vs_draft_levels(C.byref(self.c), C.byref(self.scratch.c), C.byref(self.tb), L.ptr(self.state), 1,
                                    None, int(sample_token), _stream()), "vs_draft_levels")

    @torch.no_grad()
    def prefill_python(self, hidden: torch.Tensor, inputs_embeds: torch.Tensor, image_mask: Optional[torch.Tensor],
                       sample_token: int):
        """hidden [1,L,H] target features; inputs_embeds [1,L,H] ALREADY shifted by one (cnets_ours.py:1066-1082);
        image_mask [1,L] bool in the unshifted frame or None.  Builds the packed sequence (text rows through
        fc/img_fc, the image run compressed by the adaptor to num_q-1 tokens at the run's last positions),
        runs the decoder layer over it, then expands the draft tree."""
        cfg, lib, dev, H, nh = self.cfg, self.lib, self.device, self.cfg.hidden, self.cfg.heads
        hid = hidden[0].to(dev, self.dtype).contiguous()
        emb = inputs_embeds[0].to(dev, self.dtype).contiguous()
        Lq = hid.shape[0]
        pos_all = torch.arange(Lq, device=dev, dtype=torch.int32)
        st = _stream()
        if image_mask is not None:
            im = image_mask[0, 1:].to(dev)                                  # :880 (shifted frame, L-1 entries)
            ends = torch.cat([im[:-1] & ~im[1:], im[-1:]])                  # :881-883
            last_ids = torch.nonzero(ends).flatten().tolist()
            if len(last_ids) != 1:
                raise NotImplementedError("exactly one contiguous image run is supported "
                                          "(the reference fails on more, cnets_ours.py:938-941,1021)")
            end = last_ids[0] + 1
            cur = im[:end]
            pre_idx = torch.nonzero(~cur).flatten()
            img_idx = torch.nonzero(cur).flatten()
            post_idx = torch.arange(end, Lq, device=dev)
            n_img = img_idx.numel()
            # ---- vision adaptor: K,V = Linear(image embeds); num_q learned queries cross-attend; o_proj
            n_pad = _ceil(n_img, 32)
            img = torch.zeros(n_pad, H, dtype=self.dtype, device=dev)
            img[:n_img] = emb[img_idx]
            kbuf = torch.zeros(nh, n_pad, 128, dtype=self.dtype, device=dev)
            vbuf = torch.zeros(nh, n_pad, 128, dtype=self.dtype, device=dev)
            sc = ScratchBuffers(H, cfg.inter, nh, cfg.vocab, n_pad, n_pad, dev, extra_n=2 * H)
            a = L.LinearArgs()
            a.X, a.ldx, a.W, a.ldw, a.bias = L.ptr(img), H, L.ptr(self.w_akv), H, L.ptr(self.b_akv)
            a.T, a.N, a.K = n_img, 2 * H, H
            a.splits = 1
            a.partials, a.counters, a.epi = L.ptr(sc.partials), L.ptr(sc.counters), L.EPI_HEADS
            a.n_q_tiles, a.n_k_tiles, a.rope = 0, nh, 0
            a.Kout, a.Vout, a.kv_head_stride = L.ptr(kbuf), L.ptr(vbuf), n_pad * 128
            L.check(lib.vs_linear(C.byref(a), st), "adaptor k/v projection")
            ao = torch.zeros(32, H, dtype=self.dtype, device=dev)
            at = L.AttnArgs()
            at.Q, at.q_head_stride, at.K, at.V, at.kv_head_stride = L.ptr(self.adpt_q), 32 * 128, L.ptr(kbuf), L.ptr(vbuf), n_pad * 128
            at.n_heads, at.n_kv_heads, at.Tq = nh, nh, cfg.num_q
            at.n_ctx, at.n_ctx_host, at.n_new, at.causal, at.bits, at.eager = None, n_img, 0, 0, None, 0
            at.scale = 1.0 / math.sqrt(128.0)
            at.scores, at.score_ld, at.stats, at.max_chunks = L.ptr(sc.scores), sc.score_ld, L.ptr(sc.stats), sc.max_chunks
            at.opart, at.pv_splits, at.counters, at.O, at.ldo = L.ptr(sc.opart), sc.pv_splits, L.ptr(sc.attn_counters), L.ptr(ao), H
            L.check(lib.vs_attention(C.byref(at), st), "adaptor cross-attention")
            adapted = torch.zeros(32, H, dtype=self.dtype, device=dev)
            self._linear(sc, ao, H, self.w_ao, None, cfg.num_q, H, H, adapted, H)
            g_new = adapted[cfg.num_q - 1:cfg.num_q].clone()
            n_c_img = cfg.num_q - 1
            text_idx = torch.cat([pre_idx, post_idx])
            n_pre = pre_idx.numel()
        else:
            text_idx = torch.arange(Lq, device=dev)
            n_pre, n_c_img, end = Lq, 0, Lq
            g_new = torch.zeros(1, H, dtype=self.dtype, device=dev)        # :977-981 (dummy adaptor adds 0)
            adapted = None
        n_text = text_idx.numel()
        n_c = n_text + n_c_img
        if n_c + 8 * (cfg.depth + 1) + 8 > self.cache_rows:
            raise ValueError("draft KV cache too small for this prompt")
        if Lq + cfg.depth + 2 > self.rope_rows:
            raise ValueError("prompt positions exceed the draft's rotary table")
        rows = _ceil(max(n_c, 1), 32)
        big = ScratchBuffers(H, cfg.inter, nh, cfg.vocab, rows, rows, dev)
        x1 = torch.zeros(rows, 2 * H, dtype=self.dtype, device=dev)
        x2 = torch.zeros(rows, 2 * H, dtype=self.dtype, device=dev)
        # text rows: fc(cat(emb, img_fc(cat(hid, g)))) with g = 0 before the image run and the new g after it
        if n_text:
            x1[:n_text, :H] = hid[text_idx]
            if image_mask is not None and n_text > n_pre:
                x1[n_pre:n_text, H:] = g_new
            x2[:n_text, :H] = emb[text_idx]
            self._linear(big, x1, 2 * H, self.w_img_fc, self.b_img_fc, n_text, H, 2 * H, x2[:, H:], 2 * H)
            fused = torch.zeros(rows, H, dtype=self.dtype, device=dev)
            self._linear(big, x2, 2 * H, self.w_fc, self.b_fc, n_text, H, 2 * H, fused, H)
        # packed order: pre-text, compressed image tokens, post-text (cnets_ours.py:918-950)
        pos = torch.zeros(rows, dtype=torch.int32, device=dev)
        if image_mask is not None:
            big.h[:n_pre] = fused[:n_pre]
            big.h[n_pre:n_pre + n_c_img] = adapted[:n_c_img]
            big.h[n_pre + n_c_img:n_c] = fused[n_pre:n_text]
            pos[:n_pre] = pos_all[pre_idx]
            pos[n_pre:n_pre + n_c_img] = pos_all[end - n_c_img:end]         # last num_q-1 positions of the run (:932-937)
            pos[n_pre + n_c_img:n_c] = pos_all[post_idx]
        else:
            big.h[:n_c] = fused[:n_c]
            pos[:n_c] = pos_all
        self.prefill_packed = big.h[:n_c].clone()                           # packed rows before the layer (parity tests)
        layer = self._layer_struct()
        rope = L.Rope(L.ptr(self.cos), L.ptr(self.sin))
        L.check(lib.vs_decoder_layer(C.byref(self.dims), C.byref(layer), C.byref(big.c), C.byref(rope), n_c, L.ptr(pos),
                                     None, 0, None, 0, n_c, 1, None, 0, st), "draft prefill layer")
        self.last_img_hidden = g_new
        lib.vs_broadcast_row(L.ptr(g_new), L.ptr(self.x1[:, H:]), 2 * H, 32, H, st)
        self.scratch.h[0] = big.h[n_c - 1]
        self.state[L.ST_D_LEN] = n_c
        self.state[L.ST_D_POS] = Lq
        self.state[L.ST_LAST_ROW] = 0
        self.prefill_out = big.h[:n_c]                                      # kept for tests
        self.prefill_pos = pos[:n_c]
        if use_programs():
            self.state[L.ST_NEXT_TOKEN] = int(sample_token)
            fused = True
            if self._prog_levels.get(fused) is None:
                pr = Program(dev, attn_fused=fused).register_tiled(self.tiled)
                pr.add("draft_levels", C.byref(self.c), C.byref(self.scratch.c), C.byref(self.tb),
                       L.ptr(self.state), 1, self.state.data_ptr() + 4 * L.ST_NEXT_TOKEN, 0)
                self._prog_levels[fused] = pr.commit()
            self._prog_levels[fused].run()
            return
        L.check(lib.vs_draft_levels(C.byref(self.c), C.byref(self.scratch.c), C.byref(self.tb), L.ptr(self.state), 1,
                                    None, int(sample_token), st), "vs_draft_levels")

    def _linear(self, sc: ScratchBuffers, x, ldx, w, bias, T, N, K, y, ldy, epi=L.EPI_STORE):
        run_linear(self.lib, sc, x, ldx, w, bias, T, N, K, y, ldy, epi)

    # -- later calls (cnets_ours.py:1090-1092 + tree) ------------------------------------------------------
    def tree_from_accepted(self, hidden_rows: torch.Tensor, ids: torch.Tensor, sample_token: int):
        """Host-driven variant used by the reference-shaped API: hidden_rows [m,H], ids [m] (already shifted)."""
        m = hidden_rows.shape[0]
        if m > L.VS_MAX_ACCEPT:
            raise ValueError("more accepted rows than VS_MAX_ACCEPT")
        H = self.cfg.hidden
        self.x1[:m, :H] = hidden_rows.to(self.device, self.dtype)
        if m < L.VS_MAX_ACCEPT:
            self.x1[m:L.VS_MAX_ACCEPT, :H] = self.x1[m - 1, :H]
        t = self.tree
        ids32 = ids.to(self.device, torch.int32).flatten()
        t["stable_ids"][:m] = ids32
        t["stable_ids"][m:L.VS_MAX_ACCEPT] = ids32[-1]
        dpos = int(self.state[L.ST_D_POS].item())
        t["stable_pos"][:L.VS_MAX_ACCEPT] = torch.arange(dpos, dpos + L.VS_MAX_ACCEPT, dtype=torch.int32, device=self.device)
        self.state[L.ST_M] = m
        self.state[L.ST_LAST_ROW] = m - 1
        self.state[L.ST_NEXT_TOKEN] = int(sample_token)
        if use_programs():
            d_len = int(self.state[L.ST_D_LEN].item())
            fused = True
            if self._prog_tree.get(fused) is None:
                pr = Program(self.device, attn_fused=fused).register_tiled(self.tiled)
                pr.add("draft_tree", C.byref(self.c), C.byref(self.scratch.c), C.byref(self.tb), L.ptr(self.state))
                self._prog_tree[fused] = pr.commit()
            self._prog_tree[fused].run()
            return
        L.check(self.lib.vs_draft_tree(C.byref(self.c), C.byref(self.scratch.c), C.byref(self.tb), L.ptr(self.state),
                                       _stream()), "vs_draft_tree")

    # -- host views of the tree (the return values of topK_genrate, cnets_ours.py:1238) --------------------
    def tree_outputs(self):
        cfg = self.cfg
        n = cfg.total_tokens
        st = self.state.tolist()
        n_leaves, max_depth = st[L.ST_N_LEAVES], st[L.ST_MAX_DEPTH]
        t = self.tree
        draft_tokens = t["draft_tokens"][:n].long()[None]
        retrieve = t["retrieve"].view(32, L.VS_RETRIEVE_LD)[:n_leaves, :max_depth].long().cpu()
        bits = t["tree_bits"][:n]
        tree_mask = ((bits[:, None] >> torch.arange(n, device=self.device)[None]) & 1).float()[None, None].cpu()
        tree_pos = t["tree_pos"][:n].long()
        return draft_tokens, retrieve, tree_mask, tree_pos
