#!/usr/bin/env python
"""Is the per-phase cost instruction-cache bound?  A program of 10 identical o-projection linears (GEMM + residual
epilogue): the first epilogue runs cold code, the later ones hot code.  Then the same with an attention phase pair
(different code) between the linears."""
import ctypes as C, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L
from vispec_b200.engine import ScratchBuffers
from vispec_b200.program import Program
dev, BF = "cuda:0", torch.bfloat16
T, N, K = 30, 4096, 4096
x = (torch.randn(32, K, device=dev) * 0.02).to(BF); w = (torch.randn(N, K, device=dev) * 0.02).to(BF); h = torch.randn(32, N, device=dev).to(BF)
a = L.LinearArgs(); a.X, a.ldx, a.W, a.ldw = x.data_ptr(), K, w.data_ptr(), K
a.T, a.N, a.K, a.splits, a.epi = T, N, K, 1, L.EPI_RESID
a.Y, a.ldy, a.R, a.ldr = h.data_ptr(), N, h.data_ptr(), N
nh, n_ctx = 32, 704
Q = torch.randn(nh, 32, 128, device=dev).to(BF); Kc = torch.randn(nh, 1024, 128, device=dev).to(BF); Vc = torch.randn(nh, 1024, 128, device=dev).to(BF)
sc = ScratchBuffers(nh * 128, 256, nh, 1000, 32, 1024, dev)
O = torch.zeros(T, nh * 128, dtype=BF, device=dev)
bits = torch.tensor([(1 << (i + 1)) - 1 for i in range(T)], dtype=torch.int64, device=dev)
at = L.AttnArgs()
at.Q, at.q_head_stride, at.K, at.V, at.kv_head_stride = Q.data_ptr(), 32 * 128, Kc.data_ptr(), Vc.data_ptr(), 1024 * 128
at.n_heads, at.n_kv_heads, at.Tq = nh, nh, T
at.n_ctx, at.n_ctx_host, at.n_new, at.causal, at.bits, at.eager, at.scale = None, n_ctx, T, 0, bits.data_ptr(), 1, 1 / math.sqrt(128)
at.scores, at.score_ld, at.stats, at.max_chunks = sc.scores.data_ptr(), sc.score_ld, sc.stats.data_ptr(), sc.max_chunks
at.opart, at.pv_splits, at.counters, at.O, at.ldo = sc.opart.data_ptr(), sc.pv_splits, sc.attn_counters.data_ptr(), O.data_ptr(), nh * 128
for interleave in (False, True):
    p = Program(dev, trace=True, pf_ahead=0, attn_fused=True)
    for i in range(10):
        p.add("linear", C.byref(a))
        if interleave:
            p.add("attention", C.byref(at))
    p.commit()
    for _ in range(3):
        p.run()
    torch.cuda.synchronize()
    us = p.read_trace_us()
    ty = p.phase_types()
    print("interleaved with attention" if interleave else "linears only")
    for name, code in (("gemm", L.PH_GEMM), ("epilogue", L.PH_EPI), ("attention", L.PH_ATTN_FUSED)):
        v = [round(u, 2) for u, t in zip(us, ty) if t == code]
        if v:
            print(f"  {name:12s}", v)
