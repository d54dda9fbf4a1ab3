#!/usr/bin/env python
"""GEMM-phase and epilogue-phase time of single fused linears (LLaVA-1.6-7B decode shapes, T=30) inside the persistent
program kernel, from the per-phase %globaltimer trace; sweeps the tokens-per-epilogue-item option."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import rope_tables  # noqa: E402
from vispec_b200.program import Program  # noqa: E402

dev, BF = "cuda:0", torch.bfloat16
T, H, I, V = 30, 4096, 11008, 32064
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: (torch.randn(*s, device=dev, generator=g) * 0.02).to(BF)
x = {H: rn(32, H), I: rn(32, I)}
h = rn(32, H)
cos, sin = rope_tables(128, 4096, 10000.0, BF, dev)
pos = torch.arange(700, 732, dtype=torch.int32, device=dev)
base = torch.tensor([700], dtype=torch.int32, device=dev)
Q = torch.zeros(32, 32, 128, dtype=BF, device=dev)
Kc = torch.zeros(32, 1024, 128, dtype=BF, device=dev)
Vc = torch.zeros(32, 1024, 128, dtype=BF, device=dev)
dummy = torch.zeros(4, 64, dtype=BF, device=dev)
shapes = [("qkv/HEADS", 3 * H, H, L.EPI_HEADS), ("o/RESID", H, H, L.EPI_RESID), ("gateup/SWIGLU", 2 * I, H, L.EPI_SWIGLU),
          ("down/RESID", H, I, L.EPI_RESID), ("head/STORE", V, H, L.EPI_STORE)]
for name, N, K, epi in shapes:
    w = rn(N, K)
    y = torch.zeros(32, max(N, H), dtype=BF, device=dev)
    for tpi in (0, 4, 8, 16):
        a = L.LinearArgs()
        a.X, a.ldx, a.W, a.ldw = x[K].data_ptr(), K, w.data_ptr(), K
        a.T, a.N, a.K, a.splits, a.epi = T, N, K, 1, epi
        a.Y, a.ldy = (h.data_ptr(), H) if epi == L.EPI_RESID else (y.data_ptr(), y.stride(0))
        if epi == L.EPI_RESID:
            a.R, a.ldr = h.data_ptr(), H
        if epi == L.EPI_HEADS:
            a.n_q_tiles, a.n_k_tiles, a.rope = 32, 32, 1
            a.cos, a.sin, a.pos_ids = cos.data_ptr(), sin.data_ptr(), pos.data_ptr()
            a.Qout, a.q_head_stride, a.Kout, a.Vout, a.kv_head_stride = Q.data_ptr(), 32 * 128, Kc.data_ptr(), Vc.data_ptr(), 1024 * 128
            a.kv_row_base = base.data_ptr()
        p = Program(dev, trace=True, pf_ahead=0)
        p.set_option("tpi", tpi)
        p.add("gather_rows", dummy.data_ptr(), 64, None, dummy[2:].data_ptr(), 64, 1, 64)
        p.add("linear", C.byref(a))
        p.add("gather_rows", dummy.data_ptr(), 64, None, dummy[2:].data_ptr(), 64, 1, 64)
        p.commit()
        ge, ee = [], []
        for _ in range(6):
            torch.randn(64 * 1024 * 1024, device=dev)          # evict the weights from L2
            p.run()
            torch.cuda.synchronize()
            us = p.read_trace_us()
            ge.append(us[1]); ee.append(us[2])
        mb = N * K * 2 / 1e6
        gm, em = sorted(ge)[len(ge) // 2], sorted(ee)[len(ee) // 2]
        print(f"{name:14s} tpi={tpi:2d}  W={mb:6.1f} MB  gemm {gm:6.2f} us ({mb / gm / 1e3:5.2f} TB/s)  epilogue {em:6.2f} us")
