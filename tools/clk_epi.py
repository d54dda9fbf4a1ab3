#!/usr/bin/env python
"""Checkpoint clocks of one residual-epilogue item (needs a -DVS_PROGRAM_CLK build: VS_NVCC_EXTRA=-DVS_PROGRAM_CLK)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L
from vispec_b200.program import Program
dev, BF = "cuda:0", torch.bfloat16
T, N, K = 30, 4096, 4096
x = (torch.randn(32, K, device=dev) * 0.02).to(BF); w = (torch.randn(N, K, device=dev) * 0.02).to(BF); h = torch.randn(32, N, device=dev).to(BF)
a = L.LinearArgs(); a.X, a.ldx, a.W, a.ldw = x.data_ptr(), K, w.data_ptr(), K
a.T, a.N, a.K, a.splits, a.epi = T, N, K, 1, L.EPI_RESID
a.Y, a.ldy, a.R, a.ldr = h.data_ptr(), N, h.data_ptr(), N
p = Program(dev, pf_ahead=0, trace=True)
for _ in range(6): p.add("linear", C.byref(a))
p.commit()
lib = L.load(); lib.vs_debug_clk.argtypes = [C.c_void_p]
for _ in range(3): p.run()
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
assert lib.vs_debug_clk(buf) == 64, "library not built with -DVS_PROGRAM_CLK"
ck = list(buf)
print("trace us", [round(u, 2) for u in p.read_trace_us()])
t0 = ck[32 + 10]      # barrier end of phase 10 (last GEMM); phase 11 = last epilogue
names = ["phase fn entered", "item start", "operand loads issued", "reduction loop done", "values consumed", "item end", "fn end"]
prev = t0
for i in range(7):
    print(f"  {names[i]:24s} +{ck[i] - prev:6d} cycles   (t = {(ck[i] - t0) / 1.965e3:.2f} us)")
    prev = ck[i]
