#!/usr/bin/env python
"""Checkpoint clocks of one residual-epilogue item with a folded-in RMSNorm, inside a chain of [o_proj-shaped linear + norm]
pairs whose input is the previous pair's output (needs a -DVS_PROGRAM_CLK build: tools/build_variant.py clk
-DVS_PROGRAM_CLK; VS_LIB_PATH=vispec_b200/lib/variants/clk.so).  SHAPE=o|down."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L
from vispec_b200.engine import pack_weight_tiles
from vispec_b200.program import Program
dev, BF = "cuda:0", torch.bfloat16
T, N = 30, 4096
K = 11008 if os.environ.get("SHAPE", "o") == "down" else 4096
NP = 6
xs = (torch.randn(32, K, device=dev) * 0.02).to(BF)
ws = [(torch.randn(N, K, device=dev) * 0.02).to(BF) for _ in range(NP)]
wt = [pack_weight_tiles(w) for w in ws]
h = torch.randn(32, N, device=dev).to(BF); nw = torch.ones(N, device=dev, dtype=BF)
p = Program(dev, pf_ahead=0, trace=True)
p.register_tiled(list(zip(ws, wt)))
args = []
for i in range(NP):
    a = L.LinearArgs(); a.X, a.ldx, a.W, a.ldw = xs.data_ptr(), K, ws[i].data_ptr(), K
    a.T, a.N, a.K, a.splits, a.epi = T, N, K, 1, L.EPI_RESID
    a.Y, a.ldy, a.R, a.ldr = h.data_ptr(), N, h.data_ptr(), N
    args.append(a)
    p.add("linear", C.byref(a))
    p.add("rmsnorm", h.data_ptr(), N, nw.data_ptr(), xs.data_ptr(), K, T, N, 1e-5)    # writes the first N columns of the next input
p.commit()
ty = p.phase_types()
lib = L.load(); lib.vs_debug_clk.argtypes = [C.c_void_p]
for _ in range(3): p.run()
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
assert lib.vs_debug_clk(buf) == 64, "library not built with -DVS_PROGRAM_CLK"
ck = list(buf)
us = p.read_trace_us()
print("phase types", ty)
print("trace us", [round(u, 2) for u in us])
last_epi = max(i for i, t in enumerate(ty) if t == L.PH_EPI)
t0 = ck[32 + ((last_epi - 1) & 15)]      # end of the last GEMM phase on CTA 0 (no grid barrier: its workers drained their tiles)
tb = ck[32 + ((last_epi - 2) & 15)]      # end of the grid barrier before the last GEMM phase
print("  last GEMM phase on CTA 0: %.2f us from the barrier to 'tiles drained'; tiles published at %s us" %
      ((t0 - tb) / 1.965e3, [round((ck[48 + i] - tb) / 1.965e3, 2) for i in range(4) if ck[48 + i] > tb]))
names = {0: "phase fn entered", 1: "item start", 2: "operand loads issued", 7: "contributors arrived", 3: "partials reduced",
         4: "values consumed", 8: "rows + sums of squares stored", 9: "all tiles' sums gathered", 5: "item end", 6: "fn end"}
prev = t0
for i in (0, 1, 2, 7, 3, 4, 8, 9, 5, 6):
    c = ck[16 + i]
    if c < t0: continue
    print(f"  {names[i]:30s} +{c - prev:6d} cycles   (t = {(c - t0) / 1.965e3:.2f} us)")
    prev = c
print("  grid barrier done              +%6d cycles   (t = %.2f us)" % (ck[32 + (last_epi & 15)] - prev, (ck[32 + (last_epi & 15)] - t0) / 1.965e3))
