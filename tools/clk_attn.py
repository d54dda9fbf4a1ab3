#!/usr/bin/env python
"""Checkpoint clocks of one fused-attention item (head 0, split 0 on CTA 0) inside a [linear][attention] x N program
(needs a -DVS_PROGRAM_CLK build: VS_NVCC_EXTRA=-DVS_PROGRAM_CLK python vispec_b200/build.py --force)."""
import ctypes as C, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L
from vispec_b200.engine import ScratchBuffers
from vispec_b200.program import Program
dev, BF = "cuda:0", torch.bfloat16
nh, T, n_ctx = 32, 30, int(os.environ.get("N_CTX", "704"))
N_ATTN = int(os.environ.get("N_ATTN", "1"))
Q = torch.randn(nh, 32, 128, device=dev).to(BF); K = torch.randn(N_ATTN, nh, 8192, 128, device=dev).to(BF); V = torch.randn(N_ATTN, nh, 8192, 128, device=dev).to(BF)
sc = ScratchBuffers(nh * 128, 256, nh, 1000, 32, 8192, dev)
O = torch.zeros(T, nh * 128, dtype=BF, device=dev)
bits = torch.tensor([(1 << (i + 1)) - 1 for i in range(T)], dtype=torch.int64, device=dev)
nctx = torch.tensor([n_ctx], dtype=torch.int32, device=dev)
def mk(i):
    a = L.AttnArgs()
    a.Q, a.q_head_stride, a.K, a.V, a.kv_head_stride = Q.data_ptr(), 32 * 128, K[i].data_ptr(), V[i].data_ptr(), 8192 * 128
    a.n_heads, a.n_kv_heads, a.Tq = nh, nh, T
    a.n_ctx, a.n_ctx_host, a.n_new, a.causal, a.bits, a.eager, a.scale = nctx.data_ptr(), 0, T, 0, bits.data_ptr(), 1, 1 / math.sqrt(128)
    a.scores, a.score_ld, a.stats, a.max_chunks = sc.scores.data_ptr(), sc.score_ld, sc.stats.data_ptr(), sc.max_chunks
    a.opart, a.pv_splits, a.counters, a.O, a.ldo = sc.opart.data_ptr(), sc.pv_splits, sc.attn_counters.data_ptr(), O.data_ptr(), nh * 128
    a.prefix_stable = int(os.environ.get("STABLE", "1"))
    return a
args = [mk(i) for i in range(N_ATTN)]
dummy = torch.zeros(4, 64, dtype=BF, device=dev)
p = Program(dev, trace=True, attn_fused=True)
p.add("gather_rows", dummy.data_ptr(), 64, None, dummy[2:].data_ptr(), 64, 1, 64)
for a in args:
    p.add("attention", C.byref(a))
p.add("gather_rows", dummy.data_ptr(), 64, None, dummy[2:].data_ptr(), 64, 1, 64)
p.commit()
lib = L.load(); lib.vs_debug_clk.argtypes = [C.c_void_p]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): p.run()
torch.cuda.synchronize()
if os.environ.get("FLUSH", "0") == "1":
    flush.zero_()
p.run(); torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
assert lib.vs_debug_clk(buf) == 64, "library not built with -DVS_PROGRAM_CLK"
ck = list(buf)
print("phase times us:", [round(x, 2) for x in p.read_trace_us()])
t0 = ck[32 + ((N_ATTN - 1) & 15)]          # barrier end of the phase before the last attention phase
names = ["item start", "Q in smem", "first K granule landed", "S done + statistics (this thread)", "team barrier", "-",
         "statistics exchanged, denominators", "-", "first V granule landed (after P)", "PV done", "team O reduction + partial published",
         "outputs exchanged + reduced"]
print("  SM clock during the phase: %.0f MHz" % ((ck[32 + (N_ATTN & 15)] - t0) / p.read_trace_us()[N_ATTN]))
print("  KV producer (cycles after the barrier of phase 0): start allowed %d, kv length read %d, first K issued %d, first V issued %d, all issued %d"
      % tuple(ck[i] - t0 for i in (20, 21, 22, 24, 23)))
print("  dispatch (cycles after t0): barrier released %d, loop top %d, before switch %d, kv length known %d, phase function entered %d, item start %d"
      % tuple(ck[i] - t0 for i in (52 + ((N_ATTN - 1) & 3), 56, 57, 59, 58, 0)))
prev = t0
for i in range(12):
    if ck[i] == 0: continue
    print(f"  {names[i]:28s} +{ck[i] - prev:6d} cycles   (t = {(ck[i] - t0) / 1.965e3:.2f} us)")
    prev = ck[i]
