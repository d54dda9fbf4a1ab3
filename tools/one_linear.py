#!/usr/bin/env python
"""One fused linear (o-projection shape, residual epilogue) through the program kernel; for ncu source-level profiles."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.program import Program  # noqa: E402

tpi = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev, BF = "cuda:0", torch.bfloat16
T, N, K = 30, 4096, 4096
x = (torch.randn(32, K, device=dev) * 0.02).to(BF)
w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
h = (torch.randn(32, N, device=dev)).to(BF)
a = L.LinearArgs()
a.X, a.ldx, a.W, a.ldw = x.data_ptr(), K, w.data_ptr(), K
a.T, a.N, a.K, a.splits, a.epi = T, N, K, 1, L.EPI_RESID
a.Y, a.ldy, a.R, a.ldr = h.data_ptr(), N, h.data_ptr(), N
p = Program(dev, pf_ahead=0)
p.set_option("tpi", tpi)
for _ in range(4):
    p.add("linear", C.byref(a))
p.commit()
for _ in range(3):
    p.run()
torch.cuda.synchronize()
print("done")
