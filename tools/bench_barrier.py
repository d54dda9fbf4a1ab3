#!/usr/bin/env python
"""Cost of one phase boundary of the persistent program kernel: a program of N one-row copy phases (64 bf16 each), so a
phase is essentially one grid barrier.  Prints the mean and the distribution from the per-phase %globaltimer trace."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200.program import Program  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = "cuda:0"
src = torch.randn(4, 64, device=dev).to(torch.bfloat16)
dst = torch.zeros(4, 64, dtype=torch.bfloat16, device=dev)
p = Program(dev, trace=True)
for i in range(N):
    p.add("gather_rows", src.data_ptr(), 64, None, dst.data_ptr(), 64, 1, 64)
p.commit()
for _ in range(3):
    p.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
p.run()
e1.record()
torch.cuda.synchronize()
us = sorted(p.read_trace_us()[1:])
print(f"{N} phases in {e0.elapsed_time(e1) * 1e3:.1f} us; per phase: mean {sum(us) / len(us):.2f} us, "
      f"p10 {us[len(us) // 10]:.2f}, median {us[len(us) // 2]:.2f}, p90 {us[len(us) * 9 // 10]:.2f}, max {us[-1]:.2f}")
