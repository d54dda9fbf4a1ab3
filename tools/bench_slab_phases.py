#!/usr/bin/env python
"""Where does a slab GEMM phase spend its time?  Builds ONE program of `--reps` slab linears of one shape (distinct weight
copies, so nothing is L2-resident between phases), runs it with the kernel's ablation switches (vs_program option "debug":
1 = no activation loads, 2 = no weight loads, 4 = no MMAs, 8 = workers back off) and prints the mean phase time from the
%globaltimer trace.  Results of ablated runs are garbage by construction; only the timings mean something."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import pack_gate_up, pack_weight_slab  # noqa: E402
from vispec_b200.program import Program  # noqa: E402

SHAPES = {"o": (4096, 4096, L.EPI_STORE), "down": (4096, 11008, L.EPI_STORE), "qkv": (12288, 4096, L.EPI_STORE),
          "gateup": (22016, 4096, L.EPI_SWIGLU), "head": (32064, 4096, L.EPI_STORE), "fc": (4096, 8192, L.EPI_STORE)}
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--T", type=int, default=30)
ap.add_argument("--shapes", default="o,down,gateup,head")
ap.add_argument("--debug", default="0,1,2,3,4,8")
ap.add_argument("--json", default=None)
args = ap.parse_args()
dev = "cuda:0"
out = []
for name in args.shapes.split(","):
    N, K, epi = SHAPES[name]
    g = torch.Generator(device=dev).manual_seed(1)
    ws = [(torch.randn(N, K, generator=g, device=dev) * K ** -0.5).to(torch.bfloat16) for _ in range(args.reps)]
    slabs = [(w, pack_weight_slab(w, epi), epi) for w in ws]
    x = torch.randn(32, K, generator=g, device=dev).to(torch.bfloat16)
    y = torch.zeros(32, N, dtype=torch.bfloat16, device=dev)
    for dbg in [int(v) for v in args.debug.split(",")]:
        p = Program(dev, trace=True)
        p.set_option("debug", dbg)
        p.register_tiled(slabs)
        keep = []
        for w in ws:
            a = L.LinearArgs()
            a.X, a.ldx, a.W, a.ldw, a.bias = x.data_ptr(), K, w.data_ptr(), K, None
            a.T, a.N, a.K, a.splits, a.epi = args.T, N, K, 1, epi
            a.Y, a.ldy = y.data_ptr(), (N // 2 if epi == L.EPI_SWIGLU else N)
            keep.append(a)
            p.add("linear", C.byref(a))
        p.commit()
        for _ in range(3):
            p.run()
        torch.cuda.synchronize()
        us = p.read_trace_us()[2:]
        mb = N * K * 2 / 1e6
        rec = {"shape": name, "N": N, "K": K, "T": args.T, "debug": dbg, "us_per_phase": sum(us) / len(us),
               "weights_mb": mb, "gbs": mb / (sum(us) / len(us)) * 1e3}
        out.append(rec)
        print(json.dumps(rec), flush=True)
        del p
    del ws, slabs
    torch.cuda.empty_cache()
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)
