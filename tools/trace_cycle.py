#!/usr/bin/env python
"""Per-phase timing of ONE draft+verify cycle of the persistent program kernel (BASELINE config-2 workload):
%globaltimer stamps written by CTA 0 at the end of every phase (vs_program option "trace").  Prints the time per
phase type and the phases of one target layer; `--json` dumps everything."""
import argparse
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_llava7b  # noqa: E402
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.model import utils as vu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--pf-ahead", type=int, default=None)
ap.add_argument("--cycles", type=int, default=4)
ap.add_argument("--inflight", type=int, default=None)
ap.add_argument("--brief", action="store_true")
ap.add_argument("--two-phase-attention", action="store_true")
ap.add_argument("--json", default=None)
args = ap.parse_args()

model, ids, emb, mask = build_llava7b(device="cuda:0", seed=0, layers=args.layers)
tt = model.spec_layer.total_tokens
model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, max_new_tokens=8)          # warm-up, allocates the KV views
model.spec_layer.reset_kv()
model.current_length_data.zero_()
vu.reset_tree_mode(model)
vu.initialize_tree(ids, model, model.past_key_values, None, emb, None, image_mask=mask)
torch.cuda.synchronize()
prog = model.cycle_program(trace=True, pf_ahead=args.pf_ahead, inflight=args.inflight,
                           attn_fused=not args.two_phase_attention)
types = prog.phase_types()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.cycles + 1)]
ev[0].record()
for i in range(args.cycles):
    prog.run()
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.cycles)]
us = prog.read_trace_us()          # last cycle
tot = collections.defaultdict(lambda: [0, 0.0])
for t, d in zip(types, us):
    tot[L.PHASE_NAMES[t]][0] += 1
    tot[L.PHASE_NAMES[t]][1] += d
total = sum(us)
print(f"cycle launches (CUDA events): {['%.3f' % m for m in ms]} ms; traced phases: {len(types)}, sum {total / 1e3:.3f} ms")
for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:16s} n={n:4d} total={d:9.1f} us  avg={d / n:7.2f} us  share={100 * d / total:5.1f}%")
if args.brief:
    sys.exit(0)
# one target layer in the middle of the verify step: from the GEMM in front of the 6th attention phase to the one in
# front of the 7th (works for the slab program -- 5 phases per layer -- and the round-1 stream-K program -- 13)
attn = [i for i, t in enumerate(types) if t in (L.PH_ATTN_FUSED, L.PH_ATTN_S)]
gemm_types = (L.PH_GEMM, L.PH_GEMM2)


def layer_start(ai):
    i = ai - 1
    while i > 0 and types[i] not in gemm_types:
        i -= 1
    while i > 0 and types[i - 1] in (L.PH_GEMM, L.PH_EPI) and types[i] == L.PH_EPI:
        i -= 1
    return i


start, stop = (layer_start(attn[5]), layer_start(attn[6])) if len(attn) > 7 else (0, min(len(types), 16))
print("phases of one target layer:")
for i in range(start, stop):
    print(f"  [{i}] {L.PHASE_NAMES[types[i]]:16s} {us[i]:8.2f} us")
end, done = prog.read_trace_detail()
print("per phase of that layer: work = slowest CTA's items-done minus previous barrier; sync = barrier end minus that")
for i in range(start, stop):
    t0 = end[i - 1]
    d = sorted(x - t0 for x in done[i])
    print(f"  [{i}] {L.PHASE_NAMES[types[i]]:16s} work: min {d[0] / 1e3:6.2f} median {d[len(d) // 2] / 1e3:6.2f} "
          f"max {d[-1] / 1e3:6.2f} us   sync {(end[i] - t0 - d[-1]) / 1e3:5.2f} us")
# the tail of the cycle: LM head, accept, splice, the four draft forwards
tail0 = layer_start(attn[tcfg_layers - 1]) if (tcfg_layers := model.base_model_engine.cfg.layers) <= len(attn) else 0
print("after the last target layer (phase: us):", " ".join(f"{L.PHASE_NAMES[types[i]][:9]}:{us[i]:.1f}" for i in range(tail0, len(types))))
if args.json:
    json.dump({"types": types, "us": us, "cycle_ms": ms, "end_ns": end, "done_ns": done}, open(args.json, "w"))
