#!/usr/bin/env python
"""Run a few draft+verify cycles of the BASELINE config-2 workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (see profiles/README.md for the exact commands).  Default: the product path (one
persistent program_kernel launch per cycle); VS_MULTI_KERNEL=1 profiles the one-launch-per-op path."""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_llava7b  # noqa: E402
from vispec_b200 import _lib as L  # noqa: E402
from vispec_b200.engine import use_programs  # noqa: E402
from vispec_b200.model import utils as vu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--layers", type=int, default=32)
args = ap.parse_args()

model, ids, emb, mask = build_llava7b(device="cuda:0", seed=0, layers=args.layers)
model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, max_new_tokens=8)      # warm-up
torch.cuda.synchronize()
# prefill outside the profiled region: profile only the decode cycles
model.spec_layer.reset_kv()
model.current_length_data.zero_()
vu.reset_tree_mode(model)
vu.initialize_tree(ids, model, model.past_key_values, None, emb, None, image_mask=mask)
torch.cuda.synchronize()
eng_t, eng_d = model.base_model_engine, model.spec_layer.engine
st = torch.cuda.current_stream().cuda_stream
prog = model.cycle_program() if use_programs() else None
torch.cuda.profiler.start()
for _ in range(args.cycles):
    if prog is not None:
        prog.run()
    else:
        L.check(eng_t.lib.vs_spec_cycle(C.byref(eng_t.c), C.byref(eng_t.scratch.c), C.byref(eng_d.c), C.byref(eng_d.scratch.c),
                                        C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.kv), eng_t.kv_max,
                                        L.ptr(eng_t.hidden_out), L.ptr(eng_t.logits), L.ptr(eng_t.amax), L.ptr(eng_d.out_rec), st))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.cycles, "cycles; state", eng_t.state.tolist())
