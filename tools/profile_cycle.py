#!/usr/bin/env python
"""Run a few draft+verify cycles of the BASELINE config-2 workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (see profiles/README.md for the exact commands)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_llava7b  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--layers", type=int, default=32)
args = ap.parse_args()

model, ids, emb, mask = build_llava7b(device="cuda:0", seed=0, layers=args.layers)
tt = model.spec_layer.total_tokens


def run(n):
    return model.specgenerate(ids, inputs_embeds=emb, image_mask=mask, log=True, return_acceptance_len=True,
                              max_new_tokens=10 ** 9, max_length=n + tt + 10)


run(args.warm)
torch.cuda.synchronize()
# prefill outside the profiled region: profile only the decode cycles
from vispec_b200.model import utils as vu  # noqa: E402
from vispec_b200.model.kv_cache import initialize_past_key_values  # noqa: E402
import ctypes as C  # noqa: E402
from vispec_b200 import _lib as L  # noqa: E402

model.spec_layer.reset_kv()
model.current_length_data.zero_()
vu.reset_tree_mode(model)
vu.initialize_tree(ids, model, model.past_key_values, None, emb, None, image_mask=mask)
torch.cuda.synchronize()
eng_t, eng_d = model.base_model_engine, model.spec_layer.engine
st = torch.cuda.current_stream().cuda_stream
torch.cuda.profiler.start()
for _ in range(args.cycles):
    L.check(eng_t.lib.vs_spec_cycle(C.byref(eng_t.c), C.byref(eng_t.scratch.c), C.byref(eng_d.c), C.byref(eng_d.scratch.c),
                                    C.byref(eng_d.tb), L.ptr(eng_t.state), L.ptr(eng_t.kv), eng_t.kv_max,
                                    L.ptr(eng_t.hidden_out), L.ptr(eng_t.logits), L.ptr(eng_t.amax), L.ptr(eng_d.out_rec), st))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.cycles, "cycles; state", eng_t.state.tolist())
