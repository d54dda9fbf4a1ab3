"""Golden vectors for vispec_b200.model.utils.generate_tree_buffers from the UNMODIFIED reference
(vispec/model/utils.py:90-218, vispec/model/choices.py).  Run in the build container (needs /root/reference):
    python tests/golden/make_golden_tree_buffers.py"""
import json
import os
import random
import sys

import torch  # noqa: E402

torch.cuda.synchronize = lambda *a, **k: None      # caller-side shim: the reference's Timer syncs CUDA (utils.py:30); no GPU here
sys.path.insert(0, "/root/reference")
from vispec.model import choices  # noqa: E402
from vispec.model.utils import generate_tree_buffers  # noqa: E402


def random_tree(seed, n, width=4, depth=5):
    rng = random.Random(seed)
    paths = {(rng.randrange(width),)}
    while len(paths) < n:
        p = rng.choice(sorted(paths))
        q = p + (rng.randrange(width),) if (len(p) < depth and rng.random() < 0.7) else (rng.randrange(width),)
        paths.add(q)
    # every prefix of a path must be in the tree
    full = set()
    for p in paths:
        for k in range(1, len(p) + 1):
            full.add(p[:k])
    out = [list(p) for p in full]
    rng.shuffle(out)
    return out


cases = {"mc_sim_7b_63": choices.mc_sim_7b_63, "chain": [[0], [0, 0], [0, 0, 0]], "flat": [[2], [0], [1]],
         "single": [[0]], "rand_a": random_tree(1, 12), "rand_b": random_tree(2, 40), "rand_c": random_tree(3, 25, width=9)}
out = {}
for name, tc in cases.items():
    b = generate_tree_buffers(tc, device="cpu")
    out[name] = {"tree_choices": tc, **{k: v.tolist() for k, v in b.items()}}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tree_buffers.json")
json.dump(out, open(path, "w"))
print("wrote", path, {k: len(v["tree_choices"]) for k, v in out.items()})
