#!/usr/bin/env python
"""Build an alternative libvispec_b200 with extra nvcc flags for A/B runs: tools/build_variant.py NAME -DMACRO=1 ...
-> vispec_b200/lib/variants/NAME.so (git-ignored, shipped to the GPU box; select with VS_LIB_PATH)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vispec_b200 import build as b  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(b.LIB_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, name + ".so")
cmd = [b._nvcc()] + b.NVCC_FLAGS + extra + [os.path.join(b.CSRC, s) for s in b.SOURCES] + ["-o", out]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode:
    sys.exit(r.stdout + r.stderr)
print(out)
