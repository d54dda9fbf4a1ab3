#!/bin/bash
# same-box A/B of register-split variants (tools/build_variant.py NAME -DVS_ROLE_REGS=.. -DVS_WORKER_REGS=..) against the in-tree build
mkdir -p gpurun_out
V=vispec_b200/lib/variants
bash tools/ab_bench.sh base1=base smr112a=$V/smr112.so smr48a=$V/smr48.so base2=base smr112b=$V/smr112.so smr48b=$V/smr48.so
