#!/bin/bash
# same-box A/B of register-split variants (tools/build_variant.py NAME -DVS_SETMAXNREG=1 -DVS_ROLE_REGS=.. -DVS_WORKER_REGS=..)
mkdir -p gpurun_out
V=vispec_b200/lib/variants
bash tools/ab_bench.sh base1=base smr64a=$V/smr64.so smr56a=$V/smr.so smr40a=$V/smr40.so base2=base smr64b=$V/smr64.so smr56b=$V/smr.so smr40b=$V/smr40.so
