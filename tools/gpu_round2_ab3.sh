#!/bin/bash
# same-box A/B of build variants (tools/build_variant.py NAME -D...) against the in-tree build
mkdir -p gpurun_out
V=vispec_b200/lib/variants
bash tools/ab_bench.sh base1=base epi1a=$V/epi1.so epi7a=$V/epi7.so base2=base epi1b=$V/epi1.so epi7b=$V/epi7.so
