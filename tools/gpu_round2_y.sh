#!/bin/bash
mkdir -p gpurun_out
for sp in 9 8 6 5; do echo "== VS_ATTN_SPLITS=$sp"; VS_ATTN_SPLITS=$sp timeout 300 python tools/bench_attention.py --kv 734 990 2048 8192 2>&1 | cut -c1-130; done > gpurun_out/attn_splits.txt 2>&1
cat gpurun_out/attn_splits.txt
