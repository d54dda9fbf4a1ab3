#!/bin/bash
mkdir -p gpurun_out
T="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py -q -x -k "attention" $T > gpurun_out/t_attn.log 2>&1; echo attn_rc=$?
tail -n 5 gpurun_out/t_attn.log | cut -c1-300
for pf in 0 8 16 32; do echo "== VS_ATTN_PF=$pf"; VS_ATTN_PF=$pf timeout 300 python tools/bench_attention.py --kv 734 2048 8192 2>&1 | cut -c1-200; done > gpurun_out/attn_pf.txt 2>&1
cat gpurun_out/attn_pf.txt
