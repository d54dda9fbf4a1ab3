#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:program_kernel -s 3 -c 1 -o gpurun_out/prof_attn8192_ring python tools/bench_attention.py --kv 8192 --layers 8 --reps 2 > gpurun_out/ncu_attn_ring.log 2>&1; echo ncu_rc=$?
tail -n 3 gpurun_out/ncu_attn_ring.log
