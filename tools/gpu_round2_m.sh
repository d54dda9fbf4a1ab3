#!/bin/bash
# TMA-staged attention bring-up: attention tests of the program kernel, then the attention bench
mkdir -p gpurun_out
T="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py -q -x -k "attention" $T > gpurun_out/t_attn.log 2>&1; echo attn_rc=$?
tail -n 25 gpurun_out/t_attn.log | cut -c1-400
timeout 300 python tools/bench_attention.py --kv 734 2048 8192 --no-stage > gpurun_out/attn_ring.jsonl 2>&1; cat gpurun_out/attn_ring.jsonl | cut -c1-600
