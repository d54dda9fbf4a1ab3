#!/bin/bash
mkdir -p gpurun_out
T="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py -q -x -k "attention" $T > gpurun_out/t_attn.log 2>&1; echo attn_rc=$?
tail -n 12 gpurun_out/t_attn.log | cut -c1-300
timeout 300 python tools/bench_attention.py --kv 734 2048 8192 > gpurun_out/attn_ring.jsonl 2>&1; cat gpurun_out/attn_ring.jsonl | cut -c1-400
export VS_LIB_PATH=$PWD/vispec_b200/lib/variants/clk.so
for st in 1; do for nc in 704 8000; do
  echo "== STABLE=$st N_CTX=$nc"; STABLE=$st N_CTX=$nc timeout 120 python tools/clk_attn.py 2>&1 | tail -n 14
done; done > gpurun_out/clk_attn_ring.txt 2>&1
cat gpurun_out/clk_attn_ring.txt
