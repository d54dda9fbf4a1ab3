#!/bin/bash
mkdir -p gpurun_out
export VS_LIB_PATH=$PWD/vispec_b200/lib/variants/clk.so
for st in 1 0; do for nc in 704 8000; do
  echo "== STABLE=$st N_CTX=$nc"; STABLE=$st N_CTX=$nc timeout 120 python tools/clk_attn.py 2>&1 | tail -n 16
done; done > gpurun_out/clk_attn_ring.txt 2>&1
cat gpurun_out/clk_attn_ring.txt
