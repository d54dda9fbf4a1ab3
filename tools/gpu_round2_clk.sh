#!/bin/bash
# clock checkpoints of the attention item + KV producer (needs vispec_b200/lib/variants/clk.so: tools/build_variant.py clk -DVS_PROGRAM_CLK)
mkdir -p gpurun_out
export VS_LIB_PATH=$PWD/vispec_b200/lib/variants/clk.so
for nc in 704 8000; do echo "== N_ATTN=6 N_CTX=$nc (warm, last of 6 phases)"; N_ATTN=6 N_CTX=$nc timeout 120 python tools/clk_attn.py 2>&1 | tail -n 15; done > gpurun_out/clk_attn_r2i.txt 2>&1
cat gpurun_out/clk_attn_r2i.txt
for sh in o down; do echo "== SHAPE=$sh"; SHAPE=$sh timeout 120 python tools/clk_epi.py 2>&1 | tail -n 15; done > gpurun_out/clk_epi_r2i.txt 2>&1
cat gpurun_out/clk_epi_r2i.txt
