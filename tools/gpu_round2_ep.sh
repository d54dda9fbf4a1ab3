#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:program_kernel -s 2 -c 1 -o gpurun_out/prof_epi_chain python tools/clk_epi.py > gpurun_out/ncu_epi_chain.log 2>&1; echo ncu_rc=$?
timeout 120 python tools/bench_icache.py > gpurun_out/icache.txt 2>&1; cat gpurun_out/icache.txt | cut -c1-250
