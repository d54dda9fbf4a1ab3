#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_slab_phases.py --shapes o,down,gateup --debug 3,19,7 --reps 8 > gpurun_out/slab_phases3.txt 2>&1; echo rc=$?
cut -c1-150 gpurun_out/slab_phases3.txt
