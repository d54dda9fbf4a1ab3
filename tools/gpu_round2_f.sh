#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_slab_phases.py --shapes o,gateup --debug 3,19,35,51,0,16 --reps 8 > gpurun_out/slab_phases_mma.txt 2>&1; echo rc=$?
cut -c1-160 gpurun_out/slab_phases_mma.txt
