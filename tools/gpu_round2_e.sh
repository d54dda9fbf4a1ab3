#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
timeout 600 python tools/bench_slab_phases.py --json gpurun_out/slab_phases.json > gpurun_out/slab_phases.txt 2>&1; echo phases_rc=$?
VS_X3D=0 timeout 300 python tools/bench_slab_phases.py --shapes o,gateup --debug 0,2 > gpurun_out/slab_phases_x2d.txt 2>&1; echo x2d_rc=$?
timeout 300 python tools/bench_slab_phases.py --shapes o,down --debug 0 --T 8 > gpurun_out/slab_phases_t8.txt 2>&1; echo t8_rc=$?
cat gpurun_out/slab_phases.txt | cut -c1-200; echo ---x2d; cat gpurun_out/slab_phases_x2d.txt | cut -c1-200; echo ---t8; cat gpurun_out/slab_phases_t8.txt | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -n 8 gpurun_out/t_all.log
