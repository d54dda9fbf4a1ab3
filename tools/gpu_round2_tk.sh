#!/bin/bash
mkdir -p gpurun_out
T="--timeout 240 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_e2e.py -q -x $T > gpurun_out/t_tk.log 2>&1; echo tests_rc=$?
tail -n 30 gpurun_out/t_tk.log | cut -c1-300
