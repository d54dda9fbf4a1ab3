#!/bin/bash
# quick parity pass of the in-tree build, then the same-box A/B against variants/head.so
mkdir -p gpurun_out
T="--timeout 180 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_e2e.py tests/test_gpu_sampling.py -q -x $T > gpurun_out/t_x.log 2>&1; echo tests_rc=$?
tail -n 3 gpurun_out/t_x.log | cut -c1-300
bash tools/ab_bench.sh head1=vispec_b200/lib/variants/head.so new1=base head2=vispec_b200/lib/variants/head.so new2=base
