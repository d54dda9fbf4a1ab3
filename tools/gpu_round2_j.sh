#!/bin/bash
mkdir -p gpurun_out
VS_LIB_PATH=$PWD/vispec_b200/lib/variants/mmaconv.so timeout 600 python -m pytest tests/test_gpu_program.py -q -m gpu -x --timeout 240 --timeout-method thread -k "not slab" > gpurun_out/t_mmaconv.log 2>&1; echo mmaconv_tests_rc=$? $(tail -n 1 gpurun_out/t_mmaconv.log)
bash tools/ab_bench.sh base=base mmaconv=vispec_b200/lib/variants/mmaconv.so
