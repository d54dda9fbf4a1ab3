#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_bench.sh base=base mmaconv=vispec_b200/lib/variants/mmaconv.so
