#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_bench.sh base=base nopre=vispec_b200/lib/variants/nopre.so base2=base
