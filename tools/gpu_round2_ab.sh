#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_bench.sh base=base prev=vispec_b200/lib/variants/prev.so base2=base prev2=vispec_b200/lib/variants/prev.so
