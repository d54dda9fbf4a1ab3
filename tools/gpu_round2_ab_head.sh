#!/bin/bash
# A/B of the in-tree build against vispec_b200/lib/variants/head.so (tools/build_variant.py head on the previous commit):
# bench workload and attention alone, alternating
mkdir -p gpurun_out
bash tools/ab_bench.sh head1=vispec_b200/lib/variants/head.so new1=base head2=vispec_b200/lib/variants/head.so new2=base
for r in 1 2; do
  echo "== head"; VS_LIB_PATH=$PWD/vispec_b200/lib/variants/head.so timeout 300 python tools/bench_attention.py --kv 734 8192 2>&1 | cut -c1-120
  echo "== new"; timeout 300 python tools/bench_attention.py --kv 734 8192 2>&1 | cut -c1-120
done
