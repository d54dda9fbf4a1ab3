#!/bin/bash
# GPU call: packed slab variants (address-swizzled / block-relative), M=64 A/B, full suite, bench + per-phase trace
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
slab_tests() { timeout 900 python -m pytest tests/test_gpu_program.py -q -m gpu -k "slab" $T; }
MODE=0
for m in 1 2; do
  VS_SLAB_PACKED=$m slab_tests > gpurun_out/t_slab_packed$m.log 2>&1; rc=$?; echo "packed$m rc=$rc $(tail -n 1 gpurun_out/t_slab_packed$m.log)"
  if [ $rc -eq 0 ] && [ $MODE -eq 0 ]; then MODE=$m; fi
done
VS_SLAB_PACKED=$MODE VS_SLAB_M64=0 slab_tests > gpurun_out/t_slab_m128.log 2>&1; echo "m128 rc=$? $(tail -n 1 gpurun_out/t_slab_m128.log)"
export VS_SLAB_PACKED=$MODE; echo "USING VS_SLAB_PACKED=$MODE"
timeout 900 python -m pytest tests -q -m gpu $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -n 12 gpurun_out/t_all.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo bench_rc=$?
VS_SLAB_M64=0 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2d_m128.json 2> gpurun_out/bench_r2d_m128.err; echo bench_m128_rc=$?
timeout 300 python tools/trace_cycle.py --json gpurun_out/trace_r2d.json > gpurun_out/trace_r2d.txt 2>&1; echo trace_rc=$?
cat gpurun_out/trace_r2d.txt | tail -n 40
python - <<'PY'
import json
for f in ("bench_r2d","bench_r2d_m128"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
    except Exception as e: print(f, "ERR", e)
PY
