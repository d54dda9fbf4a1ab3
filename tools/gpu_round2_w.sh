#!/bin/bash
# streamed attention change: attention tests, racecheck on a streamed shape, attention alone, bench
mkdir -p gpurun_out
T="--timeout 180 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_e2e.py -q -x $T > gpurun_out/t_w.log 2>&1; echo tests_rc=$?
tail -n 3 gpurun_out/t_w.log | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py -q -x -k "attention_tma_staged and 8-8-5000 and opts2" > gpurun_out/san_w_racecheck.log 2>&1; echo racecheck_rc=$?
tail -n 2 gpurun_out/san_w_racecheck.log
timeout 300 python tools/bench_attention.py --kv 734 2048 4096 8192 2>&1 | cut -c1-220
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2w.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
PY
