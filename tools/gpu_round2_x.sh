#!/bin/bash
mkdir -p gpurun_out
T="--timeout 180 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_e2e.py tests/test_gpu_qwen.py tests/test_gpu_draft.py tests/test_gpu_sampling.py -q -x $T > gpurun_out/t_x.log 2>&1; echo tests_rc=$?
tail -n 3 gpurun_out/t_x.log | cut -c1-300
timeout 300 python tools/bench_attention.py --kv 734 2048 8192 2>&1 | cut -c1-220
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2x.json 2> gpurun_out/bench_r2x.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2x.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
PY
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py -q -x -k "attention_tma_staged and (4-2-61 or 4-4-33) and opts2" > gpurun_out/san_r2x_racecheck.log 2>&1; echo racecheck_rc=$?
tail -n 2 gpurun_out/san_r2x_racecheck.log
