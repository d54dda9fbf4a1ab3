#!/bin/bash
mkdir -p gpurun_out
T="--timeout 120 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_e2e.py -q -x $T > gpurun_out/t_attn.log 2>&1; echo attn_rc=$?
tail -n 5 gpurun_out/t_attn.log | cut -c1-300
timeout 300 python tools/bench_attention.py --kv 734 2048 8192 2>&1 | cut -c1-220
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2s.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
print("   attn", d.get("roofline_attention")["avg_layer_us"])
PY
