#!/bin/bash
mkdir -p gpurun_out
T="--timeout 180 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_e2e.py tests/test_gpu_qwen.py tests/test_gpu_draft.py -q -x $T > gpurun_out/t_v.log 2>&1; echo tests_rc=$?
tail -n 4 gpurun_out/t_v.log | cut -c1-300
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2v.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
PY
export VS_LIB_PATH=$PWD/vispec_b200/lib/variants/clk.so
for sh in o; do echo "== SHAPE=$sh"; SHAPE=$sh timeout 120 python tools/clk_epi.py 2>&1 | tail -n 14; done > gpurun_out/clk_epi_r2v.txt 2>&1
cat gpurun_out/clk_epi_r2v.txt
