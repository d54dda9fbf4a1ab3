#!/bin/bash
mkdir -p gpurun_out
T="--timeout 180 --timeout-method thread"
timeout 900 python -m pytest tests/test_gpu_program.py tests/test_gpu_e2e.py tests/test_gpu_qwen.py tests/test_gpu_draft.py -q -x $T > gpurun_out/t_u.log 2>&1; echo tests_rc=$?
tail -n 4 gpurun_out/t_u.log | cut -c1-300
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2u.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
PY
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py -q -x -k "attention_tma_staged and (4-2-61 or 4-4-33) and opts2 or folded and 512-4" > gpurun_out/san_r2u_racecheck.log 2>&1; echo racecheck_rc=$?
tail -n 3 gpurun_out/san_r2u_racecheck.log
export VS_LIB_PATH=$PWD/vispec_b200/lib/variants/clk.so
for sh in o down; do echo "== SHAPE=$sh"; SHAPE=$sh timeout 120 python tools/clk_epi.py 2>&1 | tail -n 16; done > gpurun_out/clk_epi_r2u.txt 2>&1
cat gpurun_out/clk_epi_r2u.txt
