#!/bin/bash
# GPU call 2: slab GEMM bring-up (packed and padded variants), full suite, bench A/B against the round-1 tiled layout
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
slab_tests() { timeout 900 python -m pytest tests/test_gpu_program.py -q -m gpu -k "slab" $T; }
slab_tests > gpurun_out/t_slab_packed.log 2>&1; rc1=$?; echo packed_rc=$rc1
VS_SLAB_PACKED=0 slab_tests > gpurun_out/t_slab_padded.log 2>&1; rc2=$?; echo padded_rc=$rc2
VS_SLAB_PACKED=0 VS_X3D=0 slab_tests > gpurun_out/t_slab_padded_2d.log 2>&1; echo padded2d_rc=$?
tail -3 gpurun_out/t_slab_packed.log gpurun_out/t_slab_padded.log gpurun_out/t_slab_padded_2d.log
if [ $rc1 -ne 0 ] && [ $rc2 -eq 0 ]; then export VS_SLAB_PACKED=0; echo "USING PADDED"; fi
if [ $rc1 -ne 0 ] && [ $rc2 -ne 0 ]; then export VS_LAYOUT=tiled; echo "USING TILED"; fi
timeout 900 python -m pytest tests -q -m gpu -x $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo bench_rc=$?
VS_LAYOUT=tiled timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2b_tiled.json 2> gpurun_out/bench_r2b_tiled.err; echo bench_tiled_rc=$?
VS_SLAB_PACKED=0 timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2b_padded.json 2> gpurun_out/bench_r2b_padded.err; echo bench_padded_rc=$?
python - <<'PY'
import json
for f in ("bench_r2b","bench_r2b_tiled","bench_r2b_padded"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
    except Exception as e: print(f, "ERR", e)
PY
