#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
timeout 600 python tools/bench_slab_phases.py --shapes o,down,gateup,head,qkv --debug 0,3,4 --reps 8 > gpurun_out/slab_phases2.txt 2>&1; echo phases_rc=$?
cut -c1-150 gpurun_out/slab_phases2.txt
VS_SPIN_NS=20 timeout 300 python tools/bench_slab_phases.py --shapes o,gateup --debug 0 --reps 8 > gpurun_out/slab_phases2_spin20.txt 2>&1; cut -c1-150 gpurun_out/slab_phases2_spin20.txt
timeout 900 python -m pytest tests -q -m gpu -x $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -n 5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo bench_rc=$?
VS_LAYOUT=tiled timeout 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/bench_r2g_tiled.json 2> gpurun_out/bench_r2g_tiled.err; echo bench_tiled_rc=$?
timeout 300 python tools/trace_cycle.py --json gpurun_out/trace_r2g.json > gpurun_out/trace_r2g.txt 2>&1; echo trace_rc=$?
tail -n 22 gpurun_out/trace_r2g.txt | cut -c1-900
python - <<'PY'
import json
for f in ("bench_r2g","bench_r2g_tiled"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["roofline"]["frac"], d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
        print("   fixture", d.get("accept_fixture"), "qwen", (d.get("qwen2_5_vl_7b") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
