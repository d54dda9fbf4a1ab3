#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
timeout 900 python -m pytest tests -q -m gpu -x $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -n 6 gpurun_out/t_all.log | cut -c1-300
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; echo bench_rc=$?
python - <<'PY'
import json
for f in ("bench_r2q",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
        print("   attn", d.get("roofline_attention")); print("   long", d.get("attention_long_context"))
        print("   fixture", d.get("accept_fixture"), "qwen", (d.get("qwen2_5_vl_7b") or {}))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
