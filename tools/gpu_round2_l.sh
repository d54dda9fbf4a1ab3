#!/bin/bash
# baseline of the session: GPU tests, default bench, attention bench, one ncu capture of the streamed attention at 8192 keys
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
timeout 900 python -m pytest tests -q -m gpu -x $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -n 4 gpurun_out/t_all.log
timeout 300 python tools/bench_attention.py --kv 734 2048 8192 > gpurun_out/attn_base.jsonl 2>&1; cat gpurun_out/attn_base.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:program_kernel -s 3 -c 1 -o gpurun_out/prof_attn8192_base python tools/bench_attention.py --kv 8192 > gpurun_out/ncu_attn_base.log 2>&1; echo ncu_rc=$?
timeout 700 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo bench_rc=$?
python - <<'PY'
import json
for f in ("bench_r2l",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
        for k in ("gpu_eager_baseline","accept_fixture"):
            if d.get(k): print("   ", k, json.dumps(d[k])[:600])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
