#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
timeout 900 python -m pytest tests -q -m gpu -x $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
tail -n 4 gpurun_out/t_all.log
timeout 700 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo bench_rc=$?
timeout 600 python bench.py --workload mixed --steps 1 --warmup 1 > gpurun_out/bench_r2i_mixed.json 2> gpurun_out/bench_r2i_mixed.err; echo mixed_rc=$?
timeout 600 python bench.py --workload llava13b_tp --steps 1 --warmup 1 --layers 8 > gpurun_out/bench_r2i_tp13b.json 2> gpurun_out/bench_r2i_tp13b.err; echo tp_rc=$?
python - <<'PY'
import json
for f in ("bench_r2i","bench_r2i_mixed","bench_r2i_tp13b"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
        for k in ("streams","accept_len_hist","prompt_pass","gpu_eager_baseline","cpu_baseline","accept_fixture"):
            if d.get(k): print("   ", k, json.dumps(d[k])[:600])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
