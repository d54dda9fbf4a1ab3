#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_report.jsonl
T="--timeout 240 --timeout-method thread"
timeout 900 python -m pytest tests -q -m gpu -x $T > gpurun_out/t_all.log 2>&1; echo all_rc=$?
grep -E "passed|failed" gpurun_out/t_all.log | tail -1
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; echo bench_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2m.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
print("   attn", d["roofline_attention"]["avg_layer_us"], "long", [(x["kv_len"], x["us_per_layer"], x["frac_of_hbm_peak"]) for x in d["attention_long_context"]])
print("   gpu_eager", d["gpu_eager_baseline"]["value"], d["gpu_eager_baseline"]["speedup_decode"], "cpu", d["cpu_baseline"]["value"], "fixture", d["accept_fixture"]["value"], "qwen", d["qwen2_5_vl_7b"]["value"], d["qwen2_5_vl_7b"]["hbm_frac"])
PY
