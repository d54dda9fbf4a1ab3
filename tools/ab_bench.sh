#!/bin/bash
# A/B of alternative builds of the library on the bench workload: tools/ab_bench.sh name=path.so ...  ("base" = in-tree build)
for spec in "$@"; do
  name=${spec%%=*}; path=${spec#*=}
  if [ "$path" = "base" ]; then unset VS_LIB_PATH; else export VS_LIB_PATH=$PWD/$path; fi
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || tail -c 800 gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{name}.json").read().strip().splitlines()[-1])
    ph = d["phases_us_per_cycle"]
    print(name, "tok/s %.1f" % d["value"], "frac %.4f" % d["roofline"]["frac"], "qwen", (d.get("qwen2_5_vl_7b") or {}).get("value"),
          {k: round(v["us"]) for k, v in ph.items() if v["us"] > 50})
except Exception as e:
    print(name, "FAILED", e)
PY
done
