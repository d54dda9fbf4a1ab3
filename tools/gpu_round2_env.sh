#!/bin/bash
# re-tune the run-time switches on the final build: bench workload under VS_SPIN_NS / VS_ATTN_SPLITS / VS_EPI_NARROW
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-fixture > gpurun_out/env_$name.json 2> gpurun_out/env_$name.err
  python - "$name" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/env_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "tok/s %.1f" % d["value"], "frac %.4f" % d["roofline"]["frac"], {k: round(v["us"]) for k, v in d["phases_us_per_cycle"].items() if v["us"] > 50})
PY
}
run base1 VS_NOP=1
run spin50 VS_SPIN_NS=50
run spin200 VS_SPIN_NS=200
run splits9 VS_ATTN_SPLITS=9
run splits7 VS_ATTN_SPLITS=7
run narrow VS_EPI_NARROW=1
run base2 VS_NOP=1
