#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
( time timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err ) 2>&1 | grep real; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2_n2.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["n_gpus"], d["e2e"]["value"], d["scaling"], d["config"].get("parallelism"), list(d.keys()))
PY
( time timeout 900 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r2_ref_n2.json 2> gpurun_out/bench_r2_ref_n2.err ) 2>&1 | grep real; echo ref_rc=$?
tail -c 600 gpurun_out/bench_r2_ref_n2.json
