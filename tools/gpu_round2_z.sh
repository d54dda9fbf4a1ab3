#!/bin/bash
# TP prompt pass at world 2 with the per-step timing breakdown (VS_TP_TIMING=1)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --workload llava13b_tp --steps 3 --warmup 1 --layers 8 > gpurun_out/bench_r2z_tp13b_n2.json 2> gpurun_out/bench_r2z_tp13b_n2.err; echo tp_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2z_tp13b_n2.json").read().strip().splitlines()[-1])
for r in d["prompt_pass"]: print(json.dumps(r))
PY
tail -3 gpurun_out/bench_r2z_tp13b_n2.err
