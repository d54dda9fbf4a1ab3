#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 8 --workload mixed --steps 1 --warmup 1 > gpurun_out/bench_r2z_mixed_n8.json 2> gpurun_out/bench_r2z_mixed_n8.err; echo mixed_rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2z_mixed_n8.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["n_gpus"], d["e2e"]["value"], d.get("accept_len_hist"))
print([(s["question_id"], s["rank"], s.get("text_tokens"), round(s["tokens_per_s"])) for s in d["streams"]])
PY
