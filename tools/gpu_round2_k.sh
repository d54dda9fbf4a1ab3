#!/bin/bash
# 2-GPU smoke of the multi-GPU workloads (torchrun plumbing, NCCL, TP prompt pass)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 2 --workload llava13b_tp --steps 1 --warmup 1 --layers 8 > gpurun_out/bench_r2k_tp13b_n2.json 2> gpurun_out/bench_r2k_tp13b_n2.err; echo tp_rc=$?
timeout 900 $TR bench.py --gpus 2 --workload mixed --steps 1 --warmup 1 > gpurun_out/bench_r2k_mixed_n2.json 2> gpurun_out/bench_r2k_mixed_n2.err; echo mixed_rc=$?
python - <<'PY'
import json
for f in ("bench_r2k_tp13b_n2","bench_r2k_mixed_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],1), d["n_gpus"], d["e2e"]["value"])
        for k in ("prompt_pass","tokens_identical_on_all_ranks","accept_len_hist"):
            if k in d: print("   ", k, json.dumps(d[k])[:700])
        if "streams" in d: print("   streams", [(s["question_id"], s["rank"], round(s["tokens_per_s"])) for s in d["streams"]])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2500:])
PY
