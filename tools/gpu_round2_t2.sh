#!/bin/bash
# round-2 evidence run: bench record, launch list, ncu captures (cycle + attention alone), phase trace, sanitizer
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo bench_rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2j.csv python tools/profile_cycle.py --cycles 2 > /dev/null 2>&1; echo launches_rc=$?
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:program_kernel -c 1 -o gpurun_out/prof_cycle_r2j python tools/profile_cycle.py --cycles 2 > /dev/null 2>&1; echo ncu_cycle_rc=$?
for kv in 734 2048 8192; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:program_kernel -s 3 -c 1 -o gpurun_out/prof_attn${kv}_r2j python tools/bench_attention.py --kv $kv --layers 8 --reps 2 > /dev/null 2>&1; echo ncu_attn${kv}_rc=$?
done
timeout 300 python tools/trace_cycle.py --json gpurun_out/trace_r2j.json > gpurun_out/trace_r2j.txt 2>&1; echo trace_rc=$?
timeout 300 python tools/bench_attention.py --kv 734 2048 4096 8192 > gpurun_out/attn_r2j.jsonl 2>&1; echo attn_rc=$?
( export VS_LIB_PATH=$PWD/vispec_b200/lib/variants/clk.so
  for nc in 704 8000; do echo "== N_ATTN=6 N_CTX=$nc (warm, last of 6 phases)"; N_ATTN=6 N_CTX=$nc timeout 120 python tools/clk_attn.py 2>&1 | tail -n 15; done > gpurun_out/clk_attn_r2j.txt 2>&1
  for sh in o down; do echo "== SHAPE=$sh"; SHAPE=$sh timeout 120 python tools/clk_epi.py 2>&1 | tail -n 15; done > gpurun_out/clk_epi_r2j.txt 2>&1 )
SEL='attention_tma_staged and (32-32-734 or 28-4-700 or 4-4-33 or 8-8-5000) or attention_phases_back_to_back'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py -q -x -k "$SEL" > gpurun_out/san_attn_memcheck.log 2>&1; echo memcheck_rc=$?
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py -q -x -k "attention_tma_staged and (4-2-61 or 4-4-33) and opts2" > gpurun_out/san_attn_racecheck.log 2>&1; echo racecheck_rc=$?
tail -n 3 gpurun_out/san_attn_memcheck.log gpurun_out/san_attn_racecheck.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r2j.json").read().strip().splitlines()[-1])
print(round(d["value"],1), (d.get("roofline") or {}).get("frac"), d["e2e"]["value"], {k:v["us"] for k,v in (d.get("phases_us_per_cycle") or {}).items()})
print("   attn", d.get("roofline_attention")); print("   long", d.get("attention_long_context"))
print("   gpu_eager", d.get("gpu_eager_baseline")); print("   cpu", d.get("cpu_baseline"))
PY
cat gpurun_out/attn_r2j.jsonl | cut -c1-230
