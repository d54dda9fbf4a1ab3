#!/bin/bash
# compute-sanitizer over the phases touched in round 2 (memcheck: small shapes of every program test family + one e2e
# generation + the top-k acceptance; racecheck: folded epilogue, wide epilogue items, top-k histograms)
mkdir -p gpurun_out
SEL='linear_store and (5-1000-256 or 1-256-256) or swiglu or heads_rope and 4-2-7 or residual_in_place or chain_linear or rows_argmax or kv_splice or folded and 512-4 or attention_causal and 5-129 or attention_tree and 4-2-61'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py -q -x -k "$SEL" > gpurun_out/san_r2_memcheck_program.log 2>&1; echo memcheck_program_rc=$?
tail -n 2 gpurun_out/san_r2_memcheck_program.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_e2e.py -q -x -k "top_k or reference_api_cycle and bigram_noisy_img or crosses_the_register" > gpurun_out/san_r2_memcheck_e2e.log 2>&1; echo memcheck_e2e_rc=$?
tail -n 2 gpurun_out/san_r2_memcheck_e2e.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_program.py tests/test_gpu_sampling.py -q -x -k "folded and 512-4 or swiglu or top_k_matches and 5" > gpurun_out/san_r2_racecheck.log 2>&1; echo racecheck_rc=$?
tail -n 2 gpurun_out/san_r2_racecheck.log
