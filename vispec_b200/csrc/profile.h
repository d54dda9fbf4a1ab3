// Launch counting and optional per-launch CUDA-event timing (used by bench.py's roofline leg only).
#pragma once
#include <cuda_runtime.h>

namespace vs {
enum ProfClass { PROF_GEMM = 0, PROF_ATTN = 1 };
void count_launch(int n = 1);
// When profiling is active, brackets the launches between begin/end with events on the launching stream.
void prof_begin(int cls, double bytes, cudaStream_t st);
void prof_end(cudaStream_t st);
}  // namespace vs
