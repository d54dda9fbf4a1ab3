// Decode attention over a head-major KV slab with a fully visible prefix plus a bit-masked (tree) or
// causal tail.  One implementation serves
//   * the target's tree-attention verify  (vispec/model/modeling_llama_kv.py:527-653, mask :890-924),
//   * the draft layer's cache + SDPA        (vispec/model/cnets_ours.py:393-445, mask :781-815),
//   * the vision adaptor's cross-attention  (vispec/model/cnets_ours.py:630-661).
// The reference materialises a [q, S] fp32 additive mask with a boolean index_put (host sync) and,
// for the target, runs two batched matmuls and a softmax as separate kernels.  Here the mask is
// a 64-bit word per query, K and V are each read exactly once:
//   pass 1 (attn_scores): S = Q K^T per 64-key chunk on tensor cores, masked, scaled; per-chunk max / sum
//   pass 2 (attn_pv)    : global max / sum from the chunk stats, P = softmax in the reference's rounding,
//                          O += P V on tensor cores, deterministic last-block reduction over the KV splits.
// The "eager" flag reproduces the target's rounding points (bf16 QK^T, bf16 divide by sqrt(d), fp32 softmax
// cast to bf16, PV); otherwise scores stay fp32 and normalisation happens after PV (SDPA-style).
#include <math.h>

#include "attn_dev.cuh"
#include "kernels.h"
#include "profile.h"

namespace vs {

namespace {

struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

// Stand-alone launches: one thread block == one team (attn_dev.cuh); the persistent program kernel (program.cu)
// runs the same bodies as phases of one launch.
__global__ void __launch_bounds__(ATT_THREADS) attn_scores_kernel(const AttnArgs p) {
  __shared__ AttnSmem sm;
  const int n_ctx = p.n_ctx ? *p.n_ctx : p.n_ctx_host;
  int n_chunks = (n_ctx + p.n_new + ATT_CHUNK - 1) / ATT_CHUNK;
  // causal: keys behind the last row of this query block are invisible to all of it (prompt passes: half of the chunks)
  if (p.causal) n_chunks = min(n_chunks, (n_ctx + min(p.Tq, ((int)blockIdx.z + 1) * VS_TOK) + ATT_CHUNK - 1) / ATT_CHUNK);
  bool first = true;
  for (int c = blockIdx.y; c < n_chunks; c += gridDim.y) {
    attn_scores_item(p, sm, blockIdx.x, blockIdx.z, gridDim.z, c, first, threadIdx.x, BlockSync());
    first = false;
  }
}

__global__ void __launch_bounds__(ATT_THREADS) attn_pv_kernel(const AttnArgs p) {
  __shared__ AttnSmem sm;
  attn_pv_item(p, sm, blockIdx.x, blockIdx.z, gridDim.z, blockIdx.y, gridDim.y, threadIdx.x, BlockSync());
}

}  // namespace

cudaError_t launch_attention(const AttnArgs& a, cudaStream_t st) {
  if (a.Tq <= 0) return cudaSuccess;
  if (!a.scores || !a.stats || !a.opart || !a.counters || a.pv_splits < 1 || a.max_chunks < 1)
    return cudaErrorInvalidValue;
  int tq_blocks = (a.Tq + VS_TOK - 1) / VS_TOK;
  int gy = a.pv_splits;
  // algorithmic bytes: K and V read once = 2 * kv_len * n_kv * 128 * 2 B; kv_len is device-side, so the
  // profiler records the per-key figure and bench.py multiplies by the kv length it knows.
  prof_begin(PROF_ATTN, 2.0 * a.n_kv_heads * 128 * 2, st);
  attn_scores_kernel<<<dim3(a.n_heads, gy, tq_blocks), ATT_THREADS, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  attn_pv_kernel<<<dim3(a.n_heads, a.pv_splits, tq_blocks), ATT_THREADS, 0, st>>>(a);
  prof_end(st);
  count_launch(2);
  return cudaGetLastError();
}

}  // namespace vs
