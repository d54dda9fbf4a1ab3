// Row-wise kernels of the decode path: RMSNorm, row gathers / embedding lookup, arg-max over the vocabulary,
// and the draft head's bf16 log-softmax + top-k.
// Reference call sites: LlamaRMSNorm (vispec/model/cnets_ours.py:513-527, modeling_llama_kv.py:105-134),
// embed_tokens (cnets_ours.py:853,1066-1082), LogSoftmax + topk (cnets_ours.py:1111-1115,1145-1148),
// argmax over target logits (vispec/model/utils.py:290,438-440,554).
#include <math.h>

#include "kernels.h"
#include "profile.h"

namespace vs {

namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += sh[i];     // fixed order -> deterministic
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  v = warp_max(v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int i = 0; i < nw; ++i) r = fmaxf(r, sh[i]);
  return r;
}
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* sh) {
  v = warp_max_u64(v);
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  unsigned long long r = 0ull;
  for (int i = 0; i < nw; ++i) r = sh[i] > r ? sh[i] : r;
  return r;
}

// y = weight * bf16(x * rsqrt(mean(x^2) + eps))   (fp32 statistics, two bf16 roundings)
__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, long ldx, const bf16* __restrict__ w,
                                                      bf16* __restrict__ y, long ldy, int H, float eps) {
  __shared__ float sh[8];
  const bf16* xr = x + (long)blockIdx.x * ldx;
  bf16* yr = y + (long)blockIdx.x * ldy;
  float ss = 0.f;
  for (int i = threadIdx.x; i < H; i += 256) {
    float v = bf2f(xr[i]);
    ss += v * v;
  }
  ss = block_sum(ss, sh);
  float rstd = 1.0f / sqrtf(ss / (float)H + eps);
  for (int i = threadIdx.x; i < H; i += 256) {
    float n = rbf(bf2f(xr[i]) * rstd);
    yr[i] = f2bf(bf2f(w[i]) * n);
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const bf16* __restrict__ src, long lds,
                                                          const int* __restrict__ idx, bf16* __restrict__ dst, long ldd,
                                                          int H, int n_src_rows) {
  int t = blockIdx.x;
  int r = idx ? idx[t] : t;
  if (n_src_rows > 0) r = min(max(r, 0), n_src_rows - 1);
  const uint4* s = reinterpret_cast<const uint4*>(src + (long)r * lds);
  uint4* d = reinterpret_cast<uint4*>(dst + (long)t * ldd);
  for (int i = threadIdx.x; i < H / 8; i += 256) d[i] = s[i];
}

__global__ void __launch_bounds__(256) broadcast_row_kernel(const bf16* __restrict__ row, bf16* __restrict__ dst,
                                                            long ldd, int H) {
  const uint4* s = reinterpret_cast<const uint4*>(row);
  uint4* d = reinterpret_cast<uint4*>(dst + (long)blockIdx.x * ldd);
  for (int i = threadIdx.x; i < H / 8; i += 256) d[i] = s[i];
}

__global__ void __launch_bounds__(1024) argmax_rows_kernel(const bf16* __restrict__ x, long ldx, int V, int* out) {
  __shared__ unsigned long long sh[32];
  const bf16* xr = x + (long)blockIdx.x * ldx;
  unsigned long long best = 0ull;
  for (int i = threadIdx.x; i < V; i += 1024) {
    unsigned long long k = rank_key(bf2f(xr[i]), (uint32_t)i);
    best = k > best ? k : best;
  }
  best = block_max_u64(best, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = (int)key_index(best);
}

// One block per row.  lp = bf16((x - max) - log(sum exp(x - max))) in fp32 math, then the k best by
// (value desc, index asc).  Outputs sorted best-first, like torch.topk(sorted=True).
constexpr int TOPK_MAX = 8;
__global__ void __launch_bounds__(1024) logsoftmax_topk_kernel(const bf16* __restrict__ logits, long ldl, int V, int k,
                                                               const int* __restrict__ row_sel, float* top_val,
                                                               int* top_idx) {
  __shared__ float shf[32];
  __shared__ unsigned long long shk[32];
  const int row = row_sel ? *row_sel : blockIdx.x;
  const bf16* xr = logits + (long)row * ldl;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 1024) m = fmaxf(m, bf2f(xr[i]));
  m = block_max(m, shf);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 1024) s += expf(bf2f(xr[i]) - m);
  s = block_sum(s, shf);
  const float lse = logf(s);

  unsigned long long loc[TOPK_MAX];
#pragma unroll
  for (int j = 0; j < TOPK_MAX; ++j) loc[j] = 0ull;
  for (int i = threadIdx.x; i < V; i += 1024) {
    float lp = rbf((bf2f(xr[i]) - m) - lse);
    unsigned long long key = rank_key(lp, (uint32_t)i);
    if (key > loc[TOPK_MAX - 1]) {
      loc[TOPK_MAX - 1] = key;
#pragma unroll
      for (int j = TOPK_MAX - 1; j > 0; --j) {
        if (loc[j] > loc[j - 1]) {
          unsigned long long tmp = loc[j];
          loc[j] = loc[j - 1];
          loc[j - 1] = tmp;
        }
      }
    }
  }
  for (int r = 0; r < k; ++r) {
    unsigned long long best = block_max_u64(loc[0], shk);
    if (loc[0] == best && best != 0ull) {     // keys are unique (index in the low word): exactly one owner pops
#pragma unroll
      for (int j = 0; j < TOPK_MAX - 1; ++j) loc[j] = loc[j + 1];
      loc[TOPK_MAX - 1] = 0ull;
    }
    if (threadIdx.x == 0) {
      top_val[blockIdx.x * k + r] = key_value(best);
      top_idx[blockIdx.x * k + r] = (int)key_index(best);
    }
  }
}

}  // namespace

cudaError_t launch_rmsnorm(const bf16* x, long ldx, const bf16* w, bf16* y, long ldy, int T, int H, float eps,
                           cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  rmsnorm_kernel<<<T, 256, 0, st>>>(x, ldx, w, y, ldy, H, eps);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const bf16* src, long lds, const int* idx, bf16* dst, long ldd, int T, int H,
                               cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  if (H % 8 || lds % 8 || ldd % 8) return cudaErrorInvalidValue;
  gather_rows_kernel<<<T, 256, 0, st>>>(src, lds, idx, dst, ldd, H, 0);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_embed_rows(const bf16* table, long ldt, const int* ids, bf16* dst, long ldd, int T, int H,
                              int vocab, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  if (H % 8 || ldt % 8 || ldd % 8) return cudaErrorInvalidValue;
  gather_rows_kernel<<<T, 256, 0, st>>>(table, ldt, ids, dst, ldd, H, vocab);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_broadcast_row(const bf16* row, bf16* dst, long ldd, int T, int H, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  if (H % 8 || ldd % 8) return cudaErrorInvalidValue;
  broadcast_row_kernel<<<T, 256, 0, st>>>(row, dst, ldd, H);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_argmax_rows(const bf16* x, long ldx, int T, int V, int* out, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  argmax_rows_kernel<<<T, 1024, 0, st>>>(x, ldx, V, out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_logsoftmax_topk(const bf16* logits, long ldl, int T, int V, int k, const int* row_sel,
                                   float* top_val, int* top_idx, cudaStream_t st) {
  if (T <= 0) return cudaSuccess;
  if (k < 1 || k > TOPK_MAX) return cudaErrorInvalidValue;
  logsoftmax_topk_kernel<<<T, 1024, 0, st>>>(logits, ldl, V, k, row_sel, top_val, top_idx);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vs
