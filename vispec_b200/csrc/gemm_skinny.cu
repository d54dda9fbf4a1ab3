// Weight-streaming "skinny" GEMM for the decode path:  Y[T<=32 per block, N] = X[T,K] * W[N,K]^T.
//
// Replaces every nn.Linear on the draft-and-verify path of the reference
// (vispec/model/cnets_ours.py:363-365,461,508,921-922,982-988,1111; vispec/model/modeling_llama_kv.py:421,
// 567-569,648,1196), where the reference issues one cuBLAS call + several elementwise kernels per
// projection.  The op is HBM-bound (M<=30 tokens against 33-263 MB of weights), so the design is a
// weight stream: W is the M operand of the tensor-core MMA (128 rows per block), the <=32 tokens are
// the N operand, a 4-stage cp.async pipeline keeps 80 KB per block in flight, split-K fills the 148
// SMs, and a deterministic last-block reduction feeds a fused epilogue that reproduces the
// reference's bf16 rounding points (bias, residual add, SiLU*up, RoPE + in-place KV append).
#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "profile.h"

namespace vs {

namespace {

constexpr int BM = 128;                       // weight rows (output features) per block
constexpr int BK = 64;                        // k elements per stage (128 B per row)
constexpr int STAGES = 4;
constexpr int W_STAGE = BM * BK * 2;          // 16384
constexpr int X_STAGE = VS_TOK * BK * 2;      // 4096
constexpr int STAGE_BYTES = W_STAGE + X_STAGE;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES;   // 81920
constexpr int CT_LD = EPI_CT_LD;              // fp32 tile [32][132] staged in the pipeline smem afterwards
constexpr int THREADS = 256;

template <int EPI>
__global__ void __launch_bounds__(THREADS, 2) gemm_skinny_kernel(const GemmArgs p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_last;
  const int tile = blockIdx.x, split = blockIdx.y, tb = blockIdx.z;
  const int n_tiles = gridDim.x;
  const int n0 = tile * BM;
  const int t0 = tb * VS_TOK;
  const int tb_rows = min(VS_TOK, p.T - t0);
  const int nk = p.K / BK;
  const int kb = (int)(((long)nk * split) / p.splits);
  const int ke = (int)(((long)nk * (split + 1)) / p.splits);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int c = tid & 7;          // 16-byte chunk within the 128-byte row segment
  const int r0 = tid >> 3;        // 0..31
  auto load_stage = [&](int stage, int kc) {
    uint8_t* ws = smem + stage * STAGE_BYTES;
    uint8_t* xs = ws + W_STAGE;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = r0 + 32 * j;
      int n = n0 + r;
      bool ok = n < p.N;
      const bf16* src = p.W + (long)(ok ? n : 0) * p.ldw + (long)kc * BK + c * 8;
      cp_async16(ws + r * 128 + ((c ^ (r & 7)) << 4), src, ok);
    }
    {
      bool ok = r0 < tb_rows;
      const bf16* src = p.X + (long)(t0 + (ok ? r0 : 0)) * p.ldx + (long)kc * BK + c * 8;
      cp_async16(xs + r0 * 128 + ((c ^ (r0 & 7)) << 4), src, ok);
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (kb + s < ke) load_stage(s, kb + s);
    cp_async_commit();
  }
  for (int kc = kb; kc < ke; ++kc) {
    const int it = kc - kb;
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kc + STAGES - 1;
      if (nxt < ke) load_stage((it + STAGES - 1) % STAGES, nxt);
      cp_async_commit();
    }
    const uint8_t* ws = smem + (it % STAGES) * STAGE_BYTES;
    const uint8_t* xs = ws + W_STAGE;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      {
        int mat = lane >> 3;
        int r = warp * 16 + (lane & 7) + (mat & 1) * 8;
        int ch = kk * 2 + (mat >> 1);
        ldmatrix_x4(a, ws + r * 128 + ((ch ^ (r & 7)) << 4));
      }
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        uint32_t b[4];
        int mat = lane >> 3;
        int r = pr * 16 + (lane & 7) + (mat >> 1) * 8;
        int ch = kk * 2 + (mat & 1);
        ldmatrix_x4(b, xs + r * 128 + ((ch ^ (r & 7)) << 4));
        mma_bf16(acc[pr * 2 + 0], a, b[0], b[1]);
        mma_bf16(acc[pr * 2 + 1], a, b[2], b[3]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // accumulators -> fp32 tile in smem: ct[token][tile row]
  float* ct = reinterpret_cast<float*>(smem);
  {
    int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      int tok = grp * 8 + 2 * t;
      int row = warp * 16 + g;
      ct[tok * CT_LD + row] = acc[grp][0];
      ct[(tok + 1) * CT_LD + row] = acc[grp][1];
      ct[tok * CT_LD + row + 8] = acc[grp][2];
      ct[(tok + 1) * CT_LD + row + 8] = acc[grp][3];
    }
  }
  __syncthreads();

  if (p.splits > 1) {
    // deterministic split-K: every block publishes its partial tile; the last one to arrive sums them
    // in split order and runs the epilogue.
    const long tile_floats = (long)VS_TOK * BM;
    float* base = p.partials + ((long)tb * p.splits * n_tiles) * tile_floats;
    float* mine = base + ((long)split * n_tiles + tile) * tile_floats;
    for (int i = tid; i < VS_TOK * BM; i += THREADS) mine[i] = ct[(i >> 7) * CT_LD + (i & 127)];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      int old = atomicAdd(&p.counters[tb * n_tiles + tile], 1);
      s_last = (old == p.splits - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < VS_TOK * BM; i += THREADS) {
      float v = 0.f;
      for (int s = 0; s < p.splits; ++s) v += __ldcg(base + ((long)s * n_tiles + tile) * tile_floats + i);
      ct[(i >> 7) * CT_LD + (i & 127)] = v;
    }
    if (tid == 0) p.counters[tb * n_tiles + tile] = 0;
    __syncthreads();
  }

  Tile tl{ct, tile, n0, t0, tb_rows};
  epilogue<EPI, THREADS>(p, tl, threadIdx.x);
}

}  // namespace

template <int EPI>
static cudaError_t launch_t(const GemmArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_skinny_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid((a.N + BM - 1) / BM, a.splits, (a.T + VS_TOK - 1) / VS_TOK);
  prof_begin(PROF_GEMM, 2.0 * a.N * a.K * grid.z, st);        // algorithmic bytes: the weight matrix, once per token block
  gemm_skinny_kernel<EPI><<<grid, THREADS, SMEM_BYTES, st>>>(a);
  prof_end(st);
  count_launch();
  return cudaGetLastError();
}

// mma.sync bring-up kernel, kept selectable with VS_GEMM_LEGACY=1 for A/B debugging of the tcgen05 path
cudaError_t launch_gemm_legacy(const GemmArgs& a, cudaStream_t st) {
  switch (a.epi) {
    case EPI_STORE: return launch_t<EPI_STORE>(a, st);
    case EPI_RESID: return launch_t<EPI_RESID>(a, st);
    case EPI_SWIGLU: return launch_t<EPI_SWIGLU>(a, st);
    case EPI_HEADS: return launch_t<EPI_HEADS>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vs
