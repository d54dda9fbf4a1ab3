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
constexpr int CT_LD = BM + 4;                 // fp32 tile [32][132] staged in the pipeline smem afterwards
constexpr int THREADS = 256;

struct Tile {
  const float* ct;   // [32][CT_LD] fp32, ct[t*CT_LD + r] = acc(token t, tile row r)
  int tile, n0, t0, tb_rows;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

template <int EPI>
__device__ __forceinline__ void epilogue(const GemmArgs& p, const Tile& tl) {
  const int tid = threadIdx.x;
  if (EPI == EPI_STORE || EPI == EPI_RESID) {
    for (int i = tid; i < VS_TOK * BM; i += THREADS) {
      int t = i >> 7, r = i & 127;
      int n = tl.n0 + r;
      if (t >= tl.tb_rows || n >= p.N) continue;
      float v = tl.ct[t * CT_LD + r];
      if (p.bias) v += bf2f(p.bias[n]);
      long row = tl.t0 + t;
      if (EPI == EPI_RESID) {
        float o = rbf(v);                                         // Linear output rounds to bf16 ...
        v = bf2f(p.R[row * p.ldr + n]) + o;                       // ... then the residual add rounds again
      }
      p.Y[row * p.ldy + n] = f2bf(v);
    }
  } else if (EPI == EPI_SWIGLU) {
    // tile rows [0,64) = gate features j0..j0+63, rows [64,128) = the matching up features
    const int j0 = tl.tile * 64;
    for (int i = tid; i < VS_TOK * 64; i += THREADS) {
      int t = i >> 6, r = i & 63;
      if (t >= tl.tb_rows) continue;
      float g = rbf(tl.ct[t * CT_LD + r]);
      float u = rbf(tl.ct[t * CT_LD + 64 + r]);
      float a = rbf(silu_f(g));                                   // act_fn(gate) rounds, then * up rounds
      p.Y[(long)(tl.t0 + t) * p.ldy + j0 + r] = f2bf(a * u);
    }
  } else {  // EPI_HEADS
    const int tile = tl.tile;
    const bool is_q = tile < p.n_q_tiles;
    const bool is_k = !is_q && tile < p.n_q_tiles + p.n_k_tiles;
    const int head = is_q ? tile : (is_k ? tile - p.n_q_tiles : tile - p.n_q_tiles - p.n_k_tiles);
    const int row_base = (p.kv_row_base ? *p.kv_row_base : 0) + p.kv_row_off;
    for (int i = tid; i < VS_TOK * BM; i += THREADS) {
      int t = i >> 7, d = i & 127;
      if (t >= tl.tb_rows) continue;
      int n = tl.n0 + d;
      float b = p.bias ? bf2f(p.bias[n]) : 0.f;
      float x = rbf(tl.ct[t * CT_LD + d] + b);
      float out = x;
      if (p.rope && !(!is_q && !is_k)) {
        int d2 = d < 64 ? d + 64 : d - 64;
        float b2 = p.bias ? bf2f(p.bias[tl.n0 + d2]) : 0.f;
        float x2 = rbf(tl.ct[t * CT_LD + d2] + b2);
        float rot = d < 64 ? -x2 : x2;
        long pos = p.pos_ids[tl.t0 + t];
        float c = bf2f(p.cos[pos * 128 + d]), s = bf2f(p.sin[pos * 128 + d]);
        out = rbf(x * c) + rbf(rot * s);                          // three bf16 roundings (cnets_ours.py:117-118)
      }
      bf16 o = f2bf(out);
      if (is_q) {
        p.Qout[(long)head * p.q_head_stride + (long)(tl.t0 + t) * 128 + d] = o;
      } else {
        bf16* dst = is_k ? p.Kout : p.Vout;
        dst[(long)head * p.kv_head_stride + (long)(row_base + tl.t0 + t) * 128 + d] = o;
      }
    }
  }
}

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
  epilogue<EPI>(p, tl);
}

}  // namespace

int gemm_pick_splits(int N, int K, int T) {
  int n_tiles = (N + BM - 1) / BM;
  int tok_blocks = (T + VS_TOK - 1) / VS_TOK;
  int nk = K / BK;
  long blocks = (long)n_tiles * tok_blocks;
  int target = 148 * 2;
  int s = (int)((target + blocks - 1) / blocks);
  if (s < 1) s = 1;
  int max_s = nk / 8;               // keep >= 8 k-chunks per block so the pipeline has something to hide
  if (max_s < 1) max_s = 1;
  if (s > max_s) s = max_s;
  if (s > 16) s = 16;
  return s;
}

size_t gemm_partials_floats(int N, int T, int splits) {
  if (splits <= 1) return 0;
  size_t n_tiles = (N + BM - 1) / BM;
  size_t tok_blocks = (T + VS_TOK - 1) / VS_TOK;
  return tok_blocks * splits * n_tiles * VS_TOK * BM;
}

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

cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.T <= 0) return cudaSuccess;
  if (a.K % BK != 0 || a.ldw % 8 != 0 || a.ldx % 8 != 0) return cudaErrorInvalidValue;
  if (a.splits > 1 && (!a.partials || !a.counters)) return cudaErrorInvalidValue;
  switch (a.epi) {
    case EPI_STORE: return launch_t<EPI_STORE>(a, st);
    case EPI_RESID: return launch_t<EPI_RESID>(a, st);
    case EPI_SWIGLU: return launch_t<EPI_SWIGLU>(a, st);
    case EPI_HEADS: return launch_t<EPI_HEADS>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vs
