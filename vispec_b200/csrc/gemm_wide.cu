// Prompt-pass GEMM on the 5th-generation tensor cores:  Y[T, N] = X[T, K] * W[N, K]^T for hundreds of token rows
// (target / draft prefill: modeling_llama_kv.py:567-569,648,421; cnets_ours.py:640-641,921-922).
//
// Same contract and fused epilogues as gemm_tc.cu, which is built for decode (<= 32 rows: a pure weight stream with
// MMA N = 32).  With T = 704 that kernel runs the tensor pipe at N = 32 (1/8 of its width) and re-reads every weight tile
// once per 32-row block: 0.75 ms for the 7B q|k|v projection (94 TFLOP/s).  Here the token dimension is the wide one:
//   * work unit = (128-row weight tile, 256-token block): tcgen05.mma cta_group::1 kind::f16 M=128 x N=256 x K=16,
//     fp32 accumulators double-buffered in TMEM (2 x 256 columns = all 512), so the epilogue of one unit overlaps the
//     MMAs of the next;
//   * warp 0 = TMA producer: W tile [128 x 64] + X tile [256 x 64] per stage (48 KB), 4 stages, 128-byte swizzle; rows
//     past T are zero-filled by the tensor map;
//   * units are dealt round-robin with the token block as the fastest index, so the CTAs running together share a weight
//     tile through L2 and every weight byte leaves HBM once;
//   * warps 4-11 = epilogue, two groups of four warps (one warp per TMEM lane quarter): the 256 accumulator columns are
//     drained in eight 32-token slices, even slices by group 0 and odd ones by group 1, each through its own
//     shared-memory tile and the same `epilogue<>` bodies as the decode kernel (bias / residual / SiLU*up / RoPE + KV
//     append), so both kernels round identically.  (One group alone needs ~40 us per unit against 17 us of MMAs.)
// No split-K: K is covered by one unit.  The token block width is chosen per launch (a multiple of 16 up to 256) so that
// the unit count fills the SMs: 704 rows x 32 weight tiles run as 4 blocks of 176 tokens (128 units) rather than 3 blocks
// of 256 (96 units on 148 SMs).
#include <cuda.h>

#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "profile.h"
#include "sm100.cuh"

namespace vs {

bool gemm_get_map(const void* ptr, long ld, int rows, int cols, int box_rows, CUtensorMap* out);   // gemm_tc.cu

namespace {

constexpr int BM = 128, BK = 64, TW = 256, STAGES = 4;
constexpr int W_STAGE = BM * BK * 2;            // 16384
constexpr int X_STAGE = TW * BK * 2;            // 32768
constexpr int STAGE_BYTES = W_STAGE + X_STAGE;  // 49152 (multiple of 1024)
constexpr int CT_BYTES = VS_TOK * EPI_CT_LD * 4;
constexpr int EPI_GROUPS = 2;
constexpr int BAR_OFF = STAGES * STAGE_BYTES + EPI_GROUPS * CT_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
constexpr int THREADS = 128 + EPI_GROUPS * 128;
constexpr int EPI_THREADS = 128;
constexpr int TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N = tw (multiple of 16), M=128
__host__ __device__ constexpr uint32_t idesc_wide(int tw) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)tw >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void epi_bar(int group) { asm volatile("bar.sync %0, 128;" ::"r"(1 + group) : "memory"); }
__device__ __forceinline__ void tc_mma_wide(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_wide_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const GemmArgs p,
                 const int n_tb, const int n_units, const int tw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
  float* ct = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.K / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], EPI_GROUPS * EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int tb = u % n_tb, tile = u / n_tb;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], W_STAGE + tw * BK * 2);
          uint8_t* dst = smem + stage * STAGE_BYTES;
          tma_load_2d(&tmW, &full[stage], dst, kc * BK, tile * BM, HINT_EVICT_LAST);
          tma_load_2d(&tmX, &full[stage], dst + W_STAGE, kc * BK, tb * tw, HINT_EVICT_LAST);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint32_t idesc = idesc_wide(tw);
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TW;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_desc(smem + stage * STAGE_BYTES);
          const uint64_t db = make_desc(smem + stage * STAGE_BYTES + W_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) tc_mma_wide(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc > 0 || k > 0) ? 1u : 0u);
          tc_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (2 groups x 4 warps, TMEM lanes 32*(warp%4)...) =====================
    const int group = (warp - 4) >> 2;
    const int et = (threadIdx.x - 128) & 127;  // 0..127 == tile row
    const int lane_base = (warp & 3) * 32;
    float* ctg = ct + group * (VS_TOK * EPI_CT_LD);
    int acc = 0, acc_phase = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int tb = u % n_tb, tile = u / n_tb;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int rows_left = min(p.T - tb * tw, tw);   // > 0
#pragma unroll 1
      for (int sb = group; sb < TW / VS_TOK; sb += EPI_GROUPS) {
        const int t0 = sb * VS_TOK;
        if (t0 >= rows_left) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + acc * TW + t0, v);
#pragma unroll
        for (int t = 0; t < 32; ++t) ctg[t * EPI_CT_LD + et] = v[t];
        epi_bar(group);
        Tile tl{ctg, tile, tile * BM, tb * tw + t0, min(VS_TOK, rows_left - t0)};
        epilogue<EPI, EPI_THREADS>(p, tl, et);
        epi_bar(group);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

int g_sms = 0;

template <int EPI>
cudaError_t launch_w(const GemmArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_wide_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  const int n_tiles = (a.N + BM - 1) / BM;
  // token blocks: the count that minimises waves x (block width + the fixed weight-tile share, measured ~128 token
  // rows worth); width = multiple of 16
  int n_tb = (a.T + TW - 1) / TW, tw = TW;
  {
    long best = -1;
    for (int nb = (a.T + TW - 1) / TW; nb <= (a.T + 63) / 64; ++nb) {
      const int w = (((a.T + nb - 1) / nb) + 15) / 16 * 16;
      if (w > TW) continue;
      const long units = (long)n_tiles * nb, waves = (units + g_sms - 1) / g_sms;
      const long cost = waves * (w + 128);
      if (best < 0 || cost < best) { best = cost; n_tb = nb; tw = w; }
    }
    n_tb = (a.T + tw - 1) / tw;
  }
  CUtensorMap tmW, tmX;
  if (!gemm_get_map(a.W, a.ldw, a.N, a.K, BM, &tmW) || !gemm_get_map(a.X, a.ldx, a.T, a.K, tw, &tmX))
    return cudaErrorInvalidValue;
  const int n_units = n_tiles * n_tb;
  const int grid = n_units < g_sms ? n_units : g_sms;
  prof_begin(PROF_GEMM, 2.0 * a.N * a.K, st);
  gemm_wide_kernel<EPI><<<grid, THREADS, SMEM_BYTES, st>>>(tmW, tmX, a, n_tb, n_units, tw);
  cudaError_t e = cudaGetLastError();
  prof_end(st);
  count_launch();
  return e;
}

}  // namespace

cudaError_t launch_gemm_wide(const GemmArgs& a, cudaStream_t st) {
  switch (a.epi) {
    case EPI_STORE: return launch_w<EPI_STORE>(a, st);
    case EPI_RESID: return launch_w<EPI_RESID>(a, st);
    case EPI_SWIGLU: return launch_w<EPI_SWIGLU>(a, st);
    case EPI_HEADS: return launch_w<EPI_HEADS>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vs
