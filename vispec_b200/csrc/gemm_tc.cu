// Weight-streaming GEMM on the 5th-generation tensor cores:  Y[T<=32 per block, N] = X[T,K] * W[N,K]^T.
//
// (The mma.sync bring-up kernel this replaced was removed in round 2.)
// ncu on the mma.sync version showed the legacy tensor pipe 25 % busy at only 31 % of HBM bandwidth and a
// 1.3-wave grid, i.e. it could never reach the HBM roofline.  Here:
//   * persistent kernel, one CTA per SM, work units (128-row tile x k-split x token block) dealt round-robin
//     with the split factor chosen so the unit count is a whole number of 148-CTA waves;
//   * warp 0 = TMA producer: W tile [128 x 64] and X tile [32 x 64] per stage, 128-byte swizzle, 8 stages
//     (160 KB in flight per SM), weights hinted EVICT_FIRST, activations EVICT_LAST;
//   * warp 1 = MMA issuer: tcgen05.mma cta_group::1 kind::f16, M=128 (weight rows) N=32 (tokens) K=16,
//     fp32 accumulators double-buffered in TMEM (2 x 32 columns) so the next unit's MMAs overlap this unit's epilogue;
//   * warps 4-7 = epilogue: tcgen05.ld, deterministic split-K reduction (partials in L2, last arriver sums in split
//     order), then the fused epilogue (bias / residual / SiLU*up / RoPE + in-place KV append);
//   * programmatic dependent launch: the producer prefetches the first weight stages BEFORE griddepcontrol.wait
//     (weights never depend on the previous kernel), so the HBM stream continues across kernel boundaries.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "gemm_epilogue.cuh"
#include "kernels.h"
#include "profile.h"
#include "sm100.cuh"

namespace vs {


namespace {

constexpr int BM = 128, BK = 64, STAGES = 8;
constexpr int GEMM_WIDE_MIN_ROWS = 64;        // more rows than this: gemm_wide.cu (MMA N = 256) instead of the weight stream
constexpr int W_STAGE = BM * BK * 2;           // 16384
constexpr int X_STAGE = VS_TOK * BK * 2;       // 4096
constexpr int STAGE_BYTES = W_STAGE + X_STAGE; // 20480 (multiple of 1024: every tile stays 1024-byte aligned)
constexpr int CT_BYTES = VS_TOK * EPI_CT_LD * 4;
constexpr int BAR_OFF = STAGES * STAGE_BYTES + CT_BYTES;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;   // + barriers/flags + slack for manual 1024-byte alignment
constexpr int THREADS = 256;
constexpr int EPI_THREADS = 128;
constexpr int TMEM_COLS = 64;

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct Unit {
  int tile, split, tb, kb, ke;
};
// Unit order: k-split fastest, then the token block, then the weight-row tile.  Units are dealt round-robin, so the CTAs
// that run at the same time work on the SAME weight tile for different token blocks: with hundreds of rows (prompt
// passes) a weight tile comes from HBM once and is shared through L2 instead of being re-streamed per token block.
__device__ __forceinline__ Unit decode_unit(int u, int n_tb, int splits, int nk) {
  Unit r;
  r.split = u % splits;
  int q = u / splits;
  r.tb = q % n_tb;
  r.tile = q / n_tb;
  r.kb = (int)(((long)nk * r.split) / splits);
  r.ke = (int)(((long)nk * (r.split + 1)) / splits);
  return r;
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const GemmArgs p,
               const int n_tiles, const int n_units) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space
  float* ct = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = p.K / BK;
  const int splits = p.splits;
  const int n_tb = (p.T + VS_TOK - 1) / VS_TOK;

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
      mbar_init(&tempty[i], EPI_THREADS);
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
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0, phase = 0;
      // Phase A: the first STAGES weight tiles do not depend on the previous kernel -> issue them before the
      // dependency wait; their activation tiles follow right after it.
      int pre_u[STAGES], pre_kc[STAGES];
      int n_pre = 0;
      int u = blockIdx.x;
      Unit un = u < n_units ? decode_unit(u, n_tb, splits, nk) : Unit{0, 0, 0, 0, 0};
      int kc = un.kb;
      while (u < n_units && n_pre < STAGES) {
        mbar_expect_tx(&full[n_pre], STAGE_BYTES);
        tma_load_2d(&tmW, &full[n_pre], smem + n_pre * STAGE_BYTES, kc * BK, un.tile * BM, HINT_EVICT_FIRST);
        pre_u[n_pre] = un.tb;
        pre_kc[n_pre] = kc;
        ++n_pre;
        if (++kc >= un.ke) {
          u += gridDim.x;
          if (u < n_units) {
            un = decode_unit(u, n_tb, splits, nk);
            kc = un.kb;
          }
        }
      }
      pdl_wait();
      for (int i = 0; i < n_pre; ++i)
        tma_load_2d(&tmX, &full[i], smem + i * STAGE_BYTES + W_STAGE, pre_kc[i] * BK, pre_u[i] * VS_TOK,
                    HINT_EVICT_LAST);
      stage = n_pre % STAGES;
      phase = (n_pre == STAGES) ? 1 : 0;
      // Phase B: steady state
      while (u < n_units) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], STAGE_BYTES);
        uint8_t* dst = smem + stage * STAGE_BYTES;
        tma_load_2d(&tmW, &full[stage], dst, kc * BK, un.tile * BM, HINT_EVICT_FIRST);
        tma_load_2d(&tmX, &full[stage], dst + W_STAGE, kc * BK, un.tb * VS_TOK, HINT_EVICT_LAST);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++kc >= un.ke) {
          u += gridDim.x;
          if (u < n_units) {
            un = decode_unit(u, n_tb, splits, nk);
            kc = un.kb;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        Unit un = decode_unit(u, n_tb, splits, nk);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 32;
        for (int kc = un.kb; kc < un.ke; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_desc(smem + stage * STAGE_BYTES);
          const uint64_t db = make_desc(smem + stage * STAGE_BYTES + W_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) tc_mma(d_tmem, da + 2 * k, db + 2 * k, (kc > un.kb || k > 0) ? 1u : 0u);
          tc_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (4 warps, TMEM lanes 32*(warp%4)...) =====================
    const int et = threadIdx.x - 128;          // 0..127 == tile row
    const int lane_base = (warp & 3) * 32;
    int acc = 0, acc_phase = 0;
    pdl_wait();
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      Unit un = decode_unit(u, n_tb, splits, nk);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + acc * 32, v);
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;

      bool run_epilogue = true;
      if (splits > 1) {
        // partial tile layout [row][32 tokens]: each thread owns 128 contiguous bytes
        const long tile_floats = (long)VS_TOK * BM;
        float* base = p.partials + ((long)un.tb * splits * n_tiles) * tile_floats;
        float4* mine = reinterpret_cast<float4*>(base + ((long)un.split * n_tiles + un.tile) * tile_floats + et * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) mine[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        __threadfence();
        epi_bar();
        if (et == 0) {
          int old = atomicAdd(&p.counters[un.tb * n_tiles + un.tile], 1);
          *s_flag = (old == splits - 1);
        }
        epi_bar();
        run_epilogue = *s_flag != 0;
        if (run_epilogue) {
          __threadfence();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0.f;
          for (int s = 0; s < splits; ++s) {        // fixed order -> deterministic
            const float4* src = reinterpret_cast<const float4*>(base + ((long)s * n_tiles + un.tile) * tile_floats + et * 32);
            float4 t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = __ldcg(src + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] += t[i].x; v[4 * i + 1] += t[i].y; v[4 * i + 2] += t[i].z; v[4 * i + 3] += t[i].w;
            }
          }
          if (et == 0) p.counters[un.tb * n_tiles + un.tile] = 0;
        }
      }
      if (run_epilogue) {
#pragma unroll
        for (int t = 0; t < 32; ++t) ct[t * EPI_CT_LD + et] = v[t];
      }
      epi_bar();
      if (run_epilogue) {
        Tile tl{ct, un.tile, un.tile * BM, un.tb * VS_TOK, min(VS_TOK, p.T - un.tb * VS_TOK)};
        epilogue<EPI, EPI_THREADS>(p, tl, et);
      }
      epi_bar();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// ---- host side: tensor-map cache ------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  long ld;
  int rows, cols, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && ld == o.ld && rows == o.rows && cols == o.cols && box_rows == o.box_rows;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.ld * 0x9E3779B97F4A7C15ull + ((size_t)k.rows << 20) + ((size_t)k.cols << 4) + (size_t)k.box_rows;
    return h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D bf16 tensor [rows][cols] with row pitch ld elements; box = [box_rows][64 cols], 128-byte swizzle, OOB -> 0
bool get_map(const void* ptr, long ld, int rows, int cols, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, ld, rows, cols, box_rows};
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) {
    *out = it->second;
    return true;
  }
  EncodeFn enc = get_encode();
  if (!enc) return false;
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (g_maps.size() > 65536) g_maps.clear();
  g_maps.emplace(key, m);
  *out = m;
  return true;
}

int g_num_sms = 0;

template <int EPI>
cudaError_t launch_t(const GemmArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  CUtensorMap tmW, tmX;
  if (!get_map(a.W, a.ldw, a.N, a.K, BM, &tmW) || !get_map(a.X, a.ldx, a.T, a.K, VS_TOK, &tmX))
    return cudaErrorInvalidValue;
  const int n_tiles = (a.N + BM - 1) / BM;
  const int tok_blocks = (a.T + VS_TOK - 1) / VS_TOK;
  const int n_units = n_tiles * a.splits * tok_blocks;
  const int grid = n_units < g_num_sms ? n_units : g_num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  prof_begin(PROF_GEMM, 2.0 * a.N * a.K * tok_blocks, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<EPI>, tmW, tmX, a, n_tiles, n_units);
  prof_end(st);
  count_launch();
  return e;
}

bool use_wide() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VS_GEMM_WIDE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

}  // namespace

cudaError_t launch_gemm_wide(const GemmArgs& a, cudaStream_t st);   // gemm_wide.cu
bool gemm_get_map(const void* ptr, long ld, int rows, int cols, int box_rows, CUtensorMap* out) {
  return get_map(ptr, ld, rows, cols, box_rows, out);
}

// Split factor: as many k-splits as keep >= 4 k-chunks per unit, chosen so that the number of work units is as
// close as possible to (and not above) a whole number of waves of one CTA per SM.
int gemm_pick_splits(int N, int K, int T) {
  const int sms = 148;
  int n_tiles = (N + BM - 1) / BM;
  int tok_blocks = (T + VS_TOK - 1) / VS_TOK;
  int nk = K / BK;
  long base = (long)n_tiles * tok_blocks;
  int max_s = nk / 4;
  if (max_s < 1) max_s = 1;
  if (max_s > 16) max_s = 16;
  int best = 1;
  double best_eff = -1.0;
  for (int s = 1; s <= max_s; ++s) {
    long units = base * s;
    long waves = (units + sms - 1) / sms;
    double eff = (double)units / (double)(waves * sms);          // SM occupancy of the last wave included
    double cost = eff - 0.004 * s;                                // mild preference for fewer partial tiles
    if (units < sms) cost -= 1.0 - (double)units / sms;           // not even one full wave: strongly prefer more splits
    if (cost > best_eff) {
      best_eff = cost;
      best = s;
    }
  }
  return best;
}

size_t gemm_partials_floats(int N, int T, int splits) {
  if (splits <= 1) return 0;
  size_t n_tiles = (N + BM - 1) / BM;
  size_t tok_blocks = (T + VS_TOK - 1) / VS_TOK;
  return tok_blocks * splits * n_tiles * VS_TOK * BM;
}

cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t st) {
  if (a.T <= 0) return cudaSuccess;
  if (a.K % BK != 0 || a.ldw % 8 != 0 || a.ldx % 8 != 0) return cudaErrorInvalidValue;
  if (a.splits > 1 && (!a.partials || !a.counters)) return cudaErrorInvalidValue;
  if (a.splits > a.K / BK) return cudaErrorInvalidValue;
  if (a.T > GEMM_WIDE_MIN_ROWS && use_wide()) return launch_gemm_wide(a, st);   // prompt passes: 256-token MMA tiles
  switch (a.epi) {
    case EPI_STORE: return launch_t<EPI_STORE>(a, st);
    case EPI_RESID: return launch_t<EPI_RESID>(a, st);
    case EPI_SWIGLU: return launch_t<EPI_SWIGLU>(a, st);
    case EPI_HEADS: return launch_t<EPI_HEADS>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace vs
