// The persistent decode kernel: ONE launch of one CTA per SM executes a whole "program" -- a list of phases covering
// e.g. a complete draft+verify cycle (32 target layers + LM head + accept + KV splice + 4 draft forwards + tree build)
// -- instead of ~320 separate launches.
//
// Why: decode with <= 32 token rows is a pure weight stream (13.2 GB per verify step at 7B).  Separate
// launches stop the HBM stream at every kernel boundary (ramp-up, TMEM allocation, barrier init, tail) and ncu showed the
// split-K "last arriver" epilogue turning single CTAs into stragglers (profiles/README.md, r1a).  Here
//   * warp 0 (one lane) is the WEIGHT producer: it walks every GEMM phase of the program in order and keeps an
//     8-stage / 128 KB TMA ring full.  Weights never depend on activations, so it runs ahead across phase boundaries:
//     the HBM stream continues while the rest of the CTA is in an epilogue, an attention phase or a grid barrier;
//   * warp 3 (one lane) prefetches the weight tiles `pf_ahead` chunks further ahead into L2
//     (cp.async.bulk.prefetch.tensor), so HBM stays busy even when the smem ring is full;
//   * warp 2 (one lane) is the ACTIVATION producer: the [32 x 64] token tile of each stage is loaded once the phase
//     that produces it has completed on every CTA (it polls the grid-barrier counter);
//   * warp 1 (one lane) issues tcgen05.mma (M=128 weight rows, N=32 tokens, K=16; fp32 accumulators double-buffered in
//     TMEM) and commits stages back to the producers;
//   * warps 4-11 (two teams of 128 threads) are the workers: they drain TMEM into fp32 partial tiles, and run every
//     non-GEMM phase (reduction + fused epilogues, RMSNorm, attention, arg-max, log-softmax/top-k, tree build / accept /
//     KV splice) on team-sized work items dealt round-robin over all 2 x 148 teams.
// GEMM work is split "stream-K" style: the (tile, k-chunk) sequence of a weight matrix is cut into one contiguous,
// equally sized range per CTA, so every SM streams the same number of bytes; tiles shared by several CTAs are reduced
// in k order by the following PH_EPI phase (deterministic; no atomics on data, no last-arriver serialisation).
// Phases are separated by a grid barrier (one atomic + acquire-spin per CTA).
#include "program.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/vispec_b200.h"
#include "attn_dev.cuh"
#include "profile.h"
#include "sm100.cuh"

namespace vs {

namespace {

constexpr int PG_THREADS = 384;
constexpr int PG_WORKERS = 256;
constexpr int PG_TEAM = 128;
constexpr int BM = 128, BK = 64, NS = 8;
constexpr int W_STAGE = BM * BK * 2;             // 16384
constexpr int X_STAGE = VS_TOK * BK * 2;         // 4096
constexpr int STAGE_BYTES = W_STAGE + X_STAGE;   // 20480 (multiple of 1024)
constexpr int TEAM_SMEM = 25600;
constexpr int BAR_OFF = NS * STAGE_BYTES + 2 * TEAM_SMEM;
constexpr int SMEM_BYTES = BAR_OFF + 512 + 1024;
constexpr int TMEM_COLS = 64;
constexpr int TILE_FLOATS = VS_TOK * BM;         // one partial tile: [32 tokens][128 rows] fp32
static_assert(sizeof(AttnSmem) <= TEAM_SMEM, "team scratch too small for attention");
static_assert(sizeof(TreeSmem) <= TEAM_SMEM, "team scratch too small for the tree ops");

struct TeamSync {
  int id;
  __device__ __forceinline__ void operator()() const { named_bar_sync(id, PG_TEAM); }
};
struct WorkerSync {
  __device__ __forceinline__ void operator()() const { named_bar_sync(1, PG_WORKERS); }
};

__device__ __forceinline__ long rng_lo(int b, long C, int G) { return (long)b * C / G; }

__device__ __forceinline__ float ldcg_bf(const bf16* p) {
  unsigned short u = __ldcg(reinterpret_cast<const unsigned short*>(p));
  return __uint_as_float(((unsigned int)u) << 16);
}
__device__ __forceinline__ float bf_lo(unsigned int u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(unsigned int u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ unsigned int pack_bf2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned int*>(&v);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------
// grid barrier (workers only): bar -> one thread: release fence, arrive, acquire-spin, fence -> bar
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target, int wtid) {
  named_bar_sync(1, PG_WORKERS);
  if (wtid == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) {
    }
    __threadfence();
  }
  named_bar_sync(1, PG_WORKERS);
}

// ---------------------------------------------------------------------------------------------------------------
// PH_EPI: one weight-row tile (128 output features x T tokens) per team item
// ---------------------------------------------------------------------------------------------------------------
// Calls f(partial_tile_base) for every CTA that contributed k-chunks to `tile`, in increasing k order.
template <class F>
__device__ __forceinline__ void for_each_contributor(const Phase& ph, int tile, int G, F f) {
  const long C = (long)ph.n_tiles * ph.nk;
  const long k0 = (long)tile * ph.nk, k1 = k0 + ph.nk;
  int b = (int)(k0 * G / C);
  while (b > 0 && rng_lo(b, C, G) > k0) --b;
  while (b + 1 < G && rng_lo(b + 1, C, G) <= k0) ++b;
  for (; b < G; ++b) {
    const long lo = rng_lo(b, C, G), hi = rng_lo(b + 1, C, G);
    if (lo >= k1) break;
    if (hi <= lo || hi <= k0) continue;
    const int seg = tile - (int)(lo / ph.nk);
    f(ph.partials + ((long)b * ph.maxseg + seg) * TILE_FLOATS);
  }
}

__device__ __noinline__ void epi_store_resid(const Phase& ph, int tile, int G, int tt, float* red, TeamSync sync) {
  const int T = ph.T;
  float acc[VS_TOK];
#pragma unroll
  for (int t = 0; t < VS_TOK; ++t) acc[t] = 0.f;
  for_each_contributor(ph, tile, G, [&](const float* P) {
#pragma unroll
    for (int t = 0; t < VS_TOK; ++t)
      if (t < T) acc[t] += __ldcg(P + t * BM + tt);
  });
  const int n = tile * BM + tt;
  const bool ok = n < ph.N;
  const float bias = (ok && ph.bias) ? bf2f(ph.bias[n]) : 0.f;
  const bool resid = ph.epi == EPI_RESID;
  const int warp = tt >> 5, lane = tt & 31;
#pragma unroll
  for (int t = 0; t < VS_TOK; ++t) {
    if (t < T) {                                                   // uniform across the team
      float v = acc[t] + bias;
      if (resid) {
        float o = rbf(v);                                          // Linear output rounds to bf16 ...
        v = (ok ? ldcg_bf(ph.R + (long)t * ph.ldr + n) : 0.f) + o; // ... then the residual add rounds again
      }
      bf16 y = f2bf(v);
      if (ok) ph.Y[(long)t * ph.ldy + n] = y;
      if (ph.ssq) {
        float q = ok ? bf2f(y) * bf2f(y) : 0.f;
        q = warp_sum(q);
        if (lane == 0) red[warp * VS_TOK + t] = q;
      }
    }
  }
  if (ph.ssq) {
    sync();
    if (tt < T) ph.ssq[tile * VS_TOK + tt] = red[tt] + red[VS_TOK + tt] + red[2 * VS_TOK + tt] + red[3 * VS_TOK + tt];
    sync();
  }
}

__device__ __noinline__ void epi_swiglu(const Phase& ph, int tile, int G, int tt) {
  // tile rows [0,64) = gate features j0..j0+63, rows [64,128) = the matching up features
  const int T = ph.T;
  const int r = tt & 63, t0 = (tt >> 6) * 16;
  float g[16], u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { g[i] = 0.f; u[i] = 0.f; }
  for_each_contributor(ph, tile, G, [&](const float* P) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (t0 + i < T) {
        g[i] += __ldcg(P + (t0 + i) * BM + r);
        u[i] += __ldcg(P + (t0 + i) * BM + 64 + r);
      }
  });
  const int j = tile * 64 + r;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (t0 + i < T) {
      float a = rbf(silu_f(rbf(g[i])));                            // act_fn(gate) rounds, then * up rounds
      ph.Y[(long)(t0 + i) * ph.ldy + j] = f2bf(a * rbf(u[i]));
    }
}

__device__ __noinline__ void epi_heads(const Phase& ph, int tile, int G, int tt) {
  // one 128-row tile == one head of q | k | v; optional RoPE; q -> Qout, k/v -> KV rows in place
  const int T = ph.T;
  const int d = tt & 63, t0 = (tt >> 6) * 16;
  float a1[16], a2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a1[i] = 0.f; a2[i] = 0.f; }
  for_each_contributor(ph, tile, G, [&](const float* P) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (t0 + i < T) {
        a1[i] += __ldcg(P + (t0 + i) * BM + d);
        a2[i] += __ldcg(P + (t0 + i) * BM + 64 + d);
      }
  });
  const bool is_q = tile < ph.n_q_tiles;
  const bool is_k = !is_q && tile < ph.n_q_tiles + ph.n_k_tiles;
  const int head = is_q ? tile : (is_k ? tile - ph.n_q_tiles : tile - ph.n_q_tiles - ph.n_k_tiles);
  const int row_base = (ph.kv_row_base ? *((volatile const int*)ph.kv_row_base) : 0) + ph.kv_row_off;
  const float b1 = ph.bias ? bf2f(ph.bias[tile * BM + d]) : 0.f;
  const float b2 = ph.bias ? bf2f(ph.bias[tile * BM + 64 + d]) : 0.f;
  const bool rope = ph.rope && (is_q || is_k);
  bf16* dst = is_q ? ph.Qout + (long)head * ph.q_head_stride
                   : (is_k ? ph.Kout : ph.Vout) + (long)head * ph.kv_head_stride + (long)row_base * 128;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int t = t0 + i;
    if (t < T) {
      float x1 = rbf(a1[i] + b1), x2 = rbf(a2[i] + b2);
      float o1 = x1, o2 = x2;
      if (rope) {                                                  // three bf16 roundings (cnets_ours.py:117-118)
        const long pos = *((volatile const int*)(ph.pos_ids + t));
        const float c1 = bf2f(ph.cos[pos * 128 + d]), s1 = bf2f(ph.sin[pos * 128 + d]);
        const float c2 = bf2f(ph.cos[pos * 128 + 64 + d]), s2 = bf2f(ph.sin[pos * 128 + 64 + d]);
        o1 = rbf(x1 * c1) + rbf(-x2 * s1);
        o2 = rbf(x2 * c2) + rbf(x1 * s2);
      }
      dst[(long)t * 128 + d] = f2bf(o1);
      dst[(long)t * 128 + 64 + d] = f2bf(o2);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// row phases (one token row per team item)
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void norm_row(const Phase& ph, int t, int tt, float* red, TeamSync sync) {
  const bf16* x = ph.src + (long)t * ph.lds;
  bf16* y = ph.dst + (long)t * ph.ldd;
  const int H = ph.H;
  float ss = 0.f;
  if (ph.ssq_in) {
    for (int i = 0; i < ph.ssq_tiles; ++i) ss += __ldcg(ph.ssq_in + i * VS_TOK + t);   // fixed order, same on every thread
  } else {
    for (int i = tt * 8; i < H; i += PG_TEAM * 8) {
      uint4 v = __ldcg(reinterpret_cast<const uint4*>(x + i));
      const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = bf_lo(w[j]), b = bf_hi(w[j]);
        ss += a * a + b * b;
      }
    }
    ss = warp_sum(ss);
    sync();
    if ((tt & 31) == 0) red[tt >> 5] = ss;
    sync();
    ss = red[0] + red[1] + red[2] + red[3];
  }
  const float rstd = 1.0f / sqrtf(ss / (float)H + ph.eps);
  for (int i = tt * 8; i < H; i += PG_TEAM * 8) {
    uint4 v = __ldcg(reinterpret_cast<const uint4*>(x + i));
    uint4 wv = *reinterpret_cast<const uint4*>(ph.w + i);
    const unsigned int xi[4] = {v.x, v.y, v.z, v.w};
    const unsigned int wi[4] = {wv.x, wv.y, wv.z, wv.w};
    unsigned int o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)   // weight * bf16(x * rstd): two bf16 roundings (cnets_ours.py:523-527)
      o[j] = pack_bf2(bf_lo(wi[j]) * rbf(bf_lo(xi[j]) * rstd), bf_hi(wi[j]) * rbf(bf_hi(xi[j]) * rstd));
    *reinterpret_cast<uint4*>(y + i) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __forceinline__ void copy_row(const Phase& ph, int t, int tt) {
  int r = ph.broadcast ? 0 : (ph.idx ? *((volatile const int*)(ph.idx + t)) : t);
  if (ph.n_src_rows > 0) r = min(max(r, 0), ph.n_src_rows - 1);
  const uint4* s = reinterpret_cast<const uint4*>(ph.src + (long)r * ph.lds);
  uint4* d = reinterpret_cast<uint4*>(ph.dst + (long)t * ph.ldd);
  for (int i = tt; i < ph.H / 8; i += PG_TEAM) d[i] = __ldcg(s + i);
}

__device__ __forceinline__ unsigned long long team_max_u64(unsigned long long v, unsigned long long* red, int tt,
                                                           TeamSync sync) {
  v = warp_max_u64(v);
  sync();
  if ((tt & 31) == 0) red[tt >> 5] = v;
  sync();
  unsigned long long r = red[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) r = red[i] > r ? red[i] : r;
  return r;
}
__device__ __forceinline__ float team_max_f(float v, float* red, int tt, TeamSync sync) {
  v = warp_max(v);
  sync();
  if ((tt & 31) == 0) red[tt >> 5] = v;
  sync();
  return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
}
__device__ __forceinline__ float team_sum_f(float v, float* red, int tt, TeamSync sync) {
  v = warp_sum(v);
  sync();
  if ((tt & 31) == 0) red[tt >> 5] = v;
  sync();
  return red[0] + red[1] + red[2] + red[3];
}

// visits the V logits of one row as (value, index); 16-byte loads when the row allows it
template <class F>
__device__ __forceinline__ void scan_row(const bf16* xr, int V, int tt, F f) {
  if ((V & 7) == 0 && (reinterpret_cast<uintptr_t>(xr) & 15) == 0) {
    for (int i = tt * 8; i < V; i += PG_TEAM * 8) {
      uint4 v = __ldcg(reinterpret_cast<const uint4*>(xr + i));
      const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f(bf_lo(w[j]), i + 2 * j);
        f(bf_hi(w[j]), i + 2 * j + 1);
      }
    }
  } else {
    for (int i = tt; i < V; i += PG_TEAM) f(ldcg_bf(xr + i), i);
  }
}

__device__ __noinline__ void argmax_row(const Phase& ph, int t, int tt, void* red, TeamSync sync) {
  const bf16* xr = ph.logits + (long)t * ph.ldl;
  unsigned long long best = 0ull;
  scan_row(xr, ph.V, tt, [&](float v, int i) {
    unsigned long long k = rank_key(v, (uint32_t)i);
    best = k > best ? k : best;
  });
  best = team_max_u64(best, reinterpret_cast<unsigned long long*>(red), tt, sync);
  if (tt == 0) ph.amax_out[t] = (int)key_index(best);
}

// lp = bf16((x - max) - log(sum exp(x - max))) in fp32 math, then the k best by (value desc, index asc), best first
// (cnets_ours.py:1111-1115,1145-1148).
__device__ __noinline__ void lsm_topk_row(const Phase& ph, int item, int tt, void* red, TeamSync sync) {
  constexpr int KMAX = 8;
  const int row = ph.row_sel ? *((volatile const int*)ph.row_sel) : item;
  const bf16* xr = ph.logits + (long)row * ph.ldl;
  const int V = ph.V;
  float m = -INFINITY;
  scan_row(xr, V, tt, [&](float v, int) { m = fmaxf(m, v); });
  m = team_max_f(m, reinterpret_cast<float*>(red), tt, sync);
  float s = 0.f;
  scan_row(xr, V, tt, [&](float v, int) { s += expf(v - m); });
  s = team_sum_f(s, reinterpret_cast<float*>(red), tt, sync);
  const float lse = logf(s);
  unsigned long long loc[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) loc[j] = 0ull;
  scan_row(xr, V, tt, [&](float v, int i) {
    float lp = rbf((v - m) - lse);
    unsigned long long key = rank_key(lp, (uint32_t)i);
    if (key > loc[KMAX - 1]) {
      loc[KMAX - 1] = key;
#pragma unroll
      for (int j = KMAX - 1; j > 0; --j) {
        if (loc[j] > loc[j - 1]) {
          unsigned long long tmp = loc[j];
          loc[j] = loc[j - 1];
          loc[j - 1] = tmp;
        }
      }
    }
  });
  for (int r = 0; r < ph.k; ++r) {
    unsigned long long best = team_max_u64(loc[0], reinterpret_cast<unsigned long long*>(red), tt, sync);
    if (loc[0] == best && best != 0ull) {     // keys are unique (index in the low word): exactly one owner pops
#pragma unroll
      for (int j = 0; j < KMAX - 1; ++j) loc[j] = loc[j + 1];
      loc[KMAX - 1] = 0ull;
    }
    if (tt == 0) {
      ph.top_val[item * ph.k + r] = key_value(best);
      ph.top_idx[item * ph.k + r] = (int)key_index(best);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
struct GemmCursor {            // walks the GEMM phases of the program and this CTA's chunk range in each
  const Phase* phases;
  int n_phases;
  int p;                       // current phase index (a PH_GEMM phase) or n_phases
  long c, hi;                  // next chunk, end of this CTA's range
  int nk;
  __device__ __forceinline__ void init(const ProgramHeader& h) {
    phases = h.phases;
    n_phases = h.n_phases;
    p = -1;
    c = hi = 0;
    nk = 1;
  }
  // advance to the next phase in which this CTA has work; false when the program is exhausted
  __device__ __forceinline__ bool next_phase(int b, int G) {
    for (++p; p < n_phases; ++p) {
      if (phases[p].type != PH_GEMM) continue;
      const long C = (long)phases[p].n_tiles * phases[p].nk;
      c = rng_lo(b, C, G);
      hi = rng_lo(b + 1, C, G);
      nk = phases[p].nk;
      if (c < hi) return true;
    }
    return false;
  }
};

__global__ void __launch_bounds__(PG_THREADS, 1) program_kernel(const ProgramHeader hdr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* team_base = smem + NS * STAGE_BYTES;
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full_x = full_w + NS;
  uint64_t* empty = full_x + NS;
  uint64_t* tfull = empty + NS;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  volatile int* w_issued = reinterpret_cast<volatile int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, b = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(&full_w[i], 1);
      mbar_init(&full_x[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PG_WORKERS / 32);
    }
    *w_issued = 0;
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
    // ===================== weight producer =====================
    if (lane == 0) {
      GemmCursor cur;
      cur.init(hdr);
      int stage = 0, par = 0, issued = 0;
      while (cur.next_phase(b, G)) {
        const CUtensorMap* tm = hdr.maps + hdr.phases[cur.p].tmW;
        int tile = (int)(cur.c / cur.nk), kc = (int)(cur.c % cur.nk);
        for (long c = cur.c; c < cur.hi; ++c) {
          mbar_wait(&empty[stage], par ^ 1);
          mbar_expect_tx(&full_w[stage], W_STAGE);
          tma_load_2d(tm, &full_w[stage], smem + stage * STAGE_BYTES, kc * BK, tile * BM, HINT_EVICT_FIRST);
          *w_issued = ++issued;
          if (++stage == NS) { stage = 0; par ^= 1; }
          if (++kc == cur.nk) { kc = 0; ++tile; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      GemmCursor cur;
      cur.init(hdr);
      int stage = 0, par = 0, acc = 0, acc_par = 0;
      while (cur.next_phase(b, G)) {
        long c = cur.c;
        int tile = (int)(c / cur.nk);
        while (c < cur.hi) {
          const long seg_end = min(cur.hi, (long)(tile + 1) * cur.nk);
          mbar_wait(&tempty[acc], acc_par ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 32;
          bool first = true;
          for (; c < seg_end; ++c) {
            mbar_wait(&full_w[stage], par);
            mbar_wait(&full_x[stage], par);
            tc_fence_after();
            const uint64_t da = make_desc(smem + stage * STAGE_BYTES);
            const uint64_t db = make_desc(smem + stage * STAGE_BYTES + W_STAGE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) tc_mma(d_tmem, da + 2 * k, db + 2 * k, (!first || k > 0) ? 1u : 0u);
            tc_commit(&empty[stage]);
            first = false;
            if (++stage == NS) { stage = 0; par ^= 1; }
          }
          tc_commit(&tfull[acc]);
          acc ^= 1;
          if (acc == 0) acc_par ^= 1;
          ++tile;
        }
      }
    }
  } else if (warp == 2) {
    // ===================== activation producer =====================
    if (lane == 0) {
      GemmCursor cur;
      cur.init(hdr);
      int stage = 0, par = 0;
      while (cur.next_phase(b, G)) {
        // the token tile is produced by earlier phases of this launch: wait until all of them completed everywhere
        const unsigned int need = (unsigned int)cur.p * (unsigned int)G;
        while (ld_acquire_u32(hdr.counters) < need) __nanosleep(64);
        fence_proxy_async();
        const CUtensorMap* tm = hdr.maps + hdr.phases[cur.p].tmX;
        int kc = (int)(cur.c % cur.nk);
        for (long c = cur.c; c < cur.hi; ++c) {
          mbar_wait(&empty[stage], par ^ 1);
          mbar_expect_tx(&full_x[stage], X_STAGE);
          tma_load_2d(tm, &full_x[stage], smem + stage * STAGE_BYTES + W_STAGE, kc * BK, 0, HINT_EVICT_LAST);
          if (++stage == NS) { stage = 0; par ^= 1; }
          if (++kc == cur.nk) kc = 0;
        }
      }
    }
  } else if (warp == 3) {
    // ===================== L2 prefetcher =====================
    if (lane == 0 && hdr.pf_ahead > 0) {
      GemmCursor cur;
      cur.init(hdr);
      int idx = 0;
      while (cur.next_phase(b, G)) {
        const CUtensorMap* tm = hdr.maps + hdr.phases[cur.p].tmW;
        int tile = (int)(cur.c / cur.nk), kc = (int)(cur.c % cur.nk);
        for (long c = cur.c; c < cur.hi; ++c, ++idx) {
          while (idx >= *w_issued + hdr.pf_ahead) __nanosleep(100);
          if (idx >= *w_issued + NS) tma_prefetch_2d(tm, kc * BK, tile * BM);   // not about to be loaded anyway
          if (++kc == cur.nk) { kc = 0; ++tile; }
        }
      }
    }
  } else {
    // ===================== workers (two teams of 128 threads) =====================
    const int wtid = threadIdx.x - 4 * 32;          // 0..255
    const int team = wtid >> 7, tt = wtid & 127;
    const int wwarp = wtid >> 5;                    // 0..7
    const TeamSync tsync{2 + team};
    uint8_t* tsm = team_base + team * TEAM_SMEM;
    const int gteam = b * 2 + team, n_teams = 2 * G;
    int acc = 0, acc_par = 0;
    for (int p = 0; p < hdr.n_phases; ++p) {
      const Phase& ph = hdr.phases[p];
      switch (ph.type) {
        case PH_GEMM: {
          const long C = (long)ph.n_tiles * ph.nk;
          const long lo = rng_lo(b, C, G), hi = rng_lo(b + 1, C, G);
          if (lo < hi) {
            const int tile0 = (int)(lo / ph.nk), tile1 = (int)((hi - 1) / ph.nk);
            const int quarter = wwarp & 3, half = wwarp >> 2;
            for (int tile = tile0; tile <= tile1; ++tile) {
              mbar_wait(&tfull[acc], acc_par);
              tc_fence_after();
              float v[16];
              tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 32 + half * 16, v);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
              acc ^= 1;
              if (acc == 0) acc_par ^= 1;
              float* dst = ph.partials + ((long)b * ph.maxseg + (tile - tile0)) * TILE_FLOATS + (half * 16) * BM +
                           quarter * 32 + lane;
#pragma unroll
              for (int i = 0; i < 16; ++i) dst[i * BM] = v[i];
            }
          }
        } break;
        case PH_EPI: {
          for (int item = gteam; item < ph.n_tiles; item += n_teams) {
            if (ph.epi == EPI_SWIGLU) epi_swiglu(ph, item, G, tt);
            else if (ph.epi == EPI_HEADS) epi_heads(ph, item, G, tt);
            else epi_store_resid(ph, item, G, tt, reinterpret_cast<float*>(tsm), tsync);
          }
        } break;
        case PH_NORM: {
          for (int t = gteam; t < ph.T; t += n_teams) norm_row(ph, t, tt, reinterpret_cast<float*>(tsm), tsync);
        } break;
        case PH_ROWS: {
          for (int t = gteam; t < ph.T; t += n_teams) copy_row(ph, t, tt);
        } break;
        case PH_ATTN_S: {
          const AttnArgs& a = ph.attn;
          const int n_ctx = a.n_ctx ? *((volatile const int*)a.n_ctx) : a.n_ctx_host;
          const int n_chunks = (n_ctx + a.n_new + ATT_CHUNK - 1) / ATT_CHUNK;
          const int n_items = a.n_heads * n_chunks;
          AttnSmem& sm = *reinterpret_cast<AttnSmem*>(tsm);
          for (int item = gteam; item < n_items; item += n_teams)
            attn_scores_item(a, sm, item / n_chunks, 0, 1, item % n_chunks, true, tt, tsync);
        } break;
        case PH_ATTN_PV: {
          const AttnArgs& a = ph.attn;
          const int n_items = a.n_heads * a.pv_splits;
          AttnSmem& sm = *reinterpret_cast<AttnSmem*>(tsm);
          for (int item = gteam; item < n_items; item += n_teams)
            attn_pv_item(a, sm, item / a.pv_splits, 0, 1, item % a.pv_splits, a.pv_splits, tt, tsync);
        } break;
        case PH_ARGMAX: {
          for (int t = gteam; t < ph.T; t += n_teams) argmax_row(ph, t, tt, tsm, tsync);
        } break;
        case PH_LSM_TOPK: {
          for (int t = gteam; t < ph.T; t += n_teams) lsm_topk_row(ph, t, tt, tsm, tsync);
        } break;
        case PH_TREE: {
          if (b == 0) {
            TreeSmem& sm = *reinterpret_cast<TreeSmem*>(team_base);
            named_bar_sync(1, PG_WORKERS);
            if (ph.tree_op == TREE_LEVEL0) tree_level0_body(ph.tb, ph.state, wtid);
            else if (ph.tree_op == TREE_EXPAND) tree_expand_body(ph.tb, ph.state, ph.level, sm, wtid, WorkerSync());
            else if (ph.tree_op == TREE_BUILD)
              tree_build_body(ph.tb, ph.state, ph.sample_token_dev, ph.sample_token_host, sm, wtid, PG_WORKERS, WorkerSync());
            else if (ph.tree_op == TREE_ACCEPT) tree_accept_body(ph.tb, ph.state, ph.amax, ph.out_rec, sm, wtid, WorkerSync());
            else if (ph.tree_op == TREE_COMMIT) tree_commit_body(ph.state, wtid);
          }
        } break;
        case PH_SPLICE: {
          for (long item = gteam; item < ph.n_items; item += n_teams) kv_splice_item(ph.splice, item, tt, tsync);
        } break;
        default: break;
      }
      if (p + 1 < hdr.n_phases) grid_barrier(hdr.counters, (unsigned int)(p + 1) * (unsigned int)G, wtid);
      if (hdr.trace && b == 0 && wtid == 0) hdr.trace[p] = global_timer();
    }
    // leave the counters at zero for the next launch: the last CTA to get here resets them (every CTA has passed its
    // last barrier wait before it increments the exit counter)
    named_bar_sync(1, PG_WORKERS);
    if (wtid == 0) {
      __threadfence();
      unsigned int old = atomicAdd(hdr.counters + 1, 1u);
      if (old == (unsigned int)G - 1) {
        hdr.counters[0] = 0;
        hdr.counters[1] = 0;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace

}  // namespace vs

// =================================================================================================================
// host side: program builder (opaque vs_program of include/vispec_b200.h)
// =================================================================================================================
struct vs_program {
  std::vector<vs::Phase> phases;
  struct MapReq { const void* ptr; long ld; int rows, cols, box_rows; };
  std::vector<MapReq> maps;
  size_t partial_floats = 0;
  size_t ssq_floats = 0;
  int grid = 0;
  int pf_ahead = 24;
  bool committed = false;
  bool trace_on = false;
  // device workspace layout (set by commit)
  uint8_t* ws = nullptr;
  size_t off_maps = 0, off_phases = 0, off_counters = 0, off_trace = 0, off_ssq = 0, off_partials = 0, total = 0;
  std::string err;
};

namespace vs {
namespace {

thread_local std::string g_perr;
int pfail(int code, const char* where, const std::string& what) {
  g_perr = std::string(where) + ": " + what;
  return code;
}

int device_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  return sms;
}

int add_map(vs_program* pg, const void* ptr, long ld, int rows, int cols, int box_rows) {
  for (size_t i = 0; i < pg->maps.size(); ++i) {
    const auto& m = pg->maps[i];
    if (m.ptr == ptr && m.ld == ld && m.rows == rows && m.cols == cols && m.box_rows == box_rows) return (int)i;
  }
  pg->maps.push_back({ptr, ld, rows, cols, box_rows});
  return (int)pg->maps.size() - 1;
}

Phase blank(int type, int T) {
  Phase ph;
  memset(&ph, 0, sizeof(ph));
  ph.type = type;
  ph.T = T;
  return ph;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void layout(vs_program* pg) {
  size_t off = 0;
  pg->off_maps = off;
  off = align_up(off + pg->maps.size() * sizeof(CUtensorMap), 256);
  pg->off_phases = off;
  off = align_up(off + pg->phases.size() * sizeof(Phase), 256);
  pg->off_counters = off;
  off += 256;
  pg->off_trace = off;
  off = align_up(off + pg->phases.size() * sizeof(unsigned long long), 256);
  pg->off_ssq = off;
  off = align_up(off + pg->ssq_floats * sizeof(float), 256);
  pg->off_partials = off;
  off += pg->partial_floats * sizeof(float);
  pg->total = align_up(off, 256);
}

}  // namespace

// ---- phase constructors used by the C API below and by the composite builders ------------------------------------
int prog_add_linear(vs_program* pg, const GemmArgs& a, float* ssq) {
  if (pg->committed) return pfail(VS_ERR_INVALID, "vs_program_add_linear", "program already committed");
  if (a.T < 1 || a.T > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_linear", "programs handle 1..32 token rows");
  if (a.K % BK != 0 || a.ldw % 8 != 0 || a.ldx % 8 != 0 || a.N < 1)
    return pfail(VS_ERR_INVALID, "vs_program_add_linear", "K must be a multiple of 64, ldx/ldw of 8");
  if ((a.epi == EPI_SWIGLU || a.epi == EPI_HEADS) && a.N % 128 != 0)
    return pfail(VS_ERR_INVALID, "vs_program_add_linear", "SWIGLU / HEADS need N % 128 == 0");
  const int G = pg->grid ? pg->grid : device_sms();
  pg->grid = G;
  Phase g = blank(PH_GEMM, a.T);
  g.n_tiles = (a.N + BM - 1) / BM;
  g.nk = a.K / BK;
  g.N = a.N;
  const long C = (long)g.n_tiles * g.nk;
  const long per = (C + G - 1) / G;
  g.maxseg = (int)((per + g.nk - 1) / g.nk) + 1;
  g.tmW = add_map(pg, a.W, a.ldw, a.N, a.K, BM);
  g.tmX = add_map(pg, a.X, a.ldx, a.T, a.K, VS_TOK);
  const size_t need = (size_t)G * g.maxseg * TILE_FLOATS;
  if (need > pg->partial_floats) pg->partial_floats = need;
  pg->phases.push_back(g);
  Phase e = g;
  e.type = PH_EPI;
  e.epi = a.epi;
  e.bias = a.bias;
  e.Y = a.Y; e.ldy = a.ldy; e.R = a.R; e.ldr = a.ldr;
  e.ssq = ssq;
  if (ssq == PROG_WS_SSQ && (size_t)g.n_tiles * VS_TOK > pg->ssq_floats) pg->ssq_floats = (size_t)g.n_tiles * VS_TOK;
  e.n_q_tiles = a.n_q_tiles; e.n_k_tiles = a.n_k_tiles; e.rope = a.rope;
  e.cos = a.cos; e.sin = a.sin; e.pos_ids = a.pos_ids;
  e.Qout = a.Qout; e.q_head_stride = a.q_head_stride;
  e.Kout = a.Kout; e.Vout = a.Vout; e.kv_head_stride = a.kv_head_stride;
  e.kv_row_base = a.kv_row_base; e.kv_row_off = a.kv_row_off;
  pg->phases.push_back(e);
  return VS_OK;
}

int prog_add_norm(vs_program* pg, const bf16* x, long ldx, const bf16* w, bf16* y, long ldy, int T, int H, float eps,
                  const float* ssq_in, int ssq_tiles) {
  if (T < 1 || T > VS_TOK || H % 8 || ldx % 8 || ldy % 8)
    return pfail(VS_ERR_INVALID, "vs_program_add_rmsnorm", "1..32 rows, H and strides multiples of 8");
  Phase ph = blank(PH_NORM, T);
  ph.src = x; ph.lds = ldx; ph.w = w; ph.dst = y; ph.ldd = ldy; ph.H = H; ph.eps = eps;
  ph.ssq_in = ssq_in; ph.ssq_tiles = ssq_tiles;
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_rows(vs_program* pg, const bf16* src, long lds, const int* idx, int n_src_rows, int broadcast, bf16* dst,
                  long ldd, int T, int H) {
  if (T < 1 || H % 8 || lds % 8 || ldd % 8)
    return pfail(VS_ERR_INVALID, "vs_program_add_rows", "H and strides must be multiples of 8");
  Phase ph = blank(PH_ROWS, T);
  ph.src = src; ph.lds = lds; ph.idx = idx; ph.n_src_rows = n_src_rows; ph.broadcast = broadcast;
  ph.dst = dst; ph.ldd = ldd; ph.H = H;
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_attention(vs_program* pg, const AttnArgs& a) {
  if (a.Tq < 1 || a.Tq > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_attention", "programs handle 1..32 query rows");
  if (!a.scores || !a.stats || !a.opart || !a.counters || a.pv_splits < 1 || a.max_chunks < 1)
    return pfail(VS_ERR_INVALID, "vs_program_add_attention", "missing workspace");
  const int G = pg->grid ? pg->grid : device_sms();
  pg->grid = G;
  Phase s = blank(PH_ATTN_S, a.Tq);
  s.attn = a;
  // as many kv splits as give every team one (head, split) item, bounded by the caller's workspace
  int sp = (2 * G) / (a.n_heads > 0 ? a.n_heads : 1);
  if (sp < 1) sp = 1;
  if (sp > a.pv_splits) sp = a.pv_splits;
  s.attn.pv_splits = sp;
  pg->phases.push_back(s);
  Phase v = s;
  v.type = PH_ATTN_PV;
  pg->phases.push_back(v);
  return VS_OK;
}

int prog_add_argmax(vs_program* pg, const bf16* logits, long ldl, int T, int V, int* out) {
  Phase ph = blank(PH_ARGMAX, T);
  ph.logits = logits; ph.ldl = ldl; ph.V = V; ph.amax_out = out;
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_lsm_topk(vs_program* pg, const bf16* logits, long ldl, int T, int V, int k, const int* row_sel, float* top_val,
                      int* top_idx) {
  if (k < 1 || k > 8) return pfail(VS_ERR_INVALID, "vs_program_add_logsoftmax_topk", "k must be in 1..8");
  if (row_sel && T != 1) return pfail(VS_ERR_INVALID, "vs_program_add_logsoftmax_topk", "row_sel selects exactly one row");
  Phase ph = blank(PH_LSM_TOPK, T);
  ph.logits = logits; ph.ldl = ldl; ph.V = V; ph.k = k; ph.row_sel = row_sel; ph.top_val = top_val; ph.top_idx = top_idx;
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_tree(vs_program* pg, int op, const TreeBufs& tb, int* state, int level, const int* amax, int* out_rec,
                  const int* sample_token_dev, int sample_token_host) {
  const int n_sc = tb.top_k + tb.depth * tb.top_k * tb.top_k;
  if (tb.top_k < 1 || tb.top_k > 8 || tb.depth + 2 > VS_RETRIEVE_LD || tb.total + 1 > 64 || tb.total > n_sc ||
      n_sc > TREE_MAX_SCORES)
    return pfail(VS_ERR_INVALID, "vs_program_add_tree_op", "unsupported tree shape");
  Phase ph = blank(PH_TREE, 0);
  ph.tree_op = op; ph.tb = tb; ph.state = state; ph.level = level; ph.amax = amax; ph.out_rec = out_rec;
  ph.sample_token_dev = sample_token_dev; ph.sample_token_host = sample_token_host;
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_splice(vs_program* pg, bf16* slab, int n_slabs, int n_kv, long max_pos, const int* state, const int* sel_dev) {
  if (!slab || !state || !sel_dev) return pfail(VS_ERR_INVALID, "vs_program_add_kv_splice", "null operand");
  Phase ph = blank(PH_SPLICE, 0);
  ph.splice.slab = slab; ph.splice.head_stride = max_pos * 128; ph.splice.st = state; ph.splice.sel_dev = sel_dev;
  ph.n_items = n_slabs * n_kv;
  pg->phases.push_back(ph);
  return VS_OK;
}

const char* prog_last_error() { return g_perr.c_str(); }
int prog_fail(int code, const char* where, const char* what) { return pfail(code, where, what); }

// encode one tensor map on the host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn prog_encode_fn() {
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

}  // namespace vs

extern "C" {

int vs_program_create(vs_program** out) {
  if (!out) return vs::pfail(VS_ERR_INVALID, "vs_program_create", "null out");
  *out = new vs_program();
  const char* e = getenv("VS_PF_AHEAD");
  if (e) (*out)->pf_ahead = atoi(e);
  return VS_OK;
}

void vs_program_destroy(vs_program* pg) { delete pg; }

const char* vs_program_last_error(void) { return vs::prog_last_error(); }

int vs_program_num_phases(const vs_program* pg) { return pg ? (int)pg->phases.size() : 0; }

int vs_program_set_option(vs_program* pg, const char* name, int value) {
  if (!pg || !name) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "null argument");
  if (!strcmp(name, "pf_ahead")) pg->pf_ahead = value;
  else if (!strcmp(name, "trace")) pg->trace_on = value != 0;
  else if (!strcmp(name, "grid")) {
    if (!pg->phases.empty()) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "grid must be set before adding phases");
    pg->grid = value;
  } else return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "unknown option");
  return VS_OK;
}

size_t vs_program_workspace_bytes(vs_program* pg) {
  if (!pg) return 0;
  vs::layout(pg);
  return pg->total;
}

int vs_program_commit(vs_program* pg, void* workspace, size_t bytes, void* stream) {
  using namespace vs;
  if (!pg || !workspace) return pfail(VS_ERR_INVALID, "vs_program_commit", "null argument");
  if (pg->phases.empty()) return pfail(VS_ERR_INVALID, "vs_program_commit", "empty program");
  if (pg->grid <= 0) pg->grid = device_sms();
  layout(pg);
  if (bytes < pg->total) return pfail(VS_ERR_INVALID, "vs_program_commit", "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return pfail(VS_ERR_INVALID, "vs_program_commit", "workspace must be 256-byte aligned");
  EncodeFn enc = prog_encode_fn();
  if (!enc && !pg->maps.empty()) return pfail(VS_ERR_CUDA, "vs_program_commit", "cuTensorMapEncodeTiled unavailable");
  pg->ws = static_cast<uint8_t*>(workspace);
  std::vector<uint8_t> host(pg->off_partials, 0);
  for (size_t i = 0; i < pg->maps.size(); ++i) {
    const auto& m = pg->maps[i];
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)m.cols, (cuuint64_t)m.rows};
    cuuint64_t strides[1] = {(cuuint64_t)m.ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)m.box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(m.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return pfail(VS_ERR_CUDA, "vs_program_commit", "cuTensorMapEncodeTiled failed");
    memcpy(host.data() + pg->off_maps + i * sizeof(CUtensorMap), &tm, sizeof(tm));
  }
  float* partials = reinterpret_cast<float*>(pg->ws + pg->off_partials);
  for (size_t i = 0; i < pg->phases.size(); ++i) {
    Phase ph = pg->phases[i];
    if (ph.type == PH_GEMM || ph.type == PH_EPI) ph.partials = partials;
    if (ph.ssq == PROG_WS_SSQ) ph.ssq = reinterpret_cast<float*>(pg->ws + pg->off_ssq);
    if (ph.ssq_in == PROG_WS_SSQ) ph.ssq_in = reinterpret_cast<float*>(pg->ws + pg->off_ssq);
    memcpy(host.data() + pg->off_phases + i * sizeof(Phase), &ph, sizeof(Phase));
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(pg->ws, host.data(), host.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return pfail(VS_ERR_CUDA, "vs_program_commit", cudaGetErrorString(cudaGetLastError()));
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return pfail(VS_ERR_CUDA, "vs_program_commit", cudaGetErrorString(e));
    attr_set = true;
  }
  pg->committed = true;
  return VS_OK;
}

int vs_program_run(const vs_program* pg, void* stream) {
  using namespace vs;
  if (!pg || !pg->committed) return pfail(VS_ERR_INVALID, "vs_program_run", "program not committed");
  ProgramHeader h;
  h.phases = reinterpret_cast<const Phase*>(pg->ws + pg->off_phases);
  h.n_phases = (int)pg->phases.size();
  h.maps = reinterpret_cast<const CUtensorMap*>(pg->ws + pg->off_maps);
  h.counters = reinterpret_cast<unsigned int*>(pg->ws + pg->off_counters);
  h.trace = pg->trace_on ? reinterpret_cast<unsigned long long*>(pg->ws + pg->off_trace) : nullptr;
  h.pf_ahead = pg->pf_ahead;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pg->grid);
  cfg.blockDim = dim3(PG_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident (the phases synchronise across the grid)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, program_kernel, h);
  count_launch();
  if (e != cudaSuccess) return pfail(VS_ERR_CUDA, "vs_program_run", cudaGetErrorString(e));
  return VS_OK;
}

/* per-phase end times (ns, %globaltimer of CTA 0) of the last run; needs option "trace" = 1 */
int vs_program_read_trace(const vs_program* pg, unsigned long long* out, int cap) {
  using namespace vs;
  if (!pg || !pg->committed || !pg->trace_on) return pfail(VS_ERR_INVALID, "vs_program_read_trace", "trace not enabled");
  int n = (int)pg->phases.size();
  if (n > cap) n = cap;
  if (cudaMemcpy(out, pg->ws + pg->off_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return pfail(VS_ERR_CUDA, "vs_program_read_trace", cudaGetErrorString(cudaGetLastError()));
  return n;
}
int vs_program_phase_types(const vs_program* pg, int32_t* out, int cap) {
  if (!pg) return 0;
  int n = (int)pg->phases.size();
  if (n > cap) n = cap;
  for (int i = 0; i < n; ++i) out[i] = pg->phases[i].type;
  return n;
}

}  // extern "C"
