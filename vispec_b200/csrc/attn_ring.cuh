// Single-phase decode attention of the persistent program kernel with TMA-staged K/V (PH_ATTN_FUSED).
//
// Same algorithm, items and rounding points as the two-phase kernels of attention.cu (reference: eager attention
// modeling_llama_kv.py:602-623 with the mask of :890-924; draft SDPA cnets_ours.py:393-445):
//   item = (head, kv split) on a 256-thread team, every item of the phase resident at once;
//   pass 1  S = Q K^T, mask, scale, split (max, sum exp)  ->  exchange of the split statistics per head (arrival counter)
//   pass 2  P = softmax in the reference's rounding, O += P V  ->  exchange of the partial outputs per head.
// What is different from the cp.async bodies of attn_dev.cuh is the data path:
//   * K and V reach shared memory through ONE FIFO of 8 KB granules (32 keys x 128 dims) per team, filled by the KV
//     producer warp of the CTA with tensor-map TMA loads (`cp.async.bulk.tensor`, 128-byte swizzle, completion on
//     mbarriers).  The producer walks the program ahead of the workers: the K granules of the fully visible prefix
//     [0, n_ctx) do not depend on this launch's activations, so they are in shared memory BEFORE the phase starts; the V
//     granules follow in the same FIFO as soon as a K slot is released.
//   * pass 1: a warp owns (16 query rows) x (16 keys of every other granule); pass 2: (16 query rows) x (16 keys of every
//     granule) x (64 of the 128 output dims).  No team barrier inside either pass: warps run decoupled behind the full /
//     empty mbarriers of the FIFO.
//   * up to ATT_RING_NG granules per warp the scores stay in registers and the probabilities pass through a bf16 P tile
//     in shared memory; beyond that the scores are parked in the L2-resident score scratch between the passes (16 bytes
//     per thread and granule in eager mode, one coalesced 512-byte store per warp) and each pass-2 warp forms its
//     probabilities from them directly in the A-fragment registers of the second MMA.
#pragma once
#include "attn_dev.cuh"
#include "sm100.cuh"

namespace vs {

constexpr int ATT_GRAN = 32;                              // keys per granule
constexpr int ATT_GRAN_BYTES = ATT_GRAN * ATT_ROW_B;      // 8192: [d half][32 keys][64 dims] bf16, 128-byte swizzle
constexpr int ATT_SLOTS = 4;                              // FIFO depth per team (64 KB in flight per SM)
constexpr int ATT_RING_NG = 2;                            // granules per warp whose scores stay in registers

struct __align__(1024) AttnRingSmem {
  uint8_t ring[ATT_SLOTS][ATT_GRAN_BYTES];   // K / V granules (TMA); reused for the team's O reduction at the end of an item
  uint8_t q[VS_TOK * ATT_ROW_B];             // Q tile [32][128] bf16 (att_swz256)
  float red[ATT_WARPS][VS_TOK];              // per-warp row max
  float red2[ATT_WARPS][VS_TOK];             // per-warp sum exp
  float gM[VS_TOK], gL[VS_TOK], gR[VS_TOK];  // global row max, softmax denominator, its reciprocal
};

__device__ __forceinline__ uint32_t pack_bf2_rn(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf_lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi_f(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// byte offset of (key row, 16-byte chunk of the 256-byte row) inside a granule as the TMA unit writes it
__device__ __forceinline__ int att_gran_off(int row, int chunk16) {
  return ((chunk16 >> 3) << 12) + row * 128 + (((chunk16 & 7) ^ (row & 7)) << 4);
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3,
                                            uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
      : "memory");
}

// L2 prefetch of one box (no shared-memory destination, no completion)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// How a (head, split) item cuts the keys: granules [g_lo, g_hi) of ceil(kv_len / 32).  Producer and workers both use this.
struct AttRingCut {
  int n_gran, cps, g_lo, n;
};
__device__ __forceinline__ AttRingCut att_ring_cut(int kv_len, int split, int n_splits) {
  AttRingCut c;
  c.n_gran = (kv_len + ATT_GRAN - 1) / ATT_GRAN;
  c.cps = (c.n_gran + n_splits - 1) / n_splits;
  c.g_lo = min(split * c.cps, c.n_gran);
  c.n = min(c.g_lo + c.cps, c.n_gran) - c.g_lo;
  return c;
}

// ---------------------------------------------------------------------------------------------------------------------
// KV producer (one lane of the CTA's KV warp): the granule FIFO of both attention teams for ONE attention phase.
// `fifo[t]`: granules issued to team t so far in this launch (slot = index & 3, mbarrier parity = (index >> 2) & 1).
// Granules that hold rows >= n_ctx were written by this launch (q|k|v epilogue): they wait for `bars_before` grid
// barriers and are ordered behind the generic-proxy writes by a proxy fence; everything else is issued as soon as a slot
// is free.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void att_ring_produce(const AttnArgs& a, const CUtensorMap* mapK, const CUtensorMap* mapV, int two_loads,
                                                 int n_ctx, uint8_t* team_base, int team_stride, uint64_t* bars,
                                                 unsigned int* fifo, const unsigned int* grid_ctr, unsigned int need_tail,
                                                 bool tail_ok, int b, int G, int pf_dist) {
  const int n_splits = a.pv_splits;
  const int kv_len = n_ctx + a.n_new;
  const int tail_g = n_ctx / ATT_GRAN;                    // first granule that holds rows written by this launch
  const int group = a.n_heads / a.n_kv_heads;
  const uint64_t hint = group == 1 ? HINT_EVICT_FIRST : 0x1000000000000000ull /* evict normal */;
  int n[2], g_lo[2], kvh[2], issued[2], pf[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int gb = t * G + b;
    n[t] = 0; g_lo[t] = 0; kvh[t] = 0; issued[t] = 0; pf[t] = 0;
    if (gb < a.n_heads * n_splits) {
      const AttRingCut c = att_ring_cut(kv_len, gb % n_splits, n_splits);
      n[t] = c.n;
      g_lo[t] = c.g_lo;
      kvh[t] = (gb / n_splits) / group;
    }
  }
  while (issued[0] < 2 * n[0] || issued[1] < 2 * n[1]) {
    bool progress = false;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (issued[t] >= 2 * n[t]) continue;
      // L2 prefetch `pf_dist` granules ahead of the FIFO: four 8 KB slots per team are ~2 us of HBM latency at 15 GB/s per
      // team; with the lines already in L2 the same slots turn around several times faster.  (Rows this launch is still
      // writing may be prefetched too: L2 is the point of coherence, the later stores update the line.)
      while (pf[t] < 2 * n[t] && pf[t] < issued[t] + pf_dist) {
        const int j = pf[t];
        const bool pv = j >= n[t];
        const int pg = g_lo[t] + (pv ? j - n[t] : j);
        const CUtensorMap* pm = pv ? mapV : mapK;
        if (j >= issued[t] + ATT_SLOTS) {       // the next ATT_SLOTS granules are about to be loaded anyway
          if (two_loads) {
            tma_prefetch_3d(pm, 0, pg * ATT_GRAN, kvh[t]);
            tma_prefetch_3d(pm, 64, pg * ATT_GRAN, kvh[t]);
          } else {
            tma_prefetch_4d(pm, 0, pg * ATT_GRAN, 0, kvh[t]);
          }
        }
        ++pf[t];
      }
      const int i = issued[t];
      const bool is_v = i >= n[t];
      const int g = g_lo[t] + (is_v ? i - n[t] : i);
      if (g >= tail_g && !tail_ok) {
        if (ld_relaxed_u32(grid_ctr) < need_tail) continue;
        fence_proxy_async();
        tail_ok = true;
      }
      const unsigned int f = fifo[t] + (unsigned int)i;
      const int slot = f & (ATT_SLOTS - 1);
      uint64_t* full = bars + t * 2 * ATT_SLOTS + slot;
      uint64_t* empty = full + ATT_SLOTS;
      if (!mbar_try_wait(empty, ((f >> 2) & 1u) ^ 1u)) continue;
      uint8_t* dst = team_base + t * team_stride + slot * ATT_GRAN_BYTES;
      mbar_expect_tx(full, ATT_GRAN_BYTES);
      const CUtensorMap* m = is_v ? mapV : mapK;
      if (two_loads) {       // 3-D map {128 dims, rows, kv heads}: one box per 64-dim half
        tma_load_3d(m, full, dst, 0, g * ATT_GRAN, kvh[t], hint);
        tma_load_3d(m, full, dst + ATT_GRAN_BYTES / 2, 64, g * ATT_GRAN, kvh[t], hint);
      } else {               // 4-D map {64 dims, rows, 2 halves, kv heads}: the whole granule in one box
        tma_load_4d(m, full, dst, 0, g * ATT_GRAN, 0, kvh[t], hint);
      }
#ifdef VS_PROGRAM_CLK
      if (b == 0 && t == 0 && i == 0) g_clk[22] = clock64();
      if (b == 0 && t == 0 && i == n[0]) g_clk[24] = clock64();
#endif
      ++issued[t];
      progress = true;
    }
    if (!progress) __nanosleep(40);
  }
  fifo[0] += 2u * (unsigned int)n[0];
  fifo[1] += 2u * (unsigned int)n[1];
}

// ---------------------------------------------------------------------------------------------------------------------
// shared tails (same exchanges as attn_dev.cuh, on this shared-memory layout)
// ---------------------------------------------------------------------------------------------------------------------
constexpr float ATT_LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exp for softmax STATISTICS (row maxima / denominators): ex2.approx, relative error ~2^-22 -- far below the bf16 rounding of
// the probabilities that are divided by the result; the probabilities themselves use expf.  exp(-inf) == 0.
__device__ __forceinline__ float att_exp_stat(float x) { return att_ex2(x * ATT_LOG2E); }
// The PROBABILITIES use expf itself: a cheaper exp of the same accuracy (ex2.approx with the product's rounding error
// recovered by two FMAs, <= 2 ulp, 7 instructions) was 1 % faster at 8192 keys, but it is not BIT-identical to expf in every
// element, and on random-init weights one flipped bf16 rounding is enough to part the generated tokens from the
// reference's (bench.py gpu_eager_baseline: common prefix 4 of 256 instead of 256 of 256).  Reverted.

// split statistics -> head-wide softmax denominators; on entry sm.red / sm.red2 hold every warp's (max, sum exp) per row.
// The exchange is ONE L2 round trip: every split stores its (m, l) per row as two 8-byte words that carry the launch tag in
// their upper half (an aligned 8-byte store is single-copy atomic: value and flag arrive together, no fence, no counter),
// and every reader spins on exactly the words it needs until they carry this launch's tag.
// `ll`: [head][split][32 rows][2] words in the program workspace, private to the phase.
__device__ __forceinline__ void att_ll_store(unsigned long long* at, float m, float l, unsigned int tag) {
  const unsigned long long w0 = ((unsigned long long)tag << 32) | __float_as_uint(m);
  const unsigned long long w1 = ((unsigned long long)tag << 32) | __float_as_uint(l);
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(at), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ float2 att_ll_wait(const unsigned long long* at, unsigned int tag) {
  unsigned long long w0, w1;
  do {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(at) : "memory");
  } while ((unsigned int)(w0 >> 32) != tag || (unsigned int)(w1 >> 32) != tag);
  return make_float2(__uint_as_float((unsigned int)w0), __uint_as_float((unsigned int)w1));
}
template <class Sync>
__device__ __forceinline__ void att_ring_exchange_stats(const AttnArgs& p, AttnRingSmem& sm, int head, int split, int n_splits,
                                                        unsigned int* head_cnt, unsigned int epoch, int tid, Sync sync) {
  unsigned long long* ll = reinterpret_cast<unsigned long long*>(head_cnt + ((2 * p.n_heads + 3) & ~3)) +
                           (long)head * n_splits * VS_TOK * 2;
  const unsigned int tag = epoch + 1u;
  if (tid < VS_TOK) {
    float m = sm.red[0][tid];
#pragma unroll
    for (int w = 1; w < ATT_WARPS; ++w) m = fmaxf(m, sm.red[w][tid]);
    float l = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) {                           // fixed order
      const float mw = sm.red[w][tid];
      if (mw != -INFINITY) l += sm.red2[w][tid] * att_exp_stat(mw - m);
    }
    att_ll_store(ll + ((long)split * VS_TOK + tid) * 2, m, l, tag);
  }
  // ---- exact global denominators from the split statistics: 8 lanes per row, splits part, part+8 (fixed order)
  {
    const int q = tid >> 3, part = tid & 7;
    float m = -INFINITY, l = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (part + 8 * j < n_splits) {
        const float2 st = att_ll_wait(ll + ((long)(part + 8 * j) * VS_TOK + q) * 2, tag);
        if (st.x != -INFINITY) {
          const float mn = fmaxf(m, st.x);
          l = l * att_exp_stat(m - mn) + st.y * att_exp_stat(st.x - mn);
          m = mn;
        }
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      const float a = m == -INFINITY ? 0.f : l * att_exp_stat(m - mn);
      const float b = mo == -INFINITY ? 0.f : lo * att_exp_stat(mo - mn);
      l = a + b;
      m = mn;
    }
    if (part == 0) { sm.gM[q] = m; sm.gL[q] = l; sm.gR[q] = 1.0f / l; }
  }
  sync();
}

// second per-head arrival counter, then every split reduces its share of the head's output in split order
template <bool EAGER, class Sync>
__device__ __forceinline__ void att_ring_reduce_output(const AttnArgs& p, const AttnRingSmem& sm, int head, int split,
                                                       int n_splits, int n_act, int tq_rows, unsigned int* head_cnt,
                                                       unsigned int epoch, int tid, Sync sync) {
  const float* obase = p.opart + ((long)head * n_splits) * VS_TOK * 128;
  bf16* const O = p.O;
  const long ldo = p.ldo;
  sync();
  if (tid == 0) {
    unsigned int* cnt2 = head_cnt + p.n_heads + head;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt2) : "memory");
    const unsigned int want = (epoch + 1u) * (unsigned int)n_splits;
    while (att_ld_relaxed(cnt2) < want) {
    }
  }
  sync();
  for (int i = split + tid * n_splits; i < VS_TOK * 32; i += ATT_THREADS * n_splits) {
    const int q = i >> 5, d4 = (i & 31) * 4;
    if (q >= tq_rows) continue;
    const float* src = obase + q * 128 + d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int s0 = 0; s0 < n_act; s0 += 8) {                         // 8 splits per round trip
      float4 t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        t[j] = __ldcg(reinterpret_cast<const float4*>(src + (long)min(s0 + j, n_act - 1) * VS_TOK * 128));
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (s0 + j < n_act) { v.x += t[j].x; v.y += t[j].y; v.z += t[j].z; v.w += t[j].w; }
    }
    if (!EAGER) {
      const float l = sm.gL[q];
      v.x /= l; v.y /= l; v.z /= l; v.w /= l;
    }
    bf16 o4[4] = {f2bf(v.x), f2bf(v.y), f2bf(v.z), f2bf(v.w)};
    *reinterpret_cast<uint2*>(O + (long)q * ldo + head * 128 + d4) = *reinterpret_cast<const uint2*>(o4);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// the item.  STREAM = false: at most ATT_RING_NG granules per warp, scores stay in registers between the passes.
// EAGER: the target's rounding points (bf16 QK^T, bf16 division by sqrt(d), fp32 softmax rounded to bf16); otherwise fp32
// scores and the division by the denominator after P V (SDPA-style).
// `p`: the phase's arguments (in shared memory: the staged phase descriptor); `n_ctx`: the kv prefix length the dispatcher
// read; `bars`: full[ATT_SLOTS] | empty[ATT_SLOTS] of this team; `fifo0`: FIFO index of this item's first K granule.
// Returns the number of FIFO entries consumed (2 n).
// ---------------------------------------------------------------------------------------------------------------------

template <bool STREAM, bool EAGER, class Sync>
__device__ __forceinline__ unsigned int attn_ring_item(const AttnArgs& p, int n_ctx, AttnRingSmem& sm, uint64_t* bars,
                                                       unsigned int fifo0, int head, int split, int n_splits,
                                                       unsigned int* head_cnt, unsigned int epoch, int tid, Sync sync) {
  const int warp = tid >> 5, lane = tid & 31;
  const int mb = warp & 1, kg = (warp >> 1) & 1, sp = warp >> 2;     // query half, 16-key group of a granule, granule parity
  const int g = lane >> 2, t4 = lane & 3;
  // private copies of the arguments (every use of `p` behind an asm barrier would be a fresh load)
  const int kv_len = n_ctx + p.n_new;
  const bool causal = p.causal != 0;
  const float scale = p.scale;
  float* const scores = p.scores;
  const AttRingCut cut = att_ring_cut(kv_len, split, n_splits);
  const int n = cut.n, g_lo = cut.g_lo;
  const int tq_rows = min(VS_TOK, p.Tq);
  uint64_t* full = bars;
  uint64_t* empty = bars + ATT_SLOTS;
  const int q0 = 16 * mb + g, q1 = q0 + 8;                           // the two query rows of this thread
  const int kcol = 16 * kg + 2 * t4;                                 // key column (inside a granule) of acc[0][0]; acc[1]: +8

#ifdef VS_PROGRAM_CLK
  const bool aclk_on = blockIdx.x == 0 && tid == 0 && head == 0;
#endif
  ACLK(0);
  // ---- Q tile + tree words
  att_load_rows_async(sm.q, p.Q + (long)head * p.q_head_stride, 0, tq_rows, VS_TOK, tid);
  cp_async_commit();
  unsigned long long qb0 = 0ull, qb1 = 0ull;
  if (!causal && p.bits) {
    qb0 = __ldcg(p.bits + min(q0, tq_rows - 1));
    qb1 = __ldcg(p.bits + min(q1, tq_rows - 1));
  }
  const float inv_div = sqrtf((float)VS_HEAD_DIM), r_div = 1.0f / inv_div;
  // A-fragment address of this lane in the Q tile (k-step 0); the fragments are re-read per granule (32 registers less)
  const int qa_row = mb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, qa_ch = lane >> 4;
  // the B-fragment address of this lane inside a granule (k-step 0): row = key, 16-byte chunk = d / 8
  const int k_row = 16 * kg + (lane & 7) + ((lane >> 4) & 1) * 8, k_ch = (lane >> 3) & 1;
  const int v_row = 16 * kg + (lane & 7) + ((lane >> 3) & 1) * 8, v_ch = lane >> 4;
  // parked scores (STREAM): 8 values per thread and granule, [head][granule][128 threads of the granule][8];
  // eager scores are bf16 values already: 16 bytes per thread, otherwise 32
  const long park0 = (((long)head * (p.score_ld / ATT_GRAN) + g_lo) * 128 + (mb + 2 * kg) * 32 + lane);
  cp_async_wait<0>();
  sync();
  ACLK(1);

  // ================= pass 1: S = Q K^T =================
  float sreg[ATT_RING_NG][2][4];
  float rm0 = -INFINITY, rm1 = -INFINITY, rl0 = 0.f, rl1 = 0.f;
  // S of granule i -> acc, scaled in the reference's rounding; returns true when the granule needs masks
  auto scores_of = [&](int i, float (&acc)[2][4]) {
    const unsigned int f = fifo0 + (unsigned int)i;
    const int slot = f & (ATT_SLOTS - 1);
    mbar_wait(&full[slot], (f >> 2) & 1u);
    if (i == 0) ACLK(2);
    const uint8_t* kt = sm.ring[slot];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t bfr[4], qf[4];
      ldmatrix_x4(qf, sm.q + att_swz256(qa_row, kk * 2 + qa_ch));
      ldmatrix_x4(bfr, kt + att_gran_off(k_row, kk * 2 + k_ch));
      mma_bf16(acc[0], qf, bfr[0], bfr[1]);
      mma_bf16(acc[1], qf, bfr[2], bfr[3]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_n(&empty[slot], 2);                   // 4 of the 8 warps read a K granule, all 8 a V granule
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float s = acc[a][e];
        acc[a][e] = EAGER ? rbf(att_div(rbf(s), inv_div, r_div)) : s * scale;
      }
  };
  auto mask_scores = [&](int key0, float (&acc)[2][4]) {             // granules that are not inside the fully visible prefix
    const unsigned int v0 = (unsigned int)att_visible_mask(causal, n_ctx, kv_len, q0, qb0, key0);
    const unsigned int v1 = (unsigned int)att_visible_mask(causal, n_ctx, kv_len, q1, qb1, key0);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const unsigned int vis = (e & 2) ? v1 : v0;
        if (!((vis >> (kcol + 8 * a + (e & 1))) & 1u)) acc[a][e] = -INFINITY;
      }
  };
  if (!STREAM) {
#pragma unroll
    for (int u = 0; u < ATT_RING_NG; ++u) {
      const int i = sp + 2 * u;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) sreg[u][a][e] = -INFINITY;
      if (i < n) {
        scores_of(i, sreg[u]);
        const int key0 = (g_lo + i) * ATT_GRAN;
        if (key0 + ATT_GRAN > n_ctx) mask_scores(key0, sreg[u]);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          rm0 = fmaxf(rm0, fmaxf(sreg[u][a][0], sreg[u][a][1]));
          rm1 = fmaxf(rm1, fmaxf(sreg[u][a][2], sreg[u][a][3]));
        }
      }
    }
    // exact max over the row quad first, then the sums (exp(-inf - m) == 0: no branches; a row without any visible key
    // in this warp has m == -inf and keeps l == 0)
    rm0 = fmaxf(rm0, __shfl_xor_sync(0xffffffffu, rm0, 1));
    rm0 = fmaxf(rm0, __shfl_xor_sync(0xffffffffu, rm0, 2));
    rm1 = fmaxf(rm1, __shfl_xor_sync(0xffffffffu, rm1, 1));
    rm1 = fmaxf(rm1, __shfl_xor_sync(0xffffffffu, rm1, 2));
    // (ex2.approx as in the streamed variant: the sums only feed the softmax denominator, whose relative error of ~2^-22
    // is far below the bf16 rounding of the probabilities; 16 inlined expf were 2.5 KB of cold straight-line code)
    const float b0 = rm0 == -INFINITY ? 0.f : -rm0 * ATT_LOG2E, b1 = rm1 == -INFINITY ? 0.f : -rm1 * ATT_LOG2E;
#pragma unroll
    for (int u = 0; u < ATT_RING_NG; ++u)
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        rl0 += att_ex2(fmaf(sreg[u][a][0], ATT_LOG2E, b0)) + att_ex2(fmaf(sreg[u][a][1], ATT_LOG2E, b0));
        rl1 += att_ex2(fmaf(sreg[u][a][2], ATT_LOG2E, b1)) + att_ex2(fmaf(sreg[u][a][3], ATT_LOG2E, b1));
      }
    rl0 += __shfl_xor_sync(0xffffffffu, rl0, 1);
    rl0 += __shfl_xor_sync(0xffffffffu, rl0, 2);
    rl1 += __shfl_xor_sync(0xffffffffu, rl1, 1);
    rl1 += __shfl_xor_sync(0xffffffffu, rl1, 2);
  } else {
#pragma unroll 1
    for (int i = sp; i < n; i += 2) {
      float acc[2][4];
      scores_of(i, acc);
      const int key0 = (g_lo + i) * ATT_GRAN;
      const bool tail = key0 + ATT_GRAN > n_ctx;
      if (tail) mask_scores(key0, acc);
      // park the scores
      const long at = park0 + (long)i * 128;
      if (EAGER) {
        uint4 pk;
        pk.x = pack_bf2_rn(acc[0][0], acc[0][1]); pk.y = pack_bf2_rn(acc[0][2], acc[0][3]);
        pk.z = pack_bf2_rn(acc[1][0], acc[1][1]); pk.w = pack_bf2_rn(acc[1][2], acc[1][3]);
        __stcg(reinterpret_cast<uint4*>(scores) + at, pk);
      } else {
        float4* dst = reinterpret_cast<float4*>(scores) + at * 2;
        __stcg(dst, make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]));
        __stcg(dst + 1, make_float4(acc[1][0], acc[1][1], acc[1][2], acc[1][3]));
      }
      // online (max, sum exp): the statistics only feed the softmax denominator (a sum over thousands of terms), so
      // ex2.approx is accurate enough here; the probabilities of pass 2 use expf.  exp(-inf) == 0 keeps it branch-free as
      // long as the running max is finite, which only a fully masked tail granule can violate.
      const float m0 = fmaxf(rm0, fmaxf(fmaxf(acc[0][0], acc[0][1]), fmaxf(acc[1][0], acc[1][1])));
      const float m1 = fmaxf(rm1, fmaxf(fmaxf(acc[0][2], acc[0][3]), fmaxf(acc[1][2], acc[1][3])));
      if (!tail || (m0 != -INFINITY && m1 != -INFINITY)) {
        const float b0 = -m0 * ATT_LOG2E, b1 = -m1 * ATT_LOG2E;
        rl0 = rl0 * att_ex2(fmaf(rm0, ATT_LOG2E, b0)) +
              ((att_ex2(fmaf(acc[0][0], ATT_LOG2E, b0)) + att_ex2(fmaf(acc[0][1], ATT_LOG2E, b0))) +
               (att_ex2(fmaf(acc[1][0], ATT_LOG2E, b0)) + att_ex2(fmaf(acc[1][1], ATT_LOG2E, b0))));
        rl1 = rl1 * att_ex2(fmaf(rm1, ATT_LOG2E, b1)) +
              ((att_ex2(fmaf(acc[0][2], ATT_LOG2E, b1)) + att_ex2(fmaf(acc[0][3], ATT_LOG2E, b1))) +
               (att_ex2(fmaf(acc[1][2], ATT_LOG2E, b1)) + att_ex2(fmaf(acc[1][3], ATT_LOG2E, b1))));
        rm0 = m0;
        rm1 = m1;
      } else {
        if (m0 != -INFINITY) {
          rl0 = rl0 * __expf(rm0 - m0) + ((__expf(acc[0][0] - m0) + __expf(acc[0][1] - m0)) + (__expf(acc[1][0] - m0) + __expf(acc[1][1] - m0)));
          rm0 = m0;
        }
        if (m1 != -INFINITY) {
          rl1 = rl1 * __expf(rm1 - m1) + ((__expf(acc[0][2] - m1) + __expf(acc[0][3] - m1)) + (__expf(acc[1][2] - m1) + __expf(acc[1][3] - m1)));
          rm1 = m1;
        }
      }
    }
    // combine the four lanes of a row
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      const float mo0 = __shfl_xor_sync(0xffffffffu, rm0, o), lo0 = __shfl_xor_sync(0xffffffffu, rl0, o);
      const float mo1 = __shfl_xor_sync(0xffffffffu, rm1, o), lo1 = __shfl_xor_sync(0xffffffffu, rl1, o);
      const float mn0 = fmaxf(rm0, mo0), mn1 = fmaxf(rm1, mo1);
      rl0 = (rm0 == -INFINITY ? 0.f : rl0 * att_exp_stat(rm0 - mn0)) + (mo0 == -INFINITY ? 0.f : lo0 * att_exp_stat(mo0 - mn0));
      rl1 = (rm1 == -INFINITY ? 0.f : rl1 * att_exp_stat(rm1 - mn1)) + (mo1 == -INFINITY ? 0.f : lo1 * att_exp_stat(mo1 - mn1));
      rm0 = mn0;
      rm1 = mn1;
    }
  }
  if (t4 == 0) {           // this warp's rows; the other query half reads as "no keys"
    sm.red[warp][q0] = rm0;  sm.red2[warp][q0] = rl0;
    sm.red[warp][q1] = rm1;  sm.red2[warp][q1] = rl1;
    sm.red[warp][q0 ^ 16] = -INFINITY;  sm.red2[warp][q0 ^ 16] = 0.f;
    sm.red[warp][q1 ^ 16] = -INFINITY;  sm.red2[warp][q1 ^ 16] = 0.f;
  }
  // ---- pass 2 mapping: the same (query half, key group) but ALL granules and one 64-dim half of the output per warp
  // (32 accumulator registers instead of 64: the 96-register budget of the program kernel does not hold a [16 x 128]
  // accumulator next to the probabilities).  Both dim-half warps of a (query half, key group) need the same probabilities:
  // the streamed variant re-reads them from the parked scores, the register variant passes them through the P tile.
  const int dh = sp;
  // parked scores of granule i, two granules ahead of their use (an L2 round trip is longer than one granule of work)
  struct Parked { uint4 a, b; };
  auto fetch = [&](int i) -> Parked {
    Parked r;
    const long at = park0 + (long)i * 128;
    if (EAGER) {
      r.a = __ldcg(reinterpret_cast<const uint4*>(scores) + at);
      r.b = make_uint4(0, 0, 0, 0);
    } else {
      r.a = __ldcg(reinterpret_cast<const uint4*>(scores) + at * 2);
      r.b = __ldcg(reinterpret_cast<const uint4*>(scores) + at * 2 + 1);
    }
    return r;
  };
  ACLK(3);
  sync();                  // the statistics of all warps; in the streamed variant also: every parked score is written
  Parked nx0, nx1;
  nx0.a = nx0.b = nx1.a = nx1.b = make_uint4(0, 0, 0, 0);
  if (STREAM) {
    if (n > 0) nx0 = fetch(0);
    if (n > 1) nx1 = fetch(1);
  }
  ACLK(4);
  att_ring_exchange_stats(p, sm, head, split, n_splits, head_cnt, epoch, tid, sync);
  ACLK(6);

  // ================= pass 2: O += P V =================
  const float gm0 = sm.gM[q0], gl0 = sm.gL[q0], gr0 = sm.gR[q0];
  const float gm1 = sm.gM[q1], gl1 = sm.gL[q1], gr1 = sm.gR[q1];
  // exp(-inf) == 0: masked keys and keys past kv_len need no branch (a fully masked ROW has gm == -inf and is NaN, as in
  // the reference's softmax)
  auto prob2 = [&](float sa, float sb, float gm, float gl, float gr) -> uint32_t {
    const float ea = expf(sa - gm), eb = expf(sb - gm);
    return EAGER ? pack_bf2_rn(att_div(ea, gl, gr), att_div(eb, gl, gr)) : pack_bf2_rn(ea, eb);
  };
  if (!STREAM) {
    // P in the reference's rounding into the P tile (over the Q tile: every warp is done with Q): row q, key 32 i + ...
#pragma unroll
    for (int u = 0; u < ATT_RING_NG; ++u) {
      const int i = sp + 2 * u;
      if (i < n) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int kx = i * ATT_GRAN + kcol + 8 * a;                 // even
          *reinterpret_cast<uint32_t*>(sm.q + att_swz256(q0, kx >> 3) + (kx & 7) * 2) = prob2(sreg[u][a][0], sreg[u][a][1], gm0, gl0, gr0);
          *reinterpret_cast<uint32_t*>(sm.q + att_swz256(q1, kx >> 3) + (kx & 7) * 2) = prob2(sreg[u][a][2], sreg[u][a][3], gm1, gl1, gr1);
        }
      }
    }
    sync();
  }
  float oacc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) oacc[a][e] = 0.f;
  // the only granule that can hold rows past the kv length is the last one of the cache
  const int i_zero = (kv_len % ATT_GRAN != 0 && g_lo + n == cut.n_gran) ? n - 1 : -1;
  const uint32_t v_off = (uint32_t)att_gran_off(v_row, dh * 8 + v_ch);   // + dn * 32 bytes XOR-free: see below
  auto pv = [&](int i, const Parked& sc) {
    uint32_t pa[4];
    if (STREAM) {
      if (EAGER) {
        pa[0] = prob2(bf_lo_f(sc.a.x), bf_hi_f(sc.a.x), gm0, gl0, gr0);
        pa[1] = prob2(bf_lo_f(sc.a.y), bf_hi_f(sc.a.y), gm1, gl1, gr1);
        pa[2] = prob2(bf_lo_f(sc.a.z), bf_hi_f(sc.a.z), gm0, gl0, gr0);
        pa[3] = prob2(bf_lo_f(sc.a.w), bf_hi_f(sc.a.w), gm1, gl1, gr1);
      } else {
        pa[0] = prob2(__uint_as_float(sc.a.x), __uint_as_float(sc.a.y), gm0, gl0, gr0);
        pa[1] = prob2(__uint_as_float(sc.a.z), __uint_as_float(sc.a.w), gm1, gl1, gr1);
        pa[2] = prob2(__uint_as_float(sc.b.x), __uint_as_float(sc.b.y), gm0, gl0, gr0);
        pa[3] = prob2(__uint_as_float(sc.b.z), __uint_as_float(sc.b.w), gm1, gl1, gr1);
      }
    } else {
      ldmatrix_x4(pa, sm.q + att_swz256(qa_row, i * 4 + kg * 2 + qa_ch));
    }
    const unsigned int f = fifo0 + (unsigned int)(n + i);
    const int slot = f & (ATT_SLOTS - 1);
    mbar_wait(&full[slot], (f >> 2) & 1u);
    if (i == 0) ACLK(8);
    const uint8_t* vt = sm.ring[slot];
    // rows past the kv length hold whatever the cache holds: their probabilities are zero, but 0 x NaN is not, so the V
    // fragments of those keys are cleared in registers (fragment register 0 / 2: keys 16 kg + 2 t4, +1; 1 / 3: +8)
    uint32_t vm0 = 0xffffffffu, vm1 = 0xffffffffu;
    if (i == i_zero) {
      const int r0 = kv_len - (g_lo + i) * ATT_GRAN, k = 16 * kg + 2 * t4;
      vm0 = (k < r0 ? 0x0000ffffu : 0u) | (k + 1 < r0 ? 0xffff0000u : 0u);
      vm1 = (k + 8 < r0 ? 0x0000ffffu : 0u) | (k + 9 < r0 ? 0xffff0000u : 0u);
    }
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      uint32_t bfr[4];
      // chunk = dh * 8 + dn * 2 + v_ch: bits 1..2 of the chunk index take dn, swizzled by XOR with the row bits
      ldmatrix_x4_trans(bfr, vt + (v_off ^ (uint32_t)(dn << 5)));
      if (i == i_zero) { bfr[0] &= vm0; bfr[1] &= vm1; bfr[2] &= vm0; bfr[3] &= vm1; }
      mma_bf16(oacc[2 * dn], pa, bfr[0], bfr[1]);
      mma_bf16(oacc[2 * dn + 1], pa, bfr[2], bfr[3]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  };
  if (!STREAM) {
#pragma unroll 1
    for (int i = 0; i < n; ++i) pv(i, nx0);
  } else {
#pragma unroll 1
    for (int i = 0; i < n; i += 2) {
      const Parked c0 = nx0;
      if (i + 2 < n) nx0 = fetch(i + 2);
      pv(i, c0);
      if (i + 1 < n) {
        const Parked c1 = nx1;
        if (i + 3 < n) nx1 = fetch(i + 3);
        pv(i + 1, c1);
      }
    }
  }

  float* const opart_mine = p.opart + ((long)head * n_splits + split) * VS_TOK * 128;
  // ================= team reduction over the two key groups (fixed order), partial O of the split to global =========
  ACLK(9);
  sync();                                                            // every warp is done with the V granules
  float4* rbuf = reinterpret_cast<float4*>(sm.ring[0]) + (mb * 2 + dh) * 256 + lane;   // 4 x 4 KB, thread-private columns
  if (kg == 1) {
#pragma unroll
    for (int a = 0; a < 8; ++a) rbuf[a * 32] = make_float4(oacc[a][0], oacc[a][1], oacc[a][2], oacc[a][3]);
  }
  sync();
  if (kg == 0) {
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const float4 v = rbuf[a * 32];
      const int d = dh * 64 + a * 8 + 2 * t4;
      *reinterpret_cast<float2*>(opart_mine + q0 * 128 + d) = make_float2(oacc[a][0] + v.x, oacc[a][1] + v.y);
      *reinterpret_cast<float2*>(opart_mine + q1 * 128 + d) = make_float2(oacc[a][2] + v.z, oacc[a][3] + v.w);
    }
  }
  ACLK(10);
  att_ring_reduce_output<EAGER>(p, sm, head, split, n_splits, max(1, (cut.n_gran + cut.cps - 1) / cut.cps), tq_rows, head_cnt,
                                epoch, tid, sync);
  ACLK(11);
  return 2u * (unsigned int)n;
}

}  // namespace vs
