// Worker-side phase bodies of the persistent program kernel (program.cu).  One team = 128 threads = one work item.
//
// Everything here is written for LATENCY: a phase is a few microseconds long and sits between two grid barriers, so
// every item issues all of its global loads up front (no load -> use -> load chains), items are cut fine enough that
// one round over the 2 x 148 teams covers a phase, and data written by other CTAs during this launch is read with
// ld.cg / volatile (L2), never through a possibly stale L1 line.
#pragma once
#include <math.h>

#include "attn_dev.cuh"
#include "program.h"
#include "sm100.cuh"

namespace vs {

// 20 warps per CTA: 4 role warps (weight producer, MMA issuer, activation producer, L2 prefetcher) + 16 worker warps.
// The non-GEMM phases are chains of dependent instructions on little data; with only 8 worker warps (2 per scheduler)
// they ran at a fraction of the issue rate, 4 per scheduler hide most of the ALU / shared-memory latency.
constexpr int PG_THREADS = 640;
constexpr int PG_WORKERS = 512;
constexpr int PG_TEAM = 128;          // small team: epilogue / norm / row / scan / splice items   (4 per CTA)
constexpr int PG_BIG_TEAM = 256;      // attention team (attn_dev.cuh)                            (2 per CTA)
constexpr int PG_TEAMS = PG_WORKERS / PG_TEAM;
constexpr int PG_BIG_TEAMS = PG_WORKERS / PG_BIG_TEAM;
constexpr int PG_BM = 128, PG_BK = 64;
constexpr int TILE_FLOATS = VS_TOK * PG_BM;      // one partial tile: [32 tokens][128 rows] fp32

struct TeamSync {
  int id;
  __device__ __forceinline__ void operator()() const { named_bar_sync(id, PG_TEAM); }
};
struct BigTeamSync {
  int id;
  __device__ __forceinline__ void operator()() const { named_bar_sync(id, PG_BIG_TEAM); }
};
struct WorkerSync {
  __device__ __forceinline__ void operator()() const { named_bar_sync(1, PG_WORKERS); }
};

// The (tile, k-chunk) sequence of a weight matrix is cut into one contiguous range per CTA on the HOST (program.cu):
//   tables[range_off + 2b .. +1] = [lo, hi) of CTA b,   tables[tab_off + 4*tile .. +3] = {first contributor, count, its
//   segment index, 0}; contributors of a tile are consecutive CTAs and every later one holds the tile in its segment 0.
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
__device__ __forceinline__ float pg_silu(float x) { return x / (1.0f + expf(-x)); }
// device-state ints written earlier in this launch: L2 load, no volatile (a volatile load is also a scheduling barrier)
__device__ __forceinline__ int ld_vol(const int* p) { return __ldcg(p); }

// ---------------------------------------------------------------------------------------------------------------
// grid barrier (workers only): CTA barrier -> one thread: release-add, acquire-spin -> CTA barrier.
// The release is cumulative over the writes the other worker threads made before the CTA barrier.
// ---------------------------------------------------------------------------------------------------------------
#ifndef VS_BARRIER_ACQUIRE
#define VS_BARRIER_ACQUIRE 0
#endif
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target, int wtid) {
  named_bar_sync(1, PG_WORKERS);
  if (wtid == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    while (ld_relaxed_u32(ctr) < target) {
    }
#if VS_BARRIER_ACQUIRE
    (void)ld_acquire_u32(ctr);     // one acquire (and one L1 invalidation) per barrier, after the relaxed spin
#endif
  }
  named_bar_sync(1, PG_WORKERS);
}

// ---------------------------------------------------------------------------------------------------------------
// PH_EPI: item = (weight-row tile, group of 8 tokens).  ONE code path for all epilogues (the kernel's code must stay
// small enough to live in the instruction cache: every phase runs for a few microseconds only): thread = (column
// c < 64 paired with column c + 64, half of the token group), so each thread reduces 2 columns x 4 tokens.
// ---------------------------------------------------------------------------------------------------------------
// Optional checkpoint clocks (compile with -DVS_PROGRAM_CLK): thread 0 of team 0 on CTA 0 records clock64() at a few
// points of its first work item of a phase; tools read them back with vs_debug_clk().
#ifdef VS_PROGRAM_CLK
#define CLK_DECL const bool clk_on = blockIdx.x == 0 && tt == 0 && sync.id == 2
#define CLK(i) do { if (clk_on) g_clk[16 + i] = clock64(); } while (0)
#else
#define CLK_DECL
#define CLK(i)
#endif

// Tokens per thread of an epilogue item (item = 2 * EPI_TPT token rows): groups of 8 rows -- 30 tree tokens are 4 groups,
// and the gate|up epilogue of a 7B layer has 172 x 4 = 688 items for 592 teams (two rounds on 96 teams).
constexpr int EPI_TPT = 4;              // tokens per thread
constexpr int EPI_TPI = 2 * EPI_TPT;    // tokens per item
constexpr int EPI_BATCH = 4;        // contributors whose partial tiles are loaded together (32 loads in flight per thread)
constexpr int EPI_BATCH_WIDE = 6;   // ... for tiles with more than EPI_BATCH contributors (48 loads in flight)

// q = k / d for k * d < 2^32, with magic = floor(2^32 / d) + 1 precomputed on the host (d == 1: magic == 0 -> q = k)
__device__ __forceinline__ int fast_div(int k, unsigned int magic) {
  return magic ? (int)__umulhi((unsigned int)k, magic) : k;
}
__device__ __forceinline__ float bf_bits(unsigned short u) { return __uint_as_float(((unsigned int)u) << 16); }

// a[j][i] += sum over the contributors b0 .. b0 + n_c - 1 (k order) of their partial tile's element (token i, column j * 64):
// `pbase` = partials + this thread's (first token, column) offset; B loads per element in flight
template <int B, int TPT>
__device__ __forceinline__ void epi_reduce(float (&a)[2][TPT], const float* pbase, int b0, int n_c, int seg0, int maxseg) {
#pragma unroll 1
  for (int c0 = 0; c0 < n_c; c0 += B) {
    float v[B][2][TPT];
#pragma unroll
    for (int c = 0; c < B; ++c) {
      const int ci = min(c0 + c, n_c - 1);                                   // clamped: always a valid tile
      const float* base = pbase + ((long)(b0 + ci) * maxseg + (ci == 0 ? seg0 : 0)) * TILE_FLOATS;
#pragma unroll
      for (int i = 0; i < TPT; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) v[c][j][i] = __ldcg(base + i * PG_BM + 64 * j);   // rows >= T hold stale finite data
    }
#pragma unroll
    for (int c = 0; c < B; ++c) {
#pragma unroll
      for (int i = 0; i < TPT; ++i) {
        a[0][i] += c0 + c < n_c ? v[c][0][i] : 0.f;
        a[1][i] += c0 + c < n_c ? v[c][1][i] : 0.f;
      }
    }
  }
}

// Everything an epilogue item needs besides the partial tiles; loaded up front, before the item waits for its contributors.
template <int TPT>
struct EpiOps {
  unsigned short rraw[2][TPT];       // residual rows (RESID)
  bf16 csraw[4][TPT];                // cos / sin of the two column halves (HEADS + RoPE)
  float bias_e[2], nwf[2];               // bias and folded-RMSNorm weight of the thread's two columns
  int row_base;                          // KV append row (HEADS)
};
__device__ __forceinline__ int epi_tile_of(int item, int groups) {
  return groups == 1 ? item : (groups == 2 ? item >> 1 : (groups == 4 ? item >> 2 : item / 3));
}
// Indices are clamped instead of predicated so that the loads are plain back-to-back instructions.
template <int TPT, int KIND>
__device__ __forceinline__ void epi_load_ops(const Phase& ph, int item, int tt, EpiOps<TPT>& o) {
  const int T = ph.T, N = ph.N;
  constexpr int epi = KIND;
  const int groups = (T + (2 * TPT) - 1) / (2 * TPT);
  const int col = tt & 63, half = tt >> 6;
  const int tile = epi_tile_of(item, groups);
  const int grp = item - tile * groups;
  const int ts = grp * (2 * TPT) + half * TPT;     // first token of this thread
  const bool resid = epi == EPI_RESID, heads = epi == EPI_HEADS;
  const bool is_q = tile < ph.n_q_tiles;
  const bool is_k = !is_q && tile < ph.n_q_tiles + ph.n_k_tiles;
  const bool rope = heads && ph.rope && (is_q || is_k);
  if (resid) {
    const bf16* R = ph.R;
    const long ldr = ph.ldr;
#pragma unroll
    for (int i = 0; i < TPT; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        o.rraw[j][i] = __ldcg(reinterpret_cast<const unsigned short*>(R + (long)min(ts + i, T - 1) * ldr +
                                                                      min(tile * PG_BM + col + 64 * j, N - 1)));
  }
  if (rope) {
    int pos[TPT];
#pragma unroll
    for (int i = 0; i < TPT; ++i) pos[i] = __ldcg(ph.pos_ids + min(ts + i, T - 1));
#pragma unroll
    for (int i = 0; i < TPT; ++i) {
      o.csraw[0][i] = ph.cos[(long)pos[i] * 128 + col];      o.csraw[1][i] = ph.sin[(long)pos[i] * 128 + col];
      o.csraw[2][i] = ph.cos[(long)pos[i] * 128 + 64 + col]; o.csraw[3][i] = ph.sin[(long)pos[i] * 128 + 64 + col];
    }
  }
  // bias and (folded RMSNorm) weight of this thread's two columns, and the KV append row: with the other operands, not
  // behind the reduction (each is an L2 round trip of its own)
  {
    const bf16* bp = ph.bias;
    const bf16* nwp = (!heads && epi != EPI_SWIGLU) ? ph.w : nullptr;
    unsigned short braw[2] = {0, 0}, nraw[2] = {0, 0};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = min(tile * PG_BM + col + 64 * j, N - 1);
      if (bp) braw[j] = __ldg(reinterpret_cast<const unsigned short*>(bp + n));
      if (nwp) nraw[j] = __ldg(reinterpret_cast<const unsigned short*>(nwp + n));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) { o.bias_e[j] = bf_bits(braw[j]); o.nwf[j] = bf_bits(nraw[j]); }
  }
  o.row_base = heads ? (ph.kv_row_base ? __ldcg(ph.kv_row_base) : 0) + ph.kv_row_off : 0;
}
// TPT tokens per thread, items of 2 * TPT token rows: 4 by default; 8 (`ph.tpi` = 16) for the epilogues whose item count
// at 4 would not fit one round over the teams (gate|up: 172 tiles x 4 groups = 688 items on 592 teams; LM heads) -- an item
// is a latency chain, twice the rows per thread cost far less than a second round.
// KIND: the epilogue (GemmEpilogue) as a compile-time constant -- one compact function per kind instead of one function
// with all four bodies interleaved (every phase runs its code cold: what is fetched should be what is executed)
template <int TPT, int KIND>
__device__ __forceinline__ void epi_phase_t(const Phase& ph, unsigned int epoch, int gteam, int n_teams, int tt, float* red,
                                            TeamSync sync) {
  // everything an item needs from the descriptor, once (the descriptor lives in shared memory)
  const int T = ph.T, N = ph.N, nk = ph.nk, maxseg = ph.maxseg;
  constexpr int epi = KIND;
  const int per_lo = ph.per_lo, rem = ph.rem;
  const unsigned int m_hi = ph.magic_per_hi, m_lo = ph.magic_per_lo, m_nk = ph.magic_nk;
  const float* partials = ph.partials;
  const unsigned int* tile_cnt = ph.tile_cnt;
  const int groups = (T + (2 * TPT) - 1) / (2 * TPT);      // 1..4
  const int n_items = ph.n_tiles * groups;
  const int col = tt & 63, half = tt >> 6;
  const int warp = tt >> 5, lane = tt & 31;
  const bool resid = epi == EPI_RESID, heads = epi == EPI_HEADS;
  const int big = rem * (per_lo + 1);                  // chunks owned by the first `rem` CTAs (one extra chunk each)
  CLK_DECL;
  CLK(0);
  for (int item = gteam; item < n_items; item += n_teams) {
    CLK(1);
    const int tile = epi_tile_of(item, groups);
    const int grp = item - tile * groups;
    const int ts = grp * (2 * TPT) + half * TPT;     // first token of this thread
    // contributors of the tile: consecutive CTAs b0 .. b0+n-1; closed form of the balanced partition of program.cu
    const int k0 = tile * nk, k1 = k0 + nk - 1;
    const int b0 = k0 < big ? fast_div(k0, m_hi) : rem + fast_div(k0 - big, m_lo);
    const int b1 = k1 < big ? fast_div(k1, m_hi) : rem + fast_div(k1 - big, m_lo);
    const int n_c = b1 - b0 + 1;
    const int seg0 = tile - fast_div(b0 * per_lo + min(b0, rem), m_nk);

    // ---- all operand loads that do not depend on the reduction, issued before anything is consumed
    EpiOps<TPT> ops;
    epi_load_ops<TPT, KIND>(ph, item, tt, ops);
    const bool is_q = tile < ph.n_q_tiles;
    const bool is_k = !is_q && tile < ph.n_q_tiles + ph.n_k_tiles;
    const bool rope = heads && ph.rope && (is_q || is_k);
    const unsigned short (&rraw)[2][TPT] = ops.rraw;
    const bf16 (&csraw)[4][TPT] = ops.csraw;
    const float (&bias_e)[2] = ops.bias_e;
    const float (&nwf)[2] = ops.nwf;
    const int row_base_e = ops.row_base;
    CLK(2);
    // ---- fused GEMM+epilogue: no grid barrier separated us from the GEMM phase; wait until every contributor of this
    // tile has published its partial tile (monotonic counter: n_c arrivals per launch)
    if (tile_cnt) {
      if (tt == 0) {
        const unsigned int want = (epoch + 1u) * (unsigned int)n_c;
        while (ld_relaxed_u32(tile_cnt + tile) < want) {
        }
      }
      sync();
    }
    CLK(7);
    // ---- reduction over the contributors in k order; the loads of B contributors are in flight together.  The residual
    // epilogues (o_proj, down_proj: 5-6 contributors per tile on 148 CTAs) take the wider batch: one L2 round trip
    // instead of two.  Same order of additions either way.  (Compiling only ONE of the two variants into each function was
    // measured slower: 215.5 vs 218.4 tok/s.)
    float a[2][TPT];
#pragma unroll
    for (int i = 0; i < TPT; ++i) { a[0][i] = 0.f; a[1][i] = 0.f; }
    const float* const pbase = partials + ts * PG_BM + col;
    if (n_c > EPI_BATCH) epi_reduce<EPI_BATCH_WIDE, TPT>(a, pbase, b0, n_c, seg0, maxseg);
    else epi_reduce<EPI_BATCH, TPT>(a, pbase, b0, n_c, seg0, maxseg);

    CLK(3);
#ifdef VS_PROGRAM_CLK
    if (clk_on) g_clk[20] = (long long)(a[0][0] + a[1][3]);       // consume: the loads have landed
#endif
    CLK(4);
    if (epi == EPI_SWIGLU) {
      // tile rows [0,64) = gate features j0..j0+63, rows [64,128) = the matching up features
      bf16* Y = ph.Y;
      const long ldy = ph.ldy;
#pragma unroll
      for (int i = 0; i < TPT; ++i)
        if (ts + i < T) {
          const float g = rbf(pg_silu(rbf(a[0][i])));                  // act_fn(gate) rounds, then * up rounds
          Y[(long)(ts + i) * ldy + tile * 64 + col] = f2bf(g * rbf(a[1][i]));
        }
    } else if (heads) {
      // one 128-row tile == one head of q | k | v; optional RoPE; q -> Qout, k/v -> KV rows in place
      const int row_base = row_base_e;
      const float b1 = bias_e[0], b2 = bias_e[1];
      const int head = is_q ? tile : (is_k ? tile - ph.n_q_tiles : tile - ph.n_q_tiles - ph.n_k_tiles);
      bf16* dst = is_q ? ph.Qout + (long)head * ph.q_head_stride
                       : (is_k ? ph.Kout : ph.Vout) + (long)head * ph.kv_head_stride + (long)row_base * 128;
#pragma unroll
      for (int i = 0; i < TPT; ++i)
        if (ts + i < T) {
          const float x1 = rbf(a[0][i] + b1), x2 = rbf(a[1][i] + b2);
          float o1 = x1, o2 = x2;
          if (rope) {   // rotate-half RoPE with three bf16 roundings (cnets_ours.py:117-118)
            o1 = rbf(x1 * bf2f(csraw[0][i])) + rbf(-x2 * bf2f(csraw[1][i]));
            o2 = rbf(x2 * bf2f(csraw[2][i])) + rbf(x1 * bf2f(csraw[3][i]));
          }
          dst[(long)(ts + i) * 128 + col] = f2bf(o1);
          dst[(long)(ts + i) * 128 + 64 + col] = f2bf(o2);
        }
    } else {
      // STORE / RESID (+ per-tile sums of squares of the new rows for the RMSNorm that follows)
      bool ok[2];
      bf16* Y = ph.Y;
      const long ldy = ph.ldy;
      float* ssq = ph.ssq;
      const bf16* nw = ph.w;                                           // folded-in RMSNorm (prog_add_norm) or null
      float yv[2][TPT];
#pragma unroll
      for (int j = 0; j < 2; ++j) ok[j] = tile * PG_BM + col + 64 * j < N;
      if (ssq) sync();                                                 // previous item's readers of `red` are done
#pragma unroll
      for (int i = 0; i < TPT; ++i) {
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float v = a[j][i] + bias_e[j];
          if (resid) v = bf_bits(rraw[j][i]) + rbf(v);                 // Linear output rounds to bf16, then the add rounds again
          const bf16 y = f2bf(v);
          yv[j][i] = bf2f(y);
          if (ok[j] && ts + i < T) {
            Y[(long)(ts + i) * ldy + tile * PG_BM + col + 64 * j] = y;
            q += bf2f(y) * bf2f(y);
          }
        }
        if (ssq) {
          q = warp_sum(q);
          if (lane == 0) red[warp * TPT + i] = q;
        }
      }
      // A folded-in RMSNorm exchanges the sums of squares as tagged 8-byte words (value + (launch, phase) tag in one
      // single-copy-atomic store, behind the float entries): the readers spin on exactly the words they need -- one L2 round
      // trip instead of store -> barrier -> release-increment -> spin -> barrier -> load
      const int n_tiles = ph.n_tiles;
      unsigned long long* ssq_ll = reinterpret_cast<unsigned long long*>(ssq + (long)n_tiles * VS_TOK);
      const unsigned int ll_tag = ((epoch + 1u) << 16) | ((unsigned int)ph.phase_idx & 0xffffu);
      if (ssq) {
        sync();
        if (tt < (2 * TPT)) {                                            // token tt of the group: warps 2*h, 2*h+1 hold its half
          const int h = tt / TPT, i = tt % TPT;
          const int t = grp * (2 * TPT) + tt;
          if (t < T) {
            const float q = red[(2 * h) * TPT + i] + red[(2 * h + 1) * TPT + i];
            ssq[tile * VS_TOK + t] = q;
            if (nw) {
              const unsigned long long w = ((unsigned long long)ll_tag << 32) | __float_as_uint(q);
              asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ssq_ll + tile * VS_TOK + t), "l"(w) : "memory");
            }
          }
        }
      }
      CLK(8);
      if (TPT == EPI_TPT && nw) {      // (the builder folds a norm only into epilogues of 8-row items)
        // LlamaRMSNorm of the finished rows (same arithmetic and summation order as norm_phase): wait for the sums of
        // squares of every tile of the token group, normalise the values still in registers
#pragma unroll
        for (int q = 0; q < ((2 * TPT) + 3) / 4; ++q) {                  // warp w: tokens 2w, 2w+1 of the group (3w.. for 10)
          const int tg = warp * (((2 * TPT) + 3) / 4) + q;
          const int t = grp * (2 * TPT) + tg;
          float ss = 0.f;
          if (tg < (2 * TPT) && t < T)
            for (int i = lane; i < n_tiles; i += 32) {
              unsigned long long w;
              do {
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(ssq_ll + i * VS_TOK + t) : "memory");
              } while ((unsigned int)(w >> 32) != ll_tag);
              ss += __uint_as_float((unsigned int)w);
            }
          ss = warp_sum(ss);
          if (lane == 0 && tg < (2 * TPT)) red[32 + tg] = 1.0f / sqrtf(ss / (float)ph.H + ph.eps);
        }
        sync();
        CLK(9);
        bf16* X = ph.dst;
        const long ldd = ph.ldd;
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
          const float rstd = red[32 + half * TPT + i];
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (ok[j] && ts + i < T)                                   // weight * bf16(x * rstd): two bf16 roundings
              X[(long)(ts + i) * ldd + tile * PG_BM + col + 64 * j] = f2bf(nwf[j] * rbf(yv[j][i] * rstd));
        }
      }
    }
    CLK(5);
  }
  CLK(6);
}
// (-DVS_EPI_INLINE=mask, A/B: bit 0 inlines the residual epilogue into the kernel body, bit 1 the SwiGLU ones, bit 2 the
// q|k|v one -- inside a noinline function every global load carries two R2UR of the memory descriptor)
#ifndef VS_EPI_INLINE
#define VS_EPI_INLINE 0
#endif
#define VS_EPI_FN(name, TPT_, KIND_, ATTR_)                                                                            \
  __device__ ATTR_ void name(const Phase& ph, unsigned int epoch, int gteam, int n_teams, int tt, float* red,         \
                             TeamSync sync) {                                                                         \
    epi_phase_t<TPT_, KIND_>(ph, epoch, gteam, n_teams, tt, red, sync);                                               \
  }
#define VS_EPI_ATTR(bit) VS_EPI_ATTR_##bit
#if VS_EPI_INLINE & 1
#define VS_EPI_ATTR_0 __forceinline__
#else
#define VS_EPI_ATTR_0 __noinline__
#endif
#if VS_EPI_INLINE & 2
#define VS_EPI_ATTR_1 __forceinline__
#else
#define VS_EPI_ATTR_1 __noinline__
#endif
#if VS_EPI_INLINE & 4
#define VS_EPI_ATTR_2 __forceinline__
#else
#define VS_EPI_ATTR_2 __noinline__
#endif
VS_EPI_FN(epi_phase_store, EPI_TPT, EPI_STORE, __noinline__)
VS_EPI_FN(epi_phase_resid, EPI_TPT, EPI_RESID, VS_EPI_ATTR(0))
VS_EPI_FN(epi_phase_swiglu, EPI_TPT, EPI_SWIGLU, VS_EPI_ATTR(1))
VS_EPI_FN(epi_phase_heads, EPI_TPT, EPI_HEADS, VS_EPI_ATTR(2))
VS_EPI_FN(epi_phase_store_wide, 2 * EPI_TPT, EPI_STORE, __noinline__)
VS_EPI_FN(epi_phase_swiglu_wide, 2 * EPI_TPT, EPI_SWIGLU, VS_EPI_ATTR(1))
#undef VS_EPI_FN
__device__ __forceinline__ void epi_phase(const Phase& ph, unsigned int epoch, int gteam, int n_teams, int tt, float* red,
                                          TeamSync sync) {
  const bool wide = ph.tpi == 2 * EPI_TPI;
  switch (ph.epi) {
    case EPI_RESID: epi_phase_resid(ph, epoch, gteam, n_teams, tt, red, sync); break;
    case EPI_SWIGLU:
      if (wide) epi_phase_swiglu_wide(ph, epoch, gteam, n_teams, tt, red, sync);
      else epi_phase_swiglu(ph, epoch, gteam, n_teams, tt, red, sync);
      break;
    case EPI_HEADS: epi_phase_heads(ph, epoch, gteam, n_teams, tt, red, sync); break;
    default:
      if (wide) epi_phase_store_wide(ph, epoch, gteam, n_teams, tt, red, sync);
      else epi_phase_store(ph, epoch, gteam, n_teams, tt, red, sync);
      break;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PH_GEMM2 epilogue (slab.h): the 16 worker warps finish this CTA's output rows straight from TMEM -- nothing is
// exchanged with other CTAs except, for a residual epilogue, one row of sums of squares for the RMSNorm that follows.
// warp = (lane quarter q of the accumulators: tile rows 32q .. 32q+31, token group cg: columns 8cg .. 8cg+7);
// thread = (tile row 32q + lane, 8 tokens) of both accumulators of the pass: the two members of a gate|up or rotate-half
// pair sit in the same lane of accumulator 0 and 1.
// ---------------------------------------------------------------------------------------------------------------
struct SlabRed {
  float part[4][VS_TOK];     // per lane quarter: sums of squares of its rows, per token
  float rstd[VS_TOK];
};

// (the accumulator cursor goes in and out BY VALUE, packed acc | acc_par << 1: a reference into a noinline function pins
// the caller's variable to local memory for the whole kernel)
__device__ __noinline__ int slab_epilogue(const Phase& ph, unsigned int epoch, int b, int G, int wtid, uint32_t tmem_base,
                                          int acc_state, uint64_t* tfull, uint64_t* tempty, SlabRed& red,
                                          bool backoff = false) {
  int acc = acc_state & 1, acc_par = acc_state >> 1;
  const int wwarp = wtid >> 5, lane = wtid & 31, quarter = wwarp & 3, cg = wwarp >> 2;
  const int T = ph.T, epi = ph.epi;
  const int passes = ph.s_passes, n_tiles = ph.s_ntiles, rows_alloc = ph.s_rows_alloc, pair = ph.s_pair;
  const int e_per = ph.s_eper, e_rem = ph.s_erem;
  // accumulator lane -> row of the tile: M = 128 keeps row r in lane r; M = 64 keeps it in lane (r & 15) + 32 * (r >> 4)
  const int row = ph.s_m64 ? (lane < 16 ? quarter * 16 + lane : 1 << 20) : quarter * 32 + lane;
  const int t0 = cg * 8;                               // first token of this thread
  const bool resid = epi == EPI_RESID, heads = epi == EPI_HEADS;
  float* ssq = ph.ssq;
  const bf16* nw = ph.w;                               // folded-in RMSNorm (prog_add_norm) or null; then passes == 1
  float ssq_acc = 0.f;                                 // threads wtid < 32: token wtid, summed over the passes
  float yv[2][8];                                      // the new rows of the (single) pass, for the folded norm
  int n_keep[2] = {-1, -1};
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) yv[j][i] = 0.f;

#pragma unroll 1
  for (int ps = 0; ps < passes; ++ps) {
    const int u = b * passes + ps;
    const int e_lo = u * e_per + min(u, e_rem), cnt = e_per + (u < e_rem ? 1 : 0);
    if (cnt == 0) continue;                            // (uniform over the CTA: no stages were streamed either)
    // output feature of this thread in accumulator 0 / 1, -1 = this lane holds no row
    int n[2] = {-1, -1};
    if (pair) {
      if (row < cnt) {
        const int e = e_lo + row;
        n[0] = (e >> 6) * 128 + (e & 63);
        n[1] = n[0] + 64;
      }
    } else if (row < rows_alloc) {
      if (row < cnt) n[0] = e_lo + row;
      if (n_tiles == 2 && rows_alloc + row < cnt) n[1] = e_lo + rows_alloc + row;
    }
    // ---- operand loads that do not depend on the accumulators: in flight while the weight stream finishes
    float bias[2] = {0.f, 0.f};
    if (ph.bias) {
#pragma unroll
      for (int j = 0; j < 2; ++j) bias[j] = bf2f(ph.bias[max(n[j], 0)]);
    }
    unsigned short rraw[2][8];
    if (resid) {
      const bf16* R = ph.R;
      const long ldr = ph.ldr;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          rraw[j][i] = __ldcg(reinterpret_cast<const unsigned short*>(R + (long)min(t0 + i, T - 1) * ldr + max(n[j], 0)));
    }
    const int e_me = e_lo + min(row, cnt - 1);         // pair element of this lane (clamped)
    const int hh = e_me >> 6, d = e_me & 63;           // HEADS: head (over q | k | v) and dim < 64
    const bool is_q = hh < ph.n_q_tiles, is_k = !is_q && hh < ph.n_q_tiles + ph.n_k_tiles;
    const bool rope = heads && ph.rope && (is_q || is_k);
    bf16 cs[4][8];
    if (heads && ph.rope) {                             // (uniform branch; lanes of v heads load and ignore)
      int pos[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pos[i] = __ldcg(ph.pos_ids + min(t0 + i, T - 1));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        cs[0][i] = ph.cos[(long)pos[i] * 128 + d];      cs[1][i] = ph.sin[(long)pos[i] * 128 + d];
        cs[2][i] = ph.cos[(long)pos[i] * 128 + 64 + d]; cs[3][i] = ph.sin[(long)pos[i] * 128 + 64 + d];
      }
    }
    // ---- the accumulators of this pass
    if (backoff) mbar_wait_sleep(&tfull[acc], acc_par, 200);
    else mbar_wait(&tfull[acc], acc_par);
    tc_fence_after();
    float a[2][8];
    tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 64 + t0, a[0]);
    if (n_tiles == 2) tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 64 + 32 + t0, a[1]);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[1][i] = 0.f;
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty[acc]);
    acc ^= 1;
    if (acc == 0) acc_par ^= 1;

    if (epi == EPI_SWIGLU) {
      if (n[0] >= 0) {
        bf16* Y = ph.Y + (e_lo + row);                  // feature index = pair element
        const long ldy = ph.ldy;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (t0 + i < T) {
            const float g = rbf(pg_silu(rbf(a[0][i])));               // act_fn(gate) rounds, then * up rounds
            Y[(long)(t0 + i) * ldy] = f2bf(g * rbf(a[1][i]));
          }
      }
    } else if (heads) {
      if (n[0] >= 0) {
        const int row_base = (ph.kv_row_base ? __ldcg(ph.kv_row_base) : 0) + ph.kv_row_off;
        const int head = is_q ? hh : (is_k ? hh - ph.n_q_tiles : hh - ph.n_q_tiles - ph.n_k_tiles);
        bf16* dst = is_q ? ph.Qout + (long)head * ph.q_head_stride
                         : (is_k ? ph.Kout : ph.Vout) + (long)head * ph.kv_head_stride + (long)row_base * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (t0 + i < T) {
            const float x1 = rbf(a[0][i] + bias[0]), x2 = rbf(a[1][i] + bias[1]);
            float o1 = x1, o2 = x2;
            if (rope) {   // rotate-half RoPE with three bf16 roundings (cnets_ours.py:117-118)
              o1 = rbf(x1 * bf2f(cs[0][i])) + rbf(-x2 * bf2f(cs[1][i]));
              o2 = rbf(x2 * bf2f(cs[2][i])) + rbf(x1 * bf2f(cs[3][i]));
            }
            dst[(long)(t0 + i) * 128 + d] = f2bf(o1);
            dst[(long)(t0 + i) * 128 + 64 + d] = f2bf(o2);
          }
      }
    } else {
      // STORE / RESID (+ this CTA's sums of squares of the new rows for the RMSNorm that follows)
      bf16* Y = ph.Y;
      const long ldy = ph.ldy;
      float q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (n[j] < 0) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float v = a[j][i] + bias[j];
          if (resid) v = bf_bits(rraw[j][i]) + rbf(v);                // Linear output rounds to bf16, then the add rounds again
          const bf16 y = f2bf(v);
          yv[j][i] = bf2f(y);
          if (t0 + i < T) {
            Y[(long)(t0 + i) * ldy + n[j]] = y;
            q[i] += bf2f(y) * bf2f(y);
          }
        }
      }
      if (ssq) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s = warp_sum(q[i]);
          if (lane == 0) red.part[quarter][t0 + i] = s;
        }
        named_bar_sync(1, PG_WORKERS);
        if (wtid < VS_TOK) ssq_acc += (red.part[0][wtid] + red.part[1][wtid]) + (red.part[2][wtid] + red.part[3][wtid]);
        named_bar_sync(1, PG_WORKERS);                                 // `part` is free for the next pass
      }
      n_keep[0] = n[0];
      n_keep[1] = n[1];
    }
  }
  if (ssq) {
    if (wtid < VS_TOK) ssq[b * VS_TOK + wtid] = ssq_acc;              // [grid][32]; tokens >= T hold 0
  }
  if (nw) {
    // LlamaRMSNorm of the finished rows: publish this CTA's sums of squares, wait until every CTA of the grid has, then
    // normalise the values still held in registers (y = weight * bf16(x * rstd): two bf16 roundings, cnets_ours.py:523-527)
    float wj[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) wj[j] = bf2f(nw[max(n_keep[j], 0)]);   // in flight during the exchange
    named_bar_sync(1, PG_WORKERS);
    if (wtid == 0) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ph.norm_cnt) : "memory");
      const unsigned int want = (epoch + 1u) * (unsigned int)G;
      while (ld_relaxed_u32(ph.norm_cnt) < want) {
      }
    }
    named_bar_sync(1, PG_WORKERS);
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {                                   // warp w: tokens 2w, 2w+1
      const int t = 2 * wwarp + qq;
      float ss = 0.f;
      for (int bb = lane; bb < G; bb += 32) ss += __ldcg(ssq + bb * VS_TOK + t);   // fixed order
      ss = warp_sum(ss);
      if (lane == 0) red.rstd[t] = 1.0f / sqrtf(ss / (float)ph.H + ph.eps);
    }
    named_bar_sync(1, PG_WORKERS);
    bf16* X = ph.dst;
    const long ldd = ph.ldd;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (n_keep[j] < 0) continue;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (t0 + i < T) X[(long)(t0 + i) * ldd + n_keep[j]] = f2bf(wj[j] * rbf(yv[j][i] * red.rstd[t0 + i]));
    }
  }
  return acc | (acc_par << 1);
}

// ---------------------------------------------------------------------------------------------------------------
// PH_NORM: item = (token row, quarter of the row).  y = weight * bf16(x * rsqrt(mean(x^2) + eps)).
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void norm_phase(const Phase& ph, int gteam, int n_teams, int tt, float* red, TeamSync sync) {
  const int H = ph.H;
  const int n_items = ph.T * 4;
  const int lane = tt & 31;
  const int per_part = H / 32;                   // 16-byte vectors per quarter row
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item >> 2, part = item & 3;
    const bf16* x = ph.src + (long)t * ph.lds;
    bf16* y = ph.dst + (long)t * ph.ldd;
    // first vector of this thread: in flight while the row statistics are gathered
    const bool in0 = tt < per_part;
    const int e0 = (part * per_part + (in0 ? tt : 0)) * 8;
    uint4 x0 = in0 ? __ldcg(reinterpret_cast<const uint4*>(x + e0)) : make_uint4(0, 0, 0, 0);
    uint4 w0 = in0 ? *reinterpret_cast<const uint4*>(ph.w + e0) : make_uint4(0, 0, 0, 0);
    float ss = 0.f;
    if (ph.ssq_in) {
      // per-tile sums of squares left by the residual epilogue: lane i sums tiles i, i+32, ... (fixed order)
      for (int i = lane; i < ph.ssq_tiles; i += 32) ss += __ldcg(ph.ssq_in + i * VS_TOK + t);
      ss = warp_sum(ss);
    } else {
#pragma unroll 1
      for (int i = tt * 8; i < H; i += PG_TEAM * 8) {
        const uint4 v = __ldcg(reinterpret_cast<const uint4*>(x + i));
        const unsigned int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) ss += bf_lo(w[j]) * bf_lo(w[j]) + bf_hi(w[j]) * bf_hi(w[j]);
      }
      ss = warp_sum(ss);
      sync();
      if (lane == 0) red[tt >> 5] = ss;
      sync();
      ss = red[0] + red[1] + red[2] + red[3];
    }
    const float rstd = 1.0f / sqrtf(ss / (float)H + ph.eps);
#pragma unroll 1
    for (int v = tt; v < per_part; v += PG_TEAM) {
      const int e = (part * per_part + v) * 8;
      uint4 xq = x0, wq = w0;
      if (v != tt) {
        xq = __ldcg(reinterpret_cast<const uint4*>(x + e));
        wq = *reinterpret_cast<const uint4*>(ph.w + e);
      }
      const unsigned int xi[4] = {xq.x, xq.y, xq.z, xq.w};
      const unsigned int wi[4] = {wq.x, wq.y, wq.z, wq.w};
      unsigned int o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)   // weight * bf16(x * rstd): two bf16 roundings (cnets_ours.py:523-527)
        o[q] = pack_bf2(bf_lo(wi[q]) * rbf(bf_lo(xi[q]) * rstd), bf_hi(wi[q]) * rbf(bf_hi(xi[q]) * rstd));
      *reinterpret_cast<uint4*>(y + e) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PH_ROWS: item = row.  dst[t] = src[idx ? idx[t] : t]  (or src row 0 for a broadcast)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rows_phase(const Phase& ph, int gteam, int n_teams, int tt) {
  for (int t = gteam; t < ph.T; t += n_teams) {
    int r = ph.broadcast ? 0 : (ph.idx ? ld_vol(ph.idx + t) : t);
    if (ph.n_src_rows > 0) r = min(max(r, 0), ph.n_src_rows - 1);
    const uint4* s = reinterpret_cast<const uint4*>(ph.src + (long)r * ph.lds);
    uint4* d = reinterpret_cast<uint4*>(ph.dst + (long)t * ph.ldd);
    const int nv = ph.H / 8;
#pragma unroll 1
    for (int i0 = tt; i0 < nv; i0 += 4 * PG_TEAM) {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j * PG_TEAM < nv) v[j] = __ldcg(s + i0 + j * PG_TEAM);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + j * PG_TEAM < nv) d[i0 + j * PG_TEAM] = v[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// vocabulary scans: item = (row, slice of the vocabulary) on an attention-sized team of 256 threads; the slice is staged
// in shared memory once and scanned by rolled loops (small code); the slices of a row are merged on CTA 0 (tree phase)
// ---------------------------------------------------------------------------------------------------------------
constexpr int SCAN_MAX = 8192;     // elements per slice (16 KB of the team's scratch)

struct ScanSmem {
  bf16 x[SCAN_MAX];
  unsigned long long red64[8];
  float redf[8];
  unsigned int hist[256];            // PH_TOPK_HIST
};

__device__ __forceinline__ void slice_bounds(int V, int n_slices, int s, int& lo, int& hi) {
  const int per = ((V + n_slices - 1) / n_slices + 7) & ~7;
  lo = min(V, s * per);
  hi = min(V, lo + per);
}
// stage logits[row][lo, hi) in shared memory (16-byte loads, all in flight together)
__device__ __forceinline__ void stage_slice(ScanSmem& sm, const bf16* xr, int lo, int hi, int tid, BigTeamSync sync) {
  sync();                                           // previous item done with the buffer
  const int nv = (hi - lo + 7) / 8;
#pragma unroll 1
  for (int i = tid; i < nv; i += PG_BIG_TEAM)
    *reinterpret_cast<uint4*>(&sm.x[i * 8]) = __ldcg(reinterpret_cast<const uint4*>(xr + lo + i * 8));
  sync();
}
__device__ __forceinline__ unsigned long long bteam_max_u64(unsigned long long v, ScanSmem& sm, int tid, BigTeamSync sync) {
  v = warp_max_u64(v);
  sync();
  if ((tid & 31) == 0) sm.red64[tid >> 5] = v;
  sync();
  unsigned long long r = sm.red64[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = sm.red64[i] > r ? sm.red64[i] : r;
  return r;
}
__device__ __forceinline__ float bteam_max_f(float v, ScanSmem& sm, int tid, BigTeamSync sync) {
  v = warp_max(v);
  sync();
  if ((tid & 31) == 0) sm.redf[tid >> 5] = v;
  sync();
  float r = sm.redf[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = fmaxf(r, sm.redf[i]);
  return r;
}
__device__ __forceinline__ float bteam_sum_f(float v, ScanSmem& sm, int tid, BigTeamSync sync) {
  v = warp_sum(v);
  sync();
  if ((tid & 31) == 0) sm.redf[tid >> 5] = v;
  sync();
  float r = sm.redf[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r += sm.redf[i];      // fixed order
  return r;
}

// PH_ARGMAX: keys[row][slice] = best (value desc, index asc) key of the slice
__device__ __noinline__ void argmax_phase(const Phase& ph, int gteam, int n_teams, int tid, ScanSmem& sm, BigTeamSync sync) {
  const int n_items = ph.T * ph.n_slices;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    stage_slice(sm, ph.logits + (long)t * ph.ldl, lo, hi, tid, sync);
    unsigned long long best = 0ull;
#pragma unroll 1
    for (int i = tid; i < hi - lo; i += PG_BIG_TEAM) {
      const unsigned long long k = rank_key(bf2f(sm.x[i]), (uint32_t)(lo + i));
      best = k > best ? k : best;
    }
    best = bteam_max_u64(best, sm, tid, sync);
    if (tid == 0) ph.keys[t * ph.n_slices + s] = best;
  }
}

// PH_TOPK_HIST: hist[row][level][256] += histogram of the slice (level 0: high key byte; level 1: low byte of the keys
// whose high byte is the bin selected from level 0)
__device__ __noinline__ void topk_hist_phase(const Phase& ph, int gteam, int n_teams, int tid, ScanSmem& sm,
                                             BigTeamSync sync) {
  const int n_items = ph.T * ph.n_slices;
  const int lane = tid & 31;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    unsigned int above;
    const int b1 = ph.hist_level ? hist_select(ph.hist + (long)t * 512, (unsigned int)ph.top_k, lane, above) : -1;
    stage_slice(sm, ph.logits + (long)t * ph.ldl, lo, hi, tid, sync);
    sm.hist[tid] = 0;                                             // 256 threads, 256 bins
    sync();
#pragma unroll 1
    for (int i = tid; i < hi - lo; i += PG_BIG_TEAM) {
      const unsigned int key = bf16_key(sm.x[i]);
      if (b1 < 0) atomicAdd(&sm.hist[key >> 8], 1u);
      else if ((int)(key >> 8) == b1) atomicAdd(&sm.hist[key & 255u], 1u);
    }
    sync();
    if (sm.hist[tid]) atomicAdd(ph.hist + (long)t * 512 + ph.hist_level * 256 + tid, sm.hist[tid]);
  }
}

// PH_LSM_STATS: lstats[row][slice] = (max, sum exp(x - max)) of the slice (top_k > 0: over the logits the warper keeps)
__device__ __noinline__ void lsm_stats_phase(const Phase& ph, int gteam, int n_teams, int tid, ScanSmem& sm,
                                             BigTeamSync sync) {
  const int n_items = ph.T * ph.n_slices;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    const int row = ph.row_sel ? ld_vol(ph.row_sel) : t;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    const unsigned int theta = ph.top_k > 0 ? topk_threshold_key(ph.hist + (long)row * 512, ph.top_k, tid & 31) : 0u;
    stage_slice(sm, ph.logits + (long)row * ph.ldl, lo, hi, tid, sync);
    const float it = ph.inv_temp != 0.f ? ph.inv_temp : 1.f;     // T > 0 acceptance: statistics of logits / temperature
    float m = -INFINITY;
#pragma unroll 1
    for (int i = tid; i < hi - lo; i += PG_BIG_TEAM)
      if (bf16_key(sm.x[i]) >= theta) m = fmaxf(m, bf2f(sm.x[i]) * it);
    m = bteam_max_f(m, sm, tid, sync);
    float l = 0.f;
#pragma unroll 1
    for (int i = tid; i < hi - lo; i += PG_BIG_TEAM)
      if (bf16_key(sm.x[i]) >= theta) l += expf(bf2f(sm.x[i]) * it - m);
    l = bteam_sum_f(l, sm, tid, sync);
    if (tid == 0) ph.lstats[t * ph.n_slices + s] = make_float2(m, l);
  }
}

// PH_LSM_CAND: lp = bf16((x - max) - log(sum exp(x - max))) with the row's global statistics, then the slice's k best
// by (value desc, index asc) -> keys[row][slice][8]   (cnets_ours.py:1111-1115,1145-1148)
__device__ __noinline__ void lsm_cand_phase(const Phase& ph, int gteam, int n_teams, int tid, ScanSmem& sm,
                                            BigTeamSync sync) {
  constexpr int KMAX = 8;
  const int n_items = ph.T * ph.n_slices;
  const int lane = tid & 31;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    const int row = ph.row_sel ? ld_vol(ph.row_sel) : t;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    // global max / log-sum-exp from the slice statistics: lane i holds slices i, i+32 (fixed shuffle order, the same in
    // every warp)
    float2 st0 = lane < ph.n_slices ? __ldcg(&ph.lstats[t * ph.n_slices + lane]) : make_float2(-INFINITY, 0.f);
    float2 st1 = lane + 32 < ph.n_slices ? __ldcg(&ph.lstats[t * ph.n_slices + lane + 32]) : make_float2(-INFINITY, 0.f);
    stage_slice(sm, ph.logits + (long)row * ph.ldl, lo, hi, tid, sync);
    const float m = warp_max(fmaxf(st0.x, st1.x));
    float sum = 0.f;
    if (st0.x != -INFINITY) sum += st0.y * expf(st0.x - m);
    if (st1.x != -INFINITY) sum += st1.y * expf(st1.x - m);
    sum = warp_sum(sum);
    const float lse = logf(sum);
    unsigned long long loc[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) loc[j] = 0ull;
#pragma unroll 1
    for (int i = tid; i < hi - lo; i += PG_BIG_TEAM) {
      const float lp = rbf((bf2f(sm.x[i]) - m) - lse);
      const unsigned long long key = rank_key(lp, (uint32_t)(lo + i));
      if (key > loc[KMAX - 1]) {
        loc[KMAX - 1] = key;
#pragma unroll
        for (int j = KMAX - 1; j > 0; --j) {
          if (loc[j] > loc[j - 1]) {
            const unsigned long long tmp = loc[j];
            loc[j] = loc[j - 1];
            loc[j - 1] = tmp;
          }
        }
      }
    }
#pragma unroll 1
    for (int r = 0; r < KMAX; ++r) {
      unsigned long long best = 0ull;
      if (r < ph.k) {
        best = bteam_max_u64(loc[0], sm, tid, sync);
        if (loc[0] == best && best != 0ull) {   // keys are unique (index in the low word): exactly one owner pops
#pragma unroll
          for (int j = 0; j < KMAX - 1; ++j) loc[j] = loc[j + 1];
          loc[KMAX - 1] = 0ull;
        }
      }
      if (tid == 0) ph.keys[(t * ph.n_slices + s) * KMAX + r] = best;
    }
  }
}

// ---- merges on CTA 0 (all 256 workers), run at the start of the tree phase that consumes them ------------------
// amax[t] = index of the best key over the slices
__device__ __forceinline__ void merge_argmax(const Phase& ph, int wtid) {
  if (wtid < ph.T) {
    unsigned long long best = 0ull;
    for (int s = 0; s < ph.n_slices; ++s) {
      unsigned long long k = __ldcg(ph.keys + wtid * ph.n_slices + s);
      best = k > best ? k : best;
    }
    ph.amax_out[wtid] = (int)key_index(best);
  }
}
// top_val/top_idx[row][k] = the k best of the row's n_slices * 8 candidates, best first.  `stage`: >= T*n_slices*8 keys.
__device__ __forceinline__ void merge_topk(const Phase& ph, int wtid, unsigned long long* stage, int stage_cap,
                                           WorkerSync sync) {
  const int per_row = ph.n_slices * 8;
  const int rows_per_batch = max(1, stage_cap / per_row);
  for (int r0 = 0; r0 < ph.T; r0 += rows_per_batch) {
    const int n = min(rows_per_batch, ph.T - r0) * per_row;
    sync();
    for (int i = wtid; i < n; i += PG_WORKERS) stage[i] = __ldcg(ph.keys + r0 * per_row + i);
    sync();
    for (int i = wtid; i < n; i += PG_WORKERS) {
      const int row = i / per_row;
      const unsigned long long me = stage[i];
      if (me == 0ull) continue;
      const unsigned long long* rk = stage + row * per_row;
      int rank = 0;
      for (int j = 0; j < per_row; ++j) rank += rk[j] > me;
      if (rank < ph.k) {
        ph.top_val[(r0 + row) * ph.k + rank] = key_value(me);
        ph.top_idx[(r0 + row) * ph.k + rank] = (int)key_index(me);
      }
    }
  }
  sync();
}

}  // namespace vs
