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

constexpr int PG_THREADS = 384;
constexpr int PG_WORKERS = 256;
constexpr int PG_TEAM = 128;
constexpr int PG_BM = 128, PG_BK = 64;
constexpr int TILE_FLOATS = VS_TOK * PG_BM;      // one partial tile: [32 tokens][128 rows] fp32

struct TeamSync {
  int id;
  __device__ __forceinline__ void operator()() const { named_bar_sync(id, PG_TEAM); }
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
__device__ __forceinline__ int ld_vol(const int* p) { return *((volatile const int*)p); }

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
// PH_EPI: item = (weight-row tile, token group of TPI tokens)
// ---------------------------------------------------------------------------------------------------------------
struct Contrib {
  const float* base[8];     // partial tiles of up to 8 contributors, packed from index 0 (null beyond the count)
  int next_c, count;
};
// contributors c0 .. c0+7 of `tile` (c0 = 0 on the first call; continue with next_c while next_c < count)
__device__ __forceinline__ void contributors(const Phase& ph, const int* tables, int tile, int c0, Contrib& out) {
  const int4 e = *reinterpret_cast<const int4*>(tables + ph.tab_off + 4 * tile);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ci = c0 + c;
    const long slot = (long)(e.x + ci) * ph.maxseg + (ci == 0 ? e.z : 0);
    out.base[c] = ci < e.y ? ph.partials + slot * TILE_FLOATS : nullptr;
  }
  out.next_c = c0 + 8;
  out.count = e.y;
}

// sum over contributors (k order) of NC columns (col0 + 64 * j), tokens t0 .. t0+TPT-1 (t < T), into acc[j][]:
// all loads of four contributors are issued before the first add
template <int TPT, int NC>
__device__ __forceinline__ void accumulate(const Contrib& cb, int col0, int t0, int T, float (&acc)[NC][TPT]) {
#pragma unroll
  for (int c0 = 0; c0 < 8; c0 += 4) {
    if (!cb.base[c0]) break;                     // contributors are packed from index 0
    float v[4][NC][TPT];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int i = 0; i < TPT; ++i)
          v[c][j][i] = (cb.base[c0 + c] && t0 + i < T) ? __ldcg(cb.base[c0 + c] + (t0 + i) * PG_BM + col0 + 64 * j) : 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < NC; ++j)
#pragma unroll
        for (int i = 0; i < TPT; ++i) acc[j][i] += v[c][j][i];
  }
}

// STORE / RESID: thread = tile row (output feature), TPI tokens
template <int TPI>
__device__ __forceinline__ void epi_store_resid(const Phase& ph, const int* tables, int tile, int t0, int tt, float* red,
                                                TeamSync sync) {
  const int T = ph.T;
  const int n = tile * PG_BM + tt;
  const bool ok = n < ph.N;
  const bool resid = ph.epi == EPI_RESID;
  float acc[1][TPI], r[TPI];
#pragma unroll
  for (int i = 0; i < TPI; ++i) {
    acc[0][i] = 0.f;
    r[i] = (resid && ok && t0 + i < T) ? ldcg_bf(ph.R + (long)(t0 + i) * ph.ldr + n) : 0.f;
  }
  const float bias = (ok && ph.bias) ? bf2f(ph.bias[n]) : 0.f;
  Contrib cb;
  int c0 = 0;
  do {
    contributors(ph, tables, tile, c0, cb);
    accumulate<TPI, 1>(cb, tt, t0, T, acc);
    c0 = cb.next_c;
  } while (c0 < cb.count);
  const int warp = tt >> 5, lane = tt & 31;
  if (ph.ssq) sync();                                              // previous item's readers of `red` are done
#pragma unroll
  for (int i = 0; i < TPI; ++i) {
    if (t0 + i < T) {                                              // uniform across the team
      float v = acc[0][i] + bias;
      if (resid) v = r[i] + rbf(v);                                // Linear output rounds to bf16, then the add rounds again
      bf16 y = f2bf(v);
      if (ok) ph.Y[(long)(t0 + i) * ph.ldy + n] = y;
      if (ph.ssq) {
        float q = ok ? bf2f(y) * bf2f(y) : 0.f;
        q = warp_sum(q);
        if (lane == 0) red[warp * VS_TOK + i] = q;
      }
    }
  }
  if (ph.ssq) {
    sync();
    if (tt < TPI && t0 + tt < T)
      ph.ssq[tile * VS_TOK + t0 + tt] = red[tt] + red[VS_TOK + tt] + red[2 * VS_TOK + tt] + red[3 * VS_TOK + tt];
  }
}

// SWIGLU: tile rows [0,64) = gate features j0..j0+63, rows [64,128) = the matching up features.
// thread = (feature r < 64, half of the token group)
template <int TPI>
__device__ __forceinline__ void epi_swiglu(const Phase& ph, const int* tables, int tile, int t0, int tt) {
  constexpr int TPT = TPI / 2;
  const int T = ph.T;
  const int r = tt & 63, ts = t0 + (tt >> 6) * TPT;
  float gu[2][TPT];
#pragma unroll
  for (int i = 0; i < TPT; ++i) { gu[0][i] = 0.f; gu[1][i] = 0.f; }
  Contrib cb;
  int c0 = 0;
  do {
    contributors(ph, tables, tile, c0, cb);
    accumulate<TPT, 2>(cb, r, ts, T, gu);
    c0 = cb.next_c;
  } while (c0 < cb.count);
  const int j = tile * 64 + r;
#pragma unroll
  for (int i = 0; i < TPT; ++i)
    if (ts + i < T) {
      float a = rbf(pg_silu(rbf(gu[0][i])));                       // act_fn(gate) rounds, then * up rounds
      ph.Y[(long)(ts + i) * ph.ldy + j] = f2bf(a * rbf(gu[1][i]));
    }
}

// HEADS: one 128-row tile == one head of q | k | v; optional RoPE; q -> Qout, k/v -> KV rows in place.
// thread = (d < 64 paired with d + 64, half of the token group)
template <int TPI>
__device__ __forceinline__ void epi_heads(const Phase& ph, const int* tables, int tile, int t0, int tt) {
  constexpr int TPT = TPI / 2;
  const int T = ph.T;
  const int d = tt & 63, ts = t0 + (tt >> 6) * TPT;
  const bool is_q = tile < ph.n_q_tiles;
  const bool is_k = !is_q && tile < ph.n_q_tiles + ph.n_k_tiles;
  const bool rope = ph.rope && (is_q || is_k);
  float a[2][TPT], c1[TPT], s1[TPT], c2[TPT], s2[TPT];
#pragma unroll
  for (int i = 0; i < TPT; ++i) {
    a[0][i] = 0.f; a[1][i] = 0.f;
    c1[i] = 1.f; s1[i] = 0.f; c2[i] = 1.f; s2[i] = 0.f;
    if (rope && ts + i < T) {
      const long pos = ld_vol(ph.pos_ids + ts + i);
      c1[i] = bf2f(ph.cos[pos * 128 + d]); s1[i] = bf2f(ph.sin[pos * 128 + d]);
      c2[i] = bf2f(ph.cos[pos * 128 + 64 + d]); s2[i] = bf2f(ph.sin[pos * 128 + 64 + d]);
    }
  }
  const int row_base = (ph.kv_row_base ? ld_vol(ph.kv_row_base) : 0) + ph.kv_row_off;
  const float b1 = ph.bias ? bf2f(ph.bias[tile * PG_BM + d]) : 0.f;
  const float b2 = ph.bias ? bf2f(ph.bias[tile * PG_BM + 64 + d]) : 0.f;
  Contrib cb;
  int c0 = 0;
  do {
    contributors(ph, tables, tile, c0, cb);
    accumulate<TPT, 2>(cb, d, ts, T, a);
    c0 = cb.next_c;
  } while (c0 < cb.count);
  const int head = is_q ? tile : (is_k ? tile - ph.n_q_tiles : tile - ph.n_q_tiles - ph.n_k_tiles);
  bf16* dst = is_q ? ph.Qout + (long)head * ph.q_head_stride
                   : (is_k ? ph.Kout : ph.Vout) + (long)head * ph.kv_head_stride + (long)row_base * 128;
#pragma unroll
  for (int i = 0; i < TPT; ++i) {
    const int t = ts + i;
    if (t < T) {
      const float x1 = rbf(a[0][i] + b1), x2 = rbf(a[1][i] + b2);
      float o1 = x1, o2 = x2;
      if (rope) {                                                  // three bf16 roundings (cnets_ours.py:117-118)
        o1 = rbf(x1 * c1[i]) + rbf(-x2 * s1[i]);
        o2 = rbf(x2 * c2[i]) + rbf(x1 * s2[i]);
      }
      dst[(long)t * 128 + d] = f2bf(o1);
      dst[(long)t * 128 + 64 + d] = f2bf(o2);
    }
  }
}

template <int TPI>
__device__ __forceinline__ void epi_item(const Phase& ph, const int* tables, int tile, int t0, int tt, float* red,
                                         TeamSync sync) {
  if (ph.epi == EPI_SWIGLU) epi_swiglu<TPI>(ph, tables, tile, t0, tt);
  else if (ph.epi == EPI_HEADS) epi_heads<TPI>(ph, tables, tile, t0, tt);
  else epi_store_resid<TPI>(ph, tables, tile, t0, tt, red, sync);
}

__device__ __noinline__ void epi_phase(const Phase& ph, const int* tables, int gteam, int n_teams, int tt, float* red,
                                       TeamSync sync) {
  const int tpi = ph.tpi;
  const int groups = (ph.T + tpi - 1) / tpi;
  const int n_items = ph.n_tiles * groups;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int tile = item / groups, t0 = (item % groups) * tpi;
    if (tpi == 4) epi_item<4>(ph, tables, tile, t0, tt, red, sync);
    else if (tpi == 8) epi_item<8>(ph, tables, tile, t0, tt, red, sync);
    else epi_item<16>(ph, tables, tile, t0, tt, red, sync);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PH_NORM: item = (token row, quarter of the row).  y = weight * bf16(x * rsqrt(mean(x^2) + eps)).
// ---------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void norm_phase(const Phase& ph, int gteam, int n_teams, int tt, float* red, TeamSync sync) {
  const int H = ph.H;
  const int n_items = ph.T * 4;
  const int lane = tt & 31;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item >> 2, part = item & 3;
    const bf16* x = ph.src + (long)t * ph.lds;
    bf16* y = ph.dst + (long)t * ph.ldd;
    // this thread's slice of the row: 16-byte vectors v = part * H/32 + tt + 128 * j
    const int per_part = H / 32;                 // vectors per quarter
    uint4 xv[2], wv[2];
    int vi[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      vi[j] = tt + PG_TEAM * j;
      const bool in = vi[j] < per_part;
      const int e = (part * per_part + (in ? vi[j] : 0)) * 8;
      xv[j] = in ? __ldcg(reinterpret_cast<const uint4*>(x + e)) : make_uint4(0, 0, 0, 0);
      wv[j] = in ? *reinterpret_cast<const uint4*>(ph.w + e) : make_uint4(0, 0, 0, 0);
    }
    float ss = 0.f;
    if (ph.ssq_in) {
      // per-tile sums of squares left by the residual epilogue: lane i sums tiles i, i+32, ... (fixed order)
      for (int i = lane; i < ph.ssq_tiles; i += 32) ss += __ldcg(ph.ssq_in + i * VS_TOK + t);
      ss = warp_sum(ss);
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
      if (lane == 0) red[tt >> 5] = ss;
      sync();
      ss = red[0] + red[1] + red[2] + red[3];
    }
    const float rstd = 1.0f / sqrtf(ss / (float)H + ph.eps);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      for (int v = vi[j]; v < per_part; v += 2 * PG_TEAM) {        // first trip uses the preloaded vectors
        const int e = (part * per_part + v) * 8;
        uint4 xq = xv[j], wq = wv[j];
        if (v != vi[j]) {
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
// vocabulary scans: item = (row, slice of the vocabulary); the slices of a row are merged on CTA 0 (tree phase)
// ---------------------------------------------------------------------------------------------------------------
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

// slice s of V elements, aligned to 8: [lo, hi)
__device__ __forceinline__ void slice_bounds(int V, int n_slices, int s, int& lo, int& hi) {
  const int per = ((V + n_slices - 1) / n_slices + 7) & ~7;
  lo = min(V, s * per);
  hi = min(V, lo + per);
}
// loads the slice [lo, hi) of a row into registers: up to NV 16-byte vectors per thread (vector v = tt + 128 * j);
// rows that are not 16-byte aligned or slices longer than NV * 1024 elements are not supported by the builder.
template <int NV>
__device__ __forceinline__ void load_slice(const bf16* xr, int lo, int hi, int tt, uint4 (&v)[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int e = lo + (tt + PG_TEAM * j) * 8;
    v[j] = e < hi ? __ldcg(reinterpret_cast<const uint4*>(xr + e)) : make_uint4(0, 0, 0, 0);
  }
}
template <int NV, class F>
__device__ __forceinline__ void visit_slice(const uint4 (&v)[NV], int lo, int hi, int tt, F f) {
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int e = lo + (tt + PG_TEAM * j) * 8;
    if (e < hi) {
      const unsigned int w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (e + 2 * q < hi) f(bf_lo(w[q]), e + 2 * q);
        if (e + 2 * q + 1 < hi) f(bf_hi(w[q]), e + 2 * q + 1);
      }
    }
  }
}
constexpr int SCAN_NV = 4;     // slices of up to 4096 elements

// PH_ARGMAX: keys[row][slice] = best (value desc, index asc) key of the slice
__device__ __noinline__ void argmax_phase(const Phase& ph, int gteam, int n_teams, int tt, void* red, TeamSync sync) {
  const int n_items = ph.T * ph.n_slices;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    uint4 v[SCAN_NV];
    load_slice<SCAN_NV>(ph.logits + (long)t * ph.ldl, lo, hi, tt, v);
    unsigned long long best = 0ull;
    visit_slice<SCAN_NV>(v, lo, hi, tt, [&](float x, int i) {
      unsigned long long k = rank_key(x, (uint32_t)i);
      best = k > best ? k : best;
    });
    best = team_max_u64(best, reinterpret_cast<unsigned long long*>(red), tt, sync);
    if (tt == 0) ph.keys[t * ph.n_slices + s] = best;
  }
}

// PH_LSM_STATS: lstats[row][slice] = (max, sum exp(x - max)) of the slice
__device__ __noinline__ void lsm_stats_phase(const Phase& ph, int gteam, int n_teams, int tt, void* red, TeamSync sync) {
  const int n_items = ph.T * ph.n_slices;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    const int row = ph.row_sel ? ld_vol(ph.row_sel) : t;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    uint4 v[SCAN_NV];
    load_slice<SCAN_NV>(ph.logits + (long)row * ph.ldl, lo, hi, tt, v);
    float m = -INFINITY;
    visit_slice<SCAN_NV>(v, lo, hi, tt, [&](float x, int) { m = fmaxf(m, x); });
    m = team_max_f(m, reinterpret_cast<float*>(red), tt, sync);
    float l = 0.f;
    visit_slice<SCAN_NV>(v, lo, hi, tt, [&](float x, int) { l += expf(x - m); });
    l = team_sum_f(l, reinterpret_cast<float*>(red), tt, sync);
    if (tt == 0) ph.lstats[t * ph.n_slices + s] = make_float2(m, l);
  }
}

// PH_LSM_CAND: lp = bf16((x - max) - log(sum exp(x - max))) with the row's global statistics, then the slice's k best
// by (value desc, index asc) -> keys[row][slice][8]   (cnets_ours.py:1111-1115,1145-1148)
__device__ __noinline__ void lsm_cand_phase(const Phase& ph, int gteam, int n_teams, int tt, void* red, TeamSync sync) {
  constexpr int KMAX = 8;
  const int n_items = ph.T * ph.n_slices;
  for (int item = gteam; item < n_items; item += n_teams) {
    const int t = item / ph.n_slices, s = item % ph.n_slices;
    const int row = ph.row_sel ? ld_vol(ph.row_sel) : t;
    int lo, hi;
    slice_bounds(ph.V, ph.n_slices, s, lo, hi);
    uint4 v[SCAN_NV];
    load_slice<SCAN_NV>(ph.logits + (long)row * ph.ldl, lo, hi, tt, v);
    // global max / log-sum-exp from the slice statistics: lane i holds slices i, i+32 (fixed shuffle order, the same
    // in every warp)
    const int lane = tt & 31;
    float2 st0 = lane < ph.n_slices ? __ldcg(&ph.lstats[t * ph.n_slices + lane]) : make_float2(-INFINITY, 0.f);
    float2 st1 = lane + 32 < ph.n_slices ? __ldcg(&ph.lstats[t * ph.n_slices + lane + 32]) : make_float2(-INFINITY, 0.f);
    const float m = warp_max(fmaxf(st0.x, st1.x));
    float sum = 0.f;
    if (st0.x != -INFINITY) sum += st0.y * expf(st0.x - m);
    if (st1.x != -INFINITY) sum += st1.y * expf(st1.x - m);
    sum = warp_sum(sum);
    const float lse = logf(sum);
    unsigned long long loc[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) loc[j] = 0ull;
    visit_slice<SCAN_NV>(v, lo, hi, tt, [&](float x, int i) {
      float lp = rbf((x - m) - lse);
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
    for (int r = 0; r < KMAX; ++r) {
      unsigned long long best = 0ull;
      if (r < ph.k) {
        best = team_max_u64(loc[0], reinterpret_cast<unsigned long long*>(red), tt, sync);
        if (loc[0] == best && best != 0ull) {   // keys are unique (index in the low word): exactly one owner pops
#pragma unroll
          for (int j = 0; j < KMAX - 1; ++j) loc[j] = loc[j + 1];
          loc[KMAX - 1] = 0ull;
        }
      }
      if (tt == 0) ph.keys[(t * ph.n_slices + s) * KMAX + r] = best;
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
