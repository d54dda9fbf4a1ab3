// Device bodies of the two-pass attention (see attention.cu for the algorithm and the reference call sites).
// They are written for a "team" of 256 threads (8 warps) that owns one work item at a time, so the same code runs
//   * as a stand-alone kernel (attention.cu: team = thread block, sync = __syncthreads): prompt passes, draft prefill, and
//   * inside the persistent program kernel as the two-phase program (program.cu, option attn_fused = 0: the A/B reference
//     of the single-phase, TMA-staged attention of attn_ring.cuh).
// A phase of the program kernel is a few microseconds long and runs at low occupancy, so the bodies are written for
// short dependent-instruction chains: 8 warps per item (8 keys resp. 16 output columns per warp), the visibility of the
// 64 keys of a chunk is one 64-bit mask per query row, all global loads of an item are issued up front.
#pragma once
#include <math.h>

#include "kernels.h"

namespace vs {

// Optional checkpoint clocks (compile with -DVS_PROGRAM_CLK): see program_dev.cuh
#ifdef VS_PROGRAM_CLK
__device__ long long g_clk[64];
#define ACLK(i) do { if (aclk_on) g_clk[i] = clock64(); } while (0)
#else
#define ACLK(i)
#endif

constexpr int ATT_CHUNK = 64;      // keys per chunk
constexpr int ATT_THREADS = 256;   // one team
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_ROW_B = 256;     // bytes per 128-dim bf16 row

struct __align__(128) AttnSmem {
  union {
    struct {
      uint8_t qs[VS_TOK * ATT_ROW_B];        // Q tile [32][128] bf16, swizzled
      uint8_t ks[ATT_CHUNK * ATT_ROW_B];     // K chunk [64][128] bf16, swizzled
    } s;
    struct {
      uint8_t vsm[ATT_CHUNK * ATT_ROW_B];    // V chunk [64][128] bf16, swizzled
      uint8_t psm[VS_TOK * 128];             // P chunk [32][64] bf16, swizzled
    } v;
  };
  float red[ATT_WARPS][VS_TOK];
  float rowm[VS_TOK];
  float gM[VS_TOK], gL[VS_TOK];
  int last;
};

__device__ __forceinline__ int att_swz256(int row, int chunk16) {   // byte offset in a [rows][128] bf16 tile
  return row * ATT_ROW_B + (((chunk16 & 8) | ((chunk16 ^ row) & 7)) << 4);
}

// tile rows [0, n_rows_tile) <- src rows first_row + r, zero-filled for r >= n_rows_valid
__device__ __forceinline__ void att_load_rows_async(uint8_t* dst, const bf16* src, int first_row, int n_rows_valid,
                                                    int n_rows_tile, int tid) {
  for (int i = tid; i < n_rows_tile * 16; i += ATT_THREADS) {
    int r = i >> 4, ch = i & 15;
    bool ok = r < n_rows_valid;
    const bf16* s = src + (long)(first_row + (ok ? r : 0)) * 128 + ch * 8;
    cp_async16(dst + att_swz256(r, ch), s, ok);
  }
}

__device__ __forceinline__ unsigned long long att_low_mask(int n) {   // n lowest bits set, n clamped to [0, 64]
  return n <= 0 ? 0ull : (n >= 64 ? ~0ull : ((1ull << n) - 1ull));
}
// bit j = key (key0 + j) is visible to query row q_abs: inside the kv length, and either in the fully visible prefix,
// or causal (key < n_ctx + q_abs + 1), or allowed by the query's tree word (bit i = key n_ctx + i)
__device__ __forceinline__ unsigned long long att_visible_mask(bool causal, int n_ctx, int kv_len, int q_abs,
                                                               unsigned long long qbits, int key0) {
  unsigned long long m;
  if (causal) {
    m = att_low_mask(n_ctx + q_abs + 1 - key0);
  } else {
    m = att_low_mask(n_ctx - key0);
    const int sh = key0 - n_ctx;                       // tree bit index of key0
    if (sh >= 0) m |= sh < 64 ? (qbits >> sh) : 0ull;
    else m |= -sh < 64 ? (qbits << -sh) : 0ull;
  }
  return m & att_low_mask(kv_len - key0);
}
__device__ __forceinline__ unsigned long long att_visible_mask(const AttnArgs& p, int n_ctx, int kv_len, int q_abs,
                                                               unsigned long long qbits, int key0) {
  return att_visible_mask(p.causal != 0, n_ctx, kv_len, q_abs, qbits, key0);
}

// ---- pass 1: one (head, query block, key chunk) item: S = Q K^T, mask, scale, chunk max / sum -----------------
// `load_q`: the Q tile of (head, tqb) is not in sm.s.qs yet.
template <class Sync>
__device__ __forceinline__ void attn_scores_item(const AttnArgs& p, AttnSmem& sm, int head, int tqb, int tq_blocks,
                                                 int c, bool load_q, int tid, Sync sync) {
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int warp = tid >> 5, lane = tid & 31;
  const int n_ctx = p.n_ctx ? __ldcg(p.n_ctx) : p.n_ctx_host;
  const int kv_len = n_ctx + p.n_new;
  const int t0 = tqb * VS_TOK;
  const int tq_rows = min(VS_TOK, p.Tq - t0);
  const int g = lane >> 2, t4 = lane & 3;
  const int key0 = c * ATT_CHUNK;

  sync();   // previous item finished with the shared tiles
  if (load_q) att_load_rows_async(sm.s.qs, p.Q + (long)head * p.q_head_stride, t0, tq_rows, VS_TOK, tid);
  att_load_rows_async(sm.s.ks, p.K + (long)kvh * p.kv_head_stride, key0, min(ATT_CHUNK, kv_len - key0), ATT_CHUNK, tid);
  cp_async_commit();

  // the four query rows this thread owns in the accumulator layout: q(i) = (i>>1)*16 + g + (i&1)*8
  unsigned long long vis[4];
  {
    unsigned long long qb[4] = {0ull, 0ull, 0ull, 0ull};
    if (!p.causal && p.bits) {          // uniform; clamped indices keep the four loads unconditional and back to back
#pragma unroll
      for (int i = 0; i < 4; ++i) qb[i] = __ldcg(p.bits + t0 + min((i >> 1) * 16 + g + (i & 1) * 8, tq_rows - 1));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      vis[i] = att_visible_mask(p, n_ctx, kv_len, t0 + (i >> 1) * 16 + g + (i & 1) * 8, qb[i], key0);
  }
  float* sbase = p.scores + ((long)head * tq_blocks * VS_TOK + t0) * p.score_ld;
  float* stat_base = p.stats + (((long)head * tq_blocks + tqb) * p.max_chunks) * VS_TOK * 2;
  const float inv_div = sqrtf((float)VS_HEAD_DIM);

  cp_async_wait<0>();
  sync();

  // warp w: keys key0 + 8w .. +7;  acc[mb][e]: query = mb*16 + g + (e>>1)*8, key = key0 + 8*warp + 2*t4 + (e&1)
  float acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[a][e] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    uint32_t bfr[2];
    ldmatrix_x2(bfr, sm.s.ks + att_swz256(warp * 8 + (lane & 7), kk * 2 + ((lane >> 3) & 1)));
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      uint32_t afr[4];
      const int mat = lane >> 3;
      ldmatrix_x4(afr, sm.s.qs + att_swz256(mb * 16 + (lane & 7) + (mat & 1) * 8, kk * 2 + (mat >> 1)));
      mma_bf16(acc[mb], afr, bfr[0], bfr[1]);
    }
  }

  // scale / round / mask
  const int kofs = warp * 8 + 2 * t4;                 // key offset inside the chunk of element e&1 == 0
  float lmax[4];
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int eh = 0; eh < 2; ++eh) {
      const int qi = mb * 2 + eh;
      float s0 = acc[mb][eh * 2], s1 = acc[mb][eh * 2 + 1];
      s0 = p.eager ? rbf(rbf(s0) / inv_div) : s0 * p.scale;
      s1 = p.eager ? rbf(rbf(s1) / inv_div) : s1 * p.scale;
      if (!((vis[qi] >> kofs) & 1ull)) s0 = -INFINITY;
      if (!((vis[qi] >> (kofs + 1)) & 1ull)) s1 = -INFINITY;
      acc[mb][eh * 2] = s0;
      acc[mb][eh * 2 + 1] = s1;
      lmax[qi] = fmaxf(s0, s1);
    }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lmax[i] = fmaxf(lmax[i], __shfl_xor_sync(0xffffffffu, lmax[i], 1));
    lmax[i] = fmaxf(lmax[i], __shfl_xor_sync(0xffffffffu, lmax[i], 2));
    if (t4 == 0) sm.red[warp][(i >> 1) * 16 + g + (i & 1) * 8] = lmax[i];
  }
  sync();
  if (tid < VS_TOK) {
    float m = sm.red[0][tid];
#pragma unroll
    for (int w = 1; w < ATT_WARPS; ++w) m = fmaxf(m, sm.red[w][tid]);
    sm.rowm[tid] = m;
  }
  sync();
  float lsum[4];
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int eh = 0; eh < 2; ++eh) {
      const int qi = mb * 2 + eh;
      const int q = mb * 16 + g + eh * 8;
      const float s0 = acc[mb][eh * 2], s1 = acc[mb][eh * 2 + 1];
      const float rm = sm.rowm[q];
      float l = 0.f;
      if (s0 != -INFINITY) l += expf(s0 - rm);
      if (s1 != -INFINITY) l += expf(s1 - rm);
      lsum[qi] = l;
      *reinterpret_cast<float2*>(sbase + (long)q * p.score_ld + key0 + kofs) = make_float2(s0, s1);   // adjacent keys
    }
  sync();   // rowm consumed before red is overwritten below
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lsum[i] += __shfl_xor_sync(0xffffffffu, lsum[i], 1);
    lsum[i] += __shfl_xor_sync(0xffffffffu, lsum[i], 2);
    if (t4 == 0) sm.red[warp][(i >> 1) * 16 + g + (i & 1) * 8] = lsum[i];
  }
  sync();
  if (tid < VS_TOK) {
    float l = sm.red[0][tid];
#pragma unroll
    for (int w = 1; w < ATT_WARPS; ++w) l += sm.red[w][tid];          // fixed order
    *reinterpret_cast<float2*>(stat_base + ((long)c * VS_TOK + tid) * 2) = make_float2(sm.rowm[tid], l);
  }
}

// ---- pass 2: one (head, query block, kv split) item: P = softmax in the reference's rounding, O += P V ----------
// Splits own the chunks split, split + n_splits, ...; the last split to arrive reduces the partial tiles in split order.
template <class Sync>
__device__ __forceinline__ void attn_pv_item(const AttnArgs& p, AttnSmem& sm, int head, int tqb, int tq_blocks, int split,
                                             int n_splits, int tid, Sync sync) {
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int warp = tid >> 5, lane = tid & 31;
  const int n_ctx = p.n_ctx ? __ldcg(p.n_ctx) : p.n_ctx_host;
  const int kv_len = n_ctx + p.n_new;
  const int t0 = tqb * VS_TOK;
  const int tq_rows = min(VS_TOK, p.Tq - t0);
  // causal: the chunks behind the last row of this query block were skipped by pass 1 as well (same bits: their
  // statistics would be (-inf, 0) and their probabilities zero)
  const int n_chunks = p.causal ? min((kv_len + ATT_CHUNK - 1) / ATT_CHUNK, (n_ctx + t0 + tq_rows + ATT_CHUNK - 1) / ATT_CHUNK)
                                : (kv_len + ATT_CHUNK - 1) / ATT_CHUNK;
  const int g = lane >> 2, t4 = lane & 3;

  const float* sbase = p.scores + ((long)head * tq_blocks * VS_TOK + t0) * p.score_ld;
  const float* stat_base = p.stats + (((long)head * tq_blocks + tqb) * p.max_chunks) * VS_TOK * 2;
  const bf16* vbase = p.V + (long)kvh * p.kv_head_stride;

  sync();   // previous item finished with gM / gL / tiles
  // first V chunk of this split: in flight while the statistics are combined
  if (split < n_chunks)
    att_load_rows_async(sm.v.vsm, vbase, split * ATT_CHUNK, min(ATT_CHUNK, kv_len - split * ATT_CHUNK), ATT_CHUNK, tid);
  cp_async_commit();
  // global row max / sum from the chunk statistics: 8 lanes per query row, chunks part, part+8, ... (fixed order)
  {
    const int q = tid >> 3, part = tid & 7;
    float m = -INFINITY, l = 0.f;
    float2 st[4];
    int cc = part;
    // up to 32 chunks per trip, statistics loaded before they are combined
    for (int base = 0; base < n_chunks; base += 32) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = min(base + part + 8 * j, n_chunks - 1);       // clamped: unconditional loads
        st[j] = __ldcg(reinterpret_cast<const float2*>(stat_base + ((long)ci * VS_TOK + q) * 2));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + part + 8 * j >= n_chunks) st[j] = make_float2(-INFINITY, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (st[j].x != -INFINITY) {
          const float mn = fmaxf(m, st[j].x);
          l = l * expf(m - mn) + st[j].y * expf(st[j].x - mn);      // m == -inf: l == 0 and expf(-inf) == 0
          m = mn;
        }
      }
    }
    (void)cc;
    // combine the 8 partial (m, l) pairs of the row
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      const float a = m == -INFINITY ? 0.f : l * expf(m - mn);
      const float b = mo == -INFINITY ? 0.f : lo * expf(mo - mn);
      l = a + b;
      m = mn;
    }
    if (part == 0) { sm.gM[q] = m; sm.gL[q] = l; }
  }
  sync();

  // warp w: output columns 16w .. 16w+15;  acc[mb][nb][e]: query = mb*16 + g + (e>>1)*8, d = 16w + nb*8 + 2*t4 + (e&1)
  float acc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

  for (int c = split; c < n_chunks; c += n_splits) {
    const int key0 = c * ATT_CHUNK;
    if (c != split) {
      sync();
      att_load_rows_async(sm.v.vsm, vbase, key0, min(ATT_CHUNK, kv_len - key0), ATT_CHUNK, tid);
      cp_async_commit();
    }
    // probabilities for this chunk: [32][64] bf16, 128-byte rows, 16-byte chunks swizzled by row; 2 x 4 per thread
    float4 s4[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = tid + ATT_THREADS * j;
      s4[j] = __ldcg(reinterpret_cast<const float4*>(sbase + (long)(i >> 4) * p.score_ld + key0 + (i & 15) * 4));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = tid + ATT_THREADS * j;
      const int q = i >> 4, k4 = (i & 15) * 4;
      const float sv[4] = {s4[j].x, s4[j].y, s4[j].z, s4[j].w};
      const float gm = sm.gM[q], gl = sm.gL[q];
      bf16 pr4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pr = 0.f;
        if (sv[e] != -INFINITY) {                     // masked keys (and keys past kv_len) hold -inf
          const float ex = expf(sv[e] - gm);
          pr = p.eager ? ex / gl : ex;
        }
        pr4[e] = f2bf(pr);
      }
      const int ch = k4 >> 3;
      bf16* dst = reinterpret_cast<bf16*>(sm.v.psm + q * 128 + ((ch ^ (q & 7)) << 4)) + (k4 & 7);
      *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(pr4);
    }
    cp_async_wait<0>();
    sync();
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t afr[2][4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const int mat = lane >> 3;
        const int r = mb * 16 + (lane & 7) + (mat & 1) * 8;
        const int ch = kk * 2 + (mat >> 1);
        ldmatrix_x4(afr[mb], sm.v.psm + r * 128 + ((ch ^ (r & 7)) << 4));
      }
      uint32_t bfr[4];
      {
        const int mat = lane >> 3;
        const int r = kk * 16 + (lane & 7) + (mat & 1) * 8;        // key row
        const int ch = warp * 2 + (mat >> 1);                      // 8-wide d chunk
        ldmatrix_x4_trans(bfr, sm.v.vsm + att_swz256(r, ch));
      }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        mma_bf16(acc[mb][0], afr[mb], bfr[0], bfr[1]);
        mma_bf16(acc[mb][1], afr[mb], bfr[2], bfr[3]);
      }
    }
  }

  // publish the partial O tile [32][128] fp32; the last split to arrive reduces in split order
  float* obase = p.opart + (((long)head * tq_blocks + tqb) * n_splits) * VS_TOK * 128;
  float* mine = obase + (long)split * VS_TOK * 128;
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int eh = 0; eh < 2; ++eh) {
        const int q = mb * 16 + g + eh * 8;
        const int d = warp * 16 + nb * 8 + 2 * t4;
        *reinterpret_cast<float2*>(mine + q * 128 + d) = make_float2(acc[mb][nb][eh * 2], acc[mb][nb][eh * 2 + 1]);
      }
  __threadfence();
  sync();
  if (tid == 0) {
    int old = atomicAdd(&p.counters[head * tq_blocks + tqb], 1);
    sm.last = (old == n_splits - 1);
  }
  sync();
  if (!sm.last) return;
  __threadfence();
  // 32 x 128 outputs = 1024 float4, 4 per thread; the partial tiles of four splits are loaded before they are added
  for (int i = tid; i < VS_TOK * 32; i += ATT_THREADS) {
    const int q = i >> 5, d4 = (i & 31) * 4;
    if (q >= tq_rows) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < n_splits; s0 += 4) {
      float4 t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        t[j] = __ldcg(reinterpret_cast<const float4*>(obase + (long)min(s0 + j, n_splits - 1) * VS_TOK * 128 + q * 128 + d4));
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (s0 + j < n_splits) { v.x += t[j].x; v.y += t[j].y; v.z += t[j].z; v.w += t[j].w; }
    }
    if (!p.eager) {
      const float l = sm.gL[q];
      v.x /= l; v.y /= l; v.z /= l; v.w /= l;
    }
    bf16 o4[4] = {f2bf(v.x), f2bf(v.y), f2bf(v.z), f2bf(v.w)};
    *reinterpret_cast<uint2*>(p.O + (long)(t0 + q) * p.ldo + head * 128 + d4) = *reinterpret_cast<const uint2*>(o4);
  }
  if (tid == 0) p.counters[head * tq_blocks + tqb] = 0;
}


// x / d, correctly rounded for normal operands: q = x * r with r = RN(1 / d), then one residual correction
// (the fast path of the IEEE division sequence without its range checks; |x|, d here are O(1e-3 .. 1e3))
__device__ __forceinline__ float att_div(float x, float d, float r) {
  const float q = x * r;
  return fmaf(fmaf(-q, d, x), r, q);
}
__device__ __forceinline__ unsigned int att_ld_relaxed(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// (The single-phase attention of the persistent program kernel -- TMA-staged K/V -- lives in attn_ring.cuh.)

}  // namespace vs
