// Device bodies of the decode attention (see attention.cu for the algorithm and the reference call sites).
// They are written for a "team" of 256 threads (8 warps) that owns one work item at a time, so the same code runs
//   * as a stand-alone kernel (attention.cu: team = thread block, sync = __syncthreads), and
//   * inside the persistent program kernel (program.cu: two attention teams per CTA, sync = a named barrier).
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


// =================================================================================================================
// Fused single-phase attention for the persistent program kernel (short and medium contexts).
//
// One item = (head, kv split) and EVERY item of the phase is resident at the same time (one per 256-thread team), so
// the two passes of the algorithm above can meet in the middle of the phase instead of at a grid barrier:
//   load Q + this split's K chunks (<= 2)  ->  S = Q K^T in registers, split max / sum-exp  ->  publish the statistics,
//   start loading the V chunks into the K buffers  ->  wait for the head's other splits (per-head arrival counter)  ->
//   exact global softmax denominators  ->  P in the reference's rounding (from the registers)  ->  P V  ->  partial O,
//   last split of the head reduces in split order.
// K and V are read once, scores never leave the SM, one HBM latency is exposed instead of two plus a barrier.
// Capacity: n_splits * 2 * 64 keys (1152 at 32 heads on 148 SMs); the caller picks the two-phase path beyond it.
// =================================================================================================================
struct __align__(128) AttnFusedSmem {
  uint8_t qp[VS_TOK * ATT_ROW_B];            // Q tile [32][128] bf16 (pass 1), then P [32][<=128 keys] bf16 (pass 2)
  uint8_t kv[2][ATT_CHUNK * ATT_ROW_B];      // K chunks, then V chunks
  float red[ATT_WARPS][VS_TOK];
  float rowm[VS_TOK];
  float gM[VS_TOK], gL[VS_TOK], gR[VS_TOK];  // global row max, softmax denominator and its reciprocal
  float red2[ATT_WARPS][VS_TOK];             // per-warp sum exp (red holds the per-warp max)
  int last;
};
constexpr int ATT_FUSED_CPS = 2;             // chunks per split the fused item can hold

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

// ---- shared tails of the single-phase items (register-resident and streamed) ---------------------------------------
// Split statistics -> head-wide softmax denominators.  On entry sm.red / sm.red2 hold every warp's (max, sum exp) of its
// key columns per query row (and a team barrier has passed).  The split's (m, l) goes to p.stats [head][split][32],
// the per-head arrival counter replaces a grid barrier, then 8 lanes per row combine the splits in fixed order into
// sm.gM / sm.gL / sm.gR.
template <class Sync>
__device__ __forceinline__ void att_exchange_stats(const AttnArgs& p, AttnFusedSmem& sm, int head, int split, int n_splits,
                                                   unsigned int* head_cnt, unsigned int epoch, int tid, Sync sync) {
  float* stat_head = p.stats + (long)head * p.max_chunks * VS_TOK * 2;
  if (tid < VS_TOK) {
    float m = sm.red[0][tid];
#pragma unroll
    for (int w = 1; w < ATT_WARPS; ++w) m = fmaxf(m, sm.red[w][tid]);
    float l = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) {                           // fixed order
      const float mw = sm.red[w][tid];
      if (mw != -INFINITY) l += sm.red2[w][tid] * expf(mw - m);
    }
    *reinterpret_cast<float2*>(stat_head + ((long)split * VS_TOK + tid) * 2) = make_float2(m, l);
  }
  sync();
  // ---- publish, then wait for the other splits of this head
  if (tid == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(head_cnt + head) : "memory");
    const unsigned int want = (epoch + 1u) * (unsigned int)n_splits;
    while (att_ld_relaxed(head_cnt + head) < want) {
    }
  }
  sync();
  // ---- exact global denominators from the split statistics: 8 lanes per row, splits part, part+8 (fixed order)
  {
    const int q = tid >> 3, part = tid & 7;
    float2 st[2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
      st[j] = __ldcg(reinterpret_cast<const float2*>(stat_head + ((long)min(part + 8 * j, n_splits - 1) * VS_TOK + q) * 2));
    float m = -INFINITY, l = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (part + 8 * j < n_splits && st[j].x != -INFINITY) {
        const float mn = fmaxf(m, st[j].x);
        l = l * expf(m - mn) + st[j].y * expf(st[j].x - mn);
        m = mn;
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m, o), lo = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, mo);
      const float a = m == -INFINITY ? 0.f : l * expf(m - mn);
      const float b = mo == -INFINITY ? 0.f : lo * expf(mo - mn);
      l = a + b;
      m = mn;
    }
    if (part == 0) { sm.gM[q] = m; sm.gL[q] = l; sm.gR[q] = 1.0f / l; }
  }
  sync();
}

// Partial O [32][128] fp32 of this split -> p.opart, second per-head arrival counter, then every split reduces its
// share of the head's output in split order (float4 group i belongs to split i % n_splits) over the n_act splits that
// own keys.
template <class Sync>
__device__ __forceinline__ void att_exchange_output(const AttnArgs& p, AttnFusedSmem& sm, const float (&oacc)[2][2][4],
                                                    int head, int split, int n_splits, int n_act, int tq_rows,
                                                    unsigned int* head_cnt, unsigned int epoch, int tid, Sync sync) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  float* obase = p.opart + ((long)head * n_splits) * VS_TOK * 128;
  float* mine = obase + (long)split * VS_TOK * 128;
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int eh = 0; eh < 2; ++eh) {
        const int q = mb * 16 + g + eh * 8;
        const int d = warp * 16 + nb * 8 + 2 * t4;
        *reinterpret_cast<float2*>(mine + q * 128 + d) = make_float2(oacc[mb][nb][eh * 2], oacc[mb][nb][eh * 2 + 1]);
      }
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
    if (!p.eager) {
      const float l = sm.gL[q];
      v.x /= l; v.y /= l; v.z /= l; v.w /= l;
    }
    bf16 o4[4] = {f2bf(v.x), f2bf(v.y), f2bf(v.z), f2bf(v.w)};
    *reinterpret_cast<uint2*>(p.O + (long)q * p.ldo + head * 128 + d4) = *reinterpret_cast<const uint2*>(o4);
  }
}

// head_cnt: 2 * n_heads arrival counters (statistics published | partial O published), monotonic across launches
// (epoch = launches completed before this one).
// stats layout (p.stats): [head][split][32 rows] float2 (max, sum exp).
template <class Sync>
__device__ __forceinline__ void attn_fused_item(const AttnArgs& p, AttnFusedSmem& sm, int head, int split, int n_splits,
                                                unsigned int* head_cnt, unsigned int epoch, int tid, Sync sync) {
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int warp = tid >> 5, lane = tid & 31;
  // the two dependent-free global loads of the item (kv length, tree words) go out together, ahead of everything else:
  // behind the K/V copies of this SM they take ~2 us and the masks sit on the critical path
  const int tq_rows = min(VS_TOK, p.Tq);
  unsigned long long qb[4] = {0ull, 0ull, 0ull, 0ull};
  if (!p.causal && p.bits) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      qb[i] = __ldcg(p.bits + min((i >> 1) * 16 + (lane >> 2) + (i & 1) * 8, tq_rows - 1));
  }
  const int n_ctx = p.n_ctx ? __ldcg(p.n_ctx) : p.n_ctx_host;
  const int kv_len = n_ctx + p.n_new;
  const int n_chunks = (kv_len + ATT_CHUNK - 1) / ATT_CHUNK;
  const int cps = (n_chunks + n_splits - 1) / n_splits;           // chunks per split (contiguous)
  if (cps > ATT_FUSED_CPS) __trap();                              // the host selects the two-phase program beyond this
  const int c_lo = min(split * cps, n_chunks), c_hi = min(c_lo + cps, n_chunks);
  const int nck = c_hi - c_lo;                                    // 0, 1 or 2 chunks
  const int g = lane >> 2, t4 = lane & 3;
  const int key_lo = c_lo * ATT_CHUNK;
  const bf16* kbase = p.K + (long)kvh * p.kv_head_stride;
  const bf16* vbase = p.V + (long)kvh * p.kv_head_stride;
#ifdef VS_PROGRAM_CLK
  const bool aclk_on = blockIdx.x == 0 && tid == 0 && head == 0;
#endif
  ACLK(0);

  // ---- loads of pass 1
  att_load_rows_async(sm.qp, p.Q + (long)head * p.q_head_stride, 0, tq_rows, VS_TOK, tid);
#pragma unroll
  for (int j = 0; j < ATT_FUSED_CPS; ++j)
    if (j < nck)
      att_load_rows_async(sm.kv[j], kbase, key_lo + j * ATT_CHUNK, min(ATT_CHUNK, kv_len - key_lo - j * ATT_CHUNK),
                          ATT_CHUNK, tid);
  cp_async_commit();
  ACLK(12);
  unsigned long long vis[ATT_FUSED_CPS][4];
  {
#ifdef VS_PROGRAM_CLK
    if (aclk_on) g_clk[14] = (long long)(qb[0] + qb[1] + qb[2] + qb[3]);
#endif
    ACLK(13);
#pragma unroll
    for (int j = 0; j < ATT_FUSED_CPS; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        vis[j][i] = att_visible_mask(p, n_ctx, kv_len, (i >> 1) * 16 + g + (i & 1) * 8, qb[i], key_lo + j * ATT_CHUNK);
  }
  const float inv_div = sqrtf((float)VS_HEAD_DIM), r_div = 1.0f / inv_div;
  ACLK(1);
  cp_async_wait<0>();
  sync();
  ACLK(2);

  // ---- S = Q K^T: acc[j][mb][e]: query = mb*16 + g + (e>>1)*8, key = key_lo + 64*j + 8*warp + 2*t4 + (e&1)
  float acc[ATT_FUSED_CPS][2][4];
  const int kofs = warp * 8 + 2 * t4;
  float lmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int j = 0; j < ATT_FUSED_CPS; ++j) {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][a][e] = -INFINITY;
    if (j < nck) {
      float c[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 4; ++e) c[a][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t bfr[2];
        ldmatrix_x2(bfr, sm.kv[j] + att_swz256(warp * 8 + (lane & 7), kk * 2 + ((lane >> 3) & 1)));
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          uint32_t afr[4];
          const int mat = lane >> 3;
          ldmatrix_x4(afr, sm.qp + att_swz256(mb * 16 + (lane & 7) + (mat & 1) * 8, kk * 2 + (mat >> 1)));
          mma_bf16(c[mb], afr, bfr[0], bfr[1]);
        }
      }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int eh = 0; eh < 2; ++eh) {
          const int qi = mb * 2 + eh;
          float s0 = c[mb][eh * 2], s1 = c[mb][eh * 2 + 1];
          s0 = p.eager ? rbf(att_div(rbf(s0), inv_div, r_div)) : s0 * p.scale;
          s1 = p.eager ? rbf(att_div(rbf(s1), inv_div, r_div)) : s1 * p.scale;
          if (!((vis[j][qi] >> kofs) & 1ull)) s0 = -INFINITY;
          if (!((vis[j][qi] >> (kofs + 1)) & 1ull)) s1 = -INFINITY;
          acc[j][mb][eh * 2] = s0;
          acc[j][mb][eh * 2 + 1] = s1;
          lmax[qi] = fmaxf(lmax[qi], fmaxf(s0, s1));
        }
    }
  }
  ACLK(3);
  // ---- split statistics: every warp leaves (max, sum exp(s - max)) of its 8 key columns per row; 32 threads combine
  // the warps in fixed order (two block barriers in all)
  float lsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lmax[i] = fmaxf(lmax[i], __shfl_xor_sync(0xffffffffu, lmax[i], 1));
    lmax[i] = fmaxf(lmax[i], __shfl_xor_sync(0xffffffffu, lmax[i], 2));
  }
#pragma unroll
  for (int j = 0; j < ATT_FUSED_CPS; ++j)
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int eh = 0; eh < 2; ++eh) {
        const int qi = mb * 2 + eh;
        const float s0 = acc[j][mb][eh * 2], s1 = acc[j][mb][eh * 2 + 1];
        if (s0 != -INFINITY) lsum[qi] += expf(s0 - lmax[qi]);
        if (s1 != -INFINITY) lsum[qi] += expf(s1 - lmax[qi]);
      }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lsum[i] += __shfl_xor_sync(0xffffffffu, lsum[i], 1);
    lsum[i] += __shfl_xor_sync(0xffffffffu, lsum[i], 2);
    if (t4 == 0) {
      sm.red[warp][(i >> 1) * 16 + g + (i & 1) * 8] = lmax[i];
      sm.red2[warp][(i >> 1) * 16 + g + (i & 1) * 8] = lsum[i];
    }
  }
  sync();                                                          // also: every warp is done with Q and K in shared memory
  // the V chunks take the place of the K chunks; in flight during the statistics exchange
#pragma unroll
  for (int j = 0; j < ATT_FUSED_CPS; ++j)
    if (j < nck)
      att_load_rows_async(sm.kv[j], vbase, key_lo + j * ATT_CHUNK, min(ATT_CHUNK, kv_len - key_lo - j * ATT_CHUNK),
                          ATT_CHUNK, tid);
  cp_async_commit();
  ACLK(4);
  att_exchange_stats(p, sm, head, split, n_splits, head_cnt, epoch, tid, sync);
  ACLK(6);
  // ---- P in the reference's rounding, written over the Q tile: row q, key column 64*j + 8*warp + 2*t4 (+1)
#pragma unroll
  for (int j = 0; j < ATT_FUSED_CPS; ++j)
    if (j < nck) {
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int eh = 0; eh < 2; ++eh) {
          const int q = mb * 16 + g + eh * 8;
          const float gm = sm.gM[q], gl = sm.gL[q], gr = sm.gR[q];
          float pr[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float s = acc[j][mb][eh * 2 + e];
            float v = 0.f;
            if (s != -INFINITY) {
              const float ex = expf(s - gm);
              v = p.eager ? att_div(ex, gl, gr) : ex;
            }
            pr[e] = v;
          }
          const int kx = j * ATT_CHUNK + kofs;                      // even
          *reinterpret_cast<__nv_bfloat162*>(sm.qp + att_swz256(q, kx >> 3) + (kx & 7) * 2) = __floats2bfloat162_rn(pr[0], pr[1]);
        }
    }
  ACLK(7);
  cp_async_wait<0>();
  sync();
  ACLK(8);
  // ---- O += P V: warp w owns output columns 16w .. 16w+15
  float oacc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[a][b][e] = 0.f;
#pragma unroll
  for (int j = 0; j < ATT_FUSED_CPS; ++j)
    if (j < nck) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t afr[2][4];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int mat = lane >> 3;
          ldmatrix_x4(afr[mb], sm.qp + att_swz256(mb * 16 + (lane & 7) + (mat & 1) * 8, j * 8 + kk * 2 + (mat >> 1)));
        }
        uint32_t bfr[4];
        {
          const int mat = lane >> 3;
          ldmatrix_x4_trans(bfr, sm.kv[j] + att_swz256(kk * 16 + (lane & 7) + (mat & 1) * 8, warp * 2 + (mat >> 1)));
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          mma_bf16(oacc[mb][0], afr[mb], bfr[0], bfr[1]);
          mma_bf16(oacc[mb][1], afr[mb], bfr[2], bfr[3]);
        }
      }
    }
  ACLK(9);
  // ---- partial O exchange and distributed reduction over the splits that own keys
  att_exchange_output(p, sm, oacc, head, split, n_splits, max(1, (n_chunks + cps - 1) / cps), tq_rows, head_cnt, epoch, tid,
                      sync);
  ACLK(11);
}

// =================================================================================================================
// Streamed single-phase attention: the same (head, kv split) items, exchanges and rounding points as attn_fused_item,
// for ANY context length.  A split owns `cps` contiguous 64-key chunks; what the register-resident version keeps in
// registers goes through a 2-deep shared-memory ring instead:
//   pass 1: K chunks j+1, j+2 are in flight (cp.async groups) while S_j = Q K_j^T is computed; the masked, scaled scores
//           go to the caller's fp32 score scratch (L2-resident, each thread re-reads exactly what it wrote) and every
//           thread keeps an online (max, sum exp) of its key columns;
//   exchange of the split statistics per head (V chunks 0, 1 and the scores of chunk 0 already in flight);
//   pass 2: P_j in the reference's rounding from the re-read scores (chunk j+1 prefetched into registers) into one half
//           of the Q/P tile, O += P_j V_j with V chunks j+1, j+2 in flight;
//   partial O exchange and distributed reduction exactly as in attn_fused_item.
// K and V are read from HBM once, at full pipeline depth (2 x 16 KB per team, 64 KB per SM in flight).
// =================================================================================================================
template <class Sync>
__device__ __forceinline__ void attn_stream_item(const AttnArgs& p, AttnFusedSmem& sm, int head, int split, int n_splits,
                                                 unsigned int* head_cnt, unsigned int epoch, int tid, Sync sync) {
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int warp = tid >> 5, lane = tid & 31;
  const int n_ctx = p.n_ctx ? __ldcg(p.n_ctx) : p.n_ctx_host;
  const int kv_len = n_ctx + p.n_new;
  const int n_chunks = (kv_len + ATT_CHUNK - 1) / ATT_CHUNK;
  if (kv_len > p.score_ld) __trap();                              // the score scratch bounds the context
  const int cps = (n_chunks + n_splits - 1) / n_splits;           // chunks per split (contiguous)
  const int c_lo = min(split * cps, n_chunks), c_hi = min(c_lo + cps, n_chunks);
  const int nck = c_hi - c_lo;
  const int tq_rows = min(VS_TOK, p.Tq);
  const int g = lane >> 2, t4 = lane & 3;
  const int key_lo = c_lo * ATT_CHUNK;
  const bf16* kbase = p.K + (long)kvh * p.kv_head_stride;
  const bf16* vbase = p.V + (long)kvh * p.kv_head_stride;
  const int kofs = warp * 8 + 2 * t4;
  // this thread's score elements of a chunk: rows q(i) = (i>>1)*16 + g + (i&1)*8, key columns key0 + kofs, +1
  // (row q(i) = g + 8 i: one base pointer and one stride instead of four pointers -- the function must not spill)
  float* const srow0 = p.scores + ((long)head * VS_TOK + g) * p.score_ld + key_lo + kofs;
  const int srow_st = 8 * (int)p.score_ld;
#define SROW(i) (srow0 + (i) * srow_st)
  // eager mode: the scores are bf16 values already (rounded twice by the reference's arithmetic), so they are parked as
  // bf16 pairs -- half the L2 traffic of the round trip; element (row, key) sits at the same index of the scratch read as
  // a bf16 array.  SDPA mode keeps fp32.
  const bool half_scores = p.eager != 0;
  auto put_scores = [&](int i, int j, float s0, float s1) {
    float* at = SROW(i) + j * ATT_CHUNK;
    if (half_scores) {
      const long idx = at - p.scores;                              // even
      *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(p.scores) + idx) = __floats2bfloat162_rn(s0, s1);
    } else {
      *reinterpret_cast<float2*>(at) = make_float2(s0, s1);
    }
  };
  auto get_scores = [&](int i, int j) -> float2 {
    const float* at = SROW(i) + j * ATT_CHUNK;
    if (half_scores) {
      const long idx = at - p.scores;
      const unsigned int u = __ldcg(reinterpret_cast<const unsigned int*>(reinterpret_cast<const bf16*>(p.scores) + idx));
      return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
    }
    return __ldcg(reinterpret_cast<const float2*>(at));
  };

  // a 64-row chunk = 1024 16-byte pieces, 4 per thread: rows r0 + 16k, same 16-byte column; offsets computed once
  const int r0 = tid >> 4, ch0 = tid & 15;
  const int doff0 = att_swz256(r0, ch0);                           // rows r0 + 16 k share the swizzle: + 16 * 256 bytes each
  auto load_chunk = [&](uint8_t* dst, const bf16* base, int key0) {
    const bf16* src = base + (long)key0 * 128 + ch0 * 8;
    const int nv = kv_len - key0;                                  // valid rows (>= 64 for all but the last chunk)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool ok = r0 + 16 * k < nv;
      cp_async16(dst + doff0 + k * 16 * ATT_ROW_B, src + (ok ? (r0 + 16 * k) * 128 : 0), ok);
    }
  };
  // ---- pass 1 prologue: Q + K chunk 0 | K chunk 1
  att_load_rows_async(sm.qp, p.Q + (long)head * p.q_head_stride, 0, tq_rows, VS_TOK, tid);
  if (nck > 0) load_chunk(sm.kv[0], kbase, key_lo);
  cp_async_commit();
  if (nck > 1) load_chunk(sm.kv[1], kbase, key_lo + ATT_CHUNK);
  cp_async_commit();
  const float inv_div = sqrtf((float)VS_HEAD_DIM), r_div = 1.0f / inv_div;
  float rm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, rl[4] = {0.f, 0.f, 0.f, 0.f};

#pragma unroll 1
  for (int j = 0; j < nck; ++j) {
    const int key0 = key_lo + j * ATT_CHUNK;
    uint8_t* kt = sm.kv[j & 1];
    cp_async_wait<1>();
    sync();
    float c[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) c[a][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t bfr[2];
      ldmatrix_x2(bfr, kt + att_swz256(warp * 8 + (lane & 7), kk * 2 + ((lane >> 3) & 1)));
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        uint32_t afr[4];
        const int mat = lane >> 3;
        ldmatrix_x4(afr, sm.qp + att_swz256(mb * 16 + (lane & 7) + (mat & 1) * 8, kk * 2 + (mat >> 1)));
        mma_bf16(c[mb], afr, bfr[0], bfr[1]);
      }
    }
    sync();                                                        // every warp is done with this K tile
    if (j + 2 < nck) load_chunk(kt, kbase, key0 + 2 * ATT_CHUNK);
    cp_async_commit();
    // the statistics only feed the softmax denominator (a sum over thousands of terms): ex2.approx is accurate enough;
    // the probabilities themselves (pass 2) use expf
    if (key0 + ATT_CHUNK <= n_ctx) {
      // chunk inside the fully visible prefix (all but the last one or two): no masks, no -inf
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int eh = 0; eh < 2; ++eh) {
          const int qi = mb * 2 + eh;
          float s0 = c[mb][eh * 2], s1 = c[mb][eh * 2 + 1];
          s0 = p.eager ? rbf(att_div(rbf(s0), inv_div, r_div)) : s0 * p.scale;
          s1 = p.eager ? rbf(att_div(rbf(s1), inv_div, r_div)) : s1 * p.scale;
          put_scores(qi, j, s0, s1);
          const float mx = fmaxf(s0, s1);
          if (mx > rm[qi]) {                                       // rm == -inf: rl == 0 and exp(-inf) == 0
            rl[qi] *= __expf(rm[qi] - mx);
            rm[qi] = mx;
          }
          rl[qi] += __expf(s0 - rm[qi]) + __expf(s1 - rm[qi]);
        }
    } else {
#pragma unroll
      for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int eh = 0; eh < 2; ++eh) {
          const int qi = mb * 2 + eh;
          const int qrow = mb * 16 + g + eh * 8;
          const unsigned long long qbits = (!p.causal && p.bits) ? __ldcg(p.bits + min(qrow, tq_rows - 1)) : 0ull;
          const unsigned long long vis = att_visible_mask(p, n_ctx, kv_len, qrow, qbits, key0);
          float s0 = c[mb][eh * 2], s1 = c[mb][eh * 2 + 1];
          s0 = p.eager ? rbf(att_div(rbf(s0), inv_div, r_div)) : s0 * p.scale;
          s1 = p.eager ? rbf(att_div(rbf(s1), inv_div, r_div)) : s1 * p.scale;
          if (!((vis >> kofs) & 1ull)) s0 = -INFINITY;
          if (!((vis >> (kofs + 1)) & 1ull)) s1 = -INFINITY;
          put_scores(qi, j, s0, s1);
          const float mn = fmaxf(rm[qi], fmaxf(s0, s1));
          if (mn != -INFINITY) {
            float l = rl[qi] * __expf(rm[qi] - mn);
            if (s0 != -INFINITY) l += __expf(s0 - mn);
            if (s1 != -INFINITY) l += __expf(s1 - mn);
            rl[qi] = l;
            rm[qi] = mn;
          }
        }
    }
  }
  cp_async_wait<0>();
  sync();                                                          // (nck == 0: the Q load; otherwise nothing is pending)
  // ---- pass 2 prologue: V chunks 0 | 1 and the scores of chunk 0, in flight during the statistics exchange
  if (nck > 0) load_chunk(sm.kv[0], vbase, key_lo);
  cp_async_commit();
  if (nck > 1) load_chunk(sm.kv[1], vbase, key_lo + ATT_CHUNK);
  cp_async_commit();
  float2 snext[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) snext[i] = nck > 0 ? get_scores(i, 0) : make_float2(-INFINITY, -INFINITY);

  // ---- split statistics: lanes of a row quad, then the warps in fixed order
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, rm[i], o), lo = __shfl_xor_sync(0xffffffffu, rl[i], o);
      const float mn = fmaxf(rm[i], mo);
      const float a = rm[i] == -INFINITY ? 0.f : rl[i] * expf(rm[i] - mn);
      const float b = mo == -INFINITY ? 0.f : lo * expf(mo - mn);
      rl[i] = a + b;
      rm[i] = mn;
    }
    if (t4 == 0) {
      sm.red[warp][(i >> 1) * 16 + g + (i & 1) * 8] = rm[i];
      sm.red2[warp][(i >> 1) * 16 + g + (i & 1) * 8] = rl[i];
    }
  }
  sync();
  att_exchange_stats(p, sm, head, split, n_splits, head_cnt, epoch, tid, sync);

  // ---- pass 2: O += P V
  float oacc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[a][b][e] = 0.f;
#pragma unroll 1
  for (int j = 0; j < nck; ++j) {
    const int jj = j & 1;
    uint8_t* vt = sm.kv[jj];
    float2 cur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) cur[i] = snext[i];
    if (j + 1 < nck) {
#pragma unroll
      for (int i = 0; i < 4; ++i) snext[i] = get_scores(i, j + 1);
    }
    // P in the reference's rounding into half jj of the Q/P tile: row q, key column 64*jj + kofs (+1)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = g + 8 * i;
      const float sv[2] = {cur[i].x, cur[i].y};
      const float gm = sm.gM[q], gl = sm.gL[q], gr = sm.gR[q];
      float pr[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = 0.f;
        if (sv[e] != -INFINITY) {
          const float ex = expf(sv[e] - gm);
          v = p.eager ? att_div(ex, gl, gr) : ex;
        }
        pr[e] = v;
      }
      const int kx = jj * ATT_CHUNK + kofs;                          // even
      *reinterpret_cast<__nv_bfloat162*>(sm.qp + att_swz256(q, kx >> 3) + (kx & 7) * 2) = __floats2bfloat162_rn(pr[0], pr[1]);
    }
    cp_async_wait<1>();
    sync();                                                          // V chunk j landed, P tile complete
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t afr[2][4];
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const int mat = lane >> 3;
        ldmatrix_x4(afr[mb], sm.qp + att_swz256(mb * 16 + (lane & 7) + (mat & 1) * 8, jj * 8 + kk * 2 + (mat >> 1)));
      }
      uint32_t bfr[4];
      {
        const int mat = lane >> 3;
        ldmatrix_x4_trans(bfr, vt + att_swz256(kk * 16 + (lane & 7) + (mat & 1) * 8, warp * 2 + (mat >> 1)));
      }
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        mma_bf16(oacc[mb][0], afr[mb], bfr[0], bfr[1]);
        mma_bf16(oacc[mb][1], afr[mb], bfr[2], bfr[3]);
      }
    }
    sync();                                                          // every warp is done with this V tile and P half
    if (j + 2 < nck) load_chunk(vt, vbase, key_lo + (j + 2) * ATT_CHUNK);
    cp_async_commit();
  }
  cp_async_wait<0>();

  // ---- partial O exchange and distributed reduction over the splits that own keys
  att_exchange_output(p, sm, oacc, head, split, n_splits, max(1, (n_chunks + cps - 1) / cps), tq_rows, head_cnt, epoch, tid,
                      sync);
}
#undef SROW

}  // namespace vs
