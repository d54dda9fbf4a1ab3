// Device bodies of the decode attention (see attention.cu for the algorithm and the reference call sites).
// They are written for a "team" of 128 threads (4 warps) that owns one work item at a time, so the same code runs
//   * as a stand-alone kernel (attention.cu: team = thread block, sync = __syncthreads), and
//   * inside the persistent program kernel (program.cu: two teams per CTA, sync = a named barrier).
#pragma once
#include <math.h>

#include "kernels.h"

namespace vs {

constexpr int ATT_CHUNK = 64;      // keys per chunk
constexpr int ATT_THREADS = 128;   // one team
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
  float red[4][VS_TOK];
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

__device__ __forceinline__ bool att_visible(const AttnArgs& p, int n_ctx, int kv_len, int q_abs,
                                            unsigned long long qbits, int key) {
  if (key >= kv_len) return false;
  if (p.causal) return key < n_ctx + q_abs + 1;
  if (key < n_ctx) return true;
  int j = key - n_ctx;
  return j < 64 && ((qbits >> j) & 1ull);
}

// ---- pass 1: one (head, query block, key chunk) item: S = Q K^T, mask, scale, chunk max / sum -----------------
// `load_q`: the Q tile of (head, tqb) is not in sm.s.qs yet.
template <class Sync>
__device__ __forceinline__ void attn_scores_item(const AttnArgs& p, AttnSmem& sm, int head, int tqb, int tq_blocks,
                                                 int c, bool load_q, int tid, Sync sync) {
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int warp = tid >> 5, lane = tid & 31;
  const int n_ctx = p.n_ctx ? *((volatile const int*)p.n_ctx) : p.n_ctx_host;
  const int kv_len = n_ctx + p.n_new;
  const int t0 = tqb * VS_TOK;
  const int tq_rows = min(VS_TOK, p.Tq - t0);
  const int g = lane >> 2, t4 = lane & 3;

  sync();   // previous item finished with the shared tiles
  if (load_q) att_load_rows_async(sm.s.qs, p.Q + (long)head * p.q_head_stride, t0, tq_rows, VS_TOK, tid);

  // the four query rows this thread owns in the accumulator layout: mblk*16 + g (+8)
  unsigned long long qb[4];
  int qabs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int q = (i >> 1) * 16 + g + (i & 1) * 8;
    qabs[i] = t0 + q;
    qb[i] = (!p.causal && p.bits && q < tq_rows) ? __ldcg(p.bits + t0 + q) : 0ull;
  }

  const bf16* kbase = p.K + (long)kvh * p.kv_head_stride;
  float* sbase = p.scores + ((long)head * tq_blocks * VS_TOK + t0) * p.score_ld;
  float* stat_base = p.stats + (((long)head * tq_blocks + tqb) * p.max_chunks) * VS_TOK * 2;
  const float inv_div = sqrtf((float)VS_HEAD_DIM);

  const int key0 = c * ATT_CHUNK;
  att_load_rows_async(sm.s.ks, kbase, key0, min(ATT_CHUNK, kv_len - key0), ATT_CHUNK, tid);
  cp_async_commit();
  cp_async_wait<0>();
  sync();

  float acc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    uint32_t bfr[4];
    {
      int mat = lane >> 3;
      int r = warp * 16 + (lane & 7) + (mat >> 1) * 8;
      int ch = kk * 2 + (mat & 1);
      ldmatrix_x4(bfr, sm.s.ks + att_swz256(r, ch));
    }
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      uint32_t afr[4];
      int mat = lane >> 3;
      int r = mb * 16 + (lane & 7) + (mat & 1) * 8;
      int ch = kk * 2 + (mat >> 1);
      ldmatrix_x4(afr, sm.s.qs + att_swz256(r, ch));
      mma_bf16(acc[mb][0], afr, bfr[0], bfr[1]);
      mma_bf16(acc[mb][1], afr, bfr[2], bfr[3]);
    }
  }

  // scale / round / mask; element (mb, ng, e): query = mb*16 + g + (e>>1)*8, key = key0 + warp*16 + ng*8 + 2*t4 + (e&1)
  float lmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int ng = 0; ng < 2; ++ng)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int qi = mb * 2 + (e >> 1);
        int key = key0 + warp * 16 + ng * 8 + 2 * t4 + (e & 1);
        float s = acc[mb][ng][e];
        s = p.eager ? rbf(rbf(s) / inv_div) : s * p.scale;
        if (!att_visible(p, n_ctx, kv_len, qabs[qi], qb[qi], key)) s = -INFINITY;
        acc[mb][ng][e] = s;
        lmax[qi] = fmaxf(lmax[qi], s);
      }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lmax[i] = fmaxf(lmax[i], __shfl_xor_sync(0xffffffffu, lmax[i], 1));
    lmax[i] = fmaxf(lmax[i], __shfl_xor_sync(0xffffffffu, lmax[i], 2));
    if (t4 == 0) sm.red[warp][(i >> 1) * 16 + g + (i & 1) * 8] = lmax[i];
  }
  sync();
  if (tid < VS_TOK) sm.rowm[tid] = fmaxf(fmaxf(sm.red[0][tid], sm.red[1][tid]), fmaxf(sm.red[2][tid], sm.red[3][tid]));
  sync();
  float lsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int ng = 0; ng < 2; ++ng) {
      // the two keys (e&1) of one query row are adjacent: one 8-byte store
#pragma unroll
      for (int eh = 0; eh < 2; ++eh) {
        int qi = mb * 2 + eh;
        int q = mb * 16 + g + eh * 8;
        int key = key0 + warp * 16 + ng * 8 + 2 * t4;
        float s0 = acc[mb][ng][eh * 2], s1 = acc[mb][ng][eh * 2 + 1];
        if (s0 != -INFINITY) lsum[qi] += expf(s0 - sm.rowm[q]);
        if (s1 != -INFINITY) lsum[qi] += expf(s1 - sm.rowm[q]);
        *reinterpret_cast<float2*>(sbase + (long)q * p.score_ld + key) = make_float2(s0, s1);
      }
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
    float l = sm.red[0][tid] + sm.red[1][tid] + sm.red[2][tid] + sm.red[3][tid];
    stat_base[((long)c * VS_TOK + tid) * 2 + 0] = sm.rowm[tid];
    stat_base[((long)c * VS_TOK + tid) * 2 + 1] = l;
  }
}

// ---- pass 2: one (head, query block, kv split) item: P = softmax in the reference's rounding, O += P V ----------
// Splits own the chunks split, split + n_splits, ...; the last split to arrive reduces the partial tiles in split order.
template <class Sync>
__device__ __forceinline__ void attn_pv_item(const AttnArgs& p, AttnSmem& sm, int head, int tqb, int tq_blocks, int split,
                                             int n_splits, int tid, Sync sync) {
  const int kvh = head / (p.n_heads / p.n_kv_heads);
  const int warp = tid >> 5, lane = tid & 31;
  const int n_ctx = p.n_ctx ? *((volatile const int*)p.n_ctx) : p.n_ctx_host;
  const int kv_len = n_ctx + p.n_new;
  const int n_chunks = (kv_len + ATT_CHUNK - 1) / ATT_CHUNK;
  const int t0 = tqb * VS_TOK;
  const int tq_rows = min(VS_TOK, p.Tq - t0);
  const int g = lane >> 2, t4 = lane & 3;

  const float* sbase = p.scores + ((long)head * tq_blocks * VS_TOK + t0) * p.score_ld;
  const float* stat_base = p.stats + (((long)head * tq_blocks + tqb) * p.max_chunks) * VS_TOK * 2;

  sync();   // previous item finished with gM / gL / tiles
  // global row max / sum from the chunk statistics (4 lanes per query row, fixed order)
  {
    int q = tid >> 2, part = tid & 3;
    float m = -INFINITY;
    for (int c = part; c < n_chunks; c += 4) m = fmaxf(m, __ldcg(stat_base + ((long)c * VS_TOK + q) * 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    float l = 0.f;
    for (int c = part; c < n_chunks; c += 4) {
      float2 ml = __ldcg(reinterpret_cast<const float2*>(stat_base + ((long)c * VS_TOK + q) * 2));
      if (ml.x != -INFINITY) l += ml.y * expf(ml.x - m);
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    if (part == 0) { sm.gM[q] = m; sm.gL[q] = l; }
  }
  sync();

  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

  const bf16* vbase = p.V + (long)kvh * p.kv_head_stride;
  for (int c = split; c < n_chunks; c += n_splits) {
    const int key0 = c * ATT_CHUNK;
    sync();
    att_load_rows_async(sm.v.vsm, vbase, key0, min(ATT_CHUNK, kv_len - key0), ATT_CHUNK, tid);
    cp_async_commit();
    // probabilities for this chunk: [32][64] bf16, 128-byte rows, 16-byte chunks swizzled by row
    for (int i = tid; i < VS_TOK * (ATT_CHUNK / 4); i += ATT_THREADS) {
      int q = i >> 4, k4 = (i & 15) * 4;
      float4 s4 = __ldcg(reinterpret_cast<const float4*>(sbase + (long)q * p.score_ld + key0 + k4));
      float sv[4] = {s4.x, s4.y, s4.z, s4.w};
      bf16 pr4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pr = 0.f;
        if (sv[j] != -INFINITY && key0 + k4 + j < kv_len) {
          float e = expf(sv[j] - sm.gM[q]);
          pr = p.eager ? e / sm.gL[q] : e;
        }
        pr4[j] = f2bf(pr);
      }
      int ch = k4 >> 3;
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
        int mat = lane >> 3;
        int r = mb * 16 + (lane & 7) + (mat & 1) * 8;
        int ch = kk * 2 + (mat >> 1);
        ldmatrix_x4(afr[mb], sm.v.psm + r * 128 + ((ch ^ (r & 7)) << 4));
      }
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        uint32_t bfr[4];
        int mat = lane >> 3;
        int r = kk * 16 + (lane & 7) + (mat & 1) * 8;        // key row
        int ch = warp * 4 + pr * 2 + (mat >> 1);             // 8-wide d chunk
        ldmatrix_x4_trans(bfr, sm.v.vsm + att_swz256(r, ch));
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          mma_bf16(acc[mb][pr * 2 + 0], afr[mb], bfr[0], bfr[1]);
          mma_bf16(acc[mb][pr * 2 + 1], afr[mb], bfr[2], bfr[3]);
        }
      }
    }
  }

  // publish the partial O tile [32][128] fp32; the last split to arrive reduces in split order
  float* obase = p.opart + (((long)head * tq_blocks + tqb) * n_splits) * VS_TOK * 128;
  float* mine = obase + (long)split * VS_TOK * 128;
#pragma unroll
  for (int mb = 0; mb < 2; ++mb)
#pragma unroll
    for (int ng = 0; ng < 4; ++ng)
#pragma unroll
      for (int eh = 0; eh < 2; ++eh) {
        int q = mb * 16 + g + eh * 8;
        int d = warp * 32 + ng * 8 + 2 * t4;
        *reinterpret_cast<float2*>(mine + q * 128 + d) = make_float2(acc[mb][ng][eh * 2], acc[mb][ng][eh * 2 + 1]);
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
  for (int i = tid; i < VS_TOK * 32; i += ATT_THREADS) {
    int q = i >> 5, d4 = (i & 31) * 4;
    if (q >= tq_rows) continue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < n_splits; ++s) {
      float4 t = __ldcg(reinterpret_cast<const float4*>(obase + (long)s * VS_TOK * 128 + q * 128 + d4));
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (!p.eager) {
      float l = sm.gL[q];
      v.x /= l; v.y /= l; v.z /= l; v.w /= l;
    }
    bf16 o4[4] = {f2bf(v.x), f2bf(v.y), f2bf(v.z), f2bf(v.w)};
    *reinterpret_cast<uint2*>(p.O + (long)(t0 + q) * p.ldo + head * 128 + d4) = *reinterpret_cast<const uint2*>(o4);
  }
  if (tid == 0) p.counters[head * tq_blocks + tqb] = 0;
}

}  // namespace vs
