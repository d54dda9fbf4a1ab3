// Fused epilogues of the weight-streaming GEMM, shared by the kernels in gemm_tc.cu / gemm_wide.cu.
// Input: one fp32 accumulator tile ct[32 tokens][EPI_CT_LD] (128 weight rows) in shared memory.
#pragma once
#include "kernels.h"

namespace vs {

constexpr int EPI_CT_LD = 128 + 4;

struct Tile {
  const float* ct;   // [32][EPI_CT_LD] fp32, ct[t*EPI_CT_LD + r] = acc(token t, tile row r)
  int tile, n0, t0, tb_rows;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

// Every branch walks the 32 token rows of the tile in batches of EPI_UNROLL with all global loads of a batch issued before
// anything is consumed (clamped indices instead of predicates): the prompt-pass kernel runs this loop 8 times per work
// unit, and with one dependent load chain per row it was latency-bound (80 us per unit against 17 us of MMAs).
constexpr int EPI_UNROLL = 8;

template <int EPI, int NT>
__device__ __forceinline__ void epilogue(const GemmArgs& p, const Tile& tl, const int tid) {
  static_assert(NT % 128 == 0, "whole groups of one thread per tile row");
  constexpr int TG = NT / 128;                                    // thread groups; group g takes token rows g, g + TG, ...
  const int tg = tid >> 7;
  const int rows = tl.tb_rows;                                    // 1..32 valid token rows
  if (EPI == EPI_STORE || EPI == EPI_RESID) {
    const int r = tid & 127, n = tl.n0 + r;
    if (n >= p.N) return;
    const float bias = p.bias ? bf2f(p.bias[n]) : 0.f;
    for (int tb = tg; tb < rows; tb += EPI_UNROLL * TG) {
      float res[EPI_UNROLL];
      if (EPI == EPI_RESID) {
#pragma unroll
        for (int j = 0; j < EPI_UNROLL; ++j) res[j] = bf2f(p.R[(long)(tl.t0 + min(tb + j * TG, rows - 1)) * p.ldr + n]);
      }
#pragma unroll
      for (int j = 0; j < EPI_UNROLL; ++j) {
        const int t = tb + j * TG;
        if (t >= rows) break;
        float v = tl.ct[t * EPI_CT_LD + r] + bias;
        if (EPI == EPI_RESID) v = res[j] + rbf(v);                // Linear output rounds to bf16, then the add rounds again
        p.Y[(long)(tl.t0 + t) * p.ldy + n] = f2bf(v);
      }
    }
  } else if (EPI == EPI_SWIGLU) {
    // tile rows [0,64) = gate features j0..j0+63, rows [64,128) = the matching up features; thread = (feature, row parity)
    const int j0 = tl.tile * 64;
    const int r = tid & 63, half = tid >> 6;
    for (int t = half; t < rows; t += 2 * TG) {
      float g = rbf(tl.ct[t * EPI_CT_LD + r]);
      float u = rbf(tl.ct[t * EPI_CT_LD + 64 + r]);
      float a = rbf(silu_f(g));                                   // act_fn(gate) rounds, then * up rounds
      p.Y[(long)(tl.t0 + t) * p.ldy + j0 + r] = f2bf(a * u);
    }
  } else {  // EPI_HEADS
    const int tile = tl.tile;
    const bool is_q = tile < p.n_q_tiles;
    const bool is_k = !is_q && tile < p.n_q_tiles + p.n_k_tiles;
    const int head = is_q ? tile : (is_k ? tile - p.n_q_tiles : tile - p.n_q_tiles - p.n_k_tiles);
    const int row_base = (p.kv_row_base ? *p.kv_row_base : 0) + p.kv_row_off;
    const int d = tid & 127, d2 = d < 64 ? d + 64 : d - 64;
    const float b = p.bias ? bf2f(p.bias[tl.n0 + d]) : 0.f;
    const float b2 = p.bias ? bf2f(p.bias[tl.n0 + d2]) : 0.f;
    const bool rope = p.rope && (is_q || is_k);
    bf16* dst = is_q ? p.Qout + (long)head * p.q_head_stride + (long)tl.t0 * 128 + d
                     : (is_k ? p.Kout : p.Vout) + (long)head * p.kv_head_stride + (long)(row_base + tl.t0) * 128 + d;
    for (int tb = tg; tb < rows; tb += EPI_UNROLL * TG) {
      float c[EPI_UNROLL], s[EPI_UNROLL];
      if (rope) {
        long pos[EPI_UNROLL];
#pragma unroll
        for (int j = 0; j < EPI_UNROLL; ++j) pos[j] = p.pos_ids[tl.t0 + min(tb + j * TG, rows - 1)];
#pragma unroll
        for (int j = 0; j < EPI_UNROLL; ++j) {
          c[j] = bf2f(p.cos[pos[j] * 128 + d]);
          s[j] = bf2f(p.sin[pos[j] * 128 + d]);
        }
      }
#pragma unroll
      for (int j = 0; j < EPI_UNROLL; ++j) {
        const int t = tb + j * TG;
        if (t >= rows) break;
        const float x = rbf(tl.ct[t * EPI_CT_LD + d] + b);
        float out = x;
        if (rope) {
          const float x2 = rbf(tl.ct[t * EPI_CT_LD + d2] + b2);
          const float rot = d < 64 ? -x2 : x2;
          out = rbf(x * c[j]) + rbf(rot * s[j]);                  // three bf16 roundings (cnets_ours.py:117-118)
        }
        dst[(long)t * 128] = f2bf(out);
      }
    }
  }
}


}  // namespace vs
