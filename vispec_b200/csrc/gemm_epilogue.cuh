// Fused epilogues of the weight-streaming GEMM, shared by the kernels in gemm_skinny.cu / gemm_tc.cu.
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

template <int EPI, int NT>
__device__ __forceinline__ void epilogue(const GemmArgs& p, const Tile& tl, const int tid) {
  if (EPI == EPI_STORE || EPI == EPI_RESID) {
    for (int i = tid; i < VS_TOK * 128; i += NT) {
      int t = i >> 7, r = i & 127;
      int n = tl.n0 + r;
      if (t >= tl.tb_rows || n >= p.N) continue;
      float v = tl.ct[t * EPI_CT_LD + r];
      if (p.bias) v += bf2f(p.bias[n]);
      long row = tl.t0 + t;
      if (EPI == EPI_RESID) {
        float o = rbf(v);                                         // Linear output rounds to bf16 ...
        v = bf2f(p.R[row * p.ldr + n]) + o;                       // ... then the residual add rounds again
      }
      p.Y[row * p.ldy + n] = f2bf(v);
    }
  } else if (EPI == EPI_SWIGLU) {
    // tile rows [0,64) = gate features j0..j0+63, rows [64,128) = the matching up features
    const int j0 = tl.tile * 64;
    for (int i = tid; i < VS_TOK * 64; i += NT) {
      int t = i >> 6, r = i & 63;
      if (t >= tl.tb_rows) continue;
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
    for (int i = tid; i < VS_TOK * 128; i += NT) {
      int t = i >> 7, d = i & 127;
      if (t >= tl.tb_rows) continue;
      int n = tl.n0 + d;
      float b = p.bias ? bf2f(p.bias[n]) : 0.f;
      float x = rbf(tl.ct[t * EPI_CT_LD + d] + b);
      float out = x;
      if (p.rope && !(!is_q && !is_k)) {
        int d2 = d < 64 ? d + 64 : d - 64;
        float b2 = p.bias ? bf2f(p.bias[tl.n0 + d2]) : 0.f;
        float x2 = rbf(tl.ct[t * EPI_CT_LD + d2] + b2);
        float rot = d < 64 ? -x2 : x2;
        long pos = p.pos_ids[tl.t0 + t];
        float c = bf2f(p.cos[pos * 128 + d]), s = bf2f(p.sin[pos * 128 + d]);
        out = rbf(x * c) + rbf(rot * s);                          // three bf16 roundings (cnets_ours.py:117-118)
      }
      bf16 o = f2bf(out);
      if (is_q) {
        p.Qout[(long)head * p.q_head_stride + (long)(tl.t0 + t) * 128 + d] = o;
      } else {
        bf16* dst = is_k ? p.Kout : p.Vout;
        dst[(long)head * p.kv_head_stride + (long)(row_base + tl.t0 + t) * 128 + d] = o;
      }
    }
  }
}


}  // namespace vs
