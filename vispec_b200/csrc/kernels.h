// Internal launcher declarations shared by the .cu files.  The public C ABI is include/vispec_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace vs {

enum GemmEpilogue { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_HEADS = 3 };

// Y[T,N] = X[T,K] * W[N,K]^T (+bias) with a fused epilogue.  "Skinny": the token dimension is
// processed in blocks of 32 rows, the weight matrix is streamed once per token block.
struct GemmArgs {
  const bf16* X = nullptr;  long ldx = 0;   // [T][K]
  const bf16* W = nullptr;  long ldw = 0;   // [N][K]
  const bf16* bias = nullptr;               // [N] or null
  int T = 0, N = 0, K = 0;
  int splits = 1;                            // split-K factor (deterministic last-block reduction)
  float* partials = nullptr;                 // [tok_blocks][splits][n_tiles][32][128] fp32
  int* counters = nullptr;                   // [tok_blocks][n_tiles], zero before and after every launch
  int epi = EPI_STORE;
  bf16* Y = nullptr;  long ldy = 0;          // STORE / RESID / SWIGLU output
  const bf16* R = nullptr;  long ldr = 0;    // RESID: Y = R + bf16(acc)
  // HEADS epilogue: one 128-row tile == one attention head of q, k or v
  int n_q_tiles = 0, n_k_tiles = 0;          // remaining tiles are v heads
  int rope = 0;
  const bf16* cos = nullptr;                 // [max_pos][128]
  const bf16* sin = nullptr;
  const int* pos_ids = nullptr;              // [T] absolute positions
  bf16* Qout = nullptr;  long q_head_stride = 0;       // [n_q][rows][128]
  bf16* Kout = nullptr;  bf16* Vout = nullptr;  long kv_head_stride = 0;   // [n_kv][rows][128]
  const int* kv_row_base = nullptr;          // device int (may be null -> 0)
  int kv_row_off = 0;                        // row = *kv_row_base + kv_row_off + t
};

int gemm_pick_splits(int N, int K, int T);
size_t gemm_partials_floats(int N, int T, int splits);
cudaError_t launch_gemm(const GemmArgs& a, cudaStream_t st);

// rows ops -------------------------------------------------------------------------------------
cudaError_t launch_rmsnorm(const bf16* x, long ldx, const bf16* w, bf16* y, long ldy, int T, int H, float eps,
                           cudaStream_t st);
// dst[t] = src[idx ? idx[t] : t]  (row gather of H bf16 elements, idx on device)
cudaError_t launch_gather_rows(const bf16* src, long lds, const int* idx, bf16* dst, long ldd, int T, int H,
                               cudaStream_t st);
// dst[t] = table[ids[t]]  with int64 or int32 ids
cudaError_t launch_embed_rows(const bf16* table, long ldt, const int* ids, bf16* dst, long ldd, int T, int H,
                              int vocab, cudaStream_t st);
cudaError_t launch_broadcast_row(const bf16* row, bf16* dst, long ldd, int T, int H, cudaStream_t st);
cudaError_t launch_argmax_rows(const bf16* x, long ldx, int T, int V, int* out, cudaStream_t st);
// log-softmax (fp32 math, bf16 result) + top-k (k<=8) per row; row_sel (device) optionally picks a single row
cudaError_t launch_logsoftmax_topk(const bf16* logits, long ldl, int T, int V, int k, const int* row_sel,
                                   float* top_val, int* top_idx, cudaStream_t st);

// attention --------------------------------------------------------------------------------------
struct AttnArgs {
  const bf16* Q = nullptr;  long q_head_stride = 0;     // [nh][Tq rows][128]
  const bf16* K = nullptr;  const bf16* V = nullptr;  long kv_head_stride = 0;   // [nkv][rows][128]
  int n_heads = 0, n_kv_heads = 0;
  int Tq = 0;                                 // query rows (any; blocks of 32)
  const int* n_ctx = nullptr;                 // device int: fully visible prefix length (null -> n_ctx_host)
  int n_ctx_host = 0;
  int n_new = 0;                              // keys appended after the prefix: kv_len = n_ctx + n_new
  int causal = 0;                             // 1: query i sees keys < n_ctx + i + 1
  const unsigned long long* bits = nullptr;   // !causal: bit j of bits[i] -> key n_ctx + j visible to query i
  int eager = 0;                              // 1: target rounding points (modeling_llama_kv eager); 0: SDPA-style
  float scale = 0.f;
  float* scores = nullptr;  long score_ld = 0;          // [nh][Tq_pad][score_ld] fp32 workspace
  float* stats = nullptr;   int max_chunks = 0;         // [nh][tq_blocks][max_chunks][32][2]
  float* opart = nullptr;   int pv_splits = 0;          // [nh][tq_blocks][pv_splits][32][128]
  int* counters = nullptr;                               // [nh][tq_blocks]
  bf16* O = nullptr;  long ldo = 0;                       // [Tq][nh*128]
  // program kernel only: the caller asserts that *n_ctx and the K/V rows [0, *n_ctx) are final when the launch starts
  // (no earlier phase of the program changes them), so the KV producer may stage them ahead of the phase
  int prefix_stable = 0;
};
cudaError_t launch_attention(const AttnArgs& a, cudaStream_t st);

}  // namespace vs
