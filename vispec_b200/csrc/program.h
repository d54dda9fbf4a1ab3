// The persistent "program" kernel: a list of phases executed by ONE launch of one CTA per SM.
// Device-side phase descriptors (plain structs in global memory) shared by program.cu (kernel + builder).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "tree.h"
#include "tree_dev.cuh"

namespace vs {

enum PhaseType : int {
  PH_NOP = 0,
  PH_GEMM = 1,      // stream a weight matrix through TMA -> tcgen05; every CTA leaves fp32 partial tiles in L2
  PH_EPI = 2,       // reduce the partial tiles in k order + fused epilogue (store / residual / SiLU*up / RoPE + KV append)
  PH_NORM = 3,      // LlamaRMSNorm on T rows (optionally from per-tile sums of squares left by a residual epilogue)
  PH_ROWS = 4,      // row gather / embedding lookup / broadcast
  PH_ATTN_S = 5,    // attention pass 1: masked scores + chunk statistics
  PH_ATTN_PV = 6,   // attention pass 2: softmax in the reference's rounding, P V, split reduction
  PH_ARGMAX = 7,    // arg-max over the vocabulary: best key per (row, vocabulary slice); merged by the next tree phase
  PH_LSM_STATS = 8, // bf16 log-softmax + top-k, step 1: (max, sum exp) per (row, slice)
  PH_TREE = 9,      // integer tree bookkeeping (level0 / expand / build / accept / commit / merges) on CTA 0
  PH_SPLICE = 10,   // KV-cache splice
  PH_LSM_CAND = 11, // log-softmax + top-k, step 2: k best log-probs per (row, slice); merged by the next tree phase
  PH_ATTN_FUSED = 12, // both attention passes in one phase (every (head, split) item resident; per-head arrival counters)
  PH_TOPK_HIST = 14, // top-k logits warper, threshold search: 256-bin histogram of one byte of the order-preserving bf16 key
  PH_GEMM2 = 13      // slab GEMM (slab.h): every CTA streams whole output rows over the full K range and runs the fused
                     // epilogue (and, optionally, the following RMSNorm) straight from TMEM -- no partial tiles, no PH_EPI
};

enum TreeOp : int {
  TREE_LEVEL0 = 0, TREE_EXPAND = 1, TREE_BUILD = 2, TREE_ACCEPT = 3, TREE_COMMIT = 4,
  TREE_MERGE_ARGMAX = 5,   // only the merge of PH_ARGMAX keys -> amax_out
  TREE_MERGE_TOPK = 6,     // only the merge of PH_LSM_CAND keys -> top_val / top_idx
  TREE_ACCEPT_SAMPLE = 7   // T > 0 acceptance + next-token sampling (needs the PH_LSM_STATS of the verify logits)
};

struct Phase {
  int type;
  int T;                      // valid token rows (<= 32)
  // ---- PH_GEMM / PH_EPI
  int tmW, tmX;               // tensor-map table indices
  const bf16* Wt;             // weights in the tiled layout (see vs_program_register_tiled) or null -> tensor map tmW
  int n_tiles, nk, N;         // ceil(N/128) weight-row tiles, K/64 chunks
  int maxseg;                 // partial tiles per CTA
  int bars_before;            // grid barriers executed before this phase (every phase; set at commit)
  int phase_idx;              // position of the phase in the program (set at commit): part of the tag of tagged exchanges
  int no_barrier;             // PH_GEMM: the epilogue phase that follows waits on per-tile counters instead of a grid barrier
  unsigned int* tile_cnt;     // PH_GEMM / PH_EPI: [n_tiles] arrivals of contributors, monotonic across launches (workspace)
  int range_off;              // offset into the program's int tables: per-CTA chunk ranges [lo, hi)
  int per_lo, rem;            // balanced partition: CTA b owns per_lo (+1 if b < rem) consecutive chunks
  unsigned int magic_per_hi, magic_per_lo, magic_nk;   // floor(2^32/d)+1 for d = per_lo+1, per_lo, nk (0 when d == 1)
  float* partials;            // [grid][maxseg][32 tokens][128 rows] fp32
  int epi;
  int tpi;                    // PH_EPI: tokens per work item (4 or 8)
  const bf16* bias;
  bf16* Y; long ldy;
  const bf16* R; long ldr;
  float* ssq;                 // residual epilogue: [n_tiles][32] sums of squares of the new rows (or null)
  unsigned int* norm_cnt;     // PH_EPI with a folded-in RMSNorm (`w` != null): [4] arrivals per group of 8 token rows, monotonic
                              // (PH_GEMM2: [1] arrivals of all CTAs)
  // ---- PH_GEMM2 (slab layout, slab.h); Wt = the slab image, nk / N / T / tmX3 / the epilogue fields as above
  int s_pair, s_ntiles, s_passes, s_rows_alloc, s_blk_bytes, s_blk_smem, s_kc, s_xoff, s_eper, s_erem;
  int s_packed, s_m64;        // SlabPlan::packed / m64
  long s_unit_stride;
  int tmX3;                   // 3-D tensor map of the activations: box {64 columns, 32 rows, kc chunks} (-1: kc 2-D loads)
                              // (`ssq` of a PH_GEMM2 residual epilogue is [grid][32]: one row of partial sums per CTA)
  int n_q_tiles, n_k_tiles, rope;
  const bf16* cos; const bf16* sin; const int* pos_ids;
  bf16* Qout; long q_head_stride;
  bf16* Kout; bf16* Vout; long kv_head_stride;
  const int* kv_row_base; int kv_row_off;
  // ---- PH_NORM / PH_ROWS
  const bf16* src; long lds;
  const int* idx; int n_src_rows; int broadcast;
  bf16* dst; long ldd;
  int H;
  const bf16* w; float eps;
  const float* ssq_in; int ssq_tiles;
  // ---- PH_ATTN_*
  AttnArgs attn;
  int tmK, tmV;               // PH_ATTN_FUSED: tensor maps of the K / V rows (attn_ring.cuh granules)
  int kv_two_loads;           // the driver refused the 4-D {64, rows, 2, heads} view: two 3-D loads per granule
  int pf_bars;                // grid barriers after which the KV producer may start staging this phase's granules
  // ---- PH_ARGMAX / PH_LSM_TOPK
  const bf16* logits; long ldl; int V; int k;
  const int* row_sel; float* top_val; int* top_idx; int* amax_out;
  int n_slices;               // vocabulary slices per row
  unsigned long long* keys;   // PH_ARGMAX: [T][n_slices]; PH_LSM_CAND: [T][n_slices][8]  (program workspace)
  float2* lstats;             // PH_LSM_STATS: [T][n_slices] (max, sum exp)                 (program workspace)
  float inv_temp;             // PH_LSM_STATS / TREE_ACCEPT_SAMPLE: logits are scaled by 1 / temperature (0 -> 1)
  int top_k;                  // PH_TOPK_HIST / PH_LSM_STATS / TREE_ACCEPT_SAMPLE: TopKLogitsWarper (0 = off): only logits >= the
                              // k-th largest of their row count
  unsigned int* hist;         // [T][2][256] histograms of the high / low key byte (program workspace; zero between launches)
  int hist_level;             // PH_TOPK_HIST: 0 = high byte of every key, 1 = low byte of the keys in the selected high bin
  const float* uni_override;  // TREE_ACCEPT_SAMPLE: test hook, see vs_program_add_accept_sampling
  // ---- PH_TREE
  int tree_op, level;
  TreeBufs tb; int* state; const int* amax; int* out_rec; const int* sample_token_dev; int sample_token_host;
  // ---- PH_SPLICE
  SpliceArgs splice; int n_items;
};

struct ProgramHeader {
  const Phase* phases;
  int n_phases;
  const CUtensorMap* maps;
  const int* tables;               // host-computed chunk ranges / contributor tables of the GEMM phases
  unsigned int* counters;          // [0] grid barrier, [1] exit counter, [2] launch epoch (launches completed)
  unsigned long long* trace;       // [n_phases] globaltimer at the end of each phase (CTA 0) or null
  unsigned long long* trace_cta;   // [n_phases][grid] globaltimer when each CTA finished its items of the phase, or null
  int pf_ahead;                    // L2 prefetch distance in 16 KB weight chunks per CTA (0 = off)
  int inflight;                    // weight stages issued-but-not-landed per CTA (0 = whole ring)
  unsigned int spin_ns;            // back-off of the role warps between barrier polls
  int ns, stage_bytes;             // ring geometry: 6 x 20480 (stream-K programs) or 4 x 32768 (programs with slab GEMMs)
  int attn_pf;                     // L2 prefetch distance of the KV producer, in 8 KB granules per team (attn_ring.cuh)
  int debug;                       // ablation switches for tools/bench_slab_phases.py (results are garbage): 1 = no activation
                                   // loads, 2 = no weight loads, 4 = no MMAs, 8 = workers back off while the GEMM streams
};

// ---- host-side phase constructors (program.cu); the C API and the composite builders live in program_api.cu ----
}  // namespace vs
struct vs_program;
namespace vs {
// `ssq` / `ssq_in` == PROG_WS_SSQ: the sums-of-squares buffer lives in the program's own workspace (patched at commit)
#define PROG_WS_SSQ (reinterpret_cast<float*>(uintptr_t(1)))
int prog_add_linear(vs_program* pg, const GemmArgs& a, float* ssq);
int prog_add_norm(vs_program* pg, const bf16* x, long ldx, const bf16* w, bf16* y, long ldy, int T, int H, float eps,
                  const float* ssq_in, int ssq_tiles);
// true when the last phase of the program is a residual epilogue that wrote exactly the rows x[T][H] and left their sums
// of squares in the program's ssq workspace
bool prog_last_is_residual_of(const vs_program* pg, const bf16* x, long ldx, int T, int H);
int prog_add_rows(vs_program* pg, const bf16* src, long lds, const int* idx, int n_src_rows, int broadcast, bf16* dst,
                  long ldd, int T, int H);
int prog_add_attention(vs_program* pg, const AttnArgs& a);
#define PROG_WS_KEYS (reinterpret_cast<unsigned long long*>(uintptr_t(1)))
#define PROG_WS_LSTATS (reinterpret_cast<float2*>(uintptr_t(1)))
// The vocabulary scans leave per-slice results in the workspace; the merge runs on CTA 0 either as its own tree phase
// (`fuse` == null) or at the start of the tree phase that consumes the result (pass the MergeInfo to prog_add_tree).
struct MergeInfo { int T, n_slices, k; float* top_val; int* top_idx; int* amax_out; };
int prog_add_argmax(vs_program* pg, const bf16* logits, long ldl, int T, int V, int* out, MergeInfo* fuse);
int prog_add_lsm_topk(vs_program* pg, const bf16* logits, long ldl, int T, int V, int k, const int* row_sel, float* top_val,
                      int* top_idx, MergeInfo* fuse);
int prog_add_tree(vs_program* pg, int op, const TreeBufs& tb, int* state, int level, const int* amax, int* out_rec,
                  const int* sample_token_dev, int sample_token_host, const MergeInfo* merge);
int prog_add_accept_sampling(vs_program* pg, const TreeBufs& tb, int* state, const bf16* logits, long ldl, int V,
                             float temperature, const float* uni_override, int* out_rec);
int prog_add_splice(vs_program* pg, bf16* slab, int n_slabs, int n_kv, long max_pos, const int* state, const int* sel_dev);
int prog_fail(int code, const char* where, const char* what);
}  // namespace vs
