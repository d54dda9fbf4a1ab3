/* vispec_b200 -- C ABI of the B200-native draft-and-verify decode path of KangJialiang/ViSpec.
 *
 * Every entry point is `extern "C"`, takes plain device/host pointers and sizes, is stream-ordered on
 * the `stream` argument (a cudaStream_t passed as void*), never allocates or frees caller memory,
 * and returns 0 on success or a negative code (message via vs_last_error()).  All matrices are
 * bf16 row-major, all integer device buffers are int32 unless stated, head_dim must be 128.
 *
 * Each function names the reference interface it replaces (paths relative to the reference repo,
 * commit d0240c87).  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 */
#ifndef VISPEC_B200_H
#define VISPEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_ABI_VERSION 2
#define VS_TOK_BLOCK 32    /* token rows per block of the skinny kernels; decode paths use <= 32 rows */
#define VS_RETRIEVE_LD 8   /* row stride of retrieve_indices; tree depth + 2 must be <= 8 */
#define VS_MAX_ACCEPT 8    /* accepted path length (accept_length + 1) must be <= 8 */

enum vs_status { VS_OK = 0, VS_ERR_INVALID = -1, VS_ERR_CUDA = -2, VS_ERR_UNSUPPORTED = -3 };

/* Device-resident int32 cycle state (one array of VS_ST_SLOTS ints, owned by the caller).  It plays the
 * role of kv_cache.py:130-132 `current_length_data` (CPU tensor in the reference), of the Python ints
 * accept_length / best_candidate (utils.py:443-451) and of the draft's `stable_kv[0][2]` length counter
 * (cnets_ours.py:416-418,1108), so that a whole cycle needs no host round trip. */
enum vs_state_slot {
  VS_ST_CUR_LEN = 0,    /* committed target KV length (== len(input_ids)) */
  VS_ST_PREV_LEN = 1,   /* value of CUR_LEN before the last accept */
  VS_ST_ACCEPT = 2,     /* accept_length of the last cycle */
  VS_ST_BEST = 3,       /* best_candidate (row of retrieve_indices) */
  VS_ST_NEXT_TOKEN = 4, /* argmax(sample_p): root of the next tree */
  VS_ST_D_LEN = 5,      /* rows of the draft's stable KV */
  VS_ST_D_POS = 6,      /* next absolute position of the draft (stable_kv real_len) */
  VS_ST_M = 7,          /* rows of the pending stable forward (accept_length + 1) */
  VS_ST_N_LEAVES = 8,   /* rows of retrieve_indices */
  VS_ST_MAX_DEPTH = 9,  /* used columns of retrieve_indices */
  VS_ST_LAST_ROW = 10,  /* row of the stable forward whose hidden feeds the LM head */
  VS_ST_CYCLE = 11,     /* cycles completed */
  VS_ST_TEMP_BITS = 12, /* T > 0 only: float bits of the temperature (0 = greedy); tree build sorts the leaf paths
                           when it is non-zero (cnets_ours.py:1215-1224) */
  VS_ST_RNG_SEED = 13,  /* T > 0 only: Philox key of the acceptance / sampling uniforms */
  VS_ST_RNG_CTR = 14,   /* T > 0 only: draws so far (advanced by one per cycle) */
  VS_ST_ROPE_DELTA = 15,/* added to the verify positions: Qwen2.5-VL `rope_deltas` (utils.py:398-402); 0 otherwise */
  VS_ST_SLOTS = 16
};

const char* vs_last_error(void);
int vs_abi_version(void);

/* ------------------------------------------------------------------------------------------------
 * Fused linear: Y[T,N] = X[T,K] W[N,K]^T (+bias) with an epilogue.  Up to 64 rows: weight-streaming kernel (tcgen05 MMA
 * N = 32, split-K); more rows (prompt passes): tcgen05 MMA N = 256 token tiles, no split-K (`splits` is ignored).  Replaces nn.Linear + the elementwise
 * ops that follow it: cnets_ours.py:363-365 (+RoPE :111-119, +KV append :395-396), :461, :508, :921-922,
 * :982-988, :1111; modeling_llama_kv.py:567-569 (+RoPE :584-587, +KVCache.cat kv_cache.py:52-66), :648, :421, :1196.
 * ------------------------------------------------------------------------------------------------ */
enum vs_epilogue {
  VS_EPI_STORE = 0,  /* Y = bf16(acc + bias) */
  VS_EPI_RESID = 1,  /* Y = bf16(R + bf16(acc + bias)) */
  VS_EPI_SWIGLU = 2, /* W rows interleaved 64 gate / 64 up per 128-row tile; Y[T, N/2] = bf16(bf16(silu(g)) * u) */
  VS_EPI_HEADS = 3   /* 128-row tile = one head of q|k|v; optional RoPE; q -> Qout, k/v -> KV rows in place */
};

typedef struct vs_linear_args {
  const void* X; int64_t ldx;
  const void* W; int64_t ldw;
  const void* bias;
  int32_t T, N, K;
  int32_t splits;            /* from vs_linear_splits(); >1 needs partials + counters */
  float* partials;           /* vs_linear_partials_floats() floats */
  int32_t* counters;         /* >= ceil(T/32)*ceil(N/128) ints, zero-initialised; left zero by every call */
  int32_t epi;
  void* Y; int64_t ldy;
  const void* R; int64_t ldr;
  int32_t n_q_tiles, n_k_tiles;
  int32_t rope;
  const void* cos; const void* sin;   /* [max_pos][128] bf16 tables, fp32 math cast once (cnets_ours.py:129-155) */
  const int32_t* pos_ids;             /* [T] absolute positions (device) */
  void* Qout; int64_t q_head_stride;
  void* Kout; void* Vout; int64_t kv_head_stride;
  const int32_t* kv_row_base;         /* device int or NULL */
  int32_t kv_row_off;
} vs_linear_args;

int vs_linear_splits(int N, int K, int T);
size_t vs_linear_partials_floats(int N, int T, int splits);
int vs_linear(const vs_linear_args* a, void* stream);

/* LlamaRMSNorm (cnets_ours.py:513-527, modeling_llama_kv.py:105-134) on T rows. */
int vs_rmsnorm(const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int T, int H, float eps, void* stream);
/* dst[t] = src[idx[t]] (idx NULL: identity); embedding lookup is the same with src = table. */
int vs_gather_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int T, int H, void* stream);
int vs_embed_rows(const void* table, int64_t ldt, const int32_t* ids, void* dst, int64_t ldd, int T, int H, int vocab,
                  void* stream);
int vs_broadcast_row(const void* row, void* dst, int64_t ldd, int T, int H, void* stream);
/* argmax over the vocabulary, first index at ties (utils.py:290,440,554). */
int vs_argmax_rows(const void* logits, int64_t ldl, int T, int V, int32_t* out, void* stream);
/* bf16 LogSoftmax + top-k, k <= 8 (cnets_ours.py:1111-1115,1145-1148); row_sel (device int, may be NULL)
 * selects one row, in which case T must be 1. */
int vs_logsoftmax_topk(const void* logits, int64_t ldl, int T, int V, int k, const int32_t* row_sel, float* top_val,
                       int32_t* top_idx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention over a head-major KV slab: fully visible prefix [0, n_ctx) + n_new appended keys that are
 * causal or masked by one 64-bit word per query.  Replaces modeling_llama_kv.py:602-623 + the mask
 * build :890-924 (eager = 1), cnets_ours.py:428-433 + :781-815 and ImgAdaptor SDPA :649 (eager = 0).
 * ------------------------------------------------------------------------------------------------ */
typedef struct vs_attn_args {
  const void* Q; int64_t q_head_stride;
  const void* K; const void* V; int64_t kv_head_stride;
  int32_t n_heads, n_kv_heads;
  int32_t Tq;
  const int32_t* n_ctx; int32_t n_ctx_host;
  int32_t n_new;
  int32_t causal;
  const uint64_t* bits;
  int32_t eager;
  float scale;
  float* scores; int64_t score_ld;      /* [n_heads][ceil32(Tq)][score_ld], score_ld >= ceil64(kv_len) */
  float* stats; int32_t max_chunks;     /* [n_heads][ceil(Tq/32)][max_chunks][32][2], max_chunks >= ceil(kv_len/64) */
  float* opart; int32_t pv_splits;      /* [n_heads][ceil(Tq/32)][pv_splits][32][128] */
  int32_t* counters;                    /* [n_heads][ceil(Tq/32)] zero-initialised */
  void* O; int64_t ldo;
  int32_t prefix_stable;                /* programs only: *n_ctx and the K/V rows [0, *n_ctx) are final when the launch starts
                                           (nothing earlier in the program moves them): the kernel's KV producer may stage
                                           those rows ahead of the phase.  0 = load everything inside the phase. */
} vs_attn_args;
int vs_attention(const vs_attn_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Draft-tree bookkeeping, tree build (A7), greedy accept (A8), KV splice (A9).
 * ------------------------------------------------------------------------------------------------ */
typedef struct vs_tree_bufs {
  int32_t top_k, depth, total;        /* total = total_tokens - 1 drafted nodes (cnets_ours.py:733) */
  float* scores_flat;                 /* [top_k + depth*top_k^2] cumulative log-probs (bf16 values) */
  int32_t* tokens_flat;               /* same length */
  int32_t* parents;                   /* [1 + depth*top_k] */
  float* frontier_scores;             /* [top_k] */
  int32_t* frontier_ids;              /* [top_k] token ids fed to the next draft level */
  int32_t* cs_index;                  /* [top_k] */
  uint64_t* frontier_bits;            /* [32] draft tree-mask rows, bit j = key (stable_len + j) */
  float* top_val;                     /* [top_k][top_k] per-row top-k log-probs of the current level */
  int32_t* top_idx;                   /* [top_k][top_k] */
  int32_t* gather_idx;                /* [32] rows of the previous level's output feeding the next level */
  int32_t* draft_pos_ids;             /* [32] */
  int32_t* draft_tokens;              /* [total+1] tree tokens, slot 0 = root */
  uint64_t* tree_bits;                /* [total+1] ancestor masks (tree_mask rows) */
  int32_t* tree_pos;                  /* [total+1] tree_position_ids */
  int32_t* verify_pos;                /* [total+1] tree_pos + CUR_LEN */
  int32_t* retrieve;                  /* [total+1][VS_RETRIEVE_LD] leaf->root paths, -1 padded */
  int32_t* sel;                       /* [VS_MAX_ACCEPT] tree slots of the accepted path */
  int32_t* stable_ids;                /* [VS_MAX_ACCEPT] ids of the next stable forward */
  int32_t* stable_pos;                /* [VS_MAX_ACCEPT] positions of the next stable forward */
} vs_tree_bufs;

/* cnets_ours.py:1109-1123 (level 0) and :1126-1165 (level i bookkeeping after its forward). */
int vs_draft_level0(const vs_tree_bufs* tb, int32_t* state, void* stream);
int vs_draft_expand(const vs_tree_bufs* tb, int32_t* state, int level, void* stream);
/* cnets_ours.py:1167-1238: top-`total` selection, parent lookup, ancestor closure, leaf paths. */
int vs_tree_build(const vs_tree_bufs* tb, int32_t* state, const int32_t* sample_token_dev, int sample_token_host,
                  void* stream);
/* utils.py:438-451 (+ gathers :411, :516-526, :543-554).  amax[total+1] = argmax of the verify logits per
 * tree slot.  out_rec (device, 2 + VS_MAX_ACCEPT ints, may be NULL) = {accept_length, next_token, accepted tokens}. */
int vs_accept_greedy(const vs_tree_bufs* tb, int32_t* state, const int32_t* amax, int32_t* out_rec, void* stream);
/* utils.py:529-541 on the reference slab layout [n_slabs][1][n_kv][max_pos][128].  With state != NULL the
 * indices come from the device (state + sel_dev); otherwise from the host arguments. */
int vs_kv_splice(void* slab, int n_slabs, int n_kv, int64_t max_pos, const int32_t* state, const int32_t* sel_dev,
                 int prev_len_host, int n_host, const int32_t* sel_host, void* stream);
int vs_set_state(int32_t* state, int slot, int value, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Composite steps.
 * ------------------------------------------------------------------------------------------------ */
typedef struct vs_dims {
  int32_t hidden, inter, n_heads, n_kv_heads, vocab, n_layers;
  float eps;
} vs_dims;

typedef struct vs_layer {              /* one Llama decoder layer, bf16 device pointers */
  const void* ln_in;                   /* [H] or NULL (draft layer 0 has no input norm, cnets_ours.py:537,572) */
  const void* w_qkv; const void* b_qkv;/* [(nh+2nkv)*128][H] rows q|k|v, bias or NULL */
  const void* w_o;                     /* [H][nh*128] */
  const void* ln_post;                 /* [H] */
  const void* w_gateup;                /* [2I][H], 64 gate rows then 64 up rows per 128-row tile */
  const void* w_down;                  /* [H][I] */
  void* k_rows; void* v_rows;          /* this layer's KV: [n_kv][rows][128] */
  int64_t kv_head_stride;
} vs_layer;

typedef struct vs_scratch {            /* caller-owned work buffers for a forward of up to max_rows rows */
  int32_t max_rows;                    /* multiple of 32; decode scratch uses 32, prefill scratch is sized to the prompt */
  void* h;                             /* [max_rows][H] residual stream / layer output */
  void* x;                             /* [max_rows][H] */
  void* q;                             /* [nh][max_rows][128] */
  void* attn;                          /* [max_rows][nh*128] */
  void* act;                           /* [max_rows][I] */
  float* partials; int64_t partials_floats;
  int32_t* counters;                   /* >= (max_rows/32) * ceil(max N / 128) ints, zero */
  float* scores; int64_t score_ld;     /* [nh][max_rows][score_ld] */
  float* stats; int32_t max_chunks;    /* [nh][max_rows/32][max_chunks][32][2] */
  float* opart; int32_t pv_splits;     /* [nh][max_rows/32][pv_splits][32][128] */
  int32_t* attn_counters;              /* >= nh * max_rows/32 ints, zero */
} vs_scratch;

typedef struct vs_rope { const void* cos; const void* sin; } vs_rope;

/* One decoder layer on T <= scratch->max_rows rows held in scratch->h (updated in place).
 * Target: modeling_llama_kv.py:681-760.  Draft: cnets_ours.py:545-600. */
int vs_decoder_layer(const vs_dims* d, const vs_layer* l, const vs_scratch* s, const vs_rope* rope, int T,
                     const int32_t* pos_ids, const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx,
                     int n_ctx_host, int n_new, int causal, const uint64_t* bits, int eager, void* stream);

typedef struct vs_target {
  vs_dims dims;
  const vs_layer* layers;              /* host array [n_layers] */
  const void* embed;                   /* [V][H] */
  const void* norm;                    /* [H] */
  const void* lm_head;                 /* [V][H] */
  vs_rope rope;
  int32_t sdpa_attention;              /* 0: eager softmax with the reference's bf16 rounding points (Llama / LLaVA,
                                          modeling_llama_kv.py:602-623); 1: SDPA semantics (Qwen2.5-VL default attention
                                          class, modeling_qwen2_5_vl_kv.py:1065-1170) */
} vs_target;

/* A6: tree-attention verify of T tree tokens (utils.py:389-412 tree_decoding -> modeling_llama_kv.py:927-1080,
 * :1196-1197).  Appends T KV rows at CUR_LEN (not committed), writes post-norm hidden [32][H] bf16,
 * logits [32][V] bf16 (the reference's fp32 logits are these values cast) and amax[T]. */
int vs_verify_step(const vs_target* t, const vs_scratch* s, const int32_t* token_ids, const int32_t* pos_ids,
                   const uint64_t* bits, int T, int32_t* state, void* hidden_out, void* logits, int32_t* amax,
                   void* stream);

typedef struct vs_draft {
  vs_dims dims;                        /* n_layers = 1 */
  vs_layer layer;
  const void* embed;                   /* [V][H] */
  const void* w_fc; const void* b_fc;  /* [H][2H] */
  const void* w_img_fc; const void* b_img_fc;
  const void* lm_head;                 /* the target's head (spec_model_ours.py passes base_model.lm_head) */
  vs_rope rope;
  void* x1;                            /* [32][2H]: cat(hidden, g); the g half is written once per prompt */
  void* x2;                            /* [32][2H]: cat(embed(ids), img_fc(x1)) */
  void* logits;                        /* [32][V] */
} vs_draft;

/* A1: the image/prefill branch of the draft model in ONE call -- `ImgAdaptor.forward` (cnets_ours.py:603-661: K,V =
 * Linear(image embeds), `num_q` learned queries cross-attend, o_proj) + the prefill packing of `Model.forward`
 * (cnets_ours.py:879-975: text rows through fc(cat(emb, img_fc(cat(hidden, g)))), the image run compressed to
 * num_q - 1 tokens at the run's LAST positions, the global feature g = last adaptor row) + the draft decoder layer over
 * the packed rows (KV rows [0, n_c) of the draft cache).  `hidden` / `emb_shifted`: [L][H] target features and input
 * embeddings ALREADY shifted by one (cnets_ours.py:1066-1082); the image rows are rows [img_start, img_end) of that
 * shifted frame (exactly one contiguous run -- the reference fails on more, cnets_ours.py:938-941,1021), or
 * img_start = -1 for a text-only prompt (g = 0, cnets_ours.py:977-981).  All buffers are the caller's: nothing is
 * allocated, no host sync.  On return: g_out [H] and the g half of dm->x1, last_row_out [H] = the layer output of the
 * last packed row, state[VS_ST_D_LEN] = n_c, state[VS_ST_D_POS] = L, state[VS_ST_LAST_ROW] = 0, ws->pos = packed position
 * ids, scratch->h = the layer output [n_c][H].  Returns n_c (>= 1) or a negative VS_ERR_*. */
typedef struct vs_adaptor {
  const void* w_kv; const void* b_kv;  /* [2H][H]: k_proj rows then v_proj rows; bias or NULL */
  const void* w_o;                     /* [H][H] */
  const void* q;                       /* [n_heads][32][128]: the learned queries head-major, rows >= num_q zero */
  int32_t num_q;
} vs_adaptor;
typedef struct vs_prefill_ws {
  int32_t max_rows;                    /* multiple of 32; >= image rows and >= packed rows; <= scratch->max_rows */
  void* img_k; void* img_v;            /* [n_heads][max_rows][128] */
  void* x1; void* x2;                  /* [max_rows][2H] */
  void* fused;                         /* [max_rows][H] */
  void* adapted; void* attn_out;       /* [32][H] */
  int32_t* pos;                        /* [max_rows] */
  void* packed;                        /* [max_rows][H]: copy of the packed rows before the layer, or NULL */
} vs_prefill_ws;
int vs_img_adaptor_prefill(const vs_draft* dm, const vs_adaptor* ad, const vs_scratch* s, const vs_prefill_ws* ws,
                           const void* hidden, const void* emb_shifted, int L, int img_start, int img_end, void* g_out,
                           void* last_row_out, int32_t* state, void* stream);

/* A2-A5 decode branch (cnets_ours.py:976-1018): rows of x1[:, :H] hold the incoming hidden states. */
int vs_draft_forward(const vs_draft* dm, const vs_scratch* s, int T, const int32_t* ids, const int32_t* pos_ids,
                     const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx, int n_new, int causal,
                     const uint64_t* bits, void* stream);
/* LM head + level 0 + `depth` expansion levels + tree build (cnets_ours.py:1109-1238).  The stable
 * forward's output must be in scratch->h with VS_ST_LAST_ROW naming the row to continue from. */
int vs_draft_levels(const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb, int32_t* state, int head_rows,
                    const int32_t* sample_token_dev, int sample_token_host, void* stream);
/* topK_genrate "later calls" (cnets_ours.py:1090-1092 + the above): stable forward of VS_ST_M rows + levels. */
int vs_draft_tree(const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb, int32_t* state, void* stream);

/* One whole draft+verify cycle without a host sync (spec_model_ours.py:484-547 loop body):
 * verify -> accept -> KV splice -> gather accepted hidden -> draft tree.  out_rec as in vs_accept_greedy. */
int vs_spec_cycle(const vs_target* t, const vs_scratch* ts, const vs_draft* dm, const vs_scratch* ds,
                  const vs_tree_bufs* tb, int32_t* state, void* kv_slab, int64_t kv_max_pos, void* hidden_out,
                  void* logits, int32_t* amax, int32_t* out_rec, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Programs: the persistent decode kernel.  A program is a list of phases that ONE kernel launch (one CTA
 * per SM, co-resident, phases separated by grid barriers) executes back to back: the weight stream
 * (TMA -> tcgen05) runs ahead across phase boundaries and is prefetched into L2, so the per-op launches
 * above collapse into a single launch per draft+verify cycle.  Every vs_program_add_* mirrors the entry
 * point of the same name above (same arguments, same reference citations, same results up to the fp32
 * summation order of the k-split); token rows per phase are limited to 32, `splits`, `partials`,
 * `counters` of vs_linear_args are ignored (the program owns its split and workspace).
 *
 *   vs_program_create -> vs_program_add_* ... -> vs_program_workspace_bytes -> (caller allocates) ->
 *   vs_program_commit -> vs_program_run (any number of times) -> vs_program_destroy
 *
 * The handle is host memory owned by the library; all device memory is the caller's workspace.
 * ------------------------------------------------------------------------------------------------ */
typedef struct vs_program vs_program;
int vs_program_create(vs_program** out);
void vs_program_destroy(vs_program* pg);
const char* vs_program_last_error(void);
/* Attention inside a program: with option "attn_fused" = 1 (what every composite builder of the library's Python side
 * uses) ONE phase in which every (head, kv split) item is resident, K / V rows reach shared memory through tensor-map TMA
 * loads issued by a KV producer warp (csrc/attn_ring.cuh; vs_attn_args.prefix_stable lets it stage the cache prefix before
 * the phase starts) and the two passes meet at per-head exchanges (K and V read once).  Up to 4 granules of 32 keys per
 * split (1024 keys for 32 heads on 148 SMs) the scores never leave the SM; longer contexts park them in the caller's score
 * scratch (selected inside the kernel from the kv length).  The phase holds at most
 * vs_attention_fused_capacity(n_heads, scratch pv_splits, scratch max_chunks) keys (n_ctx + n_new) = what the scratch was
 * sized for (max_chunks * 64; score_ld must cover it and be a multiple of 32); the kernel traps beyond it.
 * "attn_fused" = 0 selects the two-phase A/B program (scores | softmax + PV on the cp.async kernels of vs_attention);
 * it is compiled in only with -DVS_TWO_PHASE_ATTN=1 (VS_ERR_UNSUPPORTED otherwise). */
int vs_attention_fused_capacity(int n_heads, int pv_splits_cap, int max_chunks);
/* options: "pf_ahead" / "inflight" (accepted and ignored since round 2: the L2 prefetcher warp is gone), "x3d" (default 1:
 * the activations of a slab stage arrive through one 3-D tensor-map load instead of one 2-D load per k-chunk),
 * formerly: "pf_ahead" (L2 prefetch distance in 16 KB weight chunks per SM, default 0 = off),
 * "trace" (1: record a %globaltimer stamp per phase), "grid" (CTAs; default = SM count; before adding phases),
 * "attn_fused" (see above; before adding phases), "fuse_epilogue" (default 1: an epilogue waits on per-tile arrival
 * counters instead of a grid barrier), "fuse_norm" (default 1; before adding phases: an RMSNorm that follows the
 * residual epilogue which produced its rows runs inside that epilogue -- same bits, one phase fewer),
 * "inflight" / "spin_ns" (tuning of the producer warps),
 * "top_k" (TopKLogitsWarper of the T > 0 acceptance phases added AFTERWARDS -- vs_program_add_accept_sampling,
 * vs_program_add_spec_cycle_sampling: every verify logit below the k-th largest of its row is dropped before the softmax,
 * utils.py:39-55; 0 = off),
 * "kv_two_loads" / "attn_prefetch" / "attn_pf" (A/B switches of the TMA-staged attention) */
int vs_program_set_option(vs_program* pg, const char* name, int value);
int vs_program_num_phases(const vs_program* pg);
int vs_program_phase_types(const vs_program* pg, int32_t* out, int cap);
size_t vs_program_workspace_bytes(vs_program* pg);
int vs_program_commit(vs_program* pg, void* workspace, size_t bytes, void* stream);   /* 256-byte aligned workspace */
int vs_program_run(const vs_program* pg, void* stream);
int vs_program_read_trace(const vs_program* pg, unsigned long long* out_ns, int cap);
/* out[phase * grid + cta] = time at which the CTA finished its work items of the phase; returns the grid size */
int vs_program_read_trace_cta(const vs_program* pg, unsigned long long* out_ns, int cap);

/* Weight layout for HBM streaming.  A row-major weight matrix W[N][K] can additionally be given in the TILED layout:
 * ceil(N/128) x (K/64) blocks of 16 KB, block (tile, kc) = rows [128*tile, +128) x columns [64*kc, +64) stored
 * contiguously in the shared-memory image the tensor cores read (row r of the block at byte 128*r, its eight 16-byte
 * chunks j stored at chunk position j ^ (r & 7); rows >= N are zero).  vispec_b200.engine.pack_weight_tiles builds
 * it.  Every linear added AFTER the registration whose W pointer matches (and ldw == K) streams the tiled copy with
 * contiguous 16 KB bulk copies and is prefetched into L2 ahead of use ("pf_ahead"); results are bit-identical. */
int vs_program_register_tiled(vs_program* pg, const void* w_row_major, const void* w_tiled);
/* Slab weights (round 2; vispec_b200/csrc/slab.h has the byte-level layout): the rows of W[N][K] are dealt over the `grid`
 * CTAs of the persistent kernel so that every CTA owns COMPLETE output rows, streams their whole K range as one
 * contiguous byte range and finishes them from TMEM -- no split-K partial tiles, no separate epilogue phase (the
 * reference computes each of these projections as one `F.linear`, e.g. modeling_llama_kv.py:563-565, 648, 421;
 * cnets_ours.py:363-365, 445, 508, 982-988).  `epi` selects the row pairing: VS_EPI_SWIGLU / VS_EPI_HEADS keep the two
 * members of a gate|up resp. rotate-half pair in the same accumulator lane.  vs_slab_bytes returns 0 for shapes the
 * layout does not cover.  A linear added after vs_program_register_slab whose W, epilogue class and grid match runs on
 * the slab path; anything else keeps the stream-K path above.  grid <= 0: the SM count of the current device. */
int vs_device_sms(void);
size_t vs_slab_bytes(int N, int K, int epi, int grid);
/* the plan behind a slab image, for tools and tests: out[0..13] = pair, packed, elements, tiles per pass, passes,
 * rows per block, block bytes (HBM), block stride in shared memory, k-chunks, k-chunks per stage, activation offset in a
 * slot, elements per unit, units with one more, grid.  Returns the number of values written (0: unsupported shape). */
int vs_slab_plan(int N, int K, int epi, int grid, int32_t* out, int cap);
int vs_slab_pack(const void* W, int64_t ldw, int N, int K, int epi, int grid, void* dst_128B_aligned, void* stream);
int vs_program_register_slab(vs_program* pg, const void* w_row_major, const void* w_slab, int epi, int grid);
int vs_program_add_linear(vs_program* pg, const vs_linear_args* a);
int vs_program_add_rmsnorm(vs_program* pg, const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int T,
                           int H, float eps);
int vs_program_add_gather_rows(vs_program* pg, const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd,
                               int T, int H);
int vs_program_add_embed_rows(vs_program* pg, const void* table, int64_t ldt, const int32_t* ids, void* dst, int64_t ldd,
                              int T, int H, int vocab);
int vs_program_add_broadcast_row(vs_program* pg, const void* row, void* dst, int64_t ldd, int T, int H);
int vs_program_add_argmax_rows(vs_program* pg, const void* logits, int64_t ldl, int T, int V, int32_t* out);
int vs_program_add_logsoftmax_topk(vs_program* pg, const void* logits, int64_t ldl, int T, int V, int k,
                                   const int32_t* row_sel, float* top_val, int32_t* top_idx);
int vs_program_add_attention(vs_program* pg, const vs_attn_args* a);
int vs_program_add_draft_level0(vs_program* pg, const vs_tree_bufs* tb, int32_t* state);
int vs_program_add_draft_expand(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, int level);
int vs_program_add_tree_build(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, const int32_t* sample_token_dev,
                              int sample_token_host);
int vs_program_add_accept_greedy(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, const int32_t* amax,
                                 int32_t* out_rec);
/* T > 0 acceptance (utils.py:453-493 with the TemperatureLogitsWarper of prepare_logits_processor :39-55) +
 * sampling of the next root token (utils.py:550-552): SpecInfer-style accept / reject along the tree with residual
 * renormalisation, on the verify logits [total+1][V] (bf16).  Uniforms come from a Philox4x32-10 stream keyed by
 * VS_ST_RNG_SEED / VS_ST_RNG_CTR (distributional parity with torch.rand_like / torch.multinomial), or -- for tests --
 * from `uni_override` (device floats: [leaf][VS_RETRIEVE_LD] acceptance uniforms followed by one sampling uniform at
 * index 32 * VS_RETRIEVE_LD).  Fills the same state / sel / stable_ids / out_rec as vs_accept_greedy. */
int vs_program_add_accept_sampling(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, const void* logits, int64_t ldl,
                                   int V, float temperature, const float* uni_override, int32_t* out_rec);
/* vs_program_add_spec_cycle with the T > 0 acceptance above instead of arg-max + greedy accept */
int vs_program_add_spec_cycle_sampling(vs_program* pg, const vs_target* t, const vs_scratch* ts, const vs_draft* dm,
                                       const vs_scratch* ds, const vs_tree_bufs* tb, int32_t* state, void* kv_slab,
                                       int64_t kv_max_pos, void* hidden_out, void* logits, float temperature,
                                       int32_t* out_rec);
int vs_program_add_kv_splice(vs_program* pg, void* slab, int n_slabs, int n_kv, int64_t max_pos, const int32_t* state,
                             const int32_t* sel_dev);
int vs_program_add_decoder_layer(vs_program* pg, const vs_dims* d, const vs_layer* l, const vs_scratch* s,
                                 const vs_rope* rope, int T, const int32_t* pos_ids, const int32_t* kv_row_base,
                                 int kv_row_off, const int32_t* n_ctx, int n_ctx_host, int n_new, int causal,
                                 const uint64_t* bits, int eager);
int vs_program_add_verify_step(vs_program* pg, const vs_target* t, const vs_scratch* s, const int32_t* token_ids,
                               const int32_t* pos_ids, const uint64_t* bits, int T, int32_t* state, void* hidden_out,
                               void* logits, int32_t* amax);
int vs_program_add_draft_forward(vs_program* pg, const vs_draft* dm, const vs_scratch* s, int T, const int32_t* ids,
                                 const int32_t* pos_ids, const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx,
                                 int n_new, int causal, const uint64_t* bits);
int vs_program_add_draft_levels(vs_program* pg, const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb,
                                int32_t* state, int head_rows, const int32_t* sample_token_dev, int sample_token_host);
int vs_program_add_draft_tree(vs_program* pg, const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb,
                              int32_t* state);
/* the whole draft+verify cycle of vs_spec_cycle as ONE launch */
int vs_program_add_spec_cycle(vs_program* pg, const vs_target* t, const vs_scratch* ts, const vs_draft* dm,
                              const vs_scratch* ds, const vs_tree_bufs* tb, int32_t* state, void* kv_slab,
                              int64_t kv_max_pos, void* hidden_out, void* logits, int32_t* amax, int32_t* out_rec);

/* ------------------------------------------------------------------------------------------------
 * Instrumentation (bench.py): number of kernels launched by this library, and optional CUDA-event timing of
 * every fused-linear (class 0, bytes = weight bytes streamed) and attention (class 1, bytes = K+V bytes per key)
 * launch on the launching stream.  vs_profile_stop synchronises the device and returns the record count.
 * ------------------------------------------------------------------------------------------------ */
long long vs_launch_count(int reset);
int vs_profile_start(int max_records);
int vs_profile_stop(float* ms_out, double* bytes_out, int32_t* cls_out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* VISPEC_B200_H */
