// extern "C" boundary of libvispec_b200.so (declared in include/vispec_b200.h) and the composite steps that
// chain the kernels of gemm_tc.cu / gemm_wide.cu / attention.cu / rowops.cu / tree.cu on one stream.
#include <math.h>
#include <stdio.h>

#include <string>

#include "../../include/vispec_b200.h"
#include "kernels.h"
#include "tree.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* where, const char* what) {
  g_err = std::string(where) + ": " + what;
  return code;
}

#define CK(expr)                                                              \
  do {                                                                        \
    cudaError_t _e = (expr);                                                  \
    if (_e != cudaSuccess) return fail(VS_ERR_CUDA, #expr, cudaGetErrorString(_e)); \
  } while (0)
#define RET(expr)                 \
  do {                            \
    int _r = (expr);              \
    if (_r != VS_OK) return _r;   \
  } while (0)
#define REQUIRE(cond, msg) \
  do {                     \
    if (!(cond)) return fail(VS_ERR_INVALID, __func__, msg); \
  } while (0)

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline const bf16* B(const void* p) { return reinterpret_cast<const bf16*>(p); }
inline bf16* B(void* p) { return reinterpret_cast<bf16*>(p); }

vs::GemmArgs to_gemm(const vs_linear_args* a) {
  vs::GemmArgs g;
  g.X = B(a->X); g.ldx = a->ldx; g.W = B(a->W); g.ldw = a->ldw; g.bias = B(a->bias);
  g.T = a->T; g.N = a->N; g.K = a->K; g.splits = a->splits < 1 ? 1 : a->splits;
  g.partials = a->partials; g.counters = a->counters; g.epi = a->epi;
  g.Y = B(a->Y); g.ldy = a->ldy; g.R = B(a->R); g.ldr = a->ldr;
  g.n_q_tiles = a->n_q_tiles; g.n_k_tiles = a->n_k_tiles; g.rope = a->rope;
  g.cos = B(a->cos); g.sin = B(a->sin); g.pos_ids = a->pos_ids;
  g.Qout = B(a->Qout); g.q_head_stride = a->q_head_stride;
  g.Kout = B(a->Kout); g.Vout = B(a->Vout); g.kv_head_stride = a->kv_head_stride;
  g.kv_row_base = a->kv_row_base; g.kv_row_off = a->kv_row_off;
  return g;
}

// split-K choice bounded by the scratch the caller gave us
int splits_for(const vs_scratch* s, int N, int K, int T) {
  int sp = vs::gemm_pick_splits(N, K, T);
  while (sp > 1 && (long)vs::gemm_partials_floats(N, T, sp) > s->partials_floats) --sp;
  return sp;
}

int linear(const vs_scratch* s, const bf16* X, long ldx, const bf16* W, const bf16* bias, int T, int N, int K,
           int epi, bf16* Y, long ldy, const bf16* R, long ldr, cudaStream_t st) {
  vs::GemmArgs g;
  g.X = X; g.ldx = ldx; g.W = W; g.ldw = K; g.bias = bias; g.T = T; g.N = N; g.K = K;
  g.splits = splits_for(s, N, K, T); g.partials = s->partials; g.counters = s->counters;
  g.epi = epi; g.Y = Y; g.ldy = ldy; g.R = R; g.ldr = ldr;
  CK(vs::launch_gemm(g, st));
  return VS_OK;
}

// packed position ids of the draft prefill (cnets_ours.py:932-948): text before the image run keeps its positions, the
// num_q - 1 compressed image tokens take the LAST positions of the run, text after the run keeps its positions
__global__ void prefill_positions_kernel(int32_t* pos, int n_pre, int n_c_img, int img_end, int n_post) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_c = n_pre + n_c_img + n_post;
  if (i >= n_c) return;
  pos[i] = i < n_pre ? i : (i < n_pre + n_c_img ? img_end - n_c_img + (i - n_pre) : img_end + (i - n_pre - n_c_img));
}

// rows [0, n) of src (ld lds) -> dst (ld ldd), `cols` bf16 elements each
cudaError_t copy_rows(bf16* dst, long ldd, const bf16* src, long lds, int n, int cols, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  return cudaMemcpy2DAsync(dst, ldd * 2, src, lds * 2, (size_t)cols * 2, n, cudaMemcpyDeviceToDevice, st);
}

}  // namespace

extern "C" {

const char* vs_last_error(void) { return g_err.c_str(); }
int vs_abi_version(void) { return VS_ABI_VERSION; }

int vs_linear_splits(int N, int K, int T) { return vs::gemm_pick_splits(N, K, T); }
size_t vs_linear_partials_floats(int N, int T, int splits) { return vs::gemm_partials_floats(N, T, splits); }

int vs_linear(const vs_linear_args* a, void* stream) {
  REQUIRE(a && a->X && a->W, "null operand");
  REQUIRE(a->K % 64 == 0, "K must be a multiple of 64");
  REQUIRE(a->epi != VS_EPI_SWIGLU || a->N % 128 == 0, "SWIGLU needs N % 128 == 0");
  REQUIRE(a->epi != VS_EPI_HEADS || a->N % 128 == 0, "HEADS needs N % 128 == 0");
  CK(vs::launch_gemm(to_gemm(a), S(stream)));
  return VS_OK;
}

int vs_rmsnorm(const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int T, int H, float eps,
               void* stream) {
  CK(vs::launch_rmsnorm(B(x), ldx, B(weight), B(y), ldy, T, H, eps, S(stream)));
  return VS_OK;
}
int vs_gather_rows(const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd, int T, int H,
                   void* stream) {
  CK(vs::launch_gather_rows(B(src), lds, idx, B(dst), ldd, T, H, S(stream)));
  return VS_OK;
}
int vs_embed_rows(const void* table, int64_t ldt, const int32_t* ids, void* dst, int64_t ldd, int T, int H, int vocab,
                  void* stream) {
  CK(vs::launch_embed_rows(B(table), ldt, ids, B(dst), ldd, T, H, vocab, S(stream)));
  return VS_OK;
}
int vs_broadcast_row(const void* row, void* dst, int64_t ldd, int T, int H, void* stream) {
  CK(vs::launch_broadcast_row(B(row), B(dst), ldd, T, H, S(stream)));
  return VS_OK;
}
int vs_argmax_rows(const void* logits, int64_t ldl, int T, int V, int32_t* out, void* stream) {
  CK(vs::launch_argmax_rows(B(logits), ldl, T, V, out, S(stream)));
  return VS_OK;
}
int vs_logsoftmax_topk(const void* logits, int64_t ldl, int T, int V, int k, const int32_t* row_sel, float* top_val,
                       int32_t* top_idx, void* stream) {
  REQUIRE(!row_sel || T == 1, "row_sel selects exactly one row");
  CK(vs::launch_logsoftmax_topk(B(logits), ldl, T, V, k, row_sel, top_val, top_idx, S(stream)));
  return VS_OK;
}

int vs_attention(const vs_attn_args* a, void* stream) {
  REQUIRE(a && a->Q && a->K && a->V && a->O, "null operand");
  REQUIRE(a->n_kv_heads > 0 && a->n_heads % a->n_kv_heads == 0, "heads must be a multiple of kv heads");
  vs::AttnArgs g;
  g.Q = B(a->Q); g.q_head_stride = a->q_head_stride; g.K = B(a->K); g.V = B(a->V); g.kv_head_stride = a->kv_head_stride;
  g.n_heads = a->n_heads; g.n_kv_heads = a->n_kv_heads; g.Tq = a->Tq; g.n_ctx = a->n_ctx; g.n_ctx_host = a->n_ctx_host;
  g.n_new = a->n_new; g.causal = a->causal; g.bits = reinterpret_cast<const unsigned long long*>(a->bits);
  g.eager = a->eager; g.scale = a->scale; g.scores = a->scores; g.score_ld = a->score_ld; g.stats = a->stats;
  g.max_chunks = a->max_chunks; g.opart = a->opart; g.pv_splits = a->pv_splits; g.counters = a->counters;
  g.O = B(a->O); g.ldo = a->ldo;
  CK(vs::launch_attention(g, S(stream)));
  return VS_OK;
}

int vs_draft_level0(const vs_tree_bufs* tb, int32_t* state, void* stream) {
  CK(vs::launch_draft_level0(*tb, state, S(stream)));
  return VS_OK;
}
int vs_draft_expand(const vs_tree_bufs* tb, int32_t* state, int level, void* stream) {
  CK(vs::launch_draft_expand(*tb, state, level, S(stream)));
  return VS_OK;
}
int vs_tree_build(const vs_tree_bufs* tb, int32_t* state, const int32_t* sample_token_dev, int sample_token_host,
                  void* stream) {
  REQUIRE(tb->depth + 2 <= VS_RETRIEVE_LD, "tree depth + 2 must be <= VS_RETRIEVE_LD");
  REQUIRE(tb->top_k >= 1 && tb->top_k <= 8, "top_k must be in 1..8");
  CK(vs::launch_tree_build(*tb, state, sample_token_dev, sample_token_host, S(stream)));
  return VS_OK;
}
int vs_accept_greedy(const vs_tree_bufs* tb, int32_t* state, const int32_t* amax, int32_t* out_rec, void* stream) {
  CK(vs::launch_accept_greedy(*tb, state, amax, out_rec, S(stream)));
  return VS_OK;
}
int vs_kv_splice(void* slab, int n_slabs, int n_kv, int64_t max_pos, const int32_t* state, const int32_t* sel_dev,
                 int prev_len_host, int n_host, const int32_t* sel_host, void* stream) {
  REQUIRE(slab, "null slab");
  REQUIRE(state || n_host <= VS_MAX_ACCEPT, "at most VS_MAX_ACCEPT rows per splice");
  CK(vs::launch_kv_splice(B(slab), n_slabs, n_kv, max_pos, state, sel_dev, prev_len_host, n_host, sel_host, S(stream)));
  return VS_OK;
}
int vs_set_state(int32_t* state, int slot, int value, void* stream) {
  REQUIRE(slot >= 0 && slot < VS_ST_SLOTS, "bad slot");
  CK(vs::launch_set_state(state, slot, value, S(stream)));
  return VS_OK;
}

int vs_decoder_layer(const vs_dims* d, const vs_layer* l, const vs_scratch* s, const vs_rope* rope, int T,
                     const int32_t* pos_ids, const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx,
                     int n_ctx_host, int n_new, int causal, const uint64_t* bits, int eager, void* stream) {
  REQUIRE(T >= 1 && T <= s->max_rows && s->max_rows % VS_TOK_BLOCK == 0, "rows exceed scratch->max_rows");
  REQUIRE(d->hidden % 64 == 0 && d->inter % 64 == 0, "hidden and inter must be multiples of 64");
  cudaStream_t st = S(stream);
  const int H = d->hidden, I = d->inter, nh = d->n_heads, nkv = d->n_kv_heads;
  bf16* h = B(s->h);
  const bf16* xin = h;
  if (l->ln_in) {
    CK(vs::launch_rmsnorm(h, H, B(l->ln_in), B(s->x), H, T, H, d->eps, st));
    xin = B(s->x);
  }
  {
    vs::GemmArgs g;
    g.X = xin; g.ldx = H; g.W = B(l->w_qkv); g.ldw = H; g.bias = B(l->b_qkv);
    g.T = T; g.N = (nh + 2 * nkv) * 128; g.K = H;
    g.splits = splits_for(s, g.N, g.K, T); g.partials = s->partials; g.counters = s->counters;
    g.epi = vs::EPI_HEADS; g.n_q_tiles = nh; g.n_k_tiles = nkv; g.rope = 1;
    g.cos = B(rope->cos); g.sin = B(rope->sin); g.pos_ids = pos_ids;
    g.Qout = B(s->q); g.q_head_stride = (long)s->max_rows * 128;
    g.Kout = B(l->k_rows); g.Vout = B(l->v_rows); g.kv_head_stride = l->kv_head_stride;
    g.kv_row_base = kv_row_base; g.kv_row_off = kv_row_off;
    CK(vs::launch_gemm(g, st));
  }
  {
    vs::AttnArgs a;
    a.Q = B(s->q); a.q_head_stride = (long)s->max_rows * 128;
    a.K = B(l->k_rows); a.V = B(l->v_rows); a.kv_head_stride = l->kv_head_stride;
    a.n_heads = nh; a.n_kv_heads = nkv; a.Tq = T; a.n_ctx = n_ctx; a.n_ctx_host = n_ctx_host; a.n_new = n_new;
    a.causal = causal; a.bits = reinterpret_cast<const unsigned long long*>(bits); a.eager = eager;
    a.scale = 1.0f / sqrtf(128.0f);
    a.scores = s->scores; a.score_ld = s->score_ld; a.stats = s->stats; a.max_chunks = s->max_chunks;
    a.opart = s->opart; a.pv_splits = s->pv_splits; a.counters = s->attn_counters;
    a.O = B(s->attn); a.ldo = (long)nh * 128;
    CK(vs::launch_attention(a, st));
  }
  RET(linear(s, B(s->attn), (long)nh * 128, B(l->w_o), nullptr, T, H, nh * 128, vs::EPI_RESID, h, H, h, H, st));
  CK(vs::launch_rmsnorm(h, H, B(l->ln_post), B(s->x), H, T, H, d->eps, st));
  RET(linear(s, B(s->x), H, B(l->w_gateup), nullptr, T, 2 * I, H, vs::EPI_SWIGLU, B(s->act), I, nullptr, 0, st));
  RET(linear(s, B(s->act), I, B(l->w_down), nullptr, T, H, I, vs::EPI_RESID, h, H, h, H, st));
  return VS_OK;
}

int vs_verify_step(const vs_target* t, const vs_scratch* s, const int32_t* token_ids, const int32_t* pos_ids,
                   const uint64_t* bits, int T, int32_t* state, void* hidden_out, void* logits, int32_t* amax,
                   void* stream) {
  REQUIRE(T >= 1 && T <= VS_TOK_BLOCK, "verify handles 1..32 tree tokens");
  cudaStream_t st = S(stream);
  const vs_dims* d = &t->dims;
  const int H = d->hidden, V = d->vocab;
  CK(vs::launch_embed_rows(B(t->embed), H, token_ids, B(s->h), H, T, H, V, st));
  const int32_t* cur = state + VS_ST_CUR_LEN;
  for (int li = 0; li < d->n_layers; ++li)
    RET(vs_decoder_layer(d, &t->layers[li], s, &t->rope, T, pos_ids, cur, 0, cur, 0, T, bits ? 0 : 1, bits, t->sdpa_attention ? 0 : 1, stream));
  CK(vs::launch_rmsnorm(B(s->h), H, B(t->norm), B(hidden_out), H, T, H, d->eps, st));
  RET(linear(s, B(hidden_out), H, B(t->lm_head), nullptr, T, V, H, vs::EPI_STORE, B(logits), V, nullptr, 0, st));
  if (amax) CK(vs::launch_argmax_rows(B(logits), V, T, V, amax, st));
  return VS_OK;
}

int vs_draft_forward(const vs_draft* dm, const vs_scratch* s, int T, const int32_t* ids, const int32_t* pos_ids,
                     const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx, int n_new, int causal,
                     const uint64_t* bits, void* stream) {
  REQUIRE(T >= 1 && T <= s->max_rows, "rows exceed scratch->max_rows");
  cudaStream_t st = S(stream);
  const vs_dims* d = &dm->dims;
  const int H = d->hidden;
  // h' = fc(cat(embed(ids), img_fc(cat(hidden, g))))   (cnets_ours.py:982-988)
  CK(vs::launch_embed_rows(B(dm->embed), H, ids, B(dm->x2), 2L * H, T, H, d->vocab, st));
  RET(linear(s, B(dm->x1), 2L * H, B(dm->w_img_fc), B(dm->b_img_fc), T, H, 2 * H, vs::EPI_STORE, B(dm->x2) + H, 2L * H,
             nullptr, 0, st));
  RET(linear(s, B(dm->x2), 2L * H, B(dm->w_fc), B(dm->b_fc), T, H, 2 * H, vs::EPI_STORE, B(s->h), H, nullptr, 0, st));
  return vs_decoder_layer(d, &dm->layer, s, &dm->rope, T, pos_ids, kv_row_base, kv_row_off, n_ctx, 0, n_new, causal, bits,
                          0, stream);
}

int vs_img_adaptor_prefill(const vs_draft* dm, const vs_adaptor* ad, const vs_scratch* s, const vs_prefill_ws* ws,
                           const void* hidden, const void* emb_shifted, int L, int img_start, int img_end, void* g_out,
                           void* last_row_out, int32_t* state, void* stream) {
  REQUIRE(dm && s && ws && hidden && emb_shifted && g_out && state, "null operand");
  REQUIRE(L >= 1, "empty prompt");
  const bool has_img = img_start >= 0;
  REQUIRE(!has_img || (ad && img_start < img_end && img_end <= L), "image run must satisfy 0 <= start < end <= L");
  cudaStream_t st = S(stream);
  const vs_dims* d = &dm->dims;
  const int H = d->hidden, nh = d->n_heads;
  const int n_img = has_img ? img_end - img_start : 0;
  const int n_pre = has_img ? img_start : L, n_post = has_img ? L - img_end : 0;
  const int n_c_img = has_img ? ad->num_q - 1 : 0;
  const int n_text = n_pre + n_post, n_c = n_text + n_c_img;
  REQUIRE(!has_img || (ad->num_q >= 1 && ad->num_q <= VS_TOK_BLOCK), "num_q must be in 1..32");
  REQUIRE(n_c >= 1 && n_c <= ws->max_rows && n_img <= ws->max_rows && ws->max_rows <= s->max_rows,
          "prompt exceeds the prefill workspace / scratch rows");
  const bf16* hid = B(hidden);
  const bf16* emb = B(emb_shifted);
  bf16* x1 = B(ws->x1);
  bf16* x2 = B(ws->x2);
  bf16* fused = B(ws->fused);
  bf16* adapted = B(ws->adapted);
  bf16* h = B(s->h);
  if (has_img) {
    // ---- ImgAdaptor.forward (cnets_ours.py:630-661): K,V = Linear(image embeds) head-major, the num_q learned queries
    // cross-attend over the image rows (no mask), o_proj.  The image rows are one contiguous run of the prompt.
    const long kvs = (long)ws->max_rows * 128;
    vs::GemmArgs g;
    g.X = emb + (long)img_start * H; g.ldx = H; g.W = B(ad->w_kv); g.ldw = H; g.bias = B(ad->b_kv);
    g.T = n_img; g.N = 2 * H; g.K = H;
    g.splits = splits_for(s, g.N, g.K, n_img); g.partials = s->partials; g.counters = s->counters;
    g.epi = vs::EPI_HEADS; g.n_q_tiles = 0; g.n_k_tiles = nh; g.rope = 0;
    g.Kout = B(ws->img_k); g.Vout = B(ws->img_v); g.kv_head_stride = kvs;
    CK(vs::launch_gemm(g, st));
    vs::AttnArgs a;
    a.Q = B(ad->q); a.q_head_stride = 32L * 128; a.K = B(ws->img_k); a.V = B(ws->img_v); a.kv_head_stride = kvs;
    a.n_heads = nh; a.n_kv_heads = nh; a.Tq = ad->num_q; a.n_ctx = nullptr; a.n_ctx_host = n_img; a.n_new = 0;
    a.causal = 0; a.bits = nullptr; a.eager = 0; a.scale = 1.0f / sqrtf(128.0f);
    a.scores = s->scores; a.score_ld = s->score_ld; a.stats = s->stats; a.max_chunks = s->max_chunks;
    a.opart = s->opart; a.pv_splits = s->pv_splits; a.counters = s->attn_counters;
    a.O = B(ws->attn_out); a.ldo = H;
    REQUIRE(n_img <= s->score_ld, "image run longer than the scratch's score rows");
    CK(vs::launch_attention(a, st));
    RET(linear(s, B(ws->attn_out), H, B(ad->w_o), nullptr, ad->num_q, H, H, vs::EPI_STORE, adapted, H, nullptr, 0, st));
    CK(cudaMemcpyAsync(g_out, adapted + (long)(ad->num_q - 1) * H, (size_t)H * 2, cudaMemcpyDeviceToDevice, st));   // :930
  } else {
    CK(cudaMemsetAsync(g_out, 0, (size_t)H * 2, st));                 // :977-981: the dummy adaptor call adds 0
  }
  // ---- text rows: fc(cat(emb, img_fc(cat(hid, g)))) with g = 0 before the image run and the new g behind it (:918-948)
  if (n_text > 0) {
    CK(copy_rows(x1, 2L * H, hid, H, n_pre, H, st));
    CK(copy_rows(x1 + (long)n_pre * 2 * H, 2L * H, hid + (long)img_end * H, H, n_post, H, st));
    CK(cudaMemset2DAsync(x1 + H, 2L * H * 2, 0, (size_t)H * 2, n_pre, st));
    if (n_post > 0) CK(vs::launch_broadcast_row(B(g_out), x1 + (long)n_pre * 2 * H + H, 2L * H, n_post, H, st));
    CK(copy_rows(x2, 2L * H, emb, H, n_pre, H, st));
    CK(copy_rows(x2 + (long)n_pre * 2 * H, 2L * H, emb + (long)img_end * H, H, n_post, H, st));
    RET(linear(s, x1, 2L * H, B(dm->w_img_fc), B(dm->b_img_fc), n_text, H, 2 * H, vs::EPI_STORE, x2 + H, 2L * H, nullptr, 0, st));
    RET(linear(s, x2, 2L * H, B(dm->w_fc), B(dm->b_fc), n_text, H, 2 * H, vs::EPI_STORE, fused, H, nullptr, 0, st));
  }
  // ---- packed order: pre-text, compressed image tokens, post-text; their position ids
  CK(copy_rows(h, H, fused, H, n_pre, H, st));
  if (n_c_img > 0) CK(copy_rows(h + (long)n_pre * H, H, adapted, H, n_c_img, H, st));
  CK(copy_rows(h + (long)(n_pre + n_c_img) * H, H, fused + (long)n_pre * H, H, n_post, H, st));
  prefill_positions_kernel<<<(n_c + 255) / 256, 256, 0, st>>>(ws->pos, n_pre, n_c_img, has_img ? img_end : 0, n_post);
  CK(cudaGetLastError());
  if (ws->packed) CK(copy_rows(B(ws->packed), H, h, H, n_c, H, st));
  // ---- the draft's decoder layer over the packed rows (causal, cnets_ours.py:971-975), KV rows [0, n_c) of the cache
  RET(vs_decoder_layer(d, &dm->layer, s, &dm->rope, n_c, ws->pos, nullptr, 0, nullptr, 0, n_c, 1, nullptr, 0, stream));
  // ---- state for the decode calls: g in the right half of x1, the last row to continue from, lengths
  CK(vs::launch_broadcast_row(B(g_out), B(dm->x1) + H, 2L * H, VS_TOK_BLOCK, H, st));
  if (last_row_out) CK(cudaMemcpyAsync(last_row_out, h + (long)(n_c - 1) * H, (size_t)H * 2, cudaMemcpyDeviceToDevice, st));
  CK(vs::launch_set_state(state, VS_ST_D_LEN, n_c, st));
  CK(vs::launch_set_state(state, VS_ST_D_POS, L, st));
  CK(vs::launch_set_state(state, VS_ST_LAST_ROW, 0, st));
  return n_c;
}

int vs_draft_levels(const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb, int32_t* state, int head_rows,
                    const int32_t* sample_token_dev, int sample_token_host, void* stream) {
  cudaStream_t st = S(stream);
  const vs_dims* d = &dm->dims;
  const int H = d->hidden, V = d->vocab, k = tb->top_k;
  REQUIRE(k >= 1 && k <= 8 && tb->depth >= 1 && tb->depth + 2 <= VS_RETRIEVE_LD, "unsupported tree shape");
  REQUIRE(k * (tb->depth + 1) <= 64, "draft tree mask is one 64-bit word per query");
  REQUIRE(head_rows >= 1 && head_rows <= VS_TOK_BLOCK, "bad head_rows");
  RET(linear(s, B(s->h), H, B(dm->lm_head), nullptr, head_rows, V, H, vs::EPI_STORE, B(dm->logits), V, nullptr, 0, st));
  CK(vs::launch_logsoftmax_topk(B(dm->logits), V, 1, V, k, state + VS_ST_LAST_ROW, tb->top_val, tb->top_idx, st));
  CK(vs::launch_draft_level0(*tb, state, st));
  const int32_t* dlen = state + VS_ST_D_LEN;
  for (int i = 0; i < tb->depth; ++i) {
    CK(vs::launch_gather_rows(B(s->h), H, tb->gather_idx, B(dm->x1), 2L * H, k, H, st));
    RET(vs_draft_forward(dm, s, k, tb->frontier_ids, tb->draft_pos_ids, dlen, k * i, dlen, k * (i + 1), 0,
                         tb->frontier_bits, stream));
    RET(linear(s, B(s->h), H, B(dm->lm_head), nullptr, k, V, H, vs::EPI_STORE, B(dm->logits), V, nullptr, 0, st));
    CK(vs::launch_logsoftmax_topk(B(dm->logits), V, k, V, k, nullptr, tb->top_val, tb->top_idx, st));
    CK(vs::launch_draft_expand(*tb, state, i, st));
  }
  return vs_tree_build(tb, state, sample_token_dev, sample_token_host, stream);
}

int vs_draft_tree(const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb, int32_t* state, void* stream) {
  cudaStream_t st = S(stream);
  const int32_t* dlen = state + VS_ST_D_LEN;
  // stable forward on the (padded) accepted rows; rows >= VS_ST_M are scratch and never read back
  RET(vs_draft_forward(dm, s, VS_MAX_ACCEPT, tb->stable_ids, tb->stable_pos, dlen, 0, dlen, VS_MAX_ACCEPT, 1, nullptr,
                       stream));
  CK(vs::launch_draft_commit(state, st));
  return vs_draft_levels(dm, s, tb, state, VS_MAX_ACCEPT, state + VS_ST_NEXT_TOKEN, 0, stream);
}

int vs_spec_cycle(const vs_target* t, const vs_scratch* ts, const vs_draft* dm, const vs_scratch* ds,
                  const vs_tree_bufs* tb, int32_t* state, void* kv_slab, int64_t kv_max_pos, void* hidden_out,
                  void* logits, int32_t* amax, int32_t* out_rec, void* stream) {
  cudaStream_t st = S(stream);
  const int T = tb->total + 1;
  REQUIRE(T <= VS_TOK_BLOCK, "tree larger than 32 nodes");
  RET(vs_verify_step(t, ts, tb->draft_tokens, tb->verify_pos, tb->tree_bits, T, state, hidden_out, logits, amax, stream));
  CK(vs::launch_accept_greedy(*tb, state, amax, out_rec, st));
  CK(vs::launch_kv_splice(B(kv_slab), 2 * t->dims.n_layers, t->dims.n_kv_heads, kv_max_pos, state, tb->sel, 0, 0, nullptr,
                          st));
  // accepted hidden states -> draft input rows (utils.py:543-546)
  CK(vs::launch_gather_rows(B(hidden_out), t->dims.hidden, tb->sel, B(dm->x1), 2L * dm->dims.hidden, VS_MAX_ACCEPT,
                            t->dims.hidden, st));
  return vs_draft_tree(dm, ds, tb, state, stream);
}

}  // extern "C"
