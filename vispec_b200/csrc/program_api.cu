// extern "C" builders of the persistent-program path (include/vispec_b200.h, "Programs"): per-op phases that mirror the
// per-op entry points of api.cu one to one, and the composite steps (decoder layer, verify step, draft forward / levels
// / tree, whole draft+verify cycle) assembled from them in the same order as api.cu chains its launches.
#include <math.h>

#include "../../include/vispec_b200.h"
#include "program.h"

namespace {

#define RET(expr)                 \
  do {                            \
    int _r = (expr);              \
    if (_r != VS_OK) return _r;   \
  } while (0)
#define REQUIRE(cond, msg) \
  do {                     \
    if (!(cond)) return vs::prog_fail(VS_ERR_INVALID, __func__, msg); \
  } while (0)


inline const bf16* B(const void* p) { return reinterpret_cast<const bf16*>(p); }
inline bf16* B(void* p) { return reinterpret_cast<bf16*>(p); }

vs::GemmArgs to_gemm(const vs_linear_args* a) {
  vs::GemmArgs g;
  g.X = B(a->X); g.ldx = a->ldx; g.W = B(a->W); g.ldw = a->ldw; g.bias = B(a->bias);
  g.T = a->T; g.N = a->N; g.K = a->K; g.splits = 1;
  g.epi = a->epi;
  g.Y = B(a->Y); g.ldy = a->ldy; g.R = B(a->R); g.ldr = a->ldr;
  g.n_q_tiles = a->n_q_tiles; g.n_k_tiles = a->n_k_tiles; g.rope = a->rope;
  g.cos = B(a->cos); g.sin = B(a->sin); g.pos_ids = a->pos_ids;
  g.Qout = B(a->Qout); g.q_head_stride = a->q_head_stride;
  g.Kout = B(a->Kout); g.Vout = B(a->Vout); g.kv_head_stride = a->kv_head_stride;
  g.kv_row_base = a->kv_row_base; g.kv_row_off = a->kv_row_off;
  return g;
}

int linear(vs_program* pg, const bf16* X, long ldx, const bf16* W, const bf16* bias, int T, int N, int K, int epi, bf16* Y,
           long ldy, const bf16* R, long ldr, float* ssq) {
  vs::GemmArgs g;
  g.X = X; g.ldx = ldx; g.W = W; g.ldw = K; g.bias = bias; g.T = T; g.N = N; g.K = K;
  g.epi = epi; g.Y = Y; g.ldy = ldy; g.R = R; g.ldr = ldr;
  return vs::prog_add_linear(pg, g, ssq);
}

// One decoder layer on the T rows of scratch->h.  `h_has_ssq`: the workspace holds the sums of squares of h (left by the
// residual epilogue that produced it); updated for the caller.
int decoder_layer(vs_program* pg, const vs_dims* d, const vs_layer* l, const vs_scratch* s, const vs_rope* rope, int T,
                  const int32_t* pos_ids, const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx, int n_ctx_host,
                  int n_new, int causal, const uint64_t* bits, int eager, bool* h_has_ssq, int prefix_stable = 0) {
  REQUIRE(T >= 1 && T <= VS_TOK_BLOCK && s->max_rows >= VS_TOK_BLOCK, "programs handle 1..32 rows");
  REQUIRE(d->hidden % 64 == 0 && d->inter % 64 == 0, "hidden and inter must be multiples of 64");
  const int H = d->hidden, I = d->inter, nh = d->n_heads, nkv = d->n_kv_heads;
  bf16* h = B(s->h);
  const bf16* xin = h;
  if (l->ln_in) {
    RET(vs::prog_add_norm(pg, h, H, B(l->ln_in), B(s->x), H, T, H, d->eps, *h_has_ssq ? PROG_WS_SSQ : nullptr, H / 128));
    xin = B(s->x);
  }
  {
    vs::GemmArgs g;
    g.X = xin; g.ldx = H; g.W = B(l->w_qkv); g.ldw = H; g.bias = B(l->b_qkv);
    g.T = T; g.N = (nh + 2 * nkv) * 128; g.K = H;
    g.epi = vs::EPI_HEADS; g.n_q_tiles = nh; g.n_k_tiles = nkv; g.rope = 1;
    g.cos = B(rope->cos); g.sin = B(rope->sin); g.pos_ids = pos_ids;
    g.Qout = B(s->q); g.q_head_stride = (long)s->max_rows * 128;
    g.Kout = B(l->k_rows); g.Vout = B(l->v_rows); g.kv_head_stride = l->kv_head_stride;
    g.kv_row_base = kv_row_base; g.kv_row_off = kv_row_off;
    RET(vs::prog_add_linear(pg, g, nullptr));
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
    a.prefix_stable = prefix_stable;
    RET(vs::prog_add_attention(pg, a));
  }
  REQUIRE(H % 128 == 0, "hidden must be a multiple of 128");
  RET(linear(pg, B(s->attn), (long)nh * 128, B(l->w_o), nullptr, T, H, nh * 128, vs::EPI_RESID, h, H, h, H, PROG_WS_SSQ));
  RET(vs::prog_add_norm(pg, h, H, B(l->ln_post), B(s->x), H, T, H, d->eps, PROG_WS_SSQ, H / 128));
  RET(linear(pg, B(s->x), H, B(l->w_gateup), nullptr, T, 2 * I, H, vs::EPI_SWIGLU, B(s->act), I, nullptr, 0, nullptr));
  RET(linear(pg, B(s->act), I, B(l->w_down), nullptr, T, H, I, vs::EPI_RESID, h, H, h, H, PROG_WS_SSQ));
  *h_has_ssq = true;
  return VS_OK;
}

int verify_step(vs_program* pg, const vs_target* t, const vs_scratch* s, const int32_t* token_ids, const int32_t* pos_ids,
                const uint64_t* bits, int T, int32_t* state, void* hidden_out, void* logits, int32_t* amax,
                vs::MergeInfo* fuse_argmax) {
  REQUIRE(T >= 1 && T <= VS_TOK_BLOCK, "verify handles 1..32 tree tokens");
  const vs_dims* d = &t->dims;
  const int H = d->hidden, V = d->vocab;
  RET(vs::prog_add_rows(pg, B(t->embed), H, token_ids, V, 0, B(s->h), H, T, H));
  const int32_t* cur = state + VS_ST_CUR_LEN;
  bool has_ssq = false;
  for (int li = 0; li < d->n_layers; ++li)
    // the kv length and the prefix rows of the verify step are final when the launch starts (prog_add_attention drops the
    // flag when a tree / splice phase precedes it in the program)
    RET(decoder_layer(pg, d, &t->layers[li], s, &t->rope, T, pos_ids, cur, 0, cur, 0, T, bits ? 0 : 1, bits, t->sdpa_attention ? 0 : 1, &has_ssq, 1));
  RET(vs::prog_add_norm(pg, B(s->h), H, B(t->norm), B(hidden_out), H, T, H, d->eps, has_ssq ? PROG_WS_SSQ : nullptr, H / 128));
  RET(linear(pg, B(hidden_out), H, B(t->lm_head), nullptr, T, V, H, vs::EPI_STORE, B(logits), V, nullptr, 0, nullptr));
  if (amax) RET(vs::prog_add_argmax(pg, B(logits), V, T, V, amax, fuse_argmax));
  return VS_OK;
}

int draft_forward(vs_program* pg, const vs_draft* dm, const vs_scratch* s, int T, const int32_t* ids, const int32_t* pos_ids,
                  const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx, int n_new, int causal,
                  const uint64_t* bits) {
  REQUIRE(T >= 1 && T <= VS_TOK_BLOCK, "programs handle 1..32 rows");
  const vs_dims* d = &dm->dims;
  const int H = d->hidden;
  // h' = fc(cat(embed(ids), img_fc(cat(hidden, g))))   (cnets_ours.py:982-988)
  RET(vs::prog_add_rows(pg, B(dm->embed), H, ids, d->vocab, 0, B(dm->x2), 2L * H, T, H));
  RET(linear(pg, B(dm->x1), 2L * H, B(dm->w_img_fc), B(dm->b_img_fc), T, H, 2 * H, vs::EPI_STORE, B(dm->x2) + H, 2L * H,
             nullptr, 0, nullptr));
  RET(linear(pg, B(dm->x2), 2L * H, B(dm->w_fc), B(dm->b_fc), T, H, 2 * H, vs::EPI_STORE, B(s->h), H, nullptr, 0, nullptr));
  bool has_ssq = false;
  return decoder_layer(pg, d, &dm->layer, s, &dm->rope, T, pos_ids, kv_row_base, kv_row_off, n_ctx, 0, n_new, causal, bits, 0,
                       &has_ssq);
}

int draft_levels(vs_program* pg, const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb, int32_t* state,
                 int head_rows, const int32_t* sample_token_dev, int sample_token_host) {
  const vs_dims* d = &dm->dims;
  const int H = d->hidden, V = d->vocab, k = tb->top_k;
  REQUIRE(k >= 1 && k <= 8 && tb->depth >= 1 && tb->depth + 2 <= VS_RETRIEVE_LD, "unsupported tree shape");
  REQUIRE(k * (tb->depth + 1) <= 64, "draft tree mask is one 64-bit word per query");
  REQUIRE(head_rows >= 1 && head_rows <= VS_TOK_BLOCK, "bad head_rows");
  RET(linear(pg, B(s->h), H, B(dm->lm_head), nullptr, head_rows, V, H, vs::EPI_STORE, B(dm->logits), V, nullptr, 0, nullptr));
  vs::MergeInfo mi;
  RET(vs::prog_add_lsm_topk(pg, B(dm->logits), V, 1, V, k, state + VS_ST_LAST_ROW, tb->top_val, tb->top_idx, &mi));
  RET(vs::prog_add_tree(pg, vs::TREE_LEVEL0, *tb, state, 0, nullptr, nullptr, nullptr, 0, &mi));
  const int32_t* dlen = state + VS_ST_D_LEN;
  for (int i = 0; i < tb->depth; ++i) {
    RET(vs::prog_add_rows(pg, B(s->h), H, tb->gather_idx, VS_TOK_BLOCK, 0, B(dm->x1), 2L * H, k, H));
    RET(draft_forward(pg, dm, s, k, tb->frontier_ids, tb->draft_pos_ids, dlen, k * i, dlen, k * (i + 1), 0,
                      tb->frontier_bits));
    RET(linear(pg, B(s->h), H, B(dm->lm_head), nullptr, k, V, H, vs::EPI_STORE, B(dm->logits), V, nullptr, 0, nullptr));
    RET(vs::prog_add_lsm_topk(pg, B(dm->logits), V, k, V, k, nullptr, tb->top_val, tb->top_idx, &mi));
    RET(vs::prog_add_tree(pg, vs::TREE_EXPAND, *tb, state, i, nullptr, nullptr, nullptr, 0, &mi));
  }
  return vs::prog_add_tree(pg, vs::TREE_BUILD, *tb, state, 0, nullptr, nullptr, sample_token_dev, sample_token_host, nullptr);
}

int draft_tree(vs_program* pg, const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb, int32_t* state) {
  const int32_t* dlen = state + VS_ST_D_LEN;
  // stable forward on the (padded) accepted rows; rows >= VS_ST_M are scratch and never read back
  RET(draft_forward(pg, dm, s, VS_MAX_ACCEPT, tb->stable_ids, tb->stable_pos, dlen, 0, dlen, VS_MAX_ACCEPT, 1, nullptr));
  RET(vs::prog_add_tree(pg, vs::TREE_COMMIT, *tb, state, 0, nullptr, nullptr, nullptr, 0, nullptr));
  return draft_levels(pg, dm, s, tb, state, VS_MAX_ACCEPT, state + VS_ST_NEXT_TOKEN, 0);
}

}  // namespace

extern "C" {

int vs_program_add_linear(vs_program* pg, const vs_linear_args* a) {
  REQUIRE(pg && a && a->X && a->W, "null operand");
  // a residual epilogue leaves the sums of squares of its new rows in the program's workspace: an RMSNorm of exactly
  // those rows added right behind it (the usual next step) starts from them / is folded into the epilogue
  return vs::prog_add_linear(pg, to_gemm(a), a->epi == vs::EPI_RESID && a->N % 128 == 0 ? PROG_WS_SSQ : nullptr);
}
int vs_program_add_rmsnorm(vs_program* pg, const void* x, int64_t ldx, const void* weight, void* y, int64_t ldy, int T,
                           int H, float eps) {
  REQUIRE(pg && x && weight && y, "null operand");
  const bool after_resid = vs::prog_last_is_residual_of(pg, B(x), ldx, T, H);
  return vs::prog_add_norm(pg, B(x), ldx, B(weight), B(y), ldy, T, H, eps, after_resid ? PROG_WS_SSQ : nullptr, H / 128);
}
int vs_program_add_gather_rows(vs_program* pg, const void* src, int64_t lds, const int32_t* idx, void* dst, int64_t ldd,
                               int T, int H) {
  REQUIRE(pg && src && dst, "null operand");
  return vs::prog_add_rows(pg, B(src), lds, idx, 0, 0, B(dst), ldd, T, H);
}
int vs_program_add_embed_rows(vs_program* pg, const void* table, int64_t ldt, const int32_t* ids, void* dst, int64_t ldd,
                              int T, int H, int vocab) {
  REQUIRE(pg && table && ids && dst, "null operand");
  return vs::prog_add_rows(pg, B(table), ldt, ids, vocab, 0, B(dst), ldd, T, H);
}
int vs_program_add_broadcast_row(vs_program* pg, const void* row, void* dst, int64_t ldd, int T, int H) {
  REQUIRE(pg && row && dst, "null operand");
  return vs::prog_add_rows(pg, B(row), H, nullptr, 0, 1, B(dst), ldd, T, H);
}
int vs_program_add_argmax_rows(vs_program* pg, const void* logits, int64_t ldl, int T, int V, int32_t* out) {
  REQUIRE(pg && logits && out, "null operand");
  return vs::prog_add_argmax(pg, B(logits), ldl, T, V, out, nullptr);
}
int vs_program_add_logsoftmax_topk(vs_program* pg, const void* logits, int64_t ldl, int T, int V, int k,
                                   const int32_t* row_sel, float* top_val, int32_t* top_idx) {
  REQUIRE(pg && logits && top_val && top_idx, "null operand");
  return vs::prog_add_lsm_topk(pg, B(logits), ldl, T, V, k, row_sel, top_val, top_idx, nullptr);
}
int vs_program_add_attention(vs_program* pg, const vs_attn_args* a) {
  REQUIRE(pg && a && a->Q && a->K && a->V && a->O, "null operand");
  REQUIRE(a->n_kv_heads > 0 && a->n_heads % a->n_kv_heads == 0, "heads must be a multiple of kv heads");
  vs::AttnArgs g;
  g.Q = B(a->Q); g.q_head_stride = a->q_head_stride; g.K = B(a->K); g.V = B(a->V); g.kv_head_stride = a->kv_head_stride;
  g.n_heads = a->n_heads; g.n_kv_heads = a->n_kv_heads; g.Tq = a->Tq; g.n_ctx = a->n_ctx; g.n_ctx_host = a->n_ctx_host;
  g.n_new = a->n_new; g.causal = a->causal; g.bits = reinterpret_cast<const unsigned long long*>(a->bits);
  g.eager = a->eager; g.scale = a->scale; g.scores = a->scores; g.score_ld = a->score_ld; g.stats = a->stats;
  g.max_chunks = a->max_chunks; g.opart = a->opart; g.pv_splits = a->pv_splits; g.counters = a->counters;
  g.O = B(a->O); g.ldo = a->ldo;
  g.prefix_stable = a->prefix_stable;
  return vs::prog_add_attention(pg, g);
}
int vs_program_add_draft_level0(vs_program* pg, const vs_tree_bufs* tb, int32_t* state) {
  REQUIRE(pg && tb && state, "null operand");
  return vs::prog_add_tree(pg, vs::TREE_LEVEL0, *tb, state, 0, nullptr, nullptr, nullptr, 0, nullptr);
}
int vs_program_add_draft_expand(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, int level) {
  REQUIRE(pg && tb && state, "null operand");
  return vs::prog_add_tree(pg, vs::TREE_EXPAND, *tb, state, level, nullptr, nullptr, nullptr, 0, nullptr);
}
int vs_program_add_tree_build(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, const int32_t* sample_token_dev,
                              int sample_token_host) {
  REQUIRE(pg && tb && state, "null operand");
  return vs::prog_add_tree(pg, vs::TREE_BUILD, *tb, state, 0, nullptr, nullptr, sample_token_dev, sample_token_host, nullptr);
}
int vs_program_add_accept_greedy(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, const int32_t* amax,
                                 int32_t* out_rec) {
  REQUIRE(pg && tb && state && amax, "null operand");
  return vs::prog_add_tree(pg, vs::TREE_ACCEPT, *tb, state, 0, amax, out_rec, nullptr, 0, nullptr);
}
int vs_program_add_accept_sampling(vs_program* pg, const vs_tree_bufs* tb, int32_t* state, const void* logits, int64_t ldl,
                                   int V, float temperature, const float* uni_override, int32_t* out_rec) {
  REQUIRE(pg && tb && state && logits, "null operand");
  return vs::prog_add_accept_sampling(pg, *tb, state, B(logits), ldl, V, temperature, uni_override, out_rec);
}
int vs_program_add_kv_splice(vs_program* pg, void* slab, int n_slabs, int n_kv, int64_t max_pos, const int32_t* state,
                             const int32_t* sel_dev) {
  REQUIRE(pg, "null program");
  return vs::prog_add_splice(pg, B(slab), n_slabs, n_kv, max_pos, state, sel_dev);
}

int vs_program_add_decoder_layer(vs_program* pg, const vs_dims* d, const vs_layer* l, const vs_scratch* s,
                                 const vs_rope* rope, int T, const int32_t* pos_ids, const int32_t* kv_row_base,
                                 int kv_row_off, const int32_t* n_ctx, int n_ctx_host, int n_new, int causal,
                                 const uint64_t* bits, int eager) {
  REQUIRE(pg && d && l && s && rope, "null operand");
  bool has_ssq = false;
  return decoder_layer(pg, d, l, s, rope, T, pos_ids, kv_row_base, kv_row_off, n_ctx, n_ctx_host, n_new, causal, bits, eager,
                       &has_ssq);
}
int vs_program_add_verify_step(vs_program* pg, const vs_target* t, const vs_scratch* s, const int32_t* token_ids,
                               const int32_t* pos_ids, const uint64_t* bits, int T, int32_t* state, void* hidden_out,
                               void* logits, int32_t* amax) {
  REQUIRE(pg && t && s && token_ids && pos_ids && state && hidden_out && logits, "null operand");
  return verify_step(pg, t, s, token_ids, pos_ids, bits, T, state, hidden_out, logits, amax, nullptr);
}
int vs_program_add_draft_forward(vs_program* pg, const vs_draft* dm, const vs_scratch* s, int T, const int32_t* ids,
                                 const int32_t* pos_ids, const int32_t* kv_row_base, int kv_row_off, const int32_t* n_ctx,
                                 int n_new, int causal, const uint64_t* bits) {
  REQUIRE(pg && dm && s && ids && pos_ids, "null operand");
  return draft_forward(pg, dm, s, T, ids, pos_ids, kv_row_base, kv_row_off, n_ctx, n_new, causal, bits);
}
int vs_program_add_draft_levels(vs_program* pg, const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb,
                                int32_t* state, int head_rows, const int32_t* sample_token_dev, int sample_token_host) {
  REQUIRE(pg && dm && s && tb && state, "null operand");
  return draft_levels(pg, dm, s, tb, state, head_rows, sample_token_dev, sample_token_host);
}
int vs_program_add_draft_tree(vs_program* pg, const vs_draft* dm, const vs_scratch* s, const vs_tree_bufs* tb,
                              int32_t* state) {
  REQUIRE(pg && dm && s && tb && state, "null operand");
  return draft_tree(pg, dm, s, tb, state);
}
int vs_program_add_spec_cycle(vs_program* pg, const vs_target* t, const vs_scratch* ts, const vs_draft* dm,
                              const vs_scratch* ds, const vs_tree_bufs* tb, int32_t* state, void* kv_slab,
                              int64_t kv_max_pos, void* hidden_out, void* logits, int32_t* amax, int32_t* out_rec) {
  REQUIRE(pg && t && ts && dm && ds && tb && state && kv_slab && hidden_out && logits && amax, "null operand");
  const int T = tb->total + 1;
  REQUIRE(T <= VS_TOK_BLOCK, "tree larger than 32 nodes");
  vs::MergeInfo mi;
  RET(verify_step(pg, t, ts, tb->draft_tokens, tb->verify_pos, tb->tree_bits, T, state, hidden_out, logits, amax, &mi));
  RET(vs::prog_add_tree(pg, vs::TREE_ACCEPT, *tb, state, 0, amax, out_rec, nullptr, 0, &mi));
  RET(vs::prog_add_splice(pg, B(kv_slab), 2 * t->dims.n_layers, t->dims.n_kv_heads, kv_max_pos, state, tb->sel));
  // accepted hidden states -> draft input rows (utils.py:543-546)
  RET(vs::prog_add_rows(pg, B(hidden_out), t->dims.hidden, tb->sel, VS_TOK_BLOCK, 0, B(dm->x1), 2L * dm->dims.hidden,
                        VS_MAX_ACCEPT, t->dims.hidden));
  return draft_tree(pg, dm, ds, tb, state);
}

int vs_program_add_spec_cycle_sampling(vs_program* pg, const vs_target* t, const vs_scratch* ts, const vs_draft* dm,
                                       const vs_scratch* ds, const vs_tree_bufs* tb, int32_t* state, void* kv_slab,
                                       int64_t kv_max_pos, void* hidden_out, void* logits, float temperature,
                                       int32_t* out_rec) {
  REQUIRE(pg && t && ts && dm && ds && tb && state && kv_slab && hidden_out && logits, "null operand");
  const int T = tb->total + 1;
  REQUIRE(T <= VS_TOK_BLOCK, "tree larger than 32 nodes");
  RET(verify_step(pg, t, ts, tb->draft_tokens, tb->verify_pos, tb->tree_bits, T, state, hidden_out, logits, nullptr, nullptr));
  RET(vs::prog_add_accept_sampling(pg, *tb, state, B(logits), t->dims.vocab, t->dims.vocab, temperature, nullptr, out_rec));
  RET(vs::prog_add_splice(pg, B(kv_slab), 2 * t->dims.n_layers, t->dims.n_kv_heads, kv_max_pos, state, tb->sel));
  RET(vs::prog_add_rows(pg, B(hidden_out), t->dims.hidden, tb->sel, VS_TOK_BLOCK, 0, B(dm->x1), 2L * dm->dims.hidden,
                        VS_MAX_ACCEPT, t->dims.hidden));
  return draft_tree(pg, dm, ds, tb, state);
}

}  // extern "C"
