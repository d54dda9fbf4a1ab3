// Device bodies of the integer steps of the draft-and-verify cycle: draft-tree expansion bookkeeping, tree build,
// greedy token-accept, draft-KV commit and the KV-cache splice.  All state stays on the device (int32), so a whole
// cycle runs without a host sync; the reference does these steps with .tolist()/.item() round trips and Python loops
// (vispec/model/cnets_ours.py:1109-1238, vispec/model/utils.py:411,438-451,516-546).
//
// Each body is written for `nthreads` cooperating threads with a caller-supplied barrier, so that it runs both as a
// stand-alone one-block kernel (tree.cu) and as a phase of the persistent program kernel (program.cu).
//
// Tie rule: wherever the reference calls torch.topk / argmax on values that can tie (bf16 log-probs), this library
// picks (value descending, index ascending).  torch's own order at exact ties is implementation-defined and differs
// between its CPU and CUDA kernels.
#pragma once
#include "kernels.h"
#include "tree.h"

namespace vs {

constexpr int TREE_MAX_SCORES = 512;     // top_k + depth * top_k^2 <= 8 + 6 * 64

struct TreeSmem {
  unsigned long long keys[TREE_MAX_SCORES];
  int selected[TREE_MAX_SCORES];
  float cu[64];
  int sel[8];
  unsigned long long old_bits[8];
  int sel_sorted[64];
  int parent_slot[64];
  unsigned long long bits[64];
  int is_parent[64];
  int acc[64];
  int n_leaf, max_depth;
  int paths[64][VS_RETRIEVE_LD];     // T > 0: leaf paths before they are sorted
};

// ---- draft level 0: after the stable forward's log-softmax/top-k on the last valid row ------------
// cnets_ours.py:1109-1123: scores = top-k log-probs, parents_list=[0], tree_mask = eye(top_k).
__device__ __forceinline__ void tree_level0_body(const TreeBufs& tb, int* st, int tid) {
  const int k = tb.top_k;
  const int j = tid;
  if (j < k) {
    float v = tb.top_val[j];
    int tok = tb.top_idx[j];
    tb.scores_flat[j] = v;
    tb.tokens_flat[j] = tok;
    tb.frontier_scores[j] = v;
    tb.frontier_ids[j] = tok;
    tb.cs_index[j] = j;
    tb.frontier_bits[j] = 1ull << j;
    tb.gather_idx[j] = st[VS_ST_LAST_ROW];            // input_hidden = last_hidden repeated top_k times
    tb.draft_pos_ids[j] = st[VS_ST_D_POS];            // all siblings share one position (cnets_ours.py:1128)
  }
  if (j == 0) tb.parents[0] = 0;
}

// ---- draft level i (0-based loop index of cnets_ours.py:1126-1165), after that level's forward ------
// needs >= 64 threads
template <class Sync>
__device__ __forceinline__ void tree_expand_body(const TreeBufs& tb, int* st, int level, TreeSmem& sm, int tid, Sync sync) {
  const int k = tb.top_k;
  const int kk = k * k;
  const int bias = 1 + kk * (level > 1 ? level - 1 : 0) + (level > 0 ? k : 0);
  if (tid < k) {
    tb.parents[1 + level * k + tid] = tb.cs_index[tid] + bias;       // uses the PREVIOUS selection
    sm.old_bits[tid] = tb.frontier_bits[tid];
  }
  if (tid < kk) {
    int r = tid / k;
    float v = rbf(tb.top_val[tid] + tb.frontier_scores[r]);           // bf16 add (cnets_ours.py:1151)
    sm.cu[tid] = v;
    sm.keys[tid] = rank_key(v, (uint32_t)tid);
    tb.scores_flat[k + level * kk + tid] = v;
    tb.tokens_flat[k + level * kk + tid] = tb.top_idx[tid];
  }
  sync();
  if (tid < kk) {
    int rank = 0;
    unsigned long long me = sm.keys[tid];
    for (int i = 0; i < kk; ++i) rank += sm.keys[i] > me;
    if (rank < k) sm.sel[rank] = tid;
  }
  sync();
  if (tid < k) {
    int ci = sm.sel[tid];
    int out_id = ci / k;
    tb.cs_index[tid] = ci;
    tb.frontier_scores[tid] = sm.cu[ci];
    tb.frontier_ids[tid] = tb.top_idx[ci];
    tb.gather_idx[tid] = out_id;
    // tree_mask = cat(tree_mask[:, :, out_ids], eye) (cnets_ours.py:1165): inherit the parent's row, add self
    tb.frontier_bits[tid] = sm.old_bits[out_id] | (1ull << (k * (level + 1) + tid));
    tb.draft_pos_ids[tid] = st[VS_ST_D_POS] + level + 1;
  }
}

// ---- tree build (cnets_ours.py:1167-1238) ---------------------------------------------------------
template <class Sync>
__device__ __forceinline__ void tree_build_body(const TreeBufs& tb, int* st, const int* sample_token_ptr,
                                                int sample_token_host, TreeSmem& sm, int tid, int nthreads, Sync sync) {
  const int k = tb.top_k;
  const int n_sc = k + tb.depth * k * k;
  const int total = tb.total;                 // drafted nodes; tree has total+1 nodes
  const int n = total + 1;

  for (int i = tid; i < n_sc; i += nthreads) sm.keys[i] = rank_key(tb.scores_flat[i], (uint32_t)i);
  if (tid < 64) sm.is_parent[tid] = 0;
  if (tid == 0) { sm.n_leaf = 0; sm.max_depth = 0; }
  sync();
  // top-`total` by (score desc, index asc), then re-sorted by flat index (torch.sort of the indices)
  for (int i = tid; i < n_sc; i += nthreads) {
    unsigned long long me = sm.keys[i];
    int rank = 0;
    for (int j = 0; j < n_sc; ++j) rank += sm.keys[j] > me;
    sm.selected[i] = rank < total;
  }
  sync();
  for (int i = tid; i < n_sc; i += nthreads) {
    if (!sm.selected[i]) continue;
    int pos = 0;
    for (int j = 0; j < i; ++j) pos += sm.selected[j];
    sm.sel_sorted[pos] = i;
  }
  sync();
  if (tid < total) {
    int fi = sm.sel_sorted[tid];
    tb.draft_tokens[tid + 1] = tb.tokens_flat[fi];
    int dp = tb.parents[fi / k];                 // node id of the parent (0 = root, else flat index + 1)
    int slot = 0;
    if (dp != 0) {
      int cnt = 0;                               // searchsorted(top_idx, dp - 1, left)
      for (int j = 0; j < total; ++j) cnt += sm.sel_sorted[j] < dp - 1;
      slot = cnt + 1;
    }
    sm.parent_slot[tid + 1] = slot;
  }
  if (tid == 0) {
    tb.draft_tokens[0] = sample_token_ptr ? *sample_token_ptr : sample_token_host;
    sm.parent_slot[0] = 0;
  }
  sync();
  if (tid < n) {
    unsigned long long b = 1ull;                 // column 0 (root) is visible to every node
    int node = tid;
    int guard = 0;
    while (node != 0 && guard++ < 64) {
      b |= 1ull << node;
      node = sm.parent_slot[node];
    }
    sm.bits[tid] = b;
    int pos = __popcll(b) - 1;
    tb.tree_bits[tid] = b;
    tb.tree_pos[tid] = pos;
    tb.verify_pos[tid] = pos + st[VS_ST_CUR_LEN] + st[VS_ST_ROPE_DELTA];
    atomicMax(&sm.max_depth, pos + 1);
    if (tid > 0) sm.is_parent[sm.parent_slot[tid]] = 1;
  }
  sync();
  const bool sort_paths = st[VS_ST_TEMP_BITS] != 0;      // T > 0: rows sorted like cnets_ours.py:1215-1224
  int my_rank = -1;
  if (tid < n && !sm.is_parent[tid]) {
    int rank = 0;
    for (int j = 0; j < tid; ++j) rank += !sm.is_parent[j];
    my_rank = rank;
    int pos = __popcll(sm.bits[tid]) - 1;
    int* row = sort_paths ? sm.paths[rank] : tb.retrieve + rank * VS_RETRIEVE_LD;
    for (int j = 0; j < VS_RETRIEVE_LD; ++j) row[j] = -1;
    int node = tid;
    for (int j = pos; j >= 0; --j) {
      row[j] = node;
      node = sm.parent_slot[node];
    }
    atomicAdd(&sm.n_leaf, 1);
  }
  sync();
  if (sort_paths && my_rank >= 0) {
    // lexicographic order with -1 (padding) sorting last; paths are distinct, so ranks are a permutation
    const int big = tb.total + 5;
    int before = 0;
    for (int o = 0; o < sm.n_leaf; ++o) {
      if (o == my_rank) continue;
      bool less = false;                                   // paths[o] < paths[my_rank] ?
      for (int j = 0; j < VS_RETRIEVE_LD; ++j) {
        const int a = sm.paths[o][j] >= 0 ? sm.paths[o][j] : big, b = sm.paths[my_rank][j] >= 0 ? sm.paths[my_rank][j] : big;
        if (a != b) { less = a < b; break; }
      }
      before += less;
    }
    int* row = tb.retrieve + before * VS_RETRIEVE_LD;
    for (int j = 0; j < VS_RETRIEVE_LD; ++j) row[j] = sm.paths[my_rank][j];
  }
  if (tid == 0) {
    st[VS_ST_N_LEAVES] = sm.n_leaf;
    st[VS_ST_MAX_DEPTH] = sm.max_depth;
  }
}

// ---- greedy accept (utils.py:438-451 + the gathers of :411, :516-526, :543-554) -------------------
// needs >= 64 threads
template <class Sync>
__device__ __forceinline__ void tree_accept_body(const TreeBufs& tb, int* st, const int* amax, int* out_rec,
                                                 TreeSmem& sm, int tid, Sync sync) {
  const int n_leaves = st[VS_ST_N_LEAVES];
  const int max_depth = st[VS_ST_MAX_DEPTH];
  if (tid < n_leaves) {
    const int* row = tb.retrieve + tid * VS_RETRIEVE_LD;
    int a = 0;
    for (int d = 1; d < max_depth; ++d) {
      int node = row[d];
      if (node < 0) break;
      if (tb.draft_tokens[node] != amax[row[d - 1]]) break;
      ++a;
    }
    sm.acc[tid] = a;
  }
  sync();
  if (tid == 0) {
    int best = 0, a = 0;
    for (int l = 0; l < n_leaves; ++l)
      if (sm.acc[l] > a) { a = sm.acc[l]; best = l; }          // first maximum, 0 when nothing is accepted
    const int* row = tb.retrieve + best * VS_RETRIEVE_LD;
    const int prev = st[VS_ST_CUR_LEN];
    st[VS_ST_PREV_LEN] = prev;
    st[VS_ST_CUR_LEN] = prev + a + 1;
    st[VS_ST_ACCEPT] = a;
    st[VS_ST_BEST] = best;
    const int next_tok = amax[row[a]];
    st[VS_ST_NEXT_TOKEN] = next_tok;
    st[VS_ST_M] = a + 1;
    st[VS_ST_LAST_ROW] = a;
    const int dpos = st[VS_ST_D_POS];
    for (int j = 0; j < VS_MAX_ACCEPT; ++j) {
      int jj = j <= a ? j : a;
      int node = row[jj];
      tb.sel[j] = node;                                   // tree slots of the accepted path (KV splice, hidden gather)
      // draft stable-forward inputs: feature of token t pairs with token t+1 (cnets_ours.py:1084-1091)
      tb.stable_ids[j] = j < a ? tb.draft_tokens[row[j + 1]] : next_tok;
      tb.stable_pos[j] = dpos + j;
    }
    if (out_rec) {                                        // host-visible record of this cycle
      out_rec[0] = a;
      out_rec[1] = next_tok;
      for (int j = 0; j < VS_MAX_ACCEPT; ++j) out_rec[2 + j] = j <= a ? tb.draft_tokens[row[j]] : -1;
    }
    st[VS_ST_CYCLE] += 1;
  }
}

// ---- T > 0 acceptance + sampling (utils.py:453-493, :550-552) -----------------------------------------------------
// Philox4x32-10 (counter-based, Salmon et al.): one call gives four 32-bit words
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float philox_uniform(unsigned int seed, unsigned int cycle, unsigned int index) {
  const uint4 r = philox4x32(make_uint4(index, cycle, 0u, 0u), make_uint2(seed, 0x5bd1e995u));
  return (float)(r.x >> 8) * (1.0f / 16777216.0f);        // [0, 1)
}

// ---- top-k logits warper (utils.py:39-55 -> transformers TopKLogitsWarper: drop every logit below the k-th largest of its
// row; ties with the k-th value stay).  Logits are bf16, so the threshold is found exactly by a two-level radix select on
// the order-preserving 16-bit key: a histogram of the high byte over the row, then of the low byte inside the selected bin.
__device__ __forceinline__ unsigned int bf16_key(bf16 v) {
  const unsigned int b = __bfloat16_as_ushort(v);
  return (b & 0x8000u) ? (~b & 0xffffu) : (b | 0x8000u);
}
// (one warp) the bin that holds the k-th largest entry of a 256-bin histogram, counting from the top bin; `above` = entries
// in higher bins.  k >= 1 and k <= total count.
__device__ __forceinline__ int hist_select(const unsigned int* h, unsigned int k, int lane, unsigned int& above) {
  const uint4 a = __ldcg(reinterpret_cast<const uint4*>(h + lane * 8)), b = __ldcg(reinterpret_cast<const uint4*>(h + lane * 8 + 4));
  const unsigned int c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  unsigned int s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j];
  unsigned int incl = s;                                         // entries in this lane's bins and above
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += t;
  }
  const unsigned int A = incl - s;
  const bool mine = A < k && k <= incl;
  const unsigned int m = __ballot_sync(0xffffffffu, mine);
  int bin = 0;
  unsigned int ab = A;
  if (mine) {
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      if (ab + c[j] >= k) { bin = lane * 8 + j; break; }
      ab += c[j];
    }
  }
  const int src = m ? __ffs(m) - 1 : 0;                          // (m == 0 only if k exceeds the count: keep everything)
  bin = __shfl_sync(0xffffffffu, bin, src);
  above = __shfl_sync(0xffffffffu, ab, src);
  if (!m) { bin = 0; above = 0; }
  return bin;
}
// (one warp) key of the k-th largest logit of a row from its two histograms
__device__ __forceinline__ unsigned int topk_threshold_key(const unsigned int* hist_row, int k, int lane) {
  unsigned int above, above2;
  const int b1 = hist_select(hist_row, (unsigned int)k, lane, above);
  const int b2 = hist_select(hist_row + 256, (unsigned int)k - above, lane, above2);
  return ((unsigned int)b1 << 8) | (unsigned int)b2;
}


struct SampleSmem {
  float px[32 * VS_RETRIEVE_LD];      // target probability of candidate (leaf, depth) under its prefix node
  float uni[32 * VS_RETRIEVE_LD];
  int cand[32 * VS_RETRIEVE_LD];      // candidate tokens (-1 = padding)
  float rowM[32], rowL[32];           // softmax statistics of the tree rows under the temperature
  float slice_mass[64];
  float warp_tot[16];
  int removed[VS_RETRIEVE_LD * 8];    // tokens rejected at the failing depth (at most top_k)
  int n_removed, row, accept, best, slice, picked;
  float target;
  unsigned int rowTheta[32];          // top-k warper: key of the k-th largest logit of each tree row (0 = keep everything)
};

// logits [T][ldl] bf16; lstats [T][n_slices] (max, sum exp) of logits * inv_temp.  512 threads of CTA 0.
template <class Sync>
__device__ __forceinline__ void tree_accept_sample_body(const TreeBufs& tb, int* st, const bf16* logits, long ldl, int V,
                                                        const float2* lstats, int n_slices, float inv_temp,
                                                        const float* uni_override, int* out_rec, SampleSmem& sm, int tid,
                                                        int nthreads, Sync sync, int top_k = 0, unsigned int* hist = nullptr) {
  const int n_leaves = st[VS_ST_N_LEAVES];
  const int max_depth = st[VS_ST_MAX_DEPTH];
  const int T = tb.total + 1;
  const unsigned int seed = (unsigned int)st[VS_ST_RNG_SEED], cycle = (unsigned int)st[VS_ST_RNG_CTR];
  // row statistics: combine the vocabulary slices (fixed order)
  if (tid < T) {
    float m = -INFINITY;
    for (int s = 0; s < n_slices; ++s) m = fmaxf(m, __ldcg(&lstats[tid * n_slices + s]).x);
    float l = 0.f;
    for (int s = 0; s < n_slices; ++s) {
      const float2 ml = __ldcg(&lstats[tid * n_slices + s]);
      if (ml.x != -INFINITY) l += ml.y * expf(ml.x - m);
    }
    sm.rowM[tid] = m;
    sm.rowL[tid] = l;
  }
  // top-k warper: the statistics above already cover the kept logits only; the same threshold gates the candidates'
  // probabilities and the residual sampling below (one warp per row)
  for (int r = tid >> 5; r < 32; r += nthreads >> 5) {
    const unsigned int th = (top_k > 0 && r < T) ? topk_threshold_key(hist + (long)r * 512, top_k, tid & 31) : 0u;
    if ((tid & 31) == 0) sm.rowTheta[r] = th;
  }
  sync();
  auto kept = [&](int row_, bf16 v) -> bool { return bf16_key(v) >= sm.rowTheta[row_]; };
  // candidate tokens, their probabilities under the prefix node, and the uniforms
  for (int idx = tid; idx < n_leaves * VS_RETRIEVE_LD; idx += nthreads) {
    const int j = idx / VS_RETRIEVE_LD, i = idx % VS_RETRIEVE_LD;
    const int node = i < max_depth ? tb.retrieve[j * VS_RETRIEVE_LD + i] : -1;
    const int tok = node >= 0 ? tb.draft_tokens[node] : -1;
    sm.cand[idx] = tok;
    float p = 0.f;
    if (i >= 1 && tok >= 0) {
      const int row = tb.retrieve[j * VS_RETRIEVE_LD + i - 1];
      const bf16 lv = logits[(long)row * ldl + tok];
      const float x = __bfloat162float(lv) * inv_temp;
      p = kept(row, lv) ? expf(x - sm.rowM[row]) / sm.rowL[row] : 0.f;
    }
    sm.px[idx] = p;
    sm.uni[idx] = uni_override ? uni_override[idx] : philox_uniform(seed, cycle, (unsigned int)idx);
  }
  sync();
  if (tid == 0) {
    // the sequential accept / reject walk of utils.py:457-488 (at most depth * top_k steps)
    int accept_length = 1, best = 0, n_removed = 0, fail_row = -1;
    bool adjust = false;
    for (int i = 1; i < max_depth; ++i) {
      if (i != accept_length) break;
      adjust = false;
      n_removed = 0;
      float removed_mass = 0.f;
      int fi = -1;
      for (int j = 0; j < n_leaves && fi < 0; ++j) {
        bool eq = true;
        for (int d = 0; d < accept_length; ++d) eq = eq && sm.cand[j * VS_RETRIEVE_LD + d] == sm.cand[best * VS_RETRIEVE_LD + d];
        if (eq) fi = j;
      }
      fail_row = tb.retrieve[fi * VS_RETRIEVE_LD + i - 1];
      bool accepted = false;
      for (int j = 0; j < n_leaves && !accepted; ++j) {
        bool eq = true;
        for (int d = 0; d < accept_length; ++d) eq = eq && sm.cand[j * VS_RETRIEVE_LD + d] == sm.cand[best * VS_RETRIEVE_LD + d];
        if (!eq) continue;
        const int x = sm.cand[j * VS_RETRIEVE_LD + i];
        if (x < 0) continue;
        bool seen = false;
        for (int r = 0; r < n_removed; ++r) seen = seen || sm.removed[r] == x;
        if (seen) continue;
        // gtp[x] after the renormalisations so far: p(x) / (1 - mass removed)
        const float px = sm.px[j * VS_RETRIEVE_LD + i] / (1.0f - removed_mass);
        if (sm.uni[j * VS_RETRIEVE_LD + i] <= px) {
          best = j;                                    // the accepted prefix now runs through leaf j
          ++accept_length;
          accepted = true;
        } else {
          sm.removed[n_removed++] = x;
          removed_mass += sm.px[j * VS_RETRIEVE_LD + i];
          adjust = true;
        }
      }
    }
    const bool residual = adjust && accept_length != max_depth;
    sm.accept = accept_length - 1;
    sm.best = best;
    sm.n_removed = residual ? n_removed : 0;
    sm.row = residual ? fail_row : tb.retrieve[best * VS_RETRIEVE_LD + accept_length - 1];
  }
  sync();
  // ---- sample the next root token from row `sm.row` with the rejected tokens removed: inverse CDF over the vocabulary
  const int row = sm.row;
  const float M = sm.rowM[row];
  const bf16* lr = logits + (long)row * ldl;
  if (tid < n_slices) {
    const int per = ((V + n_slices - 1) / n_slices + 7) & ~7;
    const int lo = min(V, tid * per), hi = min(V, lo + per);
    const float2 ml = __ldcg(&lstats[row * n_slices + tid]);
    float mass = ml.x == -INFINITY ? 0.f : ml.y * expf(ml.x - M);
    for (int r = 0; r < sm.n_removed; ++r) {
      const int x = sm.removed[r];
      if (x >= lo && x < hi && kept(row, lr[x])) mass -= expf(__bfloat162float(lr[x]) * inv_temp - M);
    }
    sm.slice_mass[tid] = fmaxf(mass, 0.f);
  }
  sync();
  if (tid == 0) {
    float total = 0.f;
    for (int s = 0; s < n_slices; ++s) total += sm.slice_mass[s];
    const float u = uni_override ? uni_override[32 * VS_RETRIEVE_LD] : philox_uniform(seed, cycle, 4096u);
    float target = u * total;
    int s = 0;
    while (s + 1 < n_slices && target >= sm.slice_mass[s]) { target -= sm.slice_mass[s]; ++s; }
    sm.slice = s;
    sm.target = target;
    sm.picked = 0x7fffffff;
  }
  sync();
  {
    const int per = ((V + n_slices - 1) / n_slices + 7) & ~7;
    const int lo = min(V, sm.slice * per), hi = min(V, lo + per);
    const int n = hi - lo;
    const int chunk = (n + nthreads - 1) / nthreads;               // consecutive elements per thread
    const int e0 = lo + tid * chunk, e1 = min(hi, e0 + chunk);
    float local = 0.f;
    for (int e = e0; e < e1; ++e) {
      bool rem = !kept(row, lr[e]);
      for (int r = 0; r < sm.n_removed; ++r) rem = rem || sm.removed[r] == e;
      if (!rem) local += expf(__bfloat162float(lr[e]) * inv_temp - M);
    }
    // exclusive prefix over the threads: warp scan, then the warp totals
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += t;
    }
    if ((tid & 31) == 31) sm.warp_tot[tid >> 5] = incl;
    sync();
    float base = 0.f;
    for (int w = 0; w < (tid >> 5); ++w) base += sm.warp_tot[w];
    float cum = base + incl - local;                              // mass before this thread's first element
    for (int e = e0; e < e1; ++e) {
      bool rem = !kept(row, lr[e]);
      for (int r = 0; r < sm.n_removed; ++r) rem = rem || sm.removed[r] == e;
      if (rem) continue;
      cum += expf(__bfloat162float(lr[e]) * inv_temp - M);
      if (cum > sm.target) { atomicMin(&sm.picked, e); break; }
    }
    sync();
    if (tid == 0 && sm.picked == 0x7fffffff) {                    // rounding at the very end of the slice: last kept token
      int e = hi - 1;
      for (;; --e) {
        bool rem = !kept(row, lr[e]);
        for (int r = 0; r < sm.n_removed; ++r) rem = rem || sm.removed[r] == e;
        if (!rem || e == lo) break;
      }
      sm.picked = e;
    }
  }
  sync();
  if (top_k > 0)                                                   // leave the histograms zero for the next launch
    for (int i = tid; i < T * 512; i += nthreads) hist[i] = 0u;
  if (tid == 0) {
    const int a = sm.accept, best = sm.best, next_tok = sm.picked;
    const int* rowp = tb.retrieve + best * VS_RETRIEVE_LD;
    const int prev = st[VS_ST_CUR_LEN];
    st[VS_ST_PREV_LEN] = prev;
    st[VS_ST_CUR_LEN] = prev + a + 1;
    st[VS_ST_ACCEPT] = a;
    st[VS_ST_BEST] = best;
    st[VS_ST_NEXT_TOKEN] = next_tok;
    st[VS_ST_M] = a + 1;
    st[VS_ST_LAST_ROW] = a;
    const int dpos = st[VS_ST_D_POS];
    for (int j = 0; j < VS_MAX_ACCEPT; ++j) {
      const int jj = j <= a ? j : a;
      tb.sel[j] = rowp[jj];
      tb.stable_ids[j] = j < a ? tb.draft_tokens[rowp[j + 1]] : next_tok;
      tb.stable_pos[j] = dpos + j;
    }
    if (out_rec) {
      out_rec[0] = a;
      out_rec[1] = next_tok;
      for (int j = 0; j < VS_MAX_ACCEPT; ++j) out_rec[2 + j] = j <= a ? tb.draft_tokens[rowp[j]] : -1;
    }
    st[VS_ST_CYCLE] += 1;
    st[VS_ST_RNG_CTR] = (int)(cycle + 1u);
  }
}

// after the stable forward: the m = accept+1 new rows become part of the draft's stable KV
__device__ __forceinline__ void tree_commit_body(int* st, int tid) {
  if (tid == 0) {
    int m = st[VS_ST_M];
    st[VS_ST_D_LEN] += m;
    st[VS_ST_D_POS] += m;
  }
}

// ---- KV splice (utils.py:529-541): one (slab, kv head) item per team of 128 threads ---------------------------
struct SpliceArgs {
  bf16* slab;               // [n_slabs][n_kv][max_pos][128]
  long head_stride;         // max_pos * 128
  const int* st;            // device state (prev_len, accept) or null
  const int* sel_dev;       // device tree slots or null
  int prev_len_host, n_host;
  int sel_host[VS_MAX_ACCEPT];
};
template <class Sync>
__device__ __forceinline__ void kv_splice_item(const SpliceArgs& p, long item, int tid, Sync sync) {
  const int prev = p.st ? __ldcg(p.st + VS_ST_PREV_LEN) : p.prev_len_host;
  const int n = p.st ? __ldcg(p.st + VS_ST_ACCEPT) + 1 : p.n_host;
  bf16* base = p.slab + item * p.head_stride;
  const int r = tid >> 4, ch = tid & 15;          // 8 rows x 16 chunks of 16 bytes
  uint4 v = make_uint4(0, 0, 0, 0);
  int src = 0;
  bool act = r < n;
  if (act) {
    src = p.sel_dev ? __ldcg(p.sel_dev + r) : p.sel_host[r];
    act = src != r;                                // row already in place
    if (act) v = __ldcg(reinterpret_cast<const uint4*>(base + (long)(prev + src) * 128 + ch * 8));
  }
  sync();                                          // gather everything before any overwrite (reference gathers into a temp)
  if (act) *reinterpret_cast<uint4*>(base + (long)(prev + r) * 128 + ch * 8) = v;
}

}  // namespace vs
