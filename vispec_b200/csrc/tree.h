// Internal declarations of the integer (tree / accept / splice) launchers.
#pragma once
#include "../../include/vispec_b200.h"
#include "kernels.h"

namespace vs {
typedef vs_tree_bufs TreeBufs;

cudaError_t launch_draft_level0(const TreeBufs& tb, int* st, cudaStream_t s);
cudaError_t launch_draft_expand(const TreeBufs& tb, int* st, int level, cudaStream_t s);
cudaError_t launch_tree_build(const TreeBufs& tb, int* st, const int* sample_token_dev, int sample_token_host,
                              cudaStream_t s);
cudaError_t launch_accept_greedy(const TreeBufs& tb, int* st, const int* amax, int* out_rec, cudaStream_t s);
cudaError_t launch_draft_commit(int* st, cudaStream_t s);
cudaError_t launch_kv_splice(bf16* slab, int n_slabs, int n_kv, long max_pos, const int* st, const int* sel_dev,
                             int prev_len_host, int n_host, const int* sel_host, cudaStream_t s);
cudaError_t launch_set_state(int* st, int slot, int value, cudaStream_t s);
}  // namespace vs
