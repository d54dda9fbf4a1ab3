// Stand-alone launchers of the integer steps of the draft-and-verify cycle (bodies and reference citations in
// tree_dev.cuh).  The persistent program kernel (program.cu) runs the same bodies as phases of one launch.
#include "kernels.h"
#include "profile.h"
#include "tree.h"
#include "tree_dev.cuh"

namespace vs {

namespace {

struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

__global__ void draft_level0_kernel(TreeBufs tb, int* st) { tree_level0_body(tb, st, threadIdx.x); }

__global__ void draft_expand_kernel(TreeBufs tb, int* st, int level) {
  __shared__ TreeSmem sm;
  tree_expand_body(tb, st, level, sm, threadIdx.x, BlockSync());
}

__global__ void __launch_bounds__(256) tree_build_kernel(TreeBufs tb, int* st, const int* sample_token_ptr,
                                                         int sample_token_host) {
  __shared__ TreeSmem sm;
  tree_build_body(tb, st, sample_token_ptr, sample_token_host, sm, threadIdx.x, blockDim.x, BlockSync());
}

__global__ void accept_greedy_kernel(TreeBufs tb, int* st, const int* amax, int* out_rec) {
  __shared__ TreeSmem sm;
  tree_accept_body(tb, st, amax, out_rec, sm, threadIdx.x, BlockSync());
}

__global__ void draft_commit_kernel(int* st) { tree_commit_body(st, threadIdx.x); }

__global__ void __launch_bounds__(128) kv_splice_kernel(SpliceArgs p) {
  kv_splice_item(p, (long)blockIdx.x * gridDim.y + blockIdx.y, threadIdx.x, BlockSync());
}

__global__ void set_state_kernel(int* st, int slot, int value) { st[slot] = value; }

}  // namespace

cudaError_t launch_draft_level0(const TreeBufs& tb, int* st, cudaStream_t s) {
  draft_level0_kernel<<<1, 32, 0, s>>>(tb, st);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_draft_expand(const TreeBufs& tb, int* st, int level, cudaStream_t s) {
  if (tb.top_k > 8) return cudaErrorInvalidValue;
  draft_expand_kernel<<<1, 64, 0, s>>>(tb, st, level);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_tree_build(const TreeBufs& tb, int* st, const int* sample_token_dev, int sample_token_host,
                              cudaStream_t s) {
  int n_sc = tb.top_k + tb.depth * tb.top_k * tb.top_k;
  if (tb.total + 1 > 64 || tb.total > n_sc || n_sc > TREE_MAX_SCORES) return cudaErrorInvalidValue;
  tree_build_kernel<<<1, 256, 0, s>>>(tb, st, sample_token_dev, sample_token_host);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_accept_greedy(const TreeBufs& tb, int* st, const int* amax, int* out_rec, cudaStream_t s) {
  accept_greedy_kernel<<<1, 64, 0, s>>>(tb, st, amax, out_rec);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_draft_commit(int* st, cudaStream_t s) {
  draft_commit_kernel<<<1, 32, 0, s>>>(st);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_kv_splice(bf16* slab, int n_slabs, int n_kv, long max_pos, const int* st, const int* sel_dev,
                             int prev_len_host, int n_host, const int* sel_host, cudaStream_t s) {
  SpliceArgs a;
  a.slab = slab;
  a.head_stride = max_pos * 128;
  a.st = st;
  a.sel_dev = sel_dev;
  a.prev_len_host = prev_len_host;
  a.n_host = n_host;
  for (int i = 0; i < VS_MAX_ACCEPT; ++i) a.sel_host[i] = (sel_host && i < n_host) ? sel_host[i] : 0;
  if (!st && (n_host < 0 || n_host > VS_MAX_ACCEPT)) return cudaErrorInvalidValue;
  kv_splice_kernel<<<dim3(n_slabs, n_kv), 128, 0, s>>>(a);
  count_launch();
  return cudaGetLastError();
}
cudaError_t launch_set_state(int* st, int slot, int value, cudaStream_t s) {
  set_state_kernel<<<1, 1, 0, s>>>(st, slot, value);
  count_launch();
  return cudaGetLastError();
}

}  // namespace vs
