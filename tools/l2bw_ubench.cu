// Micro-benchmark behind the round-2 GEMM redesign: how much L2 -> SM traffic can ride on top
// of a full-rate HBM weight stream?  One CTA per SM, one thread per CTA keeps a ring of bulk copies (cp.async.bulk,
// the same instruction the program kernel's weight producer uses) in flight and re-issues a slot as soon as it lands.
//   mode hbm   : every CTA streams its own slice of a buffer much larger than L2            -> HBM rate
//   mode l2    : every CTA streams the same small (L2-resident) buffer                      -> L2 -> SM rate
//   mode mix R : per R "weight" copies from the big buffer, one "activation" copy of the same size from the small one
//   the copy size is a parameter (16 KB = the old weight chunk, 3.5-4 KB = a 28-row x 64 chunk of the new layout)
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2bw tools/l2bw_ubench.cu && /tmp/l2bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t hint) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(hint) : "memory");
}
constexpr uint64_t HINT_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t HINT_EVICT_LAST = 0x14F0000000000000ull;
constexpr int MAX_SLOTS = 48;

// ratio_w : ratio_x copies per period; x copies come from `small` (L2 resident), w copies from this CTA's slice of `big`
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* big, size_t big_per_cta, const uint8_t* small,
                                                        size_t small_bytes, int copy_bytes, int n_slots, int n_copies,
                                                        int ratio_w, int ratio_x) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[MAX_SLOTS];
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_slots; ++i) mbar_init(&full[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint8_t* mine = big + (size_t)blockIdx.x * big_per_cta;
  size_t woff = 0, xoff = ((size_t)blockIdx.x * 4096) % small_bytes;
  const int period = ratio_w + ratio_x;
  int issued = 0;
  auto issue = [&](int slot) {
    const bool is_x = period > 0 && (issued % period) >= ratio_w;
    mbar_expect_tx(&full[slot], copy_bytes);
    if (is_x) {
      if (xoff + copy_bytes > small_bytes) xoff = 0;
      bulk_load(smem + (size_t)slot * copy_bytes, small + xoff, copy_bytes, &full[slot], HINT_EVICT_LAST);
      xoff += copy_bytes;
    } else {
      if (woff + copy_bytes > big_per_cta) woff = 0;
      bulk_load(smem + (size_t)slot * copy_bytes, mine + woff, copy_bytes, &full[slot], HINT_EVICT_FIRST);
      woff += copy_bytes;
    }
    ++issued;
  };
  for (int s = 0; s < n_slots && s < n_copies; ++s) issue(s);
  int slot = 0, par = 0;
  for (int done = 0; done < n_copies; ++done) {
    mbar_wait(&full[slot], par);
    if (issued < n_copies) issue(slot);
    if (++slot == n_slots) { slot = 0; par ^= 1; }
  }
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t big_bytes = (size_t)8 << 30;                 // 8 GiB >> 126 MB of L2
  const size_t small_bytes = 1 << 20;                       // 1 MiB: L2 resident (an activation matrix is 0.25-0.7 MB)
  uint8_t *big, *small;
  CK(cudaMalloc(&big, big_bytes));
  CK(cudaMalloc(&small, small_bytes));
  CK(cudaMemset(big, 1, big_bytes));
  CK(cudaMemset(small, 2, small_bytes));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const size_t per_cta = (big_bytes / sms) & ~(size_t)1023;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct Case { const char* name; int copy_bytes, ring_bytes, rw, rx; };
  const Case cases[] = {
      {"hbm_16k_ring96k", 16384, 96 * 1024, 1, 0},  {"hbm_16k_ring192k", 16384, 192 * 1024, 1, 0},
      {"hbm_4k_ring96k", 4096, 96 * 1024, 1, 0},    {"hbm_3584_ring96k", 3584, 96 * 1024 / 3584 * 3584, 1, 0},
      {"hbm_4k_ring160k", 4096, 160 * 1024, 1, 0},
      {"l2_16k_ring96k", 16384, 96 * 1024, 0, 1},   {"l2_4k_ring96k", 4096, 96 * 1024, 0, 1},
      {"l2_16k_ring192k", 16384, 192 * 1024, 0, 1},
      {"mix4to1_16k_ring120k", 16384, 112 * 1024, 4, 1}, {"mix4to1_4k_ring120k", 4096, 120 * 1024, 4, 1},
      {"mix2to1_4k_ring128k", 4096, 128 * 1024, 2, 1},   {"mix1to1_4k_ring128k", 4096, 128 * 1024, 1, 1},
      {"mix1to1_4k_ring192k", 4096, 192 * 1024, 1, 1},   {"mix1to1_16k_ring192k", 16384, 192 * 1024, 1, 1},
      {"mix1to2_4k_ring192k", 4096, 192 * 1024, 1, 2},
  };
  for (const Case& c : cases) {
    int n_slots = c.ring_bytes / c.copy_bytes;
    if (n_slots > MAX_SLOTS) n_slots = MAX_SLOTS;
    const size_t bytes_per_cta = (size_t)192 << 20;                     // 192 MiB per CTA per launch (28 GiB chip-wide)
    const int n_copies = (int)(bytes_per_cta / c.copy_bytes);
    const size_t smem = (size_t)n_slots * c.copy_bytes;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      stream_kernel<<<sms, 128, smem>>>(big, per_cta, small, small_bytes, c.copy_bytes, n_slots, n_copies, c.rw, c.rx);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    const double total = (double)n_copies * c.copy_bytes * sms;
    const double frac_w = c.rw + c.rx > 0 ? (double)c.rw / (c.rw + c.rx) : 0.0;
    printf("{\"case\": \"%s\", \"copy_bytes\": %d, \"slots\": %d, \"in_flight_kb_per_sm\": %.0f, \"ms\": %.3f, "
           "\"total_gbs\": %.0f, \"hbm_part_gbs\": %.0f, \"l2_part_gbs\": %.0f}\n",
           c.name, c.copy_bytes, n_slots, smem / 1024.0, best, total / best / 1e6, total * frac_w / best / 1e6,
           total * (1.0 - frac_w) / best / 1e6);
    fflush(stdout);
  }
  return 0;
}
