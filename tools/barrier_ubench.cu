// Micro-benchmark of grid-barrier variants on a persistent grid of one CTA per SM (148 x 256 threads).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/barrier_ubench tools/barrier_ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) k(unsigned* ctr, int iters, float* sink, const float* data) {
  const int tid = threadIdx.x, G = gridDim.x;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    // a little "work": one global store per thread so the release has something to order
    sink[(blockIdx.x * 256 + tid)] = acc + it;
    const unsigned target = (unsigned)(it + 1) * G;
    if (VARIANT == 0) {          // bar; t0: release-red, relaxed spin; bar
      __syncthreads();
      if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while (ld_relaxed(ctr) < target) {}
      }
      __syncthreads();
    } else if (VARIANT == 1) {   // every thread fences, then bar; t0 relaxed red + spin
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while (ld_relaxed(ctr) < target) {}
      }
      __syncthreads();
    } else if (VARIANT == 2) {   // no ordering at all (lower bound): bar; relaxed red; spin; bar
      __syncthreads();
      if (tid == 0) {
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while (ld_relaxed(ctr) < target) {}
      }
      __syncthreads();
    } else if (VARIANT == 3) {   // classic: bar; t0: threadfence, atomicAdd, acquire spin, threadfence; bar
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire(ctr) < target) {}
        __threadfence();
      }
      __syncthreads();
    } else if (VARIANT == 4) {   // no store work, no ordering: pure arrive + spin
      __syncthreads();
      if (tid == 0) {
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while (ld_relaxed(ctr) < target) {}
      }
      __syncthreads();
    } else if (VARIANT == 5) {   // one warp polls with 32 lanes? no: flag per CTA, CTA 0 gathers and broadcasts a generation
      __syncthreads();
      if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while (ld_relaxed(ctr) < target) { __nanosleep(20); }
      }
      __syncthreads();
    }
    acc += data[tid & 31];
  }
  if (acc == 12345.f) sink[0] = acc;
}

template <int V>
void run(const char* name, unsigned* ctr, float* sink, float* data, int G) {
  const int iters = 2000;
  cudaMemset(ctr, 0, 4);
  void* args[] = {&ctr, (void*)&iters, &sink, &data};
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchCooperativeKernel((void*)k<V>, dim3(G), dim3(256), args, 0, 0);   // warm
  cudaDeviceSynchronize();
  cudaMemset(ctr, 0, 4);
  cudaEventRecord(a);
  cudaLaunchCooperativeKernel((void*)k<V>, dim3(G), dim3(256), args, 0, 0);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("%-52s %7.3f us / barrier  (%s)\n", name, ms * 1e3 / iters, cudaGetErrorString(e));
}

int main() {
  int G = 0; cudaDeviceGetAttribute(&G, cudaDevAttrMultiProcessorCount, 0);
  unsigned* ctr; float *sink, *data;
  cudaMalloc(&ctr, 256); cudaMalloc(&sink, G * 256 * 4); cudaMalloc(&data, 256); cudaMemset(data, 0, 256);
  printf("grid %d x 256\n", G);
  run<0>("bar; red.release + relaxed spin; bar", ctr, sink, data, G);
  run<1>("all threads __threadfence; bar; relaxed red + spin", ctr, sink, data, G);
  run<2>("no ordering: relaxed red + relaxed spin", ctr, sink, data, G);
  run<3>("threadfence + atomicAdd + acquire spin + fence", ctr, sink, data, G);
  run<5>("red.release + relaxed spin with nanosleep(20)", ctr, sink, data, G);
  return 0;
}
