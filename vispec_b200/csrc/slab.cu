// Slab weight layout: plan + packing kernel (see slab.h) and their C ABI (include/vispec_b200.h, "Slab weights").
#include "slab.h"

#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/vispec_b200.h"
#include "common.cuh"
#include "program.h"

namespace vs {

int slab_default_packed() {
  const char* e = getenv("VS_SLAB_PACKED");
  if (e && e[0] >= '0' && e[0] <= '2') return e[0] - '0';
  return 1;
}

bool slab_plan(int N, int K, int pair, int grid, int packed, SlabPlan* out) {
  if (N < 1 || K < 64 || K % 64 != 0 || grid < 1) return false;
  if (pair && N % 128 != 0) return false;
  SlabPlan p;
  p.pair = pair ? 1 : 0;
  p.packed = packed;
  p.E = pair ? N / 2 : N;
  p.nk = K / 64;
  p.grid = grid;
  for (p.passes = 1; p.passes <= 64; ++p.passes) {
    const long units = (long)grid * p.passes;
    const int max_e = (int)((p.E + units - 1) / units);
    if (pair) {
      p.n_tiles = 2;
      p.rows_alloc = max_e;
    } else {
      p.n_tiles = max_e > SLAB_MAX_ROWS ? 2 : 1;
      p.rows_alloc = (max_e + p.n_tiles - 1) / p.n_tiles;
    }
    p.blk_bytes = p.rows_alloc * 128;
    p.blk_smem = packed ? p.blk_bytes : (p.blk_bytes + 1023) / 1024 * 1024;
    if (p.rows_alloc <= SLAB_MAX_ROWS && p.n_tiles * p.blk_smem + SLAB_X_BYTES <= SLAB_SLOT_BYTES) break;
  }
  if (p.passes > 64) return false;
  const long units = (long)grid * p.passes;
  p.e_per = (int)(p.E / units);
  p.e_rem = (int)(p.E % units);
  // chunks per stage: as many as fit one slot (weights, then the 1024-aligned activation blocks)
  p.kc = 1;
  for (int kc = 2; kc <= 8 && kc <= p.nk; ++kc) {
    const int w = (kc * p.n_tiles * p.blk_smem + 1023) / 1024 * 1024;
    if (w + kc * SLAB_X_BYTES <= SLAB_SLOT_BYTES) p.kc = kc;
  }
  p.x_off = (p.kc * p.n_tiles * p.blk_smem + 1023) / 1024 * 1024;
  {
    const char* e = getenv("VS_SLAB_M64");
    p.m64 = (p.rows_alloc <= 64 && !(e && e[0] == '0')) ? 1 : 0;
  }
  p.unit_stride = (long)p.nk * p.n_tiles * p.blk_bytes;
  p.total_bytes = units * p.unit_stride;
  *out = p;
  return true;
}

namespace {

// one block of 256 threads per (k-chunk, unit): 16-byte pieces (tile, row, chunk)
__global__ void slab_pack_kernel(const bf16* __restrict__ W, long ldw, int N, SlabPlan p, uint8_t* __restrict__ dst) {
  const int j = blockIdx.x, u = blockIdx.y;
  const int e_lo = u * p.e_per + min(u, p.e_rem), cnt = p.e_per + (u < p.e_rem ? 1 : 0);
  const int jj = j % p.kc;                                   // position of the chunk inside its stage
  uint8_t* ublk = dst + (long)u * p.unit_stride + (long)j * p.n_tiles * p.blk_bytes;
  const int pieces = p.n_tiles * p.rows_alloc * 8;
  for (int i = threadIdx.x; i < pieces; i += blockDim.x) {
    const int c = i & 7, r = (i >> 3) % p.rows_alloc, t = (i >> 3) / p.rows_alloc;
    int n = -1;
    if (p.pair) {
      if (r < cnt) {
        const int e = e_lo + r;
        n = (e / 64) * 128 + (e % 64) + 64 * t;
      }
    } else {
      const int e = e_lo + t * p.rows_alloc + r;
      if (e < e_lo + cnt) n = e;
    }
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n >= 0 && n < N) v = *reinterpret_cast<const uint4*>(W + (long)n * ldw + (long)j * 64 + c * 8);
    // shared-memory offset of the row inside its stage decides the swizzle phase (slab.h)
    const int o = (jj * p.n_tiles + t) * p.blk_smem + r * 128;
    const int cs = c ^ (p.packed == 2 ? (r & 7) : ((o >> 7) & 7));
    *reinterpret_cast<uint4*>(ublk + (long)t * p.blk_bytes + r * 128 + cs * 16) = v;
  }
}

}  // namespace
}  // namespace vs

extern "C" {

size_t vs_slab_bytes(int N, int K, int epi, int grid) {
  vs::SlabPlan p;
  const int pair = epi == vs::EPI_SWIGLU || epi == vs::EPI_HEADS;
  if (grid <= 0) grid = vs_device_sms();
  if (!vs::slab_plan(N, K, pair, grid, vs::slab_default_packed(), &p)) return 0;
  return (size_t)p.total_bytes;
}

int vs_slab_plan(int N, int K, int epi, int grid, int32_t* out, int cap) {
  vs::SlabPlan p;
  const int pair = epi == vs::EPI_SWIGLU || epi == vs::EPI_HEADS;
  if (grid <= 0) grid = vs_device_sms();
  if (!out || cap < 14 || !vs::slab_plan(N, K, pair, grid, vs::slab_default_packed(), &p)) return 0;
  const int32_t v[14] = {p.pair, p.packed, p.E, p.n_tiles, p.passes, p.rows_alloc, p.blk_bytes, p.blk_smem, p.nk, p.kc,
                         p.x_off, p.e_per, p.e_rem, p.grid};
  for (int i = 0; i < 14; ++i) out[i] = v[i];
  return 14;
}

int vs_slab_pack(const void* W, int64_t ldw, int N, int K, int epi, int grid, void* dst, void* stream) {
  vs::SlabPlan p;
  const int pair = epi == vs::EPI_SWIGLU || epi == vs::EPI_HEADS;
  if (grid <= 0) grid = vs_device_sms();
  if (!W || !dst || ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(dst) & 127))
    return vs::prog_fail(VS_ERR_INVALID, "vs_slab_pack", "W must be 16-byte aligned with ldw % 8 == 0, dst 128-byte aligned");
  if (!vs::slab_plan(N, K, pair, grid, vs::slab_default_packed(), &p))
    return vs::prog_fail(VS_ERR_UNSUPPORTED, "vs_slab_pack", "shape not supported by the slab layout");
  dim3 g(p.nk, grid * p.passes);
  vs::slab_pack_kernel<<<g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const bf16*>(W), ldw, N, p,
                                                                            reinterpret_cast<uint8_t*>(dst));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return vs::prog_fail(VS_ERR_CUDA, "vs_slab_pack", cudaGetErrorString(e));
  return VS_OK;
}

}  // extern "C"
