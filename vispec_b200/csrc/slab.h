// "Slab" weight layout of the persistent decode kernel (round 2): every CTA owns COMPLETE output rows.
//
// Round 1 cut the (128-row tile, k-chunk) sequence of a weight matrix into one range per CTA (stream-K): every SM streamed
// the same number of bytes, but a tile's k-range was spread over several CTAs, so every GEMM ended in a global-memory
// exchange of fp32 partial tiles + an epilogue phase that re-read them (157 of them per cycle, 1.44 ms of a 5.1 ms cycle,
// profiles/r1h_phase_trace_cycle.json).  Here the N output rows of a matrix are dealt over the grid instead: CTA b owns
// ~N/148 whole rows ("unit": 28 of o_proj's 4096, 2 x 75 gate|up rows, 2 x 42 q|k|v rows ...), streams their full K
// range through tcgen05 (weights = the M side of the MMA with only `rows` of its 128 lanes meaningful, the <= 32 token rows
// = N) and finishes them from TMEM: bias / residual / SiLU*up / RoPE + KV append run on the accumulators, nothing is
// reduced across CTAs.  The price is that every CTA reads the whole activation matrix (32 x K bf16, L2-resident) instead
// of its k-range: +55 % L2->SM traffic, which the L2 sustains (tools/l2bw_ubench.cu: 13 TB/s of mixed traffic).
//
// HBM image of a matrix W[N][K] for a grid of G CTAs  (vs_slab_bytes / vs_slab_pack build it from the row-major copy):
//   element : one output row -- or, for the pairing epilogues (SiLU(gate)*up, rotate-half RoPE), one PAIR of rows
//             (A, B = A + 64) of a 128-row group: A = (e / 64) * 128 + e % 64   (gate | up feature e; head e/64, dim e%64)
//   unit    : the elements of one CTA in one pass; units = G * passes, unit u owns elements
//             [u * e_per + min(u, e_rem), + e_per + (u < e_rem))            (balanced to within one element)
//   tile    : the rows of a unit that share one accumulator: plain epilogues 1 or 2 tiles of consecutive elements, pairing
//             epilogues exactly 2 (the A rows | the B rows), so lane i of both accumulators belongs to element i
//   block   : (k-chunk j, tile t) = rows_alloc rows x 64 columns, 128 B per row, stored in the order [j][t] -- i.e. a unit
//             is one contiguous stream in exactly the order the MMA consumes it
//   stage   : kc consecutive k-chunks = kc * n_tiles blocks = ONE bulk copy into a 32 KB ring slot.  Blocks are packed
//             back to back (rows_alloc need not be a multiple of 8), so a block may start in the middle of a 1024-byte
//             swizzle atom: the 16-byte chunk c of the row at byte offset `o` of the stage is stored at chunk position
//             c ^ ((o >> 7) & 7) -- the 128-byte swizzle is a function of the shared-memory ADDRESS (the slot base is
//             1024-byte aligned; descriptor base offset 0).  `packed` = 0 selects the conservative variant (blocks padded
//             to whole atoms in shared memory, one bulk copy per block: TMA-issue-bound, measured 130 vs 186 tokens/s),
//             `packed` = 2 the other reading of the ISA's base-offset field (pattern relative to the block).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace vs {

constexpr int SLAB_SLOT_BYTES = 32768;   // ring slot of the program kernel
constexpr int SLAB_X_BYTES = 4096;       // [32 tokens][64] bf16 activation block of one k-chunk
constexpr int SLAB_MAX_ROWS = 112;       // rows per tile: 2 x 112 x 128 B + 4 KB = one slot

struct SlabPlan {
  int pair;          // 1: elements are row pairs (SWIGLU / HEADS)
  int packed;        // 0: blocks padded to whole 1024-byte atoms in shared memory, one bulk copy per block;
                     // 1: blocks back to back (one bulk copy per stage), image swizzled by shared-memory address;
                     // 2: back to back, image swizzled relative to each block + descriptor base offset
  int m64;           // 1: rows_alloc <= 64 -> MMAs with M = 64 (half the shared-memory reads of the weight operand)
  int E;             // elements
  int n_tiles;       // accumulators per pass
  int passes;        // units per CTA
  int rows_alloc;    // rows per block in HBM
  int blk_bytes;     // rows_alloc * 128
  int blk_smem;      // block stride in shared memory: blk_bytes (packed) or rounded up to 1024
  int nk, kc;        // K / 64 chunks; chunks per ring stage
  int x_off;         // offset of the activation blocks inside a slot (1024-aligned, behind the kc * n_tiles weight blocks)
  int e_per, e_rem;  // balanced partition of E over grid * passes units
  int grid;
  long unit_stride;  // bytes between units
  long total_bytes;
};

// false: the shape cannot use the slab path (the caller falls back to the tiled / row-major stream-K path)
bool slab_plan(int N, int K, int pair, int grid, int packed, SlabPlan* out);
int slab_default_packed();               // VS_SLAB_PACKED (default 1)

}  // namespace vs
