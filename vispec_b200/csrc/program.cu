// The persistent decode kernel: ONE launch of one CTA per SM executes a whole "program" -- a list of phases covering
// e.g. a complete draft+verify cycle (32 target layers + LM head + accept + KV splice + 4 draft forwards + tree build)
// -- instead of ~320 separate launches.
//
// Why: decode with <= 32 token rows is a pure weight stream (13.2 GB per verify step at 7B).  Separate
// launches stop the HBM stream at every kernel boundary (ramp-up, TMEM allocation, barrier init, tail) and ncu showed the
// split-K "last arriver" epilogue turning single CTAs into stragglers (profiles/README.md, r1a).  Here
//   * warp 0 (one lane) is the WEIGHT producer: it walks every GEMM phase of the program in order and keeps a
//     6-stage / 96 KB ring full (cp.async.bulk of contiguous 16 KB blocks of the tiled weight copy, or TMA tensor loads of
//     row-major weights).  Weights never depend on activations, so it runs ahead across phase boundaries: the HBM stream
//     continues while the rest of the CTA is in an epilogue, an attention phase or a grid barrier.  (Deeper rings and an
//     L2 prefetcher on warp 3 were measured and lost in round 1; the prefetcher is gone);
//   * warp 2 (one lane) is the ACTIVATION producer: the [32 x 64] token tile of each stage is loaded once the phase
//     that produces it has completed on every CTA (it polls the grid-barrier counter);
//   * warp 1 (one lane) issues tcgen05.mma (M=128 weight rows, N=32 tokens, K=16; fp32 accumulators double-buffered in
//     TMEM) and commits stages back to the producers;
//   * warps 4-19 are the workers (four teams of 128 threads, or two attention teams of 256): they drain TMEM into fp32
//     partial tiles, and run every non-GEMM phase (reduction + fused epilogues with the following RMSNorm folded in,
//     attention, arg-max, log-softmax/top-k, tree build / accept / KV splice) on team-sized work items dealt CTA-first
//     over all teams of the grid.
// GEMM work is split "stream-K" style: the (tile, k-chunk) sequence of a weight matrix is cut into one contiguous,
// equally sized range per CTA, so every SM streams the same number of bytes; tiles shared by several CTAs are reduced
// in k order by the following PH_EPI phase, which waits on per-tile arrival counters instead of a grid barrier
// (deterministic; no atomics on data, no last-arriver serialisation).
// Other phases are separated by a grid barrier (one release-add + relaxed spin per CTA).
#include "program.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/vispec_b200.h"
#include "attn_dev.cuh"
#include "attn_ring.cuh"
#include "profile.h"
#include "program_dev.cuh"
#include "slab.h"
#include "sm100.cuh"

#ifndef VS_MMA_CONVERGED
#define VS_MMA_CONVERGED 0
#endif
// The two-phase attention program (PH_ATTN_S + PH_ATTN_PV on the cp.async bodies of attn_dev.cuh) was the A/B reference of
// the single-phase attention; it is compiled into the kernel only with -DVS_TWO_PHASE_ATTN=1 (tools/build_variant.py).
#ifndef VS_TWO_PHASE_ATTN
#define VS_TWO_PHASE_ATTN 0
#endif
// Register split between the warpgroups of the program kernel (setmaxnreg): 640 threads cap the kernel at 96 registers
// per thread; the role warps (one warpgroup) give 32 of theirs back and every worker thread takes 8.  ptxas then allocates
// the worker code on 104 registers (LDL / STL of the kernel 474 -> 316) -- same-box A/B 226.3 -> 230.8 tok/s; 56 / 104 and
// 48 / 104 measure the same within noise; 32 / 112 spills inside the producer loops: attention 491 -> 478 us per cycle but
// 230.6 -> 228.6 tok/s overall.
#ifndef VS_SETMAXNREG
#define VS_SETMAXNREG 1
#endif
#ifndef VS_ROLE_REGS
#define VS_ROLE_REGS 64
#endif
#ifndef VS_WORKER_REGS
#define VS_WORKER_REGS 104
#endif
static_assert(128 * (96 - VS_ROLE_REGS) >= 512 * (VS_WORKER_REGS - 96), "the role warpgroup must release what the workers take");

namespace vs {

namespace {

// Ring of the role warps (run-ahead buffer of the weight stream), RING_BYTES of shared memory cut at launch into
//   * 6 slots of 20 KB for a stream-K program (PH_GEMM: one 128 x 64 weight chunk + one 32 x 64 activation block per
//     stage) -- the round-1 geometry: bytes in flight, not slot count, set the streaming rate, or
//   * 4 slots of 32 KB when the program holds slab GEMMs (PH_GEMM2, slab.h: kc k-chunks of this CTA's weight rows + the
//     matching activation blocks per stage); stream-K stages then use the first 20 KB of a slot.
constexpr int BM = PG_BM, BK = PG_BK;
constexpr int NS_MAX = 6;
constexpr int RING_BYTES = 4 * SLAB_SLOT_BYTES;  // 131072
constexpr int W_STAGE = BM * BK * 2;             // 16384
constexpr int X_STAGE = VS_TOK * BK * 2;         // 4096
static_assert(NS_MAX * (W_STAGE + X_STAGE) <= RING_BYTES, "stream-K ring larger than the ring region");
static_assert(W_STAGE + X_STAGE <= SLAB_SLOT_BYTES, "stream-K stage larger than a slab slot");
constexpr int BIG_SMEM = 44032;                  // scratch of one attention team (AttnSmem); both together also stage the
                                                 // tree ops and the top-k merge on CTA 0
constexpr int TEAM_SMEM = 1024;                  // reduction scratch of one small team
constexpr int TEAM_OFF = RING_BYTES + PG_BIG_TEAMS * BIG_SMEM;
constexpr int PHASE_WORDS = (sizeof(Phase) + 3) / 4;
constexpr int PHASE_SLOT = (sizeof(Phase) + 127) / 128 * 128;
constexpr int PHBUF_OFF = TEAM_OFF + PG_TEAMS * TEAM_SMEM;             // two phase descriptors (double buffer)
constexpr int BAR_OFF = PHBUF_OFF + 2 * PHASE_SLOT;
constexpr int SMEM_BYTES = BAR_OFF + 512 + 1024;
static_assert(PHASE_WORDS <= PG_WORKERS, "phase descriptor larger than one word per worker thread");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int TMEM_COLS = 128;               // two accumulator sets x two tiles x 32 token columns
static_assert(sizeof(AttnSmem) <= BIG_SMEM, "team scratch too small for attention");
static_assert(sizeof(AttnRingSmem) <= BIG_SMEM, "team scratch too small for the TMA-staged attention");
static_assert(BIG_SMEM % 1024 == 0 && RING_BYTES % 1024 == 0, "granules are swizzled by shared-memory address");
constexpr int ATT_BAR_OFF = BAR_OFF + 256;       // full[4] | empty[4] mbarriers of the granule FIFO, per attention team
static_assert(sizeof(TreeSmem) <= BIG_SMEM, "team scratch too small for the tree ops");
static_assert(sizeof(SampleSmem) <= BIG_SMEM, "team scratch too small for the sampling accept");
static_assert(sizeof(ScanSmem) <= BIG_SMEM, "team scratch too small for the vocabulary scans");

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
// NOTE: the phase descriptor lives in shared memory; the bodies below read its fields many times between barriers and
// inline asm with memory clobbers, which would turn every use into a fresh generic load.  Each phase therefore works on
// a private (register) copy of the arguments it needs.
#if VS_TWO_PHASE_ATTN
__device__ __noinline__ void attn_s_phase(const Phase& ph, int gteam, int n_teams, int tt, uint8_t* tsm, BigTeamSync tsync) {
  const AttnArgs a = ph.attn;
  const int n_ctx = a.n_ctx ? ld_vol(a.n_ctx) : a.n_ctx_host;
  const int n_chunks = (n_ctx + a.n_new + ATT_CHUNK - 1) / ATT_CHUNK;
  const int n_items = a.n_heads * n_chunks;
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(tsm);
  for (int item = gteam; item < n_items; item += n_teams)
    attn_scores_item(a, sm, item / n_chunks, 0, 1, item % n_chunks, true, tt, tsync);
}
__device__ __noinline__ void attn_pv_phase(const Phase& ph, int gteam, int n_teams, int tt, uint8_t* tsm, BigTeamSync tsync) {
  const AttnArgs a = ph.attn;
  const int n_items = a.n_heads * a.pv_splits;
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(tsm);
  for (int item = gteam; item < n_items; item += n_teams)
    attn_pv_item(a, sm, item / a.pv_splits, 0, 1, item % a.pv_splits, a.pv_splits, tt, tsync);
}
#endif
// PH_ATTN_FUSED: one (head, kv split) item per team, K/V granules staged by the KV producer warp (attn_ring.cuh).
// `fifo`: FIFO index of the team's next granule (advances identically on the producer side).  The arguments are read
// from the staged phase descriptor in shared memory (no private copy in local memory).
template <bool STREAM, bool EAGER>
__device__ __noinline__ unsigned int attn_ring_phase(const Phase& ph, int n_ctx, AttnRingSmem& sm, uint64_t* bars,
                                                     unsigned int fifo, unsigned int epoch, int gteam, int tt, BigTeamSync tsync) {
  const int sp = ph.attn.pv_splits;
#ifdef VS_PROGRAM_CLK
  if (blockIdx.x == 0 && tt == 0 && gteam == 0) g_clk[58] = clock64();
#endif
  return fifo + attn_ring_item<STREAM, EAGER>(ph.attn, n_ctx, sm, bars, fifo, gteam / sp, gteam % sp, sp, ph.tile_cnt, epoch, tt, tsync);
}
// kv length of a prefix_stable phase: the same device int for every layer of a verify step and unchanged since the launch
// started, so one L2 round trip serves all of them
struct KvLenCache { const int* ptr; int val; };
__device__ __forceinline__ unsigned int attn_fused_phase(const Phase& ph, unsigned int epoch, int gteam, int tt, uint8_t* tsm,
                                                         uint64_t* bars, unsigned int fifo, KvLenCache& kvc, BigTeamSync tsync) {
  const AttnArgs& a = ph.attn;
  const int sp = a.pv_splits;
  if (gteam >= a.n_heads * sp) return fifo;       // one resident item per team: the items of a head wait for each other
  int n_ctx = a.n_ctx_host;
  if (a.n_ctx) {
    if (a.prefix_stable && kvc.ptr == a.n_ctx) n_ctx = kvc.val;
    else {
      n_ctx = ld_vol(a.n_ctx);
      if (a.prefix_stable) { kvc.ptr = a.n_ctx; kvc.val = n_ctx; }
    }
  }
  const int kv_len = n_ctx + a.n_new;
#ifdef VS_PROGRAM_CLK
  if (blockIdx.x == 0 && tt == 0 && gteam == 0) g_clk[59] = clock64();
#endif
  if (kv_len > a.score_ld) __trap();              // the score scratch bounds the context
  AttnRingSmem& sm = *reinterpret_cast<AttnRingSmem*>(tsm);
  // up to 2 * ATT_RING_NG granules per split the scores stay in registers; longer contexts park them (uniform over the items)
  const int n_gran = (kv_len + ATT_GRAN - 1) / ATT_GRAN;
  const bool stream = (n_gran + sp - 1) / sp > 2 * ATT_RING_NG;
  if (a.eager) {
    if (stream) return attn_ring_phase<true, true>(ph, n_ctx, sm, bars, fifo, epoch, gteam, tt, tsync);
    return attn_ring_phase<false, true>(ph, n_ctx, sm, bars, fifo, epoch, gteam, tt, tsync);
  }
  if (stream) return attn_ring_phase<true, false>(ph, n_ctx, sm, bars, fifo, epoch, gteam, tt, tsync);
  return attn_ring_phase<false, false>(ph, n_ctx, sm, bars, fifo, epoch, gteam, tt, tsync);
}
// One compact function per tree operation (each runs once per cycle on CTA 0, cold: what is fetched should be what is
// executed -- as one function with every body inlined this was 91 KB of code).
__device__ __noinline__ void tree_op_level0(const Phase& ph, int wtid) { tree_level0_body(ph.tb, ph.state, wtid); }
__device__ __noinline__ void tree_op_expand(const Phase& ph, int wtid, TreeSmem& sm) {
  tree_expand_body(ph.tb, ph.state, ph.level, sm, wtid, WorkerSync());
}
__device__ __noinline__ void tree_op_build(const Phase& ph, int wtid, TreeSmem& sm) {
  tree_build_body(ph.tb, ph.state, ph.sample_token_dev, ph.sample_token_host, sm, wtid, PG_WORKERS, WorkerSync());
}
__device__ __noinline__ void tree_op_accept(const Phase& ph, int wtid, TreeSmem& sm) {
  tree_accept_body(ph.tb, ph.state, ph.amax, ph.out_rec, sm, wtid, WorkerSync());
}
__device__ __noinline__ void tree_op_accept_sample(const Phase& ph, int wtid, uint8_t* team_base) {
  tree_accept_sample_body(ph.tb, ph.state, ph.logits, ph.ldl, ph.V, ph.lstats, ph.n_slices, ph.inv_temp, ph.uni_override,
                          ph.out_rec, *reinterpret_cast<SampleSmem*>(team_base), wtid, PG_WORKERS, WorkerSync(), ph.top_k,
                          ph.hist);
}
__device__ __noinline__ void tree_op_merge_argmax(const Phase& ph, int wtid) { merge_argmax(ph, wtid); }
__device__ __noinline__ void tree_op_merge_topk(const Phase& ph, int wtid, uint8_t* team_base) {
  merge_topk(ph, wtid, reinterpret_cast<unsigned long long*>(team_base), PG_BIG_TEAMS * BIG_SMEM / 8, WorkerSync());
}
__device__ __forceinline__ void tree_phase(const Phase& ph, int wtid, uint8_t* team_base) {
  // the tree bodies read device state with plain loads: drop whatever this SM's L1 still holds from an earlier phase of
  // this launch (gpu-scope fence = CCTL.IVALL) before anyone reads
  if (wtid == 0) __threadfence();
  named_bar_sync(1, PG_WORKERS);
  TreeSmem& sm = *reinterpret_cast<TreeSmem*>(team_base);
  if (ph.keys && (ph.tree_op == TREE_ACCEPT || ph.tree_op == TREE_MERGE_ARGMAX)) {
    tree_op_merge_argmax(ph, wtid);
    named_bar_sync(1, PG_WORKERS);
  }
  if (ph.keys && (ph.tree_op == TREE_LEVEL0 || ph.tree_op == TREE_EXPAND || ph.tree_op == TREE_MERGE_TOPK))
    tree_op_merge_topk(ph, wtid, team_base);
  switch (ph.tree_op) {
    case TREE_LEVEL0: tree_op_level0(ph, wtid); break;
    case TREE_EXPAND: tree_op_expand(ph, wtid, sm); break;
    case TREE_BUILD: tree_op_build(ph, wtid, sm); break;
    case TREE_ACCEPT: tree_op_accept(ph, wtid, sm); break;
    case TREE_COMMIT: tree_commit_body(ph.state, wtid); break;
    case TREE_ACCEPT_SAMPLE: tree_op_accept_sample(ph, wtid, team_base); break;
    default: break;
  }
}

// Every CTA of a slab GEMM needs ALL activation blocks of the phase.  Walking the k-chunks in the same order on every CTA
// makes all 148 SMs ask the L2 for the same 4-16 KB at the same moment, and requests for one line serialise in its L2
// slice: measured 8-11 GB/s per SM for the activation blocks, which made every GEMM phase activation-bound (35-65 us
// against 5-28 us of weight streaming, profiles/r2c_trace.txt).  So CTA b starts its k loop at stage b * spp / G and
// wraps around: at any time the CTAs are spread evenly over the stages.  (The k order of the fp32 accumulation of a row
// therefore depends on the CTA that owns it -- fixed per row and launch-invariant, like every other order in here.)
__device__ __forceinline__ int slab_rotation(int b, int G, int spp) {
#ifdef VS_SLAB_NO_ROTATION
  return 0;
#else
  return (int)(((long)b * spp) / G);
#endif
}
__device__ __forceinline__ int slab_stage(int s, int rot, int spp) {
  const int v = s + rot;
  return v >= spp ? v - spp : v;
}

__global__ void __launch_bounds__(PG_THREADS, 1) program_kernel(const ProgramHeader hdr) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by POINTER arithmetic on the shared array (not through an integer cast): the compiler keeps the
  // address space, so everything below is LDS / STS on 32-bit addresses instead of generic loads on 64-bit pointers
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* team_base = smem + RING_BYTES;
  const int NS = hdr.ns, STAGE_BYTES = hdr.stage_bytes;      // ring geometry of this program (see RING_BYTES)
  uint64_t* full_w = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full_x = full_w + NS_MAX;
  uint64_t* empty = full_x + NS_MAX;
  uint64_t* tfull = empty + NS_MAX;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x, b = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS_MAX; ++i) {
      mbar_init(&full_w[i], 2);     // ONE "stage full" barrier: the weight and the activation producer both arrive on it with
      mbar_init(&full_x[i], 1);     // their byte counts (the MMA warp waits once per stage instead of twice); full_x is unused
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], PG_WORKERS / 32);
    }
    uint64_t* att_bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR_OFF);
    for (int t = 0; t < PG_BIG_TEAMS; ++t)
      for (int i = 0; i < ATT_SLOTS; ++i) {
        mbar_init(&att_bars[t * 2 * ATT_SLOTS + i], 1);                      // full: the producer's expect_tx arrival
        mbar_init(&att_bars[t * 2 * ATT_SLOTS + ATT_SLOTS + i], ATT_WARPS);       // empty: 8 arrivals (attn_ring.cuh)
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // register split between the warpgroups (VS_SETMAXNREG above): 128 * (96 - VS_ROLE_REGS) >= 512 * (VS_WORKER_REGS - 96)
  if (warp < 4) {
#if VS_SETMAXNREG
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VS_ROLE_REGS));
#endif
  if (warp == 0) {
    // ===================== weight producer =====================
    // Walks every GEMM phase of the program in order.  Weights never depend on activations, so this thread runs ahead of
    // the rest of the CTA across phase boundaries: the ring is the run-ahead buffer.
    if (lane == 0) {
      int stage = 0, par = 0;
      for (int p = 0; p < hdr.n_phases; ++p) {
        const Phase& ph = hdr.phases[p];
        if (ph.type == PH_GEMM2) {
          const int passes = ph.s_passes, kc = ph.s_kc, nk = ph.nk, nt = ph.s_ntiles;
          const uint32_t blk = (uint32_t)ph.s_blk_bytes, blk_smem = (uint32_t)ph.s_blk_smem;
          const uint8_t* wbase = reinterpret_cast<const uint8_t*>(ph.Wt);
          const int spp = (nk + kc - 1) / kc, rot = slab_rotation(b, G, spp);
          for (int ps = 0; ps < passes; ++ps) {
            const int u = b * passes + ps;
            if (ph.s_eper + (u < ph.s_erem ? 1 : 0) == 0) continue;
            const uint8_t* ubase = wbase + (long)u * ph.s_unit_stride;
            for (int s = 0; s < spp; ++s) {
              const int j0 = slab_stage(s, rot, spp) * kc;          // (rotated k order, see slab_stage)
              const uint32_t nblk = (uint32_t)(min(kc, nk - j0) * nt);
              const uint8_t* src = ubase + (long)j0 * nt * blk;
              mbar_wait_sleep(&empty[stage], par ^ 1, hdr.spin_ns);
              if (hdr.debug & 2) {
                mbar_arrive(&full_w[stage]);
                if (++stage == NS) { stage = 0; par ^= 1; }
                continue;
              }
              mbar_expect_tx(&full_w[stage], nblk * blk);
              uint8_t* dst = smem + stage * STAGE_BYTES;
              if (blk == blk_smem) {
                bulk_load(dst, src, nblk * blk, &full_w[stage], HINT_EVICT_FIRST);     // the whole stage: one copy
              } else {
                for (uint32_t i = 0; i < nblk; ++i)
                  bulk_load(dst + i * blk_smem, src + i * blk, blk, &full_w[stage], HINT_EVICT_FIRST);
              }
              if (++stage == NS) { stage = 0; par ^= 1; }
            }
          }
        } else if (ph.type == PH_GEMM) {
          const int2 r = *reinterpret_cast<const int2*>(hdr.tables + ph.range_off + 2 * b);
          const int nk = ph.nk;
          const CUtensorMap* tm = hdr.maps + ph.tmW;
          const bf16* wt = ph.Wt;                            // tiled layout: chunk c is the 16 KB block c of the matrix
          int tile = r.x / nk, kc = r.x % nk;
          for (int c = r.x; c < r.y; ++c) {
            mbar_wait_sleep(&empty[stage], par ^ 1, hdr.spin_ns);
            mbar_expect_tx(&full_w[stage], W_STAGE);
            if (wt) bulk_load(smem + stage * STAGE_BYTES, wt + (long)c * (BM * BK), W_STAGE, &full_w[stage], HINT_EVICT_FIRST);
            else tma_load_2d(tm, &full_w[stage], smem + stage * STAGE_BYTES, kc * BK, tile * BM, HINT_EVICT_FIRST);
            if (++stage == NS) { stage = 0; par ^= 1; }
            if (++kc == nk) { kc = 0; ++tile; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // VS_MMA_CONVERGED=1 (build macro, A/B): the whole warp walks the program and one ELECTED lane issues the MMAs of a
    // stage back to back.  tcgen05.mma takes its operands from uniform registers; inside a single-lane branch the
    // compiler cannot prove them warp-uniform and wraps every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop
    // (~85 ns per MMA measured, tools/bench_slab_phases.py), which caps the stream at 16 KB per 4 x 85 ns = 7.1 TB/s.
#if VS_MMA_CONVERGED
#define MMA_ENTER if (true)
#define MMA_WAIT(call) do { if (lane == 0) { call; } __syncwarp(); } while (0)
#define MMA_ISSUE if (elect_one())
#define MMA_SYNC() __syncwarp()
#define MMA_UNI(v) uni(v)
#else
#define MMA_ENTER if (lane == 0)
#define MMA_WAIT(call) do { call; } while (0)
#define MMA_ISSUE
#define MMA_SYNC() do { } while (0)
#define MMA_UNI(v) (v)
#endif
    MMA_ENTER {
      int stage = 0, par = 0, acc = 0, acc_par = 0;
      const unsigned int spin = hdr.spin_ns;
      for (int p = 0; p < hdr.n_phases; ++p) {
        const Phase& ph = hdr.phases[p];
        const int type = MMA_UNI(ph.type);
        if (type == PH_GEMM2) {
          const int passes = MMA_UNI(ph.s_passes), kc = MMA_UNI(ph.s_kc), nk = MMA_UNI(ph.nk), nt = MMA_UNI(ph.s_ntiles);
          const uint32_t blk_smem = (uint32_t)MMA_UNI(ph.s_blk_smem), x_off = (uint32_t)MMA_UNI(ph.s_xoff);
          const int e_per = MMA_UNI(ph.s_eper), e_rem = MMA_UNI(ph.s_erem);
          const uint32_t idesc = MMA_UNI(ph.s_m64) ? IDESC_M64 : IDESC;
          const bool rel = MMA_UNI(ph.s_packed) == 2;
          const int spp = (nk + kc - 1) / kc, rot = slab_rotation(b, G, spp);
          for (int ps = 0; ps < passes; ++ps) {
            const int u = b * passes + ps;
            if (e_per + (u < e_rem ? 1 : 0) == 0) continue;
            MMA_WAIT(mbar_wait_sleep(&tempty[acc], acc_par ^ 1, spin));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 64;
            for (int s = 0; s < spp; ++s) {
              const int j0 = slab_stage(s, rot, spp) * kc;
              const int nj = min(kc, nk - j0);
              MMA_WAIT(mbar_wait_full(&full_w[stage], par, spin));
              tc_fence_after();
              const uint32_t base = smem_u32(smem) + stage * STAGE_BYTES;
              MMA_ISSUE {
                if (!(hdr.debug & 4)) {
                  for (int jj = 0; jj < nj; ++jj) {
                    const uint64_t db = make_desc_u32(base + x_off + jj * X_STAGE);
                    for (int t = 0; t < nt; ++t) {
                      const uint32_t blk = base + (uint32_t)(jj * nt + t) * blk_smem;
                      const uint64_t da = rel ? make_desc_rel_u32(blk) : make_desc_u32(blk);
#pragma unroll
                      for (int k = 0; k < BK / 16; ++k)
                        tc_mma(d_tmem + t * 32, da + 2 * k, db + 2 * k, (s | jj | k) ? 1u : 0u, idesc);
                    }
                  }
                }
                tc_commit(&empty[stage]);
              }
              MMA_SYNC();
              if (++stage == NS) { stage = 0; par ^= 1; }
            }
            MMA_ISSUE tc_commit(&tfull[acc]);
            MMA_SYNC();
            acc ^= 1;
            if (acc == 0) acc_par ^= 1;
          }
        } else if (type == PH_GEMM) {
          const int2 r = *reinterpret_cast<const int2*>(hdr.tables + MMA_UNI(ph.range_off) + 2 * b);
          const int lo = MMA_UNI(r.x), hi = MMA_UNI(r.y);
          const int nk = MMA_UNI(ph.nk);
          int c = lo;
          int tile = c / nk;
          while (c < hi) {
            const int seg_end = min(hi, (tile + 1) * nk);
            MMA_WAIT(mbar_wait_sleep(&tempty[acc], acc_par ^ 1, spin));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 64;
            bool first = true;
            for (; c < seg_end; ++c) {
              MMA_WAIT(mbar_wait_full(&full_w[stage], par, spin));
              tc_fence_after();
              const uint32_t base = smem_u32(smem) + stage * STAGE_BYTES;
              const uint64_t da = make_desc_u32(base);
              const uint64_t db = make_desc_u32(base + W_STAGE);
              MMA_ISSUE {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) tc_mma(d_tmem, da + 2 * k, db + 2 * k, (!first || k > 0) ? 1u : 0u);
                tc_commit(&empty[stage]);
              }
              MMA_SYNC();
              first = false;
              if (++stage == NS) { stage = 0; par ^= 1; }
            }
            MMA_ISSUE tc_commit(&tfull[acc]);
            MMA_SYNC();
            acc ^= 1;
            if (acc == 0) acc_par ^= 1;
            ++tile;
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== activation producer =====================
    if (lane == 0) {
      int stage = 0, par = 0;
      for (int p = 0; p < hdr.n_phases; ++p) {
        const Phase& ph = hdr.phases[p];
        if (ph.type != PH_GEMM2 && ph.type != PH_GEMM) continue;
        int n_stages;                                        // stages of this CTA in the phase
        int2 r = make_int2(0, 0);
        if (ph.type == PH_GEMM2) {
          int units = 0;
          for (int ps = 0; ps < ph.s_passes; ++ps) {
            const int u = b * ph.s_passes + ps;
            units += ph.s_eper + (u < ph.s_erem ? 1 : 0) > 0 ? 1 : 0;
          }
          n_stages = units * ((ph.nk + ph.s_kc - 1) / ph.s_kc);
        } else {
          r = *reinterpret_cast<const int2*>(hdr.tables + ph.range_off + 2 * b);
          n_stages = r.y - r.x;
        }
        if (n_stages <= 0) continue;
        // the token tile is produced by earlier phases of this launch: wait until all of them completed everywhere
        // (relaxed polling, no L1 invalidation storm; the tile itself is read by the TMA unit from L2, ordered behind the
        // observed counter value by the proxy fence)
        const unsigned int need = (unsigned int)ph.bars_before * (unsigned int)G;
        while (ld_relaxed_u32(hdr.counters) < need) __nanosleep(64);
        fence_proxy_async();
        if (ph.type == PH_GEMM2) {
          const int kc = ph.s_kc, nk = ph.nk;
          const uint32_t x_off = (uint32_t)ph.s_xoff;
          const CUtensorMap* tm3 = ph.tmX3 >= 0 ? hdr.maps + ph.tmX3 : nullptr;
          const CUtensorMap* tm = hdr.maps + ph.tmX;
          const int spp = (nk + kc - 1) / kc;                // stages per pass: every pass walks the same k-chunks
          const int rot = slab_rotation(b, G, spp);
          for (int s = 0; s < n_stages; ++s) {
            const int j0 = slab_stage(s % spp, rot, spp) * kc;
            const int nj = min(kc, nk - j0);
            mbar_wait_sleep(&empty[stage], par ^ 1, hdr.spin_ns);
            if (hdr.debug & 1) {
              mbar_arrive(&full_w[stage]);
              if (++stage == NS) { stage = 0; par ^= 1; }
              continue;
            }
            mbar_expect_tx(&full_w[stage], nj * X_STAGE);
            uint8_t* dst = smem + stage * STAGE_BYTES + x_off;
            if (tm3 && nj == kc) {
              tma_load_3d(tm3, &full_w[stage], dst, 0, 0, j0, HINT_EVICT_LAST);      // {64, 32, kc} box: kc blocks at once
            } else {
              for (int jj = 0; jj < nj; ++jj)
                tma_load_2d(tm, &full_w[stage], dst + jj * X_STAGE, (j0 + jj) * BK, 0, HINT_EVICT_LAST);
            }
            if (++stage == NS) { stage = 0; par ^= 1; }
          }
        } else {
          const CUtensorMap* tm = hdr.maps + ph.tmX;
          int kc = r.x % ph.nk;
          for (int c = r.x; c < r.y; ++c) {
            mbar_wait_sleep(&empty[stage], par ^ 1, hdr.spin_ns);
            mbar_expect_tx(&full_w[stage], X_STAGE);
            tma_load_2d(tm, &full_w[stage], smem + stage * STAGE_BYTES + W_STAGE, kc * BK, 0, HINT_EVICT_LAST);
            if (++stage == NS) { stage = 0; par ^= 1; }
            if (++kc == ph.nk) kc = 0;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== KV producer =====================
    // Walks the attention phases of the program in order and keeps the granule FIFO of both attention teams full
    // (attn_ring.cuh).  The fully visible prefix of a phase flagged prefix_stable does not depend on this launch, so its K
    // granules are staged as soon as the team scratch is free (ph.pf_bars grid barriers), long before the phase starts.
    if (lane == 0) {
      unsigned int fifo[2] = {0u, 0u};
      uint64_t* att_bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR_OFF);
      for (int p = 0; p < hdr.n_phases; ++p) {
        const Phase& ph = hdr.phases[p];
        if (ph.type != PH_ATTN_FUSED) continue;
        const unsigned int need_start = (unsigned int)ph.pf_bars * (unsigned int)G;
        const unsigned int need_tail = (unsigned int)ph.bars_before * (unsigned int)G;
        while (ld_relaxed_u32(hdr.counters) < need_start) __nanosleep(100);
#ifdef VS_PROGRAM_CLK
        if (b == 0) g_clk[20] = clock64();
#endif
        const bool tail_ok = need_start >= need_tail;
        if (tail_ok) fence_proxy_async();
        const AttnArgs a = ph.attn;
        const int n_ctx = a.n_ctx ? __ldcg(a.n_ctx) : a.n_ctx_host;
#ifdef VS_PROGRAM_CLK
        if (b == 0) g_clk[21] = clock64() + (n_ctx & 0);
#endif
        att_ring_produce(a, hdr.maps + ph.tmK, hdr.maps + ph.tmV, ph.kv_two_loads, n_ctx, team_base, BIG_SMEM, att_bars, fifo,
                         hdr.counters, need_tail, tail_ok, b, G, hdr.attn_pf);
#ifdef VS_PROGRAM_CLK
        if (b == 0) g_clk[23] = clock64();
#endif
      }
    }
  }
  } else {
#if VS_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(VS_WORKER_REGS));
#endif
    // ===================== workers (16 warps: four teams of 128 threads / two attention teams of 256) =====================
    const int wtid = threadIdx.x - 4 * 32;          // 0..511
    const int team = wtid >> 7, tt = wtid & 127;    // small team
    const int bteam = wtid >> 8, btt = wtid & 255;  // attention team
    const int wwarp = wtid >> 5;                    // 0..15
    const TeamSync tsync{2 + team};
    const BigTeamSync bsync{2 + PG_TEAMS + bteam};
    uint8_t* tsm = smem + TEAM_OFF + team * TEAM_SMEM;
    uint8_t* bsm = team_base + bteam * BIG_SMEM;
    // work items are dealt CTA-first: item i runs on CTA i % G, so a phase with few items spreads over all SMs
    const int gteam = team * G + b, n_teams = PG_TEAMS * G;
    const int gbteam = bteam * G + b, n_bteams = PG_BIG_TEAMS * G;
    int acc = 0, acc_par = 0;
    unsigned int att_fifo = 0;                      // granules of the attention FIFO this team consumed so far
    KvLenCache kvc{nullptr, 0};
    uint64_t* att_bars = reinterpret_cast<uint64_t*>(smem + ATT_BAR_OFF) + bteam * 2 * ATT_SLOTS;
    const unsigned int epoch = ld_relaxed_u32(hdr.counters + 2);     // launches completed before this one
    // Phase descriptors are staged through shared memory: the words of phase p+1 are copied (cp.async) while phase p
    // runs and waited for just before the barrier, so no phase starts with a chain of L2 round trips for its own arguments.
    uint8_t* phbuf = smem + PHBUF_OFF;
    {
      const unsigned int* src = reinterpret_cast<const unsigned int*>(hdr.phases);
      unsigned int* dst = reinterpret_cast<unsigned int*>(phbuf);
      for (int i = wtid; i < PHASE_WORDS; i += PG_WORKERS) dst[i] = src[i];
      named_bar_sync(1, PG_WORKERS);
    }
    int cur = 0;                                    // slot of the staged descriptor of phase p
    for (int p = 0; p < hdr.n_phases;) {
      const Phase& ph = *reinterpret_cast<const Phase*>(phbuf + cur * PHASE_SLOT);
#ifdef VS_PROGRAM_CLK
      if (b == 0 && wtid == 0 && ph.type == PH_ATTN_FUSED) g_clk[56] = clock64();
#endif
      // A GEMM whose epilogue waits on per-tile counters (no grid barrier in between) runs together with that epilogue as
      // ONE body: its descriptor carries the epilogue's fields as well (vs_program_commit), the PH_EPI descriptor that
      // follows is skipped: no descriptor switch, no CTA-wide barrier and no dispatch between the two.  (Loading the
      // epilogue's operands early, while the GEMM streams, was measured: 206 -> 194 tok/s -- the extra loads at the start of
      // the GEMM phase delay the first accumulator drain.)
      const bool pair = ph.type == PH_GEMM && ph.no_barrier && p + 1 < hdr.n_phases;
      const int step = pair ? 2 : 1;
      // the next descriptor goes global -> shared as an asynchronous copy: held in a register across the body it was
      // spilled, and the spill store waited for the load (an L2 round trip at the top of every phase)
      const bool has_next = p + step < hdr.n_phases;
      if (has_next && wtid < PHASE_WORDS)
        cp_async4(phbuf + (cur ^ 1) * PHASE_SLOT + 4 * wtid, reinterpret_cast<const unsigned int*>(hdr.phases + p + step) + wtid);
#ifdef VS_PROGRAM_CLK
      if (b == 0 && wtid == 0 && ph.type == PH_ATTN_FUSED) g_clk[57] = clock64();
#endif
      switch (ph.type) {
        case PH_GEMM: {
          const int2 rg = *reinterpret_cast<const int2*>(hdr.tables + ph.range_off + 2 * b);
          const int lo = rg.x, hi = rg.y;
          if (lo < hi) {
            const int tile0 = lo / ph.nk, tile1 = (hi - 1) / ph.nk;
            const int quarter = wwarp & 3, cg = wwarp >> 2;       // TMEM lane quarter, group of 8 token columns
            for (int tile = tile0; tile <= tile1; ++tile) {
              mbar_wait(&tfull[acc], acc_par);
              tc_fence_after();
              float v[8];
              tmem_ld8(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 64 + cg * 8, v);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
              acc ^= 1;
              if (acc == 0) acc_par ^= 1;
              float* dst = ph.partials + ((long)b * ph.maxseg + (tile - tile0)) * TILE_FLOATS + (cg * 8) * BM +
                           quarter * 32 + lane;
#pragma unroll
              for (int i = 0; i < 8; ++i) dst[i * BM] = v[i];
              if (ph.no_barrier) {       // publish this contribution: the tile's epilogue items poll the counter
                named_bar_sync(1, PG_WORKERS);
                if (wtid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ph.tile_cnt + tile) : "memory");
#ifdef VS_PROGRAM_CLK
                if (b == 0 && wtid == 0) g_clk[48 + min(tile - tile0, 3)] = clock64();     // CTA 0 published its tile
#endif
              }
            }
          }
          if (pair) {
            if (hdr.trace_cta && wtid == 0) hdr.trace_cta[(long)p * G + b] = global_timer();
            if (hdr.trace && b == 0 && wtid == 0) hdr.trace[p] = global_timer();
#ifdef VS_PROGRAM_CLK
            if (b == 0 && wtid == 0) g_clk[32 + (p & 15)] = clock64();      // "end" of the GEMM phase: this CTA's tiles are drained
#endif
            epi_phase(ph, epoch, gteam, n_teams, tt, reinterpret_cast<float*>(tsm), tsync);
          }
        } break;
        case PH_GEMM2:
        {
          const int st = slab_epilogue(ph, epoch, b, G, wtid, tmem_base, acc | (acc_par << 1), tfull, tempty,
                                       *reinterpret_cast<SlabRed*>(team_base), (hdr.debug & 8) != 0);
          acc = st & 1;
          acc_par = st >> 1;
        }
          break;
        case PH_EPI: epi_phase(ph, epoch, gteam, n_teams, tt, reinterpret_cast<float*>(tsm), tsync); break;
        case PH_NORM: norm_phase(ph, gteam, n_teams, tt, reinterpret_cast<float*>(tsm), tsync); break;
        case PH_ROWS: rows_phase(ph, gteam, n_teams, tt); break;
#if VS_TWO_PHASE_ATTN
        case PH_ATTN_S: attn_s_phase(ph, gbteam, n_bteams, btt, bsm, bsync); break;
        case PH_ATTN_PV: attn_pv_phase(ph, gbteam, n_bteams, btt, bsm, bsync); break;
#endif
        case PH_ATTN_FUSED: att_fifo = attn_fused_phase(ph, epoch, gbteam, btt, bsm, att_bars, att_fifo, kvc, bsync); break;
        case PH_ARGMAX: argmax_phase(ph, gbteam, n_bteams, btt, *reinterpret_cast<ScanSmem*>(bsm), bsync); break;
        case PH_LSM_STATS: lsm_stats_phase(ph, gbteam, n_bteams, btt, *reinterpret_cast<ScanSmem*>(bsm), bsync); break;
        case PH_TOPK_HIST: topk_hist_phase(ph, gbteam, n_bteams, btt, *reinterpret_cast<ScanSmem*>(bsm), bsync); break;
        case PH_LSM_CAND: lsm_cand_phase(ph, gbteam, n_bteams, btt, *reinterpret_cast<ScanSmem*>(bsm), bsync); break;
        case PH_TREE:
          if (b == 0) tree_phase(ph, wtid, team_base);
          break;
        case PH_SPLICE: {
          for (long item = gteam; item < ph.n_items; item += n_teams) kv_splice_item(ph.splice, item, tt, tsync);
        } break;
        default: break;
      }
      const int pl = p + step - 1;                  // the last phase this body executed
      if (has_next) cp_async_wait_all();            // the staged descriptor (made visible by the barrier below)
      if (hdr.trace_cta && wtid == 0) hdr.trace_cta[(long)pl * G + b] = global_timer();   // this CTA's items are done
      if (has_next) {
        if (!pair && ph.type == PH_GEMM && ph.no_barrier) named_bar_sync(1, PG_WORKERS);   // only the staged descriptor needs it
        else grid_barrier(hdr.counters, (unsigned int)(ph.bars_before + 1) * (unsigned int)G, wtid);
      }
#ifdef VS_PROGRAM_CLK
      if (b == 0 && wtid == 0) g_clk[52 + (pl & 3)] = clock64();
#endif
      if (hdr.trace && b == 0 && wtid == 0) hdr.trace[pl] = global_timer();
#ifdef VS_PROGRAM_CLK
      if (b == 0 && wtid == 0) g_clk[32 + (pl & 15)] = clock64();        // end of the barrier of phase pl
#endif
      p += step;
      cur ^= 1;
    }
    // leave the counters at zero for the next launch: the last CTA to get here resets them (every CTA has passed its
    // last barrier wait before it increments the exit counter)
    named_bar_sync(1, PG_WORKERS);
    if (wtid == 0) {
      __threadfence();
      unsigned int old = atomicAdd(hdr.counters + 1, 1u);
      if (old == (unsigned int)G - 1) {
        hdr.counters[0] = 0;
        hdr.counters[1] = 0;
        hdr.counters[2] = epoch + 1;
        __threadfence();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

}  // namespace

}  // namespace vs

// =================================================================================================================
// host side: program builder (opaque vs_program of include/vispec_b200.h)
// =================================================================================================================
struct vs_program {
  std::vector<vs::Phase> phases;
  // box_depth > 0: 3-D {64, rows, depth}; kv != 0: K / V rows of an attention phase (ld = head stride in elements,
  // rows = rows per head, cols = kv heads): 4-D {64 dims, rows, 2 halves, heads} with a {64, 32, 2, 1} box
  struct MapReq { const void* ptr; long ld; int rows, cols, box_rows, box_depth, kv; };
  struct SlabReg { const void* w; const void* slab; int pair, grid, packed; };
  std::vector<SlabReg> slabs; // row-major weight pointer -> slab image (vs_program_register_slab)
  int last_ssq_rows = 0;      // rows of [.][32] partial sums the last residual epilogue left in the ssq workspace
  bool x3d = true;            // activations of a slab GEMM through ONE 3-D tensor-map load per stage
  std::vector<MapReq> maps;
  std::vector<int> tables;    // chunk ranges per CTA + contributor info per tile of every GEMM phase
  std::vector<std::pair<const void*, const void*>> tiled;   // row-major weight pointer -> tiled copy
  size_t partial_floats = 0;
  size_t ssq_floats = 0;
  size_t tile_cnt_count = 0;  // per-tile arrival counters of all fused GEMM+epilogue pairs
  bool fuse_epilogue = true;  // epilogue phases wait on per-tile counters instead of a grid barrier
  bool fuse_norm = true;      // an RMSNorm that follows a residual epilogue runs inside that epilogue (no PH_NORM phase)
  bool attn_fused = false;    // attention as ONE phase (kv length limited to vs_attention_fused_capacity)
  size_t keys_count = 0;      // u64 entries of the arg-max / top-k candidate buffer
  size_t lstats_count = 0;    // float2 entries of the log-softmax slice statistics
  size_t hist_count = 0;      // u32 entries of the top-k warper's histograms ([rows][2][256])
  int top_k = 0;              // option "top_k": TopKLogitsWarper of the T > 0 acceptance phases added afterwards (0 = off)
  int grid = 0;
  int pf_ahead = 0;            // L2 prefetch distance (tensor-map prefetch costs as much TMA issue as the load itself: off)
  int inflight = 0;           // weight stages in flight per SM (0 = the whole ring)
  int spin_ns = 100;          // back-off of the single-lane role warps between barrier polls
  int debug = 0;              // ProgramHeader::debug
  int kv_two_loads = 0;       // option "kv_two_loads": force the 3-D form of the K / V tensor maps (A/B, tests)
  bool attn_prefetch = true;  // option "attn_prefetch": stage the stable K prefix of an attention phase ahead of the phase
  int attn_pf = 0;            // option "attn_pf": L2 prefetch distance of the KV producer in granules (0 = off; measured
                              // slower at every distance: 8192 keys 60.1 -> 62.6 / 65.7 / 69.4 us per layer at 8 / 16 / 32)
  bool committed = false;
  bool trace_on = false;
  // device workspace layout (set by commit)
  uint8_t* ws = nullptr;
  size_t off_trace_cta = 0;
  size_t off_tables = 0;
  size_t off_maps = 0, off_phases = 0, off_counters = 0, off_trace = 0, off_ssq = 0, off_tilecnt = 0, off_keys = 0, off_lstats = 0, off_hist = 0, off_partials = 0, total = 0;
  std::string err;
};

namespace vs {
namespace {

thread_local std::string g_perr;
int pfail(int code, const char* where, const std::string& what) {
  g_perr = std::string(where) + ": " + what;
  return code;
}

int device_sms() {     // of the CURRENT device (one process may drive several GPUs)
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cache[dev] = sms;
  }
  return cache[dev];
}

int add_map(vs_program* pg, const void* ptr, long ld, int rows, int cols, int box_rows, int box_depth = 0) {
  for (size_t i = 0; i < pg->maps.size(); ++i) {
    const auto& m = pg->maps[i];
    if (m.ptr == ptr && m.ld == ld && m.rows == rows && m.cols == cols && m.box_rows == box_rows && m.box_depth == box_depth &&
        !m.kv)
      return (int)i;
  }
  pg->maps.push_back({ptr, ld, rows, cols, box_rows, box_depth, 0});
  return (int)pg->maps.size() - 1;
}
int add_kv_map(vs_program* pg, const void* ptr, long head_stride, int n_kv) {
  const int rows = (int)(head_stride / 128);
  for (size_t i = 0; i < pg->maps.size(); ++i) {
    const auto& m = pg->maps[i];
    if (m.kv && m.ptr == ptr && m.ld == head_stride && m.cols == n_kv) return (int)i;
  }
  pg->maps.push_back({ptr, head_stride, rows, n_kv, ATT_GRAN, 0, 1});
  return (int)pg->maps.size() - 1;
}

Phase blank(int type, int T) {
  Phase ph;
  memset(&ph, 0, sizeof(ph));
  ph.type = type;
  ph.T = T;
  return ph;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void layout(vs_program* pg) {
  size_t off = 0;
  pg->off_maps = off;
  off = align_up(off + pg->maps.size() * sizeof(CUtensorMap), 256);
  pg->off_phases = off;
  off = align_up(off + pg->phases.size() * sizeof(Phase), 256);
  pg->off_tables = off;
  off = align_up(off + pg->tables.size() * sizeof(int), 256);
  pg->off_counters = off;
  off += 256;
  pg->off_trace = off;
  off = align_up(off + pg->phases.size() * sizeof(unsigned long long), 256);
  pg->off_trace_cta = off;
  if (pg->trace_on) off = align_up(off + pg->phases.size() * sizeof(unsigned long long) * (pg->grid > 0 ? pg->grid : device_sms()), 256);
  pg->off_ssq = off;
  off = align_up(off + pg->ssq_floats * sizeof(float), 256);
  pg->off_tilecnt = off;
  off = align_up(off + pg->tile_cnt_count * sizeof(unsigned int), 256);
  pg->off_keys = off;
  off = align_up(off + pg->keys_count * sizeof(unsigned long long), 256);
  pg->off_lstats = off;
  off = align_up(off + pg->lstats_count * sizeof(float2), 256);
  pg->off_hist = off;
  off = align_up(off + pg->hist_count * sizeof(unsigned int), 256);
  pg->off_partials = off;
  off += pg->partial_floats * sizeof(float);
  pg->total = align_up(off, 256);
}

}  // namespace

// ---- phase constructors used by the C API below and by the composite builders ------------------------------------
int prog_add_linear(vs_program* pg, const GemmArgs& a, float* ssq) {
  if (pg->committed) return pfail(VS_ERR_INVALID, "vs_program_add_linear", "program already committed");
  if (a.T < 1 || a.T > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_linear", "programs handle 1..32 token rows");
  if (a.K % BK != 0 || a.ldw % 8 != 0 || a.ldx % 8 != 0 || a.N < 1)
    return pfail(VS_ERR_INVALID, "vs_program_add_linear", "K must be a multiple of 64, ldx/ldw of 8");
  if ((a.epi == EPI_SWIGLU || a.epi == EPI_HEADS) && a.N % 128 != 0)
    return pfail(VS_ERR_INVALID, "vs_program_add_linear", "SWIGLU / HEADS need N % 128 == 0");
  const int G = pg->grid ? pg->grid : device_sms();
  pg->grid = G;
  // ---- slab path (slab.h): a registered slab image of these weights for this grid -> whole rows per CTA, epilogue from
  // TMEM inside the same phase
  {
    const int pair = (a.epi == EPI_SWIGLU || a.epi == EPI_HEADS) ? 1 : 0;
    const vs_program::SlabReg* reg = nullptr;
    if (a.ldw == a.K)
      for (const auto& r : pg->slabs)
        if (r.w == a.W && r.pair == pair && r.grid == G) reg = &r;
    SlabPlan sp;
    if (reg && slab_plan(a.N, a.K, pair, G, reg->packed, &sp)) {
      Phase g = blank(PH_GEMM2, a.T);
      g.N = a.N;
      g.nk = sp.nk;
      g.Wt = static_cast<const bf16*>(reg->slab);
      g.s_pair = sp.pair; g.s_ntiles = sp.n_tiles; g.s_passes = sp.passes; g.s_rows_alloc = sp.rows_alloc;
      g.s_blk_bytes = sp.blk_bytes; g.s_blk_smem = sp.blk_smem; g.s_kc = sp.kc; g.s_xoff = sp.x_off;
      g.s_eper = sp.e_per; g.s_erem = sp.e_rem; g.s_unit_stride = sp.unit_stride;
      g.s_packed = sp.packed; g.s_m64 = sp.m64;
      g.tmX = add_map(pg, a.X, a.ldx, a.T, a.K, VS_TOK);
      g.tmX3 = (pg->x3d && sp.kc > 1) ? add_map(pg, a.X, a.ldx, a.T, a.K, VS_TOK, sp.kc) : -1;
      g.epi = a.epi;
      g.bias = a.bias;
      g.Y = a.Y; g.ldy = a.ldy; g.R = a.R; g.ldr = a.ldr;
      g.ssq = ssq;
      if (ssq == PROG_WS_SSQ) {
        if ((size_t)G * VS_TOK > pg->ssq_floats) pg->ssq_floats = (size_t)G * VS_TOK;
        pg->last_ssq_rows = G;
      }
      g.n_q_tiles = a.n_q_tiles; g.n_k_tiles = a.n_k_tiles; g.rope = a.rope;
      g.cos = a.cos; g.sin = a.sin; g.pos_ids = a.pos_ids;
      g.Qout = a.Qout; g.q_head_stride = a.q_head_stride;
      g.Kout = a.Kout; g.Vout = a.Vout; g.kv_head_stride = a.kv_head_stride;
      g.kv_row_base = a.kv_row_base; g.kv_row_off = a.kv_row_off;
      pg->phases.push_back(g);
      return VS_OK;
    }
  }
  Phase g = blank(PH_GEMM, a.T);
  g.n_tiles = (a.N + BM - 1) / BM;
  g.nk = a.K / BK;
  g.N = a.N;
  const long C = (long)g.n_tiles * g.nk;
  if (C * G >= (1L << 31)) return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_linear", "matrix too large for 32-bit chunk ranges");
  const long per = (C + G - 1) / G;
  g.maxseg = (int)((per + g.nk - 1) / g.nk) + 1;
  // stream-K partition: CTA b owns per (+1 for the first rem CTAs) consecutive (tile, k-chunk) pairs
  {
    const int per_lo = (int)(C / G), rem = (int)(C % G);
    g.range_off = (int)pg->tables.size();
    std::vector<int> lo(G + 1);
    for (int b = 0; b <= G; ++b) lo[b] = b * per_lo + (b < rem ? b : rem);
    for (int b = 0; b < G; ++b) { pg->tables.push_back(lo[b]); pg->tables.push_back(lo[b + 1]); }
    while (pg->tables.size() % 4) pg->tables.push_back(0);
    auto magic = [](long d) -> unsigned int { return d <= 1 ? 0u : (unsigned int)((1ull << 32) / (unsigned long long)d + 1ull); };
    g.per_lo = per_lo;
    g.rem = rem;
    g.magic_per_hi = magic(per_lo + 1);
    g.magic_per_lo = magic(per_lo);
    g.magic_nk = magic(g.nk);
  }
  g.tmW = add_map(pg, a.W, a.ldw, a.N, a.K, BM);
  g.Wt = nullptr;
  if (a.ldw == a.K)
    for (const auto& pr : pg->tiled)
      if (pr.first == a.W) g.Wt = static_cast<const bf16*>(pr.second);
  g.tmX = add_map(pg, a.X, a.ldx, a.T, a.K, VS_TOK);
  const size_t need = (size_t)G * g.maxseg * TILE_FLOATS;
  if (need > pg->partial_floats) pg->partial_floats = need;
  g.no_barrier = pg->fuse_epilogue ? 1 : 0;
  // offset (in counters) into the workspace region, +1 so that 0 stays "none"; patched to a pointer at commit
  g.tile_cnt = pg->fuse_epilogue ? reinterpret_cast<unsigned int*>(uintptr_t(pg->tile_cnt_count + 1)) : nullptr;
  if (pg->fuse_epilogue) pg->tile_cnt_count += g.n_tiles;
  pg->phases.push_back(g);
  Phase e = g;
  e.type = PH_EPI;
  // token rows per epilogue item: 8; 16 where 8 would need a second round over the teams and the epilogue keeps no
  // per-group state (SiLU*up, plain store: gate|up of a 7B layer = 172 tiles x 4 groups = 688 items on 592 teams, LM heads)
  e.tpi = EPI_TPI;
  if ((a.epi == EPI_SWIGLU || (a.epi == EPI_STORE && !ssq)) && (long)g.n_tiles * ((a.T + EPI_TPI - 1) / EPI_TPI) > (long)PG_TEAMS * G &&
      !getenv("VS_EPI_NARROW"))
    e.tpi = 2 * EPI_TPI;
  e.epi = a.epi;
  e.bias = a.bias;
  e.Y = a.Y; e.ldy = a.ldy; e.R = a.R; e.ldr = a.ldr;
  e.ssq = ssq;
  if (ssq == PROG_WS_SSQ) {
    // [n_tiles][32] floats + the same entries as tagged 8-byte words for a folded-in RMSNorm (program_dev.cuh)
    if ((size_t)3 * g.n_tiles * VS_TOK > pg->ssq_floats) pg->ssq_floats = (size_t)3 * g.n_tiles * VS_TOK;
    pg->last_ssq_rows = g.n_tiles;
  }
  e.n_q_tiles = a.n_q_tiles; e.n_k_tiles = a.n_k_tiles; e.rope = a.rope;
  e.cos = a.cos; e.sin = a.sin; e.pos_ids = a.pos_ids;
  e.Qout = a.Qout; e.q_head_stride = a.q_head_stride;
  e.Kout = a.Kout; e.Vout = a.Vout; e.kv_head_stride = a.kv_head_stride;
  e.kv_row_base = a.kv_row_base; e.kv_row_off = a.kv_row_off;
  pg->phases.push_back(e);
  return VS_OK;
}

int prog_add_norm(vs_program* pg, const bf16* x, long ldx, const bf16* w, bf16* y, long ldy, int T, int H, float eps,
                  const float* ssq_in, int ssq_tiles) {
  if (T < 1 || T > VS_TOK || H % 8 || ldx % 8 || ldy % 8)
    return pfail(VS_ERR_INVALID, "vs_program_add_rmsnorm", "1..32 rows, H and strides multiples of 8");
  // RMSNorm right behind the residual epilogue that produced the rows (and their per-tile sums of squares): fold it
  // into that epilogue.  Each (tile, 8-token group) item publishes its sums of squares, waits on the group's arrival
  // counter for the other tiles and normalises the values it still holds in registers -- one phase, one grid barrier
  // and one read of the rows less.  All items of a group wait for each other, so every item needs its own team.
  if (ssq_in == PROG_WS_SSQ && pg->last_ssq_rows > 0) ssq_tiles = pg->last_ssq_rows;
  if (pg->fuse_norm && ssq_in && !pg->phases.empty()) {
    Phase& e = pg->phases.back();
    if (e.type == PH_GEMM2 && e.epi == EPI_RESID && e.ssq == ssq_in && e.Y == x && e.ldy == ldx && e.T == T && e.N == H &&
        e.s_passes == 1 && !e.w) {
      e.w = w; e.dst = y; e.ldd = ldy; e.H = H; e.eps = eps;
      e.norm_cnt = reinterpret_cast<unsigned int*>(uintptr_t(pg->tile_cnt_count + 1));
      pg->tile_cnt_count += 1;
      return VS_OK;
    }
  }
  if (pg->fuse_norm && pg->fuse_epilogue && ssq_in && !pg->phases.empty()) {
    Phase& e = pg->phases.back();
    const int G = pg->grid ? pg->grid : device_sms();
    const int groups = (T + EPI_TPI - 1) / EPI_TPI;
    if (e.type == PH_EPI && e.epi == EPI_RESID && e.ssq == ssq_in && e.Y == x && e.ldy == ldx && e.T == T && e.N == H &&
        e.n_tiles == ssq_tiles && !e.w && e.tpi == EPI_TPI && e.n_tiles * groups <= PG_TEAMS * G) {
      e.w = w; e.dst = y; e.ldd = ldy; e.H = H; e.eps = eps;
      e.norm_cnt = reinterpret_cast<unsigned int*>(uintptr_t(pg->tile_cnt_count + 1));
      pg->tile_cnt_count += 4;
      return VS_OK;
    }
  }
  Phase ph = blank(PH_NORM, T);
  ph.src = x; ph.lds = ldx; ph.w = w; ph.dst = y; ph.ldd = ldy; ph.H = H; ph.eps = eps;
  ph.ssq_in = ssq_in; ph.ssq_tiles = ssq_tiles;
  pg->phases.push_back(ph);
  return VS_OK;
}

bool prog_last_is_residual_of(const vs_program* pg, const bf16* x, long ldx, int T, int H) {
  if (!pg || pg->phases.empty()) return false;
  const Phase& e = pg->phases.back();
  return (e.type == PH_EPI || e.type == PH_GEMM2) && e.epi == EPI_RESID && e.ssq == PROG_WS_SSQ && e.Y == x && e.ldy == ldx &&
         e.T == T && e.N == H;
}

int prog_add_rows(vs_program* pg, const bf16* src, long lds, const int* idx, int n_src_rows, int broadcast, bf16* dst,
                  long ldd, int T, int H) {
  if (T < 1 || H % 8 || lds % 8 || ldd % 8)
    return pfail(VS_ERR_INVALID, "vs_program_add_rows", "H and strides must be multiples of 8");
  Phase ph = blank(PH_ROWS, T);
  ph.src = src; ph.lds = lds; ph.idx = idx; ph.n_src_rows = n_src_rows; ph.broadcast = broadcast;
  ph.dst = dst; ph.ldd = ldd; ph.H = H;
  pg->phases.push_back(ph);
  return VS_OK;
}

// kv splits of the fused single-phase attention: every (head, split) item needs its own resident 256-thread team
static int fused_splits(int G, int n_heads, int pv_splits_cap, int max_chunks) {
  int sp = (PG_BIG_TEAMS * G) / (n_heads > 0 ? n_heads : 1);
  if (sp > 8) sp = 8;     // measured: 8 splits beat 9 up to a few thousand keys (734 keys: 16.0 vs 17.1 us per layer; the
                          // statistics of a head fit one load per lane, item -> (head, split) is a shift); 8192 keys: 58.8 vs 57.0
  if (sp > pv_splits_cap) sp = pv_splits_cap;
  if (sp > max_chunks) sp = max_chunks;
  const char* e = getenv("VS_ATTN_SPLITS");       // A/B: fewer, longer splits per head
  if (e && atoi(e) >= 1 && atoi(e) < sp) sp = atoi(e);
  return sp;
}

int prog_add_attention(vs_program* pg, const AttnArgs& a) {
  if (a.Tq < 1 || a.Tq > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_attention", "programs handle 1..32 query rows");
  if (!a.scores || !a.stats || !a.opart || !a.counters || a.pv_splits < 1 || a.max_chunks < 1)
    return pfail(VS_ERR_INVALID, "vs_program_add_attention", "missing workspace");
  const int G = pg->grid ? pg->grid : device_sms();
  pg->grid = G;
  if (pg->attn_fused) {
    const int sp = fused_splits(G, a.n_heads, a.pv_splits, a.max_chunks);
    if (sp < 1 || a.n_heads * sp > PG_BIG_TEAMS * G)
      return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_attention", "too many heads for the fused attention phase");
    if (a.kv_head_stride % 128 != 0 || a.kv_head_stride < 128 || reinterpret_cast<uintptr_t>(a.K) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(a.V) % 16 != 0 || a.score_ld % ATT_GRAN != 0)
      return pfail(VS_ERR_INVALID, "vs_program_add_attention", "K / V rows must be 16-byte aligned with a head stride that is a multiple of 128 elements, score_ld a multiple of 32");
    Phase f = blank(PH_ATTN_FUSED, a.Tq);
    f.attn = a;
    f.attn.pv_splits = sp;
    f.tmK = add_kv_map(pg, a.K, a.kv_head_stride, a.n_kv_heads);
    f.tmV = add_kv_map(pg, a.V, a.kv_head_stride, a.n_kv_heads);
    // the prefix may be staged ahead only if nothing earlier in the program moves the kv length or the rows
    for (const auto& ph : pg->phases)
      if (ph.type == PH_TREE || ph.type == PH_SPLICE) f.attn.prefix_stable = 0;
    // per-head arrival counters live with the per-tile counters of the fused epilogues (+1: 0 stays "none")
    // (+ the tagged split statistics of attn_ring.cuh: [heads][splits][32 rows] x 16 bytes, 16-byte aligned)
    pg->tile_cnt_count = (pg->tile_cnt_count + 3) & ~size_t(3);
    f.tile_cnt = reinterpret_cast<unsigned int*>(uintptr_t(pg->tile_cnt_count + 1));
    pg->tile_cnt_count += ((2 * a.n_heads + 3) & ~3) + (size_t)a.n_heads * sp * VS_TOK * 4;
    pg->phases.push_back(f);
    return VS_OK;
  }
#if !VS_TWO_PHASE_ATTN
  return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_attention",
               "the two-phase attention program is not compiled into this build (option attn_fused = 0 needs -DVS_TWO_PHASE_ATTN=1)");
#endif
  Phase s = blank(PH_ATTN_S, a.Tq);
  s.attn = a;
  // as many kv splits as give every team one (head, split) item, bounded by the caller's workspace
  int sp = (PG_BIG_TEAMS * G) / (a.n_heads > 0 ? a.n_heads : 1);
  if (sp < 1) sp = 1;
  if (sp > a.pv_splits) sp = a.pv_splits;
  s.attn.pv_splits = sp;
  pg->phases.push_back(s);
  Phase v = s;
  v.type = PH_ATTN_PV;
  pg->phases.push_back(v);
  return VS_OK;
}

static int pick_slices(int V) {
  int n = 16;
  while ((long)n * (SCAN_MAX - 8) < V) n *= 2;      // a slice must fit the team's staging buffer
  return n;
}

int prog_add_argmax(vs_program* pg, const bf16* logits, long ldl, int T, int V, int* out, MergeInfo* fuse) {
  if (T < 1 || T > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_argmax_rows", "programs handle 1..32 rows");
  if (V % 8 || ldl % 8 || (reinterpret_cast<uintptr_t>(logits) & 15))
    return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_argmax_rows", "rows must be 16-byte aligned with V % 8 == 0");
  const int ns = pick_slices(V);
  if (ns > 64) return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_argmax_rows", "vocabulary too large");
  Phase ph = blank(PH_ARGMAX, T);
  ph.logits = logits; ph.ldl = ldl; ph.V = V; ph.amax_out = out; ph.n_slices = ns; ph.keys = PROG_WS_KEYS;
  pg->phases.push_back(ph);
  if ((size_t)T * ns > pg->keys_count) pg->keys_count = (size_t)T * ns;
  MergeInfo mi = {T, ns, 0, nullptr, nullptr, out};
  if (fuse) { *fuse = mi; return VS_OK; }
  TreeBufs none;
  memset(&none, 0, sizeof(none));
  return prog_add_tree(pg, TREE_MERGE_ARGMAX, none, nullptr, 0, nullptr, nullptr, nullptr, 0, &mi);
}

int prog_add_lsm_topk(vs_program* pg, const bf16* logits, long ldl, int T, int V, int k, const int* row_sel, float* top_val,
                      int* top_idx, MergeInfo* fuse) {
  if (k < 1 || k > 8) return pfail(VS_ERR_INVALID, "vs_program_add_logsoftmax_topk", "k must be in 1..8");
  if (T < 1 || T > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_logsoftmax_topk", "programs handle 1..32 rows");
  if (row_sel && T != 1) return pfail(VS_ERR_INVALID, "vs_program_add_logsoftmax_topk", "row_sel selects exactly one row");
  if (V % 8 || ldl % 8 || (reinterpret_cast<uintptr_t>(logits) & 15))
    return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_logsoftmax_topk", "rows must be 16-byte aligned with V % 8 == 0");
  const int ns = pick_slices(V);
  if (ns > 64) return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_logsoftmax_topk", "vocabulary too large");
  Phase ph = blank(PH_LSM_STATS, T);
  ph.logits = logits; ph.ldl = ldl; ph.V = V; ph.k = k; ph.row_sel = row_sel; ph.top_val = top_val; ph.top_idx = top_idx;
  ph.n_slices = ns; ph.keys = PROG_WS_KEYS; ph.lstats = PROG_WS_LSTATS;
  pg->phases.push_back(ph);
  ph.type = PH_LSM_CAND;
  pg->phases.push_back(ph);
  if ((size_t)T * ns * 8 > pg->keys_count) pg->keys_count = (size_t)T * ns * 8;
  if ((size_t)T * ns > pg->lstats_count) pg->lstats_count = (size_t)T * ns;
  MergeInfo mi = {T, ns, k, top_val, top_idx, nullptr};
  if (fuse) { *fuse = mi; return VS_OK; }
  TreeBufs none;
  memset(&none, 0, sizeof(none));
  return prog_add_tree(pg, TREE_MERGE_TOPK, none, nullptr, 0, nullptr, nullptr, nullptr, 0, &mi);
}

int prog_add_tree(vs_program* pg, int op, const TreeBufs& tb, int* state, int level, const int* amax, int* out_rec,
                  const int* sample_token_dev, int sample_token_host, const MergeInfo* merge) {
  if (op != TREE_MERGE_ARGMAX && op != TREE_MERGE_TOPK) {
    const int n_sc = tb.top_k + tb.depth * tb.top_k * tb.top_k;
    if (tb.top_k < 1 || tb.top_k > 8 || tb.depth + 2 > VS_RETRIEVE_LD || tb.total + 1 > 64 || tb.total > n_sc ||
        n_sc > TREE_MAX_SCORES)
      return pfail(VS_ERR_INVALID, "vs_program_add_tree_op", "unsupported tree shape");
  }
  Phase ph = blank(PH_TREE, 0);
  ph.tree_op = op; ph.tb = tb; ph.state = state; ph.level = level; ph.amax = amax; ph.out_rec = out_rec;
  ph.sample_token_dev = sample_token_dev; ph.sample_token_host = sample_token_host;
  if (merge) {
    ph.T = merge->T; ph.n_slices = merge->n_slices; ph.k = merge->k; ph.keys = PROG_WS_KEYS;
    ph.top_val = merge->top_val; ph.top_idx = merge->top_idx; ph.amax_out = merge->amax_out;
  }
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_accept_sampling(vs_program* pg, const TreeBufs& tb, int* state, const bf16* logits, long ldl, int V,
                             float temperature, const float* uni_override, int* out_rec) {
  if (!(temperature > 0.f)) return pfail(VS_ERR_INVALID, "vs_program_add_accept_sampling", "temperature must be > 0");
  const int T = tb.total + 1;
  if (T < 1 || T > VS_TOK) return pfail(VS_ERR_INVALID, "vs_program_add_accept_sampling", "tree larger than 32 nodes");
  if (V % 8 || ldl % 8 || (reinterpret_cast<uintptr_t>(logits) & 15))
    return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_accept_sampling", "rows must be 16-byte aligned with V % 8 == 0");
  const int ns = pick_slices(V);
  if (ns > 64) return pfail(VS_ERR_UNSUPPORTED, "vs_program_add_accept_sampling", "vocabulary too large");
  Phase st = blank(PH_LSM_STATS, T);
  st.logits = logits; st.ldl = ldl; st.V = V; st.n_slices = ns; st.lstats = PROG_WS_LSTATS; st.inv_temp = 1.0f / temperature;
  if (pg->top_k > 0) {
    // TopKLogitsWarper (utils.py:39-55): the key of the k-th largest logit of every tree row by a two-level radix select
    // (two histogram phases), then statistics, candidate probabilities and the residual sampling over the kept logits
    st.top_k = pg->top_k < V ? pg->top_k : V;
    st.hist = reinterpret_cast<unsigned int*>(uintptr_t(1));                 // patched at commit
    if ((size_t)VS_TOK * 512 > pg->hist_count) pg->hist_count = (size_t)VS_TOK * 512;
    for (int level = 0; level < 2; ++level) {
      Phase h = st;
      h.type = PH_TOPK_HIST;
      h.hist_level = level;
      pg->phases.push_back(h);
    }
  }
  pg->phases.push_back(st);
  if ((size_t)T * ns > pg->lstats_count) pg->lstats_count = (size_t)T * ns;
  Phase ph = st;
  ph.type = PH_TREE;
  ph.T = 0;
  ph.tree_op = TREE_ACCEPT_SAMPLE; ph.tb = tb; ph.state = state; ph.out_rec = out_rec; ph.uni_override = uni_override;
  pg->phases.push_back(ph);
  return VS_OK;
}

int prog_add_splice(vs_program* pg, bf16* slab, int n_slabs, int n_kv, long max_pos, const int* state, const int* sel_dev) {
  if (!slab || !state || !sel_dev) return pfail(VS_ERR_INVALID, "vs_program_add_kv_splice", "null operand");
  Phase ph = blank(PH_SPLICE, 0);
  ph.splice.slab = slab; ph.splice.head_stride = max_pos * 128; ph.splice.st = state; ph.splice.sel_dev = sel_dev;
  ph.n_items = n_slabs * n_kv;
  pg->phases.push_back(ph);
  return VS_OK;
}

const char* prog_last_error() { return g_perr.c_str(); }
int prog_fail(int code, const char* where, const char* what) { return pfail(code, where, what); }

// encode one tensor map on the host
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn prog_encode_fn() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(sym);
  }
  return fn;
}

}  // namespace vs

extern "C" {

int vs_program_create(vs_program** out) {
  if (!out) return vs::pfail(VS_ERR_INVALID, "vs_program_create", "null out");
  *out = new vs_program();
  const char* e = getenv("VS_PF_AHEAD");
  if (e) (*out)->pf_ahead = atoi(e);
  e = getenv("VS_INFLIGHT");
  if (e) (*out)->inflight = atoi(e);
  e = getenv("VS_FUSE_EPILOGUE");
  if (e) (*out)->fuse_epilogue = atoi(e) != 0;
  e = getenv("VS_FUSE_NORM");
  if (e) (*out)->fuse_norm = atoi(e) != 0;
  e = getenv("VS_SPIN_NS");
  if (e) (*out)->spin_ns = atoi(e);
  e = getenv("VS_X3D");
  if (e) (*out)->x3d = atoi(e) != 0;
  e = getenv("VS_PG_DEBUG");
  if (e) (*out)->debug = atoi(e);
  e = getenv("VS_ATTN_PF");           // A/B switches of the TMA-staged attention (attn_ring.cuh)
  if (e) (*out)->attn_pf = atoi(e);
  e = getenv("VS_ATTN_PREFETCH");
  if (e) (*out)->attn_prefetch = atoi(e) != 0;
  return VS_OK;
}

void vs_program_destroy(vs_program* pg) { delete pg; }

const char* vs_program_last_error(void) { return vs::prog_last_error(); }

int vs_program_num_phases(const vs_program* pg) { return pg ? (int)pg->phases.size() : 0; }

int vs_program_set_option(vs_program* pg, const char* name, int value) {
  if (!pg || !name) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "null argument");
  if (!strcmp(name, "pf_ahead")) pg->pf_ahead = value;
  else if (!strcmp(name, "trace")) pg->trace_on = value != 0;
  else if (!strcmp(name, "inflight")) pg->inflight = value;
  else if (!strcmp(name, "attn_fused")) {
    if (!pg->phases.empty()) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "attn_fused must be set before adding phases");
    pg->attn_fused = value != 0;
  } else if (!strcmp(name, "fuse_epilogue")) {
    if (!pg->phases.empty()) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "fuse_epilogue must be set before adding phases");
    pg->fuse_epilogue = value != 0;
  }
  else if (!strcmp(name, "fuse_norm")) {
    if (!pg->phases.empty()) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "fuse_norm must be set before adding phases");
    pg->fuse_norm = value != 0;
  }
  else if (!strcmp(name, "spin_ns")) pg->spin_ns = value;
  else if (!strcmp(name, "x3d")) pg->x3d = value != 0;
  else if (!strcmp(name, "debug")) pg->debug = value;
  else if (!strcmp(name, "kv_two_loads")) pg->kv_two_loads = value != 0;
  else if (!strcmp(name, "attn_prefetch")) pg->attn_prefetch = value != 0;
  else if (!strcmp(name, "attn_pf")) pg->attn_pf = value;
  else if (!strcmp(name, "top_k")) pg->top_k = value > 0 ? value : 0;

  else if (!strcmp(name, "grid")) {
    if (!pg->phases.empty()) return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "grid must be set before adding phases");
    pg->grid = value;
  } else return vs::pfail(VS_ERR_INVALID, "vs_program_set_option", "unknown option");
  return VS_OK;
}

int vs_attention_fused_capacity(int n_heads, int pv_splits_cap, int max_chunks) {
  // beyond 2 * ATT_RING_NG granules per split (scores in registers) the phase parks its scores: the scratch is the limit
  return vs::fused_splits(vs::device_sms(), n_heads, pv_splits_cap, max_chunks) < 1 ? 0 : max_chunks * vs::ATT_CHUNK;
}

int vs_program_register_tiled(vs_program* pg, const void* w_row_major, const void* w_tiled) {
  if (!pg || !w_row_major || !w_tiled) return vs::pfail(VS_ERR_INVALID, "vs_program_register_tiled", "null argument");
  if (reinterpret_cast<uintptr_t>(w_tiled) % 1024) return vs::pfail(VS_ERR_INVALID, "vs_program_register_tiled", "tiled weights must be 1024-byte aligned");
  pg->tiled.emplace_back(w_row_major, w_tiled);
  return VS_OK;
}

int vs_device_sms(void) { return vs::device_sms(); }

int vs_program_register_slab(vs_program* pg, const void* w_row_major, const void* w_slab, int epi, int grid) {
  if (!pg || !w_row_major || !w_slab) return vs::pfail(VS_ERR_INVALID, "vs_program_register_slab", "null argument");
  if (reinterpret_cast<uintptr_t>(w_slab) % 128) return vs::pfail(VS_ERR_INVALID, "vs_program_register_slab", "slab image must be 128-byte aligned");
  if (grid <= 0) grid = vs::device_sms();
  const int pair = (epi == vs::EPI_SWIGLU || epi == vs::EPI_HEADS) ? 1 : 0;
  pg->slabs.push_back({w_row_major, w_slab, pair, grid, vs::slab_default_packed()});
  return VS_OK;
}

size_t vs_program_workspace_bytes(vs_program* pg) {
  if (!pg) return 0;
  vs::layout(pg);
  return pg->total;
}

int vs_program_commit(vs_program* pg, void* workspace, size_t bytes, void* stream) {
  using namespace vs;
  if (!pg || !workspace) return pfail(VS_ERR_INVALID, "vs_program_commit", "null argument");
  if (pg->phases.empty()) return pfail(VS_ERR_INVALID, "vs_program_commit", "empty program");
  if (pg->grid <= 0) pg->grid = device_sms();
  layout(pg);
  if (bytes < pg->total) return pfail(VS_ERR_INVALID, "vs_program_commit", "workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) % 256) return pfail(VS_ERR_INVALID, "vs_program_commit", "workspace must be 256-byte aligned");
  EncodeFn enc = prog_encode_fn();
  if (!enc && !pg->maps.empty()) return pfail(VS_ERR_CUDA, "vs_program_commit", "cuTensorMapEncodeTiled unavailable");
  pg->ws = static_cast<uint8_t*>(workspace);
  std::vector<uint8_t> host(pg->off_partials, 0);
  for (size_t i = 0; i < pg->maps.size(); ++i) {
    const auto& m = pg->maps[i];
    CUtensorMap tm;
    CUresult r;
    if (m.kv) {
      // K or V rows [heads][rows][128]: {64 dims, rows, 2 halves of a row, heads}; one box = 32 rows x both halves lands as
      // [half][32 rows][64 dims] = two 128-byte-swizzled 4 KB tiles (attn_ring.cuh).  If the driver refuses the view
      // (the half stride is smaller than the row stride), a 3-D {128, rows, heads} map with two loads per granule.
      cuuint64_t dims[4] = {64, (cuuint64_t)m.rows, 2, (cuuint64_t)m.cols};
      cuuint64_t strides[3] = {256, 128, (cuuint64_t)m.ld * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)m.box_rows, 2, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if (!pg->kv_two_loads) {
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(m.ptr), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {       // every K / V map of the program takes the 3-D form: start over
          pg->kv_two_loads = 1;
          i = (size_t)-1;
          continue;
        }
      } else {
        cuuint64_t dims3[3] = {128, (cuuint64_t)m.rows, (cuuint64_t)m.cols};
        cuuint64_t strides3[2] = {256, (cuuint64_t)m.ld * 2};
        cuuint32_t box3[3] = {64, (cuuint32_t)m.box_rows, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(m.ptr), dims3, strides3, box3, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return pfail(VS_ERR_CUDA, "vs_program_commit", "cuTensorMapEncodeTiled failed (K / V rows)");
      }
    } else if (m.box_depth > 0) {
      // [rows][cols] viewed as {64 columns, rows, cols / 64 chunks}: one box = box_depth consecutive [rows x 64] blocks,
      // each landing as its own 128-byte-swizzled 4 KB tile (the activation side of a slab stage)
      cuuint64_t dims[3] = {(cuuint64_t)BK, (cuuint64_t)m.rows, (cuuint64_t)(m.cols / BK)};
      cuuint64_t strides[2] = {(cuuint64_t)m.ld * 2, (cuuint64_t)BK * 2};
      cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)m.box_rows, (cuuint32_t)m.box_depth};
      cuuint32_t estr[3] = {1, 1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(m.ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dims[2] = {(cuuint64_t)m.cols, (cuuint64_t)m.rows};
      cuuint64_t strides[1] = {(cuuint64_t)m.ld * 2};
      cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)m.box_rows};
      cuuint32_t estr[2] = {1, 1};
      r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(m.ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) {
      if (m.box_depth > 0) {   // the driver refused the 3-D view: fall back to kc 2-D loads per stage
        for (auto& ph : pg->phases)
          if (ph.type == PH_GEMM2 && ph.tmX3 == (int)i) ph.tmX3 = -1;
        memset(&tm, 0, sizeof(tm));
      } else {
        return pfail(VS_ERR_CUDA, "vs_program_commit", "cuTensorMapEncodeTiled failed");
      }
    }
    memcpy(host.data() + pg->off_maps + i * sizeof(CUtensorMap), &tm, sizeof(tm));
  }
  memcpy(host.data() + pg->off_tables, pg->tables.data(), pg->tables.size() * sizeof(int));
  float* partials = reinterpret_cast<float*>(pg->ws + pg->off_partials);
  int bars = 0;
  int scratch_free_bars = 0;     // grid barriers after which the attention teams' scratch is no longer in use by a phase
  for (size_t i = 0; i < pg->phases.size(); ++i) {
    Phase ph = pg->phases[i];
    ph.bars_before = bars;
    ph.phase_idx = (int)i;
    if (ph.type == PH_ATTN_FUSED) {
      ph.kv_two_loads = pg->kv_two_loads;
      ph.pf_bars = (ph.attn.prefix_stable && pg->attn_prefetch) ? scratch_free_bars : bars;
    }
    if (ph.type != PH_GEMM && ph.type != PH_EPI && ph.type != PH_NORM && ph.type != PH_ROWS && ph.type != PH_SPLICE)
      scratch_free_bars = bars + 1;
    if (!(ph.type == PH_GEMM && ph.no_barrier)) ++bars;
    if (ph.tile_cnt)
      ph.tile_cnt = reinterpret_cast<unsigned int*>(pg->ws + pg->off_tilecnt) + (reinterpret_cast<uintptr_t>(ph.tile_cnt) - 1);
    if (ph.norm_cnt)
      ph.norm_cnt = reinterpret_cast<unsigned int*>(pg->ws + pg->off_tilecnt) + (reinterpret_cast<uintptr_t>(ph.norm_cnt) - 1);
    if (ph.type == PH_GEMM || ph.type == PH_EPI) ph.partials = partials;
    if (ph.type == PH_GEMM2 && ph.tmX3 < 0) ph.tmX3 = -1;
    if (ph.ssq == PROG_WS_SSQ) ph.ssq = reinterpret_cast<float*>(pg->ws + pg->off_ssq);
    if (ph.ssq_in == PROG_WS_SSQ) ph.ssq_in = reinterpret_cast<float*>(pg->ws + pg->off_ssq);
    if (ph.keys == PROG_WS_KEYS) ph.keys = reinterpret_cast<unsigned long long*>(pg->ws + pg->off_keys);
    if (ph.lstats == PROG_WS_LSTATS) ph.lstats = reinterpret_cast<float2*>(pg->ws + pg->off_lstats);
    if (ph.hist) ph.hist = reinterpret_cast<unsigned int*>(pg->ws + pg->off_hist);
    memcpy(host.data() + pg->off_phases + i * sizeof(Phase), &ph, sizeof(Phase));
  }
  // A GEMM that is fused with its epilogue (no grid barrier in between) is executed together with it from ONE descriptor:
  // the epilogue's (patched) descriptor was copied from the GEMM's when both were built, so it holds every GEMM field too.
  {
    Phase* hp = reinterpret_cast<Phase*>(host.data() + pg->off_phases);
    for (size_t i = 0; i + 1 < pg->phases.size(); ++i)
      if (hp[i].type == PH_GEMM && hp[i].no_barrier && hp[i + 1].type == PH_EPI) {
        Phase m = hp[i + 1];
        m.type = PH_GEMM;
        m.bars_before = hp[i].bars_before;
        hp[i] = m;
      }
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(pg->ws, host.data(), host.size(), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess)
    return pfail(VS_ERR_CUDA, "vs_program_commit", cudaGetErrorString(cudaGetLastError()));
  {
    static bool attr_set[64] = {false};      // a function attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e != cudaSuccess) return pfail(VS_ERR_CUDA, "vs_program_commit", cudaGetErrorString(e));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  pg->committed = true;
  return VS_OK;
}

int vs_program_run(const vs_program* pg, void* stream) {
  using namespace vs;
  if (!pg || !pg->committed) return pfail(VS_ERR_INVALID, "vs_program_run", "program not committed");
  ProgramHeader h;
  h.phases = reinterpret_cast<const Phase*>(pg->ws + pg->off_phases);
  h.n_phases = (int)pg->phases.size();
  h.maps = reinterpret_cast<const CUtensorMap*>(pg->ws + pg->off_maps);
  h.tables = reinterpret_cast<const int*>(pg->ws + pg->off_tables);
  h.counters = reinterpret_cast<unsigned int*>(pg->ws + pg->off_counters);
  h.trace = pg->trace_on ? reinterpret_cast<unsigned long long*>(pg->ws + pg->off_trace) : nullptr;
  h.trace_cta = pg->trace_on ? reinterpret_cast<unsigned long long*>(pg->ws + pg->off_trace_cta) : nullptr;
  h.pf_ahead = pg->pf_ahead;
  h.inflight = pg->inflight;
  h.spin_ns = pg->spin_ns;
  h.debug = pg->debug;
  h.attn_pf = pg->attn_pf;
  bool has_slab = false;
  for (const auto& ph : pg->phases) has_slab = has_slab || ph.type == PH_GEMM2;
  h.ns = has_slab ? RING_BYTES / SLAB_SLOT_BYTES : NS_MAX;
  h.stage_bytes = has_slab ? SLAB_SLOT_BYTES : W_STAGE + X_STAGE;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pg->grid);
  cfg.blockDim = dim3(PG_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;       // all CTAs co-resident (the phases synchronise across the grid)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, program_kernel, h);
  count_launch();
  if (e != cudaSuccess) return pfail(VS_ERR_CUDA, "vs_program_run", cudaGetErrorString(e));
  return VS_OK;
}

/* per-phase end times (ns, %globaltimer of CTA 0) of the last run; needs option "trace" = 1 */
int vs_program_read_trace(const vs_program* pg, unsigned long long* out, int cap) {
  using namespace vs;
  if (!pg || !pg->committed || !pg->trace_on) return pfail(VS_ERR_INVALID, "vs_program_read_trace", "trace not enabled");
  int n = (int)pg->phases.size();
  if (n > cap) n = cap;
  if (cudaMemcpy(out, pg->ws + pg->off_trace, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return pfail(VS_ERR_CUDA, "vs_program_read_trace", cudaGetErrorString(cudaGetLastError()));
  return n;
}
/* per-phase, per-CTA "items done" stamps (ns) of the last run: out[phase * grid + cta]; returns the grid size */
int vs_program_read_trace_cta(const vs_program* pg, unsigned long long* out, int cap) {
  using namespace vs;
  if (!pg || !pg->committed || !pg->trace_on) return pfail(VS_ERR_INVALID, "vs_program_read_trace_cta", "trace not enabled");
  long n = (long)pg->phases.size() * pg->grid;
  if (n > cap) return pfail(VS_ERR_INVALID, "vs_program_read_trace_cta", "buffer too small");
  if (cudaMemcpy(out, pg->ws + pg->off_trace_cta, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return pfail(VS_ERR_CUDA, "vs_program_read_trace_cta", cudaGetErrorString(cudaGetLastError()));
  return pg->grid;
}
/* -DVS_PROGRAM_CLK builds only: the checkpoint clocks of program_dev.cuh (64 values); returns 0 when not compiled in */
int vs_debug_clk(long long* out) {
#ifdef VS_PROGRAM_CLK
  cudaMemcpyFromSymbol(out, vs::g_clk, 64 * sizeof(long long));
  return 64;
#else
  (void)out;
  return 0;
#endif
}
int vs_program_phase_types(const vs_program* pg, int32_t* out, int cap) {
  if (!pg) return 0;
  int n = (int)pg->phases.size();
  if (n > cap) n = cap;
  for (int i = 0; i < n; ++i) out[i] = pg->phases[i].type;
  return n;
}

}  // extern "C"
