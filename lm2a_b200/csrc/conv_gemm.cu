// Implicit-GEMM Conv1d / Linear for channels-last bf16 slabs on sm_100a.
//
//   out[m, n] = sum_{seg, tap, c} f_seg(A_seg)[m + shift(tap), c] * W[n, k(seg,tap,c)] + bias[n]
//
// Replaces nn.Conv1d (k=1, k=3 p=1, k=4 s=2 p=1) and every nn.Linear / MHA projection on the
// sampling path (reference models/unet1d_ultimate.py:87-88,115,216-221,255-261,295,364;
// models/cross_attention.py:19-36). Fused around the contraction:
//   * consumer side: the GroupNorm + SiLU in front of a conv (unet1d_ultimate.py:136-137,
//     146-147, 362-363) is applied to the operand tile in shared memory between its TMA
//     landing and the MMA (f_seg above; per-(clip-row, group) mean / rstd from the producer's
//     exact integer sums), so the normalised tensor never exists in HBM;
//   * producer side (epilogue): bias, FiLM h*(1+scale)+shift (:141-143), the ResBlock skip 1x1
//     conv (second K segment accumulating into the same TMEM tile), the residual add (:159),
//     and the GroupNorm sums of the OUTPUT for the next consumer.
//
// One CTA per SM, persistent over 128 x BLOCK_N output tiles (CG = 2: a CTA pair per
// 256 x BLOCK_N tile, cta_group::2 UMMA). Two pipelines behind one epilogue (template XF):
//   plain (XF = 0)   warp 0 TMA producer: per K block one 128-slot A box (a tap is a row shift
//               of the flattened slab; the zero slot between clips and TMA's out-of-bounds zero
//               fill are the conv padding) and one W box in the same stage; warp 1 MMA issuer
//               (one thread; one barrier wait + one commit per K block), fp32 accumulators
//               double-buffered in TMEM; warps 2-9 epilogue
//   operand transform (XF = 1)   an "A block" is ONE box of 128 (+ 2 halo) slots x 64 channels
//               that serves every tap: a row-shifted view of a 128B-swizzled tile is the same
//               UMMA descriptor with its start address advanced by 128 B per row (the tensor
//               core applies the swizzle to absolute shared-memory address bits, see
//               umma_desc_sw128_rows). W boxes ride in a ring of their own. Warps 2-5 wait for
//               an A block, apply y = SiLU(GroupNorm(x)) to it in place (one pass per element,
//               not per tap), fence to the async proxy and hand it to the MMA warp; in a CTA
//               pair warp 6 forwards the hand-off to the leader at cluster scope; the next 8
//               warps are the epilogue
//   epilogue    tcgen05.ld -> bias / FiLM / residual -> GroupNorm sums -> bf16 slab via TMA
//               store (or the final fp32 [R, C, T] eps tensor); warp w reads TMEM lane quadrant
//               w % 4 and one half of the columns
//
// GroupNorm sums are exact: every lane's per-slot channel sums (a fixed-order fp32 sum that does
// not depend on where the slot sits in the batch) are converted to 40.24 / 44.20 fixed point and
// accumulated with 64-bit integer adds (warp shuffle tree, then one atomic per (clip-row,
// group)). Integer addition is associative, so the statistics - and with them every output bit -
// are independent of tile shape, batch position and sharding.
#include "../../include/lm2a_b200.h"
#include <stdlib.h>

#include "common.cuh"

namespace lm2a {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kASlotRows = 136;                      // 128 + 2 halo rows, rounded to 8-row atoms
constexpr int kASlotBytes = kASlotRows * kBlockK * 2;  // 17408 = 17 * 1024
constexpr int kATileBytes = kBlockM * kBlockK * 2;     // 16384: a plain 128-row box
// Operand-transform warps of an XF launch: one per scheduler. (Two per scheduler, with the
// register file re-balanced by setmaxnreg - 40 / 80 / 136 registers for the producer+MMA /
// transform / epilogue warpgroups - measured SLOWER on B200: 39.5 vs 36.7 us at M 4160 x N 1024
// x K 3072. The transform is not latency-bound but shares the SM's 128 B/clk of shared-memory
// bandwidth with the UMMA operand reads and the TMA fills, which already take ~107 B/clk.)
constexpr int kEpiWarps = 8;  // two per TMEM lane quadrant, each takes half of the columns
constexpr int kXformWarps = 4;
constexpr int kFirstXformWarp = 2;
constexpr int kRowStride = 4 * kXformWarps;                          // rows between a thread's rows
constexpr int kRowsPerThread = (130 + kRowStride - 1) / kRowStride;  // of a 130-row A block
constexpr int kRowBatch = 3;                                         // rows in flight per thread
// warps: 0 TMA producer, 1 MMA issuer, [2, 6) operand transform, 6 pair hand-off (XF pair
// launches only), then 8 epilogue warps; plain launches: 0, 1, then the epilogue
__host__ __device__ constexpr int first_epi_warp(bool xf, int cg) {
  return xf ? kFirstXformWarp + kXformWarps + (cg == 2 ? 1 : 0) : 2;
}
__host__ __device__ constexpr int num_threads(bool xf, int cg) {
  return 32 * (first_epi_warp(xf, cg) + kEpiWarps);
}
constexpr int kMaxGnChannels = 2048;   // gamma / beta of the input GroupNorm staged in smem
constexpr int kMaxGnEntries = 512;     // (clip-rows touched by a tile) x groups
constexpr float kStatScale1 = 16777216.0f;   // 2^24: sum
constexpr float kStatScale2 = 1048576.0f;    // 2^20: sum of squares

struct ConvArgs {
  int seg_cblk[2];   // cin / 64 per segment
  int seg_taps[2];   // LM2A_TAPS_*
  int seg_half[2];   // K4S2: channel offset of the odd slot inside a slot pair
  int num_kb;        // 64-deep K blocks per output tile (all segments and taps)
  int m_tiles, n_tiles;
  long long m;
  int tp, t_valid, n_valid;
  const float* bias;
  const float* film;
  int film_ld, film_shift_off;
  const __nv_bfloat16* residual;
  int res_ld;
  void* out;
  int out_ld;
  int out_mode;
  unsigned long long* stats;  // exact GroupNorm sums of the output (or null): [R][stats_pitch][2]
  int stats_pitch;            // groups per clip-row in the stats buffer
  int stats_cg;               // channels per group
  int stats_c0;               // channel of the normalised tensor that output column 0 maps to
  // GroupNorm (+ SiLU) applied to segment 0 on the fly
  const long long* gn_stats;  // null = raw operand
  const float* gn_gamma;
  const float* gn_beta;
  int gn_pitch, gn_groups, gn_cg;
  unsigned int gn_cg_magic;   // ceil(2^22 / gn_cg): c / gn_cg == (c * magic) >> 22 for c < 2048
  float gn_eps;
  // x2 linear upsampling (align_corners) of segment 0 on the fly (XF = 2): the low-res slab
  const __nv_bfloat16* up_src;
  int up_ld, up_tp_in, up_t_in;
  int share_taps;             // 1: one A block serves all taps (row-shifted views); 0: one
                              //    128-slot box per tap
  int tap_outer;              // plain launches: K walk order (1: tap outer, channel block inner)
  int dbg_stats;              // timing experiments only: 1 = sums without atomics, 2 = atomics only
  int dbg_noshift;            // timing experiments only: every tap reads the unshifted view
  int dbg_noxform;            // timing experiments only: transform warps pass blocks through
};

// XF: the launch builds segment 0's operand tiles on the fly (transform warps active): 1 =
// GroupNorm + SiLU of the landed tile, 2 = x2 linear upsampling of a lower-resolution slab.
// Plain launches (0) have no such warps and a single ring of combined stages.
template <int BLOCK_N, int CG, int XF>
struct SmemLayout {
  static constexpr int kBSlotBytes = (BLOCK_N / CG) * kBlockK * 2;  // a CTA pair splits W along N
  // XF: a ring of A blocks (136 rows: one box serves every tap) and a ring of W blocks.
  // Plain: ONE ring of stages, each a 128-row A box plus its W box behind one barrier pair
  // (one wait and one commit per K block for the MMA-issuing thread).
  static constexpr int kASlot = XF ? kASlotBytes : kATileBytes;
  static constexpr int kBStages = XF ? (kBSlotBytes == 32768 ? 4 : (kBSlotBytes == 16384 ? 6 : 8))
                                     : (kBSlotBytes == 32768 ? 4 : (kBSlotBytes == 16384 ? 6 : 8));
  static constexpr int kAStages = XF ? (kBSlotBytes == 32768 ? 3 : 4) : kBStages;
  static constexpr int kAOffset = 0;
  static constexpr int kBOffset = kAStages * kASlot;
  // epilogue: per warp a [32 rows][32 cols] bf16 staging tile (TMA store source, 64B swizzle)
  // and a table of per-column (scale, offset) pairs for its BLOCK_N / 2 columns
  static constexpr int kOutOffset = kBOffset + kBStages * kBSlotBytes;
  static constexpr int kOutBytesPerWarp = 32 * 32 * 2;
  static constexpr int kTabOffset = kOutOffset + kEpiWarps * kOutBytesPerWarp;
  static constexpr int kTabBytesPerWarp = (BLOCK_N / 2) * 8;
  // operand transform: gamma | beta of the input GroupNorm, (mean, rstd) per (clip-row, group)
  // of the current tile, clip-row index of every slot of the A block
  static constexpr int kGammaOffset = kTabOffset + kEpiWarps * kTabBytesPerWarp;
  static constexpr int kMrOffset = kGammaOffset + (XF == 1 ? 2 * kMaxGnChannels * 4 : 0);
  static constexpr int kRowInfoOffset = kMrOffset + (XF == 1 ? kMaxGnEntries * 8 : 0);
  static constexpr int kBarOffset = kRowInfoOffset + (XF == 1 ? kASlotRows * 4 : 0);
  static constexpr int kNumBars = 4 * kAStages + 2 * kBStages + 4;
  static constexpr int kBytes = kBarOffset + 8 * kNumBars + 16 + 1024;  // + tmem slot + align
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
};

// 128B-swizzled K-major operand descriptor of a tile that starts `row_shift` (< 8) rows into a
// 1024-byte aligned slot: the start address advances by whole 128-byte rows and nothing else
// changes. The tensor core applies the swizzle XOR to the absolute shared-memory address bits
// (measured on B200: with the matrix-base-offset field, bits 49-51, set to the row phase the
// result is wrong; with the field left 0 every tap view is exact), exactly like the +32 B
// K-step advance inside a swizzle row.
__device__ __forceinline__ uint64_t umma_desc_sw128_rows(uint32_t slot_addr, uint32_t row_shift) {
  return umma_desc_sw128_kmajor(slot_addr + row_shift * 128u);
}

__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      printf("lm2a: mbarrier (cluster) wait timeout (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Exact 64-bit sum over the warp with three hardware integer reductions (REDUX): the value is
// split into a signed high word and two unsigned 16-bit limbs of the low word, whose 32-term
// sums cannot overflow 32 bits.
__device__ __forceinline__ long long warp_sum_ll(long long v) {
  const unsigned lo = (unsigned)(unsigned long long)v;
  const int hi = (int)(v >> 32);
  const unsigned s0 = __reduce_add_sync(0xffffffffu, lo & 0xffffu);
  const unsigned s1 = __reduce_add_sync(0xffffffffu, lo >> 16);
  const int s2 = __reduce_add_sync(0xffffffffu, hi);
  return ((long long)s2 << 32) + ((long long)s1 << 16) + (long long)s0;
}

// The K loop of one output tile as a sequence of A blocks, each followed by the W blocks of
// the taps it serves. on_a(seg, cb, row0, chan): box at slot m0 + row0, channel `chan` of the
// segment's tensor map. on_b(shift, kb): tap view `shift` rows into the A block, W K-block kb.
// share_taps == 0: every tap gets a 128-slot box of its own (row0 = the tap's shift, view 0).
template <typename FA, typename FB>
__device__ __forceinline__ void walk_tile(const ConvArgs& p, FA&& on_a, FB&& on_b) {
  int kb_base = 0;
#pragma unroll 1
  for (int seg = 0; seg < 2; ++seg) {
    const int cblk = p.seg_cblk[seg];
    if (cblk == 0) continue;
    const int mode = p.seg_taps[seg];
    const int ntaps_all = mode == LM2A_TAPS_K1 ? 1 : (mode == LM2A_TAPS_K3 ? 3 : 4);
    if (!p.share_taps && p.tap_outer) {
      // tap outer, channel block inner: consecutive boxes walk the channels of the same slots
      // and W is read contiguously along K
#pragma unroll 1
      for (int tap = 0; tap < ntaps_all; ++tap) {
        int row0 = 0, choff = 0;
        if (mode == LM2A_TAPS_K3) {
          row0 = tap - 1;
        } else if (mode == LM2A_TAPS_K4S2) {
          row0 = tap == 0 ? -1 : (tap == 3 ? 1 : 0);
          choff = (tap == 0 || tap == 2) ? p.seg_half[seg] : 0;
        }
#pragma unroll 1
        for (int cb = 0; cb < cblk; ++cb) {
          on_a(seg, cb, row0, choff + cb * kBlockK);
          on_b(0, kb_base + tap * cblk + cb);
        }
      }
      kb_base += cblk * ntaps_all;
      continue;
    }
    if (!p.share_taps) {
      // same K order as the shared-block walk (channel block outer, tap inner; for k4s2 the odd
      // half's taps 0, 2 before the even half's 1, 3): the accumulation order, and with it every
      // output bit, does not depend on how the operand is staged
#pragma unroll 1
      for (int cb = 0; cb < cblk; ++cb) {
#pragma unroll 1
        for (int j = 0; j < ntaps_all; ++j) {
          int tap = j, row0 = 0, choff = 0;
          if (mode == LM2A_TAPS_K3) {
            row0 = tap - 1;
          } else if (mode == LM2A_TAPS_K4S2) {
            tap = (j & 1) * 2 + (j >> 1);          // 0, 2, 1, 3
            row0 = tap == 0 ? -1 : (tap == 3 ? 1 : 0);
            choff = (tap == 0 || tap == 2) ? p.seg_half[seg] : 0;
          }
          on_a(seg, cb, row0, choff + cb * kBlockK);
          on_b(0, kb_base + tap * cblk + cb);
        }
      }
      kb_base += cblk * ntaps_all;
      continue;
    }
    const int nhalf = mode == LM2A_TAPS_K4S2 ? 2 : 1;
    const int ntap = mode == LM2A_TAPS_K1 ? 1 : (mode == LM2A_TAPS_K3 ? 3 : 2);
#pragma unroll 1
    for (int cb = 0; cb < cblk; ++cb) {
#pragma unroll 1
      for (int hf = 0; hf < nhalf; ++hf) {
        // K3: slots m-1, m, m+1 -> one box from m0 - 1, tap j = view shifted by j rows.
        // K4S2 over slot pairs: x[2t-1+tap]; hf = 0 is the odd half of a pair (taps 0, 2 =
        // pairs t-1, t -> box from m0 - 1), hf = 1 the even half (taps 1, 3 = pairs t, t+1).
        const int row0 = (mode == LM2A_TAPS_K3 || (mode == LM2A_TAPS_K4S2 && hf == 0)) ? -1 : 0;
        const int chan = ((mode == LM2A_TAPS_K4S2 && hf == 0) ? p.seg_half[seg] : 0) + cb * kBlockK;
        on_a(seg, cb, row0, chan);
#pragma unroll 1
        for (int j = 0; j < ntap; ++j) {
          const int tap = mode == LM2A_TAPS_K4S2 ? 2 * j + hf : j;
          on_b(p.dbg_noshift ? 0 : j, kb_base + tap * cblk + cb);
        }
      }
    }
    kb_base += cblk * ntaps_all;
  }
}

__device__ __forceinline__ uint32_t a_box_bytes(int mode, int share_taps) {
  const int rows =
      (mode == LM2A_TAPS_K1 || !share_taps) ? 128 : (mode == LM2A_TAPS_K3 ? 130 : 129);
  return (uint32_t)rows * kBlockK * 2;
}

template <int BLOCK_N, int CG, int XF>
__global__ void __launch_bounds__(num_threads(XF != 0, CG), 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0,
                 const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut, const ConvArgs p) {
  using L = SmemLayout<BLOCK_N, CG, XF>;
  constexpr int NA = L::kAStages, NB = L::kBStages;
  constexpr int kFirstEpiWarp = first_epi_warp(XF != 0, CG);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));  // generic view of smem_base
  const uint32_t bar_base = smem_base + L::kBarOffset;
  // barrier slots (8 B each)
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_ready = [&](int s) { return bar_base + 8u * (NA + s); };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 * NA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (3 * NA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (3 * NA + NB + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (3 * NA + 2 * NB + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (3 * NA + 2 * NB + 2 + s); };
  auto x_done = [&](int s) { return bar_base + 8u * (3 * NA + 2 * NB + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * L::kNumBars;
  auto a_slot = [&](int s) { return smem_base + L::kAOffset + s * L::kASlot; };
  auto b_slot = [&](int s) { return smem_base + L::kBOffset + s * L::kBSlotBytes; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = CG == 2 ? (int)blockIdx.x >> 1 : (int)blockIdx.x;       // tile-walking unit
  const int num_units = CG == 2 ? (int)gridDim.x >> 1 : (int)gridDim.x;
  // XF launches hand every A block to the transform warps (a_full -> transform / pass-through
  // -> a_ready); plain launches let the TMA complete straight on the barrier the MMA waits on
  constexpr bool xform = XF == 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < NA; ++s) {
      mbar_init(a_full(s), 1);
      // single CTA: the four transform warps arrive; pair: one hand-off thread per CTA
      mbar_init(a_ready(s), CG == 2 ? 2 : kXformWarps);
      mbar_init(x_done(s), kXformWarps);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 32 * kEpiWarps * CG);
    }
    mbar_fence_init();
  }
  constexpr uint32_t tmem_cols = 2u * BLOCK_N;
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_slot, tmem_cols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, tmem_cols);
      tmem_relinquish();
    }
  }
  if (warp >= kFirstXformWarp && warp < kFirstXformWarp + kXformWarps && xform) {
    // gamma / beta of the input GroupNorm (parameters: never written by a kernel of the stream)
    float* sg = reinterpret_cast<float*>(smem_gen + L::kGammaOffset);
    const int gn_c = p.gn_groups * p.gn_cg;
    for (int i = threadIdx.x - 32 * kFirstXformWarp; i < gn_c; i += 32 * kXformWarps) {
      sg[i] = __ldg(p.gn_gamma + i);
      sg[kMaxGnChannels + i] = __ldg(p.gn_beta + i);
    }
  }
  tc_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // Everything above overlapped the previous kernel's tail. The weights are never written by
  // a kernel, so the producer also starts the W boxes of the first pipeline stages before
  // waiting for the previous kernel (their HBM latency hides under its tail); activations
  // (A boxes, residual, FiLM table, statistics) are only touched after griddepcontrol.wait.
  int w_prefetched = 0;
  if (warp == 0 && lane == 0 && unit < total_tiles) {
    const int n0 = (unit % p.n_tiles) * BLOCK_N + cta_rank * (BLOCK_N / CG);
    walk_tile(p, [](int, int, int, int) {},
              [&](int, int kb) {
                if (w_prefetched >= NB) return;
                const int s = w_prefetched++;
                // plain launches: the stage's barrier also covers its A box (128 rows), which
                // follows after griddepcontrol.wait
                const uint32_t bytes = L::kBSlotBytes + (XF ? 0 : kATileBytes);
                if (CG == 2) {
                  if (cta_rank == 0) mbar_expect_tx(b_full(s), 2 * bytes);
                  tma_load_2d_cg2(b_slot(s), &tmB, kb * kBlockK, n0, mapa_shared(b_full(s), 0));
                } else {
                  mbar_expect_tx(b_full(s), bytes);
                  tma_load_2d(b_slot(s), &tmB, kb * kBlockK, n0, b_full(s));
                }
              });
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      int b_issued = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        const int m0 = (tile / p.n_tiles) * (kBlockM * CG) + cta_rank * kBlockM;
        const int n0 = (tile % p.n_tiles) * BLOCK_N + cta_rank * (BLOCK_N / CG);
        const CUtensorMap* cur_tm = &tmA0;
        int cur_chan = 0, cur_row = 0;
        walk_tile(
            p,
            [&](int seg, int, int row0, int chan) {
              const CUtensorMap* tm = seg ? &tmA1 : &tmA0;
              if (!XF) {   // plain: the A box travels with its W box (below)
                cur_tm = tm;
                cur_chan = chan;
                cur_row = m0 + row0;
                return;
              }
              mbar_wait(a_empty(sa), pa ^ 1u);
              if (XF == 2 && seg == 0) {
                // the transform warps build this block themselves: the slot is theirs now
                mbar_arrive(a_full(sa));
              } else {
                mbar_expect_tx(a_full(sa), a_box_bytes(p.seg_taps[seg], 1));
                tma_load_2d(a_slot(sa), tm, chan, m0 + row0, a_full(sa));
              }
              if (++sa == NA) {
                sa = 0;
                pa ^= 1u;
              }
            },
            [&](int, int kb) {
              // first tile: the W boxes of the first stages are already in flight (see above)
              const bool w_done = b_issued < w_prefetched;
              if (!w_done) mbar_wait(b_empty(sb), pb ^ 1u);
              const uint32_t bytes = L::kBSlotBytes + (XF ? 0 : kATileBytes);
              if (CG == 2) {
                // both CTAs' boxes complete on the LEADER's barrier; only it arms the count
                const uint32_t fb = mapa_shared(b_full(sb), 0);
                if (cta_rank == 0 && !w_done) mbar_expect_tx(b_full(sb), 2 * bytes);
                if (!XF) tma_load_2d_cg2(a_slot(sb), cur_tm, cur_chan, cur_row, fb);
                if (!w_done) tma_load_2d_cg2(b_slot(sb), &tmB, kb * kBlockK, n0, fb);
              } else {
                if (!w_done) mbar_expect_tx(b_full(sb), bytes);
                if (!XF) tma_load_2d(a_slot(sb), cur_tm, cur_chan, cur_row, b_full(sb));
                if (!w_done) tma_load_2d(b_slot(sb), &tmB, kb * kBlockK, n0, b_full(sb));
              }
              ++b_issued;
              if (++sb == NB) {
                sb = 0;
                pb ^= 1u;
              }
            });
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer (pair: the leader CTA only)
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, BLOCK_N);
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, acc = 0, acc_phase = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint32_t started = 0;   // 0 until the first MMA of the tile (which overwrites D)
        if constexpr (XF == 0) {
          // Plain launches: every stage holds one A box and its W box, consumed strictly in
          // ring order - the issuing thread does not need to know segments or taps. Its own
          // instruction stream sits between two barrier hand-offs of every K block, so the loop
          // is as short as it gets: descriptors advance by a constant per stage.
          constexpr uint64_t kStageStep = (uint64_t)(L::kASlot >> 4);
          constexpr uint64_t kBStageStep = (uint64_t)(L::kBSlotBytes >> 4);
          const uint64_t adesc0 = umma_desc_sw128_kmajor(a_slot(0));
          const uint64_t bdesc0 = umma_desc_sw128_kmajor(b_slot(0));
#pragma unroll 1
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(b_full(sb), pb);
            tc_fence_after_sync();
            const uint64_t adesc = adesc0 + (uint64_t)sb * kStageStep;
            const uint64_t bdesc = bdesc0 + (uint64_t)sb * kBStageStep;
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              if (CG == 2)
                umma_bf16_ss_cg2(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, started);
              else
                umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, started);
              started = 1u;
            }
            if (CG == 2) umma_commit_cg2(b_empty(sb)); else umma_commit(b_empty(sb));
            if (++sb == NB) {
              sb = 0;
              pb ^= 1u;
            }
          }
          if (CG == 2) umma_commit_cg2(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
          continue;
        }
        uint32_t cur_a = 0;
        bool a_open = false;
        walk_tile(
            p,
            [&](int, int, int, int) {
              if (!XF) return;   // plain: the A box arrives with its W box
              if (a_open) {  // every tap of the previous A block has been issued: free its slot
                if (CG == 2) umma_commit_cg2(a_empty(cur_a)); else umma_commit(a_empty(cur_a));
              }
              if (CG == 2) mbar_wait_cluster(a_ready(sa), pa);
              else mbar_wait(a_ready(sa), pa);
              tc_fence_after_sync();
              cur_a = sa;
              a_open = true;
              if (++sa == NA) {
                sa = 0;
                pa ^= 1u;
              }
            },
            [&](int shift, int) {
              mbar_wait(b_full(sb), pb);
              tc_fence_after_sync();
              const uint64_t adesc = XF ? umma_desc_sw128_rows(a_slot(cur_a), (uint32_t)shift)
                                        : umma_desc_sw128_kmajor(a_slot(sb));
              const uint64_t bdesc = umma_desc_sw128_kmajor(b_slot(sb));
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                // +32 B per K=16 step inside the 128 B swizzle row (address field is >>4)
                if (CG == 2)
                  umma_bf16_ss_cg2(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, started);
                else
                  umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, started);
                started = 1u;
              }
              if (CG == 2) umma_commit_cg2(b_empty(sb)); else umma_commit(b_empty(sb));
              if (++sb == NB) {
                sb = 0;
                pb ^= 1u;
              }
            });
        if (a_open) {
          if (CG == 2) umma_commit_cg2(a_empty(cur_a)); else umma_commit(a_empty(cur_a));
        }
        if (CG == 2) umma_commit_cg2(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= kFirstXformWarp && warp < kFirstXformWarp + kXformWarps && XF) {
    // ------------------------------------------------- operand transform (128 threads)
    // One warp per scheduler and no other warp to hide its latencies behind: the loop is laid
    // out for instruction-level parallelism. A thread owns one 16-byte chunk (8 channels) of
    // the rows rl, rl + 16, ... of every A block of the tile; which of those rows are real
    // slots, and of which clip, is the same for all blocks of a tile and lives in registers;
    // rows are processed three at a time with their loads issued up front.
    const int xt = threadIdx.x - 32 * kFirstXformWarp;
    const int chunk = xt & 7;    // 16-byte chunk (8 channels) of a 128-byte operand row
    const int rl = xt >> 3;      // row lane: rows rl, rl + kRowStride, ...
    const float* sg = reinterpret_cast<const float*>(smem_gen + L::kGammaOffset);
    float2* mr = reinterpret_cast<float2*>(smem_gen + L::kMrOffset);
    int* row_info = reinterpret_cast<int*>(smem_gen + L::kRowInfoOffset);
    const int mode0 = p.seg_taps[0];
    // XF launches always share taps: segment 0 is one 130-slot (k3) / 128-slot (k1) block
    const int rows0 = mode0 == LM2A_TAPS_K3 ? 130 : 128;
    const int roff0 = mode0 == LM2A_TAPS_K3 ? -1 : 0;
    // byte offset of this thread's chunk inside row rl (rows rl + 16 k share its swizzle phase)
    const uint32_t thr_off = (uint32_t)rl * 128u + (((uint32_t)(chunk ^ (rl & 7))) << 4);
    if constexpr (XF == 2) {
      // ---- x2 linear upsampling, align_corners = True (F.interpolate of reference
      // unet1d_ultimate.py:231-236) of the low-resolution slab straight into the operand block:
      // out slot t of a clip reads low slots i0 = floor(t * (T-1)/(2T-1)) and min(i0+1, T-1) with
      // weights (1 - w, w); slots t >= 2T stay zero (F.pad, :409-413). The arithmetic is
      // upsample2x_kernel's, so fusing does not change a bit.
      const __nv_bfloat16* src = p.up_src + chunk * 8;
      const int t_in = p.up_t_in;
      const float scale = p.t_valid > 1 ? (float)(t_in - 1) / (float)(p.t_valid - 1) : 0.f;
      uint32_t sa = 0, pa = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        const int m_base = (tile / p.n_tiles) * (kBlockM * CG) + cta_rank * kBlockM + roff0;
        // per row of this thread: element offsets of its two source slots, the lerp weight and
        // the store mask (same for every channel block of the tile)
        uint32_t off0[kRowsPerThread], off1[kRowsPerThread], keep[kRowsPerThread];
        float w1[kRowsPerThread];
#pragma unroll
        for (int k = 0; k < kRowsPerThread; ++k) {
          const long long m = (long long)m_base + rl + kRowStride * k;
          keep[k] = 0u;
          off0[k] = off1[k] = 0u;
          w1[k] = 0.f;
          if (rl + kRowStride * k < rows0 && m >= 0 && m < p.m) {
            const int r = (int)m / p.tp;
            const int t = (int)m - r * p.tp;
            if (t < p.t_valid) {
              const float sp = scale * (float)t;
              const int i0 = (int)sp;
              const int i1 = i0 + (i0 < t_in - 1 ? 1 : 0);
              w1[k] = sp - (float)i0;
              off0[k] = (uint32_t)((r * p.up_tp_in + i0) * p.up_ld);
              off1[k] = (uint32_t)((r * p.up_tp_in + i1) * p.up_ld);
              keep[k] = 0xffffffffu;
            }
          }
        }
        walk_tile(
            p,
            [&](int seg, int cb, int, int) {
              mbar_wait(a_full(sa), pa);
              if (seg == 0) {
                const uint32_t base = a_slot(sa) + thr_off;
                const __nv_bfloat16* sc = src + cb * kBlockK;
                uint4 qa[kRowsPerThread], qb[kRowsPerThread];
#pragma unroll
                for (int k = 0; k < kRowsPerThread; ++k) {
                  qa[k] = __ldg(reinterpret_cast<const uint4*>(sc + off0[k]));
                  qb[k] = __ldg(reinterpret_cast<const uint4*>(sc + off1[k]));
                }
#pragma unroll
                for (int k = 0; k < kRowsPerThread; ++k) {
                  if (k == kRowsPerThread - 1 && rl >= kASlotRows - 128) continue;
                  const float wb = w1[k], wa = 1.0f - wb;
                  const uint32_t aw[4] = {qa[k].x, qa[k].y, qa[k].z, qa[k].w};
                  const uint32_t bw[4] = {qb[k].x, qb[k].y, qb[k].z, qb[k].w};
                  uint32_t ow[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 fa = unpack_bf16x2(aw[e]);
                    const float2 fb = unpack_bf16x2(bw[e]);
                    ow[e] = pack_bf16x2(__fmaf_rn(wb, fb.x, __fmul_rn(wa, fa.x)),
                                        __fmaf_rn(wb, fb.y, __fmul_rn(wa, fa.y))) & keep[k];
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                                   base + (uint32_t)k * (uint32_t)(kRowStride * 128)),
                               "r"(ow[0]), "r"(ow[1]), "r"(ow[2]), "r"(ow[3])
                               : "memory");
                }
                fence_proxy_async_smem();
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(CG == 2 ? x_done(sa) : a_ready(sa));
              if (++sa == NA) {
                sa = 0;
                pa ^= 1u;
              }
            },
            [](int, int) {});
      }
    } else {
    uint32_t sa = 0, pa = 0;
    for (int tile = unit; tile < total_tiles; tile += num_units) {
      const int m_base = (tile / p.n_tiles) * (kBlockM * CG) + cta_rank * kBlockM + roff0;
      // clip-row of every slot of the block and (mean, rstd) of every (clip-row, group) it
      // touches, from the producer's exact integer sums
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kXformWarps) : "memory");
      const int m_lo = m_base > 0 ? m_base : 0;
      const long long m_hi_ll = (long long)m_base + rows0 - 1;
      const int m_hi = m_hi_ll < p.m - 1 ? (int)m_hi_ll : (int)(p.m - 1);
      const int r_first = m_lo / p.tp;
      const int r_last = m_hi >= m_lo ? m_hi / p.tp : r_first - 1;
      for (int i = xt; i < kASlotRows; i += 32 * kXformWarps) {
        const long long m = (long long)m_base + i;
        int info = -1;
        if (i < rows0 && m >= 0 && m < p.m) {
          const int r = (int)m / p.tp;
          const int t = (int)m - r * p.tp;
          if (t < p.t_valid) info = r - r_first;
        }
        row_info[i] = info;
      }
      const int nent = (r_last - r_first + 1) * p.gn_groups;
      const double inv_n = 1.0 / ((double)p.gn_cg * (double)p.t_valid);
      for (int e = xt; e < nent; e += 32 * kXformWarps) {
        const int rr = r_first + e / p.gn_groups, g = e % p.gn_groups;
        const long long* sp = p.gn_stats + ((size_t)rr * p.gn_pitch + g) * 2;
        // (rstd, -mean * rstd): u = x * rstd - mean * rstd is one FMA per element
        mr[e] = gn_rstd_cm(__ldcg(sp), __ldcg(sp + 1), inv_n, p.gn_eps);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kXformWarps) : "memory");
      // per row of this thread: byte offset of its clip's (rstd, -mean rstd) table row and an
      // all-ones / all-zeros store mask (pad slots and rows outside the slab must stay zero:
      // they are the conv padding). Branch-free below: every row runs the same instructions.
      uint32_t mr_off[kRowsPerThread], keep[kRowsPerThread];
#pragma unroll
      for (int k = 0; k < kRowsPerThread; ++k) {
        const int i = rl + kRowStride * k;
        const int inf = i < kASlotRows ? row_info[i] : -1;
        keep[k] = inf >= 0 ? 0xffffffffu : 0u;
        mr_off[k] = (uint32_t)((inf > 0 ? inf : 0) * p.gn_groups) * 8u;
      }
      const uint32_t mr_base = smem_base + L::kMrOffset;
      walk_tile(
          p,
          [&](int seg, int cb, int, int) {
            mbar_wait(a_full(sa), pa);
            if (seg == 0 && !p.dbg_noxform) {
              const uint32_t base = a_slot(sa) + thr_off;
              const int c0 = cb * kBlockK + chunk * 8;
              const uint32_t g = ((uint32_t)c0 * (uint32_t)p.gn_cg_magic) >> 22;   // c0 / gn_cg
              // y = SiLU(v), v = u * gamma + beta, u = x * rstd - mean * rstd, evaluated as
              // vh = u * (gamma/2) + beta/2, y = vh * tanh(vh) + vh (the halves fold exactly)
              float ga[8], be[8];
#pragma unroll
              for (int e = 0; e < 8; e += 4) {
                const float4 g4 = *reinterpret_cast<const float4*>(sg + c0 + e);
                const float4 b4 = *reinterpret_cast<const float4*>(sg + kMaxGnChannels + c0 + e);
                ga[e] = g4.x * 0.5f; ga[e + 1] = g4.y * 0.5f; ga[e + 2] = g4.z * 0.5f; ga[e + 3] = g4.w * 0.5f;
                be[e] = b4.x * 0.5f; be[e + 1] = b4.y * 0.5f; be[e + 2] = b4.z * 0.5f; be[e + 3] = b4.w * 0.5f;
              }
              const uint32_t mr_g = mr_base + g * 8u;
#pragma unroll
              for (int k0 = 0; k0 < kRowsPerThread; k0 += kRowBatch) {
                uint32_t w[kRowBatch][4];
                float2 sc[kRowBatch];
#pragma unroll
                for (int j = 0; j < kRowBatch; ++j) {
                  if (k0 + j >= kRowsPerThread) continue;
                  // the last row of a thread (rl + 128) exists only for rl < 8: the slot has 136
                  // rows (warp-uniform: a warp holds four consecutive row lanes)
                  if (k0 + j == kRowsPerThread - 1 && rl >= kASlotRows - 128) continue;
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(w[j][0]), "=r"(w[j][1]), "=r"(w[j][2]), "=r"(w[j][3])
                               : "r"(base + (uint32_t)(k0 + j) * (uint32_t)(kRowStride * 128)));
                  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                               : "=f"(sc[j].x), "=f"(sc[j].y)
                               : "r"(mr_g + mr_off[k0 + j]));
                }
#pragma unroll
                for (int j = 0; j < kRowBatch; ++j) {
                  if (k0 + j >= kRowsPerThread) continue;
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 x = unpack_bf16x2(w[j][e]);
                    const float v0 = fmaf(fmaf(x.x, sc[j].x, sc[j].y), ga[2 * e], be[2 * e]);
                    const float v1 = fmaf(fmaf(x.y, sc[j].x, sc[j].y), ga[2 * e + 1], be[2 * e + 1]);
                    w[j][e] = pack_bf16x2(silu_from_half(v0), silu_from_half(v1)) & keep[k0 + j];
                  }
                }
#pragma unroll
                for (int j = 0; j < kRowBatch; ++j) {
                  if (k0 + j >= kRowsPerThread) continue;
                  if (!(k0 + j == kRowsPerThread - 1 && rl >= kASlotRows - 128))
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                                     base + (uint32_t)(k0 + j) * (uint32_t)(kRowStride * 128)),
                                 "r"(w[j][0]), "r"(w[j][1]), "r"(w[j][2]), "r"(w[j][3])
                                 : "memory");
                }
              }
              fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(CG == 2 ? x_done(sa) : a_ready(sa));
            if (++sa == NA) {
              sa = 0;
              pa ^= 1u;
            }
          },
          [](int, int) {});
    }
    }  // XF == 1
  } else if (warp == kFirstXformWarp + kXformWarps && XF && CG == 2) {
    // --------------------------------------------------------- pair: hand-off warp
    // The leader's MMA reads both CTAs' transformed blocks, so each CTA has to release its
    // block at cluster scope - a fence that costs several hundred cycles. One otherwise idle
    // thread pays it, off the transform warps' critical path: they arrive on a CTA-local
    // barrier (cheap), this thread forwards the arrival to the leader's a_ready.
    if (lane == 0) {
      uint32_t sa = 0, pa = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        walk_tile(
            p,
            [&](int, int, int, int) {
              mbar_wait(x_done(sa), pa);
              mbar_arrive_cluster(mapa_shared(a_ready(sa), 0));
              if (++sa == NA) {
                sa = 0;
                pa ^= 1u;
              }
            },
            [](int, int) {});
      }
    }
  } else if (warp < kFirstEpiWarp) {
    // (plain launches have no such warps)
  } else {
    // ----------------------------------------------------------------- epilogue
    const int ew = warp - kFirstEpiWarp;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int half = ew >> 2;
    const int row = quad * 32 + lane;
    const uint32_t stg = smem_base + L::kOutOffset + ew * L::kOutBytesPerWarp;
    const uint32_t tab_base = smem_base + L::kTabOffset + ew * L::kTabBytesPerWarp;
    const bool film_tab = p.film != nullptr && p.film_ld == 0;
    const bool film_row = p.film != nullptr && p.film_ld != 0;
    constexpr int kHalfN = BLOCK_N / 2;  // columns handled by this warp
    // channel granularity of the GroupNorm sums inside a 32-column chunk
    const int gmask = p.stats_cg | p.stats_c0;
    const int gran = gmask % 32 == 0 ? 32 : (gmask % 16 == 0 ? 16 : 8);

    uint32_t acc = 0, acc_phase = 0;
    for (int tile = unit; tile < total_tiles; tile += num_units) {
      const int m_tile0 = (tile / p.n_tiles) * (kBlockM * CG) + cta_rank * kBlockM;
      const long long m = (long long)m_tile0 + row;
      const int n0 = (tile % p.n_tiles) * BLOCK_N;
      const bool in_range = m < p.m;
      const int r = in_range ? (int)((int)m / p.tp) : 0;
      const int t = in_range ? (int)m - r * p.tp : 0;
      const bool valid = in_range && t < p.t_valid;
      // clip-rows touched by this warp's 32 slots (GroupNorm sums are per clip-row)
      const int m_first = m_tile0 + quad * 32;
      const int m_last = m_first + 31 < (int)p.m - 1 ? m_first + 31 : (int)p.m - 1;
      const int r_lo = m_first / p.tp;
      const int r_hi = m_first < (int)p.m ? m_last / p.tp : r_lo - 1;

      // per-column (a, b) with out = acc * a + b: bias and the FiLM modulation
      // (acc + bias) * (1 + scale) + shift folded; built while the MMAs of this tile still run.
      // (per-row FiLM tables, film_ld != 0, are applied per lane further down)
      {
        __syncwarp();  // previous tile's table reads are done
        const int cl = lane * 4;                        // column inside this warp's half
        const int ncol = n0 + half * kHalfN + cl;
        if (cl < kHalfN) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol));
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), h4 = s4;
          if (film_tab) {
            s4 = __ldg(reinterpret_cast<const float4*>(p.film + ncol));
            h4 = __ldg(reinterpret_cast<const float4*>(p.film + p.film_shift_off + ncol));
          }
          const uint32_t ta = tab_base + (uint32_t)cl * 8u;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "f"(1.f + s4.x),
                       "f"(fmaf(b4.x, 1.f + s4.x, h4.x)), "f"(1.f + s4.y),
                       "f"(fmaf(b4.y, 1.f + s4.y, h4.y))
                       : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta + 16u),
                       "f"(1.f + s4.z), "f"(fmaf(b4.z, 1.f + s4.z, h4.z)), "f"(1.f + s4.w),
                       "f"(fmaf(b4.w, 1.f + s4.w, h4.w))
                       : "memory");
        }
        __syncwarp();
      }

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(quad * 32) << 16);

      // GroupNorm sums of the output: per-lane exact integer sums of the group being walked
      int st_g = -1;
      long long st_a1 = 0, st_a2 = 0;
      auto flush_stats = [&]() {
        if (st_g >= 0) {
          for (int rr = r_lo; rr <= r_hi; ++rr) {
            const bool mine = in_range && r == rr;
            const long long a = warp_sum_ll(mine ? st_a1 : 0ll);
            const long long b = warp_sum_ll(mine ? st_a2 : 0ll);
            if (lane == 0 && p.dbg_stats != 1) {
              unsigned long long* sp = p.stats + ((size_t)rr * p.stats_pitch + st_g) * 2;
              atomicAdd(sp, (unsigned long long)a);
              atomicAdd(sp + 1, (unsigned long long)b);
            }
          }
        }
        st_a1 = st_a2 = 0;
      };
#pragma unroll 1
      for (int c0 = half * kHalfN; c0 < (half + 1) * kHalfN; c0 += 32) {
        const int n = n0 + c0;
        if (n >= p.n_valid) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        float f[32];
        const uint32_t tcol = tab_base + (uint32_t)(c0 - half * kHalfN) * 8u;
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float4 ab;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(ab.x), "=f"(ab.y), "=f"(ab.z), "=f"(ab.w)
                       : "r"(tcol + 16u * j));
          f[2 * j] = fmaf(__uint_as_float(v[2 * j]), ab.x, ab.y);
          f[2 * j + 1] = fmaf(__uint_as_float(v[2 * j + 1]), ab.z, ab.w);
        }
        if (film_row) {
          const float* sc = p.film + (size_t)r * p.film_ld + n;
          const float* sh = sc + p.film_shift_off;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc + j));
            const float4 h4 = __ldg(reinterpret_cast<const float4*>(sh + j));
            f[j + 0] = fmaf(f[j + 0], 1.0f + s4.x, h4.x);
            f[j + 1] = fmaf(f[j + 1], 1.0f + s4.y, h4.y);
            f[j + 2] = fmaf(f[j + 2], 1.0f + s4.z, h4.z);
            f[j + 3] = fmaf(f[j + 3], 1.0f + s4.w, h4.w);
          }
        }
        if (p.residual != nullptr && valid) {
          const uint4* rp =
              reinterpret_cast<const uint4*>(p.residual + (size_t)m * p.res_ld + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 q = __ldg(rp + j);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = unpack_bf16x2(w[e]);
              f[j * 8 + e * 2 + 0] += a.x;
              f[j * 8 + e * 2 + 1] += a.y;
            }
          }
        }
        if (p.stats != nullptr && p.dbg_stats == 2) {
          st_g = (p.stats_c0 + n) / p.stats_cg;
          st_a1 += 1;
        } else if (p.stats != nullptr) {
          // Sum / sum of squares of this lane's slot over each `gran`-channel sub-block (fp32,
          // fixed order: independent of where the slot sits in the batch), converted to fixed
          // point and added - exactly, as integers - to the lane's running sums of the current
          // group; a group is flushed (warp reduction + one atomic pair per clip-row) when the
          // walk over the warp's columns leaves it.
          // (a lane is one slot: pad slots and slots past the slab hold finite values computed
          // from zero / real operands, so they are summed like the others and dropped as a whole)
          float q1[4], q2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = f[q * 8 + j];
              a += x;
              b = fmaf(x, x, b);
            }
            q1[q] = valid ? a : 0.f;
            q2[q] = valid ? b : 0.f;
          }
          const int nsub = 32 / gran;  // sub-blocks inside this 32-channel chunk
          if (nsub == 1) {
            q1[0] = (q1[0] + q1[1]) + (q1[2] + q1[3]);
            q2[0] = (q2[0] + q2[1]) + (q2[2] + q2[3]);
          } else if (nsub == 2) {
            q1[0] += q1[1];
            q2[0] += q2[1];
            q1[1] = q1[2] + q1[3];
            q2[1] = q2[2] + q2[3];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q < nsub) {
              const int g = (p.stats_c0 + n + q * gran) / p.stats_cg;
              if (g != st_g) {   // warp-uniform
                flush_stats();
                st_g = g;
              }
              st_a1 += __float2ll_rn(q1[q] * kStatScale1);
              st_a2 += __float2ll_rn(q2[q] * kStatScale2);
            }
          }
        }
        if (p.out_mode == LM2A_OUT_BF16_SLAB) {
          // bf16 -> this warp's [32 slots][32 channels] staging tile (64B swizzle) -> one TMA
          // store; rows past the slab and channels past n_valid are clipped by the tensor map,
          // pad slots (t >= t_valid) are written as zeros to keep the conv padding intact.
          if (lane == 0) tma_store_wait_read<0>();  // the previous box has left the tile
          __syncwarp();
          const uint32_t srow = stg + (uint32_t)lane * 64u;
          const uint32_t sx = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t q0 = 0u, q1w = 0u, q2w = 0u, q3 = 0u;
            if (valid) {
              q0 = pack_bf16x2(f[j * 8 + 0], f[j * 8 + 1]);
              q1w = pack_bf16x2(f[j * 8 + 2], f[j * 8 + 3]);
              q2w = pack_bf16x2(f[j * 8 + 4], f[j * 8 + 5]);
              q3 = pack_bf16x2(f[j * 8 + 6], f[j * 8 + 7]);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                             srow + (((uint32_t)j ^ sx) << 4)),
                         "r"(q0), "r"(q1w), "r"(q2w), "r"(q3)
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && m_first < (int)p.m) {
            tma_store_2d(&tmOut, stg, n, m_first);
            tma_store_commit();
          }
        } else {
          // fp32 [R, n_valid, t_valid]: lanes are consecutive t -> coalesced per channel
          if (valid) {
            float* op = reinterpret_cast<float*>(p.out) +
                        ((size_t)r * p.n_valid + n) * p.t_valid + t;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (n + j < p.n_valid) op[(size_t)j * p.t_valid] = f[j];
            }
          }
        }
      }
      tc_fence_before_sync();
      if (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
      else mbar_arrive(tempty_bar(acc));
      flush_stats();   // after the accumulator is handed back: off the next tile's critical path
      acc ^= 1u;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) tma_store_wait<0>();  // all boxes written before the grid completes
  }

  tc_fence_before_sync();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------ host side
// bf16 output slab as seen by the epilogue's TMA stores: [32 channels x 32 slots] boxes, 64B swizzle
int encode_2d_out(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                  uint64_t pitch_elems) {
  EncodeTiledFn fn = get_encode_fn();
  LM2A_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult res = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LM2A_REQUIRE(res == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled (output) failed (%d): base=%p inner=%llu outer=%llu "
               "pitch=%llu", (int)res, base, (unsigned long long)inner,
               (unsigned long long)outer, (unsigned long long)pitch_elems);
  return 0;
}

// 2-D bf16 operand map: inner = channels (box 64 -> 128 B, SWIZZLE_128B), outer = slots / rows.
int encode_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
              uint64_t pitch_elems, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  LM2A_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult res = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LM2A_REQUIRE(res == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu pitch=%llu",
               (int)res, base, (unsigned long long)inner, (unsigned long long)outer,
               (unsigned long long)pitch_elems);
  return 0;
}

template <int BLOCK_N, int CG, int XF>
int launch(cudaStream_t stream, const CUtensorMap& a0, const CUtensorMap& a1,
           const CUtensorMap& b, const CUtensorMap& o, const ConvArgs& args) {
  using L = SmemLayout<BLOCK_N, CG, XF>;
  auto kern = conv_gemm_kernel<BLOCK_N, CG, XF>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    LM2A_CUDA_OK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
  const int tiles = args.m_tiles * args.n_tiles;
  const int units = num_sms() / CG;  // CTAs (CG = 1) or CTA pairs (CG = 2) that fit the chip
  const int grid = (tiles < units ? tiles : units) * CG;
  LM2A_CUDA_OK(launch_kernel_cluster(kern, dim3(grid), dim3(num_threads(XF != 0, CG)), L::kBytes, stream,
                                     (unsigned)CG, a0, a1, b, o, args));
  count_launch();
  return 0;
}

// Tile shape choice by a wave model. Per 64-deep K block a CTA needs max(MMA cycles, operand
// bytes / ~64 B per clock an SM pulls from L2):
//   1 CTA,  128 x 256: max(512, 48 KB / 64) = 768      1 CTA,  128 x 128: max(256, 512) = 512
//   pair,   256 x 256: max(512, 32 KB / 64) = 512      pair,   256 x 128: max(256, 384) = 384
struct TileChoice {
  int block_n, cg;
};
TileChoice choose_tile(long long m, int n_pad) {
  const long long sms = num_sms();
  TileChoice best{128, 1};
  long long best_cost = -1;
  const int bns[2] = {256, 128};
  for (int cg = 2; cg >= 1; --cg) {
    for (int bi = 0; bi < 2; ++bi) {
      const int bn = bns[bi];
      if (n_pad % bn != 0) continue;
      const long long tiles = ((m + 128 * cg - 1) / (128 * cg)) * (n_pad / bn);
      const long long units = sms / cg;
      const long long waves = (tiles + units - 1) / units;
      const long long per_kb = cg == 2 ? (bn == 256 ? 512 : 384) : (bn == 256 ? 768 : 512);
      const long long cost = waves * per_kb;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best = TileChoice{bn, cg};
      }
    }
  }
  return best;
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_conv1d_bf16(void* stream, const lm2a_conv_desc* d) {
  using namespace lm2a;
  LM2A_REQUIRE(d != nullptr, "conv1d: null descriptor");
  LM2A_REQUIRE(d->seg[0].x != nullptr && d->w != nullptr && d->bias != nullptr &&
                   d->out != nullptr,
               "conv1d: null x / w / bias / out pointer");
  LM2A_REQUIRE(d->m > 0 && d->tp > 0 && d->t_valid > 0 && d->t_valid <= d->tp &&
                   d->m % d->tp == 0,
               "conv1d: bad slab geometry m=%lld tp=%d t_valid=%d", (long long)d->m, d->tp,
               d->t_valid);
  LM2A_REQUIRE(d->n_pad > 0 && d->n_pad % 128 == 0 && d->n_valid > 0 && d->n_valid <= d->n_pad,
               "conv1d: n_pad=%d must be a positive multiple of 128 (n_valid=%d)", d->n_pad,
               d->n_valid);
  LM2A_REQUIRE(d->m < (1ll << 31) - 512, "conv1d: too many slots");

  ConvArgs a{};
  CUtensorMap tmA[2];
  int k_total = 0;
  // Launches that normalise their operand on the fly stage ONE A block per (segment, 64
  // channels) and serve every tap through row-shifted views of it (the transform then runs once
  // per element, not once per tap). Plain launches are MMA-bound, not load-bound: measured on
  // B200 at the production shapes, one 128-slot box per tap riding in the same stage as its W
  // box (one barrier wait + one commit per K block for the MMA-issuing thread) is equal or
  // faster than shared blocks with split rings (50.6 vs 53.7 us at M 8320 x N 1024 x K 3584).
  static const int noshift_env = [] {
    const char* e = getenv("LM2A_CONV_DBG_NOSHIFT");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  const bool up2x = d->in_up_t > 0;
  a.share_taps = (d->in_gn_stats != nullptr || up2x) ? 1 : 0;
  a.dbg_noshift = noshift_env;
  static const int dbg_stats_env = [] {
    const char* e = getenv("LM2A_CONV_DBG_STATS");
    return e != nullptr ? atoi(e) : 0;
  }();
  a.dbg_stats = dbg_stats_env;
  // K walk of plain launches: taps outer (default; measured 5-10 % faster on B200 at the
  // production shapes: W is read contiguously along K and the MMA-issuing thread's loop is
  // shorter) or channel blocks outer (k_order = 1: the order the operand-transform launches
  // use, for bit-for-bit comparisons against them). LM2A_CONV_TAP_OUTER=0 forces the latter.
  static const int tap_outer_env = [] {
    const char* e = getenv("LM2A_CONV_TAP_OUTER");
    return (e != nullptr && e[0] == '0') ? 0 : 1;
  }();
  LM2A_REQUIRE(d->k_order == 0 || d->k_order == 1, "conv1d: k_order=%d (0 or 1)", d->k_order);
  a.tap_outer = d->k_order == 0 ? tap_outer_env : 0;
  static const int noxform_env = [] {
    const char* e = getenv("LM2A_CONV_DBG_NOXFORM");
    return (e != nullptr && e[0] == '1') ? 1 : 0;
  }();
  a.dbg_noxform = noxform_env;
  for (int s = 0; s < 2; ++s) {
    const lm2a_conv_seg& g = d->seg[s];
    if (g.x == nullptr) {
      a.seg_cblk[s] = 0;
      a.seg_taps[s] = 0;
      a.seg_half[s] = 0;
      tmA[s] = tmA[0];
      continue;
    }
    LM2A_REQUIRE(g.cin > 0 && g.cin % 64 == 0, "conv1d: seg %d cin=%d not a multiple of 64", s,
                 g.cin);
    LM2A_REQUIRE(g.ld >= g.cin && g.ld % 8 == 0, "conv1d: seg %d ld=%d invalid (cin=%d)", s,
                 g.ld, g.cin);
    LM2A_REQUIRE((reinterpret_cast<uintptr_t>(g.x) & 15) == 0,
                 "conv1d: seg %d base not 16-byte aligned", s);
    LM2A_REQUIRE(g.taps >= LM2A_TAPS_K1 && g.taps <= LM2A_TAPS_K4S2, "conv1d: seg %d taps=%d",
                 s, g.taps);
    a.seg_cblk[s] = g.cin / 64;
    a.seg_taps[s] = g.taps;
    int ntaps = 1;
    if (s == 0 && up2x) {
      // seg[0] is the LOW-resolution slab: the operand tiles are interpolated from it by the
      // transform warps (no tensor map; a placeholder keeps the kernel signature uniform)
      LM2A_REQUIRE(g.taps == LM2A_TAPS_K3 && 2 * g.rows == d->m && d->tp == 2 * d->in_up_tp &&
                       d->t_valid == 2 * d->in_up_t && d->in_up_t <= d->in_up_tp &&
                       d->in_gn_stats == nullptr,
                   "conv1d: fused x2 upsampling needs a k3 segment over the low-res slab with "
                   "m = 2 * rows, tp = 2 * in_up_tp, t_valid = 2 * in_up_t and no input GroupNorm");
      LM2A_REQUIRE(g.rows * (long long)g.ld < (1ll << 31), "conv1d: low-res slab too large");
      a.seg_half[s] = 0;
      a.up_src = reinterpret_cast<const __nv_bfloat16*>(g.x);
      a.up_ld = g.ld;
      a.up_tp_in = d->in_up_tp;
      a.up_t_in = d->in_up_t;
      k_total += 3 * g.cin;
      continue;
    }
    if (g.taps == LM2A_TAPS_K4S2) {
      LM2A_REQUIRE(g.rows == 2 * d->m,
                   "conv1d: k4s2 needs input slots (%lld) == 2 * output slots (%lld)",
                   (long long)g.rows, (long long)d->m);
      a.seg_half[s] = g.ld;
      ntaps = 4;
      // slot pairs: one row of the [rows / 2, 2 * ld] view = (even slot | odd slot); a box
      // covers the pairs t (+1 halo) of one half
      if (encode_2d(&tmA[s], g.x, (uint64_t)g.ld + g.cin, (uint64_t)g.rows / 2,
                    (uint64_t)g.ld * 2, a.share_taps ? kBlockM + 1 : kBlockM))
        return 1;
    } else {
      LM2A_REQUIRE(g.rows == d->m, "conv1d: seg %d slots (%lld) != output slots (%lld)", s,
                   (long long)g.rows, (long long)d->m);
      a.seg_half[s] = 0;
      ntaps = g.taps == LM2A_TAPS_K3 ? 3 : 1;
      if (encode_2d(&tmA[s], g.x, (uint64_t)g.cin, (uint64_t)g.rows, (uint64_t)g.ld,
                    (g.taps == LM2A_TAPS_K3 && a.share_taps) ? kBlockM + 2 : kBlockM))
        return 1;
    }
    k_total += ntaps * g.cin;
  }
  int block_n = d->block_n, cg = d->cta_group;
  if (cg == 0) {
    // LM2A_CONV_CG=1|2 pins the auto choice (A/B measurements); unset = wave model
    static const int forced = [] {
      const char* e = getenv("LM2A_CONV_CG");
      return (e != nullptr && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0;
    }();
    if (forced != 0) {
      cg = forced;
      if (block_n == 0) block_n = d->n_pad % 256 == 0 ? 256 : 128;
    }
  }
  if (block_n == 0 || cg == 0) {
    const TileChoice c = choose_tile(d->m, d->n_pad);
    if (block_n == 0 && cg == 0) {
      block_n = c.block_n;
      cg = c.cg;
    } else if (block_n == 0) {
      block_n = d->n_pad % 256 == 0 ? 256 : 128;
    } else {
      cg = c.cg;
    }
  }
  LM2A_REQUIRE((block_n == 128 || block_n == 256) && d->n_pad % block_n == 0,
               "conv1d: block_n=%d incompatible with n_pad=%d", block_n, d->n_pad);
  LM2A_REQUIRE(cg == 1 || cg == 2, "conv1d: cta_group=%d (0 = auto, 1 or 2)", cg);
  LM2A_REQUIRE((reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d->out) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d->bias) & 15) == 0,
               "conv1d: w / out / bias must be 16-byte aligned");
  CUtensorMap tmB;
  if (encode_2d(&tmB, d->w, (uint64_t)k_total, (uint64_t)d->n_pad, (uint64_t)k_total,
                (uint32_t)(block_n / cg)))
    return 1;

  if (up2x) {
    tmA[0] = tmB;
    if (d->seg[1].x == nullptr) tmA[1] = tmB;
  }
  a.num_kb = k_total / kBlockK;
  a.m_tiles = (int)((d->m + kBlockM * cg - 1) / (kBlockM * cg));
  a.n_tiles = d->n_pad / block_n;
  a.m = d->m;
  a.tp = d->tp;
  a.t_valid = d->t_valid;
  a.n_valid = d->n_valid;
  a.bias = d->bias;
  a.film = d->film;
  a.film_ld = d->film_ld;
  a.film_shift_off = d->film_shift_off;
  if (d->film != nullptr) {
    LM2A_REQUIRE((reinterpret_cast<uintptr_t>(d->film) & 15) == 0 && d->film_ld % 4 == 0 &&
                     d->film_shift_off % 4 == 0,
                 "conv1d: film table must be 16-byte aligned with ld / offset multiples of 4");
  }
  a.residual = reinterpret_cast<const __nv_bfloat16*>(d->residual);
  a.res_ld = d->res_ld;
  if (d->residual != nullptr) {
    LM2A_REQUIRE((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0 && d->res_ld % 8 == 0,
                 "conv1d: residual slab must be 16-byte aligned, ld multiple of 8");
  }
  a.out = d->out;
  a.out_ld = d->out_ld;
  a.out_mode = d->out_mode;
  a.stats = reinterpret_cast<unsigned long long*>(d->stats);
  a.stats_pitch = d->stats_pitch;
  a.stats_cg = d->stats_cg > 0 ? d->stats_cg : 32;
  a.stats_c0 = d->stats_c0;
  if (d->stats != nullptr) {
    LM2A_REQUIRE(d->out_mode == LM2A_OUT_BF16_SLAB, "conv1d: stats need a bf16 slab output");
    LM2A_REQUIRE(d->stats_cg > 0 && d->stats_cg % 8 == 0 && d->stats_pitch > 0 &&
                     d->stats_c0 >= 0 && d->stats_c0 % 8 == 0 &&
                     (d->stats_c0 + d->n_valid + d->stats_cg - 1) / d->stats_cg <= d->stats_pitch &&
                     (reinterpret_cast<uintptr_t>(d->stats) & 15) == 0,
                 "conv1d: bad stats layout (channels per group=%d, groups per row=%d, first "
                 "channel=%d, n=%d)", d->stats_cg, d->stats_pitch, d->stats_c0, d->n_valid);
  }
  a.gn_stats = reinterpret_cast<const long long*>(d->in_gn_stats);
  a.gn_gamma = d->in_gn_gamma;
  a.gn_beta = d->in_gn_beta;
  a.gn_pitch = d->in_gn_pitch;
  a.gn_groups = d->in_gn_groups;
  a.gn_eps = d->in_gn_eps;
  if (d->in_gn_stats != nullptr) {
    const lm2a_conv_seg& g = d->seg[0];
    LM2A_REQUIRE(d->in_gn_gamma != nullptr && d->in_gn_beta != nullptr && d->in_gn_groups > 0 &&
                     d->in_gn_pitch >= d->in_gn_groups && g.cin % d->in_gn_groups == 0 &&
                     (g.cin / d->in_gn_groups) % 8 == 0 && g.cin <= kMaxGnChannels,
                 "conv1d: input GroupNorm needs gamma / beta, groups dividing cin=%d into "
                 "multiples of 8 channels, cin <= %d (groups=%d pitch=%d)", g.cin,
                 kMaxGnChannels, d->in_gn_groups, d->in_gn_pitch);
    LM2A_REQUIRE(g.taps == LM2A_TAPS_K1 || g.taps == LM2A_TAPS_K3,
                 "conv1d: input GroupNorm is available for k1 / k3 segments only");
    LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(d->in_gn_stats) |
                   reinterpret_cast<uintptr_t>(d->in_gn_gamma) |
                   reinterpret_cast<uintptr_t>(d->in_gn_beta)) & 15) == 0,
                 "conv1d: input GroupNorm stats / gamma / beta must be 16-byte aligned");
    LM2A_REQUIRE(d->in_gn_silu != 0,
                 "conv1d: the operand transform is GroupNorm + SiLU (in_gn_silu = 0: use "
                 "lm2a_gn_apply_bf16 in front of the conv)");
    a.gn_cg = g.cin / d->in_gn_groups;
    a.gn_cg_magic = (unsigned int)(((1u << 22) + a.gn_cg - 1) / a.gn_cg);
    const int clip_rows = (kBlockM + 2 - 1) / d->tp + 2;   // clip-rows a 130-slot block can touch
    LM2A_REQUIRE(clip_rows * d->in_gn_groups <= kMaxGnEntries,
                 "conv1d: input GroupNorm: %d clip-rows x %d groups per tile exceed %d entries "
                 "(clips of %d slots are too short for the fused path)", clip_rows,
                 d->in_gn_groups, kMaxGnEntries, d->tp);
  }
  if (d->out_mode == LM2A_OUT_BF16_SLAB) {
    LM2A_REQUIRE(d->out_ld % 8 == 0 && d->out_ld >= d->n_valid && d->n_valid % 32 == 0,
                 "conv1d: bf16 slab output needs ld %% 8 == 0 and n_valid %% 32 == 0 (ld=%d n=%d)",
                 d->out_ld, d->n_valid);
  } else {
    LM2A_REQUIRE(d->out_mode == LM2A_OUT_F32_NCT, "conv1d: bad out_mode %d", d->out_mode);
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmOut = tmB;  // unused in the fp32 output mode
  if (d->out_mode == LM2A_OUT_BF16_SLAB &&
      encode_2d_out(&tmOut, d->out, (uint64_t)d->n_valid, (uint64_t)d->m, (uint64_t)d->out_ld))
    return 1;
  const int xf = a.gn_stats != nullptr ? 1 : (a.up_src != nullptr ? 2 : 0);
#define LM2A_CONV_LAUNCH(BN, CGV)                                                         \
  (xf == 1 ? launch<BN, CGV, 1>(st, tmA[0], tmA[1], tmB, tmOut, a)                        \
           : (xf == 2 ? launch<BN, CGV, 2>(st, tmA[0], tmA[1], tmB, tmOut, a)             \
                      : launch<BN, CGV, 0>(st, tmA[0], tmA[1], tmB, tmOut, a)))
  if (cg == 2) return block_n == 256 ? LM2A_CONV_LAUNCH(256, 2) : LM2A_CONV_LAUNCH(128, 2);
  return block_n == 256 ? LM2A_CONV_LAUNCH(256, 1) : LM2A_CONV_LAUNCH(128, 1);
#undef LM2A_CONV_LAUNCH
}
