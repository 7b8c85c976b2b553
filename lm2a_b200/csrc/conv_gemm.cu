// Implicit-GEMM Conv1d / Linear for channels-last bf16 slabs on sm_100a.
//
//   out[m, n] = sum_{seg, tap, c} A_seg[m + shift(tap), c] * W[n, k(seg,tap,c)] + bias[n]
//
// Replaces nn.Conv1d (k=1, k=3 p=1, k=4 s=2 p=1) and every nn.Linear / MHA
// projection on the sampling path (reference models/unet1d_ultimate.py:87-88,
// 115,216-221,255-261,295,364; models/cross_attention.py:19-36). The FiLM
// modulation h*(1+scale)+shift (:141-143), the ResBlock skip 1x1 conv (second K
// segment accumulating into the same TMEM tile) and the residual add (:159) are
// fused into the epilogue.
//
// Structure (one CTA per SM, persistent over 128 x BLOCK_N output tiles):
//   warp 0     TMA producer: per K-block one 128x64 A box (a tap is a row shift of
//              the flattened slab; the zero slot between clips and TMA's
//              out-of-bounds zero fill are the conv padding) and one BLOCK_Nx64 W box
//   warp 1     tcgen05.mma issuer (single thread), accumulators in TMEM,
//              double-buffered so tile i+1's MMAs overlap tile i's epilogue
//   warps 2-9  epilogue: tcgen05.ld -> bias/FiLM/residual (+ partial GroupNorm sums of the
//              output) -> bf16 slab (or the final fp32 [R, C, T] eps tensor); warp w reads
//              TMEM lane quadrant w % 4 and the column half (w - 2) / 4
#include "../../include/lm2a_b200.h"
#include <stdlib.h>

#include "common.cuh"

namespace lm2a {

#ifdef LM2A_CONV_TIMING
// Instrumented build (python -m lm2a_b200.build with LM2A_NVCC_DEFS=-DLM2A_CONV_TIMING, see
// tools/conv_stall_probe.py): cycles the MMA-issuing thread waits for operands (full barriers) /
// for a free accumulator, cycles the TMA producer waits for a free stage, and the CTA lifetime,
// summed over all CTAs. Not part of the product build.
__device__ unsigned long long g_conv_timing[8];
#endif

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;
constexpr int kEpiWarps = 8;  // two per TMEM lane quadrant, each takes half of the columns
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct ConvArgs {
  int seg_cblk[2];   // cin / 64 per segment
  int seg_taps[2];   // LM2A_TAPS_*
  int seg_half[2];   // K4S2: channel offset of the odd slot inside a slot pair
  int num_kb;
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
  float2* stats;     // partial GroupNorm sums of the output (or null), see lm2a_conv_desc
  int stats_sub;     // sub-blocks per clip-row in the stats buffer (= its row pitch)
  int stats_ns;      // slices per (clip-row, sub-block)
  int stats_gran;    // channels per sub-block: 8, 16 or 32
  // fused GroupNorm + SiLU of the output (see lm2a_conv_desc.gn_*): the CTA keeps its tiles in
  // TMEM, all CTAs meet at a grid barrier once the partial sums are written, then the tiles
  // are normalised straight out of TMEM
  const float* gn_gamma;
  const float* gn_beta;
  int gn_groups;
  float gn_eps;
  unsigned int* gn_barrier;  // {arrival count, generation}
};

template <int BLOCK_N, int STAGES, int CG>
struct SmemLayout {
  static constexpr int kBTileBytes = (BLOCK_N / CG) * kBlockK * 2;  // a CTA pair splits B along N
  static constexpr int kStageBytes = kATileBytes + kBTileBytes;
  // epilogue: per warp a [32 rows][32 cols] bf16 staging tile (TMA store source, 64B swizzle)
  // and a table of per-column (scale, offset) pairs for its BLOCK_N / 2 columns
  static constexpr int kOutOffset = STAGES * kStageBytes;
  static constexpr int kOutBytesPerWarp = 32 * 32 * 2;
  static constexpr int kTabOffset = kOutOffset + kEpiWarps * kOutBytesPerWarp;
  static constexpr int kTabBytesPerWarp = 2 * (BLOCK_N / 2) * 8;  // x2: one table per clip-row
  static constexpr int kBarOffset = kTabOffset + kEpiWarps * kTabBytesPerWarp;
  static constexpr int kBytes = kBarOffset + 256 + 1024;  // + barriers + align slack
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
  // fused GroupNorm pass 1 stages every output box of the CTA in the (then idle) pipeline area
  static_assert(kEpiWarps * (512 / BLOCK_N) * (BLOCK_N / 2 / 32) * kOutBytesPerWarp <=
                    STAGES * kStageBytes, "pass-1 staging must fit the operand pipeline");
};

// CG = 1: one CTA per 128 x BLOCK_N tile. CG = 2: a CTA pair (cluster of 2) per 256 x BLOCK_N
// tile with cta_group::2 UMMA: each CTA stages its own 128 rows of A and half of the W rows,
// so the operand bytes an SM pulls from L2 per MMA cycle drop from 96 to 64 (BLOCK_N = 256).
template <int BLOCK_N, int STAGES, int CG>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0,
                 const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmOut2, const ConvArgs p) {
  using L = SmemLayout<BLOCK_N, STAGES, CG>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + L::kBarOffset;
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tfull[4], tempty[2], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 6);
  // fused GroupNorm: every tile of this CTA keeps its own accumulator (512 / BLOCK_N of them)
  const bool fuse_gn = p.gn_gamma != nullptr;
  constexpr int kMaxAcc = 512 / BLOCK_N;
  auto a_tile = [&](int s) { return smem_base + s * L::kStageBytes; };
  auto b_tile = [&](int s) { return smem_base + s * L::kStageBytes + kATileBytes; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = CG == 2 ? (int)blockIdx.x >> 1 : (int)blockIdx.x;       // tile-walking unit
  const int num_units = CG == 2 ? (int)gridDim.x >> 1 : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmOut2);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(tfull_bar(s), 1);
    for (int s = 0; s < 2; ++s)
      mbar_init(tempty_bar(s), 32 * kEpiWarps * CG);  // pair: the leader collects both CTAs
    mbar_fence_init();
  }
  const uint32_t tmem_cols = fuse_gn ? 512u : 2u * BLOCK_N;
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_slot, tmem_cols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, tmem_cols);
      tmem_relinquish();
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
  // (A boxes, residual, FiLM table) are only touched after griddepcontrol.wait.
  int w_prefetched = 0;
  if (warp == 0 && lane == 0 && unit < total_tiles) {
    const int n0 = (unit % p.n_tiles) * BLOCK_N + cta_rank * (BLOCK_N / CG);
    w_prefetched = p.num_kb < STAGES ? p.num_kb : STAGES;
    for (int kb = 0; kb < w_prefetched; ++kb) {
      if (CG == 2) {
        if (cta_rank == 0) mbar_expect_tx(full_bar(kb), 2 * L::kStageBytes);
        tma_load_2d_cg2(b_tile(kb), &tmB, kb * kBlockK, n0, mapa_shared(full_bar(kb), 0));
      } else {
        mbar_expect_tx(full_bar(kb), L::kStageBytes);
        tma_load_2d(b_tile(kb), &tmB, kb * kBlockK, n0, full_bar(kb));
      }
    }
  }
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
#ifdef LM2A_CONV_TIMING
      long long t_prod_wait = 0;
#endif
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        const int m0 = (tile / p.n_tiles) * (kBlockM * CG) + cta_rank * kBlockM;
        const int n0 = (tile % p.n_tiles) * BLOCK_N + cta_rank * (BLOCK_N / CG);
        int kb = 0;
#pragma unroll 1
        for (int seg = 0; seg < 2; ++seg) {
          const int cblk = p.seg_cblk[seg];
          if (cblk == 0) continue;
          const CUtensorMap* tm = seg ? &tmA1 : &tmA0;
          const int mode = p.seg_taps[seg];
          const int ntaps = mode == LM2A_TAPS_K1 ? 1 : (mode == LM2A_TAPS_K3 ? 3 : 4);
#pragma unroll 1
          for (int tap = 0; tap < ntaps; ++tap) {
            int shift = 0, choff = 0;
            if (mode == LM2A_TAPS_K3) {
              shift = tap - 1;
            } else if (mode == LM2A_TAPS_K4S2) {
              // x[2t-1+tap] over slot pairs: tap0 -> pair t-1 odd half, tap1 -> pair t
              // even half, tap2 -> pair t odd half, tap3 -> pair t+1 even half
              shift = tap == 0 ? -1 : (tap == 3 ? 1 : 0);
              choff = (tap == 0 || tap == 2) ? p.seg_half[seg] : 0;
            }
#pragma unroll 1
            for (int cb = 0; cb < cblk; ++cb, ++kb) {
              // first tile: the W box of the first stages is already in flight (see above)
              const bool w_done = tile == unit && kb < w_prefetched;
#ifdef LM2A_CONV_TIMING
              const long long tw0 = clock64();
#endif
              if (!w_done) mbar_wait(empty_bar(stage), phase ^ 1u);
#ifdef LM2A_CONV_TIMING
              t_prod_wait += clock64() - tw0;
#endif
              if (CG == 2) {
                // both CTAs' boxes complete on the LEADER's barrier; only it arms the count
                const uint32_t fb = mapa_shared(full_bar(stage), 0);
                if (cta_rank == 0 && !w_done) mbar_expect_tx(full_bar(stage), 2 * L::kStageBytes);
                tma_load_2d_cg2(a_tile(stage), tm, choff + cb * kBlockK, m0 + shift, fb);
                if (!w_done) tma_load_2d_cg2(b_tile(stage), &tmB, kb * kBlockK, n0, fb);
              } else {
                if (!w_done) mbar_expect_tx(full_bar(stage), L::kStageBytes);
                tma_load_2d(a_tile(stage), tm, choff + cb * kBlockK, m0 + shift,
                            full_bar(stage));
                if (!w_done) tma_load_2d(b_tile(stage), &tmB, kb * kBlockK, n0, full_bar(stage));
              }
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
#ifdef LM2A_CONV_TIMING
      atomicAdd(&g_conv_timing[2], (unsigned long long)t_prod_wait);
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer (pair: the leader CTA only)
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM * CG, BLOCK_N);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
#ifdef LM2A_CONV_TIMING
      long long t_full_wait = 0, t_acc_wait = 0;
      const long long t_begin = clock64();
#endif
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        if (!fuse_gn) {
#ifdef LM2A_CONV_TIMING
          const long long ta0 = clock64();
#endif
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
#ifdef LM2A_CONV_TIMING
          t_acc_wait += clock64() - ta0;
#endif
          tc_fence_after_sync();
        }
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
#pragma unroll 1
        for (int kb = 0; kb < p.num_kb; ++kb) {
#ifdef LM2A_CONV_TIMING
          const long long tf0 = clock64();
#endif
          mbar_wait(full_bar(stage), phase);
#ifdef LM2A_CONV_TIMING
          t_full_wait += clock64() - tf0;
#endif
          tc_fence_after_sync();
          const uint64_t adesc = umma_desc_sw128_kmajor(a_tile(stage));
          const uint64_t bdesc = umma_desc_sw128_kmajor(b_tile(stage));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (address field is >>4)
            if (CG == 2)
              umma_bf16_ss_cg2(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc,
                               (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc,
                           (kb | k) != 0 ? 1u : 0u);
          }
          if (CG == 2) umma_commit_cg2(empty_bar(stage)); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (CG == 2) umma_commit_cg2(tfull_bar(acc)); else umma_commit(tfull_bar(acc));
        if (fuse_gn) {
          ++acc;  // one accumulator per tile, kept until the normalisation pass
        } else {
          acc ^= 1u;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
#ifdef LM2A_CONV_TIMING
      atomicAdd(&g_conv_timing[0], (unsigned long long)t_full_wait);
      atomicAdd(&g_conv_timing[1], (unsigned long long)t_acc_wait);
      atomicAdd(&g_conv_timing[3], (unsigned long long)(clock64() - t_begin));
      atomicAdd(&g_conv_timing[4], 1ull);
#endif
    }
  } else {
    // ----------------------------------------------------------------- epilogue
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t stg_base = smem_base + L::kOutOffset + (warp - 2) * L::kOutBytesPerWarp;
    const uint32_t tab_base = smem_base + L::kTabOffset + (warp - 2) * L::kTabBytesPerWarp;
    const bool film_tab = p.film != nullptr && p.film_ld == 0;
    const bool film_row = p.film != nullptr && p.film_ld != 0;
    constexpr int kHalfN = BLOCK_N / 2;  // columns handled by this warp

    // One output tile. pass 0: out = acc * a + b (bias / FiLM folded per column) + residual,
    // partial GroupNorm sums, raw output. pass 1 (fused GroupNorm only): the same accumulator
    // again, now with the per-(clip-row, column) normalisation folded into the table, SiLU, and
    // the normalised tile goes out through tmOut2.
    uint32_t box_ord = 0;  // pass 1: running index of this warp's output boxes
    auto run_tile = [&](int tile, uint32_t acc, uint32_t wait_parity, int pass) {
      const int m_tile0 = (tile / p.n_tiles) * (kBlockM * CG) + cta_rank * kBlockM;
      const long long m = (long long)m_tile0 + row;
      const int n0 = (tile % p.n_tiles) * BLOCK_N;
      const bool in_range = m < p.m;
      const int r = in_range ? (int)((int)m / p.tp) : 0;
      const int t = in_range ? (int)m - r * p.tp : 0;
      const bool valid = in_range && t < p.t_valid;
      // clip-rows touched by this warp's 32 slots (GroupNorm partial sums are per clip-row)
      const int m_first = m_tile0 + quad * 32;
      const int m_last = m_first + 31 < (int)p.m - 1 ? m_first + 31 : (int)p.m - 1;
      const int r_lo = m_first / p.tp;
      const int r_hi = m_first < (int)p.m ? m_last / p.tp : r_lo - 1;

      // per-column (a, b) with out = acc * a + b: bias and the FiLM modulation
      // (acc + bias) * (1 + scale) + shift folded; built while the MMAs of this tile still run.
      // (per-row FiLM tables, film_ld != 0, are applied per lane further down)
      {
        __syncwarp();  // previous tile's table reads are done
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        const int cl = lane * 4;                        // column inside this warp's half
        const int ncol = n0 + half * kHalfN + cl;
        if (cl < kHalfN) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol));
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), h4 = s4;
          if (film_tab) {
            s4 = __ldg(reinterpret_cast<const float4*>(p.film + ncol));
            h4 = __ldg(reinterpret_cast<const float4*>(p.film + p.film_shift_off + ncol));
          }
          lo = make_float4(1.f + s4.x, fmaf(b4.x, 1.f + s4.x, h4.x), 1.f + s4.y,
                           fmaf(b4.y, 1.f + s4.y, h4.y));
          hi = make_float4(1.f + s4.z, fmaf(b4.z, 1.f + s4.z, h4.z), 1.f + s4.w,
                           fmaf(b4.w, 1.f + s4.w, h4.w));
        }
        if (pass == 0) {
          if (cl < kHalfN) {
            const uint32_t ta = tab_base + (uint32_t)cl * 8u;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "f"(lo.x),
                         "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta + 16u), "f"(hi.x),
                         "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
          }
        } else {
          // GroupNorm folded in: y = (v - mean) * rstd * gamma + beta with v = acc * a + b
          //   => A = a * rstd * gamma, B = (b - mean) * rstd * gamma + beta, one table per clip-row
          const int cg = p.n_valid / p.gn_groups;          // channels per group (multiple of 32)
          const int spg = cg >> 5;                         // 32-channel sub-blocks per group
          const int cnt = spg * p.stats_ns;                // partial sums per (clip-row, group)
          const int g_first = (n0 + half * kHalfN) / cg;
          const int g_last = (n0 + half * kHalfN + kHalfN - 1) / cg;
          const double inv_n = 1.0 / ((double)cg * (double)p.t_valid);
          float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), be4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cl < kHalfN && ncol < p.n_valid) {
            g4 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + ncol));
            be4 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + ncol));
          }
          const int my_g = ncol / cg;                      // 4 consecutive columns: one group
          // all partial-sum loads of the (<= 2 clip-rows) x (<= 4 groups) of this warp are issued
          // before any reduction, so their L2 latency is paid once
          constexpr int kMaxG = kHalfN / 32;
          float2 part[2][kMaxG];
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            const int rr = ci == 0 ? r_lo : r_hi;
            const bool live = r_hi >= r_lo && (ci == 0 || r_hi != r_lo);
#pragma unroll
            for (int gi = 0; gi < kMaxG; ++gi) {
              part[ci][gi] = make_float2(0.f, 0.f);
              const int g = g_first + gi;
              if (live && g <= g_last && g < p.gn_groups) {
                const float2* sp =
                    p.stats + ((size_t)rr * p.stats_sub + (size_t)g * spg) * p.stats_ns;
                for (int i = lane; i < cnt; i += 32) {
                  const float2 v = __ldcg(sp + i);
                  part[ci][gi].x += v.x;
                  part[ci][gi].y += v.y;
                }
              }
            }
          }
          for (int ci = 0; ci < 2; ++ci) {
            float mean_c = 0.f, rstd_c = 0.f;
#pragma unroll
            for (int gi = 0; gi < kMaxG; ++gi) {
              const int g = g_first + gi;
              if (g > g_last) break;  // warp-uniform
              double sa = (double)part[ci][gi].x, sb = (double)part[ci][gi].y;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                sa += __shfl_xor_sync(0xffffffffu, sa, o);
                sb += __shfl_xor_sync(0xffffffffu, sb, o);
              }
              const double mean = sa * inv_n;
              double var = sb * inv_n - mean * mean;
              var = var > 0.0 ? var : 0.0;
              if (g == my_g) {
                mean_c = (float)mean;
                rstd_c = (float)(1.0 / sqrt(var + (double)p.gn_eps));
              }
            }
            if (cl < kHalfN) {
              const float ga[4] = {g4.x * rstd_c, g4.y * rstd_c, g4.z * rstd_c, g4.w * rstd_c};
              const float4 tlo = make_float4(lo.x * ga[0], fmaf(lo.y - mean_c, ga[0], be4.x),
                                             lo.z * ga[1], fmaf(lo.w - mean_c, ga[1], be4.y));
              const float4 thi = make_float4(hi.x * ga[2], fmaf(hi.y - mean_c, ga[2], be4.z),
                                             hi.z * ga[3], fmaf(hi.w - mean_c, ga[3], be4.w));
              const uint32_t ta = tab_base + (uint32_t)(ci * kHalfN + cl) * 8u;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta), "f"(tlo.x),
                           "f"(tlo.y), "f"(tlo.z), "f"(tlo.w) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ta + 16u), "f"(thi.x),
                           "f"(thi.y), "f"(thi.z), "f"(thi.w) : "memory");
            }
          }
        }
        __syncwarp();
      }
      // pass 1: lanes of the second clip-row of this warp read the second table
      const uint32_t tab_lane =
          tab_base + ((pass == 1 && in_range && r != r_lo) ? (uint32_t)kHalfN * 8u : 0u);

      if (pass == 0) {
        mbar_wait(tfull_bar(acc), wait_parity);
        tc_fence_after_sync();
      }
      const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(quad * 32) << 16);
      const CUtensorMap* tm_out = pass == 0 ? &tmOut : &tmOut2;

#pragma unroll 1
      for (int c0 = half * kHalfN; c0 < (half + 1) * kHalfN; c0 += 32) {
        const int n = n0 + c0;
        if (n >= p.n_valid) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(taddr + c0, v);
        float f[32];
        const uint32_t tcol = tab_lane + (uint32_t)(c0 - half * kHalfN) * 8u;
        float4 ab[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(ab[j].x), "=f"(ab[j].y), "=f"(ab[j].z), "=f"(ab[j].w)
                       : "r"(tcol + 16u * j));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          f[2 * j] = fmaf(__uint_as_float(v[2 * j]), ab[j].x, ab[j].y);
          f[2 * j + 1] = fmaf(__uint_as_float(v[2 * j + 1]), ab[j].z, ab[j].w);
        }
        if (pass == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = silu_tanh(f[j]);
        }
        if (film_row && pass == 0) {
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
        if (pass == 0 && p.residual != nullptr && valid) {
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
        if (pass == 0 && p.stats != nullptr) {
          // Partial sum / sum of squares of this warp's 32 slots x 32 channels, per clip-row
          // and per `gran`-channel sub-block, written (not accumulated) to a slot that only
          // this warp owns: the consumer (gn_apply) adds the slices in a fixed order, so the
          // statistics are deterministic and need no zeroing between steps.
          float q1[4], q2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = valid ? f[q * 8 + j] : 0.f;
              a += v;
              b = fmaf(v, v, b);
            }
            q1[q] = a;
            q2[q] = b;
          }
          const int nsub = 32 / p.stats_gran;  // sub-blocks inside this 32-channel chunk
          if (nsub == 1) {
            q1[0] = (q1[0] + q1[1]) + (q1[2] + q1[3]);
            q2[0] = (q2[0] + q2[1]) + (q2[2] + q2[3]);
          } else if (nsub == 2) {
            q1[0] += q1[1];
            q2[0] += q2[1];
            q1[1] = q1[2] + q1[3];
            q2[1] = q2[2] + q2[3];
          }
          const int sub0 = (n >> 3) / (p.stats_gran >> 3);
          for (int rr = r_lo; rr <= r_hi; ++rr) {
            const bool mine = in_range && r == rr;
            const int t_first = m_first - rr * p.tp;
            const int slice = t_first > 0 ? (t_first + 31) >> 5 : 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < nsub) {
                const float a = warp_sum(mine ? q1[q] : 0.f);
                const float b = warp_sum(mine ? q2[q] : 0.f);
                if (lane == 0)
                  p.stats[((size_t)rr * p.stats_sub + sub0 + q) * p.stats_ns + slice] =
                      make_float2(a, b);
              }
            }
          }
        }
        if (p.out_mode == LM2A_OUT_BF16_SLAB) {
          // bf16 -> this warp's [32 slots][32 channels] staging tile (64B swizzle) -> one TMA
          // store; rows past the slab and channels past n_valid are clipped by the tensor map,
          // pad slots (t >= t_valid) are written as zeros to keep the conv padding intact.
          // (fused GroupNorm with no raw output requested: pass 0 stores nothing)
          if (pass == 1 || p.out != nullptr) {
            // pass 1 runs after every MMA of the CTA (pair): the operand pipeline's shared
            // memory is idle, so each box gets a staging tile of its own there and no store
            // waits for the previous one; pass 0 reuses this warp's single staging tile
            uint32_t stg = stg_base;
            if (pass == 1) {
              stg = smem_base + ((uint32_t)((warp - 2) * kMaxAcc * (kHalfN / 32)) + box_ord) *
                                    (uint32_t)L::kOutBytesPerWarp;
              ++box_ord;
            } else {
              if (lane == 0) tma_store_wait_read<0>();  // the previous box has left the tile
              __syncwarp();
            }
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
              tma_store_2d(tm_out, stg, n, m_first);
              tma_store_commit();
            }
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
    };

    if (!fuse_gn) {
      uint32_t acc = 0, acc_phase = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        run_tile(tile, acc, acc_phase, 0);
        tc_fence_before_sync();
        if (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
        acc ^= 1u;
        if (acc == 0) acc_phase ^= 1u;
      }
    } else {
      // pass 0 over this CTA's tiles (their accumulators stay in TMEM), grid barrier once the
      // partial GroupNorm sums of every CTA are in global memory, then the normalising pass
      uint32_t ord = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) run_tile(tile, ord++, 0, 0);
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      if (warp == 2 && lane == 0) {
        volatile unsigned int* cnt = p.gn_barrier;
        volatile unsigned int* gen = p.gn_barrier + 1;
        const unsigned int g0 = *gen;
        __threadfence();
        if (atomicAdd(p.gn_barrier, 1u) == gridDim.x - 1) {
          *cnt = 0u;
          __threadfence();
          atomicAdd(p.gn_barrier + 1, 1u);
        } else {
          const uint64_t t0 = global_timer_ns();
          unsigned int spins = 0;
          while (*gen == g0) {
            if ((++spins & 0xff) == 0 && global_timer_ns() - t0 > 4000000000ull) {
              printf("lm2a: conv grid barrier timeout (block %d)\n", (int)blockIdx.x);
              __trap();
            }
          }
        }
        __threadfence();
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      tc_fence_after_sync();
      ord = 0;
      for (int tile = unit; tile < total_tiles; tile += num_units) run_tile(tile, ord++, 0, 1);
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
// 2-D bf16 map: inner = channels (box 64 -> 128 B, SWIZZLE_128B), outer = slots / rows.
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

template <int BLOCK_N, int STAGES, int CG>
int launch(cudaStream_t stream, const CUtensorMap& a0, const CUtensorMap& a1,
           const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& o2,
           const ConvArgs& args) {
  using L = SmemLayout<BLOCK_N, STAGES, CG>;
  auto kern = conv_gemm_kernel<BLOCK_N, STAGES, CG>;
  static bool configured = false;
  if (!configured) {
    LM2A_CUDA_OK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
    configured = true;
  }
  const int tiles = args.m_tiles * args.n_tiles;
  const int units = num_sms() / CG;  // CTAs (CG = 1) or CTA pairs (CG = 2) that fit the chip
  const int grid = (tiles < units ? tiles : units) * CG;
  LM2A_CUDA_OK(launch_kernel_cluster(kern, dim3(grid), dim3(kThreads), L::kBytes, stream,
                                     (unsigned)CG, a0, a1, b, o, o2, args));
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
// max_tiles_per_unit > 0 restricts the choice to shapes where no CTA (pair) gets more tiles than
// it has TMEM accumulators (512 / block_n): the fused GroupNorm keeps every tile resident.
// Returns block_n = 0 when nothing fits.
TileChoice choose_tile(long long m, int n_pad, bool keep_tiles_in_tmem = false) {
  const long long sms = num_sms();
  TileChoice best{keep_tiles_in_tmem ? 0 : 128, 1};
  long long best_cost = -1;
  const int bns[2] = {256, 128};
  for (int cg = 2; cg >= 1; --cg) {
    for (int bi = 0; bi < 2; ++bi) {
      const int bn = bns[bi];
      if (n_pad % bn != 0) continue;
      const long long tiles = ((m + 128 * cg - 1) / (128 * cg)) * (n_pad / bn);
      const long long units = sms / cg;
      const long long waves = (tiles + units - 1) / units;
      if (keep_tiles_in_tmem && waves > 512 / bn) continue;
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
                   (d->out != nullptr || d->gn_gamma != nullptr),
               "conv1d: null x / w / bias / out pointer");
  LM2A_REQUIRE(d->m > 0 && d->tp > 0 && d->t_valid > 0 && d->t_valid <= d->tp &&
                   d->m % d->tp == 0,
               "conv1d: bad slab geometry m=%lld tp=%d t_valid=%d", (long long)d->m, d->tp,
               d->t_valid);
  LM2A_REQUIRE(d->n_pad > 0 && d->n_pad % 128 == 0 && d->n_valid > 0 && d->n_valid <= d->n_pad,
               "conv1d: n_pad=%d must be a positive multiple of 128 (n_valid=%d)", d->n_pad,
               d->n_valid);
  LM2A_REQUIRE(d->m < (1ll << 31) - 256, "conv1d: too many slots");

  ConvArgs a{};
  CUtensorMap tmA[2];
  int k_total = 0;
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
    if (g.taps == LM2A_TAPS_K4S2) {
      LM2A_REQUIRE(g.rows == 2 * d->m,
                   "conv1d: k4s2 needs input slots (%lld) == 2 * output slots (%lld)",
                   (long long)g.rows, (long long)d->m);
      a.seg_half[s] = g.ld;
      ntaps = 4;
      if (encode_2d(&tmA[s], g.x, (uint64_t)g.ld + g.cin, (uint64_t)g.rows / 2,
                    (uint64_t)g.ld * 2, kBlockM))
        return 1;
    } else {
      LM2A_REQUIRE(g.rows == d->m, "conv1d: seg %d slots (%lld) != output slots (%lld)", s,
                   (long long)g.rows, (long long)d->m);
      a.seg_half[s] = 0;
      ntaps = g.taps == LM2A_TAPS_K3 ? 3 : 1;
      if (encode_2d(&tmA[s], g.x, (uint64_t)g.cin, (uint64_t)g.rows, (uint64_t)g.ld, kBlockM))
        return 1;
    }
    k_total += ntaps * g.cin;
  }
  const bool fuse_gn = d->gn_gamma != nullptr;
  if (fuse_gn) {
    LM2A_REQUIRE(d->gn_beta != nullptr && d->gn_out != nullptr && d->gn_barrier != nullptr &&
                     d->stats != nullptr && d->stats_gran == 32,
                 "conv1d: fused GroupNorm needs gn_beta, gn_out, gn_barrier and 32-channel stats");
    LM2A_REQUIRE(d->out_mode == LM2A_OUT_BF16_SLAB && d->gn_groups > 0 &&
                     d->n_valid % d->gn_groups == 0 && (d->n_valid / d->gn_groups) % 32 == 0 &&
                     d->tp >= 32 && (d->film == nullptr || d->film_ld == 0),
                 "conv1d: fused GroupNorm needs >= 32-channel groups, tp >= 32 and a uniform "
                 "FiLM table (n=%d groups=%d tp=%d film_ld=%d)", d->n_valid, d->gn_groups,
                 d->tp, d->film_ld);
    LM2A_REQUIRE(d->gn_out_ld % 8 == 0 && d->gn_out_ld >= d->n_valid &&
                     ((reinterpret_cast<uintptr_t>(d->gn_out) |
                       reinterpret_cast<uintptr_t>(d->gn_gamma) |
                       reinterpret_cast<uintptr_t>(d->gn_beta)) & 15) == 0,
                 "conv1d: gn_out / gn_gamma / gn_beta alignment");
  }
  int block_n = d->block_n, cg = d->cta_group;
  if (fuse_gn && (block_n == 0 || cg == 0)) {
    const TileChoice c = choose_tile(d->m, d->n_pad, true);
    LM2A_REQUIRE(c.block_n != 0, "conv1d: fused GroupNorm: %lld x %d output does not fit the "
                 "chip's TMEM accumulators (see lm2a_conv_gn_fusable)", (long long)d->m, d->n_pad);
    block_n = c.block_n;
    cg = c.cg;
  }
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

  a.num_kb = k_total / 64;
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
  a.stats = reinterpret_cast<float2*>(d->stats);
  a.stats_sub = d->stats_sub;
  a.stats_ns = d->stats_ns;
  a.stats_gran = d->stats_gran;
  if (d->stats != nullptr) {
    LM2A_REQUIRE(d->out_mode == LM2A_OUT_BF16_SLAB, "conv1d: stats need a bf16 slab output");
    LM2A_REQUIRE((d->stats_gran == 8 || d->stats_gran == 16 || d->stats_gran == 32) &&
                     d->stats_sub > 0 && d->stats_ns >= d->tp / 32 + 2 &&
                     (reinterpret_cast<uintptr_t>(d->stats) & 7) == 0,
                 "conv1d: bad stats layout (gran=%d sub=%d ns=%d, need ns >= tp/32+2 = %d)",
                 d->stats_gran, d->stats_sub, d->stats_ns, d->tp / 32 + 2);
  }
  if (d->out_mode == LM2A_OUT_BF16_SLAB) {
    LM2A_REQUIRE((d->out == nullptr || (d->out_ld % 8 == 0 && d->out_ld >= d->n_valid)) &&
                     d->n_valid % 32 == 0,
                 "conv1d: bf16 slab output needs ld %% 8 == 0 and n_valid %% 32 == 0 (ld=%d n=%d)",
                 d->out_ld, d->n_valid);
  } else {
    LM2A_REQUIRE(d->out_mode == LM2A_OUT_F32_NCT, "conv1d: bad out_mode %d", d->out_mode);
  }

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUtensorMap tmOut = tmB, tmOut2 = tmB;  // unused in the fp32 output mode / without fusion
  if (d->out_mode == LM2A_OUT_BF16_SLAB && d->out != nullptr &&
      encode_2d_out(&tmOut, d->out, (uint64_t)d->n_valid, (uint64_t)d->m, (uint64_t)d->out_ld))
    return 1;
  a.gn_gamma = d->gn_gamma;
  a.gn_beta = d->gn_beta;
  a.gn_groups = d->gn_groups;
  a.gn_eps = d->gn_eps;
  a.gn_barrier = reinterpret_cast<unsigned int*>(d->gn_barrier);
  if (fuse_gn) {
    const long long tiles = (long long)a.m_tiles * a.n_tiles;
    const long long units = num_sms() / cg;
    LM2A_REQUIRE((tiles + units - 1) / units <= 512 / block_n,
                 "conv1d: fused GroupNorm: %lld tiles over %lld CTA%s exceed the %d TMEM "
                 "accumulators", tiles, units, cg == 2 ? " pairs" : "s", 512 / block_n);
    if (encode_2d_out(&tmOut2, d->gn_out, (uint64_t)d->n_valid, (uint64_t)d->m,
                      (uint64_t)d->gn_out_ld))
      return 1;
  }
  if (cg == 2) {
    if (block_n == 256) return launch<256, 6, 2>(st, tmA[0], tmA[1], tmB, tmOut, tmOut2, a);
    return launch<128, 8, 2>(st, tmA[0], tmA[1], tmB, tmOut, tmOut2, a);
  }
  if (block_n == 256) return launch<256, 4, 1>(st, tmA[0], tmA[1], tmB, tmOut, tmOut2, a);
  return launch<128, 6, 1>(st, tmA[0], tmA[1], tmB, tmOut, tmOut2, a);
}

extern "C" int lm2a_conv_gn_fusable(int64_t m, int32_t n_pad) {
  if (m <= 0 || n_pad <= 0 || n_pad % 128 != 0) return 0;
  return lm2a::choose_tile(m, n_pad, true).block_n != 0 ? 1 : 0;
}

#ifdef LM2A_CONV_TIMING
// instrumented build only: read (and clear) the counters of g_conv_timing into out[8]
extern "C" int lm2a_conv_timing_read(unsigned long long* out) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, lm2a::g_conv_timing, 8 * sizeof(unsigned long long)) != cudaSuccess)
    return 1;
  unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  return cudaMemcpyToSymbol(lm2a::g_conv_timing, zero, sizeof(zero)) == cudaSuccess ? 0 : 1;
}
#endif
