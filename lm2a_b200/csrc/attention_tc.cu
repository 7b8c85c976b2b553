// Cross-attention core softmax(q k^T) v on the 5th-gen tensor cores (tcgen05 + TMEM),
// both condition streams (motion, lyrics) of CrossAttentionFusion in one launch
// (reference models/cross_attention.py:50-61, i.e. the bmm / softmax / bmm inside
// nn.MultiheadAttention). K and V^T are per-clip caches built once per clip; q arrives
// pre-scaled by log2(e)/sqrt(d_h) so the softmax is a bare exp2. The [Tq, Lk] probability
// matrix never leaves the SM (the reference materialises it, head-averages it and throws
// it away).
//
// One CTA = 128 queries of one (clip-row, stream, head); keys are walked in tiles of 64.
//   warp 4     TMA producer: Q tile once + a ring of K tiles and a ring of V^T tiles (separate
//              rings: a K slot is free as soon as S_j is done, a V slot only after P_j V_j)
//   warp 5     tcgen05.mma issuer (one thread): S_j = Q K_j^T into a double-buffered TMEM
//              tile, O += P_j V_j with P_j read from TENSOR MEMORY (.ts operand form: the
//              softmax warps write the bf16 probabilities over the S tile they came from - no
//              shared-memory round trip, no proxy fence); S_{j+1} is issued before P_j V_j so
//              the tensor pipe works while the softmax warps run
//   warps 0-3  softmax: one query row per thread (TMEM lane), tcgen05.ld of S, running
//              max with lazy rescaling (O in TMEM is only corrected when the row max grows
//              by more than 2^8), exp2, row sum, P -> bf16 pairs -> tcgen05.st;
//              finally O / l -> bf16 slab
// Query tails (T = 516 / 258 / 129 leave 4 / 2 / 1 rows in a tile of their own): handing those
// rows to the producer warps of the full tiles (CUDA-core and mma.sync variants, commit
// "attention: tail query rows on the producer warp") measured neutral to slower on B200
// (92 vs 94 / 67 vs 67 / 64 vs 60 us per launch at levels 0 / 1 / 2) and was removed. Three CTAs
// per SM at dh = 32 by CAPPING the registers (112-register cap, S walked in two 32-column halves,
// no spills) measured 112 vs 91 us and was removed as well; three CTAs per SM by RE-BALANCING
// the registers between the roles (setmaxnreg, below) is what dh = 32 / 64 run today.
// TMEM: S[0] cols 0-63, S[1] cols 64-127, O in dh further columns (dh = 64 three-CTA build: one
// S buffer, O behind it, P in a second 32-column allocation); two CTAs per SM for 64 < dh <= 128,
// three for dh = 32 / 64. Head dims: any multiple of 64 up to 384 (64-channel operand panels, 128B
// swizzle) plus 32 and 96 (32-channel panels, 64B swizzle) — the legacy UNet1D
// (reference models/unet1d.py:17-29: 4 heads over 256..1536 channels) needs 192, 256 and 384.
#include "../../include/lm2a_b200.h"
#include <type_traits>

#include "common.cuh"

namespace lm2a {
namespace {

constexpr int kBQ = 128;
constexpr int kBK = 64;
// d_h = 32: THREE CTAs per SM. The softmax warps need ~136 registers, so three CTAs of 192 threads
// do not fit the register file (two earlier attempts capped the registers instead: spills, or S
// re-read from TMEM in halves - both slower). Instead the CTA is launched as two warpgroups at
// 80 registers per thread and re-balanced with setmaxnreg: the softmax warpgroup grows to 128 - 136,
// the producer / issuer warpgroup (two of its warps only pad the warpgroup) shrinks to 32 - 24.
// 12 softmax warps per SM instead of 8: the kernel is latency-bound (ncu: XU pipe 42 % busy,
// 1.46 IPC per SM, a quarter of the softmax warps' samples waiting for the first S tile of
// their CTA), not MUFU-bound - moving a quarter of the exp2 to the FMA pipe changed nothing.
#ifndef LM2A_ATTN_TRI
#define LM2A_ATTN_TRI 1
#endif
constexpr uint32_t kTmemColsS = 128;  // two 64-column S tiles
constexpr float kRescaleThreshold = 8.0f;  // log2 units

template <int DH>
struct AttnSmem {
  static constexpr bool kTri = LM2A_ATTN_TRI && (DH == 32 || DH == 64);   // three CTAs per SM
  // d_h = 64: S (2 x 64 columns) + O (64) = 192 tensor-memory columns per CTA would only admit two
  // CTAs. Three fit with ONE S buffer, a separate 32-column P buffer and O behind S (64 + 64 in a
  // 128-column allocation, + 32): the softmax warps release S as soon as it is in registers
  // (s_free), so S_{j+1} is still computed under the exp2 work of chunk j; P_j waits for
  // P_{j-1} V_{j-1} (issued a whole chunk earlier).
  static constexpr bool kSingleS = kTri && DH == 64;
  static constexpr int kThreads = kTri ? 256 : 192;  // 4 softmax warps, TMA producer, MMA issuer
  static constexpr int kCtasPerSm = kTri ? 3 : 2;
  static constexpr int kPanelW = DH % 64 == 0 ? 64 : 32;  // channel panels of the Q / K tiles
  static constexpr int kPanels = DH / kPanelW;
  static constexpr bool kSw64 = kPanelW == 32;
  // V^T tile = DH rows x 64 keys; a TMA box / UMMA N is at most 256 rows
  static constexpr int kVBoxRows = DH > 256 ? DH / 2 : DH;
  static constexpr int kVBoxes = DH / kVBoxRows;
  static constexpr int kQPanelBytes = kBQ * kPanelW * 2;
  static constexpr int kKPanelBytes = kBK * kPanelW * 2;
  static constexpr int kQBytes = kBQ * DH * 2;
  static constexpr int kKBytes = kBK * DH * 2;
  static constexpr int kVBytes = DH * kBK * 2;
  static constexpr int kPBytes = 0;   // P_j lives in tensor memory (over its S tile)
  // dh = 128: two K + two V stages keep a CTA at 96 KB so that two fit an SM;
  // dh = 384: Q alone is 96 KB, one stage each (208 KB)
  static constexpr int kKStages = DH > 256 ? 1 : (DH > 64 ? 2 : 3);
  static constexpr int kVStages = DH > 256 ? 1 : (DH > 64 ? 2 : 3);
  static constexpr int kPBufs = kSingleS ? 1 : 2;   // p_full / pv_done pairs (one per P buffer)
  static constexpr int kPOff = kQBytes;
  static constexpr int kKOff = kPOff + kPBufs * kPBytes;
  static constexpr int kVOff = kKOff + kKStages * kKBytes;
  static constexpr int kBarOff = kVOff + kVStages * kVBytes;
  static constexpr int kNumBars = 1 + 2 * kKStages + 2 * kVStages + 2 + 2 * kPBufs + 2;
  static constexpr int kNeeded = kBarOff + 8 * kNumBars + 8;
  // two CTAs per SM (register budget of the softmax warps): ask for enough shared memory that a
  // third is never scheduled
  // (kTri: the register file admits exactly three 256-thread CTAs at 80 registers)
  static constexpr int kBytes = kTri ? kNeeded : (kNeeded < 80 * 1024 ? 80 * 1024 : kNeeded);
  static_assert(DH % 32 == 0 && DH >= 32 && DH <= 384, "head dim");
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
  // TMEM: S = 128 columns + O = DH columns rounded up to a power of two; dh > 256 takes the
  // whole 512 columns in one allocation (S first, O behind it)
  static constexpr bool kOneAlloc = DH > 256;
  static constexpr uint32_t kTmemColsO = DH <= 32 ? 32 : DH <= 64 ? 64 : DH <= 128 ? 128 : 256;
  static constexpr uint32_t kTmemColsP = 32;   // kSingleS: the second allocation holds P
};

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, bool sw64) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>((sw64 ? 512 : 1024) >> 4) << 32;  // 8 rows x swizzle span
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(sw64 ? 4 : 2) << 61;              // SWIZZLE_64B / SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: P_j is read from tensor memory, where the softmax warps wrote
// it over the S tile it was computed from (packed bf16 pairs: 8 columns per 16 keys)
__device__ __forceinline__ void umma_bf16_ts_tc(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
        "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7])
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Register split of a three-CTA build (256 threads launched at 80 registers: softmax warpgroup +
// other warpgroup = 160). d_h = 32: 128 / 32 - at 24 registers the producer / issuer loops spill
// and the operand pipeline slows (72.8 vs 68.3 us per launch); d_h = 64: 136 / 24 - its softmax
// (one S buffer, separate P buffer) spills at 128 (64.0 vs 58.1 us).
template <int DH>
__device__ __forceinline__ void tri_regs_softmax() {
  if constexpr (AttnSmem<DH>::kTri) {
    if constexpr (DH == 32) asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
  }
}
template <int DH>
__device__ __forceinline__ void tri_regs_other() {
  if constexpr (AttnSmem<DH>::kTri) {
    if constexpr (DH == 32) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
  }
}

template <int DH>
__global__ void __launch_bounds__(AttnSmem<DH>::kThreads, AttnSmem<DH>::kCtasPerSm)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                     const __grid_constant__ CUtensorMap tmKm,
                     const __grid_constant__ CUtensorMap tmKt,
                     const __grid_constant__ CUtensorMap tmVm,
                     const __grid_constant__ CUtensorMap tmVt, __nv_bfloat16* __restrict__ o,
                     int o_ld, const int* __restrict__ kv_slot, int tp, int t_valid, int lk,
                     int e, int heads) {
  using L = AttnSmem<DH>;
  constexpr bool kSw64 = L::kSw64;
  constexpr int KS = L::kKStages, VS = L::kVStages, PB = L::kPBufs;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t q_base = smem_base;
  auto k_tile = [&](int s) { return smem_base + L::kKOff + s * L::kKBytes; };
  auto v_tile = [&](int s) { return smem_base + L::kVOff + s * L::kVBytes; };
  const uint32_t bar_base = smem_base + L::kBarOff;
  // barrier slots (8 B each)
  const uint32_t q_full = bar_base;
  auto k_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (1 + KS + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (1 + 2 * KS + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (1 + 2 * KS + VS + s); };
  auto s_full = [&](int b) { return bar_base + 8u * (1 + 2 * KS + 2 * VS + b); };
  auto p_full = [&](int b) { return bar_base + 8u * (3 + 2 * KS + 2 * VS + b); };
  auto pv_done = [&](int b) { return bar_base + 8u * (3 + 2 * KS + 2 * VS + PB + b); };
  const uint32_t o_full = bar_base + 8u * (3 + 2 * KS + 2 * VS + 2 * PB);
  const uint32_t s_free = bar_base + 8u * (4 + 2 * KS + 2 * VS + 2 * PB);     // kSingleS only
  const uint32_t tmem_slot = bar_base + 8u * (5 + 2 * KS + 2 * VS + 2 * PB);  // S, then O (+4 B)
  constexpr bool kSingleS = L::kSingleS;
  constexpr uint32_t kTmemColsO = L::kTmemColsO;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.z;
  const int stream = blockIdx.y / heads;
  const int h = blockIdx.y % heads;
  const int q0 = blockIdx.x * kBQ;
  const int ntiles = (lk + kBK - 1) / kBK;
  // softmax warps whose 32 query rows are all past t_valid do nothing at all (their P rows
  // stay garbage: MMA rows are independent and those O rows are never stored)
  const int valid_warps = min(4, (t_valid - q0 + 31) >> 5);
  // a last key tile of <= 16 keys (Lk = 516 = 8 * 64 + 4) is computed 16 keys wide: S with
  // N = 16, softmax over 16 columns, one K = 16 step of P.V
  const bool short_last = lk - (ntiles - 1) * kBK <= 16;

  if (warp == 4 && lane == 0) {
    if ((smem_base & 1023u) != 0) {
      printf("lm2a: attention shared memory base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(stream ? &tmKt : &tmKm);
    tma_prefetch_desc(stream ? &tmVt : &tmVm);
    mbar_init(q_full, 1);
    for (int s = 0; s < KS; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
    }
    for (int s = 0; s < VS; ++s) {
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) mbar_init(s_full(b), 1);
    for (int b = 0; b < PB; ++b) {
      mbar_init(p_full(b), 32 * valid_warps);
      mbar_init(pv_done(b), 1);
    }
    mbar_init(o_full, 1);
    mbar_init(s_free, 32 * valid_warps);
    mbar_fence_init();
  }
  if (warp == 5) {
    if (L::kOneAlloc) {
      tmem_alloc(tmem_slot, 512);
    } else {
      tmem_alloc(tmem_slot, kTmemColsS);
      tmem_alloc(tmem_slot + 4, kSingleS ? L::kTmemColsP : kTmemColsO);
    }
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base, tmem_o, tmem_p = 0;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (L::kOneAlloc) tmem_o = tmem_base + kTmemColsS;
  else asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_o) : "r"(tmem_slot + 4));
  if (kSingleS) {   // S in columns [0, 64) and O in [64, 128) of the first allocation, P apart
    tmem_p = tmem_o;
    tmem_o = tmem_base + kBK;
  }
  // S buffer of chunk j and the parity of its s_full phase
  auto s_buf = [&](int j) { return kSingleS ? 0 : (j & 1); };
  auto s_par = [&](int j) { return (uint32_t)(kSingleS ? j : (j >> 1)) & 1u; };
  // everything above overlapped the previous kernel's tail; from here on we touch its output
  pdl_wait();
  pdl_launch_dependents();

  // kTri: each role's code is dominated by its own setmaxnreg (ptxas budgets the registers of a
  // region from the setmaxnreg that dominates it); warps 6, 7 only complete the warpgroup
  // producer / issuer: the lean wait in the three-CTA builds (24 - 32 registers per thread)
  auto role_wait = [&](uint32_t bar, uint32_t parity) {
    if constexpr (L::kTri) mbar_wait_lean(bar, parity);
    else mbar_wait(bar, parity);
  };
  if (warp == 4) {
    tri_regs_other<DH>();
    // ------------------------------------------------------------- TMA producer
    // Q, then K_{j+1} before V_j: a K slot frees when S_{j+1-KS} is done, a V slot when
    // P_{j-VS} V_{j-VS} is done, which happen in this order
    const int slot = kv_slot[r];
    const CUtensorMap* km = stream ? &tmKt : &tmKm;
    const CUtensorMap* vm = stream ? &tmVt : &tmVm;
    auto load_k = [&](int j) {
      const int st = j % KS;
      role_wait(k_empty(st), ((uint32_t)(j / KS) & 1u) ^ 1u);
      mbar_expect_tx(k_full(st), L::kKBytes);
#pragma unroll
      for (int p = 0; p < L::kPanels; ++p)
        tma_load_2d(k_tile(st) + p * L::kKPanelBytes, km, h * DH + p * L::kPanelW,
                    slot * lk + j * kBK, k_full(st));
    };
    auto load_v = [&](int j) {
      const int st = j % VS;
      role_wait(v_empty(st), ((uint32_t)(j / VS) & 1u) ^ 1u);
      mbar_expect_tx(v_full(st), L::kVBytes);
#pragma unroll
      for (int c = 0; c < L::kVBoxes; ++c)
        tma_load_2d(v_tile(st) + c * L::kVBoxRows * kBK * 2, vm, j * kBK,
                    slot * e + h * DH + c * L::kVBoxRows, v_full(st));
    };
    if (lane == 0) {
      mbar_expect_tx(q_full, L::kQBytes);
#pragma unroll
      for (int p = 0; p < L::kPanels; ++p)
        tma_load_2d(q_base + p * L::kQPanelBytes, &tmQ, stream * e + h * DH + p * L::kPanelW,
                    r * tp + q0, q_full);
      load_k(0);
    }
    if (lane == 0) {
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 5) {
    tri_regs_other<DH>();
    // --------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kBQ, kBK);
      constexpr uint32_t idesc_s16 = umma_idesc_bf16(kBQ, 16);   // short last key tile
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kBQ, L::kVBoxRows);
      auto issue_pv = [&](int jj) {
        const int pb = jj % PB, st = jj % VS;
        role_wait(v_full(st), (uint32_t)(jj / VS) & 1u);
        role_wait(p_full(pb), (uint32_t)(jj / PB) & 1u);
        tc_fence_after_sync();
        const int keys = min(kBK, lk - jj * kBK);
        const int ksteps = (keys + 15) >> 4;
        const uint32_t p_tmem =
            kSingleS ? tmem_p : tmem_base + (uint32_t)(jj & 1) * kBK;   // P_jj over S_jj
        const uint64_t bdesc = umma_desc_kmajor(v_tile(st), false);
#pragma unroll
        for (int c = 0; c < L::kVBoxes; ++c) {
          // rows [c * kVBoxRows, ...) of the V^T tile -> O columns of the same range
          const uint64_t bd = bdesc + (uint64_t)((c * L::kVBoxRows * kBK * 2) >> 4);
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts_tc(tmem_o + c * L::kVBoxRows, p_tmem + 8u * k, bd + 2u * k, idesc_pv,
                            (jj | k) != 0 ? 1u : 0u);
        }
        umma_commit(v_empty(st));
        umma_commit(pv_done(pb));
      };
      role_wait(q_full, 0);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % KS, b = s_buf(j);
        role_wait(k_full(st), (uint32_t)(j / KS) & 1u);
        // one S buffer: S_{j-1} must be in the softmax warps' registers
        if (kSingleS && j >= 1) role_wait(s_free, (uint32_t)(j - 1) & 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const int panel = (k * 16) / L::kPanelW, kk = k % (L::kPanelW / 16);
          const uint64_t adesc =
              umma_desc_kmajor(q_base + panel * L::kQPanelBytes + kk * 32, kSw64);
          const uint64_t bdesc =
              umma_desc_kmajor(k_tile(st) + panel * L::kKPanelBytes + kk * 32, kSw64);
          umma_bf16_ss(tmem_base + b * kBK, adesc, bdesc,
                       (short_last && j == ntiles - 1) ? idesc_s16 : idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(k_empty(st));
        umma_commit(s_full(b));
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(ntiles - 1);
      umma_commit(o_full);
    }
  } else if (warp > 5) {
    tri_regs_other<DH>();
  } else {
    tri_regs_softmax<DH>();
    if (warp < valid_warps) {
    // ------------------------------------------------------------------ softmax
    const int row = warp * 32 + lane;  // TMEM lane == query row of the tile
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    const uint32_t p_row = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);

    // MASK: only the last key tile of a row can hold keys past Lk; the other tiles carry no
    // masking code at all (the compiler turns a guarded mask into 64 unconditional selects:
    // a quarter of the loop's instructions)
    auto softmax_tile = [&](auto nc_tag, auto mask_tag, int j) {
      constexpr int NC = decltype(nc_tag)::value;   // S columns (keys) of this tile: 64 or 16
      constexpr bool MASK = decltype(mask_tag)::value;
      const int b = s_buf(j), pb = j % PB;
      mbar_wait(s_full(b), s_par(j));
      tc_fence_after_sync();
      float s[NC];
      if constexpr (NC == 64) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(tmem_base + lane_off + b * kBK, v0);
        tmem_ld_32x32(tmem_base + lane_off + b * kBK + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          s[c] = __uint_as_float(v0[c]);
          s[32 + c] = __uint_as_float(v1[c]);
        }
      } else {
        uint32_t v0[16];
        tmem_ld_32x16(tmem_base + lane_off + b * kBK, v0);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) s[c] = __uint_as_float(v0[c]);
      }
      if constexpr (kSingleS) {   // S_j is in registers: S_{j+1} may overwrite the buffer
        tc_fence_before_sync();
        mbar_arrive(s_free);
      }
      if constexpr (MASK) {
        const int keys = lk - j * kBK;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c >= keys) s[c] = -INFINITY;
      }
      float mxa[4] = {s[0], s[1], s[2], s[3]};  // four independent chains
#pragma unroll
      for (int c = 4; c < NC; c += 4) {
        mxa[0] = fmaxf(mxa[0], s[c]);
        mxa[1] = fmaxf(mxa[1], s[c + 1]);
        mxa[2] = fmaxf(mxa[2], s[c + 2]);
        mxa[3] = fmaxf(mxa[3], s[c + 3]);
      }
      const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));

      // lazy rescale: keep the stale max while the new one is within 2^8 of it
      const bool grow = mx > m_used + kRescaleThreshold;
      float corr = 1.0f;
      if (grow) {
        corr = ex2_approx(m_used - mx);  // first tile: exp2(-inf) = 0
        m_used = mx;
        l_run *= corr;
      }
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        mbar_wait(pv_done((j - 1) % PB), (uint32_t)((j - 1) / PB) & 1u);
        tc_fence_after_sync();
        // (16 columns at a time: the 64 scores of the chunk are live across this rare path, and
        // 32 more registers would push them to local memory in the three-CTA builds)
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
          uint32_t ov[16];
          tmem_ld_32x16(tmem_o + lane_off + c0, ov);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * corr);
          tmem_st_32x16(tmem_o + lane_off + c0, ov);
        }
        tmem_st_wait();
      }

      // P_j overwrites the S tile it was computed from (this thread's own row; the S values are
      // in registers). S_{j+2}, the next product into this buffer, is issued after P_j V_j and
      // tcgen05.mma executes in issue order, so nothing else has to be waited for.
      uint32_t pk[NC / 2];
      l_run += softmax_probs<NC, MASK>(s, m_used, pk);   // exp2 split over the XU and FMA pipes
      uint32_t p_addr = tmem_base + lane_off + b * kBK;
      if constexpr (kSingleS) {
        // the one P buffer is free once P_{j-1} V_{j-1} has read it (issued a chunk ago)
        if (j >= 1) {
          mbar_wait(pv_done(0), (uint32_t)(j - 1) & 1u);
          tc_fence_after_sync();
        }
        p_addr = tmem_p + lane_off;
      }
      if constexpr (NC == 64) tmem_st_32x32(p_addr, pk);
      else tmem_st_32x8(p_addr, pk);
      tmem_st_wait();
      tc_fence_before_sync();
      mbar_arrive(p_full(pb));
    };
    using W64 = std::integral_constant<int, kBK>;
    using W16 = std::integral_constant<int, 16>;
#pragma unroll 1
    for (int j = 0; j < ntiles - 1; ++j) softmax_tile(W64{}, std::false_type{}, j);
    if (short_last) softmax_tile(W16{}, std::true_type{}, ntiles - 1);
    else softmax_tile(W64{}, std::true_type{}, ntiles - 1);

    // ---- finalise: O / l -> bf16 slab
    mbar_wait(o_full, 0);
    tc_fence_after_sync();
    const float inv = 1.0f / l_run;
    const int t = q0 + row;
    __nv_bfloat16* op = o + ((size_t)r * tp + t) * o_ld + stream * e + h * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t ov[32];
      tmem_ld_32x32(tmem_o + lane_off + c0, ov);
      tmem_ld_wait();
      if (t < t_valid) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(ov[c + 0]) * inv, __uint_as_float(ov[c + 1]) * inv);
          q.y = pack_bf16x2(__uint_as_float(ov[c + 2]) * inv, __uint_as_float(ov[c + 3]) * inv);
          q.z = pack_bf16x2(__uint_as_float(ov[c + 4]) * inv, __uint_as_float(ov[c + 5]) * inv);
          q.w = pack_bf16x2(__uint_as_float(ov[c + 6]) * inv, __uint_as_float(ov[c + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c0 + c) = q;
        }
      }
    }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after_sync();
    if (L::kOneAlloc) {
      tmem_dealloc(tmem_base, 512);
    } else {
      tmem_dealloc(tmem_base, kTmemColsS);
      if (kSingleS) tmem_dealloc(tmem_p, L::kTmemColsP);
      else tmem_dealloc(tmem_o, kTmemColsO);
    }
  }
}

// ------------------------------------------------------------------ host side
int encode_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
               uint64_t pitch_elems, uint32_t box_inner, uint32_t box_outer, bool sw64) {
  EncodeTiledFn fn = get_encode_fn();
  LM2A_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult res = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LM2A_REQUIRE(res == CUDA_SUCCESS,
               "cross_attn: cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu "
               "pitch=%llu box=%ux%u",
               (int)res, base, (unsigned long long)inner, (unsigned long long)outer,
               (unsigned long long)pitch_elems, box_inner, box_outer);
  return 0;
}

template <int DH>
int launch_attn(cudaStream_t st, const void* q, int q_ld, void* o, int o_ld, const void* k_m,
                const void* vt_m, const void* k_t, const void* vt_t, int k_ld, int vt_ld,
                const int32_t* kv_slot, int slots, int rows, int tp, int t_valid, int lk, int e,
                int heads, int n_streams) {
  using L = AttnSmem<DH>;
  constexpr bool sw64 = L::kSw64;
  auto kern = cross_attn_tc_kernel<DH>;
  constexpr int smem_bytes = L::kBytes;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    LM2A_CUDA_OK(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CUtensorMap tq, tkm, tkt, tvm, tvt;
  if (encode_map(&tq, q, (uint64_t)n_streams * e, (uint64_t)rows * tp, (uint64_t)q_ld,
                 L::kPanelW, kBQ, sw64))
    return 1;
  if (encode_map(&tkm, k_m, (uint64_t)e, (uint64_t)slots * lk, (uint64_t)k_ld, L::kPanelW, kBK,
                 sw64) ||
      encode_map(&tkt, k_t, (uint64_t)e, (uint64_t)slots * lk, (uint64_t)k_ld, L::kPanelW, kBK,
                 sw64))
    return 1;
  if (encode_map(&tvm, vt_m, (uint64_t)lk, (uint64_t)slots * e, (uint64_t)vt_ld, kBK,
                 L::kVBoxRows, false) ||
      encode_map(&tvt, vt_t, (uint64_t)lk, (uint64_t)slots * e, (uint64_t)vt_ld, kBK,
                 L::kVBoxRows, false))
    return 1;
  dim3 grid((t_valid + kBQ - 1) / kBQ, n_streams * heads, rows);
  LM2A_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(L::kThreads), smem_bytes, st, tq, tkm, tkt, tvm, tvt,
                                          reinterpret_cast<__nv_bfloat16*>(o), o_ld, kv_slot, tp,
                                          t_valid, lk, e, heads));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

// attention_res.cu: keys / values resident in shared memory (d_h = 32 / 64 when they fit)
int cross_attn_resident(cudaStream_t st, const void* q, int q_ld, void* o, int o_ld,
                        const void* k_m, const void* vt_m, const void* k_t, const void* vt_t,
                        int k_ld, int vt_ld, const int32_t* kv_slot, int slots, int rows, int tp,
                        int t_valid, int lk, int e, int heads, int n_streams);
}  // namespace lm2a

extern "C" int lm2a_cross_attn_streams_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                                            int32_t o_ld, const void* k_motion,
                                            const void* vt_motion, const void* k_text,
                                            const void* vt_text, int32_t k_ld, int32_t vt_ld,
                                            const int32_t* kv_slot, int32_t slots, int32_t rows,
                                            int32_t tp, int32_t t_valid, int32_t lk, int32_t e,
                                            int32_t heads, int32_t n_streams) {
  using namespace lm2a;
  LM2A_REQUIRE(n_streams == 1 || n_streams == 2, "cross_attn: n_streams=%d (1 or 2)", n_streams);
  LM2A_REQUIRE(q && o && k_motion && vt_motion && k_text && vt_text && kv_slot,
               "cross_attn: null pointer");
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && slots > 0 && tp > 0 && t_valid > 0 &&
                   t_valid <= tp && lk > 0,
               "cross_attn: bad geometry");
  LM2A_REQUIRE(heads > 0 && e % heads == 0, "cross_attn: e=%d not divisible by heads=%d", e,
               heads);
  LM2A_REQUIRE(q_ld % 8 == 0 && o_ld % 8 == 0 && k_ld % 8 == 0 && vt_ld % 8 == 0 &&
                   q_ld >= n_streams * e && o_ld >= n_streams * e && k_ld >= e && vt_ld >= lk,
               "cross_attn: bad pitches (q_ld=%d o_ld=%d k_ld=%d vt_ld=%d e=%d lk=%d)", q_ld,
               o_ld, k_ld, vt_ld, e, lk);
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(o) |
                 reinterpret_cast<uintptr_t>(k_motion) | reinterpret_cast<uintptr_t>(vt_motion) |
                 reinterpret_cast<uintptr_t>(k_text) | reinterpret_cast<uintptr_t>(vt_text)) &
                15) == 0,
               "cross_attn: tensors must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int dh = e / heads;
  {
    const int rc = cross_attn_resident(st, q, q_ld, o, o_ld, k_motion, vt_motion, k_text, vt_text,
                                       k_ld, vt_ld, kv_slot, slots, rows, tp, t_valid, lk, e,
                                       heads, n_streams);
    if (rc >= 0) return rc;   // -1: shape not covered by the resident kernel
  }
#define LM2A_ATTN_CASE(D)                                                                     \
  case D:                                                                                     \
    return launch_attn<D>(st, q, q_ld, o, o_ld, k_motion, vt_motion, k_text, vt_text, k_ld,   \
                          vt_ld, kv_slot, slots, rows, tp, t_valid, lk, e, heads, n_streams)
  switch (dh) {
    LM2A_ATTN_CASE(32);
    LM2A_ATTN_CASE(64);
    LM2A_ATTN_CASE(96);
    LM2A_ATTN_CASE(128);
    LM2A_ATTN_CASE(192);
    LM2A_ATTN_CASE(256);
    LM2A_ATTN_CASE(384);
    default:
      LM2A_REQUIRE(false, "cross_attn: head dim %d unsupported (32, 64, 96, 128, 192, 256, 384)",
                   dh);
  }
#undef LM2A_ATTN_CASE
  return 0;
}

extern "C" int lm2a_cross_attn_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                                    int32_t o_ld, const void* k_motion, const void* vt_motion,
                                    const void* k_text, const void* vt_text, int32_t k_ld,
                                    int32_t vt_ld, const int32_t* kv_slot, int32_t slots,
                                    int32_t rows, int32_t tp, int32_t t_valid, int32_t lk,
                                    int32_t e, int32_t heads) {
  return lm2a_cross_attn_streams_bf16(stream, q, q_ld, o, o_ld, k_motion, vt_motion, k_text,
                                      vt_text, k_ld, vt_ld, kv_slot, slots, rows, tp, t_valid, lk,
                                      e, heads, 2);
}
