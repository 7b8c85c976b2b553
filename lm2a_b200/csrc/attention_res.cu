// Cross-attention core with the keys / values RESIDENT in shared memory (tcgen05 + TMEM), the
// fast path of lm2a_cross_attn_streams_bf16 and the whole of lm2a_cross_attn_cond_bf16
// (reference models/cross_attention.py:50-61: the bmm / softmax / bmm inside
// nn.MultiheadAttention, for both condition streams of CrossAttentionFusion in one launch).
//
// Why resident: a clip has Lk = 516 condition frames, so the K / V^T panel of one (clip-row,
// stream, head) is 33 KB (d_h = 32) or 66 KB (d_h = 64) - it fits next to two query tiles. The
// streaming kernel (attention_tc.cu) re-reads it once per 128-query tile (4.3x crossbar traffic
// at level 0) and pays the prologue and the latency chain of a CTA per tile.
//
// One CTA = one K/V group (clip-row, stream, head) - or, in COND mode, one (clip-row, stream)
// whose 8 heads all attend to the same operand - times a 1/nz share of its query tiles:
//   producer    one warp, TMA: K chunks (64 keys each, one mbarrier per chunk so the first S
//               product starts when the first chunk lands), V^T chunks, then the Q tiles
//   issuers     one warp (one thread) per tile slot, tcgen05.mma: S_j = Q K_j^T into a
//               double-buffered TMEM tile per slot; O += P_j V_j with P_j read from TENSOR
//               MEMORY (the .ts operand form: the softmax warps write the bf16 probabilities
//               over the S tile they were computed from - no shared-memory round trip, no proxy
//               fence); S_{j+1} is issued before P_j V_j
//   softmax     four warps per tile slot: one query row per thread, running
//               max with lazy rescaling, exp2, row sum, P -> bf16 -> tcgen05.st; at the end of a
//               tile O / l -> bf16 slab. Two warps per scheduler keep the MUFU pipe fed.
// TMEM (512 columns): per slot S[0] 64 | S[1] 64 | O up to 128.
//
// COND mode (d_h = cond_dim = 128; UNet levels 2 and 3): K_h = C Wk_h^T + b and V_h = C Wv_h^T + b
// are rank-128 images of the SAME condition sequence C [Lk, 128]. The engine folds Wk_h into the
// query projection and Wv_h into the output projection (exact algebra on the host in fp64; the
// key bias shifts every score of a row equally and drops out of the softmax), so every head
// computes softmax(Q'_h C^T) C against the raw bf16 condition slab: one 132 KB operand per
// (clip-row, stream) instead of 2 MB of per-head K / V, shared by all heads and all levels
// (L2-resident), and used for BOTH products - K-major for Q' C^T, MN-major (the b_major bit of
// the instruction descriptor over the very same swizzled tile) for P C, so no transposed copy.
// Query tiles pack heads: T = 64 puts two heads in one 128-row tile, the leftover row of T = 129
// of all 8 heads shares one tile.
#include "../../include/lm2a_b200.h"
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"

namespace lm2a {

#ifdef LM2A_ATTN_TIMING
// Probe build only (tools/attn_probe.py): cycles summed over CTAs. [0] softmax warp 0 of slot 0
// waiting for S, [1] its whole chunk loop, [2] waiting for a free P buffer, [3] slot-0 issuer
// waiting for P, [4] waiting for a free S buffer, [5] issuer lifetime, [6] chunks, [7] CTAs,
// [8] softmax: S load (tcgen05.ld + wait), [9] P store + wait::st
__device__ unsigned long long g_attn_timing[16];
#define LM2A_T0(v) const long long v = clock64()
#define LM2A_TACC(acc, v) acc += clock64() - v
#else
#define LM2A_T0(v)
#define LM2A_TACC(acc, v)
#endif

namespace {

constexpr int kBQ = 128;
constexpr int kBK = 64;
// per tile slot: 4 softmax warps and one MMA-issuing warp; one TMA producer warp per CTA
// Two-slot CTAs carry a twelfth, idle warp: the 8 softmax warps are two warpgroups, and
// producer + two issuers + the idle warp a third, so the register file can be re-balanced with
// setmaxnreg (launched at 168 registers per thread - what an 11-warp CTA is granted anyway,
// warps being allocated in fours - the softmax warpgroups grow to 216, the third shrinks to 72):
// the softmax / epilogue code of the d_h = 128 build spilled at 168 (ncu: its per-tile epilogue
// stalled on local-memory reloads).
__host__ __device__ constexpr int res_threads(int slots) {
  return slots == 2 ? 32 * 12 : 32 * (5 * slots + 1);
}
#ifndef LM2A_RES_REBALANCE
#define LM2A_RES_REBALANCE 1
#endif
constexpr int kMaxChunks = 20;         // keys resident: up to 1280
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int kQBoxRows = 16;          // rows per Q TMA box (tail tiles pack heads in 16-row units)

struct ResArgs {
  __nv_bfloat16* o;
  int o_ld;
  const int* kv_slot;
  int tp, t_valid, lk, heads;
  int e;         // channels per stream of the q / o slabs (heads * DH)
  int ekv;       // rows per cache slot of the V^T operand (per-head mode: heads * DH)
  int nz;        // CTAs sharing a group's tiles
  int split_tail;  // 1: the group's tail tiles run in a CTA of their own (the last of the nz)
  int n_full;    // full 128-row tiles per head
  int rb;        // rows per head in a tail tile (16, 32, 64 or 128)
  int gpt;       // heads per tail tile
  int n_tail;    // tail tiles per group
  int k_rows;    // key rows allocated in shared memory (multiple of 16)
  int nchunks;   // ceil(lk / 64)
  int short_last;  // last chunk has <= 16 keys: computed 16 wide
};

template <int DH, bool COND>
struct ResCfg {
  static constexpr int kPW = DH % 64 == 0 ? 64 : 32;   // channel panel width of Q / K tiles
  static constexpr int kPanels = DH / kPW;
  static constexpr bool kSw64 = kPW == 32;
  static constexpr int kRowBytes = kPW * 2;
  static constexpr int kQPanelBytes = kBQ * kRowBytes;
  static constexpr int kQBytes = kBQ * DH * 2;
  static constexpr int kVChunkBytes = COND ? 0 : DH * kBK * 2;  // V^T: DH rows x 64 keys
  static constexpr int kHeadsPerGroupIsAll = COND ? 1 : 0;
  static_assert(DH == 32 || DH == 64 || DH == 128, "resident attention: head dim");
  static_assert(!COND || DH == 128, "cond mode: d_h = cond_dim = 128");
};

__device__ __forceinline__ uint64_t res_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                             uint32_t sbo_bytes, bool sw64) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(sw64 ? 4 : 2) << 61;   // SWIZZLE_64B / SWIZZLE_128B
  return d;
}
// K-major operand tile: rows of kPW bf16, 8-row swizzle atoms
__device__ __forceinline__ uint64_t res_desc_k(uint32_t smem_addr, bool sw64) {
  return res_desc(smem_addr, 16, sw64 ? 512 : 1024, sw64);
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
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
__device__ __forceinline__ void res_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void res_tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
        "r"(v[7])
      : "memory");
}
__device__ __forceinline__ void res_tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float res_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One query tile of a group: heads h0 .. h0 + g - 1, `rb` rows each starting at slot q0.
struct TileInfo {
  int h0, g, q0, rb;
};
__device__ __forceinline__ TileInfo decode_tile(const ResArgs& p, int hpg, int ti) {
  TileInfo t;
  const int nf = hpg * p.n_full;
  if (ti < nf) {
    t.h0 = ti / p.n_full;
    t.q0 = (ti - t.h0 * p.n_full) * kBQ;
    t.g = 1;
    t.rb = kBQ;
  } else {
    const int tt = ti - nf;
    t.h0 = tt * p.gpt;
    t.q0 = p.n_full * kBQ;
    t.g = min(p.gpt, hpg - t.h0);
    t.rb = p.rb;
  }
  return t;
}

// SLOTS: query tiles in flight per CTA. 2 (one CTA per SM, all 512 TMEM columns) when the
// resident operand fills the SM (COND, d_h = 64); 1 with two CTAs per SM where two operands fit
// (d_h = 32: 88 KB each), so that the CTAs hide each other's prologue, loads and tail tiles.
template <int DH, bool COND, int SLOTS>
__global__ void __launch_bounds__(res_threads(SLOTS), SLOTS == 1 ? 2 : 1)
cross_attn_res_kernel(const __grid_constant__ CUtensorMap tmQ,
                      const __grid_constant__ CUtensorMap tmKm,
                      const __grid_constant__ CUtensorMap tmKt,
                      const __grid_constant__ CUtensorMap tmVm,
                      const __grid_constant__ CUtensorMap tmVt, const ResArgs p) {
  using C = ResCfg<DH, COND>;
  constexpr bool kSw64 = C::kSw64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  // layout: Q[2] | K panels (k_rows rows each) | V^T chunks | barriers
  const uint32_t k_panel_bytes = (uint32_t)p.k_rows * C::kRowBytes;
  constexpr int kSlots = SLOTS;
  constexpr int kProdWarp = 4 * SLOTS;       // warps [0, 4 SLOTS): softmax; then producer; then issuers
  const uint32_t k_base = smem_base + kSlots * C::kQBytes;
  const uint32_t v_base = k_base + C::kPanels * k_panel_bytes;
  const uint32_t bar_base = v_base + (uint32_t)p.nchunks * C::kVChunkBytes;
  auto q_tile = [&](int sl) { return smem_base + sl * C::kQBytes; };
  // barrier slots (8 B each)
  auto k_full = [&](int j) { return bar_base + 8u * j; };
  auto v_full = [&](int j) { return bar_base + 8u * (kMaxChunks + j); };
  auto slot_bar = [&](int sl, int i) { return bar_base + 8u * (2 * kMaxChunks + sl * 12 + i); };
  auto q_full = [&](int sl) { return slot_bar(sl, 0); };
  auto q_free = [&](int sl) { return slot_bar(sl, 1); };
  auto s_full = [&](int sl, int b) { return slot_bar(sl, 2 + b); };
  auto p_full = [&](int sl, int b) { return slot_bar(sl, 4 + b); };
  auto pv_done = [&](int sl, int b) { return slot_bar(sl, 6 + b); };
  auto o_full = [&](int sl) { return slot_bar(sl, 8); };
  auto o_free = [&](int sl) { return slot_bar(sl, 9); };
  auto s_free = [&](int sl, int b) { return slot_bar(sl, 10 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxChunks + 2 * 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid (group, clip-row, z): the z index varies slowest, so the CTAs of the last z - with
  // split_tail the light ones - are dispatched after all others
  const int r = blockIdx.y;
  const int hpg = COND ? p.heads : 1;                 // heads per K/V group
  const int stream = COND ? (int)blockIdx.x : (int)blockIdx.x / p.heads;
  const int hg = COND ? 0 : (int)blockIdx.x % p.heads;  // the group's head (per-head mode)
  const int n_tiles = hpg * p.n_full + p.n_tail;
  // this CTA's tiles: t_start, t_start + t_stride, ... Normally the nz CTAs of a group take its
  // tiles round-robin. split_tail: nz - 1 CTAs share the full tiles, the last CTA takes the tail
  // tiles (T mod 128 rows of every head: a few rows per tile, one softmax warp active) - that
  // CTA is short and runs on an SM the single-wave grid leaves idle, instead of costing one of
  // the other CTAs a serial round of its own
  const int z = blockIdx.z;
  int t_start = z, t_stride = p.nz, my_tiles = z < n_tiles ? (n_tiles - z + p.nz - 1) / p.nz : 0;
  if (p.split_tail) {
    const int n_main = hpg * p.n_full, nzm = p.nz - 1;
    if (z == nzm) {
      t_start = n_main;
      t_stride = 1;
      my_tiles = p.n_tail;
    } else {
      t_stride = nzm;
      my_tiles = z < n_main ? (n_main - z + nzm - 1) / nzm : 0;
    }
  }
  const int nch = p.nchunks;

  if (warp == kProdWarp && lane == 0) {
    if ((smem_base & 1023u) != 0) {
      printf("lm2a: attention shared memory base not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(stream ? &tmKt : &tmKm);
    if (!COND) tma_prefetch_desc(stream ? &tmVt : &tmVm);
    for (int j = 0; j < nch; ++j) {
      mbar_init(k_full(j), 1);
      mbar_init(v_full(j), 1);
    }
    for (int sl = 0; sl < kSlots; ++sl) {
      mbar_init(q_full(sl), 1);
      mbar_init(q_free(sl), 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(s_full(sl, b), 1);
        mbar_init(p_full(sl, b), 128);
        mbar_init(s_free(sl, b), 128);
        mbar_init(pv_done(sl, b), 1);
      }
      mbar_init(o_full(sl), 1);
      mbar_init(o_free(sl), 128);
    }
    mbar_fence_init();
  }
  if (warp == kProdWarp + 1) {
    tmem_alloc(tmem_slot, 256u * SLOTS);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  auto tm_s = [&](int sl, int b) { return tmem_base + (uint32_t)(sl * 256 + b * kBK); };
  auto tm_o = [&](int sl) { return tmem_base + (uint32_t)(sl * 256 + 128); };
  // P_j: bf16 pairs, 32 columns. d_h <= 64 leaves room for two dedicated buffers behind O, so
  // S_{j+2} does not have to wait for P_j V_j; at d_h = 128 P_j overwrites the S tile it came from
  constexpr bool kSepP = DH <= 64;
  auto tm_p = [&](int sl, int b) {
    return kSepP ? tmem_base + (uint32_t)(sl * 256 + 192 + b * 32) : tm_s(sl, b);
  };
  // everything above overlapped the previous kernel's tail; from here on we touch its output
  // (the K / V caches are written once per batch, long before, but the kv_slot table and Q are
  // read after the wait)
  pdl_wait();
  pdl_launch_dependents();

  constexpr bool kRebalance = LM2A_RES_REBALANCE && SLOTS == 2;
  if (warp == kProdWarp) {
    if constexpr (kRebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    // ------------------------------------------------------------- TMA producer
    if (lane == 0 && my_tiles > 0) {
      const int slot = p.kv_slot[r];
      const CUtensorMap* km = stream ? &tmKt : &tmKm;
      const CUtensorMap* vm = stream ? &tmVt : &tmVm;
      auto load_q = [&](int sl, int ti) {
        const TileInfo t = decode_tile(p, hpg, ti);
        mbar_expect_tx(q_full(sl), C::kQBytes);
        const int per_head = t.rb / kQBoxRows;   // boxes per head and panel
#pragma unroll 1
        for (int bx = 0; bx < kBQ / kQBoxRows; ++bx) {
          const int g = min(bx / per_head, t.g - 1);   // unused sub-blocks re-read the last head
          const int row = r * p.tp + t.q0 + (bx % per_head) * kQBoxRows;
          const int col = stream * p.e + (hg + t.h0 + g) * DH;
#pragma unroll
          for (int pn = 0; pn < C::kPanels; ++pn)
            tma_load_2d(q_tile(sl) + pn * C::kQPanelBytes + bx * kQBoxRows * C::kRowBytes, &tmQ,
                        col + pn * C::kPW, row, q_full(sl));
        }
      };
      auto load_k = [&](int j) {
        // (the last chunk loads a full 64-row box as well: k_rows reserves the rows; keys past Lk
        // are another clip's rows or TMA zero fill, and are masked in the softmax)
        mbar_expect_tx(k_full(j), C::kPanels * kBK * C::kRowBytes);
#pragma unroll
        for (int pn = 0; pn < C::kPanels; ++pn)
          tma_load_2d(k_base + pn * k_panel_bytes + j * kBK * C::kRowBytes, km,
                      (COND ? 0 : hg * DH) + pn * C::kPW, slot * p.lk + j * kBK, k_full(j));
      };
      load_k(0);
      load_q(0, t_start);
      if (kSlots > 1 && my_tiles > 1) load_q(1, t_start + t_stride);
      for (int j = 1; j < nch; ++j) load_k(j);
      if (!COND) {
        for (int j = 0; j < nch; ++j) {
          mbar_expect_tx(v_full(j), C::kVChunkBytes);
          tma_load_2d(v_base + j * C::kVChunkBytes, vm, j * kBK, slot * p.ekv + hg * DH,
                      v_full(j));
        }
      }
      // remaining Q tiles: slot sl is reloaded once the last S product of its tile is done
      for (int i = kSlots; i < my_tiles; ++i) {
        const int sl = i % kSlots;
        mbar_wait(q_free(sl), (uint32_t)(i / kSlots - 1) & 1u);
        load_q(sl, t_start + i * t_stride);
      }
    }
  } else if (warp > kProdWarp) {
    // ------------------------------------------- MMA issuers: warp 9 -> slot 0, warp 10 -> slot 1
    // One thread per tile slot, so neither slot waits behind the other's barriers, and a lean
    // loop: every descriptor is precomputed and advanced by a constant per chunk (the issuing
    // thread's own instruction latency is on the critical path S_j -> softmax -> P_j V_j).
    if constexpr (kRebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int sl = warp - kProdWarp - 1;   // (== kSlots: the idle warp that completes the warpgroup)
    const int slot_tiles =
        (sl < kSlots && my_tiles > sl) ? (my_tiles - sl + kSlots - 1) / kSlots : 0;
    if (lane == 0 && slot_tiles > 0) {
      constexpr uint32_t kMajorB = COND ? (1u << 16) : 0u;   // P.V: B = C read MN-major
      constexpr uint32_t idesc_s = umma_idesc_bf16(kBQ, kBK);
      constexpr uint32_t idesc_s16 = umma_idesc_bf16(kBQ, 16);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kBQ, DH) | kMajorB;
      constexpr int KS = DH / 16;
      uint64_t qd[KS], kd[KS];
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        const int panel = (k * 16) / C::kPW, kk = k % (C::kPW / 16);
        qd[k] = res_desc_k(q_tile(sl) + panel * C::kQPanelBytes + kk * 32, kSw64);
        kd[k] = res_desc_k(k_base + panel * k_panel_bytes + kk * 32, kSw64);
      }
      constexpr uint32_t k_chunk_step = (kBK * C::kRowBytes) >> 4;
      // P.V operand: COND reads the C tile MN-major (K = keys advances by rows, N = channels
      // spans the two 64-channel panels: LBO = panel stride); else the V^T chunk K-major
      const uint64_t vd0 = COND ? res_desc(k_base, k_panel_bytes, 1024, false)
                                : res_desc_k(v_base, false);
      constexpr uint32_t v_chunk_step = COND ? (kBK * C::kRowBytes) >> 4 : C::kVChunkBytes >> 4;
      constexpr uint32_t v_k_step = COND ? (16 * C::kRowBytes) >> 4 : 2u;
      const uint32_t ts[2] = {tm_s(sl, 0), tm_s(sl, 1)};
      const uint32_t tpb[2] = {tm_p(sl, 0), tm_p(sl, 1)};
      const uint32_t to = tm_o(sl);
      const uint32_t sf[2] = {s_full(sl, 0), s_full(sl, 1)};
      const uint32_t pf[2] = {p_full(sl, 0), p_full(sl, 1)};
      const uint32_t fr[2] = {s_free(sl, 0), s_free(sl, 1)};
      const uint32_t pd[2] = {pv_done(sl, 0), pv_done(sl, 1)};
      uint32_t php0 = 0, php1 = 0;   // parity of the current p_full / s_free phase per S buffer
      const int last_keys = p.lk - (nch - 1) * kBK;
      const int last_ksteps = (last_keys + 15) >> 4;
      auto issue_s = [&](int j, bool first) {
        if (first) mbar_wait(k_full(j), 0);
        const uint32_t b = (uint32_t)j & 1u;
        const uint32_t idesc = (p.short_last && j == nch - 1) ? idesc_s16 : idesc_s;
        const uint64_t adv = (uint64_t)((uint32_t)j * k_chunk_step);
#pragma unroll
        for (int k = 0; k < KS; ++k)
          umma_bf16_ss(b ? ts[1] : ts[0], qd[k], kd[k] + adv, idesc, k != 0 ? 1u : 0u);
        umma_commit(b ? sf[1] : sf[0]);
        if (j == nch - 1) umma_commit(q_free(sl));   // the Q tile has been read for the last time
      };
#ifdef LM2A_ATTN_TIMING
      long long t_pw = 0, t_fw = 0;
      const long long t_begin = clock64();
#endif
      for (int it = 0; it < slot_tiles; ++it) {
        const bool first = it == 0;
        mbar_wait(q_full(sl), (uint32_t)it & 1u);
        tc_fence_after_sync();
        // both S buffers are free at a tile boundary: every chunk of the previous tile was
        // loaded before its P arrived, and all P.V products have been issued
        issue_s(0, first);
        if (kSepP && nch > 1) issue_s(1, first);
#pragma unroll 1
        for (int j = 0; j < nch; ++j) {
          const uint32_t b = (uint32_t)j & 1u;
          const uint32_t par = b ? php1 : php0;
          if (b) php1 ^= 1u; else php0 ^= 1u;
          if (kSepP) {
            // S_{j+2} as soon as the softmax warps have pulled S_j out of its buffer
            if (j + 2 < nch) {
              LM2A_T0(tf);
              mbar_wait(b ? fr[1] : fr[0], par);
              LM2A_TACC(t_fw, tf);
              tc_fence_after_sync();
              issue_s(j + 2, first);
            }
          } else if (j + 1 < nch) {
            issue_s(j + 1, first);   // buffer (j+1)&1: P_{j-1} V_{j-1} was issued before (in order)
          }
          if (!COND && first) mbar_wait(v_full(j), 0);
          // first P.V of a tile overwrites O: the previous tile's O has been read out of TMEM
          if (j == 0 && it > 0) mbar_wait(o_free(sl), (uint32_t)(it - 1) & 1u);
          LM2A_T0(tp0);
          mbar_wait(b ? pf[1] : pf[0], par);
          LM2A_TACC(t_pw, tp0);
          tc_fence_after_sync();
          const int ksteps = j == nch - 1 ? last_ksteps : kBK / 16;
          const uint64_t vd = vd0 + (uint64_t)((uint32_t)j * v_chunk_step);
          const uint32_t pa = b ? tpb[1] : tpb[0];
          // P_j: packed bf16 pairs, 8 columns per 16 keys
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            if (k < ksteps)
              umma_bf16_ts(to, pa + 8u * k, vd + (uint64_t)(k * v_k_step), idesc_pv,
                           (j | k) != 0 ? 1u : 0u);
          umma_commit(b ? pd[1] : pd[0]);
          if (j == nch - 1) umma_commit(o_full(sl));
        }
      }
#ifdef LM2A_ATTN_TIMING
      if (sl == 0) {
        atomicAdd(&g_attn_timing[3], (unsigned long long)t_pw);
        atomicAdd(&g_attn_timing[4], (unsigned long long)t_fw);
        atomicAdd(&g_attn_timing[5], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_attn_timing[7], 1ull);
      }
#endif
    }
  } else {
    if constexpr (kRebalance) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------------------------------------------ softmax
    const int sl = warp >> 2;            // tile slot of this warpgroup
    const int wq = warp & 3;             // TMEM lane quadrant
    const int row = wq * 32 + lane;      // query row of the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    uint32_t phs0 = 0, phs1 = 0;   // parity of the current s_full / pv_done phase per S buffer
    const int my_slot_tiles = my_tiles > sl ? (my_tiles - sl + kSlots - 1) / kSlots : 0;
#ifdef LM2A_ATTN_TIMING
    long long t_sw = 0, t_loop = 0, t_pb = 0, t_ld = 0, t_st = 0, n_chunks = 0;
#endif
    for (int it = 0; it < my_slot_tiles; ++it) {
      const int ti = t_start + (kSlots * it + sl) * t_stride;
      const TileInfo t = decode_tile(p, hpg, ti);
      const int g = row / t.rb;
      const int tq = t.q0 + (row - g * t.rb);
      const bool valid = g < t.g && tq < p.t_valid;
      const bool active = __any_sync(0xffffffffu, valid);
      float m_used = -INFINITY, l_run = 0.f;

      // MASK: only the last chunk of a row can hold keys past Lk; the other chunks carry no
      // masking code at all (the compiler turns a guarded mask into 64 unconditional selects)
      auto softmax_chunk = [&](auto nc_tag, auto mask_tag, int j) {
        constexpr int NC = decltype(nc_tag)::value;   // keys of this chunk: 64 or 16
        constexpr bool MASK = decltype(mask_tag)::value;
        const int b = j & 1;
        // parity of this chunk's phase on the barriers of buffer b (one phase per chunk), and of
        // the other buffer's latest phase (chunk j - 1)
        const uint32_t par = b ? phs1 : phs0;
        const uint32_t par_prev = (b ? phs0 : phs1) ^ 1u;
        if (b) phs1 ^= 1u; else phs0 ^= 1u;
        LM2A_T0(tsw);
        mbar_wait(s_full(sl, b), par);
        LM2A_TACC(t_sw, tsw);
        if (active) {
          tc_fence_after_sync();
          LM2A_T0(tld);
          float s[NC];
          if constexpr (NC == 64) {
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(tm_s(sl, b) + lane_off, v0);
            tmem_ld_32x32(tm_s(sl, b) + lane_off + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              s[c] = __uint_as_float(v0[c]);
              s[32 + c] = __uint_as_float(v1[c]);
            }
          } else {
            uint32_t v0[16];
            tmem_ld_32x16(tm_s(sl, b) + lane_off, v0);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) s[c] = __uint_as_float(v0[c]);
          }
          LM2A_TACC(t_ld, tld);
          if (kSepP) {   // the S buffer may be overwritten (S_{j+2}) from here on
            tc_fence_before_sync();
            mbar_arrive(s_free(sl, b));
          }
          if constexpr (MASK) {
            const int keys = p.lk - j * kBK;
#pragma unroll
            for (int c = 0; c < NC; ++c)
              if (c >= keys) s[c] = -INFINITY;
          }
          float mxa[4] = {s[0], s[1], s[2], s[3]};
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
            corr = res_ex2(m_used - mx);  // first chunk: exp2(-inf) = 0
            m_used = mx;
            l_run *= corr;
          }
          if (j > 0 && __any_sync(0xffffffffu, grow)) {
            // O holds P_0 V_0 .. P_{j-1} V_{j-1}: wait for the last of them, then correct it
            mbar_wait(pv_done(sl, b ^ 1), par_prev);
            tc_fence_after_sync();
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 32) {
              uint32_t ov[32];
              tmem_ld_32x32(tm_o(sl) + lane_off + c0, ov);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * corr);
              res_tmem_st32(tm_o(sl) + lane_off + c0, ov);
            }
          }
          uint32_t pk[NC / 2];
          l_run += softmax_probs<NC, MASK>(s, m_used, pk);   // exp2 over the XU and FMA pipes
          // a dedicated P buffer is free once P_{j-2} V_{j-2} (the previous phase of this
          // buffer's barrier) has completed; within a tile only - tiles end on o_full
          LM2A_T0(tpb);
          if (kSepP && j >= 2) mbar_wait(pv_done(sl, b), par ^ 1u);
          LM2A_TACC(t_pb, tpb);
          LM2A_T0(tst);
          if constexpr (NC == 64) res_tmem_st32(tm_p(sl, b) + lane_off, pk);
          else res_tmem_st8(tm_p(sl, b) + lane_off, pk);
          res_tmem_st_wait();
          LM2A_TACC(t_st, tst);
          tc_fence_before_sync();
        } else if (kSepP) {
          // a warp without valid rows only keeps the barriers' phases in step. It must not run
          // ahead of them: S_{j+2} is ready before P_j V_j has been issued, so without this wait
          // its arrival for chunk j + 2 could land in the p_full phase of chunk j
          mbar_arrive(s_free(sl, b));
          if (j >= 2) mbar_wait(pv_done(sl, b), par ^ 1u);
        }
        mbar_arrive(p_full(sl, b));
      };
      using W64 = std::integral_constant<int, kBK>;
      using W16 = std::integral_constant<int, 16>;
      LM2A_T0(tloop);
#pragma unroll 1
      for (int j = 0; j < nch - 1; ++j) softmax_chunk(W64{}, std::false_type{}, j);
      if (p.short_last) softmax_chunk(W16{}, std::true_type{}, nch - 1);
      else softmax_chunk(W64{}, std::true_type{}, nch - 1);
      LM2A_TACC(t_loop, tloop);
#ifdef LM2A_ATTN_TIMING
      n_chunks += nch;
#endif

      // ---- finalise: O / l -> bf16 slab
      mbar_wait(o_full(sl), (uint32_t)it & 1u);
      if (active) {
        tc_fence_after_sync();
        const float inv = 1.0f / l_run;
        __nv_bfloat16* op = p.o + ((size_t)r * p.tp + tq) * p.o_ld + stream * p.e +
                            (hg + t.h0 + min(g, t.g - 1)) * DH;
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t ov[32];
          tmem_ld_32x32(tm_o(sl) + lane_off + c0, ov);
          tmem_ld_wait();
          if (valid) {
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
        tc_fence_before_sync();
      }
      mbar_arrive(o_free(sl));
    }
#ifdef LM2A_ATTN_TIMING
    if (warp == 0 && lane == 0) {
      atomicAdd(&g_attn_timing[0], (unsigned long long)t_sw);
      atomicAdd(&g_attn_timing[1], (unsigned long long)t_loop);
      atomicAdd(&g_attn_timing[2], (unsigned long long)t_pb);
      atomicAdd(&g_attn_timing[6], (unsigned long long)n_chunks);
      atomicAdd(&g_attn_timing[8], (unsigned long long)t_ld);
      atomicAdd(&g_attn_timing[9], (unsigned long long)t_st);
    }
#endif
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == kProdWarp + 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 256u * SLOTS);
  }
}

// ------------------------------------------------------------------ host side
int res_encode_map(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
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
               "cross_attn (resident): cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu "
               "outer=%llu pitch=%llu box=%ux%u",
               (int)res, base, (unsigned long long)inner, (unsigned long long)outer,
               (unsigned long long)pitch_elems, box_inner, box_outer);
  return 0;
}

struct ResPlan {
  ResArgs a;
  int smem_bytes;
  int groups_y;
};

// Shared-memory footprint of the resident operands; false when the keys do not fit (the
// caller falls back to the streaming kernel). No device query: usable for planning on the host.
template <int DH, bool COND, int SLOTS>
bool res_fits(int lk, ResPlan* out) {
  using C = ResCfg<DH, COND>;
  ResArgs& a = out->a;
  a.nchunks = (lk + kBK - 1) / kBK;
  if (lk <= 0 || a.nchunks > kMaxChunks) return false;
  a.short_last = (lk - (a.nchunks - 1) * kBK) <= 16 ? 1 : 0;
  a.k_rows = a.nchunks * kBK;   // the last chunk's box is always 64 rows
  const long long smem = (long long)SLOTS * C::kQBytes +
                         (long long)C::kPanels * a.k_rows * C::kRowBytes +
                         (long long)a.nchunks * C::kVChunkBytes + 8ll * (2 * kMaxChunks + 24) + 16;
  if (smem > 227 * 1024) return false;
  out->smem_bytes = (int)smem;
  return true;
}

// Tile enumeration and the CTA split of one launch.
template <int DH, bool COND, int SLOTS>
bool res_plan(int rows, int n_streams, int heads, int t_valid, int lk, ResPlan* out) {
  if (!res_fits<DH, COND, SLOTS>(lk, out)) return false;
  ResArgs& a = out->a;
  const int hpg = COND ? heads : 1;
  a.n_full = t_valid / kBQ;
  const int rem = t_valid - a.n_full * kBQ;
  a.rb = kBQ;
  a.gpt = 1;
  a.n_tail = 0;
  if (rem > 0) {
    a.rb = rem <= 16 ? 16 : (rem <= 32 ? 32 : (rem <= 64 ? 64 : 128));
    a.gpt = kBQ / a.rb < hpg ? kBQ / a.rb : hpg;
    a.n_tail = (hpg + a.gpt - 1) / a.gpt;
  }
  const int n_tiles = hpg * a.n_full + a.n_tail;
  out->groups_y = COND ? n_streams : n_streams * heads;
  // CTA split: waves x (rounds of two tiles + the fixed cost of a CTA: prologue, K/V load)
  static const int forced_nz = [] {
    const char* e = getenv("LM2A_ATTN_NZ");
    return e != nullptr ? atoi(e) : 0;
  }();
  const long long groups = (long long)rows * out->groups_y;
  const int sms = (num_sms() > 0 ? num_sms() : 148) * (SLOTS == 1 ? 2 : 1);   // resident CTAs
  int best = 1;
  double best_cost = -1.0;
  for (int nz = 1; nz <= n_tiles; ++nz) {
    const long long waves = (groups * nz + sms - 1) / sms;
    const int per = (n_tiles + nz - 1) / nz;
    const double cost = (double)waves * ((per + SLOTS - 1) / SLOTS + 0.3) * (SLOTS == 1 ? 0.5 : 1.0);
    if (best_cost < 0 || cost < best_cost - 1e-9) {
      best_cost = cost;
      best = nz;
    }
  }
  a.nz = (forced_nz > 0 && forced_nz <= n_tiles) ? forced_nz : best;
  a.split_tail = 0;
  // Tail tiles in a CTA of their own (see the kernel): worth it when the CTAs of the full tiles
  // form a single wave that leaves SMs idle, and the light CTAs (a tail tile costs ~0.3 of a
  // full round; fixed cost of a CTA as above) get through those SMs within the main CTAs' time.
  // LM2A_ATTN_SPLIT_TAIL: 0 never (default), 1 cost model, 2 whenever possible (tests); read per
  // call. Measured neutral on B200 at the production shape (T = 129, 8 heads, B = 32: 43.6 vs
  // 43.3 us): the tail tile spreads its 8 valid rows (one per head, 16-row sub-blocks) over all
  // four softmax warps and a chunk's time is latency, not work - a "light" CTA takes as long as
  // a full round, and four waves of them on the 20 idle SMs last as long as the main CTAs.
  const char* sev = getenv("LM2A_ATTN_SPLIT_TAIL");
  const int split_mode = sev != nullptr ? atoi(sev) : 0;
  const int n_main = hpg * a.n_full;
  if (split_mode == 2) best_cost = 1e30;
  if (split_mode != 0 && forced_nz == 0 && SLOTS > 1 && a.n_tail > 0 && n_main > 0) {
    for (int nzm = 1; nzm <= n_main; ++nzm) {
      const long long main_ctas = groups * nzm;
      if (main_ctas >= sms) break;
      const long long free_sms = sms - main_ctas;
      const int per = (n_main + nzm - 1) / nzm;
      const double main_cost = (per + SLOTS - 1) / SLOTS + 0.3;
      const double light = 0.3 * ((a.n_tail + SLOTS - 1) / SLOTS) + 0.3;
      const double light_cost = (double)((groups + free_sms - 1) / free_sms) * light;
      const double cost = main_cost > light_cost ? main_cost : light_cost;
      if (cost < best_cost - 1e-9) {
        best_cost = cost;
        a.nz = nzm + 1;
        a.split_tail = 1;
      }
    }
  }
  return true;
}

template <int DH, bool COND, int SLOTS>
int launch_res(cudaStream_t st, const ResPlan& pl, const void* q, int q_ld, const void* k_m,
               const void* vt_m, const void* k_t, const void* vt_t, int k_ld, int vt_ld, int slots,
               int rows, int tp, int lk, int e, int ekv, int n_streams) {
  using C = ResCfg<DH, COND>;
  auto kern = cross_attn_res_kernel<DH, COND, SLOTS>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    LM2A_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
  CUtensorMap tq, tkm, tkt, tvm, tvt;
  if (res_encode_map(&tq, q, (uint64_t)n_streams * e, (uint64_t)rows * tp, (uint64_t)q_ld, C::kPW,
                     kQBoxRows, C::kSw64))
    return 1;
  const uint64_t kcols = COND ? (uint64_t)DH : (uint64_t)e;
  if (res_encode_map(&tkm, k_m, kcols, (uint64_t)slots * lk, (uint64_t)k_ld, C::kPW, kBK,
                     C::kSw64) ||
      res_encode_map(&tkt, k_t, kcols, (uint64_t)slots * lk, (uint64_t)k_ld, C::kPW, kBK, C::kSw64))
    return 1;
  if (COND) {
    tvm = tkm;
    tvt = tkt;
  } else {
    if (res_encode_map(&tvm, vt_m, (uint64_t)lk, (uint64_t)slots * ekv, (uint64_t)vt_ld, kBK, DH,
                       false) ||
        res_encode_map(&tvt, vt_t, (uint64_t)lk, (uint64_t)slots * ekv, (uint64_t)vt_ld, kBK, DH,
                       false))
      return 1;
  }
  dim3 grid(pl.groups_y, rows, pl.a.nz);
  LM2A_CUDA_OK(launch_kernel(kern, grid, dim3(res_threads(SLOTS)), (size_t)pl.smem_bytes, st, tq, tkm, tkt,
                             tvm, tvt, pl.a));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace

// Resident-K/V fast path of lm2a_cross_attn_streams_bf16: returns -1 when the shape is not
// covered (head dim other than 32 / 64, keys that do not fit in shared memory, or
// LM2A_ATTN_RESIDENT=0), else the launch status.
int cross_attn_resident(cudaStream_t st, const void* q, int q_ld, void* o, int o_ld,
                        const void* k_m, const void* vt_m, const void* k_t, const void* vt_t,
                        int k_ld, int vt_ld, const int32_t* kv_slot, int slots, int rows, int tp,
                        int t_valid, int lk, int e, int heads, int n_streams) {
  // Opt-in (LM2A_ATTN_RESIDENT=1). Measured on B200 at the production shapes (B = 32): 109 vs
  // 90 us at d_h = 32 (T = 516) and 90 vs 66 us at d_h = 64 (T = 258) against the streaming
  // kernel. Both are bound by the softmax warps (64 MUFU.EX2 + ~290 other instructions per
  // thread and 64-key chunk: 1230 of the 1570 cycles a chunk takes, tools/attn_probe.py), and the
  // streaming kernel's two CTAs per SM hide each other's prologue, K/V latency and the nearly
  // empty tail tiles (T mod 128 = 4 / 2 rows), which one resident CTA per SM cannot. The
  // resident design pays where the operand is shared by many tiles: COND mode below.
  const char* ev = getenv("LM2A_ATTN_RESIDENT");   // read per call: tests toggle it
  if (!(ev != nullptr && ev[0] == '1')) return -1;
  const int dh = e / heads;
  ResPlan pl{};
  bool ok = false;
  // d_h = 32: two single-slot CTAs per SM when two operand sets fit (else one two-slot CTA)
  bool two_ctas = false;
  if (dh == 32) {
    ResPlan probe{};
    two_ctas = res_fits<32, false, 1>(lk, &probe) && 2 * probe.smem_bytes + 4096 <= 227 * 1024;
    ok = two_ctas ? res_plan<32, false, 1>(rows, n_streams, heads, t_valid, lk, &pl)
                  : res_plan<32, false, 2>(rows, n_streams, heads, t_valid, lk, &pl);
  } else if (dh == 64) {
    ok = res_plan<64, false, 2>(rows, n_streams, heads, t_valid, lk, &pl);
  }
  if (!ok) return -1;
  pl.a.o = reinterpret_cast<__nv_bfloat16*>(o);
  pl.a.o_ld = o_ld;
  pl.a.kv_slot = kv_slot;
  pl.a.tp = tp;
  pl.a.t_valid = t_valid;
  pl.a.lk = lk;
  pl.a.heads = heads;
  pl.a.e = e;
  pl.a.ekv = e;
  if (dh == 32 && two_ctas)
    return launch_res<32, false, 1>(st, pl, q, q_ld, k_m, vt_m, k_t, vt_t, k_ld, vt_ld, slots, rows,
                                    tp, lk, e, e, n_streams);
  if (dh == 32)
    return launch_res<32, false, 2>(st, pl, q, q_ld, k_m, vt_m, k_t, vt_t, k_ld, vt_ld, slots, rows,
                                    tp, lk, e, e, n_streams);
  return launch_res<64, false, 2>(st, pl, q, q_ld, k_m, vt_m, k_t, vt_t, k_ld, vt_ld, slots, rows,
                                  tp, lk, e, e, n_streams);
}

}  // namespace lm2a

#ifdef LM2A_ATTN_TIMING
extern "C" int lm2a_attn_timing_read(unsigned long long* out16) {
  unsigned long long zero[16] = {};
  if (cudaMemcpyFromSymbol(out16, lm2a::g_attn_timing, sizeof(zero)) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(lm2a::g_attn_timing, zero, sizeof(zero)) != cudaSuccess;
}
#endif

extern "C" int lm2a_cross_attn_cond_supported(int32_t lk) {
  using namespace lm2a;
  ResPlan pl{};
  return res_fits<128, true, 2>(lk, &pl) ? 1 : 0;
}

extern "C" int lm2a_cross_attn_cond_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                                         int32_t o_ld, const void* cond_motion,
                                         const void* cond_text, int32_t cond_ld,
                                         const int32_t* kv_slot, int32_t slots, int32_t rows,
                                         int32_t tp, int32_t t_valid, int32_t lk, int32_t heads,
                                         int32_t n_streams) {
  using namespace lm2a;
  constexpr int DH = 128;
  LM2A_REQUIRE(n_streams == 1 || n_streams == 2, "cross_attn_cond: n_streams=%d (1 or 2)",
               n_streams);
  LM2A_REQUIRE(q && o && cond_motion && cond_text && kv_slot, "cross_attn_cond: null pointer");
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && slots > 0 && tp > 0 && t_valid > 0 &&
                   t_valid <= tp && lk > 0 && heads > 0 && heads <= 64,
               "cross_attn_cond: bad geometry");
  const int e = heads * DH;
  LM2A_REQUIRE(q_ld % 8 == 0 && o_ld % 8 == 0 && cond_ld % 8 == 0 && q_ld >= n_streams * e &&
                   o_ld >= n_streams * e && cond_ld >= DH,
               "cross_attn_cond: bad pitches (q_ld=%d o_ld=%d cond_ld=%d, %d heads of 128)", q_ld,
               o_ld, cond_ld, heads);
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(o) |
                 reinterpret_cast<uintptr_t>(cond_motion) |
                 reinterpret_cast<uintptr_t>(cond_text)) & 15) == 0,
               "cross_attn_cond: tensors must be 16-byte aligned");
  ResPlan pl{};
  LM2A_REQUIRE((res_plan<DH, true, 2>(rows, n_streams, heads, t_valid, lk, &pl)),
               "cross_attn_cond: %d keys do not fit in shared memory (see "
               "lm2a_cross_attn_cond_supported)", lk);
  pl.a.o = reinterpret_cast<__nv_bfloat16*>(o);
  pl.a.o_ld = o_ld;
  pl.a.kv_slot = kv_slot;
  pl.a.tp = tp;
  pl.a.t_valid = t_valid;
  pl.a.lk = lk;
  pl.a.heads = heads;
  pl.a.e = e;
  pl.a.ekv = 0;
  return launch_res<DH, true, 2>(reinterpret_cast<cudaStream_t>(stream), pl, q, q_ld, cond_motion,
                              nullptr, cond_text, nullptr, cond_ld, 0, slots, rows, tp, lk, e, 0,
                              n_streams);
}
