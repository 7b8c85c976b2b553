// Cross-attention core softmax(q k^T) v for both condition streams (motion, lyrics)
// of CrossAttentionFusion (reference models/cross_attention.py:50-61, i.e. the
// bmm / softmax / bmm inside nn.MultiheadAttention). K and V are per-clip caches
// built once per clip (they do not depend on x or t); q arrives pre-scaled by
// log2(e)/sqrt(d_h) so the softmax is a bare exp2. The [Tq, Lk] probability
// matrix never leaves the SM (the reference materialises it, head-averages it
// and throws it away).
//
// v1: flash-style, 64 queries x 64 keys per step, 4 warps, mma.sync m16n8k16
// bf16 with fp32 accumulation, cp.async double-buffered K/V tiles with an XOR
// swizzle that keeps ldmatrix conflict-free.
#include "../../include/lm2a_b200.h"
#include "common.cuh"

namespace lm2a {
namespace {

constexpr int kBQ = 64;
constexpr int kBK = 64;
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int DH>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  // byte offset of 16-byte chunk `chunk` of row `row` in a [64][DH] bf16 tile
  constexpr int CPR = DH / 8;
  const int sw = (DH == 32) ? (chunk ^ ((row >> 1) & 3)) : (chunk ^ (row & 7));
  return (uint32_t)(row * CPR + sw) * 16u;
}

// rows [row0, row0+64) of a row-major bf16 matrix (pitch ld elements) -> swizzled tile;
// rows >= row_end are zero-filled.
template <int DH>
__device__ __forceinline__ void load_tile(uint32_t smem_tile, const __nv_bfloat16* base,
                                          long long ld, int row0, int row_end) {
  constexpr int CPR = DH / 8;
  constexpr int ITER = 64 * CPR / kAttnThreads;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int idx = threadIdx.x + i * kAttnThreads;
    const int row = idx / CPR, chunk = idx % CPR;
    const bool ok = row0 + row < row_end;
    const __nv_bfloat16* src = base + (long long)(ok ? row0 + row : 0) * ld + chunk * 8;
    cp_async16(smem_tile + tile_off<DH>(row, chunk), src, ok);
  }
}

template <int DH>
__global__ void __launch_bounds__(kAttnThreads)
cross_attn_kernel(const __nv_bfloat16* __restrict__ q, int q_ld, __nv_bfloat16* __restrict__ o,
                  int o_ld, const __nv_bfloat16* __restrict__ k_m,
                  const __nv_bfloat16* __restrict__ v_m, const __nv_bfloat16* __restrict__ k_t,
                  const __nv_bfloat16* __restrict__ v_t, int kv_ld,
                  const int* __restrict__ kv_slot, int tp, int t_valid, int lk, int e,
                  int heads) {
  extern __shared__ __align__(128) uint8_t attn_smem[];
  constexpr int TILE = 64 * DH * 2;
  const uint32_t sQ = smem_u32(attn_smem);
  auto sK = [&](int s) { return sQ + TILE * (1 + 2 * s); };
  auto sV = [&](int s) { return sQ + TILE * (2 + 2 * s); };

  const int r = blockIdx.z;
  const int stream = blockIdx.y / heads;
  const int h = blockIdx.y % heads;
  const int q0 = blockIdx.x * kBQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = kv_slot[r];
  const int chan = stream * e + h * DH;

  const __nv_bfloat16* qb = q + (long long)r * tp * q_ld + chan;
  const __nv_bfloat16* kb =
      (stream == 0 ? k_m : k_t) + (long long)slot * lk * kv_ld + h * DH;
  const __nv_bfloat16* vb =
      (stream == 0 ? v_m : v_t) + (long long)slot * lk * kv_ld + h * DH;

  const int ntiles = (lk + kBK - 1) / kBK;
  load_tile<DH>(sQ, qb, q_ld, q0, t_valid);
  load_tile<DH>(sK(0), kb, kv_ld, 0, lk);
  load_tile<DH>(sV(0), vb, kv_ld, 0, lk);
  cp_async_commit();

  float oacc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int it = 0; it < ntiles; ++it) {
    const int st = it & 1;
    if (it + 1 < ntiles) {
      load_tile<DH>(sK(st ^ 1), kb, kv_ld, (it + 1) * kBK, lk);
      load_tile<DH>(sV(st ^ 1), vb, kv_ld, (it + 1) * kBK, lk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // ---- S = Q K^T for this warp's 16 queries x 64 keys
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t a[4];
      ldsm_x4(sQ + tile_off<DH>(warp * 16 + (lane & 15), kk * 2 + (lane >> 4)), a);
#pragma unroll
      for (int nb = 0; nb < 8; nb += 2) {
        uint32_t b[4];
        ldsm_x4(sK(st) + tile_off<DH>(nb * 8 + (lane & 7) + ((lane >> 4) << 3),
                                      kk * 2 + ((lane >> 3) & 1)),
                b);
        mma_bf16(s[nb], a, b[0], b[1]);
        mma_bf16(s[nb + 1], a, b[2], b[3]);
      }
    }
    // ---- mask keys past lk (only the last tile can have them)
    const int key0 = it * kBK + 2 * (lane & 3);
    if (it * kBK + kBK > lk) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int j = key0 + nb * 8;
        if (j >= lk) s[nb][0] = s[nb][2] = -INFINITY;
        if (j + 1 >= lk) s[nb][1] = s[nb][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows lane/4 and lane/4 + 8)
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nb][0], s[nb][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nb][2], s[nb][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    float corr[2], rs[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      corr[i] = exp2f(m_run[i] - mx[i]);  // first tile: exp2(-inf) = 0
      m_run[i] = mx[i];
    }
    uint32_t pa[4][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = exp2f(s[nb][0] - mx[0]);
      const float p1 = exp2f(s[nb][1] - mx[0]);
      const float p2 = exp2f(s[nb][2] - mx[1]);
      const float p3 = exp2f(s[nb][3] - mx[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pa[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) l_run[i] = l_run[i] * corr[i] + rs[i];
#pragma unroll
    for (int nb = 0; nb < DH / 8; ++nb) {
      oacc[nb][0] *= corr[0];
      oacc[nb][1] *= corr[0];
      oacc[nb][2] *= corr[1];
      oacc[nb][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int nb = 0; nb < DH / 8; nb += 2) {
        uint32_t b[4];
        ldsm_x4_t(sV(st) + tile_off<DH>(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3),
                                        nb + (lane >> 4)),
                  b);
        mma_bf16(oacc[nb], pa[kk], b[0], b[1]);
        mma_bf16(oacc[nb + 1], pa[kk], b[2], b[3]);
      }
    }
    __syncthreads();  // all warps done with stage st before it is refilled
  }

  // ---- finalise: O / l, bf16 store
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  const int t_a = q0 + warp * 16 + (lane >> 2);
  const int t_b = t_a + 8;
  __nv_bfloat16* ob = o + (long long)r * tp * o_ld + chan + 2 * (lane & 3);
#pragma unroll
  for (int nb = 0; nb < DH / 8; ++nb) {
    if (t_a < t_valid)
      *reinterpret_cast<uint32_t*>(ob + (long long)t_a * o_ld + nb * 8) =
          pack_bf16x2(oacc[nb][0] * inv0, oacc[nb][1] * inv0);
    if (t_b < t_valid)
      *reinterpret_cast<uint32_t*>(ob + (long long)t_b * o_ld + nb * 8) =
          pack_bf16x2(oacc[nb][2] * inv1, oacc[nb][3] * inv1);
  }
}

template <int DH>
int launch_attn(cudaStream_t st, const void* q, int q_ld, void* o, int o_ld, const void* k_m,
                const void* v_m, const void* k_t, const void* v_t, int kv_ld,
                const int32_t* kv_slot, int rows, int tp, int t_valid, int lk, int e,
                int heads) {
  constexpr int smem = 5 * 64 * DH * 2;
  auto kern = cross_attn_kernel<DH>;
  static bool configured = false;
  if (!configured) {
    LM2A_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((t_valid + kBQ - 1) / kBQ, 2 * heads, rows);
  kern<<<grid, kAttnThreads, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), q_ld, reinterpret_cast<__nv_bfloat16*>(o), o_ld,
      reinterpret_cast<const __nv_bfloat16*>(k_m), reinterpret_cast<const __nv_bfloat16*>(v_m),
      reinterpret_cast<const __nv_bfloat16*>(k_t), reinterpret_cast<const __nv_bfloat16*>(v_t),
      kv_ld, kv_slot, tp, t_valid, lk, e, heads);
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_cross_attn_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                                    int32_t o_ld, const void* k_motion, const void* v_motion,
                                    const void* k_text, const void* v_text, int32_t kv_ld,
                                    const int32_t* kv_slot, int32_t rows, int32_t tp,
                                    int32_t t_valid, int32_t lk, int32_t e, int32_t heads) {
  using namespace lm2a;
  LM2A_REQUIRE(q && o && k_motion && v_motion && k_text && v_text && kv_slot,
               "cross_attn: null pointer");
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && tp > 0 && t_valid > 0 && t_valid <= tp && lk > 0,
               "cross_attn: bad geometry");
  LM2A_REQUIRE(heads > 0 && e % heads == 0, "cross_attn: e=%d not divisible by heads=%d", e,
               heads);
  LM2A_REQUIRE(q_ld % 8 == 0 && o_ld % 8 == 0 && kv_ld % 8 == 0 && q_ld >= 2 * e &&
                   o_ld >= 2 * e && kv_ld >= e,
               "cross_attn: bad pitches");
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(o) |
                 reinterpret_cast<uintptr_t>(k_motion) | reinterpret_cast<uintptr_t>(v_motion) |
                 reinterpret_cast<uintptr_t>(k_text) | reinterpret_cast<uintptr_t>(v_text)) & 15) == 0,
               "cross_attn: tensors must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int dh = e / heads;
  switch (dh) {
    case 32:
      return launch_attn<32>(st, q, q_ld, o, o_ld, k_motion, v_motion, k_text, v_text, kv_ld,
                             kv_slot, rows, tp, t_valid, lk, e, heads);
    case 64:
      return launch_attn<64>(st, q, q_ld, o, o_ld, k_motion, v_motion, k_text, v_text, kv_ld,
                             kv_slot, rows, tp, t_valid, lk, e, heads);
    case 128:
      return launch_attn<128>(st, q, q_ld, o, o_ld, k_motion, v_motion, k_text, v_text, kv_ld,
                              kv_slot, rows, tp, t_valid, lk, e, heads);
    default:
      LM2A_REQUIRE(false, "cross_attn: head dim %d unsupported (32, 64 or 128)", dh);
  }
  return 0;
}
