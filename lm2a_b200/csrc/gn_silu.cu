// GroupNorm (biased variance, per-channel affine) + SiLU over a channels-last
// bf16 slab. One CTA per (clip-row, group): the (t_valid x C/G) sub-slab is streamed
// from HBM/L2 once into shared memory with cp.async (every 16-byte request of the CTA
// is in flight at once, no registers held), mean and variance are computed two-pass in
// fp32 from that copy, and the normalised/activated result is written once.
// Replaces nn.GroupNorm + nn.SiLU of the reference (models/unet1d_ultimate.py:
// 91-95,136-137,146-147,362-363; eps = 1e-5, F.group_norm semantics).
// HBM-bound: algorithmic bytes = 2 B read + 2 B written per element. Small CTAs
// (256 threads, 33-66 KB of shared memory) keep 3-6 sub-slabs in flight per SM.
#include "../../include/lm2a_b200.h"
#include "common.cuh"

namespace lm2a {
namespace {

constexpr int kGnThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < kGnThreads / 32; ++w) s += red[w];
  return s;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <bool CACHED>
__global__ void __launch_bounds__(kGnThreads)
gn_silu_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y,
               int y_ld, const float* __restrict__ gamma, const float* __restrict__ beta,
               int tp, int t_valid, int cg, float eps, int apply_silu) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint4 cache[];
  __shared__ float red[kGnThreads / 32];
  const int g = blockIdx.x, r = blockIdx.y;
  const int vpr = cg >> 3;              // 16-byte vectors per slot
  const int nvec = t_valid * vpr;
  const int cv = threadIdx.x % vpr;     // fixed per thread: kGnThreads % vpr == 0
  const int tstep = kGnThreads / vpr;
  const int t0 = threadIdx.x / vpr;
  const size_t row_base = (size_t)r * tp;
  const __nv_bfloat16* xg = x + (size_t)g * cg + cv * 8;
  __nv_bfloat16* yg = y + (size_t)g * cg + cv * 8;

  if (CACHED) {
    const uint32_t cbase = smem_u32(cache);
    for (int t = t0, i = threadIdx.x; t < t_valid; t += tstep, i += kGnThreads)
      cp_async16(cbase + (uint32_t)i * 16u, xg + (row_base + t) * x_ld);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  // pass 1: sum
  float s = 0.f;
  for (int t = t0, i = threadIdx.x; t < t_valid; t += tstep, i += kGnThreads) {
    const uint4 q = CACHED ? cache[i]
                           : __ldg(reinterpret_cast<const uint4*>(xg + (row_base + t) * x_ld));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack_bf16x2(w[e]);
      s += a.x + a.y;
    }
  }
  const float inv_n = 1.0f / (float)(nvec * 8);
  const float mean = block_sum(s, red) * inv_n;

  // pass 2: centred sum of squares
  float ss = 0.f;
  for (int t = t0, i = threadIdx.x; t < t_valid; t += tstep, i += kGnThreads) {
    const uint4 q = CACHED ? cache[i]
                           : __ldg(reinterpret_cast<const uint4*>(xg + (row_base + t) * x_ld));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = unpack_bf16x2(w[e]);
      const float d0 = a.x - mean, d1 = a.y - mean;
      ss += d0 * d0 + d1 * d1;
    }
  }
  const float var = block_sum(ss, red) * inv_n;
  const float rstd = rsqrtf(var + eps);

  float ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gm = __ldg(gamma + g * cg + cv * 8 + e) * rstd;
    ga[e] = gm;
    be[e] = __ldg(beta + g * cg + cv * 8 + e) - mean * gm;
  }

  // pass 3: normalise + SiLU, write; slots t >= t_valid are written as zero
  for (int t = t0, i = threadIdx.x; t < tp; t += tstep, i += kGnThreads) {
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (t < t_valid) {
      const uint4 q = CACHED ? cache[i]
                             : __ldg(reinterpret_cast<const uint4*>(xg + (row_base + t) * x_ld));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      uint32_t ow[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = unpack_bf16x2(w[e]);
        float v0 = fmaf(a.x, ga[2 * e], be[2 * e]);
        float v1 = fmaf(a.y, ga[2 * e + 1], be[2 * e + 1]);
        if (apply_silu) {
          v0 = silu_f(v0);
          v1 = silu_f(v1);
        }
        ow[e] = pack_bf16x2(v0, v1);
      }
      o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    *reinterpret_cast<uint4*>(yg + (row_base + t) * y_ld) = o;
  }
}

// ---------------------------------------------------------------------------
// GroupNorm + SiLU as a pure streaming pass: the per-(clip-row, group) sums were already
// produced by the epilogue of the kernel that wrote x (lm2a_conv1d_bf16 / lm2a_bias_add_bf16
// `stats`: exact 64-bit fixed-point sums, 2^24 / 2^20 scaled), so this kernel only derives
// mean / rstd and applies y = SiLU(x * a_c + b_c). One CTA = a few consecutive slots of one
// clip-row, all channels; every thread owns one 8-channel vector column (4 slots of it).
// The UNet plan applies the same transform inside the consuming conv (lm2a_conv_desc.in_gn_*);
// this stand-alone pass serves callers that need the normalised slab itself.
constexpr int kApplyVec = 4;  // 16-byte vectors per thread and pass

__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y,
                int y_ld, const long long* __restrict__ stats, int stats_pitch,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                int tp, int t_valid, int c, int groups, float eps, int apply_silu) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float2 s_sc[64];   // (rstd, -mean * rstd) per group
  const int r = blockIdx.y;
  const int cg = c / groups;
  // thread -> (vector column, slot) mapping; the first pass's loads are issued before the
  // statistics are read so their latency overlaps
  const int vpr = c >> 3;
  const int lanes = vpr < kGnThreads ? vpr : kGnThreads;
  const int passes = (vpr + kGnThreads - 1) / kGnThreads;
  const int tstep = kGnThreads / lanes;
  const int cvl = threadIdx.x % lanes, tl = threadIdx.x / lanes;
  // vector-column counts that do not divide the CTA (c = 768, 1536: legacy UNet1D concat
  // slabs) leave the last kGnThreads % lanes threads / the tail of the last pass idle
  const bool active = tl < tstep;
  const int t_begin = blockIdx.x * (kApplyVec * tstep);
  const size_t row_base = (size_t)r * tp;
  uint4 q[kApplyVec];
#pragma unroll
  for (int k = 0; k < kApplyVec; ++k) {
    const int t = t_begin + tl + k * tstep;
    q[k] = make_uint4(0u, 0u, 0u, 0u);
    if (active && t < t_valid)
      q[k] = __ldg(reinterpret_cast<const uint4*>(x + (row_base + t) * x_ld + cvl * 8));
  }
  if ((int)threadIdx.x < groups) {
    const int g = threadIdx.x;
    const long long* sp = stats + ((size_t)r * stats_pitch + g) * 2;
    const double inv_n = 1.0 / ((double)cg * (double)t_valid);
    s_sc[g] = gn_rstd_cm(__ldcg(sp), __ldcg(sp + 1), inv_n, eps);
  }
  __syncthreads();

  for (int pass = 0; pass < passes; ++pass) {
    const int cv = pass * kGnThreads + cvl;
    if (!active || cv >= vpr) break;
    if (pass > 0) {
#pragma unroll
      for (int k = 0; k < kApplyVec; ++k) {
        const int t = t_begin + tl + k * tstep;
        if (t < t_valid)
          q[k] = __ldg(reinterpret_cast<const uint4*>(x + (row_base + t) * x_ld + cv * 8));
      }
    }
    const int g = (cv * 8) / cg;
    // y = SiLU(v), v = u * gamma + beta, u = x * rstd - mean * rstd, evaluated as
    // vh = u * (gamma/2) + beta/2, y = vh * tanh(vh) + vh (the same arithmetic as the conv's
    // operand transform: the two paths are bit-identical)
    const float rstd = s_sc[g].x, cm = s_sc[g].y;
    const float hs = apply_silu ? 0.5f : 1.0f;
    float ga[8], be[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ga[e] = __ldg(gamma + cv * 8 + e) * hs;
      be[e] = __ldg(beta + cv * 8 + e) * hs;
    }
#pragma unroll
    for (int k = 0; k < kApplyVec; ++k) {
      const int t = t_begin + tl + k * tstep;
      if (t >= tp) break;
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (t < t_valid) {
        const uint32_t w[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
        uint32_t ow[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = unpack_bf16x2(w[e]);
          float v0 = fmaf(fmaf(a.x, rstd, cm), ga[2 * e], be[2 * e]);
          float v1 = fmaf(fmaf(a.y, rstd, cm), ga[2 * e + 1], be[2 * e + 1]);
          if (apply_silu) {
            v0 = silu_from_half(v0);
            v1 = silu_from_half(v1);
          }
          ow[e] = pack_bf16x2(v0, v1);
        }
        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      *reinterpret_cast<uint4*>(y + (row_base + t) * y_ld + cv * 8) = o;
    }
  }
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_gn_apply_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                                  int32_t y_ld, const void* stats, int32_t stats_pitch,
                                  const float* gamma, const float* beta, int32_t rows,
                                  int32_t tp, int32_t t_valid, int32_t c, int32_t groups,
                                  float eps, int32_t apply_silu) {
  using namespace lm2a;
  LM2A_REQUIRE(x && y && stats && gamma && beta, "gn_apply: null pointer");
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && tp > 0 && t_valid > 0 && t_valid <= tp,
               "gn_apply: bad geometry");
  LM2A_REQUIRE(groups > 0 && groups <= 64 && c % groups == 0 && stats_pitch >= groups,
               "gn_apply: c=%d not divisible by groups=%d (<= 64, stats pitch %d)", c, groups,
               stats_pitch);
  const int vpr = c / 8;
  LM2A_REQUIRE(c % 8 == 0 && vpr > 0 && (c / groups) % 8 == 0,
               "gn_apply: c=%d / channels per group must be positive multiples of 8", c);
  LM2A_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && x_ld >= c && y_ld >= c,
               "gn_apply: ld must be a multiple of 8 and >= c");
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                 reinterpret_cast<uintptr_t>(stats)) & 15) == 0,
               "gn_apply: slabs and statistics must be 16-byte aligned");
  const int lanes = vpr < kGnThreads ? vpr : kGnThreads;
  const int slots_per_cta = kApplyVec * (kGnThreads / lanes);
  dim3 grid((tp + slots_per_cta - 1) / slots_per_cta, rows);
  LM2A_CUDA_OK(launch_kernel(gn_apply_kernel, dim3(grid), dim3(kGnThreads), 0,
                             reinterpret_cast<cudaStream_t>(stream),
                             reinterpret_cast<const __nv_bfloat16*>(x), x_ld,
                             reinterpret_cast<__nv_bfloat16*>(y), y_ld,
                             reinterpret_cast<const long long*>(stats), stats_pitch, gamma, beta,
                             tp, t_valid, c, groups, eps, apply_silu));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_gn_silu_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                                 int32_t y_ld, const float* gamma, const float* beta,
                                 int32_t rows, int32_t tp, int32_t t_valid, int32_t c,
                                 int32_t groups, float eps, int32_t apply_silu) {
  using namespace lm2a;
  LM2A_REQUIRE(x && y && gamma && beta, "gn_silu: null pointer");
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && tp > 0 && t_valid > 0 && t_valid <= tp,
               "gn_silu: bad geometry");
  LM2A_REQUIRE(groups > 0 && c % groups == 0, "gn_silu: c=%d not divisible by groups=%d", c,
               groups);
  const int cg = c / groups;
  LM2A_REQUIRE(cg % 8 == 0 && kGnThreads % (cg / 8) == 0 && cg / 8 <= kGnThreads,
               "gn_silu: channels per group (%d) must be 8 * a divisor of %d", cg, kGnThreads);
  LM2A_REQUIRE(x_ld % 8 == 0 && y_ld % 8 == 0 && x_ld >= c && y_ld >= c,
               "gn_silu: ld must be a multiple of 8 and >= c");
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
               "gn_silu: slabs must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t cache_bytes = (size_t)t_valid * cg * 2;
  dim3 grid(groups, rows);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  if (cache_bytes <= 200 * 1024) {
    static bool configured[kMaxDevices] = {};
    if (first_use_on_device(configured))
      LM2A_CUDA_OK(cudaFuncSetAttribute(gn_silu_kernel<true>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
    LM2A_CUDA_OK(launch_kernel(gn_silu_kernel<true>, dim3(grid), dim3(kGnThreads), cache_bytes, st, 
        xp, x_ld, yp, y_ld, gamma, beta, tp, t_valid, cg, eps, apply_silu));
  } else {
    LM2A_CUDA_OK(launch_kernel(gn_silu_kernel<false>, dim3(grid), dim3(kGnThreads), 0, st, xp, x_ld, yp, y_ld, gamma, beta, tp,
                                                       t_valid, cg, eps, apply_silu));
  }
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
