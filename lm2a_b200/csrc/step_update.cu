// The per-step update of the sampling loop as ONE kernel (reference sample.py:167-210 and the
// torch.cat([x, x]) of :162 for the next step):
//   CFG blend with both clamps -> DDPM posterior update -> noise injection, the noise drawn
//   inside the kernel (counter-based Philox4x32-10 keyed per clip) -> new x written back as fp32
//   [B, c, T] AND as the next step's bf16 input slab [copies*B, tp, ld] (what lm2a_ingest_x
//   would produce) -> the step's GroupNorm statistics arena cleared -> device timestep advanced.
// One launch replaces randn + cfg_posterior + ingest_x of the next step; a trajectory is then a
// pure replay of [UNet launches, this kernel].
//
// Noise: z[b, ch, t] at timestep s = BoxMuller(Philox4x32-10(counter = (ch * ceil(T/4) + t/4, s,
// 0, 0), key = clip_seed[b])) component t % 4. It depends on the clip's own seed, the element and
// the timestep only - not on the batch the clip is sampled in - so a clip's trajectory is
// bit-identical under any sharding of a dataset over GPUs.
// The update arithmetic keeps the reference's operation order with explicit round-to-nearest
// mul / add / sub (no FMA contraction), exactly like lm2a_cfg_posterior.
#include "../../include/lm2a_b200.h"
#include "common.cuh"

namespace lm2a {
namespace {

__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
  return v != v ? v : fminf(fmaxf(v, lo), hi);
}

// Philox4x32-10 (Salmon et al., SC'11): 10 rounds of two 32x32->64 multiplies, key bumped by
// the Weyl constants each round
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kM0, ctr.x), lo0 = kM0 * ctr.x;
    const uint32_t hi1 = __umulhi(kM1, ctr.z), lo1 = kM1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += kW0;
    key.y += kW1;
  }
  return ctr;
}

// the four standard normals of counter block (ch, t / 4) of a clip at counter word `step`:
// components 0 / 1 = Box-Muller (cos, sin) of (r.x, r.y), components 2 / 3 of (r.z, r.w)
__device__ __forceinline__ void philox_normal4(uint2 key, int ch, int tq, int t4, uint32_t step,
                                               float (&z)[4]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)(ch * t4 + tq), step, 0u, 0u), key);
  const uint32_t a[2] = {r.x, r.z}, b[2] = {r.y, r.w};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    // u in (0, 1): 24 random bits + half a step, exactly representable in fp32
    const float u1 = __fmaf_rn((float)(a[h] >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f);
    const float u2 = __fmaf_rn((float)(b[h] >> 8), 5.9604644775390625e-08f, 2.98023223876953125e-08f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    z[2 * h] = rad * cs;
    z[2 * h + 1] = rad * sn;
  }
}
// standard normal for element (ch, t)
__device__ __forceinline__ float philox_normal(uint2 key, int ch, int t, int t4, uint32_t step) {
  float z[4];
  philox_normal4(key, ch, t >> 2, t4, step, z);
  return z[t & 3];
}

// One CTA = one clip x 32 consecutive slots x all channels (the tile lm2a_ingest_x uses). A
// thread owns (channel, group of four consecutive slots) items: one Philox block yields the four
// normals it needs, x / eps move as 16-byte vectors when T is a multiple of 4; the updated tile
// is transposed through shared memory into channels-last bf16 slab rows written 16 bytes at a time.
__global__ void __launch_bounds__(256)
cfg_step_kernel(float* __restrict__ x, const float* __restrict__ eps,
                const float* __restrict__ noise, const unsigned long long* __restrict__ clip_seed,
                const float* __restrict__ sched, int64_t* __restrict__ t_dev, int n_t,
                unsigned int* __restrict__ ticket, int batch, int c, int T, float gw, int guided,
                int advance, __nv_bfloat16* __restrict__ slab, int copies, int tp, int ld,
                uint4* __restrict__ zero, long long zero_vec, float* __restrict__ eps_out) {
  pdl_wait();
  pdl_launch_dependents();
  {
    const long long nthr = (long long)gridDim.x * gridDim.y * blockDim.x;
    const long long tid = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    for (long long i = tid; i < zero_vec; i += nthr) zero[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __shared__ float tile[32][129];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * 32;
  const long long t_now = t_dev[b];
  const float4 co = __ldg(reinterpret_cast<const float4*>(sched) + t_now);
  const float coef1 = co.x, coef2 = co.y, sigma = co.z;
  const bool add_noise = t_now > 0 && (noise != nullptr || clip_seed != nullptr);
  uint2 key = make_uint2(0u, 0u);
  if (clip_seed != nullptr) {
    const unsigned long long s = clip_seed[b];
    key = make_uint2((uint32_t)s, (uint32_t)(s >> 32));
  }
  const int t4 = (T + 3) >> 2;
  const size_t clip = (size_t)c * T;
  const bool vec = (T & 3) == 0;   // rows of x / eps are 16-byte aligned
  // items: (channel, slot group): 8 groups of four slots per tile, groups fastest (coalesced)
  for (int it = threadIdx.x; it < c * 8; it += 256) {
    const int ch = it >> 3, g = it & 7;
    const int tb = t0 + g * 4;
    float xn[4] = {0.f, 0.f, 0.f, 0.f};
    if (tb < T) {
      const size_t i0 = (size_t)b * clip + (size_t)ch * T + tb;
      const int nv = min(4, T - tb);
      float xv[4] = {0.f, 0.f, 0.f, 0.f}, eu[4] = {0.f, 0.f, 0.f, 0.f}, ec[4] = {0.f, 0.f, 0.f, 0.f};
      float nz[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) {
        const float4 a = *reinterpret_cast<const float4*>(x + i0);
        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        const float4 e0 = __ldg(reinterpret_cast<const float4*>(eps + i0));
        eu[0] = e0.x; eu[1] = e0.y; eu[2] = e0.z; eu[3] = e0.w;
        if (guided) {
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(eps + i0 + (size_t)batch * clip));
          ec[0] = e1.x; ec[1] = e1.y; ec[2] = e1.z; ec[3] = e1.w;
        }
        if (add_noise && noise != nullptr) {
          const float4 n4 = __ldg(reinterpret_cast<const float4*>(noise + i0));
          nz[0] = n4.x; nz[1] = n4.y; nz[2] = n4.z; nz[3] = n4.w;
        }
      } else {
        for (int k = 0; k < nv; ++k) {
          xv[k] = x[i0 + k];
          eu[k] = __ldg(eps + i0 + k);
          if (guided) ec[k] = __ldg(eps + i0 + k + (size_t)batch * clip);
          if (add_noise && noise != nullptr) nz[k] = __ldg(noise + i0 + k);
        }
      }
      if (add_noise && noise == nullptr) philox_normal4(key, ch, tb >> 2, t4, (uint32_t)t_now, nz);
      float ev[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float e = eu[k];
        if (guided) {
          const float d = clamp_nan(__fsub_rn(ec[k], eu[k]), -5.0f, 5.0f);
          e = clamp_nan(__fadd_rn(eu[k], __fmul_rn(gw, d)), -10.0f, 10.0f);
        }
        ev[k] = e;
        xn[k] = __fadd_rn(__fmul_rn(coef1, __fsub_rn(xv[k], __fmul_rn(coef2, e))),
                          __fmul_rn(sigma, nz[k]));
        if (k >= nv) xn[k] = 0.f;
      }
      if (vec) {
        *reinterpret_cast<float4*>(x + i0) = make_float4(xn[0], xn[1], xn[2], xn[3]);
        if (eps_out != nullptr)
          *reinterpret_cast<float4*>(eps_out + i0) = make_float4(ev[0], ev[1], ev[2], ev[3]);
      } else {
        for (int k = 0; k < nv; ++k) {
          x[i0 + k] = xn[k];
          if (eps_out != nullptr) eps_out[i0 + k] = ev[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) tile[g * 4 + k][ch] = xn[k];
  }
  if (slab != nullptr) {
    __syncthreads();
    // 16-byte chunks of 8 channels: ld / 8 chunks per slot
    const int cpr = ld >> 3;
    for (int it = threadIdx.x; it < 32 * cpr; it += 256) {
      const int tl = it / cpr, cg8 = (it - tl * cpr) * 8;
      const int ts = t0 + tl;
      if (ts >= tp) continue;
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c0 = cg8 + 2 * e;
        const float v0 = (c0 < c && ts < T) ? tile[tl][c0] : 0.f;
        const float v1 = (c0 + 1 < c && ts < T) ? tile[tl][c0 + 1] : 0.f;
        w[e] = pack_bf16x2(v0, v1);
      }
      const uint4 q = make_uint4(w[0], w[1], w[2], w[3]);
      for (int k = 0; k < copies; ++k)
        *reinterpret_cast<uint4*>(slab + ((size_t)(k * batch + b) * tp + ts) * ld + cg8) = q;
    }
  }
  if (advance) {
    // the last CTA to finish steps the device-side timestep (graph replay needs no host)
    __shared__ unsigned int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(ticket, 1u);
      is_last = (done == gridDim.x * gridDim.y - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      for (int k = threadIdx.x; k < n_t; k += blockDim.x) t_dev[k] = t_dev[k] - 1;
      if (threadIdx.x == 0) *ticket = 0u;
    }
  }
}

// out[b, ch, t] = the same Philox normal the step kernel injects at counter word `step`
// (x_T of a trajectory is drawn with step = T, one past the largest timestep)
__global__ void __launch_bounds__(256)
philox_normal_kernel(float* __restrict__ out, const unsigned long long* __restrict__ clip_seed,
                     int c, int T, uint32_t step) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const unsigned long long s = clip_seed[b];
  const uint2 key = make_uint2((uint32_t)s, (uint32_t)(s >> 32));
  const int t4 = (T + 3) >> 2;
  const int n = c * T;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ch = i / T, t = i - ch * T;
    out[(size_t)b * n + i] = philox_normal(key, ch, t, t4, step);
  }
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_cfg_step(void* stream, float* x, const float* eps, const float* noise,
                             const uint64_t* clip_seed, const float* sched, int64_t* t_dev,
                             int32_t n_t, uint32_t* ticket, int32_t batch, int32_t c, int32_t t,
                             float guidance, int32_t guided, int32_t advance, void* slab,
                             int32_t copies, int32_t tp, int32_t ld, void* zero,
                             int64_t zero_bytes, float* eps_out) {
  using namespace lm2a;
  LM2A_REQUIRE(x && eps && sched && t_dev, "cfg_step: null pointer");
  LM2A_REQUIRE(batch > 0 && batch <= 65535 && c > 0 && c <= 128 && t > 0,
               "cfg_step: bad geometry (batch=%d c=%d t=%d; c <= 128)", batch, c, t);
  LM2A_REQUIRE(n_t >= batch, "cfg_step: t_dev must hold one timestep per clip (n_t=%d)", n_t);
  LM2A_REQUIRE(!advance || ticket != nullptr, "cfg_step: advance needs a ticket counter");
  LM2A_REQUIRE((reinterpret_cast<uintptr_t>(sched) & 15) == 0,
               "cfg_step: the schedule table must be 16-byte aligned");
  if (slab != nullptr) {
    LM2A_REQUIRE(copies > 0 && tp >= t && ld >= c && ld <= 128 && ld % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(slab) & 15) == 0,
                 "cfg_step: bad slab geometry (copies=%d tp=%d ld=%d; ld %% 8 == 0, 16-byte base)",
                 copies, tp, ld);
  }
  LM2A_REQUIRE(zero_bytes >= 0 && zero_bytes % 16 == 0 &&
                   (zero_bytes == 0 || (zero != nullptr &&
                                        (reinterpret_cast<uintptr_t>(zero) & 15) == 0)),
               "cfg_step: the region to clear must be 16-byte aligned and sized");
  const int span = slab != nullptr ? tp : t;
  dim3 grid((span + 31) / 32, batch);
  LM2A_CUDA_OK(launch_kernel(cfg_step_kernel, grid, dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), x, eps, noise,
                             reinterpret_cast<const unsigned long long*>(clip_seed), sched, t_dev,
                             n_t, ticket, batch, c, t, guidance, guided, advance,
                             reinterpret_cast<__nv_bfloat16*>(slab), copies, tp, ld,
                             reinterpret_cast<uint4*>(zero), (long long)(zero_bytes / 16),
                             eps_out));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_philox_normal(void* stream, float* out, const uint64_t* clip_seed,
                                  int32_t batch, int32_t c, int32_t t, uint32_t step) {
  using namespace lm2a;
  LM2A_REQUIRE(out && clip_seed, "philox_normal: null pointer");
  LM2A_REQUIRE(batch > 0 && batch <= 65535 && c > 0 && t > 0, "philox_normal: bad geometry");
  const int n = c * t;
  int blocks = (n + 255) / 256;
  if (blocks > 148) blocks = 148;
  LM2A_CUDA_OK(launch_kernel(philox_normal_kernel, dim3(blocks, batch), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), out,
                             reinterpret_cast<const unsigned long long*>(clip_seed), c, t, step));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
