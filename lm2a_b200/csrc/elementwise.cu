// HBM-bound helpers of the sampling step: layout ingest, x2 linear upsampling,
// timestep embedding MLP, FiLM tables, and the fused CFG blend + DDPM posterior
// update. All launch on the caller's stream and never synchronise.
#include "../../include/lm2a_b200.h"
#include "common.cuh"

namespace lm2a {
namespace {

// torch.clamp semantics: NaN propagates (fminf / fmaxf alone would turn a NaN eps into a bound,
// and the non-finite early stop of the sampling loop, reference sample.py:216-223, could never
// trigger)
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
  return v != v ? v : fminf(fmaxf(v, lo), hi);
}

// ---------------------------------------------------------------------------
// x fp32 [B, c, T] -> bf16 slab [copies*B, tp, ld]. CFG doubles the batch by
// feeding the same x to the uncond and cond rows (reference sample.py:162), so
// one read of x feeds `copies` rows. First kernel of a UNet step: it also clears the
// step's GroupNorm statistics arena (the conv / bias_add epilogues accumulate into it).
__global__ void __launch_bounds__(256)
ingest_x_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ slab, int batch,
                int copies, int c, int T, int tp, int ld, uint4* __restrict__ zero,
                long long zero_vec) {
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
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int ch = ty; ch < c; ch += 8) {
    const int t = t0 + tx;
    tile[tx][ch] = (t < T) ? __ldg(x + ((size_t)b * c + ch) * T + t) : 0.f;
  }
  __syncthreads();
  for (int tl = ty; tl < 32; tl += 8) {
    const int t = t0 + tl;
    if (t >= tp) break;
    for (int cc = tx; cc < ld; cc += 32) {
      const float v = (cc < c && t < T) ? tile[tl][cc] : 0.f;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      for (int k = 0; k < copies; ++k)
        slab[((size_t)(k * batch + b) * tp + t) * ld + cc] = h;
    }
  }
}

// fp32 [rows, T, c] -> bf16 slab [rows, tp, ld]
__global__ void __launch_bounds__(256)
ingest_seq_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ slab, long long total,
                  int T, int c, int tp, int ld) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cc = (int)(i % ld);
  const long long slot = i / ld;
  const int t = (int)(slot % tp);
  const long long r = slot / tp;
  float v = 0.f;
  if (t < T && cc < c) v = __ldg(x + ((size_t)r * T + t) * c + cc);
  slab[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------
// Condition resampling to the mel length: match_len(arr, T, mode='interp') of the reference
// (datasetcode/dataset.py:49-87 -> interpolate_seq, called from sample.py:124-125), i.e. per
// feature np.interp(linspace(0, L-1, T), arange(L), arr[:, d]) cast to float32. The arithmetic
// follows numpy step by step in fp64 with explicit round-to-nearest ops (no FMA contraction):
//   step = (L-1)/(T-1); x = i*step, x[T-1] = L-1; j = floor(x);
//   y = (f[j+1]-f[j]) * (x - j) + f[j]   (f[L-1] when j == L-1)
// so the fp32 result is bit-identical to the host code. Optionally also writes the bf16 slab
// the CondProjection GEMM consumes (channels >= c and slots >= t_out zeroed).
__global__ void __launch_bounds__(256)
resample_seq_kernel(const float* __restrict__ x, const int* __restrict__ lens,
                    float* __restrict__ out, __nv_bfloat16* __restrict__ slab, long long total,
                    int t_in_max, int c, int t_out, int tp, int ld) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cc = (int)(i % ld);
  const long long slot = i / ld;
  const int t = (int)(slot % tp);
  const int r = (int)(slot / tp);
  float v = 0.f;
  if (t < t_out && cc < c) {
    const int len = lens != nullptr ? lens[r] : t_in_max;
    const float* f = x + (size_t)r * t_in_max * c + cc;
    if (len == t_out) {
      v = __ldg(f + (size_t)t * c);
    } else {
      const double stop = (double)(len - 1);
      const double step = t_out > 1 ? __ddiv_rn(stop, (double)(t_out - 1)) : 0.0;
      const double xn = (t == t_out - 1 && t_out > 1) ? stop : __dmul_rn((double)t, step);
      int j = (int)xn;
      if (j >= len - 1) {
        v = __ldg(f + (size_t)(len - 1) * c);
      } else {
        const double f0 = (double)__ldg(f + (size_t)j * c);
        const double f1 = (double)__ldg(f + (size_t)(j + 1) * c);
        const double slope = __dsub_rn(f1, f0);
        v = __double2float_rn(__dadd_rn(__dmul_rn(slope, __dsub_rn(xn, (double)j)), f0));
      }
    }
    if (out != nullptr) out[((size_t)r * t_out + t) * c + cc] = v;
  }
  if (slab != nullptr) slab[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------
// F.interpolate(scale_factor=2, mode='linear', align_corners=True) on a slab:
// src = i * (T-1)/(2T-1); out = (1-w) x[floor(src)] + w x[min(floor(src)+1, T-1)]
// (reference models/unet1d_ultimate.py:231-236).
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y,
                  int y_ld, long long total_vec, int tp_in, int t_in, int tp_out, int c) {
  pdl_wait();
  pdl_launch_dependents();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int vpr = c >> 3;
  const int cv = (int)(i % vpr);
  const long long slot = i / vpr;
  const int t = (int)(slot % tp_out);
  const long long r = slot / tp_out;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  const int t_out = 2 * t_in;
  if (t < t_out) {
    const float scale = t_out > 1 ? (float)(t_in - 1) / (float)(t_out - 1) : 0.f;
    const float src = scale * (float)t;
    const int i0 = (int)src;
    const int i1 = i0 + (i0 < t_in - 1 ? 1 : 0);
    const float w1 = src - (float)i0;
    const float w0 = 1.0f - w1;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(
        x + ((size_t)r * tp_in + i0) * x_ld + cv * 8));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(
        x + ((size_t)r * tp_in + i1) * x_ld + cv * 8));
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t ow[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = unpack_bf16x2(aw[e]);
      const float2 fb = unpack_bf16x2(bw[e]);
      // explicit rounding order (shared with the conv's fused upsampling: bit-identical)
      ow[e] = pack_bf16x2(__fmaf_rn(w1, fb.x, __fmul_rn(w0, fa.x)),
                          __fmaf_rn(w1, fb.y, __fmul_rn(w0, fa.y)));
    }
    o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  *reinterpret_cast<uint4*>(y + ((size_t)r * tp_out + t) * y_ld + cv * 8) = o;
}

// y[slot, :] = x[slot, :] + bias for valid slots (zero for the pad slots): the
// identity-skip ResBlock of an all-zero-condition row, whose attention output is a
// constant vector (reference unet1d_ultimate.py:152-159 with cross_attention.py:38-67).
// One CTA = one 32-slot segment; thread = one 8-channel vector column, every `tstep`-th slot.
// Optionally accumulates the exact GroupNorm sums of y (see lm2a_conv_desc.stats): per (slot,
// vector) the fp32 sum / sum of squares of its 8 channels in a fixed order, converted to
// 64-bit fixed point (2^24 / 2^20) and added with integer arithmetic, so the statistics do
// not depend on how slots are grouped into CTAs.
__global__ void __launch_bounds__(256)
bias_add_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, __nv_bfloat16* __restrict__ y,
                int y_ld, const float* __restrict__ bias, long long slots, int tp, int t_valid,
                int c, unsigned long long* __restrict__ stats, int stats_pitch, int stats_cg,
                int stats_c0) {
  pdl_wait();
  pdl_launch_dependents();
  const long long m_first = (long long)blockIdx.x * 32;
  const int vpr = c >> 3;
  const int lanes = vpr < 256 ? vpr : 256;
  const int tstep = 256 / lanes;
  const int cvl = threadIdx.x % lanes, ts = threadIdx.x / lanes;
  if (ts >= tstep) return;   // vector-column counts that do not divide the CTA
  // lanes of a warp that hold vectors of one group are combined by a shuffle tree first
  // (only when every warp sees one slot lane and groups are power-of-two runs of lanes)
  int lg = 1;
  if (stats != nullptr && lanes % 32 == 0 && vpr % 32 == 0) {
    const int l = stats_cg >> 3;
    if (l <= 32 && (l & (l - 1)) == 0 && stats_c0 % stats_cg == 0) lg = l;
  }
  for (int cv = cvl; cv < vpr; cv += 256) {
    float bv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bv[e] = __ldg(bias + cv * 8 + e);
    long long s1 = 0, s2 = 0;
    int r_cur = (int)((m_first + ts) / tp);
    auto flush = [&](int rr) {
      if (stats == nullptr) return;
      long long a = s1, b = s2;
      for (int o = 1; o < lg; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((cv & (lg - 1)) == 0 && (a != 0 || b != 0)) {
        unsigned long long* sp =
            stats + ((size_t)rr * stats_pitch + (stats_c0 + cv * 8) / stats_cg) * 2;
        atomicAdd(sp, (unsigned long long)a);
        atomicAdd(sp + 1, (unsigned long long)b);
      }
      s1 = s2 = 0;
    };
    for (int sidx = ts; sidx < 32; sidx += tstep) {
      const long long m = m_first + sidx;
      if (m >= slots) break;
      const int r = (int)(m / tp);
      const int t = (int)(m - (long long)r * tp);
      if (r != r_cur) {
        flush(r_cur);
        r_cur = r;
      }
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (t < t_valid) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(x + (size_t)m * x_ld + cv * 8));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t ow[4];
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack_bf16x2(w[e]);
          const float v0 = f.x + bv[2 * e], v1 = f.y + bv[2 * e + 1];
          a += v0 + v1;
          b = fmaf(v0, v0, fmaf(v1, v1, b));
          ow[e] = pack_bf16x2(v0, v1);
        }
        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        s1 += __float2ll_rn(a * 16777216.0f);
        s2 += __float2ll_rn(b * 1048576.0f);
      }
      *reinterpret_cast<uint4*>(y + (size_t)m * y_ld + cv * 8) = o;
    }
    flush(r_cur);
  }
}

// V cache transpose, once per clip: src bf16 [slots * lk, src_ld] (c channels used) ->
// dst bf16 [slots * c, dst_ld] (lk keys used). The tcgen05 attention kernel wants V^T with
// keys contiguous (K-major B operand of the P V product).
__global__ void __launch_bounds__(256)
transpose_kv_kernel(const __nv_bfloat16* __restrict__ src, int src_ld,
                    __nv_bfloat16* __restrict__ dst, int dst_ld, int lk, int c) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __nv_bfloat16 tile[32][34];
  const int slot = blockIdx.z;
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int key = k0 + i;
    tile[i][tx] = (key < lk) ? src[((size_t)slot * lk + key) * src_ld + c0 + tx]
                             : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int key = k0 + tx;
    if (key < lk) dst[((size_t)slot * c + c0 + i) * dst_ld + key] = tile[tx][i];
  }
}

// ---------------------------------------------------------------------------
// SinusoidalPosEmb -> Linear -> SiLU (reference models/embedding.py:19-43) and the
// SiLU that opens every FiLM net (unet1d_ultimate.py:50-53). One CTA per row;
// each warp produces outputs with a shuffle reduction over the input dim.
__global__ void __launch_bounds__(256)
time_mlp_kernel(const int64_t* __restrict__ t, const float* __restrict__ w,
                const float* __restrict__ b, float* __restrict__ out, int dim, int fold_silu) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float emb[];
  const int r = blockIdx.x;
  const int half = dim / 2;
  const float tv = (float)t[r];
  const float step = (float)(-(log(10000.0) / (double)(half - 1)));
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    const int k = i < half ? i : i - half;
    const float f = expf((float)k * step);
    const float a = tv * f;
    emb[i] = i < half ? sinf(a) : cosf(a);
  }
  __syncthreads();
  // blockIdx.y selects a group of 8 outputs (one per warp): the weight rows of different
  // outputs stream in parallel instead of serially through one CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = blockIdx.y * 8 + warp; j < dim; j += gridDim.y * 8) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc = fmaf(emb[k], __ldg(w + (size_t)j * dim + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float v = silu_accurate(acc + __ldg(b + j));
      out[(size_t)r * dim + j] = fold_silu ? silu_accurate(v) : v;
    }
  }
}

// film[r, j] = <silu_temb[r, :], w[j, :]> + b[j]; one warp per output column j,
// the weight row lives in registers and is reused across all rows.
template <int KPL>  // dim / 32
__global__ void __launch_bounds__(256)
film_kernel(const float* __restrict__ s, const float* __restrict__ w,
            const float* __restrict__ b, float* __restrict__ film, int rows, int cols) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= cols) return;
  constexpr int dim = KPL * 32;
  float wr[KPL];
#pragma unroll
  for (int i = 0; i < KPL; ++i) wr[i] = __ldg(w + (size_t)j * dim + lane + 32 * i);
  const float bj = __ldg(b + j);
  for (int r = 0; r < rows; ++r) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) acc = fmaf(wr[i], __ldg(s + (size_t)r * dim + lane + 32 * i), acc);
    acc = warp_sum(acc);
    if (lane == 0) film[(size_t)r * cols + j] = acc + bj;
  }
}

// ---------------------------------------------------------------------------
// Classifier-free-guidance blend with both clamps (reference sample.py:167-174)
// and the DDPM posterior update (sample.py:186-210 == models/diffusion.py:71-102)
// in one pass: 5 fp32 streams per element (x, eps_u, eps_c, noise in; x out).
// Arithmetic keeps the reference's operation order with explicit round-to-
// nearest mul/add/sub (no FMA contraction), so given the same eps and noise the
// update is bit-identical to the PyTorch elementwise sequence.
__global__ void __launch_bounds__(256)
cfg_posterior_kernel(float* __restrict__ x, const float* __restrict__ eps,
                     const float* __restrict__ noise, const float* __restrict__ sched,
                     int64_t* __restrict__ t_dev, int n_t, unsigned int* __restrict__ ticket,
                     long long total_vec, long long clip_vec, int batch, float gw, int guided,
                     int advance, float* __restrict__ eps_out) {
  pdl_wait();
  pdl_launch_dependents();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long uncond_to_cond = (long long)batch * clip_vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += stride) {
    // per-clip timestep (GaussianDiffusion.p_sample accepts a (B,) tensor of mixed t)
    const long long t_now = t_dev[i / clip_vec];
    const float4 co = __ldg(reinterpret_cast<const float4*>(sched) + t_now);
    const float coef1 = co.x, coef2 = co.y, sigma = co.z;
    const bool add_noise = t_now > 0 && noise != nullptr;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float4 e;
    if (guided) {
      const float4 eu = __ldg(reinterpret_cast<const float4*>(eps) + i);
      const float4 ec = __ldg(reinterpret_cast<const float4*>(eps) + i + uncond_to_cond);
      const float u[4] = {eu.x, eu.y, eu.z, eu.w};
      const float c[4] = {ec.x, ec.y, ec.z, ec.w};
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float d = __fsub_rn(c[k], u[k]);
        d = clamp_nan(d, -5.0f, 5.0f);
        float g = __fadd_rn(u[k], __fmul_rn(gw, d));
        o[k] = clamp_nan(g, -10.0f, 10.0f);
      }
      e = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      e = __ldg(reinterpret_cast<const float4*>(eps) + i);
    }
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add_noise) nz = __ldg(reinterpret_cast<const float4*>(noise) + i);
    float4 o;
    o.x = __fadd_rn(__fmul_rn(coef1, __fsub_rn(xv.x, __fmul_rn(coef2, e.x))), __fmul_rn(sigma, nz.x));
    o.y = __fadd_rn(__fmul_rn(coef1, __fsub_rn(xv.y, __fmul_rn(coef2, e.y))), __fmul_rn(sigma, nz.y));
    o.z = __fadd_rn(__fmul_rn(coef1, __fsub_rn(xv.z, __fmul_rn(coef2, e.z))), __fmul_rn(sigma, nz.z));
    o.w = __fadd_rn(__fmul_rn(coef1, __fsub_rn(xv.w, __fmul_rn(coef2, e.w))), __fmul_rn(sigma, nz.w));
    reinterpret_cast<float4*>(x)[i] = o;
    if (eps_out != nullptr) reinterpret_cast<float4*>(eps_out)[i] = e;
  }
  if (advance) {
    // the last CTA to finish steps the device-side timestep (graph replay needs no host)
    __shared__ unsigned int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(ticket, 1u);
      is_last = (done == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      for (int k = threadIdx.x; k < n_t; k += blockDim.x) t_dev[k] = t_dev[k] - 1;
      if (threadIdx.x == 0) *ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// CFG blend + DDIM update (reference models/diffusion.py:124-165, `ddim_sample`) over a
// strided timestep sequence. Row k of `table` (fp32 x 8) holds, for step k of the sequence,
//   {sqrt(1 - abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1 - abar_prev - sigma^2), sigma,
//    noise gate (1 if t_prev > 0 else 0), 0, 0}
// computed by the caller with the reference's torch expressions, so with the same eps and
// noise the result is bit-identical to the torch op sequence (explicit rn ops, no FMA):
//   x0 = clamp((x - eps * c0) / c1, -2, 2);  x = (c2 * x0 + c3 * eps) + sigma * z
// The last CTA advances the device-side step index and loads the next timestep of the
// sequence into t_dev (graph replay needs no host).
__global__ void __launch_bounds__(256)
cfg_ddim_kernel(float* __restrict__ x, const float* __restrict__ eps,
                const float* __restrict__ noise, const float* __restrict__ table,
                const int64_t* __restrict__ t_seq, int* __restrict__ step_idx,
                int64_t* __restrict__ t_dev, int n_t, unsigned int* __restrict__ ticket,
                long long total_vec, long long clip_vec, int batch, float gw, int guided,
                int advance, float* __restrict__ x0_out) {
  pdl_wait();
  pdl_launch_dependents();
  const int k = *step_idx;
  const float4 ca = __ldg(reinterpret_cast<const float4*>(table) + 2 * k);
  const float4 cb = __ldg(reinterpret_cast<const float4*>(table) + 2 * k + 1);
  const float c0 = ca.x, c1 = ca.y, c2 = ca.z, c3 = ca.w, sigma = cb.x;
  const bool add_noise = cb.y != 0.f && noise != nullptr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long uncond_to_cond = (long long)batch * clip_vec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += stride) {
    const float4 xv4 = reinterpret_cast<const float4*>(x)[i];
    float e[4];
    if (guided) {
      const float4 eu = __ldg(reinterpret_cast<const float4*>(eps) + i);
      const float4 ec = __ldg(reinterpret_cast<const float4*>(eps) + i + uncond_to_cond);
      const float u[4] = {eu.x, eu.y, eu.z, eu.w};
      const float c[4] = {ec.x, ec.y, ec.z, ec.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float d = __fsub_rn(c[q], u[q]);
        d = clamp_nan(d, -5.0f, 5.0f);
        const float g = __fadd_rn(u[q], __fmul_rn(gw, d));
        e[q] = clamp_nan(g, -10.0f, 10.0f);
      }
    } else {
      const float4 ev = __ldg(reinterpret_cast<const float4*>(eps) + i);
      e[0] = ev.x; e[1] = ev.y; e[2] = ev.z; e[3] = ev.w;
    }
    float4 nz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (add_noise) nz = __ldg(reinterpret_cast<const float4*>(noise) + i);
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
    const float zv[4] = {nz.x, nz.y, nz.z, nz.w};
    float o[4], x0[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float p0 = __fdiv_rn(__fsub_rn(xv[q], __fmul_rn(e[q], c0)), c1);
      p0 = clamp_nan(p0, -2.0f, 2.0f);
      x0[q] = p0;
      o[q] = __fadd_rn(__fadd_rn(__fmul_rn(c2, p0), __fmul_rn(c3, e[q])),
                       __fmul_rn(sigma, zv[q]));
    }
    reinterpret_cast<float4*>(x)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (x0_out != nullptr)
      reinterpret_cast<float4*>(x0_out)[i] = make_float4(x0[0], x0[1], x0[2], x0[3]);
  }
  if (advance) {
    __shared__ unsigned int is_last;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int done = atomicAdd(ticket, 1u);
      is_last = (done == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (is_last) {
      const int64_t t_next = t_seq[k + 1];
      for (int q = threadIdx.x; q < n_t; q += blockDim.x) t_dev[q] = t_next;
      if (threadIdx.x == 0) {
        *step_idx = k + 1;
        *ticket = 0u;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Mel-spectrogram evaluation metrics of reference val.py:25-113 (`compute_metrics`), one CTA
// per clip, fp64 accumulation: MSE, SSIM (skimage.metrics.structural_similarity as val.py
// calls it: per mel band 1-D, Gaussian window sigma 1.5 / radius 5, data_range 1, population
// covariance, border of 5 frames cropped, mean over bands; inputs min-max normalised with the
// REAL mel's range and clipped to [0, 1]), mean frame cosine, |mean error|, |std error|, SNR.
// gen is de-normalised on the fly (gen * std + mean, sample.py:230). out: fp64 [B, 8]
// {mse, ssim, avg_cos_sim, mean_error, std_error, snr, real_var, 0}.
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];  // fixed order
  return t;
}
__device__ __forceinline__ float block_minmax(float v, bool is_max, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : fminf(v, u);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) t = is_max ? fmaxf(t, red[w]) : fminf(t, red[w]);
  return t;
}

__global__ void __launch_bounds__(1024)
mel_metrics_kernel(const float* __restrict__ gen, const float* __restrict__ real,
                   double* __restrict__ out, int n_mels, int T, float g_scale, float g_shift) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ double red[32];
  __shared__ float redf[32];
  const int b = blockIdx.x;
  const float* g = gen + (size_t)b * n_mels * T;
  const float* r = real + (size_t)b * n_mels * T;
  const int n = n_mels * T;
  double se = 0.0, sr = 0.0, srr = 0.0, sg = 0.0, sgg = 0.0;
  float rmin = INFINITY, rmax = -INFINITY, gmin = INFINITY, gmax = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float rv = r[i];
    const float gv = __fadd_rn(__fmul_rn(g[i], g_scale), g_shift);
    const double d = (double)rv - (double)gv;
    se += d * d;
    sr += rv; srr += (double)rv * rv;
    sg += gv; sgg += (double)gv * gv;
    rmin = fminf(rmin, rv); rmax = fmaxf(rmax, rv);
    gmin = fminf(gmin, gv); gmax = fmaxf(gmax, gv);
  }
  se = block_sum(se, red);
  sr = block_sum(sr, red); srr = block_sum(srr, red);
  sg = block_sum(sg, red); sgg = block_sum(sgg, red);
  rmin = block_minmax(rmin, false, redf); rmax = block_minmax(rmax, true, redf);
  gmin = block_minmax(gmin, false, redf); gmax = block_minmax(gmax, true, redf);
  // frame cosine (sklearn cosine_similarity of the two 80-vectors of a frame), averaged
  double cs = 0.0;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double ab = 0.0, aa = 0.0, bb = 0.0;
    for (int c = 0; c < n_mels; ++c) {
      const double rv = r[(size_t)c * T + t];
      const double gv = __fadd_rn(__fmul_rn(g[(size_t)c * T + t], g_scale), g_shift);
      ab += rv * gv; aa += rv * rv; bb += gv * gv;
    }
    const double na = sqrt(aa), nb = sqrt(bb);
    cs += ab / ((na == 0.0 ? 1.0 : na) * (nb == 0.0 ? 1.0 : nb));
  }
  cs = block_sum(cs, red);
  // SSIM
  float lo = rmin, hi = rmax;
  if (hi - lo < 1e-6f) {
    lo = fminf(rmin, gmin);
    hi = fmaxf(rmax, gmax);
  }
  const double inv_range = 1.0 / ((double)hi - (double)lo + 1e-8);
  double w[11];
  {
    double ws = 0.0;
    for (int k = -5; k <= 5; ++k) {
      w[k + 5] = exp(-0.5 * (double)(k * k) / (1.5 * 1.5));
      ws += w[k + 5];
    }
    for (int k = 0; k < 11; ++k) w[k] /= ws;
  }
  const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
  const int span = T - 10;  // frames 5 .. T-6 survive the crop
  double ss = 0.0;
  if (span > 0) {
    for (int i = threadIdx.x; i < n_mels * span; i += blockDim.x) {
      const int c = i / span, t = 5 + i % span;
      double ux = 0, uy = 0, uxx = 0, uyy = 0, uxy = 0;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const size_t idx = (size_t)c * T + t - 5 + k;
        double xv = ((double)r[idx] - (double)lo) * inv_range;
        double yv = ((double)__fadd_rn(__fmul_rn(g[idx], g_scale), g_shift) - (double)lo) * inv_range;
        xv = fmin(fmax(xv, 0.0), 1.0);
        yv = fmin(fmax(yv, 0.0), 1.0);
        ux += w[k] * xv; uy += w[k] * yv;
        uxx += w[k] * xv * xv; uyy += w[k] * yv * yv; uxy += w[k] * xv * yv;
      }
      const double vx = uxx - ux * ux, vy = uyy - uy * uy, vxy = uxy - ux * uy;
      ss += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2));
    }
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) {
    const double inv = 1.0 / (double)n;
    const double mse = se * inv;
    const double mr = sr * inv, mg = sg * inv;
    double vr = srr * inv - mr * mr, vg = sgg * inv - mg * mg;
    vr = vr > 0 ? vr : 0; vg = vg > 0 ? vg : 0;
    double ssim = span > 0 ? ss / ((double)n_mels * span) : 0.0;
    ssim = fmin(fmax(ssim, 0.0), 1.0);
    double* o = out + (size_t)b * 8;
    o[0] = mse;
    o[1] = ssim;
    o[2] = cs / (double)T;
    o[3] = fabs(mr - mg);
    o[4] = fabs(sqrt(vr) - sqrt(vg));
    o[5] = vr < 1e-8 ? 0.0 : 10.0 * log10(vr / (mse + 1e-8));
    o[6] = vr;
    o[7] = 0.0;
  }
}

// ---------------------------------------------------------------------------
// Fused Adan step for every parameter tensor of the model in ONE launch (reference
// models/adan.py:34-114, restart_cond = None; train.py:99-100 builds it over all UNet +
// CondProjection parameters) with the EMA shadow update of train.py:177-180 folded in. The
// reference runs ~20 elementwise launches per tensor (306 + 4 tensors); here a chunk table
// maps CTAs onto (tensor, 64 K-element chunk) pairs. 13 fp32 streams per element with EMA
// (p, g, prev_g, m, v, n, ema read; p, prev_g, m, v, n, ema written) = 52 B, 44 B without.
// Operation order and roundings follow the reference's torch calls (mul_ then add_(alpha) =
// one FMA, scalar operands pre-rounded to fp32 by the caller, reciprocal * lr for `lr / t`,
// addcmul_ = fma(value * t1, t2, self), IEEE sqrt, true division by the weight-decay
// denominator). Moments are bit-identical to torch on CPU and CUDA; p agrees within a few ulp
// (torch's CPU sqrt is not correctly rounded, torch's CUDA div_ by a scalar multiplies by the
// reciprocal: the two torch back ends do not agree bit for bit with each other either).
struct AdanScalars {
  float om_b1, b1, om_b2, b2, om_b3, b3;   // (1 - beta_i), beta_i
  float cm, cv, cn;                         // bias corrections 1 / (1 - (1 - beta_i)^step)
  float lr, eps, denom;                     // denom = 1 + weight_decay * lr
  float ema_decay, om_ema_decay;
  int first_step;                           // state["step"] == 0: moments are not updated
};

__device__ __forceinline__ void adan_elem(float& p, float g, float& prev, float& m, float& v,
                                          float& n, float* ema, const AdanScalars& s) {
  if (!s.first_step) {
    m = fmaf(s.b1, g, __fmul_rn(m, s.om_b1));
    const float diff = __fsub_rn(g, prev);
    v = fmaf(s.b2, diff, __fmul_rn(v, s.om_b2));
    const float u = __fadd_rn(g, __fmul_rn(diff, s.om_b2));
    n = fmaf(s.b3, __fmul_rn(u, u), __fmul_rn(n, s.om_b3));
  }
  const float t = __fadd_rn(__fsqrt_rn(__fmul_rn(n, s.cn)), s.eps);
  const float wss = __fmul_rn(__frcp_rn(t), s.lr);
  const float upd = __fadd_rn(__fmul_rn(m, s.cm), __fmul_rn(__fmul_rn(v, s.om_b2), s.cv));
  p = __fdiv_rn(fmaf(__fmul_rn(-1.0f, wss), upd, p), s.denom);
  prev = g;
  if (ema != nullptr) *ema = __fadd_rn(__fmul_rn(*ema, s.ema_decay), __fmul_rn(p, s.om_ema_decay));
}

__global__ void __launch_bounds__(256)
adan_step_kernel(const lm2a_adan_tensor* __restrict__ tensors,
                 const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_index,
                 int chunk_elems, AdanScalars s) {
  pdl_wait();
  pdl_launch_dependents();
  const lm2a_adan_tensor t = tensors[chunk_tensor[blockIdx.x]];
  const long long begin = (long long)chunk_index[blockIdx.x] * chunk_elems;
  const long long end = begin + chunk_elems < t.numel ? begin + chunk_elems : t.numel;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                     reinterpret_cast<uintptr_t>(t.prev_g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v) | reinterpret_cast<uintptr_t>(t.n) |
                     reinterpret_cast<uintptr_t>(t.ema)) & 15) == 0;
  long long i = begin + (long long)threadIdx.x * 4;
  if (vec) {
    for (; i + 3 < end; i += 256 * 4) {
      float4 p = *reinterpret_cast<float4*>(t.p + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(t.g + i));
      float4 pr = *reinterpret_cast<float4*>(t.prev_g + i);
      float4 m = *reinterpret_cast<float4*>(t.m + i);
      float4 v = *reinterpret_cast<float4*>(t.v + i);
      float4 n = *reinterpret_cast<float4*>(t.n + i);
      float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t.ema != nullptr) e = *reinterpret_cast<float4*>(t.ema + i);
      adan_elem(p.x, g.x, pr.x, m.x, v.x, n.x, t.ema ? &e.x : nullptr, s);
      adan_elem(p.y, g.y, pr.y, m.y, v.y, n.y, t.ema ? &e.y : nullptr, s);
      adan_elem(p.z, g.z, pr.z, m.z, v.z, n.z, t.ema ? &e.z : nullptr, s);
      adan_elem(p.w, g.w, pr.w, m.w, v.w, n.w, t.ema ? &e.w : nullptr, s);
      *reinterpret_cast<float4*>(t.p + i) = p;
      *reinterpret_cast<float4*>(t.prev_g + i) = pr;
      *reinterpret_cast<float4*>(t.m + i) = m;
      *reinterpret_cast<float4*>(t.v + i) = v;
      *reinterpret_cast<float4*>(t.n + i) = n;
      if (t.ema != nullptr) *reinterpret_cast<float4*>(t.ema + i) = e;
    }
    // ragged tail of the tensor (< 4 elements): the thread whose vector straddles `end`
    for (long long k = i; k < end && k < i + 4; ++k)
      adan_elem(t.p[k], t.g[k], t.prev_g[k], t.m[k], t.v[k], t.n[k],
                t.ema ? t.ema + k : nullptr, s);
  } else {
    for (long long k = begin + threadIdx.x; k < end; k += 256)
      adan_elem(t.p[k], t.g[k], t.prev_g[k], t.m[k], t.v[k], t.n[k],
                t.ema ? t.ema + k : nullptr, s);
  }
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_ingest_x(void* stream, const float* x, void* slab, int32_t batch,
                             int32_t copies, int32_t c, int32_t t, int32_t tp, int32_t ld,
                             void* zero, int64_t zero_bytes) {
  using namespace lm2a;
  LM2A_REQUIRE(x && slab, "ingest_x: null pointer");
  LM2A_REQUIRE(batch > 0 && copies > 0 && c > 0 && c <= 128 && ld <= 128 && ld >= c && t > 0 &&
                   tp >= t,
               "ingest_x: bad geometry (c=%d ld=%d t=%d tp=%d; c, ld <= 128)", c, ld, t, tp);
  LM2A_REQUIRE(zero_bytes >= 0 && zero_bytes % 16 == 0 &&
                   (zero_bytes == 0 || (zero != nullptr &&
                                        (reinterpret_cast<uintptr_t>(zero) & 15) == 0)),
               "ingest_x: the region to clear must be 16-byte aligned and sized");
  dim3 grid((tp + 31) / 32, batch);
  LM2A_CUDA_OK(launch_kernel(ingest_x_kernel, dim3(grid), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), x,
                             reinterpret_cast<__nv_bfloat16*>(slab), batch, copies, c, t, tp, ld,
                             reinterpret_cast<uint4*>(zero), (long long)(zero_bytes / 16)));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_ingest_seq(void* stream, const float* x, void* slab, int32_t rows,
                               int32_t t, int32_t c, int32_t tp, int32_t ld) {
  using namespace lm2a;
  LM2A_REQUIRE(x && slab, "ingest_seq: null pointer");
  LM2A_REQUIRE(rows > 0 && t > 0 && tp >= t && c > 0 && ld >= c, "ingest_seq: bad geometry");
  const long long total = (long long)rows * tp * ld;
  const int blocks = (int)((total + 255) / 256);
  LM2A_CUDA_OK(launch_kernel(ingest_seq_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<__nv_bfloat16*>(slab), total, t, c, tp, ld));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_resample_seq(void* stream, const float* x, const int32_t* lens,
                                 float* out_f32, void* out_slab, int32_t rows, int32_t t_in_max,
                                 int32_t c, int32_t t_out, int32_t tp, int32_t ld) {
  using namespace lm2a;
  LM2A_REQUIRE(x && (out_f32 || out_slab), "resample_seq: null pointer");
  LM2A_REQUIRE(rows > 0 && t_in_max > 0 && c > 0 && t_out > 0 && tp >= t_out && ld >= c,
               "resample_seq: bad geometry (rows=%d t_in=%d c=%d t_out=%d tp=%d ld=%d)", rows,
               t_in_max, c, t_out, tp, ld);
  const long long total = (long long)rows * tp * ld;
  const int blocks = (int)((total + 255) / 256);
  LM2A_CUDA_OK(launch_kernel(resample_seq_kernel, dim3(blocks), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), x, lens, out_f32,
                             reinterpret_cast<__nv_bfloat16*>(out_slab), total, t_in_max, c,
                             t_out, tp, ld));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_transpose_kv_bf16(void* stream, const void* src, int32_t src_ld, void* dst,
                                      int32_t dst_ld, int32_t slots, int32_t lk, int32_t c) {
  using namespace lm2a;
  LM2A_REQUIRE(src && dst, "transpose_kv: null pointer");
  LM2A_REQUIRE(slots > 0 && slots <= 65535 && lk > 0 && c > 0 && c % 32 == 0 && src_ld >= c &&
                   dst_ld >= lk,
               "transpose_kv: bad geometry (slots=%d lk=%d c=%d src_ld=%d dst_ld=%d)", slots, lk,
               c, src_ld, dst_ld);
  dim3 grid((lk + 31) / 32, c / 32, slots);
  LM2A_CUDA_OK(launch_kernel(transpose_kv_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(src), src_ld, reinterpret_cast<__nv_bfloat16*>(dst),
      dst_ld, lk, c));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_upsample2x_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                                    int32_t y_ld, int32_t rows, int32_t tp_in, int32_t t_in,
                                    int32_t tp_out, int32_t c) {
  using namespace lm2a;
  LM2A_REQUIRE(x && y, "upsample2x: null pointer");
  LM2A_REQUIRE(rows > 0 && t_in > 0 && tp_in >= t_in && tp_out >= 2 * t_in && c % 8 == 0 &&
                   x_ld % 8 == 0 && y_ld % 8 == 0 && x_ld >= c && y_ld >= c,
               "upsample2x: bad geometry");
  const long long total_vec = (long long)rows * tp_out * (c / 8);
  const int blocks = (int)((total_vec + 255) / 256);
  LM2A_CUDA_OK(launch_kernel(upsample2x_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), x_ld, reinterpret_cast<__nv_bfloat16*>(y), y_ld,
      total_vec, tp_in, t_in, tp_out, c));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_bias_add_bf16(void* stream, const void* x, int32_t x_ld, void* y,
                                  int32_t y_ld, const float* bias, int64_t slots, int32_t tp,
                                  int32_t t_valid, int32_t c, void* stats, int32_t stats_pitch,
                                  int32_t stats_cg, int32_t stats_c0) {
  using namespace lm2a;
  LM2A_REQUIRE(x && y && bias, "bias_add: null pointer");
  LM2A_REQUIRE(slots > 0 && tp > 0 && t_valid > 0 && t_valid <= tp && slots % tp == 0 &&
                   c > 0 && c % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && x_ld >= c && y_ld >= c,
               "bias_add: bad geometry");
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                 reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
               "bias_add: tensors must be 16-byte aligned");
  if (stats != nullptr) {
    LM2A_REQUIRE(stats_cg > 0 && stats_cg % 8 == 0 && stats_pitch > 0 && stats_c0 >= 0 &&
                     stats_c0 % 8 == 0 &&
                     (stats_c0 + c + stats_cg - 1) / stats_cg <= stats_pitch &&
                     (reinterpret_cast<uintptr_t>(stats) & 15) == 0,
                 "bias_add: bad stats layout (channels per group=%d, groups per row=%d, first "
                 "channel=%d)", stats_cg, stats_pitch, stats_c0);
  }
  const long long blocks = (slots + 31) / 32;
  LM2A_REQUIRE(slots < (1ll << 31) - 64, "bias_add: too many slots");
  LM2A_CUDA_OK(launch_kernel(bias_add_kernel, dim3((unsigned)blocks), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream),
                             reinterpret_cast<const __nv_bfloat16*>(x), x_ld,
                             reinterpret_cast<__nv_bfloat16*>(y), y_ld, bias, (long long)slots, tp,
                             t_valid, c, reinterpret_cast<unsigned long long*>(stats), stats_pitch,
                             stats_cg, stats_c0));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_time_embed(void* stream, const int64_t* t, const float* w, const float* b,
                               float* out, int32_t rows, int32_t dim, int32_t fold_silu) {
  using namespace lm2a;
  LM2A_REQUIRE(t && w && b && out, "time_mlp: null pointer");
  LM2A_REQUIRE(rows > 0 && dim >= 4 && dim % 2 == 0 && dim <= 4096, "time_mlp: bad dim %d", dim);
  LM2A_CUDA_OK(launch_kernel(time_mlp_kernel, dim3(dim3(rows, (dim + 7) / 8)), dim3(256),
                             dim * sizeof(float), reinterpret_cast<cudaStream_t>(stream), t, w, b,
                             out, dim, fold_silu != 0 ? 1 : 0));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_time_mlp(void* stream, const int64_t* t, const float* w, const float* b,
                             float* silu_temb, int32_t rows, int32_t dim) {
  return lm2a_time_embed(stream, t, w, b, silu_temb, rows, dim, 1);
}

extern "C" int lm2a_film(void* stream, const float* silu_temb, const float* w, const float* b,
                         float* film, int32_t rows, int32_t dim, int32_t cols) {
  using namespace lm2a;
  LM2A_REQUIRE(silu_temb && w && b && film, "film: null pointer");
  LM2A_REQUIRE(rows > 0 && cols > 0, "film: bad geometry");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (cols + 7) / 8;
  switch (dim) {
    case 128: LM2A_CUDA_OK(launch_kernel(film_kernel<4>, dim3(blocks), dim3(256), 0, st, silu_temb, w, b, film, rows, cols)); break;
    case 256: LM2A_CUDA_OK(launch_kernel(film_kernel<8>, dim3(blocks), dim3(256), 0, st, silu_temb, w, b, film, rows, cols)); break;
    case 512: LM2A_CUDA_OK(launch_kernel(film_kernel<16>, dim3(blocks), dim3(256), 0, st, silu_temb, w, b, film, rows, cols)); break;
    default:
      LM2A_REQUIRE(false, "film: time_emb_dim %d unsupported (128, 256 or 512)", dim);
  }
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_cfg_posterior(void* stream, float* x, const float* eps, const float* noise,
                                  const float* sched, int64_t* t_dev, int32_t n_t,
                                  uint32_t* ticket, int32_t batch, int64_t elems_per_clip,
                                  float guidance, int32_t guided, int32_t advance,
                                  float* eps_out) {
  using namespace lm2a;
  LM2A_REQUIRE(x && eps && sched && t_dev, "cfg_posterior: null pointer");
  LM2A_REQUIRE(batch > 0 && elems_per_clip > 0 && elems_per_clip % 4 == 0,
               "cfg_posterior: elems_per_clip must be a positive multiple of 4");
  LM2A_REQUIRE(!advance || (ticket != nullptr && n_t > 0),
               "cfg_posterior: advance needs a ticket counter and n_t > 0");
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps) |
                 reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(sched) |
                 reinterpret_cast<uintptr_t>(eps_out)) & 15) == 0,
               "cfg_posterior: tensors must be 16-byte aligned");
  const long long clip_vec = elems_per_clip / 4;
  const long long total_vec = clip_vec * batch;
  long long blocks = (total_vec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  LM2A_CUDA_OK(launch_kernel(cfg_posterior_kernel, dim3((int)blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, eps, noise, sched, t_dev, n_t, ticket, total_vec, clip_vec, batch, guidance, guided,
      advance, eps_out));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_cfg_ddim(void* stream, float* x, const float* eps, const float* noise,
                             const float* table, const int64_t* t_seq, int32_t* step_idx,
                             int64_t* t_dev, int32_t n_t, uint32_t* ticket, int32_t batch,
                             int64_t elems_per_clip, float guidance, int32_t guided,
                             int32_t advance, float* x0_out) {
  using namespace lm2a;
  LM2A_REQUIRE(x && eps && table && step_idx, "cfg_ddim: null pointer");
  LM2A_REQUIRE(batch > 0 && elems_per_clip > 0 && elems_per_clip % 4 == 0,
               "cfg_ddim: elems_per_clip must be a positive multiple of 4");
  LM2A_REQUIRE(!advance || (ticket != nullptr && t_seq != nullptr && t_dev != nullptr && n_t > 0),
               "cfg_ddim: advance needs ticket, t_seq, t_dev and n_t > 0");
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps) |
                 reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(table) |
                 reinterpret_cast<uintptr_t>(x0_out)) & 15) == 0,
               "cfg_ddim: tensors must be 16-byte aligned");
  const long long clip_vec = elems_per_clip / 4;
  const long long total_vec = clip_vec * batch;
  long long blocks = (total_vec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  LM2A_CUDA_OK(launch_kernel(cfg_ddim_kernel, dim3((int)blocks), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), x, eps, noise, table, t_seq,
                             step_idx, t_dev, n_t, ticket, total_vec, clip_vec, batch, guidance,
                             guided, advance, x0_out));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_mel_metrics(void* stream, const float* gen, const float* real, double* out,
                                int32_t batch, int32_t n_mels, int32_t t, float gen_scale,
                                float gen_shift) {
  using namespace lm2a;
  LM2A_REQUIRE(gen && real && out, "mel_metrics: null pointer");
  LM2A_REQUIRE(batch > 0 && n_mels > 0 && t > 0, "mel_metrics: bad geometry");
  LM2A_REQUIRE(t >= 11, "mel_metrics: SSIM window (11 frames) exceeds the clip length %d", t);
  LM2A_CUDA_OK(launch_kernel(mel_metrics_kernel, dim3(batch), dim3(1024), 0,
                             reinterpret_cast<cudaStream_t>(stream), gen, real, out, n_mels, t,
                             gen_scale, gen_shift));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_adan_step(void* stream, const lm2a_adan_tensor* tensors,
                              const int32_t* chunk_tensor, const int32_t* chunk_index,
                              int32_t n_chunks, int32_t chunk_elems, int32_t first_step,
                              const float* scalars) {
  using namespace lm2a;
  LM2A_REQUIRE(tensors && chunk_tensor && chunk_index && scalars, "adan_step: null pointer");
  LM2A_REQUIRE(n_chunks > 0 && chunk_elems > 0 && chunk_elems % 1024 == 0,
               "adan_step: chunk_elems must be a positive multiple of 1024");
  AdanScalars s;
  s.om_b1 = scalars[0]; s.b1 = scalars[1]; s.om_b2 = scalars[2]; s.b2 = scalars[3];
  s.om_b3 = scalars[4]; s.b3 = scalars[5]; s.cm = scalars[6]; s.cv = scalars[7];
  s.cn = scalars[8]; s.lr = scalars[9]; s.eps = scalars[10]; s.denom = scalars[11];
  s.ema_decay = scalars[12]; s.om_ema_decay = scalars[13];
  s.first_step = first_step != 0 ? 1 : 0;
  LM2A_CUDA_OK(launch_kernel(adan_step_kernel, dim3(n_chunks), dim3(256), 0,
                             reinterpret_cast<cudaStream_t>(stream), tensors, chunk_tensor,
                             chunk_index, chunk_elems, s));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
