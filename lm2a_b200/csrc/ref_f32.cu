// fp32 validation path: the same launch plan (descriptors, slab geometry, folds, exact GroupNorm
// sums) executed with plain fp32 CUDA-core kernels on fp32 slabs and fp32 weights. It exists for
// one purpose - BASELINE.json's tolerance (i): the single-step eps prediction must agree with the
// reference's fp32 PyTorch path within 1e-4 relative - and is selected with
// UNet1D_ultimate(..., precision="fp32"). The production path is the bf16 tcgen05 one; these
// kernels are simple tiled SIMT code (no tensor cores: kind::tf32 has a 10-bit mantissa), a few
// TFLOP/s, not tuned. Same reference lines as their bf16 counterparts:
//   lm2a_conv1d_f32        unet1d_ultimate.py:87-88,115,136-159,216-239,255-261,295,362-364;
//                          cross_attention.py:19-36 (projections)
//   lm2a_cross_attn_f32    cross_attention.py:50-61 (nn.MultiheadAttention core)
//   lm2a_bias_add_f32, lm2a_ingest_x_f32, lm2a_ingest_seq_f32: slab plumbing of engine.py
#include "../../include/lm2a_b200.h"
#include <math.h>

#include "common.cuh"

namespace lm2a {
namespace {

constexpr int kBM = 64, kBN = 64, kBKc = 16;
constexpr int kMaxGnEnt = 1024;
constexpr float kScale1 = 16777216.0f;   // 2^24
constexpr float kScale2 = 1048576.0f;    // 2^20

struct F32Args {
  const float* x[2];
  long long rows[2];
  int ld[2], cin[2], taps[2];
  const float* w;
  int k_total;
  long long m;
  int tp, t_valid, n_valid;
  const float* bias;
  const float* film;
  int film_ld, film_shift_off;
  const float* residual;
  int res_ld;
  float* out;
  int out_ld, out_mode;
  unsigned long long* stats;
  int stats_pitch, stats_cg, stats_c0;
  const long long* gn_stats;
  const float* gn_gamma;
  const float* gn_beta;
  int gn_pitch, gn_groups, gn_cg;
  float gn_eps;
  int up_tp_in, up_t_in;
};

__device__ __forceinline__ float silu_exact(float v) { return v / (1.0f + expf(-v)); }

__global__ void __launch_bounds__(256)
conv_f32_kernel(const F32Args p) {
  __shared__ float As[kBKc][kBM + 4];
  __shared__ float Bs[kBKc][kBN + 4];
  __shared__ float2 gn_tab[kMaxGnEnt];   // (rstd, -mean * rstd) per (clip-row, group) of the tile
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * kBM;
  const int n0 = blockIdx.y * kBN;
  const int ty = tid >> 4, tx = tid & 15;
  const bool up2x = p.up_t_in > 0;
  const bool gn = p.gn_stats != nullptr;
  // clip-rows the (shifted) operand rows of this tile can touch
  const long long lo = m0 - 1 > 0 ? m0 - 1 : 0;
  const long long hi = m0 + kBM < p.m ? m0 + kBM : p.m - 1;
  const int r_first = (int)(lo / p.tp);
  if (gn) {
    const int nrow = (int)(hi / p.tp) - r_first + 1;
    const double inv_n = 1.0 / ((double)p.gn_cg * (double)p.t_valid);
    for (int e = tid; e < nrow * p.gn_groups; e += 256) {
      const int rr = r_first + e / p.gn_groups, g = e % p.gn_groups;
      const long long* sp = p.gn_stats + ((size_t)rr * p.gn_pitch + g) * 2;
      gn_tab[e] = gn_rstd_cm(__ldcg(sp), __ldcg(sp + 1), inv_n, p.gn_eps);
    }
  }
  __syncthreads();

  float acc[4][4] = {};
  // A loader: thread -> (row lr, 4 consecutive channels); B loader: (weight row, 4 consecutive k)
  const int lr = tid >> 2, lc = (tid & 3) * 4;
  int kbase = 0;
  for (int seg = 0; seg < 2; ++seg) {
    const int cin = p.cin[seg];
    if (cin == 0) continue;
    const int mode = p.taps[seg];
    const int ntaps = mode == LM2A_TAPS_K1 ? 1 : (mode == LM2A_TAPS_K3 ? 3 : 4);
    for (int tap = 0; tap < ntaps; ++tap) {
      // source slot of output slot m for this tap, in the segment's own slab
      const long long m = m0 + lr;
      long long src = m;
      if (mode == LM2A_TAPS_K3) src = m + tap - 1;
      else if (mode == LM2A_TAPS_K4S2) src = 2 * m - 1 + tap;
      const long long src_rows = (seg == 0 && up2x) ? p.m : p.rows[seg];
      const bool in_slab = m < p.m && src >= 0 && src < src_rows;
      // operand transforms act on segment 0 at the OUTPUT resolution (k1 / k3 only)
      int r_src = 0, t_src = 0;
      if (in_slab && seg == 0 && (gn || up2x)) {
        r_src = (int)(src / p.tp);
        t_src = (int)(src - (long long)r_src * p.tp);
      }
      const bool live = in_slab && !(seg == 0 && (gn || up2x) && t_src >= p.t_valid);
      // x2 linear upsampling (align_corners): source slots of the low-resolution slab
      long long off0 = 0, off1 = 0;
      float wa = 1.f, wb = 0.f;
      if (live && seg == 0 && up2x) {
        const float scale = p.t_valid > 1 ? (float)(p.up_t_in - 1) / (float)(p.t_valid - 1) : 0.f;
        const float sp = scale * (float)t_src;
        const int i0 = (int)sp;
        const int i1 = i0 + (i0 < p.up_t_in - 1 ? 1 : 0);
        wb = sp - (float)i0;
        wa = 1.0f - wb;
        off0 = ((long long)r_src * p.up_tp_in + i0) * p.ld[0];
        off1 = ((long long)r_src * p.up_tp_in + i1) * p.ld[0];
      }
      for (int c0 = 0; c0 < cin; c0 += kBKc) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
          const int c = c0 + lc;
          if (seg == 0 && up2x) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(p.x[0] + off0 + c));
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.x[0] + off1 + c));
            a = make_float4(fmaf(wb, v.x, wa * u.x), fmaf(wb, v.y, wa * u.y),
                            fmaf(wb, v.z, wa * u.z), fmaf(wb, v.w, wa * u.w));
          } else {
            a = __ldg(reinterpret_cast<const float4*>(p.x[seg] + src * p.ld[seg] + c));
          }
          if (seg == 0 && gn) {
            const float2 sc = gn_tab[(r_src - r_first) * p.gn_groups + c / p.gn_cg];
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + c));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + c));
            a.x = silu_exact(fmaf(fmaf(a.x, sc.x, sc.y), g4.x, b4.x));
            a.y = silu_exact(fmaf(fmaf(a.y, sc.x, sc.y), g4.y, b4.y));
            a.z = silu_exact(fmaf(fmaf(a.z, sc.x, sc.y), g4.z, b4.z));
            a.w = silu_exact(fmaf(fmaf(a.w, sc.x, sc.y), g4.w, b4.w));
          }
        }
        const int kk = kbase + tap * cin + c0;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + lr < p.n_valid)
          b = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)(n0 + lr) * p.k_total + kk + lc));
        __syncthreads();   // previous chunk's reads are done
        As[lc + 0][lr] = a.x; As[lc + 1][lr] = a.y; As[lc + 2][lr] = a.z; As[lc + 3][lr] = a.w;
        Bs[lc + 0][lr] = b.x; Bs[lc + 1][lr] = b.y; Bs[lc + 2][lr] = b.z; Bs[lc + 3][lr] = b.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kBKc; ++k) {
          const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          const float ar[4] = {av.x, av.y, av.z, av.w};
          const float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
      }
    }
    kbase += ntaps * cin;
  }

  // epilogue: bias / FiLM, residual, exact GroupNorm sums, store
  const int n = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.m) continue;
    const int r = (int)(m / p.tp);
    const int t = (int)(m - (long long)r * p.tp);
    const bool valid = t < p.t_valid;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n + j;
      float o = 0.f;
      if (nn < p.n_valid && valid) {
        o = acc[i][j] + __ldg(p.bias + nn);
        if (p.film != nullptr) {
          const float* f = p.film + (p.film_ld != 0 ? (size_t)r * p.film_ld : 0);
          o = fmaf(o, 1.0f + __ldg(f + nn), __ldg(f + p.film_shift_off + nn));
        }
        if (p.residual != nullptr) o += __ldg(p.residual + (size_t)m * p.res_ld + nn);
      }
      v[j] = o;
    }
    if (p.stats != nullptr && valid && n < p.n_valid) {
      // the four columns of a thread lie in one group (groups are multiples of 8 channels)
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < p.n_valid) {
          s1 += v[j];
          s2 = fmaf(v[j], v[j], s2);
        }
      }
      unsigned long long* sp =
          p.stats + ((size_t)r * p.stats_pitch + (p.stats_c0 + n) / p.stats_cg) * 2;
      atomicAdd(sp, (unsigned long long)__float2ll_rn(s1 * kScale1));
      atomicAdd(sp + 1, (unsigned long long)__float2ll_rn(s2 * kScale2));
    }
    if (p.out_mode == LM2A_OUT_BF16_SLAB) {   // (fp32 slab in this path)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.n_valid) p.out[(size_t)m * p.out_ld + n + j] = v[j];
    } else if (valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.n_valid) p.out[((size_t)r * p.n_valid + n + j) * p.t_valid + t] = v[j];
    }
  }
}

// One warp per query row: scores for keys lane, lane + 32, ... (q broadcast from shared memory),
// exact softmax in base 2 (q arrives pre-scaled by log2(e) / sqrt(d_h)), then O[c] for channels
// lane, lane + 32, ... with the probabilities read back from shared memory.
constexpr int kAttnWarps = 4;
__global__ void __launch_bounds__(32 * kAttnWarps)
cross_attn_f32_kernel(const float* __restrict__ q, int q_ld, float* __restrict__ o, int o_ld,
                      const float* __restrict__ kv_m, const float* __restrict__ kv_t, int kv_ld,
                      const int* __restrict__ kv_slot, int tp, int t_valid, int lk, int e,
                      int heads) {
  extern __shared__ float sm[];
  const int dh = e / heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = sm + warp * (dh + lk);
  float* ps = qs + dh;
  const int r = blockIdx.z;
  const int stream = blockIdx.y / heads, h = blockIdx.y % heads;
  const int t = blockIdx.x * kAttnWarps + warp;
  if (t >= t_valid) return;
  const float* kv = (stream ? kv_t : kv_m) + (size_t)kv_slot[r] * lk * kv_ld;
  const float* qp = q + ((size_t)r * tp + t) * q_ld + stream * e + h * dh;
  for (int c = lane; c < dh; c += 32) qs[c] = qp[c];
  __syncwarp();
  float mx = -INFINITY;
  for (int k = lane; k < lk; k += 32) {
    const float* kp = kv + (size_t)k * kv_ld + h * dh;
    float s = 0.f;
    for (int c = 0; c < dh; c += 4) {
      const float4 kk = __ldg(reinterpret_cast<const float4*>(kp + c));
      s = fmaf(qs[c], kk.x, s);
      s = fmaf(qs[c + 1], kk.y, s);
      s = fmaf(qs[c + 2], kk.z, s);
      s = fmaf(qs[c + 3], kk.w, s);
    }
    ps[k] = s;
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float l = 0.f;
  for (int k = lane; k < lk; k += 32) {
    const float pv = exp2f(ps[k] - mx);
    ps[k] = pv;
    l += pv;
  }
  l = warp_sum(l);
  __syncwarp();
  const float inv = 1.0f / l;
  float* op = o + ((size_t)r * tp + t) * o_ld + stream * e + h * dh;
  for (int c = lane; c < dh; c += 32) {
    const float* vp = kv + e + h * dh + c;
    float a = 0.f;
    for (int k = 0; k < lk; ++k) a = fmaf(ps[k], __ldg(vp + (size_t)k * kv_ld), a);
    op[c] = a * inv;
  }
}

__global__ void __launch_bounds__(256)
bias_add_f32_kernel(const float* __restrict__ x, int x_ld, float* __restrict__ y, int y_ld,
                    const float* __restrict__ bias, long long slots, int tp, int t_valid, int c,
                    unsigned long long* __restrict__ stats, int stats_pitch, int stats_cg,
                    int stats_c0) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vpr = c >> 2;
  if (i >= slots * vpr) return;
  const long long m = i / vpr;
  const int cc = (int)(i - m * vpr) * 4;
  const int r = (int)(m / tp), t = (int)(m - (long long)r * tp);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < t_valid) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + (size_t)m * x_ld + cc));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + cc));
    v = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    if (stats != nullptr) {
      const float s1 = (v.x + v.y) + (v.z + v.w);
      const float s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
      unsigned long long* sp = stats + ((size_t)r * stats_pitch + (stats_c0 + cc) / stats_cg) * 2;
      atomicAdd(sp, (unsigned long long)__float2ll_rn(s1 * kScale1));
      atomicAdd(sp + 1, (unsigned long long)__float2ll_rn(s2 * kScale2));
    }
  }
  *reinterpret_cast<float4*>(y + (size_t)m * y_ld + cc) = v;
}

// fp32 [B, c, T] -> fp32 slab [copies * B, tp, ld] (zero pads), optionally clearing a region
__global__ void __launch_bounds__(256)
ingest_x_f32_kernel(const float* __restrict__ x, float* __restrict__ slab, int batch, int copies,
                    int c, int T, int tp, int ld, uint4* __restrict__ zero, long long zero_vec) {
  const long long nthr = (long long)gridDim.x * blockDim.x;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (long long i = tid; i < zero_vec; i += nthr) zero[i] = make_uint4(0u, 0u, 0u, 0u);
  const long long total = (long long)batch * tp * ld;
  for (long long i = tid; i < total; i += nthr) {
    const int cc = (int)(i % ld);
    const long long slot = i / ld;
    const int t = (int)(slot % tp);
    const int b = (int)(slot / tp);
    const float v = (cc < c && t < T) ? __ldg(x + ((size_t)b * c + cc) * T + t) : 0.f;
    for (int k = 0; k < copies; ++k) slab[((size_t)(k * batch + b) * tp + t) * ld + cc] = v;
  }
}

// fp32 [rows, T, c] -> fp32 slab [rows, tp, ld]
__global__ void __launch_bounds__(256)
ingest_seq_f32_kernel(const float* __restrict__ x, float* __restrict__ slab, long long total,
                      int T, int c, int tp, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cc = (int)(i % ld);
  const long long slot = i / ld;
  const int t = (int)(slot % tp);
  const long long r = slot / tp;
  slab[i] = (t < T && cc < c) ? __ldg(x + ((size_t)r * T + t) * c + cc) : 0.f;
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_conv1d_f32(void* stream, const lm2a_conv_desc* d) {
  using namespace lm2a;
  LM2A_REQUIRE(d != nullptr, "conv1d_f32: null descriptor");
  LM2A_REQUIRE(d->seg[0].x != nullptr && d->w != nullptr && d->bias != nullptr && d->out != nullptr,
               "conv1d_f32: null x / w / bias / out pointer");
  LM2A_REQUIRE(d->m > 0 && d->tp > 0 && d->t_valid > 0 && d->t_valid <= d->tp && d->m % d->tp == 0,
               "conv1d_f32: bad slab geometry m=%lld tp=%d t_valid=%d", (long long)d->m, d->tp,
               d->t_valid);
  LM2A_REQUIRE(d->n_valid > 0 && d->n_valid <= d->n_pad, "conv1d_f32: n_valid=%d n_pad=%d",
               d->n_valid, d->n_pad);
  F32Args a{};
  const bool up2x = d->in_up_t > 0;
  for (int s = 0; s < 2; ++s) {
    const lm2a_conv_seg& g = d->seg[s];
    if (g.x == nullptr) continue;
    LM2A_REQUIRE(g.cin > 0 && g.cin % 16 == 0 && g.ld >= g.cin && g.ld % 4 == 0 &&
                     (reinterpret_cast<uintptr_t>(g.x) & 15) == 0,
                 "conv1d_f32: seg %d cin=%d ld=%d (cin %% 16, ld %% 4, 16-byte base)", s, g.cin,
                 g.ld);
    LM2A_REQUIRE(g.taps >= LM2A_TAPS_K1 && g.taps <= LM2A_TAPS_K4S2, "conv1d_f32: seg %d taps=%d",
                 s, g.taps);
    if (s == 0 && up2x) {
      LM2A_REQUIRE(g.taps == LM2A_TAPS_K3 && 2 * g.rows == d->m && d->tp == 2 * d->in_up_tp &&
                       d->t_valid == 2 * d->in_up_t && d->in_gn_stats == nullptr,
                   "conv1d_f32: fused x2 upsampling geometry");
    } else if (g.taps == LM2A_TAPS_K4S2) {
      LM2A_REQUIRE(g.rows == 2 * d->m, "conv1d_f32: k4s2 needs 2 * m input slots");
    } else {
      LM2A_REQUIRE(g.rows == d->m, "conv1d_f32: seg %d slots != output slots", s);
    }
    a.x[s] = reinterpret_cast<const float*>(g.x);
    a.rows[s] = g.rows;
    a.ld[s] = g.ld;
    a.cin[s] = g.cin;
    a.taps[s] = g.taps;
    a.k_total += (g.taps == LM2A_TAPS_K1 ? 1 : (g.taps == LM2A_TAPS_K3 ? 3 : 4)) * g.cin;
  }
  a.w = reinterpret_cast<const float*>(d->w);
  a.m = d->m;
  a.tp = d->tp;
  a.t_valid = d->t_valid;
  a.n_valid = d->n_valid;
  a.bias = d->bias;
  a.film = d->film;
  a.film_ld = d->film_ld;
  a.film_shift_off = d->film_shift_off;
  a.residual = reinterpret_cast<const float*>(d->residual);
  a.res_ld = d->res_ld;
  a.out = reinterpret_cast<float*>(d->out);
  a.out_ld = d->out_ld;
  a.out_mode = d->out_mode;
  LM2A_REQUIRE(d->out_mode == LM2A_OUT_BF16_SLAB || d->out_mode == LM2A_OUT_F32_NCT,
               "conv1d_f32: bad out_mode %d", d->out_mode);
  a.stats = reinterpret_cast<unsigned long long*>(d->stats);
  a.stats_pitch = d->stats_pitch;
  a.stats_cg = d->stats_cg > 0 ? d->stats_cg : 32;
  a.stats_c0 = d->stats_c0;
  if (d->stats != nullptr)
    LM2A_REQUIRE(d->stats_cg % 4 == 0 && d->stats_c0 % 4 == 0 && d->out_mode == LM2A_OUT_BF16_SLAB,
                 "conv1d_f32: stats need a slab output and groups of multiples of 4 channels");
  a.gn_stats = reinterpret_cast<const long long*>(d->in_gn_stats);
  a.gn_gamma = d->in_gn_gamma;
  a.gn_beta = d->in_gn_beta;
  a.gn_pitch = d->in_gn_pitch;
  a.gn_groups = d->in_gn_groups;
  a.gn_eps = d->in_gn_eps;
  if (d->in_gn_stats != nullptr) {
    LM2A_REQUIRE(d->in_gn_gamma && d->in_gn_beta && d->in_gn_groups > 0 &&
                     d->seg[0].cin % d->in_gn_groups == 0 && d->in_gn_silu != 0 &&
                     (d->seg[0].taps == LM2A_TAPS_K1 || d->seg[0].taps == LM2A_TAPS_K3),
                 "conv1d_f32: bad input GroupNorm");
    a.gn_cg = d->seg[0].cin / d->in_gn_groups;
    LM2A_REQUIRE(a.gn_cg % 4 == 0, "conv1d_f32: channels per group must be a multiple of 4");
    LM2A_REQUIRE(((kBM + 2) / d->tp + 2) * d->in_gn_groups <= kMaxGnEnt,
                 "conv1d_f32: clips too short for the in-kernel GroupNorm table");
  }
  a.up_tp_in = d->in_up_tp;
  a.up_t_in = d->in_up_t;
  dim3 grid((unsigned)((d->m + kBM - 1) / kBM), (unsigned)((d->n_valid + kBN - 1) / kBN));
  conv_f32_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_cross_attn_f32(void* stream, const float* q, int32_t q_ld, float* o,
                                   int32_t o_ld, const float* kv_motion, const float* kv_text,
                                   int32_t kv_ld, const int32_t* kv_slot, int32_t slots,
                                   int32_t rows, int32_t tp, int32_t t_valid, int32_t lk, int32_t e,
                                   int32_t heads, int32_t n_streams) {
  using namespace lm2a;
  LM2A_REQUIRE(q && o && kv_motion && kv_text && kv_slot, "cross_attn_f32: null pointer");
  LM2A_REQUIRE(n_streams == 1 || n_streams == 2, "cross_attn_f32: n_streams=%d", n_streams);
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && slots > 0 && tp > 0 && t_valid > 0 && t_valid <= tp &&
                   lk > 0 && heads > 0 && e % heads == 0 && (e / heads) % 4 == 0,
               "cross_attn_f32: bad geometry");
  LM2A_REQUIRE(kv_ld >= 2 * e && kv_ld % 4 == 0 && q_ld >= n_streams * e && o_ld >= n_streams * e,
               "cross_attn_f32: bad pitches");
  const size_t smem = (size_t)kAttnWarps * (e / heads + lk) * sizeof(float);
  LM2A_REQUIRE(smem <= 200 * 1024, "cross_attn_f32: %d keys do not fit", lk);
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    LM2A_CUDA_OK(cudaFuncSetAttribute(cross_attn_f32_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((t_valid + kAttnWarps - 1) / kAttnWarps, n_streams * heads, rows);
  cross_attn_f32_kernel<<<grid, 32 * kAttnWarps, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      q, q_ld, o, o_ld, kv_motion, kv_text, kv_ld, kv_slot, tp, t_valid, lk, e, heads);
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_bias_add_f32(void* stream, const float* x, int32_t x_ld, float* y,
                                 int32_t y_ld, const float* bias, int64_t slots, int32_t tp,
                                 int32_t t_valid, int32_t c, void* stats, int32_t stats_pitch,
                                 int32_t stats_cg, int32_t stats_c0) {
  using namespace lm2a;
  LM2A_REQUIRE(x && y && bias, "bias_add_f32: null pointer");
  LM2A_REQUIRE(slots > 0 && tp > 0 && t_valid > 0 && t_valid <= tp && c > 0 && c % 4 == 0 &&
                   x_ld % 4 == 0 && y_ld % 4 == 0 && x_ld >= c && y_ld >= c,
               "bias_add_f32: bad geometry");
  if (stats != nullptr)
    LM2A_REQUIRE(stats_cg > 0 && stats_cg % 4 == 0 && stats_c0 % 4 == 0 && stats_pitch > 0,
                 "bias_add_f32: bad stats layout");
  const long long total = (long long)slots * (c / 4);
  bias_add_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0,
                        reinterpret_cast<cudaStream_t>(stream)>>>(
      x, x_ld, y, y_ld, bias, (long long)slots, tp, t_valid, c,
      reinterpret_cast<unsigned long long*>(stats), stats_pitch, stats_cg, stats_c0);
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_ingest_x_f32(void* stream, const float* x, float* slab, int32_t batch,
                                 int32_t copies, int32_t c, int32_t t, int32_t tp, int32_t ld,
                                 void* zero, int64_t zero_bytes) {
  using namespace lm2a;
  LM2A_REQUIRE(x && slab, "ingest_x_f32: null pointer");
  LM2A_REQUIRE(batch > 0 && copies > 0 && c > 0 && t > 0 && tp >= t && ld >= c,
               "ingest_x_f32: bad geometry");
  LM2A_REQUIRE(zero_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(zero) & 15) == 0,
               "ingest_x_f32: zero region must be 16-byte aligned / sized");
  const long long total = (long long)batch * tp * ld;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  ingest_x_f32_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, slab, batch, copies, c, t, tp, ld, reinterpret_cast<uint4*>(zero),
      (long long)(zero_bytes / 16));
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

extern "C" int lm2a_ingest_seq_f32(void* stream, const float* x, float* slab, int32_t rows,
                                   int32_t t, int32_t c, int32_t tp, int32_t ld) {
  using namespace lm2a;
  LM2A_REQUIRE(x && slab, "ingest_seq_f32: null pointer");
  LM2A_REQUIRE(rows > 0 && t > 0 && tp >= t && c > 0 && ld >= c, "ingest_seq_f32: bad geometry");
  const long long total = (long long)rows * tp * ld;
  ingest_seq_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(x, slab, total, t, c, tp, ld);
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
