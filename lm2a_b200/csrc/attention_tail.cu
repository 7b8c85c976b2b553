// The query rows that do not fill a 128-row tile of the tensor-core attention kernels.
//
// Production lengths leave T mod 128 = 4 / 2 / 1 query rows per (clip-row, stream, head) at
// levels 0 / 1 / 2 (T = 516 / 258 / 129). In cross_attn_tc_kernel those rows occupy a CTA of
// their own (20 / 33 % of the launch's CTAs, each holding a CTA slot for the whole key walk with
// one of its four softmax warps active); in the condition-slab kernel the one leftover row of
// the eight heads forms a ninth tile that costs its CTA a third serial round. This kernel
// computes exactly those rows on the CUDA cores - softmax(q k^T) v of reference
// models/cross_attention.py:50-61 for n_tail <= 8 queries per (clip-row, stream, head), a few
// hundred MFLOP per launch - so that the tensor-core launch covers whole tiles only. The launch
// plan puts it on a parallel graph branch next to the tensor-core launch (both read the same Q
// slab and write disjoint rows of O); it fits beside that kernel's CTAs (256 threads, 9 - 25 KB
// of shared memory).
//
// One CTA of 8 warps per (clip-row, stream, head) - in condition mode per (clip-row, stream,
// group of heads): the heads share K = V = C - handling up to 8 queries against one K / V pair:
//   scores  one KEY per thread (its row read with independent 16-byte loads), the dot products
//           with all queries of the CTA (fp32 copies in shared memory, broadcast reads); q arrives
//           pre-scaled by log2(e)/sqrt(d_h)
//   softmax one warp per query: max / exp2 / sum over its score row in shared memory, in fp32
//           (the probabilities are NOT rounded to bf16 here)
//   P V     V rows [Lk, d_h] (per-head mode: the V half of the K | V projection output; condition
//           mode: C again): warp = key segment, 16-byte loads, all queries at once; the segments
//           are summed in a fixed order (deterministic: no atomics)
#include "../../include/lm2a_b200.h"

#include "common.cuh"

namespace lm2a {
namespace {

constexpr int kTailWarps = 8;
constexpr int kTailThreads = kTailWarps * 32;
constexpr int kTailQ = 8;   // queries per CTA

struct TailArgs {
  const __nv_bfloat16* q;
  __nv_bfloat16* o;
  const __nv_bfloat16* k[2];    // per stream: K rows [slots * lk, k_ld] (COND: the condition slab)
  const __nv_bfloat16* v[2];    // per stream: V rows [slots * lk, v_ld]
  const int* kv_slot;
  int q_ld, o_ld, k_ld, v_ld;
  int tp, t0, n_tail, lk, lk_pad, e, heads;
  int hg;         // heads per CTA (1 unless the heads share K / V)
  int shared_kv;  // 1: every head reads channels [0, D) of k / v (condition mode)
};

template <int D>
__global__ void __launch_bounds__(kTailThreads)
cross_attn_tail_kernel(const TailArgs a) {
  extern __shared__ __align__(16) float tail_smem[];
  float* q_s = tail_smem;                          // [kTailQ][D]
  float* red = q_s + kTailQ * D;                   // [kTailWarps][kTailQ][D]: P V partial sums
  float* l_s = red + kTailWarps * kTailQ * D;      // [kTailQ]
  float* s_s = l_s + kTailQ;                       // [nq][lk_pad]: scores, then probabilities
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = blockIdx.z, stream = blockIdx.y, h0 = blockIdx.x * a.hg;
  const int hg = min(a.hg, a.heads - h0);
  const int nq = hg * a.n_tail;                    // <= kTailQ
  pdl_wait();
  pdl_launch_dependents();
  const int slot = a.kv_slot[r];
  const int kv_col = a.shared_kv ? 0 : h0 * D;
  const __nv_bfloat16* kbase = a.k[stream] + (size_t)slot * a.lk * a.k_ld + kv_col;
  const __nv_bfloat16* vbase = a.v[stream] + (size_t)slot * a.lk * a.v_ld + kv_col;
  // query qi = (head h0 + qi / n_tail, row t0 + qi % n_tail)
  auto q_row = [&](int qi) { return (size_t)r * a.tp + a.t0 + qi % a.n_tail; };
  auto q_col = [&](int qi) { return stream * a.e + (h0 + qi / a.n_tail) * D; };
  for (int i = tid; i < kTailQ * D; i += kTailThreads) {
    const int qi = i / D, d = i - qi * D;
    q_s[i] = qi < nq ? __bfloat162float(a.q[q_row(qi) * a.q_ld + q_col(qi) + d]) : 0.f;
  }
  __syncthreads();
  // ---- scores: one key per thread
  for (int key = tid; key < a.lk; key += kTailThreads) {
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)key * a.k_ld);
    uint4 kv[D / 8];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) kv[i] = __ldg(kr + i);
    float acc[kTailQ];
#pragma unroll
    for (int qi = 0; qi < kTailQ; ++qi) acc[qi] = 0.f;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      const float2 v0 = unpack_bf16x2(kv[i].x), v1 = unpack_bf16x2(kv[i].y);
      const float2 v2 = unpack_bf16x2(kv[i].z), v3 = unpack_bf16x2(kv[i].w);
#pragma unroll
      for (int qi = 0; qi < kTailQ; ++qi) {
        const float4 qa = *reinterpret_cast<const float4*>(q_s + qi * D + 8 * i);
        const float4 qb = *reinterpret_cast<const float4*>(q_s + qi * D + 8 * i + 4);
        float s = acc[qi];
        s = fmaf(qa.x, v0.x, s);
        s = fmaf(qa.y, v0.y, s);
        s = fmaf(qa.z, v1.x, s);
        s = fmaf(qa.w, v1.y, s);
        s = fmaf(qb.x, v2.x, s);
        s = fmaf(qb.y, v2.y, s);
        s = fmaf(qb.z, v3.x, s);
        s = fmaf(qb.w, v3.y, s);
        acc[qi] = s;
      }
    }
#pragma unroll
    for (int qi = 0; qi < kTailQ; ++qi)
      if (qi < nq) s_s[qi * a.lk_pad + key] = acc[qi];
  }
  __syncthreads();
  // ---- softmax: warp qi owns query qi
  if (warp < nq) {
    float* sr = s_s + warp * a.lk_pad;
    float mx = -INFINITY;
    for (int key = lane; key < a.lk; key += 32) mx = fmaxf(mx, sr[key]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float l = 0.f;
    for (int key = lane; key < a.lk; key += 32) {
      const float p = exp2f(sr[key] - mx);
      sr[key] = p;
      l += p;
    }
    l = warp_sum(l);
    if (lane == 0) l_s[warp] = l;
  }
  __syncthreads();
  // ---- P V: warp = key segment, every query of the CTA at once. A key's V row is D / 8 lanes
  // of 16 bytes, so one warp load covers 32 / (D / 8) keys; each lane keeps 8 channels x 8
  // queries, the key sub-lanes are folded with shuffles, the 8 segments through shared memory
  // in a fixed order (deterministic)
  constexpr int LPK = D / 8, KPL = 32 / LPK;
  const int cg = lane % LPK, ks = lane / LPK;
  const int per = (a.lk + kTailWarps - 1) / kTailWarps;
  const int k0 = warp * per, k1 = min(a.lk, k0 + per);
  float acc[kTailQ][8];
#pragma unroll
  for (int qq = 0; qq < kTailQ; ++qq)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[qq][c] = 0.f;
  const __nv_bfloat16* vb = vbase + 8 * cg;
#pragma unroll 2
  for (int kb = k0; kb < k1; kb += KPL) {
    const int key = kb + ks;
    if (key < k1) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(vb + (size_t)key * a.v_ld));
      const float2 v0 = unpack_bf16x2(v.x), v1 = unpack_bf16x2(v.y);
      const float2 v2 = unpack_bf16x2(v.z), v3 = unpack_bf16x2(v.w);
#pragma unroll
      for (int qq = 0; qq < kTailQ; ++qq) {
        const float p = qq < nq ? s_s[qq * a.lk_pad + key] : 0.f;
        acc[qq][0] = fmaf(p, v0.x, acc[qq][0]);
        acc[qq][1] = fmaf(p, v0.y, acc[qq][1]);
        acc[qq][2] = fmaf(p, v1.x, acc[qq][2]);
        acc[qq][3] = fmaf(p, v1.y, acc[qq][3]);
        acc[qq][4] = fmaf(p, v2.x, acc[qq][4]);
        acc[qq][5] = fmaf(p, v2.y, acc[qq][5]);
        acc[qq][6] = fmaf(p, v3.x, acc[qq][6]);
        acc[qq][7] = fmaf(p, v3.y, acc[qq][7]);
      }
    }
  }
#pragma unroll
  for (int off = LPK; off < 32; off <<= 1)
#pragma unroll
    for (int qq = 0; qq < kTailQ; ++qq)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[qq][c] += __shfl_xor_sync(0xffffffffu, acc[qq][c], off);
  if (ks == 0) {
#pragma unroll
    for (int qq = 0; qq < kTailQ; ++qq) {
      float4* dst = reinterpret_cast<float4*>(red + (warp * kTailQ + qq) * D + 8 * cg);
      dst[0] = make_float4(acc[qq][0], acc[qq][1], acc[qq][2], acc[qq][3]);
      dst[1] = make_float4(acc[qq][4], acc[qq][5], acc[qq][6], acc[qq][7]);
    }
  }
  __syncthreads();
  // segments summed in a fixed order, scaled by 1 / l, stored as bf16
  for (int i = tid; i < nq * D; i += kTailThreads) {
    const int qq = i / D, d = i - qq * D;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kTailWarps; ++w) sum += red[(w * kTailQ + qq) * D + d];
    a.o[q_row(qq) * a.o_ld + q_col(qq) + d] = __float2bfloat16(sum / l_s[qq]);
  }
}

}  // namespace
}  // namespace lm2a

extern "C" int lm2a_cross_attn_tail_bf16(void* stream, const void* q, int32_t q_ld, void* o,
                                         int32_t o_ld, const void* k_motion,
                                         const void* v_motion, const void* k_text,
                                         const void* v_text, int32_t k_ld, int32_t v_ld,
                                         const int32_t* kv_slot, int32_t slots, int32_t rows,
                                         int32_t tp, int32_t t0, int32_t n_tail, int32_t lk,
                                         int32_t e, int32_t heads, int32_t n_streams,
                                         int32_t shared_kv) {
  using namespace lm2a;
  LM2A_REQUIRE(n_streams == 1 || n_streams == 2, "cross_attn_tail: n_streams=%d (1 or 2)",
               n_streams);
  LM2A_REQUIRE(q && o && k_motion && k_text && v_motion && v_text && kv_slot,
               "cross_attn_tail: null pointer");
  LM2A_REQUIRE(rows > 0 && rows <= 65535 && slots > 0 && tp > 0 && t0 >= 0 && n_tail > 0 &&
                   n_tail <= kTailQ && t0 + n_tail <= tp && lk > 0 && heads > 0 && heads <= 65535,
               "cross_attn_tail: bad geometry (t0=%d n_tail=%d tp=%d lk=%d)", t0, n_tail, tp, lk);
  LM2A_REQUIRE(e % heads == 0, "cross_attn_tail: e=%d not divisible by heads=%d", e, heads);
  const int dh = e / heads;
  LM2A_REQUIRE(dh == 32 || dh == 64 || dh == 128, "cross_attn_tail: head dim %d (32, 64 or 128)",
               dh);
  const int lk_pad = (lk + 7) / 8 * 8;
  const int kv_cols = shared_kv ? dh : e;
  LM2A_REQUIRE(q_ld % 8 == 0 && o_ld % 8 == 0 && k_ld % 8 == 0 && v_ld % 8 == 0 &&
                   q_ld >= n_streams * e && o_ld >= n_streams * e && k_ld >= kv_cols &&
                   v_ld >= kv_cols,
               "cross_attn_tail: bad pitches (q_ld=%d o_ld=%d k_ld=%d v_ld=%d e=%d)", q_ld, o_ld,
               k_ld, v_ld, e);
  LM2A_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(o) |
                 reinterpret_cast<uintptr_t>(k_motion) | reinterpret_cast<uintptr_t>(k_text) |
                 reinterpret_cast<uintptr_t>(v_motion) | reinterpret_cast<uintptr_t>(v_text)) &
                15) == 0,
               "cross_attn_tail: tensors must be 16-byte aligned");
  TailArgs a;
  a.q = reinterpret_cast<const __nv_bfloat16*>(q);
  a.o = reinterpret_cast<__nv_bfloat16*>(o);
  a.k[0] = reinterpret_cast<const __nv_bfloat16*>(k_motion);
  a.k[1] = reinterpret_cast<const __nv_bfloat16*>(k_text);
  a.v[0] = reinterpret_cast<const __nv_bfloat16*>(v_motion);
  a.v[1] = reinterpret_cast<const __nv_bfloat16*>(v_text);
  a.kv_slot = kv_slot;
  a.q_ld = q_ld;
  a.o_ld = o_ld;
  a.k_ld = k_ld;
  a.v_ld = v_ld;
  a.tp = tp;
  a.t0 = t0;
  a.n_tail = n_tail;
  a.lk = lk;
  a.lk_pad = lk_pad;
  a.e = e;
  a.heads = heads;
  a.shared_kv = shared_kv ? 1 : 0;
  a.hg = 1;
  if (shared_kv) {   // heads that share K / V share a CTA: up to kTailQ queries
    a.hg = kTailQ / n_tail;
    if (a.hg > heads) a.hg = heads;
  }
  const int nq_max = a.hg * n_tail;
  const size_t smem = ((size_t)(kTailQ + kTailWarps * kTailQ) * dh + kTailQ +
                       (size_t)nq_max * lk_pad) * sizeof(float);
  LM2A_REQUIRE(smem <= 200 * 1024, "cross_attn_tail: lk=%d does not fit in shared memory", lk);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((heads + a.hg - 1) / a.hg, n_streams, rows);
#define LM2A_TAIL_CASE(D)                                                                      \
  case D: {                                                                                    \
    static bool configured[kMaxDevices] = {};                                                  \
    if (first_use_on_device(configured))                                                       \
      LM2A_CUDA_OK(cudaFuncSetAttribute(cross_attn_tail_kernel<D>,                             \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                        200 * 1024));                                          \
    LM2A_CUDA_OK(                                                                              \
        launch_kernel(cross_attn_tail_kernel<D>, grid, dim3(kTailThreads), smem, st, a));      \
    break;                                                                                     \
  }
  switch (dh) {
    LM2A_TAIL_CASE(32)
    LM2A_TAIL_CASE(64)
    LM2A_TAIL_CASE(128)
  }
#undef LM2A_TAIL_CASE
  LM2A_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}
