// Shared device helpers for the sm_100a kernels: raw PTX for mbarrier, TMA,
// tcgen05 (MMA / TMEM), plus the host-side error plumbing of the C ABI.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace lm2a {

// ---------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
constexpr int kMaxDevices = 64;
int current_device();
int num_sms();                          // of the current device
bool first_use_on_device(bool* flags);  // flags: a static bool[kMaxDevices] of the call site

// cuTensorMapEncodeTiled resolved through the runtime (no -lcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();

// Every kernel of the library is launched with programmatic dependent launch (PDL): its
// prologue (barrier init, TMEM allocation, tensor-map prefetch) overlaps the tail of the
// previous kernel in the stream; it calls pdl_wait() before touching global memory.
// LM2A_PDL=0 in the environment turns the attribute off (plain stream order).
bool pdl_enabled();

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                  cudaStream_t stream, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          cudaStream_t stream, Args&&... args) {
  return launch_kernel_cluster(kern, grid, block, smem, stream, 1u,
                               static_cast<Args&&>(args)...);
}
#endif

#define LM2A_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::lm2a::set_error(__VA_ARGS__);  \
      return 1;                        \
    }                                  \
  } while (0)

#define LM2A_CUDA_OK(expr)                                                   \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      ::lm2a::set_error("%s failed: %s (%s:%d)", #expr,                      \
                        cudaGetErrorString(_e), __FILE__, __LINE__);         \
      return 2;                                                              \
    }                                                                        \
  } while (0)

// -------------------------------------------------------------- device side
#ifdef __CUDACC__

// PDL: block until the preceding kernel in the stream has completed and its writes are
// visible (no-op when launched without the attribute); then let the next kernel's CTAs be
// scheduled as soon as resources free up.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug traps (cudaErrorLaunchFailure) after ~4 s
// instead of hanging the GPU until an external watchdog kills the box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      printf("lm2a: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// The same bounded wait without the diagnostic printf: for single-thread roles that run at a
// small setmaxnreg budget (the printf's argument marshalling costs registers and a stack frame
// in every loop that waits). A protocol bug still traps after ~4 s.
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > 4000000000ull) __trap();
  }
}

// Wait that may last for a whole mainloop (epilogue warps waiting for an accumulator): the
// try_wait carries a suspend-time hint, so the eight waiting warps do not keep the MIO queue
// busy with polls (the default time limit re-polls every ~50 ns) while other warps of the CTA
// work through shared memory. The thread still wakes as soon as the phase completes.
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(2000u)
        : "memory");
    if (done) return;
    if ((++spins & 0xff) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      printf("lm2a: mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---- TMA -----------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m,
                                            int32_t c0, int32_t c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m,
                                            int32_t c0, int32_t c1, int32_t c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src,
                                             int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
      :
      : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- tcgen05 / TMEM --------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a,
                                             uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every previously issued tcgen05.mma of this thread is done
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          bar)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns; thread i of the warp gets lane
// (taddr.lane + i), columns taddr.col .. +31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, "
      "[%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
        "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
        "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
        "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
        "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster share one UMMA -----------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile(
      "barrier.cluster.arrive.release.aligned;\n\t"
      "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(
                   cluster_addr)
               : "memory");
}
// TMA load whose completion may be signalled on the peer CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t smem_dst, const CUtensorMap* m,
                                                int32_t c0, int32_t c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr),
               "r"(ncols)
               : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem of both]; issued by the leader CTA only
__device__ __forceinline__ void umma_bf16_ss_cg2(uint32_t tmem_d, uint64_t desc_a,
                                                 uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(bar),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor: rows of 64 bf16
// (128 B), 8-row swizzle atoms 1024 B apart (SBO), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);   // start address
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO
  d |= static_cast<uint64_t>(1) << 46;                      // version
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: bf16 A/B (K-major), fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---- misc math -------------------------------------------------------------
__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
__device__ __forceinline__ float silu_accurate(float v) { return v / (1.0f + expf(-v)); }

// v * sigmoid(v) with sigmoid(v) = 0.5 + 0.5 tanh(v / 2): one MUFU op instead of two
__device__ __forceinline__ float silu_tanh(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
  return v * fmaf(0.5f, t, 0.5f);
}

// GroupNorm statistics of one (clip-row, group) from the producer's exact fixed-point sums
// (sum * 2^24, sum of squares * 2^20; see lm2a_conv_desc.stats): rstd and -mean * rstd, so that
// the normalisation is one FMA per element. Explicit round-to-nearest operations: every kernel
// that consumes the sums derives bit-identical scalars (no compiler-chosen FMA contraction).
__device__ __forceinline__ float2 gn_rstd_cm(long long s1, long long s2, double inv_n, float eps) {
  const double mean = __dmul_rn(__dmul_rn((double)s1, 1.0 / 16777216.0), inv_n);
  double var = __dsub_rn(__dmul_rn(__dmul_rn((double)s2, 1.0 / 1048576.0), inv_n),
                         __dmul_rn(mean, mean));
  var = var > 0.0 ? var : 0.0;
  const float rstd = (float)__ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(var, (double)eps)));
  return make_float2(rstd, __fmul_rn(-(float)mean, rstd));
}

// SiLU of v given vh = v / 2: v * sigmoid(v) = vh * tanh(vh) + vh. Callers fold the 1/2 into the
// affine map in front of it (GroupNorm scale / shift), which is exact in fp32.
__device__ __forceinline__ float silu_from_half(float vh) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(vh));
  return fmaf(vh, t, vh);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}

// ---------------------------------------------------------------- softmax probabilities
// p_c = exp2(s_c - m) for one query row's NC scores of a key tile, packed to bf16 pairs, and
// their fp32 sum (reference models/cross_attention.py:50-61: the softmax inside
// nn.MultiheadAttention; log2(e)/sqrt(d_h) is folded into W_q, so it is a bare exp2).
// The attention kernels are bound by the XU pipe (MUFU.EX2: 4 lanes/clk per scheduler = 8 cycles
// per warp instruction, 64 per thread and key tile) while the FMA pipe idles. LM2A_SOFTMAX_POLY =
// k in 0..4: of every four score PAIRS, k are evaluated on the FMA pipe instead - round to
// nearest integer n with the 1.5 * 2^23 trick, degree-3 polynomial for 2^f on f = x - n in
// [-0.5, 0.5] (max relative error 7.5e-5, 50x below the bf16 rounding of P that follows), n
// added into the exponent field - as packed f32x2 instructions (FADD2 / FFMA2: two lanes per
// issue slot), so that the loop's issue rate stays below the MUFU time it removes.
// Measured on B200 (profiles/r2_attn_tri_poly.txt): k = 1 is neutral (84.0 -> 83.6 us at d_h = 32,
// 72.9 -> 72.7 with three CTAs per SM), k = 2 / 3 slower (85.9 / 90.2 us): the softmax warps are
// latency-bound, not XU-bound (ncu: XU pipe 42 % busy, 0.36 IPC per scheduler). Default 0.
#ifndef LM2A_SOFTMAX_POLY
#define LM2A_SOFTMAX_POLY 0
#endif
__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f32x2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_sub(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 2^x for a pair, x <= 8 (lazy-rescale bound); x below -126 is clamped (result 2^-126 ~ 0)
__device__ __forceinline__ void exp2_pair_fma(float x0, float x1, float& p0, float& p1) {
  constexpr float kMagic = 12582912.0f;   // 1.5 * 2^23: x + kMagic rounds x to an integer
  const uint64_t x = f32x2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t magic = f32x2_pack(kMagic, kMagic);
  const uint64_t t = f32x2_add(x, magic);            // low mantissa bits = n (two's complement)
  const uint64_t f = f32x2_sub(x, f32x2_sub(t, magic));
  uint64_t p = f32x2_fma(f32x2_pack(0.0551716611f, 0.0551716611f), f,
                         f32x2_pack(0.2426111251f, 0.2426111251f));
  p = f32x2_fma(p, f, f32x2_pack(0.6932609677f, 0.6932609677f));
  p = f32x2_fma(p, f, f32x2_pack(0.9999280572f, 0.9999280572f));
  float t0, t1, q0, q1;
  f32x2_unpack(t, t0, t1);
  f32x2_unpack(p, q0, q1);
  p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// MASKED tiles (scores may be -inf) stay on the MUFU path
template <int NC, bool MASKED>
__device__ __forceinline__ float softmax_probs(const float (&s)[NC], float m,
                                               uint32_t (&pk)[NC / 2]) {
  constexpr int kPoly = MASKED ? 0 : LM2A_SOFTMAX_POLY;
  if constexpr (kPoly == 0) {
    float suma[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) {
      const float p0 = ex2_approx_ftz(s[2 * c] - m);
      const float p1 = ex2_approx_ftz(s[2 * c + 1] - m);
      suma[c & 3] += p0 + p1;
      pk[c] = pack_bf16x2(p0, p1);
    }
    return (suma[0] + suma[1]) + (suma[2] + suma[3]);
  } else {
    const uint64_t m2 = f32x2_pack(m, m);
    uint64_t sum2[2] = {0ull, 0ull};
#pragma unroll
    for (int c = 0; c < NC / 2; ++c) {
      float x0, x1, p0, p1;
      f32x2_unpack(f32x2_sub(f32x2_pack(s[2 * c], s[2 * c + 1]), m2), x0, x1);
      if ((c & 3) < kPoly) {
        exp2_pair_fma(x0, x1, p0, p1);
      } else {
        p0 = ex2_approx_ftz(x0);
        p1 = ex2_approx_ftz(x1);
      }
      sum2[c & 1] = f32x2_add(sum2[c & 1], f32x2_pack(p0, p1));
      pk[c] = pack_bf16x2(p0, p1);
    }
    float a, b;
    f32x2_unpack(f32x2_add(sum2[0], sum2[1]), a, b);
    return a + b;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

}  // namespace lm2a
