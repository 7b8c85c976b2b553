// Development probe (not part of the product library): checks two tcgen05 operand forms the
// attention kernel relies on, against a host reference:
//   mode 0: D = A[smem, K-major SW128] * B[smem, MN-major SW128]      (b_major bit of the idesc)
//   mode 1: D = A[TMEM, packed bf16 pairs] * B[smem, MN-major SW128]  (.ts form)
// P: [128 x 64] bf16 row-major (queries x keys), C: [64 x 128] bf16 row-major (keys x channels),
// O = P C: [128 x 128] fp32.
#include "../../lm2a_b200/csrc/common.cuh"

using namespace lm2a;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
      "r"(v[30]), "r"(v[31])
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __nv_bfloat16* __restrict__ P, const __nv_bfloat16* __restrict__ C,
             float* __restrict__ O, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // C: two 64-channel panels, each [64 keys][128 B], 128B swizzle; P (mode 0): [128][128 B]
  const uint32_t c_off = 0, p_off = 16384, bar_off = 32768;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 16; i += 128) {   // 64 keys x 16 chunks of 16 B
    const int key = i >> 4, ch = i & 15, panel = ch >> 3, c8 = ch & 7;
    const uint4 q = *reinterpret_cast<const uint4*>(C + key * 128 + ch * 8);
    *reinterpret_cast<uint4*>(gen + c_off + panel * 8192 + key * 128 + ((c8 ^ (key & 7)) << 4)) = q;
  }
  for (int i = tid; i < 128 * 8; i += 128) {   // P: 128 rows x 8 chunks
    const int row = i >> 3, c8 = i & 7;
    const uint4 q = *reinterpret_cast<const uint4*>(P + row * 64 + c8 * 8);
    *reinterpret_cast<uint4*>(gen + p_off + row * 128 + ((c8 ^ (row & 7)) << 4)) = q;
  }
  const uint32_t bar = base + bar_off, slot = base + bar_off + 8;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  if (mode == 1) {
    uint32_t v[32];
    const uint32_t* pr = reinterpret_cast<const uint32_t*>(P + tid * 64);
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = pr[c];
    tmem_st32(tmem + lane_off, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before_sync();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    // kind::f16: c f32 (bit 4), a bf16 (bit 7), b bf16 (bit 10), b_major = MN (bit 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) |
                           ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      // B: 16 keys from row 16 k; MN blocks (64 channels) 8192 B apart, 8-key groups 1024 B apart
      const uint64_t bdesc = desc_sw128(base + c_off + k * 2048, 8192, 1024);
      if (mode == 0) {
        const uint64_t adesc = desc_sw128(base + p_off + k * 32, 16, 1024);
        umma_bf16_ss(tmem + 128, adesc, bdesc, idesc, k != 0);
      } else {
        umma_ts(tmem + 128, tmem + 8 * k, bdesc, idesc, k != 0);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after_sync();
  for (int c0 = 0; c0 < 128; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + 128 + lane_off + c0, v);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) O[tid * 128 + c0 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

extern "C" int umma_probe(const void* P, const void* C, float* O, int mode) {
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  probe_kernel<<<1, 128, 40000>>>(reinterpret_cast<const __nv_bfloat16*>(P),
                                  reinterpret_cast<const __nv_bfloat16*>(C), O, mode);
  return (int)cudaDeviceSynchronize();
}
