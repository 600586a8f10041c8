// Bring-up probe for the shared-memory / tensor-memory operand forms the attention kernels rely on:
// un-swizzled ("interleaved") core-matrix layouts for K-major and MN-major operands, and A-from-TMEM.
// D[128, N] = A[128, K] * B[N, K]^T with one CTA; used only by tests (tests/test_kernels_gpu.py).
#include "common.cuh"
#include "ptx.cuh"

namespace swinb200 {
using namespace ptx;

// un-swizzled operand tile: element (row, col) of a [rows x cols] bf16 tile lives at
//   (col / 8) * chunk_stride + row * 16 + (col % 8) * 2        ("[16-byte column chunk][row][8 elements]")
// K-major view  (rows = m/n, cols = k): core matrix = 8 rows x 16 B contiguous; LBO = chunk_stride, SBO = 128.
// MN-major view (rows = k, cols = m/n): same bytes;                                LBO = 128, SBO = chunk_stride.
__host__ __device__ constexpr uint64_t umma_smem_desc_nosw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D, int N, int K,
                  int a_mode, int b_mode, int pad16) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // A tile: K-major -> rows = 128 (m), cols = K ; MN-major -> rows = K (k), cols = 128 (m)
  const int a_rows = (a_mode == 1) ? K : 128, a_cols = (a_mode == 1) ? 128 : K;
  const int b_rows = (b_mode == 1) ? K : N, b_cols = (b_mode == 1) ? N : K;
  const uint32_t csa = a_rows * 16 + pad16 * 16, csb = b_rows * 16 + pad16 * 16;
  unsigned char* sA = smem;
  unsigned char* sB = smem + (((a_cols / 8) * csa + 1023) / 1024) * 1024;
  if (a_mode != 2)
    for (int i = tid; i < a_rows * (a_cols / 8); i += 128) {
      const int r = i / (a_cols / 8), c = i % (a_cols / 8);
      *reinterpret_cast<uint4*>(sA + c * csa + r * 16) = *reinterpret_cast<const uint4*>(A + (size_t)r * a_cols + c * 8);
    }
  for (int i = tid; i < b_rows * (b_cols / 8); i += 128) {
    const int r = i / (b_cols / 8), c = i % (b_cols / 8);
    *reinterpret_cast<uint4*>(sB + c * csb + r * 16) = *reinterpret_cast<const uint4*>(B + (size_t)r * b_cols + c * 8);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_a = tmem_base + 256;
  if (a_mode == 2) {
    // A from tensor memory: lane = row, 32-bit column c holds (k = 2c, 2c+1) as packed bf16
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + (size_t)tid * K);
    for (int c0 = 0; c0 < K / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = arow[c0 + j];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                       tmem_a + ((uint32_t)(warp * 32) << 16) + c0),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N, a_mode == 1, b_mode == 1);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bdesc = (b_mode == 0) ? umma_smem_desc_nosw(b0 + k * 2 * csb, csb, 128)
                                           : umma_smem_desc_nosw(b0 + k * 256, 128, csb);
      if (a_mode == 2) {
        umma_bf16_ts(tmem_base, tmem_a + k * 8, bdesc, idesc, k > 0);
      } else {
        const uint64_t adesc = (a_mode == 0) ? umma_smem_desc_nosw(a0 + k * 2 * csa, csa, 128)
                                             : umma_smem_desc_nosw(a0 + k * 256, 128, csa);
        umma_bf16_ss(tmem_base, adesc, bdesc, idesc, k > 0);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0, 900);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_32x16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace swinb200

using namespace swinb200;

extern "C" int swinb200_debug_umma_probe(const void* A, const void* B, float* D, int N, int K, int a_mode, int b_mode, int pad16,
                                         void* stream) {
  SWB_CHECK_ARG(A && B && D, "umma_probe: null pointer");
  SWB_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, "umma_probe: bad N/K");
  SWB_CHECK_ARG(a_mode >= 0 && a_mode <= 2 && (b_mode == 0 || b_mode == 1), "umma_probe: bad mode");
  const size_t smem = 200 * 1024;
  SWB_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, N, K, a_mode,
                                                            b_mode, pad16);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}
