// Gather-rate probe: how fast can 148 persistent CTAs pull 9x18-token windows of one 96-channel head into shared memory
// (TMA boxes [32 ch x 18 x 9], 64B-swizzled, three per operand) from
//   layout 0: token-major   qkv[B*H*W][2304]            (rows 4608 B apart, 192 useful bytes each)
//   layout 1: plane-major   qkv[B][24 planes][H*W][96]  (a head's rows 192 B apart: 3456-byte contiguous runs)
// nops operands (3 = q,k,v) per (window, head) item, double-buffered.  Prints GB/s.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe tools/gather_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma5(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

constexpr int kWh = 9, kWw = 18, kL = 162, kBox = kL * 64;   // bytes per box
constexpr int kStages = 2;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, int layout, int nops, int B, int H, int W, int heads,
                                             int order, unsigned long long* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kStages * 9 * 11264);
  const int nWw = W / kWw, nW = (H / kWh) * nWw;
  const int nitems = B * nW * heads;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int item, int stage) {
    int head, w, b;
    if (order == 0) { head = item % heads; w = (item / heads) % nW; b = item / (heads * nW); }      // heads fastest
    else { w = item % nW; head = (item / nW) % heads; b = item / (heads * nW); }                     // windows fastest
    const int wh = w / nWw, ww = w % nWw;
    mbar_expect(&bar[stage], (uint32_t)(nops * 3 * kBox));
    for (int op = 0; op < nops; ++op)
      for (int c = 0; c < 3; ++c) {
        unsigned char* dst = smem + (stage * 9 + op * 3 + c) * 11264;
        if (layout == 0) tma5(dst, &tm, &bar[stage], 0, (op * heads * 96 + head * 96) / 32 + c, ww * kWw, wh * kWh, b);
        else tma5(dst, &tm, &bar[stage], 0, c, ww * kWw, wh * kWh, (b * 3 + op) * heads + head);
      }
  };
  unsigned long long acc = 0;
  int it = 0;
  if (threadIdx.x == 0 && (int)blockIdx.x < nitems) issue(blockIdx.x, 0);
  for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
    const int stage = it & 1;
    if (threadIdx.x == 0 && item + (int)gridDim.x < nitems) issue(item + gridDim.x, stage ^ 1);
    while (!mbar_try(&bar[stage], (uint32_t)((it >> 1) & 1))) {}
    acc += *reinterpret_cast<const unsigned long long*>(smem + stage * 9 * 11264 + threadIdx.x * 64);
    __syncthreads();
  }
  if (acc == 0x1234567ull) *sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 1, H = 180, W = 360, heads = 8, C = 768;
  const size_t T = (size_t)B * H * W;
  void* buf;
  CK(cudaMalloc(&buf, T * 3 * C * 2));
  CK(cudaMemset(buf, 1, T * 3 * C * 2));
  void* fl;
  CK(cudaMalloc(&fl, 256 << 20));
  unsigned long long* sink;
  CK(cudaMalloc(&sink, 8));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * 9 * 11264 + 64));
  for (int layout = 0; layout < 2; ++layout)
    for (int nops = 3; nops >= 1; nops -= 2)
      for (int order = 0; order < 2; ++order) {
        CUtensorMap tm;
        cuuint32_t box[5] = {32, 1, (cuuint32_t)kWw, (cuuint32_t)kWh, 1}, estr[5] = {1, 1, 1, 1, 1};
        CUresult r;
        if (layout == 0) {
          const cuuint64_t row = 3 * C * 2;
          cuuint64_t dims[5] = {32, (cuuint64_t)3 * C / 32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
          cuuint64_t strides[4] = {64, row, row * W, row * W * H};
          r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
          const cuuint64_t row = 96 * 2;
          cuuint64_t dims[5] = {32, 3, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * 3 * heads};
          cuuint64_t strides[4] = {64, row, row * W, row * W * H};
          r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
          CK(cudaMemset(fl, rep, 256 << 20));     // flush L2
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0);
          probe<<<148, 128, kStages * 9 * 11264 + 64>>>(tm, layout, nops, B, H, W, heads, order, sink);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (ms < best) best = ms;
        }
        const double bytes = (double)T * heads * 96 * 2 * nops;
        printf("layout %s  operands %d  order %s : %.1f us  %.0f GB/s\n", layout ? "plane-major" : "token-major", nops,
               order ? "windows-fastest" : "heads-fastest", best * 1e3, bytes / best / 1e6);
      }
  return 0;
}
