// Bandwidth-bound kernels of the SwinV2 hot path: parameter staging, patchify / unpatchify,
// LayerNorm + residual (fwd/bwd), bias-gradient column sums, q/k L2-normalisation,
// latitude-weighted L2 loss (fwd/bwd).  All accesses are 128-bit vectorised and coalesced;
// reductions use warp shuffles first, shared memory second, fp32 atomics last.
#include <stdarg.h>
#include "common.cuh"

namespace swinb200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast
// ------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    ld8(src + i * 8, v);
    st8(dst + i * 8, v);
  }
  if (blockIdx.x == 0) {
    for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 2-D transpose (pos_embed (C, T) <-> (T, C)), optional sum over a leading batch of sources
// ------------------------------------------------------------------------------------------------
// dst[c, r] = sum_b src[b, r, c]   (src: (nb, R, Cc) ; dst: (Cc, R))
__global__ void transpose_sum_kernel(const float* __restrict__ src, float* __restrict__ dst, int nb, int R, int Cc) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // (32, 8)
  for (int j = ty; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    float acc = 0.f;
    if (r < R && c < Cc)
      for (int b = 0; b < nb; ++b) acc += src[((size_t)b * R + r) * Cc + c];
    tile[j][tx] = acc;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < R && c < Cc) dst[(size_t)c * R + r] = tile[tx][j];
  }
}

// 64 x 64 tiles with 128-bit global accesses on both sides (R and Cc multiples of 4, 16-byte aligned pointers)
__global__ void __launch_bounds__(256) transpose_sum_vec_kernel(const float* __restrict__ src, float* __restrict__ dst, int nb,
                                                                int R, int Cc) {
  __shared__ float tile[64][65];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + ty + 16 * j, c = c0 + tx * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R && c < Cc)
      for (int b = 0; b < nb; ++b) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + ((size_t)b * R + r) * Cc + c));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    float* t = &tile[ty + 16 * j][tx * 4];
    t[0] = acc.x; t[1] = acc.y; t[2] = acc.z; t[3] = acc.w;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + ty + 16 * j, r = r0 + tx * 4;
    if (r < R && c < Cc) {
      const int cc = ty + 16 * j;
      *reinterpret_cast<float4*>(dst + (size_t)c * R + r) =
          make_float4(tile[tx * 4][cc], tile[tx * 4 + 1][cc], tile[tx * 4 + 2][cc], tile[tx * 4 + 3][cc]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// patchify / unpatchify
// ------------------------------------------------------------------------------------------------
constexpr int kPatchTok = 16;  // tokens per CTA

// img (B, C, Hi, Wi) fp32 -> out (T, C*P*P).  P == 4 (float4 along the patch row).
template <typename T, int ORDER>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, T* __restrict__ out, int B, int C,
                                                       int Hi, int Wi, int ntok) {
  constexpr int P = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int K = C * P * P;
  const int pitch = K + 8;  // keeps rows 16-byte aligned, de-phases banks
  const int H = Hi / P, W = Wi / P;
  const int t0 = blockIdx.x * kPatchTok;
  const int ntile = min(kPatchTok, ntok - t0);
  // gather: one float4 (4 q-values) per (c, p, token)
  const int items = C * P * kPatchTok;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tok = it % kPatchTok;
    const int cp = it / kPatchTok;
    const int p = cp % P, c = cp / P;
    if (tok >= ntile) continue;
    const int t = t0 + tok;
    const int j = t % W, i = (t / W) % H, b = t / (W * H);
    const float4 v = *reinterpret_cast<const float4*>(img + (((size_t)b * C + c) * Hi + (i * P + p)) * Wi + j * P);
    T* row = tile + (size_t)tok * pitch;
    if (ORDER == 0) {
      T* d = row + (c * P + p) * P;
      Act<T>::st(d + 0, v.x); Act<T>::st(d + 1, v.y); Act<T>::st(d + 2, v.z); Act<T>::st(d + 3, v.w);
    } else {
      Act<T>::st(row + (p * P + 0) * C + c, v.x);
      Act<T>::st(row + (p * P + 1) * C + c, v.y);
      Act<T>::st(row + (p * P + 2) * C + c, v.z);
      Act<T>::st(row + (p * P + 3) * C + c, v.w);
    }
  }
  __syncthreads();
  // contiguous store of ntile rows x K elements, 8 elements per thread-iteration
  const int k8 = K / 8;
  for (int it = threadIdx.x; it < ntile * k8; it += blockDim.x) {
    const int tok = it / k8, kc = it % k8;
    float v[8];
    ld8(tile + (size_t)tok * pitch + kc * 8, v);
    st8(out + (size_t)(t0 + tok) * K + kc * 8, v);
  }
}

// y (T, P*P*Co) -> out (B, Co, Hi, Wi) (+ skip[:, :Co]); ORDER 1: columns (p, q, c) (output head), ORDER 0: (c, p, q)
// (gradient of the PatchEmbed im2col, i.e. dL/d image for multi-step rollouts)
// Same im2col (order 0) reading the image channels from up to 8 separate tensors -- the field, the zenith-angle channel
// and the static land-mask / orography features (utils/preprocess_utils.py:50-68, networks/helpers.py:36-40) -- so the
// concatenated (B, 77, 720, 1440) input is never written: a source with batch stride 0 is shared by every sample.
constexpr int kPatchMaxSrc = 8;
struct PatchSrcs {
  const float* ptr[kPatchMaxSrc];
  long long bstride[kPatchMaxSrc];   // elements between samples (0: broadcast over the batch)
  int c0[kPatchMaxSrc + 1];          // first concatenated-channel index of each source; c0[n] = C
  int n;
};
template <typename T>
__global__ void __launch_bounds__(256) patchify_cat_kernel(const __grid_constant__ PatchSrcs src, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, T* __restrict__ out, int B, int C,
                                                           int Hi, int Wi, int ntok) {
  constexpr int P = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int K = C * P * P;
  const int pitch = K + 8;
  const int H = Hi / P, W = Wi / P;
  const int t0 = blockIdx.x * kPatchTok;
  const int ntile = min(kPatchTok, ntok - t0);
  const int items = C * P * kPatchTok;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tok = it % kPatchTok;
    const int cp = it / kPatchTok;
    const int p = cp % P, c = cp / P;
    if (tok >= ntile) continue;
    int s = 0;
    while (s + 1 < src.n && c >= src.c0[s + 1]) ++s;
    const int t = t0 + tok;
    const int j = t % W, i = (t / W) % H, b = t / (W * H);
    const float* base = src.ptr[s] + (size_t)b * src.bstride[s] + ((size_t)(c - src.c0[s]) * Hi + (i * P + p)) * Wi + j * P;
    float4 v = *reinterpret_cast<const float4*>(base);
    if (mean != nullptr) {   // z-score of the loaders (data_loader_era5_dali.py:77-90): (x - mean_c) / std_c, true division
      const float m = __ldg(mean + c), sd = __ldg(stdv + c);
      v.x = (v.x - m) / sd; v.y = (v.y - m) / sd; v.z = (v.z - m) / sd; v.w = (v.w - m) / sd;
    }
    T* d = tile + (size_t)tok * pitch + (c * P + p) * P;
    Act<T>::st(d + 0, v.x); Act<T>::st(d + 1, v.y); Act<T>::st(d + 2, v.z); Act<T>::st(d + 3, v.w);
  }
  __syncthreads();
  const int k8 = K / 8;
  for (int it = threadIdx.x; it < ntile * k8; it += blockDim.x) {
    const int tok = it / k8, kc = it % k8;
    float v[8];
    ld8(tile + (size_t)tok * pitch + kc * 8, v);
    st8(out + (size_t)(t0 + tok) * K + kc * 8, v);
  }
}

template <typename T, int ORDER>
__global__ void __launch_bounds__(256) unpatchify_kernel(const T* __restrict__ y, const float* __restrict__ skip,
                                                         int skip_chans, const float* __restrict__ smean,
                                                         const float* __restrict__ sstd, float* __restrict__ out, int B, int Co,
                                                         int Hi, int Wi, int ntok) {
  constexpr int P = 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int K = Co * P * P;
  const int pitch = K + 8;
  const int H = Hi / P, W = Wi / P;
  const int t0 = blockIdx.x * kPatchTok;
  const int ntile = min(kPatchTok, ntok - t0);
  const int k8 = K / 8;
  for (int it = threadIdx.x; it < ntile * k8; it += blockDim.x) {
    const int tok = it / k8, kc = it % k8;
    float v[8];
    ld8(y + (size_t)(t0 + tok) * K + kc * 8, v);
    st8(tile + (size_t)tok * pitch + kc * 8, v);
  }
  __syncthreads();
  const int items = Co * P * kPatchTok;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int tok = it % kPatchTok;
    const int cp = it / kPatchTok;
    const int p = cp % P, c = cp / P;
    if (tok >= ntile) continue;
    const int t = t0 + tok;
    const int j = t % W, i = (t / W) % H, b = t / (W * H);
    const T* row = tile + (size_t)tok * pitch;
    float4 v;
    if (ORDER == 1) {
      v.x = Act<T>::ld(row + (p * P + 0) * Co + c);
      v.y = Act<T>::ld(row + (p * P + 1) * Co + c);
      v.z = Act<T>::ld(row + (p * P + 2) * Co + c);
      v.w = Act<T>::ld(row + (p * P + 3) * Co + c);
    } else {
      const T* s4 = row + (c * P + p) * P;
      v.x = Act<T>::ld(s4 + 0); v.y = Act<T>::ld(s4 + 1); v.z = Act<T>::ld(s4 + 2); v.w = Act<T>::ld(s4 + 3);
    }
    const size_t pix = (size_t)(i * P + p) * Wi + j * P;
    if (skip != nullptr) {
      float4 s = *reinterpret_cast<const float4*>(skip + ((size_t)b * skip_chans + c) * Hi * Wi + pix);
      if (smean != nullptr) {   // the skip is the raw field: z-score it exactly as the im2col does
        const float m = __ldg(smean + c), sd = __ldg(sstd + c);
        s.x = (s.x - m) / sd; s.y = (s.y - m) / sd; s.z = (s.z - m) / sd; s.w = (s.w - m) / sd;
      }
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    *reinterpret_cast<float4*>(out + ((size_t)b * Co + c) * Hi * Wi + pix) = v;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm + residual
// ------------------------------------------------------------------------------------------------
// One warp per row; lane owns chunks of 8 channels at c = lane*8 + 256*j.
template <typename T, int NCHUNK>
__global__ void __launch_bounds__(256) ln_residual_fwd_kernel(const T* __restrict__ z, const float* __restrict__ x_in,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ sample_scale,
                                                              const float* __restrict__ pos, float* __restrict__ x_out,
                                                              T* __restrict__ xb_out, float* __restrict__ stats, int rows,
                                                              int C, int rows_per_sample, float eps) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int warp_global = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * warps_per_block;
  const float invC = 1.0f / (float)C;
  // rows are visited from the END of the tensor: the GEMM that produced z wrote its last rows last, so they are still in L2
  for (int ri = warp_global; ri < rows; ri += nwarps) {
    const int row = rows - 1 - ri;
    float v[NCHUNK][8], xi[NCHUNK][8];
    float sum = 0.f;
    // every global load of the row (branch output and residual stream) is issued before the first reduction
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int c = lane * 8 + j * 256;
      if (c < C) {
        ld8(z + (size_t)row * C + c, v[j]);
        if (x_in) ld8(x_in + (size_t)row * C + c, xi[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int c = lane * 8 + j * 256;
      if (c < C) {
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += v[j][e];
      }
    }
    const float mean = warp_sum(sum) * invC;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int c = lane * 8 + j * 256;
      if (c < C) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[j][e] - mean;
          sq += d * d;
        }
      }
    }
    const float var = warp_sum(sq) * invC;
    const float rstd = rsqrtf(var + eps);
    const float sc = sample_scale ? sample_scale[row / rows_per_sample] : 1.0f;
    const int prow = row % rows_per_sample;
#pragma unroll
    for (int j = 0; j < NCHUNK; ++j) {
      const int c = lane * 8 + j * 256;
      if (c < C) {
        float g[8], b[8], r[8];
        ld8(gamma + c, g);
        ld8(beta + c, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = (v[j][e] - mean) * rstd * g[e] + b[e];
        if (pos) {
          float pe[8];
          ld8(pos + (size_t)prow * C + c, pe);
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] += pe[e];
        }
        if (sample_scale) {
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] *= sc;
        }
        if (x_in) {
#pragma unroll
          for (int e = 0; e < 8; ++e) r[e] += xi[j][e];
        }
        st8(x_out + (size_t)row * C + c, r);
        st8(xb_out + (size_t)row * C + c, r);
      }
    }
    if (lane == 0) {
      stats[2 * (size_t)row] = mean;
      stats[2 * (size_t)row + 1] = rstd;
    }
  }
}

// backward: dz = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dx*scale*gamma
//
// Persistent CTAs stream row tiles through shared memory with the bulk-copy engine: a tile of TR consecutive rows is one
// contiguous region of dx (fp32) and of z, so each tile is two cp.async.bulk loads (double-buffered, mbarrier
// complete_tx) and one cp.async.bulk store of dz -- HBM traffic never waits on the arithmetic.  Per tile:
//   pass 1 (warp per row)            row means of g and g*xhat via warp shuffles -> shared scalars
//   pass 2 (thread per 8 channels)   dz rows into the staged output tile; per-channel sums for dgamma / dbeta /
//                                    dbias_prev stay in 24 registers for the CTA's whole lifetime
// and one fp32 atomic per channel per thread group at the end.
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(smem_addr_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_addr_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void lnb_mbar_init(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void lnb_mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lnb_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_addr_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) {
      printf("swinb200: ln_residual_bwd tile wait timed out (block %d)\n", (int)blockIdx.x);
      __trap();
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) ln_residual_bwd_kernel(const float* __restrict__ dx, const T* __restrict__ z,
                                                                 const float* __restrict__ stats, const float* __restrict__ gamma,
                                                                 const float* __restrict__ sample_scale, T* __restrict__ dz,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 float* __restrict__ dbias_prev, int rows, int C,
                                                                 int rows_per_sample, int TR) {
  extern __shared__ __align__(128) unsigned char lsm[];
  const size_t dx_bytes = (size_t)TR * C * 4, z_bytes = (size_t)TR * C * sizeof(T);
  float* dxs[2] = {reinterpret_cast<float*>(lsm), reinterpret_cast<float*>(lsm + dx_bytes)};
  T* zs[2] = {reinterpret_cast<T*>(lsm + 2 * dx_bytes), reinterpret_cast<T*>(lsm + 2 * dx_bytes + z_bytes)};
  T* outs[2] = {reinterpret_cast<T*>(lsm + 2 * dx_bytes + 2 * z_bytes), reinterpret_cast<T*>(lsm + 2 * dx_bytes + 3 * z_bytes)};
  float* rsc = reinterpret_cast<float*>(lsm + 2 * dx_bytes + 4 * z_bytes);   // [TR][4]: rstd*sc?, see below
  uint64_t* full = reinterpret_cast<uint64_t*>(rsc + TR * 4);                // [2]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (rows + TR - 1) / TR;
  const float invC = 1.0f / (float)C;
  const int nchunks = C / 8;
  const int ngrp = 256 / nchunks > 0 ? 256 / nchunks : 1;    // pass-2 thread groups (each covers all channels)
  const int grp = tid / nchunks, chunk = tid - grp * nchunks;
  const bool p2_active = grp < ngrp && nchunks <= 256;
  const int c2 = chunk * 8;

  if (tid == 0) {
    lnb_mbar_init(&full[0], 1);
    lnb_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int tile, int stage) {
    const int r0 = tile * TR;
    const int nr = min(TR, rows - r0);
    const uint32_t b1 = (uint32_t)((size_t)nr * C * 4), b2 = (uint32_t)((size_t)nr * C * sizeof(T));
    lnb_mbar_expect(&full[stage], b1 + b2);
    bulk_load(dxs[stage], dx + (size_t)r0 * C, b1, &full[stage]);
    bulk_load(zs[stage], z + (size_t)r0 * C, b2, &full[stage]);
  };

  float gm2[8], a_g[8], a_b[8], a_z[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { gm2[e] = 0.f; a_g[e] = 0.f; a_b[e] = 0.f; a_z[e] = 0.f; }
  if (p2_active) ld8(gamma + c2, gm2);
  // pass-1 ownership: lane holds channels lane*8 + 256*j (j < 4) of every row its warp visits
  float gm1[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int e = 0; e < 8; ++e) gm1[j][e] = 0.f;
    if (lane * 8 + j * 256 < C) ld8(gamma + lane * 8 + j * 256, gm1[j]);
  }

  int it = 0;
  if (tid == 0 && (int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const int r0 = tile * TR;
    const int nr = min(TR, rows - r0);
    if (tid == 0) {
      const int nxt = tile + gridDim.x;
      if (nxt < ntiles) issue(nxt, stage ^ 1);     // that stage was fully consumed before the previous iteration's last barrier
    }
    const float* dxt = dxs[stage];
    const T* zt = zs[stage];
    T* ot = outs[stage];
    lnb_mbar_wait(&full[stage], (uint32_t)((it >> 1) & 1));
    // ---- pass 1: row statistics (warp per row): A = sum dx*gamma, Bq = sum dx*gamma*z -------------------------------
    for (int rr = warp; rr < nr; rr += 8) {
      const int row = r0 + rr;
      float A = 0.f, Bq = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane * 8 + j * 256;
        if (c < C) {
          float d8[8], z8[8];
          ld8(dxt + (size_t)rr * C + c, d8);
          ld8(zt + (size_t)rr * C + c, z8);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float g = d8[e] * gm1[j][e];
            A += g;
            Bq = fmaf(g, z8[e], Bq);
          }
        }
      }
      A = warp_sum(A);
      Bq = warp_sum(Bq);
      if (lane == 0) {
        const float mean = __ldg(stats + 2 * (size_t)row), rstd = __ldg(stats + 2 * (size_t)row + 1);
        const float sc = sample_scale ? __ldg(sample_scale + row / rows_per_sample) : 1.0f;
        // mean_c(g) and mean_c(g * xhat) with g = dx*sc*gamma, xhat = (z - mean) * rstd
        rsc[rr * 4 + 0] = -mean * rstd;                       // xhat = z * rstd + this
        rsc[rr * 4 + 1] = rstd;
        rsc[rr * 4 + 2] = sc * A * invC;
        rsc[rr * 4 + 3] = sc * rstd * (Bq - mean * A) * invC;
      }
    }
    // the output stage written two tiles ago must have been read out by the copy engine
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    // ---- pass 2: dz + per-channel sums (thread per 8 channels) ------------------------------------------------------
    if (p2_active) {
      for (int rr = grp; rr < nr; rr += ngrp) {
        const float4 st = *reinterpret_cast<const float4*>(rsc + rr * 4);
        const float sc = sample_scale ? __ldg(sample_scale + (r0 + rr) / rows_per_sample) : 1.0f;
        float d8[8], z8[8], o8[8];
        ld8(dxt + (size_t)rr * C + c2, d8);
        ld8(zt + (size_t)rr * C + c2, z8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float du = d8[e] * sc;
          const float xh = fmaf(z8[e], st.y, st.x);
          o8[e] = st.y * (fmaf(du, gm2[e], -st.z) - xh * st.w);
          a_b[e] += du;
          a_g[e] = fmaf(du, xh, a_g[e]);
          a_z[e] += o8[e];
        }
        st8(ot + (size_t)rr * C + c2, o8);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      bulk_store(dz + (size_t)r0 * C, ot, (uint32_t)((size_t)nr * C * sizeof(T)));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (p2_active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(dgamma + c2 + e, a_g[e]);
      atomicAdd(dbeta + c2 + e, a_b[e]);
      if (dbias_prev) atomicAdd(dbias_prev + c2 + e, a_z[e]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// column sums (bias gradients)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int rows, int cols,
                                                     int ld) {
  __shared__ float red[8][256 + 8];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + tx) * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c < cols) {
    // four independent 16-byte loads in flight per thread: the pass is pure HBM streaming.  Rows are visited from the END of
    // the tensor: the kernel that produced x wrote its last rows last, so that part is still in L2.
    const int step = gridDim.y * 8;
    const T* xe = x + (size_t)(rows - 1) * ld + c;
    int r = blockIdx.y * 8 + ty;
    for (; r + 3 * step < rows; r += 4 * step) {
      float v0[8], v1[8], v2[8], v3[8];
      ld8(xe - (size_t)r * ld, v0);
      ld8(xe - (size_t)(r + step) * ld, v1);
      ld8(xe - (size_t)(r + 2 * step) * ld, v2);
      ld8(xe - (size_t)(r + 3 * step) * ld, v3);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += (v0[e] + v1[e]) + (v2[e] + v3[e]);
    }
    for (; r < rows; r += step) {
      float v[8];
      ld8(xe - (size_t)r * ld, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[ty][tx * 8 + e] = acc[e];
  __syncthreads();
  const int col = threadIdx.x;  // 256 columns per block
  float s = 0.f;
#pragma unroll
  for (int y = 0; y < 8; ++y) s += red[y][col];
  const int gc = blockIdx.x * 256 + col;
  if (gc < cols) atomicAdd(out + gc, s);
}

// ------------------------------------------------------------------------------------------------
// q/k L2 normalisation (in place), LPV lanes per (token, head) vector
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) qk_normalize_kernel(T* __restrict__ qkv, float* __restrict__ inv_norm, int Tn, int C,
                                                           int heads, int lpv, int cpl) {
  const int d = C / heads;
  const int nvec_per_tok = 2 * heads;
  const size_t nvec = (size_t)Tn * nvec_per_tok;
  const int groups_per_block = blockDim.x / lpv;
  const int gl = threadIdx.x / lpv, li = threadIdx.x % lpv;
  for (size_t g0 = (size_t)blockIdx.x * groups_per_block; g0 < nvec; g0 += (size_t)gridDim.x * groups_per_block) {
    // every thread of the block runs the same trip count so the shuffles below stay warp-converged
    const size_t g = g0 + gl;
    const bool valid = g < nvec;
    const size_t tok = valid ? g / nvec_per_tok : 0;
    const int v = valid ? (int)(g % nvec_per_tok) : 0;
    T* p = qkv + tok * 3 * (size_t)C + (size_t)v * d;
    float x[6][8];
    float ss = 0.f;
    for (int j = 0; j < cpl; ++j) {
      if (valid) {
        ld8(p + (li + j * lpv) * 8, x[j]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += x[j][e] * x[j][e];
      }
    }
    for (int o = lpv >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (valid) {
      for (int j = 0; j < cpl; ++j) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[j][e] *= inv;
        st8(p + (li + j * lpv) * 8, x[j]);
      }
      if (li == 0) inv_norm[g] = inv;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// shifted-window mask, materialised as the reference's (nW, L, L) buffer
// ------------------------------------------------------------------------------------------------
__global__ void shift_mask_kernel(float* __restrict__ mask, int H, int W, int Wh, int Ww, int s0) {
  const int L = Wh * Ww;
  const int nWw = W / Ww;
  const size_t total = (size_t)(H / Wh) * nWw * L * L;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i % L);
    const int n = (int)((i / L) % L);
    const int w = (int)(i / ((size_t)L * L));
    const int wh = w / nWw;
    const int ln = shift_region_label(wh * Wh + n / Ww, H, s0);
    const int lm = shift_region_label(wh * Wh + m / Ww, H, s0);
    mask[i] = (ln != lm) ? -100.0f : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// latitude-weighted L2 loss
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latw_l2_fwd_kernel(const float* __restrict__ prd, const float* __restrict__ tar,
                                                          const float* __restrict__ qw, float* __restrict__ num,
                                                          float* __restrict__ den, int H, int W, int rows_per_block) {
  __shared__ float rn[8], rd[8];
  const int plane = blockIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(H, r0 + rows_per_block);
  const int w4 = W / 4;
  const float4* p4 = reinterpret_cast<const float4*>(prd + (size_t)plane * H * W);
  const float4* t4 = reinterpret_cast<const float4*>(tar + (size_t)plane * H * W);
  float an = 0.f, ad = 0.f;
  const int n4 = (r1 - r0) * w4;
  const size_t base = (size_t)r0 * w4;
  for (int i = threadIdx.x; i < n4; i += blockDim.x * 4) {
    float4 p[4], t[4];
    float q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ii = i + u * blockDim.x;
      if (ii < n4) {
        p[u] = __ldg(p4 + base + ii);
        t[u] = __ldg(t4 + base + ii);
        q[u] = __ldg(qw + r0 + ii / w4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ii = i + u * blockDim.x;
      if (ii < n4) {
        const float dx = p[u].x - t[u].x, dy = p[u].y - t[u].y, dz = p[u].z - t[u].z, dw = p[u].w - t[u].w;
        an += q[u] * ((dx * dx + dy * dy) + (dz * dz + dw * dw));
        ad += q[u] * ((t[u].x * t[u].x + t[u].y * t[u].y) + (t[u].z * t[u].z + t[u].w * t[u].w));
      }
    }
  }
  an = warp_sum(an);
  ad = warp_sum(ad);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { rn[wid] = an; rd[wid] = ad; }
  __syncthreads();
  if (wid == 0) {
    an = lane < (int)(blockDim.x >> 5) ? rn[lane] : 0.f;
    ad = lane < (int)(blockDim.x >> 5) ? rd[lane] : 0.f;
    an = warp_sum(an);
    ad = warp_sum(ad);
    if (lane == 0) {
      atomicAdd(num + plane, an);
      atomicAdd(den + plane, ad);
    }
  }
}

__global__ void latw_l2_finish_kernel(const float* __restrict__ num, const float* __restrict__ den,
                                      const float* __restrict__ chw, int relative, int squared,
                                      float* __restrict__ loss, int BC, int C) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < BC; i += blockDim.x) {
    float v = relative ? num[i] / den[i] : num[i];
    if (!squared) v = sqrtf(v);
    acc += chw[i % C] * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) *loss = acc;
  }
}

__global__ void __launch_bounds__(256) latw_l2_bwd_kernel(const float* __restrict__ prd, const float* __restrict__ tar,
                                                          const float* __restrict__ qw, const float* __restrict__ chw,
                                                          const float* __restrict__ num, const float* __restrict__ den,
                                                          const float* __restrict__ gloss, int relative, int squared,
                                                          float* __restrict__ dprd, int C, int H, int W, int rows_per_block) {
  const int plane = blockIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(H, r0 + rows_per_block);
  const int w4 = W / 4;
  float coef = 2.0f * gloss[0] * chw[plane % C] / (relative ? den[plane] : 1.0f);
  if (!squared) coef *= 0.5f * rsqrtf(relative ? num[plane] / den[plane] : num[plane]);
  const float4* p4 = reinterpret_cast<const float4*>(prd + (size_t)plane * H * W);
  const float4* t4 = reinterpret_cast<const float4*>(tar + (size_t)plane * H * W);
  float4* d4 = reinterpret_cast<float4*>(dprd + (size_t)plane * H * W);
  const int n4 = (r1 - r0) * w4;
  const size_t base = (size_t)r0 * w4;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 p = __ldg(p4 + base + i), t = __ldg(t4 + base + i);
    const float s = coef * __ldg(qw + r0 + i / w4);
    d4[base + i] = make_float4(s * (p.x - t.x), s * (p.y - t.y), s * (p.z - t.z), s * (p.w - t.w));
  }
}

}  // namespace swinb200

// =================================================================================================
// C ABI
// =================================================================================================
using namespace swinb200;

extern "C" int swinb200_version(void) { return SWINB200_VERSION; }
extern "C" const char* swinb200_last_error(void) { return g_err; }

extern "C" int swinb200_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream) {
  SWB_CHECK_ARG(src && dst, "cast: null pointer");
  SWB_CHECK_ARG(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "cast: pointers must be 16-byte aligned");
  if (n == 0) return SWINB200_OK;
  const int blocks = (int)min((size_t)sm_count() * 8, (n / 8 + 255) / 256 + 1);
  cast_f32_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_pos_embed_grad(const float* dx, float* dpos, int B, int rows_per_sample, int C, void* stream) {
  SWB_CHECK_ARG(dx && dpos && B > 0 && rows_per_sample > 0 && C > 0, "pos_embed_grad: bad arguments");
  if (rows_per_sample % 4 == 0 && C % 4 == 0 && (uintptr_t)dx % 16 == 0 && (uintptr_t)dpos % 16 == 0) {
    dim3 grid((rows_per_sample + 63) / 64, (C + 63) / 64);
    transpose_sum_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dx, dpos, B, rows_per_sample, C);
  } else {
    dim3 grid((rows_per_sample + 31) / 32, (C + 31) / 32), block(32, 8);
    transpose_sum_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dx, dpos, B, rows_per_sample, C);
  }
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_transpose_f32(const float* src, float* dst, int R, int Cc, void* stream) {
  // dst (Cc, R) = src (R, Cc)^T
  return swinb200_pos_embed_grad(src, dst, 1, R, Cc, stream);
}

template <typename T>
static int launch_patchify(const float* img, T* out, int B, int C, int Hi, int Wi, int order, cudaStream_t s) {
  const int ntok = B * (Hi / 4) * (Wi / 4);
  const int K = C * 16;
  const size_t smem = (size_t)kPatchTok * (K + 8) * sizeof(T);
  const int blocks = (ntok + kPatchTok - 1) / kPatchTok;
  if (order == 0) {
    SWB_CUDA(cudaFuncSetAttribute(patchify_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patchify_kernel<T, 0><<<blocks, 256, smem, s>>>(img, out, B, C, Hi, Wi, ntok);
  } else {
    SWB_CUDA(cudaFuncSetAttribute(patchify_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patchify_kernel<T, 1><<<blocks, 256, smem, s>>>(img, out, B, C, Hi, Wi, ntok);
  }
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_patchify(const float* img, void* out, int act_dtype, int B, int C, int Hi, int Wi, int P, int order,
                                 void* stream) {
  SWB_CHECK_ARG(img && out, "patchify: null pointer");
  SWB_CHECK_ARG(P == 4, "patchify: only patch_size 4 is supported (got %d)", P);
  SWB_CHECK_ARG(B > 0 && C > 0 && Hi % 4 == 0 && Wi % 4 == 0, "patchify: bad shape B=%d C=%d Hi=%d Wi=%d", B, C, Hi, Wi);
  SWB_CHECK_ARG((C * 16) % 8 == 0 && (size_t)kPatchTok * (C * 16 + 8) * 4 <= 220 * 1024, "patchify: C=%d too large", C);
  SWB_CHECK_ARG(order == 0 || order == 1, "patchify: order must be 0 or 1");
  if (act_dtype == SWINB200_BF16) return launch_patchify<__nv_bfloat16>(img, (__nv_bfloat16*)out, B, C, Hi, Wi, order, (cudaStream_t)stream);
  if (act_dtype == SWINB200_F32) return launch_patchify<float>(img, (float*)out, B, C, Hi, Wi, order, (cudaStream_t)stream);
  SWB_CHECK_ARG(false, "patchify: bad act_dtype %d", act_dtype);
}

extern "C" int swinb200_patchify_cat(int n_src, const float* const* srcs, const int* chans, const long long* batch_strides, void* out,
                                     int act_dtype, int B, int Hi, int Wi, int P, void* stream) {
  return swinb200_patchify_cat_norm(n_src, srcs, chans, batch_strides, nullptr, nullptr, out, act_dtype, B, Hi, Wi, P, stream);
}

extern "C" int swinb200_patchify_cat_norm(int n_src, const float* const* srcs, const int* chans, const long long* batch_strides,
                                          const float* mean, const float* stdv, void* out, int act_dtype, int B, int Hi, int Wi, int P,
                                          void* stream) {
  SWB_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "patchify_cat: mean and std must be given together");
  SWB_CHECK_ARG(n_src >= 1 && n_src <= kPatchMaxSrc && srcs && chans && batch_strides && out, "patchify_cat: 1..%d sources", kPatchMaxSrc);
  SWB_CHECK_ARG(P == 4, "patchify_cat: only patch_size 4 is supported (got %d)", P);
  SWB_CHECK_ARG(B > 0 && Hi % 4 == 0 && Wi % 4 == 0, "patchify_cat: bad shape B=%d Hi=%d Wi=%d", B, Hi, Wi);
  PatchSrcs ps;
  int C = 0;
  for (int s = 0; s < n_src; ++s) {
    SWB_CHECK_ARG(srcs[s] && chans[s] > 0 && batch_strides[s] >= 0 && ((uintptr_t)srcs[s] % 16 == 0) && batch_strides[s] % 4 == 0,
                  "patchify_cat: source %d is null, empty or not 16-byte aligned", s);
    ps.ptr[s] = srcs[s]; ps.bstride[s] = batch_strides[s]; ps.c0[s] = C;
    C += chans[s];
  }
  ps.c0[n_src] = C;
  ps.n = n_src;
  SWB_CHECK_ARG((size_t)kPatchTok * (C * 16 + 8) * 4 <= 220 * 1024, "patchify_cat: C=%d too large", C);
  const int ntok = B * (Hi / 4) * (Wi / 4);
  const int blocks = (ntok + kPatchTok - 1) / kPatchTok;
  if (act_dtype == SWINB200_BF16) {
    const size_t smem = (size_t)kPatchTok * (C * 16 + 8) * 2;
    SWB_CUDA(cudaFuncSetAttribute(patchify_cat_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patchify_cat_kernel<__nv_bfloat16><<<blocks, 256, smem, (cudaStream_t)stream>>>(ps, mean, stdv, (__nv_bfloat16*)out, B, C, Hi, Wi, ntok);
  } else if (act_dtype == SWINB200_F32) {
    const size_t smem = (size_t)kPatchTok * (C * 16 + 8) * 4;
    SWB_CUDA(cudaFuncSetAttribute(patchify_cat_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patchify_cat_kernel<float><<<blocks, 256, smem, (cudaStream_t)stream>>>(ps, mean, stdv, (float*)out, B, C, Hi, Wi, ntok);
  } else {
    SWB_CHECK_ARG(false, "patchify_cat: bad act_dtype %d", act_dtype);
  }
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

template <typename T>
static int launch_unpatchify(const T* y, const float* skip, int skip_chans, const float* smean, const float* sstd, float* out, int B,
                             int Co, int Hi, int Wi, int order, cudaStream_t s) {
  const int ntok = B * (Hi / 4) * (Wi / 4);
  const int K = Co * 16;
  const size_t smem = (size_t)kPatchTok * (K + 8) * sizeof(T);
  const int blocks = (ntok + kPatchTok - 1) / kPatchTok;
  if (order == 1) {
    SWB_CUDA(cudaFuncSetAttribute(unpatchify_kernel<T, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unpatchify_kernel<T, 1><<<blocks, 256, smem, s>>>(y, skip, skip_chans, smean, sstd, out, B, Co, Hi, Wi, ntok);
  } else {
    SWB_CUDA(cudaFuncSetAttribute(unpatchify_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unpatchify_kernel<T, 0><<<blocks, 256, smem, s>>>(y, skip, skip_chans, smean, sstd, out, B, Co, Hi, Wi, ntok);
  }
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_unpatchify(const void* y, int act_dtype, const float* skip, int skip_chans, float* out, int B, int Co,
                                   int Hi, int Wi, int P, int order, void* stream) {
  return swinb200_unpatchify_norm(y, act_dtype, skip, skip_chans, nullptr, nullptr, out, B, Co, Hi, Wi, P, order, stream);
}

extern "C" int swinb200_unpatchify_norm(const void* y, int act_dtype, const float* skip, int skip_chans, const float* skip_mean,
                                        const float* skip_std, float* out, int B, int Co, int Hi, int Wi, int P, int order,
                                        void* stream) {
  SWB_CHECK_ARG((skip_mean == nullptr) == (skip_std == nullptr), "unpatchify: skip mean and std must be given together");
  SWB_CHECK_ARG(skip_mean == nullptr || skip != nullptr, "unpatchify: skip statistics without a skip tensor");
  SWB_CHECK_ARG(order == 0 || order == 1, "unpatchify: order must be 0 or 1");
  SWB_CHECK_ARG(y && out, "unpatchify: null pointer");
  SWB_CHECK_ARG(P == 4, "unpatchify: only patch_size 4 is supported (got %d)", P);
  SWB_CHECK_ARG(B > 0 && Co > 0 && Hi % 4 == 0 && Wi % 4 == 0, "unpatchify: bad shape");
  SWB_CHECK_ARG(skip == nullptr || skip_chans >= Co, "unpatchify: skip has fewer channels (%d) than the output (%d)", skip_chans, Co);
  SWB_CHECK_ARG((size_t)kPatchTok * (Co * 16 + 8) * 4 <= 220 * 1024, "unpatchify: Co=%d too large", Co);
  if (act_dtype == SWINB200_BF16) return launch_unpatchify<__nv_bfloat16>((const __nv_bfloat16*)y, skip, skip_chans, skip_mean, skip_std, out, B, Co, Hi, Wi, order, (cudaStream_t)stream);
  if (act_dtype == SWINB200_F32) return launch_unpatchify<float>((const float*)y, skip, skip_chans, skip_mean, skip_std, out, B, Co, Hi, Wi, order, (cudaStream_t)stream);
  SWB_CHECK_ARG(false, "unpatchify: bad act_dtype %d", act_dtype);
}

template <typename T>
static int launch_ln_fwd(const T* z, const float* x_in, const float* gamma, const float* beta, const float* ss,
                         const float* pos, float* x_out, T* xb, float* stats, int rows, int C, int rps, float eps,
                         cudaStream_t s) {
  const int blocks = min((rows + 7) / 8, sm_count() * 8);
  const int nchunk = (C + 255) / 256;
#define SWB_LN_FWD(N) ln_residual_fwd_kernel<T, N><<<blocks, 256, 0, s>>>(z, x_in, gamma, beta, ss, pos, x_out, xb, stats, rows, C, rps, eps)
  switch (nchunk) {
    case 1: SWB_LN_FWD(1); break;
    case 2: SWB_LN_FWD(2); break;
    case 3: SWB_LN_FWD(3); break;
    case 4: SWB_LN_FWD(4); break;
    default: set_error("ln_residual_fwd: C=%d > 1024 unsupported", C); return SWINB200_ERR_UNSUPPORTED;
  }
#undef SWB_LN_FWD
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_ln_residual_fwd(const void* z, int act_dtype, const float* x_in, const float* gamma, const float* beta,
                                        const float* sample_scale, const float* pos, float* x_out, void* xb_out, float* stats,
                                        int rows, int C, int rows_per_sample, float eps, void* stream) {
  SWB_CHECK_ARG(z && gamma && beta && x_out && xb_out && stats, "ln_residual_fwd: null pointer");
  SWB_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && rows_per_sample > 0, "ln_residual_fwd: bad shape rows=%d C=%d", rows, C);
  if (act_dtype == SWINB200_BF16)
    return launch_ln_fwd<__nv_bfloat16>((const __nv_bfloat16*)z, x_in, gamma, beta, sample_scale, pos, x_out, (__nv_bfloat16*)xb_out, stats, rows, C, rows_per_sample, eps, (cudaStream_t)stream);
  if (act_dtype == SWINB200_F32)
    return launch_ln_fwd<float>((const float*)z, x_in, gamma, beta, sample_scale, pos, x_out, (float*)xb_out, stats, rows, C, rows_per_sample, eps, (cudaStream_t)stream);
  SWB_CHECK_ARG(false, "ln_residual_fwd: bad act_dtype %d", act_dtype);
}

// ---- bf16 fast path for C = 128*NK channels (the model's C = 768) ---------------------------------------------------
// One producer warp keeps a 3-stage ring of 10-row tiles (dx fp32 + z bf16, two cp.async.bulk per tile) in flight; each of
// the 10 consumer warps owns one row of every tile and does the whole row from registers: lane l holds channels
// {l*4 + 128*k + e} (conflict-free 16-byte / 8-byte shared loads), two warp reductions give the row means, dz goes to the
// warp's own staging row and leaves with a 1.5 KB cp.async.bulk store.  No CTA-wide barrier in the loop; the ring slot is
// released as soon as the row sits in registers.  Per-channel sums (dgamma, dbeta, bias gradient of the previous Linear)
// stay in registers, are folded through shared memory once per CTA and reach HBM as one atomic per channel per CTA.
constexpr int kLnbRows = 10, kLnbStages = 3, kLnbThreads = 32 * (kLnbRows + 1);
__device__ __forceinline__ void lnb_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(bar)) : "memory");
}
template <int NK>
struct LnbSmem {
  static constexpr int C = 128 * NK;
  static constexpr int kDx = kLnbRows * C * 4, kZ = kLnbRows * C * 2, kStage = kDx + kZ;
  static constexpr int kOffOut = kLnbStages * kStage;                 // [warp][2][C] bf16
  static constexpr int kOffRed = kOffOut + kLnbRows * 2 * C * 2;      // [3][C] fp32 sums
  static constexpr int kOffGamma = kOffRed + 3 * C * 4;               // [C] gamma
  static constexpr int kOffBar = kOffGamma + C * 4;
  static constexpr int kBytes = kOffBar + 64;
};

template <int NK>
__global__ void __launch_bounds__(kLnbThreads, 1)
ln_bwd_rowwarp_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ z, const float* __restrict__ stats,
                      const float* __restrict__ gamma, const float* __restrict__ sample_scale, __nv_bfloat16* __restrict__ dz,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias_prev, int rows,
                      int rows_per_sample) {
  using SM = LnbSmem<NK>;
  constexpr int C = SM::C;
  extern __shared__ __align__(128) unsigned char lsm[];
  float* red = reinterpret_cast<float*>(lsm + SM::kOffRed);
  uint64_t* full = reinterpret_cast<uint64_t*>(lsm + SM::kOffBar);   // [stages]
  uint64_t* empty = full + kLnbStages;                                // [stages]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntiles = (rows + kLnbRows - 1) / kLnbRows;

  if (tid == 0) {
    for (int s = 0; s < kLnbStages; ++s) {
      lnb_mbar_init(&full[s], 1);
      lnb_mbar_init(&empty[s], kLnbRows);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 3 * C; i += kLnbThreads) red[i] = 0.f;
  float* sgamma = reinterpret_cast<float*>(lsm + SM::kOffGamma);      // read per row from smem: 24 registers per lane stay free
  for (int i = tid; i < C; i += kLnbThreads) sgamma[i] = __ldg(gamma + i);
  __syncthreads();

  if (warp == kLnbRows) {
    // ------------------------------------------------ producer ------------------------------------------------
    if (lane == 0) {
      int it = 0;
      for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
        const int tile = ntiles - 1 - ti;          // from the end: the tail of dx (just written by the dgrad GEMM) is still in L2
        const int stage = it % kLnbStages;
        if (it >= kLnbStages) lnb_mbar_wait(&empty[stage], (uint32_t)((it / kLnbStages - 1) & 1));
        const int r0 = tile * kLnbRows;
        const int nr = min(kLnbRows, rows - r0);
        unsigned char* st = lsm + stage * SM::kStage;
        lnb_mbar_expect(&full[stage], (uint32_t)(nr * C * 6));
        bulk_load(st, dx + (size_t)r0 * C, (uint32_t)(nr * C * 4), &full[stage]);
        bulk_load(st + SM::kDx, z + (size_t)r0 * C, (uint32_t)(nr * C * 2), &full[stage]);
      }
    }
  } else {
    // ------------------------------------------------ consumers: warp w <-> row w of every tile ----------------------
    float a_g[NK][4], a_b[NK][4], a_z[NK][4];
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) { a_g[k][e] = 0.f; a_b[k][e] = 0.f; a_z[k][e] = 0.f; }
    const float4* gm4 = reinterpret_cast<const float4*>(sgamma);
    unsigned char* obuf = lsm + SM::kOffOut + warp * 2 * C * 2;
    const float invC = 1.0f / (float)C;
    int it = 0, nstores = 0;
    for (int ti = blockIdx.x; ti < ntiles; ti += gridDim.x, ++it) {
      const int tile = ntiles - 1 - ti;
      const int stage = it % kLnbStages;
      const int row = tile * kLnbRows + warp;
      const bool active = row < rows;
      float mean = 0.f, rstd = 0.f, sc = 1.f;
      if (active) {     // issued before the wait: the latency hides behind the tile's arrival
        mean = __ldg(stats + 2 * (size_t)row);
        rstd = __ldg(stats + 2 * (size_t)row + 1);
        if (sample_scale) sc = __ldg(sample_scale + row / rows_per_sample);
      }
      lnb_mbar_wait(&full[stage], (uint32_t)((it / kLnbStages) & 1));
      float d[NK][4];
      uint2 zp[NK];                                  // z stays packed (4 x bf16) until it is used
      if (active) {
        const unsigned char* st = lsm + stage * SM::kStage;
        const float4* dxr = reinterpret_cast<const float4*>(st + (size_t)warp * C * 4);
        const uint2* zr = reinterpret_cast<const uint2*>(st + SM::kDx + (size_t)warp * C * 2);
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          const float4 d4 = dxr[lane + 32 * k];
          zp[k] = zr[lane + 32 * k];
          d[k][0] = d4.x; d[k][1] = d4.y; d[k][2] = d4.z; d[k][3] = d4.w;
        }
      }
      // the row is in registers: the slot may be refilled.  The refill is an async-proxy write after these generic-proxy
      // reads, hence the proxy fence ahead of the release.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) lnb_mbar_arrive(&empty[stage]);
      if (!active) continue;
      float A = 0.f, Bq = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const float4 g4 = gm4[lane + 32 * k];
        const float gmk[4] = {g4.x, g4.y, g4.z, g4.w};
        const float zk[4] = {__uint_as_float(zp[k].x << 16), __uint_as_float(zp[k].x & 0xffff0000u),
                             __uint_as_float(zp[k].y << 16), __uint_as_float(zp[k].y & 0xffff0000u)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g = d[k][e] * gmk[e];
          A += g;
          Bq = fmaf(g, zk[e], Bq);
        }
      }
      A = warp_sum(A);
      Bq = warp_sum(Bq);
      const float xo = -mean * rstd;                          // xhat = z * rstd + xo
      const float mg = sc * A * invC;                         // mean_c(g),  g = dx*sc*gamma
      const float mgx = sc * rstd * (Bq - mean * A) * invC;   // mean_c(g * xhat)
      unsigned char* ob = obuf + (nstores & 1) * C * 2;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store issued two rows ago has drained
      __syncwarp();
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        float o4[4];
        const float4 g4 = gm4[lane + 32 * k];
        const float gmk[4] = {g4.x, g4.y, g4.z, g4.w};
        const float zk[4] = {__uint_as_float(zp[k].x << 16), __uint_as_float(zp[k].x & 0xffff0000u),
                             __uint_as_float(zp[k].y << 16), __uint_as_float(zp[k].y & 0xffff0000u)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float du = d[k][e] * sc;
          const float xh = fmaf(zk[e], rstd, xo);
          o4[e] = rstd * (fmaf(du, gmk[e], -mg) - xh * mgx);
          a_b[k][e] += du;
          a_g[k][e] = fmaf(du, xh, a_g[k][e]);
          a_z[k][e] += o4[e];
        }
        uint2 pk;
        pk.x = pack_bf16x2(o4[0], o4[1]);
        pk.y = pack_bf16x2(o4[2], o4[3]);
        reinterpret_cast<uint2*>(ob)[lane + 32 * k] = pk;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        bulk_store(dz + (size_t)row * C, ob, (uint32_t)(C * 2));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      ++nstores;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = lane * 4 + 128 * k + e;
        atomicAdd(red + c, a_g[k][e]);
        atomicAdd(red + C + c, a_b[k][e]);
        atomicAdd(red + 2 * C + c, a_z[k][e]);
      }
  }
  __syncthreads();
  for (int c = tid; c < C; c += kLnbThreads) {
    atomicAdd(dgamma + c, red[c]);
    atomicAdd(dbeta + c, red[C + c]);
    if (dbias_prev) atomicAdd(dbias_prev + c, red[2 * C + c]);
  }
}

template <int NK>
static int launch_ln_bwd_rowwarp(const float* dx, const __nv_bfloat16* z, const float* stats, const float* gamma, const float* ss,
                                 __nv_bfloat16* dz, float* dgamma, float* dbeta, float* dbias_prev, int rows, int rps,
                                 cudaStream_t s) {
  using SM = LnbSmem<NK>;
  static bool configured = false;
  if (!configured) {
    SWB_CUDA(cudaFuncSetAttribute(ln_bwd_rowwarp_kernel<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
    configured = true;
  }
  const int ntiles = (rows + kLnbRows - 1) / kLnbRows;
  ln_bwd_rowwarp_kernel<NK><<<max(1, min(ntiles, sm_count())), kLnbThreads, SM::kBytes, s>>>(dx, z, stats, gamma, ss, dz, dgamma,
                                                                                           dbeta, dbias_prev, rows, rps);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

template <typename T>
static int launch_ln_bwd(const float* dx, const T* z, const float* stats, const float* gamma, const float* ss, T* dz,
                         float* dgamma, float* dbeta, float* dbias_prev, int rows, int C, int rps, cudaStream_t s) {
  if constexpr (sizeof(T) == 2) {
    const bool aligned = ((uintptr_t)dx % 16 == 0) && ((uintptr_t)z % 16 == 0) && ((uintptr_t)dz % 16 == 0) && ((uintptr_t)gamma % 16 == 0);
    if (C == 768 && aligned) return launch_ln_bwd_rowwarp<6>(dx, z, stats, gamma, ss, dz, dgamma, dbeta, dbias_prev, rows, rps, s);
  }
  if (C / 8 > 256) {
    set_error("ln_residual_bwd: C=%d > 2048 unsupported", C);
    return SWINB200_ERR_UNSUPPORTED;
  }
  // rows per tile: two stages of (dx fp32 + z) plus two output stages within ~100 KB (two CTAs per SM), at most 8
  const size_t per_row = (size_t)C * (2 * 4 + 4 * sizeof(T));
  int TR = (int)min((size_t)8, (size_t)(100 * 1024) / per_row);
  if (TR < 1) {
    set_error("ln_residual_bwd: C=%d too large for the shared-memory tile", C);
    return SWINB200_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)TR * per_row + (size_t)TR * 16 + 64;
  static bool configured = false;
  if (!configured) {
    SWB_CUDA(cudaFuncSetAttribute(ln_residual_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const int ntiles = (rows + TR - 1) / TR;
  const int blocks = max(1, min(ntiles, 2 * sm_count()));
  ln_residual_bwd_kernel<T><<<blocks, 256, smem, s>>>(dx, z, stats, gamma, ss, dz, dgamma, dbeta, dbias_prev, rows, C, rps, TR);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_ln_residual_bwd(const float* dx, const void* z, int act_dtype, const float* stats, const float* gamma,
                                        const float* sample_scale, void* dz, float* dgamma, float* dbeta, float* dbias_prev,
                                        int rows, int C, int rows_per_sample, void* stream) {
  SWB_CHECK_ARG(dx && z && stats && gamma && dz && dgamma && dbeta, "ln_residual_bwd: null pointer");
  SWB_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && rows_per_sample > 0, "ln_residual_bwd: bad shape rows=%d C=%d", rows, C);
  if (act_dtype == SWINB200_BF16)
    return launch_ln_bwd<__nv_bfloat16>(dx, (const __nv_bfloat16*)z, stats, gamma, sample_scale, (__nv_bfloat16*)dz, dgamma, dbeta, dbias_prev, rows, C, rows_per_sample, (cudaStream_t)stream);
  if (act_dtype == SWINB200_F32)
    return launch_ln_bwd<float>(dx, (const float*)z, stats, gamma, sample_scale, (float*)dz, dgamma, dbeta, dbias_prev, rows, C, rows_per_sample, (cudaStream_t)stream);
  SWB_CHECK_ARG(false, "ln_residual_bwd: bad act_dtype %d", act_dtype);
}

extern "C" int swinb200_colsum(const void* x, int act_dtype, float* out, int rows, int cols, int ld, void* stream) {
  SWB_CHECK_ARG(x && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0, "colsum: bad arguments rows=%d cols=%d ld=%d", rows, cols, ld);
  const int gx = (cols + 255) / 256;
  int gy = max(1, min((rows + 63) / 64, (sm_count() * 4 + gx - 1) / gx));
  dim3 grid(gx, gy);
  if (act_dtype == SWINB200_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, out, rows, cols, ld);
  else if (act_dtype == SWINB200_F32)
    colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, out, rows, cols, ld);
  else
    SWB_CHECK_ARG(false, "colsum: bad act_dtype %d", act_dtype);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_qk_normalize(void* qkv, int act_dtype, float* inv_norm, int T, int C, int heads, void* stream) {
  SWB_CHECK_ARG(qkv && inv_norm && T > 0 && heads > 0 && C % heads == 0, "qk_normalize: bad arguments");
  const int d = C / heads;
  SWB_CHECK_ARG(d % 8 == 0, "qk_normalize: head_dim %d must be a multiple of 8", d);
  const int chunks = d / 8;
  int lpv = 1;
  while (lpv < 8 && chunks % (lpv * 2) == 0) lpv *= 2;
  const int cpl = chunks / lpv;
  SWB_CHECK_ARG(cpl <= 6, "qk_normalize: head_dim %d unsupported (chunks per lane %d > 6)", d, cpl);
  const size_t nvec = (size_t)T * 2 * heads;
  const int gpb = 256 / lpv;
  const int blocks = (int)min((nvec + gpb - 1) / gpb, (size_t)sm_count() * 16);
  if (act_dtype == SWINB200_BF16)
    qk_normalize_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)qkv, inv_norm, T, C, heads, lpv, cpl);
  else if (act_dtype == SWINB200_F32)
    qk_normalize_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)qkv, inv_norm, T, C, heads, lpv, cpl);
  else
    SWB_CHECK_ARG(false, "qk_normalize: bad act_dtype %d", act_dtype);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_shift_mask(float* mask, int H, int W, int Wh, int Ww, int s0, int s1, void* stream) {
  SWB_CHECK_ARG(mask && H > 0 && W > 0 && Wh > 0 && Ww > 0 && H % Wh == 0 && W % Ww == 0, "shift_mask: bad geometry");
  SWB_CHECK_ARG(s0 > 0 || s1 > 0, "shift_mask: un-shifted blocks have no mask");
  shift_mask_kernel<<<sm_count() * 4, 256, 0, (cudaStream_t)stream>>>(mask, H, W, Wh, Ww, s0);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_latw_l2_fwd(const float* prd, const float* tar, const float* qw, const float* chw, int relative,
                                    int squared, float* num, float* den, float* loss, int B, int C, int H, int W,
                                    void* stream) {
  SWB_CHECK_ARG(prd && tar && qw && chw && num && den && loss, "latw_l2_fwd: null pointer");
  SWB_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "latw_l2_fwd: bad shape (W must be a multiple of 4)");
  cudaStream_t s = (cudaStream_t)stream;
  SWB_CUDA(cudaMemsetAsync(num, 0, sizeof(float) * B * C, s));
  SWB_CUDA(cudaMemsetAsync(den, 0, sizeof(float) * B * C, s));
  const int planes = B * C;
  int split = max(1, (sm_count() * 8 + planes - 1) / planes);
  int rpb = max(1, (H + split - 1) / split);
  split = (H + rpb - 1) / rpb;
  latw_l2_fwd_kernel<<<dim3(planes, split), 256, 0, s>>>(prd, tar, qw, num, den, H, W, rpb);
  SWB_LAUNCH_CHECK();
  latw_l2_finish_kernel<<<1, 256, 0, s>>>(num, den, chw, relative, squared, loss, planes, C);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

extern "C" int swinb200_latw_l2_bwd(const float* prd, const float* tar, const float* qw, const float* chw, const float* num,
                                    const float* den, const float* gloss, int relative, int squared, float* dprd, int B, int C,
                                    int H, int W, void* stream) {
  SWB_CHECK_ARG(prd && tar && qw && chw && gloss && dprd && (den || !relative) && (num || squared), "latw_l2_bwd: null pointer");
  SWB_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && W % 4 == 0, "latw_l2_bwd: bad shape");
  const int planes = B * C;
  int split = max(1, (sm_count() * 8 + planes - 1) / planes);
  int rpb = max(1, (H + split - 1) / split);
  split = (H + rpb - 1) / rpb;
  latw_l2_bwd_kernel<<<dim3(planes, split), 256, 0, (cudaStream_t)stream>>>(prd, tar, qw, chw, num, den, gloss, relative, squared, dprd, C, H, W, rpb);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}
