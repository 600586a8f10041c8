// Fused multi-tensor Adam step: one streaming pass per parameter updates the fp32 master weight and
// both moments AND refreshes the bf16 shadow the tensor-core GEMMs read, so the optimizer is pure HBM traffic
// (16 B read + 12 B written per element, + 2 B for a shadow) with no separate re-cast pass afterwards.
// Arithmetic follows torch.optim.Adam(fused=True) (train.py:175-176: lr, betas=(0.9, 0.95), eps 1e-8, no amsgrad):
//   g  = grad / grad_scale (+ weight_decay * p)
//   m  = m + (1 - beta1) * (g - m)              (std::lerp with weight < 0.5)
//   v  = beta2 * v + (1 - beta2) * g * g
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps),     bc_i = 1 - beta_i^step
// and honours GradScaler's found_inf (the whole step is skipped on device, no host sync).
#include "common.cuh"

namespace swinb200 {

constexpr int kAdamMaxTensors = 384;     // tensors per launch (the whole 162-tensor model is one launch; 20 KB of kernel parameters)
constexpr int kAdamChunk = 16384;        // elements per CTA
constexpr int kAdamThreads = 256;

struct AdamArgs {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  __nv_bfloat16* shadow[kAdamMaxTensors];   // nullable per tensor
  long long numel[kAdamMaxTensors];
  int first_chunk[kAdamMaxTensors + 1];     // prefix sums: CTA b works on tensor t with first_chunk[t] <= b < first_chunk[t+1]
  int n_tensors;
  float step_size, bc2_sqrt, beta1, beta2, eps, weight_decay;
  const float* grad_scale;   // device scalar or null
  const float* found_inf;    // device scalar or null
};

__device__ __forceinline__ float adam_update(float& p, float g, float& m, float& v, const AdamArgs& a, float inv_scale) {
  g *= inv_scale;
  if (a.weight_decay != 0.f) g = fmaf(p, a.weight_decay, g);
  m = m + (1.0f - a.beta1) * (g - m);
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - (a.step_size * m) / denom;
  return p;
}

__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const __grid_constant__ AdamArgs a) {
  if (a.found_inf != nullptr && *a.found_inf != 0.f) return;           // GradScaler: skip the step, device side
  const float inv_scale = a.grad_scale != nullptr ? 1.0f / *a.grad_scale : 1.0f;
  int lo = 0, hi = a.n_tensors;              // binary search of this CTA's tensor
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (a.first_chunk[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const int t = lo;
  const long long c0 = (long long)((int)blockIdx.x - a.first_chunk[t]) * kAdamChunk;
  const int n = (int)min((long long)kAdamChunk, a.numel[t] - c0);
  float* p = a.p[t] + c0;
  const float* g = a.g[t] + c0;
  float* m = a.m[t] + c0;
  float* v = a.v[t] + c0;
  __nv_bfloat16* sh = a.shadow[t] ? a.shadow[t] + c0 : nullptr;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0) && ((uintptr_t)sh % 8 == 0);
  if (vec) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += kAdamThreads) {
      float4 p4 = reinterpret_cast<float4*>(p)[i];
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
      float4 m4 = reinterpret_cast<float4*>(m)[i];
      float4 v4 = reinterpret_cast<float4*>(v)[i];
      adam_update(p4.x, g4.x, m4.x, v4.x, a, inv_scale);
      adam_update(p4.y, g4.y, m4.y, v4.y, a, inv_scale);
      adam_update(p4.z, g4.z, m4.z, v4.z, a, inv_scale);
      adam_update(p4.w, g4.w, m4.w, v4.w, a, inv_scale);
      reinterpret_cast<float4*>(p)[i] = p4;
      reinterpret_cast<float4*>(m)[i] = m4;
      reinterpret_cast<float4*>(v)[i] = v4;
      if (sh) {
        uint2 pk;
        pk.x = pack_bf16x2(p4.x, p4.y);
        pk.y = pack_bf16x2(p4.z, p4.w);
        reinterpret_cast<uint2*>(sh)[i] = pk;
      }
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += kAdamThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_update(pp, g[i], mm, vv, a, inv_scale);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (sh) sh[i] = __float2bfloat16(pp);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kAdamThreads) {
      float pp = p[i], mm = m[i], vv = v[i];
      adam_update(pp, g[i], mm, vv, a, inv_scale);
      p[i] = pp; m[i] = mm; v[i] = vv;
      if (sh) sh[i] = __float2bfloat16(pp);
    }
  }
}

}  // namespace swinb200

using namespace swinb200;

extern "C" int swinb200_adam_step(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                                  void* const* exp_avg_sq, void* const* shadows, const long long* numel, double lr, double beta1,
                                  double beta2, double eps, double weight_decay, long long step, const float* grad_scale,
                                  const float* found_inf, void* stream) {
  SWB_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || (params && grads && exp_avg && exp_avg_sq && numel)), "adam_step: null table");
  SWB_CHECK_ARG(step >= 1, "adam_step: step counts from 1 (got %lld)", step);
  SWB_CHECK_ARG(lr >= 0 && beta1 >= 0 && beta1 < 1 && beta2 >= 0 && beta2 < 1 && eps >= 0, "adam_step: bad hyper-parameters");
  AdamArgs a;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
  a.grad_scale = grad_scale; a.found_inf = found_inf;
  int nt = 0;
  long long nb = 0;
  a.first_chunk[0] = 0;
  auto flush = [&]() -> int {
    if (nb > 0) {
      a.n_tensors = nt;
      adam_multi_kernel<<<(unsigned)nb, kAdamThreads, 0, (cudaStream_t)stream>>>(a);
      SWB_LAUNCH_CHECK();
    }
    nt = 0; nb = 0;
    return SWINB200_OK;
  };
  for (int i = 0; i < n_tensors; ++i) {
    SWB_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] >= 0, "adam_step: tensor %d has a null pointer", i);
    if (numel[i] == 0) continue;
    const long long chunks = (numel[i] + kAdamChunk - 1) / kAdamChunk;
    if (nt == kAdamMaxTensors || nb + chunks > 0x7fffffffLL) { if (int e = flush()) return e; }
    a.p[nt] = (float*)params[i]; a.g[nt] = (const float*)grads[i]; a.m[nt] = (float*)exp_avg[i]; a.v[nt] = (float*)exp_avg_sq[i];
    a.shadow[nt] = shadows ? (__nv_bfloat16*)shadows[i] : nullptr;
    a.numel[nt] = numel[i];
    nb += chunks;
    a.first_chunk[++nt] = (int)nb;
  }
  return flush();
}
