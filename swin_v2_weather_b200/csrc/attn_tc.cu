// tcgen05 windowed cosine attention (placeholder until the kernel lands; reports UNSUPPORTED loudly).
#include "common.cuh"

namespace swinb200 {

int attn_tcgen05_fwd(const void*, const float*, const float*, void*, float*, int, int, int, int, int, int, int, int, int,
                     cudaStream_t) {
  set_error("window_attn_fwd: the tcgen05 back end is not built in this version");
  return SWINB200_ERR_UNSUPPORTED;
}
int attn_tcgen05_bwd(const void*, const float*, const float*, const float*, const void*, const void*, const float*, void*,
                     float*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t) {
  set_error("window_attn_bwd: the tcgen05 back end is not built in this version");
  return SWINB200_ERR_UNSUPPORTED;
}

}  // namespace swinb200
