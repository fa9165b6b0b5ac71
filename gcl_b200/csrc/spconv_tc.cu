// K3 tensor-core path placeholder: filled in by the tcgen05 kind::tf32 kernel.
#include "common.cuh"

namespace gclb {
struct ConvParams;
bool spconv_tc_supported(const ConvParams&) { return false; }
int spconv_fwd_tc(const ConvParams&, int64_t, cudaStream_t) { return GCLB_ERR_UNSUPPORTED; }
bool nn_tc_supported(int) { return false; }
int nn_tc(const float*, const float*, int, const int64_t*, const int64_t*, int, int64_t, int64_t, unsigned long long*,
          unsigned long long*, cudaStream_t) {
  return GCLB_ERR_UNSUPPORTED;
}
}  // namespace gclb

extern "C" int gclb_has_tcgen05(void) { return 0; }
