// Fused tcgen05 render forward (placeholder until the kernel lands in this file).
#include "fnr_common.cuh"
#include "fnr_kernels.h"
namespace fnr {
bool tc_supported(Family, const KField&, const KRays&) { return false; }
int launch_tc_render_forward(Family, const KField&, const KParams&, const KRays&, const KFieldOut&, const KComposite&,
                             cudaStream_t) {
  set_error("tcgen05 render kernel not built");
  return FNR_ERR_UNSUPPORTED;
}
}  // namespace fnr
