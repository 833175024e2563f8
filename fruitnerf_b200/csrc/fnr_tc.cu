// Dispatch of the tensor-core render / export forward (fnr_render_forward, fnr_export_forward with impl = tcgen05 / auto):
//   fruit_nerf family (geo 15, semantic 15-64-64, colour 63-64-64-3)           -> fnr_tc_ws.cu  (warp-specialised: cp.async gather
//                                                                                 warps, TMEM-resident MLP chains, per-group compositing)
//   fruit_nerf_big / _huge (geo 30, semantic 30-128-128-64, colour 78-64-64-3)  -> fnr_tc_big.cu
// everything else is served by the simt kernels (tc_supported() == false).
// (The round-1 kernel of this file -- gather and MLP chain in the same 16 warps, activations in shared memory -- was replaced by
// fnr_tc_ws.cu in round 2: 0.529 -> 0.284 ms on the 4096 x 192 bench batch.)
#include "fnr_common.cuh"
#include "fnr_kernels.h"

namespace fnr {

bool tc_supported(Family fam, const KField& F, const KRays& Rr) {
  (void)F;
  if (fam == kFamilyBig) return tc_big_supported(Rr.S);
  return fam == kFamilySmall && tc_ws_supported(Rr.S);
}

bool tc_export_supported(Family fam, const KExport& E) {
  if (fam == kFamilyBig) return tc_big_supported(E.S);
  return fam == kFamilySmall && tc_ws_supported(E.S);
}

int launch_tc_render_forward(Family fam, const KField& F, const KParams& P, const KRays& Rr, const KFieldOut& O,
                             const KComposite& Cm, cudaStream_t st) {
  if (!tc_supported(fam, F, Rr)) {
    set_error("tcgen05 render kernel does not support this shape");
    return FNR_ERR_UNSUPPORTED;
  }
  if (Rr.R == 0) return FNR_OK;
  if (fam == kFamilyBig) return launch_tc_render_forward_big(F, P, Rr, O, Cm, st);
  return launch_tc_render_forward_ws(F, P, Rr, O, Cm, st);
}

// The export path reuses the fused forward (AABB positions, mean appearance embedding) and replaces the
// compositing stage by the three threshold selections + stream compaction.
int launch_tc_export(Family fam, const KField& F, const KParams& P, const KExport& E, cudaStream_t st) {
  if (!tc_export_supported(fam, E)) {
    set_error("tcgen05 export kernel does not support this shape");
    return FNR_ERR_UNSUPPORTED;
  }
  if (E.B == 0) return FNR_OK;
  if (fam == kFamilyBig) return launch_tc_export_big(F, P, E, st);
  return launch_tc_export_ws(F, P, E, st);
}

}  // namespace fnr
