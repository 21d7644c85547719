// placeholder until the fused tile kernels land
#include "vw_internal.cuh"
int vw_fused_forward(vw_ctx *, const VwFusedFwd &, const VwFilt &) { return VW_EUNSUPPORTED; }
int vw_fused_inverse(vw_ctx *, const VwFusedInv &, const VwFilt &) { return VW_EUNSUPPORTED; }
