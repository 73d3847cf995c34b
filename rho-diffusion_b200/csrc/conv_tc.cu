// placeholder replaced below by the tcgen05 engine
#include "conv_geom.cuh"
namespace rho {
bool conv_fwd_tc_supported(const ConvGeom&, const rho_conv_params*) { return false; }
int conv_fwd_tc(const ConvGeom&, const rho_conv_params*, cudaStream_t) { return fail(RHO_ERR_UNSUPPORTED, "tcgen05 engine not built"); }
bool conv_dgrad_tc_supported(const ConvGeom&, const rho_conv_bwd_params*) { return false; }
int conv_dgrad_tc(const ConvGeom&, const rho_conv_bwd_params*, cudaStream_t) { return fail(RHO_ERR_UNSUPPORTED, "tcgen05 engine not built"); }
size_t conv_dgrad_tc_workspace(const ConvGeom&, const rho_conv_bwd_params*) { return 0; }
}  // namespace rho
