// C-ABI entry points of conv_nd and engine selection (tcgen05 implicit GEMM vs CUDA-core fallback).
#include "conv_geom.cuh"

namespace rho {
// conv_simt.cu
int conv_fwd_simt(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st);
int conv_dgrad_simt(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st);
int conv_wgrad_simt(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st);
size_t conv_wgrad_simt_workspace(const ConvGeom& g);
// conv_small.cu
bool conv_small_supported(const ConvGeom& g, const rho_conv_params* p);
int conv_small_fwd(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st);
bool conv_small_dgrad_supported(const ConvGeom& g, const rho_conv_bwd_params* p);
int conv_small_dgrad(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st);
bool conv_small_wgrad_supported(const ConvGeom& g, const rho_conv_bwd_params* p);
size_t conv_small_wgrad_workspace(const ConvGeom& g, const rho_conv_bwd_params* p);
int conv_small_wgrad(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st);
// conv_tc.cu
bool conv_fwd_tc_supported(const ConvGeom& g, const rho_conv_params* p);
int conv_fwd_tc_stats_rows(const ConvGeom& g, const rho_conv_params* p);
int conv_fwd_tc(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st);
bool conv_dgrad_tc_supported(const ConvGeom& g, const rho_conv_bwd_params* p);
int conv_dgrad_tc(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st);
size_t conv_dgrad_tc_workspace(const ConvGeom& g, const rho_conv_bwd_params* p);
// conv_wgrad_tc.cu
bool conv_wgrad_tc_supported(const ConvGeom& g, const rho_conv_bwd_params* p);
size_t conv_wgrad_tc_workspace(const ConvGeom& g, const rho_conv_bwd_params* p);
int conv_wgrad_tc(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st);
size_t conv_aux_weight_bytes(const ConvGeom& g, const rho_conv_params* p);
int conv_pack_weight_aux(const ConvGeom& g, const rho_conv_params* p, const float* w, void* out, cudaStream_t st);

static int validate_fwd(const rho_conv_params* p, ConvGeom& g) {
  int rc = make_geom(p, g);
  if (rc) return rc;
  RHO_REQUIRE(p->x && p->w_packed && p->y, "conv_fwd: null tensor");
  RHO_REQUIRE((p->C2 == 0) == (p->x2 == nullptr), "conv_fwd: x2/C2 mismatch");
  RHO_REQUIRE(g.M > 0, "conv_fwd: empty output");
  return RHO_OK;
}
}  // namespace rho

using namespace rho;

extern "C" {

int rho_conv_engine_for(const rho_conv_params* p) {
  ConvGeom g;
  if (make_geom(p, g)) return RHO_ERR_INVALID_ARG;
  if (p->engine == RHO_CONV_SIMT) return RHO_CONV_SIMT;
  return conv_fwd_tc_supported(g, p) ? RHO_CONV_TCGEN05 : RHO_CONV_SIMT;
}

size_t rho_conv_aux_weight_bytes(const rho_conv_params* p) {
  ConvGeom g;
  if (make_geom(p, g)) return 0;
  return conv_aux_weight_bytes(g, p);
}

int rho_conv_pack_weight_aux(const rho_conv_params* p, const float* w, void* aux, rho_stream_t stream) {
  ConvGeom g;
  int rc = make_geom(p, g);
  if (rc) return rc;
  RHO_REQUIRE(w && aux, "conv_pack_weight_aux: null pointer");
  return conv_pack_weight_aux(g, p, w, aux, as_stream(stream));
}

int32_t rho_conv_stats_rows(const rho_conv_params* p) {
  ConvGeom g;
  if (!p || make_geom(p, g)) return 0;
  if (p->engine == RHO_CONV_SIMT) return 0;
  return conv_fwd_tc_stats_rows(g, p);
}

size_t rho_conv_workspace_bytes(const rho_conv_params* p) {
  (void)p;
  return 0;
}

int rho_conv_fwd(const rho_conv_params* p, rho_stream_t stream) {
  ConvGeom g;
  int rc = validate_fwd(p, g);
  if (rc) return rc;
  const bool tc_ok = conv_fwd_tc_supported(g, p);
  if (p->engine == RHO_CONV_TCGEN05 && !tc_ok) return fail(RHO_ERR_UNSUPPORTED, "conv_fwd: tcgen05 engine does not take this shape");
  if (p->engine != RHO_CONV_SIMT && tc_ok) return conv_fwd_tc(g, p, as_stream(stream));
  // degenerate first / last layers (Cin <= 4 or Cout <= 4): HBM-bound direct convolution
  if (p->engine == RHO_CONV_AUTO && p->act_dtype == p->w_dtype && conv_small_supported(g, p)) return conv_small_fwd(g, p, as_stream(stream));
  return conv_fwd_simt(g, p, as_stream(stream));
}

size_t rho_conv_bwd_workspace_bytes(const rho_conv_bwd_params* p) {
  ConvGeom g;
  if (!p || make_geom(&p->fwd, g)) return 0;
  size_t a = conv_wgrad_simt_workspace(g);
  size_t b = conv_dgrad_tc_workspace(g, p);
  size_t c = conv_wgrad_tc_workspace(g, p);
  size_t d = conv_small_wgrad_workspace(g, p);
  a = a > b ? a : b;
  a = a > c ? a : c;
  return a > d ? a : d;
}

int rho_conv_dgrad(const rho_conv_bwd_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p, "conv_dgrad: null params");
  ConvGeom g;
  int rc = make_geom(&p->fwd, g);
  if (rc) return rc;
  RHO_REQUIRE(p->dy && p->dx && p->fwd.w_packed, "conv_dgrad: null tensor");
  RHO_REQUIRE((g.C2 == 0) || p->dx2, "conv_dgrad: dx2 required for a concat input");
  const bool tc_ok = conv_dgrad_tc_supported(g, p);
  if (p->fwd.engine == RHO_CONV_TCGEN05 && !tc_ok) return fail(RHO_ERR_UNSUPPORTED, "conv_dgrad: tcgen05 engine does not take this shape");
  if (p->fwd.engine != RHO_CONV_SIMT && tc_ok) return conv_dgrad_tc(g, p, as_stream(stream));
  if (p->fwd.engine == RHO_CONV_AUTO && conv_small_dgrad_supported(g, p)) return conv_small_dgrad(g, p, as_stream(stream));
  return conv_dgrad_simt(g, p, as_stream(stream));
}

int rho_conv_wgrad(const rho_conv_bwd_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p, "conv_wgrad: null params");
  ConvGeom g;
  int rc = make_geom(&p->fwd, g);
  if (rc) return rc;
  RHO_REQUIRE(p->dy, "conv_wgrad: null dy");
  RHO_REQUIRE(p->dw || p->dbias, "conv_wgrad: nothing to compute");
  if (p->dw && p->fwd.engine != RHO_CONV_SIMT && conv_wgrad_tc_supported(g, p)) {
    RHO_REQUIRE(p->fwd.x, "conv_wgrad: x required");
    rc = conv_wgrad_tc(g, p, as_stream(stream));
    if (rc || !p->dbias) return rc;
    rho_conv_bwd_params q = *p;     // bias gradient: column sums of dy (deterministic two-stage reduction)
    q.dw = nullptr;
    return conv_wgrad_simt(g, &q, as_stream(stream));
  }
  if (p->fwd.engine == RHO_CONV_TCGEN05 && p->dw) return fail(RHO_ERR_UNSUPPORTED, "conv_wgrad: tcgen05 engine does not take this shape");
  // degenerate first / last layers: HBM-bound streaming reduction
  if (p->dw && p->fwd.engine == RHO_CONV_AUTO && conv_small_wgrad_supported(g, p)) {
    RHO_REQUIRE(p->fwd.x, "conv_wgrad: x required");
    rc = conv_small_wgrad(g, p, as_stream(stream));
    if (rc || !p->dbias) return rc;
    rho_conv_bwd_params q = *p;
    q.dw = nullptr;
    return conv_wgrad_simt(g, &q, as_stream(stream));
  }
  return conv_wgrad_simt(g, p, as_stream(stream));
}

}  // extern "C"
