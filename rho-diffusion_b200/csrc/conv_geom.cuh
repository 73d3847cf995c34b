// Geometry shared by the conv engines (SIMT and tcgen05).
#pragma once
#include "common.cuh"

namespace rho {

struct ConvGeom {
  int N, D, H, W;     // input dims before the optional nearest upsample
  int Du, Hu, Wu;     // after upsample
  int Do, Ho, Wo;     // output dims
  int C1, C2, Cin, Cout;
  int kd, kh, kw, taps;
  int sd, sh, sw;
  int ud, uh, uw;
  int pd, ph, pw;
  long long M;        // N*Do*Ho*Wo
  long long Min;      // N*D*H*W
  int Ktot;           // taps*Cin
};

inline int make_geom(const rho_conv_params* p, ConvGeom& g) {
  RHO_REQUIRE(p, "conv: null params");
  RHO_REQUIRE(p->N > 0 && p->D > 0 && p->H > 0 && p->W > 0 && p->C1 > 0 && p->C2 >= 0 && p->Cout > 0, "conv: bad dims");
  auto ok_k = [](int k) { return k == 1 || k == 3; };
  auto ok_s = [](int s) { return s == 1 || s == 2; };
  RHO_REQUIRE(ok_k(p->kd) && ok_k(p->kh) && ok_k(p->kw), "conv: kernel extents must be 1 or 3");
  RHO_REQUIRE(ok_s(p->sd) && ok_s(p->sh) && ok_s(p->sw), "conv: strides must be 1 or 2");
  RHO_REQUIRE(ok_s(p->ud) && ok_s(p->uh) && ok_s(p->uw), "conv: upsample factors must be 1 or 2");
  g.N = p->N; g.D = p->D; g.H = p->H; g.W = p->W;
  g.ud = p->ud; g.uh = p->uh; g.uw = p->uw;
  g.Du = p->D * p->ud; g.Hu = p->H * p->uh; g.Wu = p->W * p->uw;
  g.kd = p->kd; g.kh = p->kh; g.kw = p->kw; g.taps = p->kd * p->kh * p->kw;
  g.sd = p->sd; g.sh = p->sh; g.sw = p->sw;
  g.pd = p->kd / 2; g.ph = p->kh / 2; g.pw = p->kw / 2;
  g.Do = (g.Du + 2 * g.pd - g.kd) / g.sd + 1;
  g.Ho = (g.Hu + 2 * g.ph - g.kh) / g.sh + 1;
  g.Wo = (g.Wu + 2 * g.pw - g.kw) / g.sw + 1;
  g.C1 = p->C1; g.C2 = p->C2; g.Cin = p->C1 + p->C2; g.Cout = p->Cout;
  g.M = (long long)g.N * g.Do * g.Ho * g.Wo;
  g.Min = (long long)g.N * g.D * g.H * g.W;
  g.Ktot = g.taps * g.Cin;
  return RHO_OK;
}

}  // namespace rho
