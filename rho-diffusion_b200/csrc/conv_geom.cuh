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
  // A pointwise conv (every kernel extent 1, no stride, no upsample) sees the positions of a sample as one flat row-major list, so an
  // extent beyond the tensor-core engines' 32767-coordinate limit folds into the next dimension: the attention projections of the
  // 128^3 config are Conv1d(k = 1) over T = 128 x 16 x 16 = 32768 tokens (unet_v2.py:296-304) and would otherwise fall to the CUDA cores.
  if (p->kd * p->kh * p->kw == 1 && p->sd * p->sh * p->sw == 1 && p->ud * p->uh * p->uw == 1) {
    while (g.W > 32767 && g.W % 2 == 0 && g.H <= 16383) { g.W /= 2; g.H *= 2; }
    while (g.H > 32767 && g.H % 2 == 0 && g.D <= 16383) { g.H /= 2; g.D *= 2; }
  }
  g.ud = p->ud; g.uh = p->uh; g.uw = p->uw;
  g.Du = g.D * p->ud; g.Hu = g.H * p->uh; g.Wu = g.W * p->uw;
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

struct RowPos {
  int n, d, h, w;
  bool valid;
};

__device__ __forceinline__ RowPos decode_out(const ConvGeom& g, long long m) {
  RowPos r;
  r.valid = m < g.M;
  long long t = r.valid ? m : 0;
  r.w = (int)(t % g.Wo); t /= g.Wo;
  r.h = (int)(t % g.Ho); t /= g.Ho;
  r.d = (int)(t % g.Do); t /= g.Do;
  r.n = (int)t;
  return r;
}
__device__ __forceinline__ RowPos decode_in(const ConvGeom& g, long long m) {
  RowPos r;
  r.valid = m < g.Min;
  long long t = r.valid ? m : 0;
  r.w = (int)(t % g.W); t /= g.W;
  r.h = (int)(t % g.H); t /= g.H;
  r.d = (int)(t % g.D); t /= g.D;
  r.n = (int)t;
  return r;
}

// forward gather: input position (pre-upsample) feeding output `r` through tap (kd,kh,kw); -1 if padding
__device__ __forceinline__ long long fwd_src(const ConvGeom& g, const RowPos& r, int tap) {
  int kw = tap % g.kw, kh = (tap / g.kw) % g.kh, kd = tap / (g.kw * g.kh);
  int pd = r.d * g.sd + kd - g.pd, ph = r.h * g.sh + kh - g.ph, pw = r.w * g.sw + kw - g.pw;
  if (!r.valid || pd < 0 || pd >= g.Du || ph < 0 || ph >= g.Hu || pw < 0 || pw >= g.Wu) return -1;
  int id = pd / g.ud, ih = ph / g.uh, iw = pw / g.uw;
  return (((long long)r.n * g.D + id) * g.H + ih) * g.W + iw;
}

// dgrad gather: output position whose tap `tap` reads upsampled child `child` of input position `r`; -1 if none
__device__ __forceinline__ long long dgrad_src(const ConvGeom& g, const RowPos& r, int vt) {
  int tap = vt % g.taps, child = vt / g.taps;
  int kw = tap % g.kw, kh = (tap / g.kw) % g.kh, kd = tap / (g.kw * g.kh);
  int cw = child % g.uw, ch = (child / g.uw) % g.uh, cd = child / (g.uw * g.uh);
  int td = r.d * g.ud + cd + g.pd - kd, th = r.h * g.uh + ch + g.ph - kh, tw = r.w * g.uw + cw + g.pw - kw;
  if (!r.valid || td < 0 || th < 0 || tw < 0) return -1;
  if ((td % g.sd) | (th % g.sh) | (tw % g.sw)) return -1;
  int od = td / g.sd, oh = th / g.sh, ow = tw / g.sw;
  if (od >= g.Do || oh >= g.Ho || ow >= g.Wo) return -1;
  return (((long long)r.n * g.Do + od) * g.Ho + oh) * g.Wo + ow;
}

}  // namespace rho
