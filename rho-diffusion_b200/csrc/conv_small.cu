// conv_nd for the two degenerate layers of every UNet: the first conv (Cin = in_channels = 1..4 -> model_channels,
// rho_diffusion/models/unet_v2.py:535) and the last one (model_channels -> out_channels = 1..4, unet_v2.py:682).
// K = taps*Cin (resp. N = Cout) is far too small for a GEMM tile: both are HBM-bound and are written as direct
// convolutions whose only job is to stream the wide tensor once with coalesced 16-byte accesses.
//   small-Cin : 8 threads per output position, each produces 8 consecutive output channels (16-byte bf16 stores,
//               a warp writes 4 x 128 contiguous bytes); the taps*Cin*Cout weights live in shared memory.
//   small-Cout: one warp per output position; lanes stride over (tap, channel-pair) so that every tap is one
//               coalesced row read; warp-shuffle reduction.
#include "conv_geom.cuh"

namespace rho {

constexpr int SC_MAX_W = 12288;  // floats of weights staged in shared memory (48 KB)

template <typename TA, typename TW, typename TO>
__global__ void __launch_bounds__(256) conv_small_cin_kernel(ConvGeom g, const TA* __restrict__ x, const TW* __restrict__ w,
                                                             const float* __restrict__ bias, TO* __restrict__ y) {
  __shared__ __align__(16) float ws[SC_MAX_W];   // [tap*Cin + ci][Cout]
  const int K = g.taps * g.Cin;
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) {
    const int co = i / K, k = i % K;
    ws[k * g.Cout + co] = to_f32(w[i]);
  }
  __syncthreads();
  const int groups = g.Cout >> 3;                      // 8-channel groups per position
  const long long total = g.M * groups;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long m = idx / groups;
    const int c0 = (int)(idx % groups) * 8;
    const RowPos rp = decode_out(g, m);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[c0 + j] : 0.f;
    int t = 0;
    for (int kd = 0; kd < g.kd; ++kd) {
      const int pd = rp.d * g.sd + kd - g.pd;
      const bool okd = pd >= 0 && pd < g.Du;
      const long long bd = ((long long)rp.n * g.D + pd / g.ud) * g.H;
      for (int kh = 0; kh < g.kh; ++kh) {
        const int ph = rp.h * g.sh + kh - g.ph;
        const bool okh = okd && ph >= 0 && ph < g.Hu;
        const long long bh = (bd + ph / g.uh) * g.W;
        for (int kw = 0; kw < g.kw; ++kw, ++t) {
          const int pw = rp.w * g.sw + kw - g.pw;
          if (!(okh && pw >= 0 && pw < g.Wu)) continue;
          const long long src = bh + pw / g.uw;
          for (int ci = 0; ci < g.Cin; ++ci) {
            const float xv = to_f32(x[src * g.Cin + ci]);
            const float4 w0 = *reinterpret_cast<const float4*>(&ws[(t * g.Cin + ci) * g.Cout + c0]);
            const float4 w1 = *reinterpret_cast<const float4*>(&ws[(t * g.Cin + ci) * g.Cout + c0 + 4]);
            acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
            acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
          }
        }
      }
    }
    const float lo[4] = {acc[0], acc[1], acc[2], acc[3]}, hi[4] = {acc[4], acc[5], acc[6], acc[7]};
    store4(y + m * g.Cout + c0, lo);
    store4(y + m * g.Cout + c0 + 4, hi);
  }
}

template <typename TA, typename TW, typename TO>
__global__ void __launch_bounds__(256) conv_small_cout_kernel(ConvGeom g, const TA* __restrict__ x, const TW* __restrict__ w,
                                                              const float* __restrict__ bias, TO* __restrict__ y) {
  __shared__ float ws[SC_MAX_W];   // [co][tap*Cin + ci]
  const int K = g.taps * g.Cin;
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) ws[i] = to_f32(w[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int half = g.Cin >> 1;                 // channel pairs per tap
  const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < g.M; m += warps_total) {
    const RowPos rp = decode_out(g, m);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int t = 0;
    for (int kd = 0; kd < g.kd; ++kd) {
      const int pd = rp.d * g.sd + kd - g.pd;
      const bool okd = pd >= 0 && pd < g.Du;
      const long long bd = ((long long)rp.n * g.D + pd / g.ud) * g.H;
      for (int kh = 0; kh < g.kh; ++kh) {
        const int ph = rp.h * g.sh + kh - g.ph;
        const bool okh = okd && ph >= 0 && ph < g.Hu;
        const long long bh = (bd + ph / g.uh) * g.W;
        for (int kw = 0; kw < g.kw; ++kw, ++t) {
          const int pw = rp.w * g.sw + kw - g.pw;
          if (!(okh && pw >= 0 && pw < g.Wu)) continue;       // warp-uniform
          const TA* row = x + (bh + pw / g.uw) * g.Cin;
          for (int cp = lane; cp < half; cp += 32) {          // one coalesced row read per tap
            const float x0 = to_f32(row[2 * cp]), x1 = to_f32(row[2 * cp + 1]);
            const int k = t * g.Cin + 2 * cp;
#pragma unroll
            for (int co = 0; co < 4; ++co)
              if (co < g.Cout) acc[co] = fmaf(x0, ws[co * K + k], fmaf(x1, ws[co * K + k + 1], acc[co]));
          }
        }
      }
    }
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      if (co < g.Cout) {
        const float s = warp_sum(acc[co]);
        if (lane == 0) y[m * g.Cout + co] = from_f32<TO>(s + (bias ? bias[co] : 0.f));
      }
    }
  }
}

bool conv_small_supported(const ConvGeom& g, const rho_conv_params* p) {
  if (p->residual || g.C2 != 0) return false;
  if ((long long)g.taps * g.Cin * g.Cout > SC_MAX_W) return false;
  if (g.Cin <= 4 && g.Cout % 8 == 0) return true;
  if (g.Cout <= 4 && g.Cin % 2 == 0) return true;
  return false;
}

template <typename TA, typename TW, typename TO>
static int launch_small(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  if (g.Cin <= 4 && g.Cout % 8 == 0) {
    long long total = g.M * (g.Cout >> 3);
    int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 16);
    conv_small_cin_kernel<TA, TW, TO><<<blocks, 256, 0, st>>>(g, (const TA*)p->x, (const TW*)p->w_packed, p->bias, (TO*)p->y);
    RHO_CHECK_LAUNCH("conv_small_cin");
  } else {
    int blocks = (int)std::min<long long>(ceil_div(g.M * 32, 256), (long long)sm_count() * 16);
    conv_small_cout_kernel<TA, TW, TO><<<blocks, 256, 0, st>>>(g, (const TA*)p->x, (const TW*)p->w_packed, p->bias, (TO*)p->y);
    RHO_CHECK_LAUNCH("conv_small_cout");
  }
  return RHO_OK;
}

int conv_small_fwd(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  const int a = p->act_dtype, w = p->w_dtype, f = p->y_f32;
  if (a == RHO_F32 && w == RHO_F32) return launch_small<float, float, float>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16 && !f) return launch_small<bf16, bf16, bf16>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16 && f) return launch_small<bf16, bf16, float>(g, p, st);
  return fail(RHO_ERR_UNSUPPORTED, "conv_small: dtype combination");
}

}  // namespace rho
