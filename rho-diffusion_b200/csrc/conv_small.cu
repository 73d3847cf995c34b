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


// ---- fast paths (bf16 activations, stride 1, no upsample: the shapes every shipped config produces) ------------
// small-Cout: L lanes share one output position, lane v owns the 16-byte channel octets v, v+L, ...: every warp load
// instruction covers 32/L complete rows (full 128-byte lines when Cin = 64), every input row is fetched from L1/L2
// once per tap by exactly one lane group, and the L partial sums meet in a log2(L)-step shuffle at the end.
template <int L, typename TO>
__global__ void __launch_bounds__(256) conv_small_cout_vec_kernel(ConvGeom g, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const float* __restrict__ bias, TO* __restrict__ y) {
  extern __shared__ __align__(16) float ws_dyn[];   // [co][tap*Cin + ci]
  const int K = g.taps * g.Cin;
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) ws_dyn[i] = to_f32(w[i]);
  __syncthreads();
  const int v = threadIdx.x % L;
  const long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const bool valid = m < g.M;
  unsigned t = valid ? (unsigned)m : 0u;                      // host guarantees M < 2^31
  const int ow = (int)(t % (unsigned)g.Wo); t /= (unsigned)g.Wo;
  const int oh = (int)(t % (unsigned)g.Ho); t /= (unsigned)g.Ho;
  const int od = (int)(t % (unsigned)g.Do);
  const int n = (int)(t / (unsigned)g.Do);
  const int octets = g.Cin >> 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int tap = 0;
  for (int kd = 0; kd < g.kd; ++kd) {
    const int pd = od + kd - g.pd;
    const bool okd = valid && pd >= 0 && pd < g.D;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int ph = oh + kh - g.ph;
      const bool okh = okd && ph >= 0 && ph < g.H;
      const long long base = (((long long)n * g.D + pd) * g.H + ph) * g.W;
      for (int kw = 0; kw < g.kw; ++kw, ++tap) {
        const int pw = ow + kw - g.pw;
        if (!(okh && pw >= 0 && pw < g.W)) continue;
        const uint4* row = reinterpret_cast<const uint4*>(x + (base + pw) * g.Cin);
        for (int o = v; o < octets; o += L) {
          const uint4 q = __ldg(row + o);
          const uint32_t u[4] = {q.x, q.y, q.z, q.w};
          float xv[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) { xv[2 * j] = __uint_as_float(u[j] << 16); xv[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u); }
          const int k = tap * g.Cin + o * 8;
#pragma unroll
          for (int co = 0; co < 4; ++co) {
            if (co < g.Cout) {
              const float4 w0 = *reinterpret_cast<const float4*>(&ws_dyn[co * K + k]);
              const float4 w1 = *reinterpret_cast<const float4*>(&ws_dyn[co * K + k + 4]);
              float a = acc[co];
              a = fmaf(xv[0], w0.x, a); a = fmaf(xv[1], w0.y, a); a = fmaf(xv[2], w0.z, a); a = fmaf(xv[3], w0.w, a);
              a = fmaf(xv[4], w1.x, a); a = fmaf(xv[5], w1.y, a); a = fmaf(xv[6], w1.z, a); a = fmaf(xv[7], w1.w, a);
              acc[co] = a;
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    if (co < g.Cout) {
      float s = acc[co];
#pragma unroll
      for (int off = L >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (v == 0 && valid) y[m * g.Cout + co] = from_f32<TO>(s + (bias ? bias[co] : 0.f));
    }
  }
}

// small-Cin: 8 lanes share a run of PW = 4 consecutive output positions along W, each lane produces 8 output channels
// for all four (the (kw + j) input column is loaded once and feeds up to three taps; each weight vector is read from
// shared memory once per four positions).  A warp store instruction writes four complete 128-byte lines at Cout = 64.
// `flip`: the kernel computes the DATA gradient of a small-Cout conv (the roles swap: g.Cin = forward Cout, g.Cout = forward
// Cin, x = dY) and `w` is the forward packed weight [g.Cin][taps][g.Cout], read with the taps reversed.
template <int CIN>
__global__ void __launch_bounds__(256) conv_small_cin_row4_kernel(ConvGeom g, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const float* __restrict__ bias, bf16* __restrict__ y, int flip) {
  extern __shared__ __align__(16) float ws_dyn[];   // [tap*Cin + ci][Cout]
  const int K = g.taps * CIN;
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) {
    if (!flip) {
      const int co = i / K, k = i % K;
      ws_dyn[k * g.Cout + co] = to_f32(w[i]);
    } else {
      const int co = i % g.Cout, tap = (i / g.Cout) % g.taps, ci = i / (g.Cout * g.taps);   // w[ci][tap][co]
      ws_dyn[((g.taps - 1 - tap) * CIN + ci) * g.Cout + co] = to_f32(w[i]);
    }
  }
  __syncthreads();
  const int groups = g.Cout >> 3;
  const int wq = g.Wo >> 2;                                   // runs per output row
  const long long total = (g.M >> 2) * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % groups) * 8;
  unsigned t = (unsigned)(idx / groups);                      // run index; host guarantees M < 2^31
  const int ow0 = (int)(t % (unsigned)wq) * 4; t /= (unsigned)wq;
  const int oh = (int)(t % (unsigned)g.Ho); t /= (unsigned)g.Ho;
  const int od = (int)(t % (unsigned)g.Do);
  const int n = (int)(t / (unsigned)g.Do);
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float b = bias ? bias[c0 + j] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][j] = b;
  }
  for (int kd = 0; kd < g.kd; ++kd) {
    const int pd = od + kd - g.pd;
    if (pd < 0 || pd >= g.D) continue;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int ph = oh + kh - g.ph;
      if (ph < 0 || ph >= g.H) continue;
      const bf16* row = x + ((((long long)n * g.D + pd) * g.H + ph) * g.W) * CIN;
      const int tap0 = (kd * g.kh + kh) * g.kw;
      // input columns ow0 - pw .. ow0 + 3 + (kw - 1 - pw)
      for (int col = 0; col < 4 + g.kw - 1; ++col) {
        const int pw = ow0 + col - g.pw;
        if (pw < 0 || pw >= g.W) continue;
        float xv[CIN];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xv[ci] = to_f32(row[(long long)pw * CIN + ci]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kw = col - q;                             // tap that maps input column `col` to output q
          if (kw < 0 || kw >= g.kw) continue;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float* wp = &ws_dyn[((tap0 + kw) * CIN + ci) * g.Cout + c0];
            const float4 w0 = *reinterpret_cast<const float4*>(wp);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
            acc[q][0] = fmaf(xv[ci], w0.x, acc[q][0]); acc[q][1] = fmaf(xv[ci], w0.y, acc[q][1]);
            acc[q][2] = fmaf(xv[ci], w0.z, acc[q][2]); acc[q][3] = fmaf(xv[ci], w0.w, acc[q][3]);
            acc[q][4] = fmaf(xv[ci], w1.x, acc[q][4]); acc[q][5] = fmaf(xv[ci], w1.y, acc[q][5]);
            acc[q][6] = fmaf(xv[ci], w1.z, acc[q][6]); acc[q][7] = fmaf(xv[ci], w1.w, acc[q][7]);
          }
        }
      }
    }
  }
  const long long m0 = (((long long)n * g.Do + od) * g.Ho + oh) * g.Wo + ow0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(acc[q][2 * j], acc[q][2 * j + 1]);
    *reinterpret_cast<uint4*>(y + (m0 + q) * g.Cout + c0) = o;
  }
}


// ---- run kernels (kw = 3, pad 1, unit stride): every thread owns a RUN of consecutive output positions along W so that
// an input column loaded once feeds the three kw taps, and the weights of one (kd, kh) filter row sit in registers.
// small-Cout: lane v of an L-lane group owns channel octet v (Cin = 8 L) for a run of 4 positions: per filter row
// 6 x 16-byte loads and 96 FMAs per output channel; the L partial sums meet in log2(L) shuffles.
template <int L, typename TO>
__global__ void __launch_bounds__(256) conv_small_cout_run4_kernel(ConvGeom g, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                   const float* __restrict__ bias, TO* __restrict__ y) {
  extern __shared__ __align__(16) float ws_dyn[];   // [co][tap*Cin + ci]
  const int K = g.taps * g.Cin;
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) ws_dyn[i] = to_f32(w[i]);
  __syncthreads();
  const int v = threadIdx.x % L;
  const long long run = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / L;
  const bool valid = run < (g.M >> 2);
  unsigned t = valid ? (unsigned)run : 0u;
  const unsigned wq = (unsigned)g.Wo >> 2;
  const int ow0 = (int)(t % wq) * 4; t /= wq;
  const int oh = (int)(t % (unsigned)g.Ho); t /= (unsigned)g.Ho;
  const int od = (int)(t % (unsigned)g.Do);
  const int n = (int)(t / (unsigned)g.Do);
  float acc[4][4];
#pragma unroll
  for (int co = 0; co < 4; ++co)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[co][q] = 0.f;
  for (int kd = 0; kd < g.kd; ++kd) {
    const int pd = od + kd - g.pd;
    if (pd < 0 || pd >= g.D) continue;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int ph = oh + kh - g.ph;
      if (ph < 0 || ph >= g.H) continue;
      const bf16* row = x + ((((long long)n * g.D + pd) * g.H + ph) * g.W) * g.Cin + v * 8;
      uint4 raw[6];
#pragma unroll
      for (int col = 0; col < 6; ++col) {
        const int pw = ow0 + col - 1;
        raw[col] = (pw >= 0 && pw < g.W) ? __ldg(reinterpret_cast<const uint4*>(row + (long long)pw * g.Cin)) : make_uint4(0u, 0u, 0u, 0u);
      }
      float xv[6][8];
#pragma unroll
      for (int col = 0; col < 6; ++col) {
        const uint32_t u[4] = {raw[col].x, raw[col].y, raw[col].z, raw[col].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { xv[col][2 * j] = __uint_as_float(u[j] << 16); xv[col][2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u); }
      }
      const int k0 = ((kd * g.kh + kh) * 3) * g.Cin + v * 8;
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        if (co < g.Cout) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 w0 = *reinterpret_cast<const float4*>(&ws_dyn[co * K + k0 + kw * g.Cin]);
            const float4 w1 = *reinterpret_cast<const float4*>(&ws_dyn[co * K + k0 + kw * g.Cin + 4]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float* xx = xv[q + kw];
              float a = acc[co][q];
              a = fmaf(xx[0], w0.x, a); a = fmaf(xx[1], w0.y, a); a = fmaf(xx[2], w0.z, a); a = fmaf(xx[3], w0.w, a);
              a = fmaf(xx[4], w1.x, a); a = fmaf(xx[5], w1.y, a); a = fmaf(xx[6], w1.z, a); a = fmaf(xx[7], w1.w, a);
              acc[co][q] = a;
            }
          }
        }
      }
    }
  }
  const long long m0 = run * 4;
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    if (co < g.Cout) {
      const float b = bias ? bias[co] : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float s = acc[co][q];
#pragma unroll
        for (int off = L >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (v == 0 && valid) y[(m0 + q) * g.Cout + co] = from_f32<TO>(s + b);
      }
    }
  }
}

// small-Cin: a thread owns 8 output channels for a run of 8 positions; per filter row the 10 input columns and, per input
// channel, the 3 x 8 weights are in registers: 6 shared-memory loads per 192 FMAs.
template <int CIN>
__global__ void __launch_bounds__(256) conv_small_cin_row8_kernel(ConvGeom g, const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                  const float* __restrict__ bias, bf16* __restrict__ y, int flip) {
  extern __shared__ __align__(16) float ws_dyn[];   // [tap*Cin + ci][Cout]
  const int K = g.taps * CIN;
  for (int i = threadIdx.x; i < K * g.Cout; i += blockDim.x) {
    if (!flip) {
      const int co = i / K, k = i % K;
      ws_dyn[k * g.Cout + co] = to_f32(w[i]);
    } else {
      const int co = i % g.Cout, tap = (i / g.Cout) % g.taps, ci = i / (g.Cout * g.taps);   // w[ci][tap][co]
      ws_dyn[((g.taps - 1 - tap) * CIN + ci) * g.Cout + co] = to_f32(w[i]);
    }
  }
  __syncthreads();
  const int groups = g.Cout >> 3;
  const int wq = g.Wo >> 3;
  const long long total = (g.M >> 3) * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % groups) * 8;
  unsigned t = (unsigned)(idx / groups);
  const int ow0 = (int)(t % (unsigned)wq) * 8; t /= (unsigned)wq;
  const int oh = (int)(t % (unsigned)g.Ho); t /= (unsigned)g.Ho;
  const int od = (int)(t % (unsigned)g.Do);
  const int n = (int)(t / (unsigned)g.Do);
  float acc[8][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float b = bias ? bias[c0 + j] : 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q][j] = b;
  }
  for (int kd = 0; kd < g.kd; ++kd) {
    const int pd = od + kd - g.pd;
    if (pd < 0 || pd >= g.D) continue;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int ph = oh + kh - g.ph;
      if (ph < 0 || ph >= g.H) continue;
      const bf16* row = x + ((((long long)n * g.D + pd) * g.H + ph) * g.W) * CIN;
      float xv[10][CIN];
#pragma unroll
      for (int col = 0; col < 10; ++col) {
        const int pw = ow0 + col - 1;
        const bool ok = pw >= 0 && pw < g.W;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) xv[col][ci] = ok ? to_f32(row[(long long)pw * CIN + ci]) : 0.f;
      }
      const int tap0 = (kd * g.kh + kh) * 3;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float wr[3][8];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* wp = &ws_dyn[((tap0 + kw) * CIN + ci) * g.Cout + c0];
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
          wr[kw][0] = w0.x; wr[kw][1] = w0.y; wr[kw][2] = w0.z; wr[kw][3] = w0.w;
          wr[kw][4] = w1.x; wr[kw][5] = w1.y; wr[kw][6] = w1.z; wr[kw][7] = w1.w;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float xx = xv[q + kw][ci];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(xx, wr[kw][j], acc[q][j]);
          }
      }
    }
  }
  const long long m0 = (((long long)n * g.Do + od) * g.Ho + oh) * g.Wo + ow0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(acc[q][2 * j], acc[q][2 * j + 1]);
    *reinterpret_cast<uint4*>(y + (m0 + q) * g.Cout + c0) = o;
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

// bf16 fast paths: unit stride, no fused upsample, fewer than 2^31 output positions
static bool small_fast_geom(const ConvGeom& g) {
  return g.sd == 1 && g.sh == 1 && g.sw == 1 && g.ud == 1 && g.uh == 1 && g.uw == 1 && g.M < (1ll << 31);
}
static bool run_geom(const ConvGeom& g) { return g.kw == 3 && g.pw == 1 && (g.kh == 1 || g.ph == g.kh / 2) && (g.kd == 1 || g.pd == g.kd / 2); }
template <typename TO>
static int launch_small_cout_vec(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  const int octets = g.Cin >> 3;
  if (run_geom(g) && g.Wo % 4 == 0 && (octets == 4 || octets == 8 || octets == 16 || octets == 32)) {
    const size_t smem_r = (size_t)g.taps * g.Cin * g.Cout * sizeof(float);
    const unsigned blocks_r = (unsigned)ceil_div((g.M >> 2) * octets, 256);
#define RHO_SCRUN(LL) conv_small_cout_run4_kernel<LL, TO><<<blocks_r, 256, smem_r, st>>>(g, (const bf16*)p->x, (const bf16*)p->w_packed, p->bias, (TO*)p->y)
    if (octets == 4) RHO_SCRUN(4);
    else if (octets == 8) RHO_SCRUN(8);
    else if (octets == 16) RHO_SCRUN(16);
    else RHO_SCRUN(32);
#undef RHO_SCRUN
    RHO_CHECK_LAUNCH("conv_small_cout_run4");
    return RHO_OK;
  }
  const int L = octets % 8 == 0 ? 8 : octets % 4 == 0 ? 4 : octets % 2 == 0 ? 2 : 1;
  const size_t smem = (size_t)g.taps * g.Cin * g.Cout * sizeof(float);
  const unsigned blocks = (unsigned)ceil_div(g.M * L, 256);
#define RHO_SCV(LL) conv_small_cout_vec_kernel<LL, TO><<<blocks, 256, smem, st>>>(g, (const bf16*)p->x, (const bf16*)p->w_packed, p->bias, (TO*)p->y)
  if (L == 8) RHO_SCV(8);
  else if (L == 4) RHO_SCV(4);
  else if (L == 2) RHO_SCV(2);
  else RHO_SCV(1);
#undef RHO_SCV
  RHO_CHECK_LAUNCH("conv_small_cout_vec");
  return RHO_OK;
}
static int launch_small_cin_row4(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  const size_t smem = (size_t)g.taps * g.Cin * g.Cout * sizeof(float);
  if (run_geom(g) && g.Wo % 8 == 0) {
    const unsigned blocks8 = (unsigned)ceil_div((g.M >> 3) * (g.Cout >> 3), 256);
#define RHO_SCR8(CI) conv_small_cin_row8_kernel<CI><<<blocks8, 256, smem, st>>>(g, (const bf16*)p->x, (const bf16*)p->w_packed, p->bias, (bf16*)p->y, 0)
    switch (g.Cin) {
      case 1: RHO_SCR8(1); break;
      case 2: RHO_SCR8(2); break;
      case 3: RHO_SCR8(3); break;
      default: RHO_SCR8(4); break;
    }
#undef RHO_SCR8
    RHO_CHECK_LAUNCH("conv_small_cin_row8");
    return RHO_OK;
  }
  const unsigned blocks = (unsigned)ceil_div((g.M >> 2) * (g.Cout >> 3), 256);
#define RHO_SCR(CI) conv_small_cin_row4_kernel<CI><<<blocks, 256, smem, st>>>(g, (const bf16*)p->x, (const bf16*)p->w_packed, p->bias, (bf16*)p->y, 0)
  switch (g.Cin) {
    case 1: RHO_SCR(1); break;
    case 2: RHO_SCR(2); break;
    case 3: RHO_SCR(3); break;
    default: RHO_SCR(4); break;
  }
#undef RHO_SCR
  RHO_CHECK_LAUNCH("conv_small_cin_row4");
  return RHO_OK;
}


// ---- weight gradient of the two degenerate layers (bf16, unit stride) ------------------------------------------------
//   first conv (Cin <= 4):  dW[co][tap][ci] = sum_pos dY[pos][co] * x[pos + off(tap)][ci]     wide = dY, narrow = x,  SIGN = +1
//   last  conv (Cout <= 4): dW[co][tap][ci] = sum_q   x[q][ci]   * dY[q - off(tap)][co]       wide = x,  narrow = dY, SIGN = -1
// Both stream the wide [M x CW] tensor once per (kd, narrow channel) with 16-byte loads and keep the KH*KW x 8 partial
// sums of a thread in registers; the position lanes of a CTA meet in shuffles + one shared-memory pass, the CTAs in a
// fixed-order second kernel (deterministic, no atomics).
template <int SIGN, int KH, int KW>
__global__ void __launch_bounds__(256) conv_small_wgrad_kernel(ConvGeom g, const bf16* __restrict__ wide, const bf16* __restrict__ narrow,
                                                               int CW, int CN, int chunks, float* __restrict__ partial) {
  constexpr int NT = KH * KW;
  __shared__ float red[8][64 * 8];                 // [warp][octet-in-warp * 8 + channel] (octets per warp <= 32)
  const int octets = CW >> 3;
  const int P = 256 / octets;                      // position lanes per CTA (octets is a power of two <= 32: whole warps)
  const int tid = threadIdx.x;
  const int octet = tid % octets, lp = tid / octets;
  const int kd = blockIdx.y, cn = blockIdx.z, chunk = blockIdx.x;
  const int dd = SIGN * (kd - g.pd);
  const unsigned M = (unsigned)g.M;
  const unsigned per = (M + chunks - 1) / chunks;
  const unsigned p0 = chunk * per, p1 = min(M, p0 + per);
  float acc[NT][8];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  for (unsigned m = p0 + lp; m < p1; m += P) {
    unsigned t = m;
    const int w = (int)(t % (unsigned)g.W); t /= (unsigned)g.W;
    const int h = (int)(t % (unsigned)g.H); t /= (unsigned)g.H;
    const int d = (int)(t % (unsigned)g.D);
    const int n = (int)(t / (unsigned)g.D);
    const int nd = d + dd;
    if (nd < 0 || nd >= g.D) continue;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(wide + (long long)m * CW) + octet);
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    float wv[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { wv[2 * j] = __uint_as_float(u[j] << 16); wv[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u); }
    const long long base = ((long long)n * g.D + nd) * g.H;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int nh = h + SIGN * (j / KW - KH / 2), nw = w + SIGN * (j % KW - KW / 2);
      float nv = 0.f;
      if (nh >= 0 && nh < g.H && nw >= 0 && nw < g.W) nv = __bfloat162float(narrow[((base + nh) * g.W + nw) * CN + cn]);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(nv, wv[e], acc[j][e]);
    }
  }
  // lanes with the same octet inside a warp: xor-shuffle over the position bits; then the 8 warps through shared memory
  const int warp = tid >> 5, lane = tid & 31;
  const int opw = octets < 32 ? octets : 32;       // octets per warp
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[j][e];
      for (int off = 16; off >= opw; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[j][e] = v;
    }
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if (lane < opw) {
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][lane * 8 + e] = acc[j][e];
    }
    __syncthreads();
    // channel cw = octet * 8 + e; octet = (warp % sets) * 32 + lane when octets > 32
    for (int cw = tid; cw < CW; cw += 256) {
      float s = 0.f;
      if (octets <= 32) {
        for (int wi = 0; wi < 8; ++wi) s += red[wi][cw];
      } else {
        const int set = (cw >> 3) / 32, idx = ((cw >> 3) % 32) * 8 + (cw & 7), sets = octets / 32;
        for (int wi = set; wi < 8; wi += sets) s += red[wi][idx];
      }
      const int tap = kd * NT + j;
      partial[(((long long)chunk * g.taps + tap) * CN + cn) * CW + cw] = s;
    }
    __syncthreads();
  }
}

// dw[co][ci][tap] = sum_chunk partial[chunk][tap][cn][cw];  wide_is_cout: cw = co, cn = ci (first conv), else cw = ci, cn = co
__global__ void conv_small_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int taps, int CN, int CW, int wide_is_cout,
                                               float* __restrict__ dw) {
  const int total = taps * CN * CW;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cw = i % CW, cn = (i / CW) % CN, tap = i / (CW * CN);
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += partial[(long long)c * total + i];
  const int co = wide_is_cout ? cw : cn, ci = wide_is_cout ? cn : cw;
  const int Cin = wide_is_cout ? CN : CW;
  dw[((long long)co * Cin + ci) * taps + tap] = s;
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static int small_wgrad_chunks(const ConvGeom& g, int CW, int CN) {
  const int P = 256 / (CW >> 3);
  long long want = std::max<long long>(1, (long long)sm_count() * 2 / ((long long)g.kd * CN));
  long long cap = std::max<long long>(1, g.M / ((long long)P * 8));
  return (int)std::min<long long>(std::min(want, cap), 1024);
}
bool conv_small_wgrad_supported(const ConvGeom& g, const rho_conv_bwd_params* p) {
  if (p->fwd.act_dtype != RHO_BF16 || !small_fast_geom(g) || g.C2 != 0) return false;
  if (!((g.kh == 3 && g.kw == 3) || (g.kh == 1 && g.kw == 3) || (g.kh == 1 && g.kw == 1))) return false;
  const bool first = g.Cin <= 4 && g.Cout % 8 == 0 && pow2(g.Cout >> 3) && g.Cout <= 2048;
  const bool last = g.Cout <= 4 && g.Cin % 8 == 0 && pow2(g.Cin >> 3) && g.Cin <= 2048;
  return first || last;
}
size_t conv_small_wgrad_workspace(const ConvGeom& g, const rho_conv_bwd_params* p) {
  if (!conv_small_wgrad_supported(g, p)) return 0;
  const bool first = g.Cin <= 4 && g.Cout % 8 == 0;
  const int CW = first ? g.Cout : g.Cin, CN = first ? g.Cin : g.Cout;
  return (size_t)small_wgrad_chunks(g, CW, CN) * g.taps * CN * CW * sizeof(float);
}
int conv_small_wgrad(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  const bool first = g.Cin <= 4 && g.Cout % 8 == 0;
  const int CW = first ? g.Cout : g.Cin, CN = first ? g.Cin : g.Cout;
  const int chunks = small_wgrad_chunks(g, CW, CN);
  const size_t need = (size_t)chunks * g.taps * CN * CW * sizeof(float);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "conv_small_wgrad: workspace %zu < %zu", p->workspace_bytes, need);
  float* partial = (float*)p->workspace;
  const bf16* wide = (const bf16*)(first ? p->dy : p->fwd.x);
  const bf16* narrow = (const bf16*)(first ? p->fwd.x : p->dy);
  dim3 grid((unsigned)chunks, (unsigned)g.kd, (unsigned)CN);
#define RHO_SWG(KH, KW)                                                                                                   \
  do {                                                                                                                    \
    if (first) conv_small_wgrad_kernel<1, KH, KW><<<grid, 256, 0, st>>>(g, wide, narrow, CW, CN, chunks, partial);       \
    else conv_small_wgrad_kernel<-1, KH, KW><<<grid, 256, 0, st>>>(g, wide, narrow, CW, CN, chunks, partial);            \
  } while (0)
  if (g.kh == 3 && g.kw == 3) RHO_SWG(3, 3);
  else if (g.kw == 3) RHO_SWG(1, 3);
  else RHO_SWG(1, 1);
#undef RHO_SWG
  RHO_CHECK_LAUNCH("conv_small_wgrad");
  const int total = g.taps * CN * CW;
  conv_small_wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial, chunks, g.taps, CN, CW, first ? 1 : 0, p->dw);
  RHO_CHECK_LAUNCH("conv_small_wgrad_reduce");
  return RHO_OK;
}

// data gradient of the last conv (Cout <= 4): the first-conv kernel on dY with the flipped filter
bool conv_small_dgrad_supported(const ConvGeom& g, const rho_conv_bwd_params* p) {
  if (p->fwd.act_dtype != RHO_BF16 || p->fwd.w_dtype != RHO_BF16 || !small_fast_geom(g) || g.C2 != 0) return false;
  if ((long long)g.taps * g.Cin * g.Cout > SC_MAX_W) return false;
  return g.Cout <= 4 && g.Cin % 8 == 0 && g.W % 4 == 0 && g.kd % 2 == 1 && g.kh % 2 == 1 && g.kw % 2 == 1;
}
int conv_small_dgrad(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  ConvGeom d = g;                       // swapped roles; unit stride + "same" padding: the grids coincide
  d.Cin = g.Cout; d.C1 = g.Cout; d.Cout = g.Cin;
  d.Ktot = d.taps * d.Cin;
  const size_t smem = (size_t)d.taps * d.Cin * d.Cout * sizeof(float);
  if (run_geom(d) && d.Wo % 8 == 0) {
    const unsigned blocks8 = (unsigned)ceil_div((d.M >> 3) * (d.Cout >> 3), 256);
#define RHO_SCD8(CI) conv_small_cin_row8_kernel<CI><<<blocks8, 256, smem, st>>>(d, (const bf16*)p->dy, (const bf16*)p->fwd.w_packed, nullptr, (bf16*)p->dx, 1)
    switch (d.Cin) {
      case 1: RHO_SCD8(1); break;
      case 2: RHO_SCD8(2); break;
      case 3: RHO_SCD8(3); break;
      default: RHO_SCD8(4); break;
    }
#undef RHO_SCD8
    RHO_CHECK_LAUNCH("conv_small_dgrad8");
    return RHO_OK;
  }
  const unsigned blocks = (unsigned)ceil_div((d.M >> 2) * (d.Cout >> 3), 256);
#define RHO_SCD(CI) conv_small_cin_row4_kernel<CI><<<blocks, 256, smem, st>>>(d, (const bf16*)p->dy, (const bf16*)p->fwd.w_packed, nullptr, (bf16*)p->dx, 1)
  switch (d.Cin) {
    case 1: RHO_SCD(1); break;
    case 2: RHO_SCD(2); break;
    case 3: RHO_SCD(3); break;
    default: RHO_SCD(4); break;
  }
#undef RHO_SCD
  RHO_CHECK_LAUNCH("conv_small_dgrad");
  return RHO_OK;
}

int conv_small_fwd(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  const int a = p->act_dtype, w = p->w_dtype, f = p->y_f32;
  if (a == RHO_BF16 && w == RHO_BF16 && small_fast_geom(g)) {
    if (g.Cout <= 4 && g.Cin % 8 == 0) return f ? launch_small_cout_vec<float>(g, p, st) : launch_small_cout_vec<bf16>(g, p, st);
    if (g.Cin <= 4 && g.Cout % 8 == 0 && g.Wo % 4 == 0 && !f) return launch_small_cin_row4(g, p, st);
  }
  if (a == RHO_F32 && w == RHO_F32) return launch_small<float, float, float>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16 && !f) return launch_small<bf16, bf16, bf16>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16 && f) return launch_small<bf16, bf16, float>(g, p, st);
  return fail(RHO_ERR_UNSUPPORTED, "conv_small: dtype combination");
}

}  // namespace rho
