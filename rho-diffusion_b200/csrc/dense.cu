// Embedding path (fp32, tiny): sinusoidal timestep embedding (rho_diffusion/models/common.py:27-43), the
// nn.Linear layers of `time_embed` and `emb_layers` (rho_diffusion/models/unet_v2.py:521-525, 229-235)
// with the preceding activation fused, and MultiEmbeddings.forward
// (rho_diffusion/models/conditioning.py:115-139).  B is the batch (<= a few hundred), so these are
// latency-bound GEMVs: one warp per output element, coalesced over the reduction dimension.
#include "common.cuh"

namespace rho {

// omega[k] = wavelength^(2k/dim) is a per-model constant table built on the host exactly as the reference builds
// it (float32 torch.pow); powf on the device may differ by an ulp, which sin/cos at t ~ 1000 would amplify to 1e-4.
__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, int B, int dim, const float* __restrict__ omega_tab,
                                          float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int half = dim / 2;
  if (i >= B * half) return;
  int b = i / half, k = i % half;
  float arg = (float)t[b] / omega_tab[k];
  out[(long long)b * dim + 2 * k] = sinf(arg);
  out[(long long)b * dim + 2 * k + 1] = cosf(arg);
}

// y[b][o] = sum_i act(x[b][i]) * w[o][i] + bias[o] (+ add[b][o]); one warp per (b, o)
__global__ void __launch_bounds__(256) linear_fwd_kernel(rho_linear_params p) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p.B * p.O) return;
  const int b = warp / p.O, o = warp % p.O;
  const float* x = p.x + (long long)b * p.x_stride;
  const float* w = p.w + (long long)o * p.I;
  float s = 0.f;
  for (int i = lane; i < p.I; i += 32) s = fmaf(apply_act(x[i], p.act_in), w[i], s);
  s = warp_sum(s);
  if (lane == 0) {
    if (p.b) s += p.b[o];
    if (p.add) s += p.add[(long long)b * p.add_stride + o];
    p.y[(long long)b * p.y_stride + o] = s;
  }
}

// dx[b][i] = act'(x[b][i]) * sum_o dy[b][o] * w[o][i]; thread per (b, i) (coalesced over i in w rows)
// CTA = 32 inputs x 8 output lanes (grid: I/32 x B): w rows are read as 128-byte segments, the 8 partial sums per input
// meet in shared memory in a fixed order
__global__ void __launch_bounds__(256) linear_bwd_dx_kernel(rho_linear_params p, const float* __restrict__ dy, int dy_stride,
                                                            float* __restrict__ dx, int dx_stride) {
  __shared__ float red[8][32];
  const int b = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (i < p.I) {
    const float* dyb = dy + (long long)b * dy_stride;
#pragma unroll 4
    for (int o = ty; o < p.O; o += 8) s = fmaf(dyb[o], p.w[(long long)o * p.I + i], s);
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < p.I) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    dx[(long long)b * dx_stride + i] = t * act_grad(p.x[(long long)b * p.x_stride + i], p.act_in);
  }
}

// dw[o][i] = sum_b dy[b][o] * act(x[b][i]); thread per (o, i)
__global__ void linear_bwd_dw_kernel(rho_linear_params p, const float* __restrict__ dy, int dy_stride, float* __restrict__ dw) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.O * p.I) return;
  int o = (int)(idx / p.I), i = (int)(idx % p.I);
  float s = 0.f;
  for (int b = 0; b < p.B; ++b) s = fmaf(dy[(long long)b * dy_stride + o], apply_act(p.x[(long long)b * p.x_stride + i], p.act_in), s);
  dw[idx] = s;
}

// Batches of 32 rows and more (config C samples 512 sequences at a time) turn the GEMVs above into real GEMMs: the warp-per-output
// kernels re-read a weight row per (b, o) from L2.  One tiled fp32 kernel serves the three products
//     forward   y[b][o]  = sum_i act(x[b][i]) w[o][i]  (+ bias, + add)      M = B, N = O, K = I   A, B contiguous along K
//     dx        dx[b][i] = act'(x[b][i]) sum_o dy[b][o] w[o][i]              M = B, N = I, K = O   A along K, B along N
//     dw        dw[o][i] = sum_b dy[b][o] act(x[b][i])                       M = O, N = I, K = B   A along M, B along N
// 64 x 64 tile, K step 16, 256 threads with 4 x 4 register tiles; the K order is fixed, so results do not depend on the grid.
enum { LIN_FWD = 0, LIN_DX = 1, LIN_DW = 2 };

template <int MODE>
__global__ void __launch_bounds__(256) linear_tiled_kernel(rho_linear_params p, const float* __restrict__ dy, int dy_stride,
                                                           float* __restrict__ out, int out_stride) {
  constexpr int T = 64, KS = 16, LD = T + 4;
  __shared__ float As[KS][LD], Bs[KS][LD];
  const int M = MODE == LIN_DW ? p.O : p.B;
  const int N = MODE == LIN_FWD ? p.O : p.I;
  const int K = MODE == LIN_FWD ? p.I : (MODE == LIN_DX ? p.O : p.B);
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += KS) {
    // ---- A tile -> As[k][m] ----
    if (MODE == LIN_DW) {                      // A(m, k) = dy[k][m]: contiguous along m
      const int k = k0 + (t >> 4), m = m0 + (t & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) As[t >> 4][(t & 15) * 4 + j] = (k < K && m + j < M) ? dy[(long long)k * dy_stride + m + j] : 0.f;
    } else {                                   // contiguous along k: act(x[m][k]) resp. dy[m][k]
      const int m = m0 + (t >> 2), k = k0 + (t & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (m < M && k + j < K) v = MODE == LIN_FWD ? apply_act(p.x[(long long)m * p.x_stride + k + j], p.act_in) : dy[(long long)m * dy_stride + k + j];
        As[(t & 3) * 4 + j][t >> 2] = v;
      }
    }
    // ---- B tile -> Bs[k][n] ----
    if (MODE == LIN_FWD) {                     // B(k, n) = w[n][k]: contiguous along k
      const int n = n0 + (t >> 2), k = k0 + (t & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[(t & 3) * 4 + j][t >> 2] = (n < N && k + j < K) ? p.w[(long long)n * p.I + k + j] : 0.f;
    } else {                                   // contiguous along n: w[k][n] resp. act(x[k][n])
      const int k = k0 + (t >> 4), n = n0 + (t & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.f;
        if (k < K && n + j < N) v = MODE == LIN_DX ? p.w[(long long)k * p.I + n + j] : apply_act(p.x[(long long)k * p.x_stride + n + j], p.act_in);
        Bs[t >> 4][(t & 15) * 4 + j] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (MODE == LIN_FWD) {
        if (p.b) v += p.b[n];
        if (p.add) v += p.add[(long long)m * p.add_stride + n];
      } else if (MODE == LIN_DX) {
        v *= act_grad(p.x[(long long)m * p.x_stride + n], p.act_in);
      }
      out[(long long)m * out_stride + n] = v;
    }
  }
}

constexpr int LIN_TILED_MIN_B = 32;

__global__ void linear_bwd_db_kernel(int B, int O, const float* __restrict__ dy, int dy_stride, float* __restrict__ db) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= O) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dy[(long long)b * dy_stride + o];
  db[o] = s;
}

__global__ void embedding_match_kernel(const float* __restrict__ y, int y_stride, int B, const float* __restrict__ values, int V,
                                       const float* __restrict__ table, int dim, float* __restrict__ out, int accumulate,
                                       int* __restrict__ idx_out) {
  const int b = blockIdx.x;
  __shared__ int s_idx;
  if (threadIdx.x == 0) {
    const float v = y[(long long)b * y_stride];
    int found = -1;
    for (int k = 0; k < V; ++k)
      if (values[k] == v) { found = k; break; }
    s_idx = found;
    if (idx_out) idx_out[b] = found;
  }
  __syncthreads();
  const int k = s_idx;
  for (int d = threadIdx.x; d < dim; d += blockDim.x) {
    // a label that is not a point of the parameter space has no embedding: the reference fails on it (conditioning.py:132-137);
    // here the row is poisoned with NaN (loud downstream, no host sync) and idx_out carries -1 for the host-side check
    float e = k >= 0 ? table[(long long)k * dim + d] : __int_as_float(0x7fc00000);
    float* o = out + (long long)b * dim + d;
    *o = accumulate ? *o + e : e;
  }
}

// thread per (v, d): deterministic accumulation over the batch in index order
__global__ void embedding_match_bwd_kernel(const int* __restrict__ idx, int B, const float* __restrict__ dout, int dim,
                                           float* __restrict__ dtable, int V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V * dim) return;
  int v = i / dim, d = i % dim;
  float s = 0.f;
  for (int b = 0; b < B; ++b)
    if (idx[b] == v) s += dout[(long long)b * dim + d];
  dtable[i] += s;
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_timestep_embedding(const int64_t* t, int32_t B, int32_t dim, const float* omega, float* out, rho_stream_t stream) {
  RHO_REQUIRE(t && out && omega && B > 0 && dim > 0 && dim % 2 == 0, "timestep_embedding: bad args (dim must be even)");
  int total = B * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(t, B, dim, omega, out);
  RHO_CHECK_LAUNCH("timestep_embedding");
  return RHO_OK;
}

int rho_linear_fwd(const rho_linear_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p && p->x && p->w && p->y && p->B > 0 && p->I > 0 && p->O > 0, "linear_fwd: bad args");
  if (p->B >= LIN_TILED_MIN_B) {
    linear_tiled_kernel<LIN_FWD><<<dim3((unsigned)ceil_div(p->O, 64), (unsigned)ceil_div(p->B, 64)), 256, 0, as_stream(stream)>>>(
        *p, nullptr, 0, p->y, p->y_stride);
    RHO_CHECK_LAUNCH("linear_fwd_tiled");
    return RHO_OK;
  }
  long long warps = (long long)p->B * p->O;
  linear_fwd_kernel<<<(unsigned)ceil_div(warps * 32, 256), 256, 0, as_stream(stream)>>>(*p);
  RHO_CHECK_LAUNCH("linear_fwd");
  return RHO_OK;
}

int rho_linear_bwd(const rho_linear_params* p, const float* dy, int32_t dy_stride, float* dx, int32_t dx_stride, float* dw, float* db,
                   rho_stream_t stream) {
  RHO_REQUIRE(p && p->x && p->w && dy && p->B > 0 && p->I > 0 && p->O > 0, "linear_bwd: bad args");
  cudaStream_t st = as_stream(stream);
  const bool tiled = p->B >= LIN_TILED_MIN_B;
  if (dx && tiled) {
    linear_tiled_kernel<LIN_DX><<<dim3((unsigned)ceil_div(p->I, 64), (unsigned)ceil_div(p->B, 64)), 256, 0, st>>>(*p, dy, dy_stride, dx, dx_stride);
    RHO_CHECK_LAUNCH("linear_bwd_dx_tiled");
  } else if (dx) {
    linear_bwd_dx_kernel<<<dim3((unsigned)ceil_div(p->I, 32), (unsigned)p->B), 256, 0, st>>>(*p, dy, dy_stride, dx, dx_stride);
    RHO_CHECK_LAUNCH("linear_bwd_dx");
  }
  if (dw && tiled) {
    linear_tiled_kernel<LIN_DW><<<dim3((unsigned)ceil_div(p->I, 64), (unsigned)ceil_div(p->O, 64)), 256, 0, st>>>(*p, dy, dy_stride, dw, p->I);
    RHO_CHECK_LAUNCH("linear_bwd_dw_tiled");
  } else if (dw) {
    linear_bwd_dw_kernel<<<(unsigned)ceil_div((long long)p->O * p->I, 128), 128, 0, st>>>(*p, dy, dy_stride, dw);
    RHO_CHECK_LAUNCH("linear_bwd_dw");
  }
  if (db) {
    linear_bwd_db_kernel<<<(p->O + 127) / 128, 128, 0, st>>>(p->B, p->O, dy, dy_stride, db);
    RHO_CHECK_LAUNCH("linear_bwd_db");
  }
  return RHO_OK;
}

int rho_embedding_match(const float* y, int32_t y_stride, int32_t B, const float* values, int32_t V, const float* table, int32_t dim,
                        float* out, int32_t accumulate, int32_t* idx_out, rho_stream_t stream) {
  RHO_REQUIRE(y && values && table && out && B > 0 && V > 0 && dim > 0, "embedding_match: bad args");
  embedding_match_kernel<<<B, 128, 0, as_stream(stream)>>>(y, y_stride, B, values, V, table, dim, out, accumulate, idx_out);
  RHO_CHECK_LAUNCH("embedding_match");
  return RHO_OK;
}

int rho_embedding_match_bwd(const int32_t* idx, int32_t B, const float* dout, int32_t dim, float* dtable, int32_t V, rho_stream_t stream) {
  RHO_REQUIRE(idx && dout && dtable && B > 0 && V > 0 && dim > 0, "embedding_match_bwd: bad args");
  embedding_match_bwd_kernel<<<(V * dim + 127) / 128, 128, 0, as_stream(stream)>>>(idx, B, dout, dim, dtable, V);
  RHO_CHECK_LAUNCH("embedding_match_bwd");
  return RHO_OK;
}

}  // extern "C"
