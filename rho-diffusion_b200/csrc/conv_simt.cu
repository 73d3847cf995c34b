// conv_nd on CUDA cores: the exact-fp32 engine (TF32 off, as scripts/training_ddp.py:42-43 demands) and
// the general fallback for shapes the tcgen05 engine does not take (Cin = 1/3, Cout = 1/3, odd tails).
//
// Implicit GEMM on channels-last activations:
//   forward : M = N*Do*Ho*Wo output positions, N = Cout, K = taps*Cin
//   dgrad   : M = N*D*H*W input positions,     N = Cin,  K = (upsample children x taps)*Cout
//   wgrad   : M = Cout, N = taps*Cin, K = output positions (split-K, deterministic two-stage reduction)
// The gather handles "same" padding, stride 1/2, fused nearest x2 upsample and the two-source channel
// concat; the epilogue adds bias and the residual.  64x64x16 tiles, 256 threads, 4x4 register tiles,
// register-prefetched shared-memory staging.
#include "conv_geom.cuh"

namespace rho {

constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;

template <typename T, bool VEC>
__device__ __forceinline__ void load_a4(const T* __restrict__ x, const T* __restrict__ x2, int C1, int C2, long long pos, int c,
                                        float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (pos < 0) return;
  const int Cin = C1 + C2;
  if (VEC) {
    if (c >= Cin) return;
    if (c < C1) load4(x + pos * C1 + c, v);
    else load4(x2 + pos * C2 + (c - C1), v);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int cc = c + j;
      if (cc < C1) v[j] = to_f32(x[pos * C1 + cc]);
      else if (cc < Cin) v[j] = to_f32(x2[pos * C2 + (cc - C1)]);
    }
  }
}

// MODE 0: forward (src = x/x2, out channels = Cout, K channels = Cin)
// MODE 1: dgrad   (src = dy,   out channels = Cin,  K channels = Cout)
template <typename TA, typename TW, typename TO, int MODE, bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvGeom g, const TA* __restrict__ src, const TA* __restrict__ src2,
                                                        const TW* __restrict__ w, const float* __restrict__ bias,
                                                        const TA* __restrict__ residual, TO* __restrict__ out, TO* __restrict__ out2) {
  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  const int KC = MODE == 0 ? g.Cin : g.Cout;   // channels reduced per (virtual) tap
  const int NC = MODE == 0 ? g.Cout : g.Cin;   // output channels
  const int sC1 = MODE == 0 ? g.C1 : g.Cout;   // source tensor channel split
  const int sC2 = MODE == 0 ? g.C2 : 0;
  const int vtaps = MODE == 0 ? g.taps : g.taps * g.ud * g.uh * g.uw;

  // A-load role: one row, 4 consecutive k
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;
  const RowPos rp = MODE == 0 ? decode_out(g, m0 + a_row) : decode_in(g, m0 + a_row);
  // B-load role
  const int b_n = MODE == 0 ? (tid >> 2) : (tid & 15) * 4;
  const int b_k = MODE == 0 ? (tid & 3) * 4 : (tid >> 4);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int kblocks = (KC + BK - 1) / BK;
  const int total = vtaps * kblocks;
  float ra[4], rb[4];

  auto fetch = [&](int it) {
    const int vt = it / kblocks, c0 = (it % kblocks) * BK;
    const long long pos = MODE == 0 ? fwd_src(g, rp, vt) : dgrad_src(g, rp, vt);
    load_a4<TA, VEC>(src, src2, sC1, sC2, pos, c0 + a_k, ra);
    const int tap = MODE == 0 ? vt : vt % g.taps;
    rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
    if (MODE == 0) {
      // B[k = ci][n = co] = w[co][tap][ci]; thread reads 4 consecutive ci of one co
      const int co = n0 + b_n, ci = c0 + b_k;
      if (co < g.Cout) {
        const TW* wp = w + (long long)co * g.Ktot + (long long)tap * g.Cin + ci;
        if (VEC) { if (ci < g.Cin) load4(wp, rb); }
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ci + j < g.Cin) rb[j] = to_f32(wp[j]);
        }
      }
    } else {
      // B[k = co][n = ci] = w[co][tap][ci]; thread reads 4 consecutive ci of one co
      const int co = c0 + b_k, ci = n0 + b_n;
      if (co < g.Cout) {
        const TW* wp = w + (long long)co * g.Ktot + (long long)tap * g.Cin + ci;
        if (VEC) { if (ci < g.Cin) load4(wp, rb); }
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (ci + j < g.Cin) rb[j] = to_f32(wp[j]);
        }
      }
    }
  };

  fetch(0);
  for (int it = 0; it < total; ++it) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[a_k + j][a_row] = ra[j];
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[b_k + j][b_n] = rb[j];
    } else {
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
    __syncthreads();
    if (it + 1 < total) fetch(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
  const long long Mtot = MODE == 0 ? g.M : g.Min;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= Mtot) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= NC) continue;
      float v = acc[i][j];
      if (MODE == 0) {
        if (bias) v += bias[c];
        if (residual) v += to_f32(residual[m * g.Cout + c]);
        out[m * g.Cout + c] = from_f32<TO>(v);
      } else {
        if (c < g.C1) out[m * g.C1 + c] = from_f32<TO>(v);
        else out2[m * g.C2 + (c - g.C1)] = from_f32<TO>(v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: partial[split][co][tap*Cin + ci] = sum over this split's output positions of dy[pos][co] * x[src(pos,tap)][ci]
// grid: (ceil(Cout/64), taps * ceil(Cin/64), splits)
// ------------------------------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(ConvGeom g, const TA* __restrict__ x, const TA* __restrict__ x2,
                                                              const TA* __restrict__ dy, int splits, long long per_split,
                                                              float* __restrict__ partial) {
  __shared__ __align__(16) float As[BK][LDS];  // [pos][co]
  __shared__ __align__(16) float Bs[BK][LDS];  // [pos][ci]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int cin_blocks = (g.Cin + BN - 1) / BN;
  const int co0 = blockIdx.x * BM;
  const int tap = blockIdx.y / cin_blocks;
  const int ci0 = (blockIdx.y % cin_blocks) * BN;
  const int split = blockIdx.z;
  const long long pbeg = (long long)split * per_split;
  const long long pend = min(g.M, pbeg + per_split);

  const int l_pos = tid >> 4, l_c = (tid & 15) * 4;  // each thread: one position, 4 consecutive channels
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto fetch = [&](long long p) {
    const long long pos = p + l_pos;
    ra[0] = ra[1] = ra[2] = ra[3] = 0.f;
    rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
    if (pos < pend) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int co = co0 + l_c + j;
        if (co < g.Cout) ra[j] = to_f32(dy[pos * g.Cout + co]);
      }
      RowPos rp = decode_out(g, pos);
      long long src = fwd_src(g, rp, tap);
      if (src >= 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int ci = ci0 + l_c + j;
          if (ci < g.C1) rb[j] = to_f32(x[src * g.C1 + ci]);
          else if (ci < g.Cin) rb[j] = to_f32(x2[src * g.C2 + (ci - g.C1)]);
        }
      }
    }
  };

  if (pbeg < pend) fetch(pbeg);
  for (long long p = pbeg; p < pend; p += BK) {
    *reinterpret_cast<float4*>(&As[l_pos][l_c]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&Bs[l_pos][l_c]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    __syncthreads();
    if (p + BK < pend) fetch(p + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + (long long)split * g.Cout * g.Ktot;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int ci = ci0 + tx * 4 + j;
      if (ci < g.Cin) out[(long long)co * g.Ktot + (long long)tap * g.Cin + ci] = acc[i][j];
    }
  }
}

// dw[co][ci][tap] = sum_split partial[split][co][tap][ci]
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cout, int Cin, int taps,
                                         float* __restrict__ dw) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Cout * Cin * taps;
  if (i >= total) return;
  // i indexes the packed layout [co][tap][ci] for coalesced reads
  int ci = (int)(i % Cin);
  int tap = (int)((i / Cin) % taps);
  int co = (int)(i / ((long long)Cin * taps));
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(long long)k * total + i];
  dw[((long long)co * Cin + ci) * taps + tap] = s;
}

// dbias[c] = sum_m dy[m][c]  (two-stage, deterministic)
template <typename TA>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const TA* __restrict__ dy, long long M, int C, int chunks,
                                                             float* __restrict__ partial) {
  const long long per = (M + chunks - 1) / chunks;
  const long long m0 = (long long)blockIdx.x * per, m1 = min(M, m0 + per);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long m = m0; m < m1; ++m) s += to_f32(dy[m * C + c]);
    partial[(long long)blockIdx.x * C + c] = s;
  }
}
// bf16, C % 8 == 0, C <= 2048: 16-byte loads, 4 rows in flight per thread, rows of a CTA meet in shared memory
__global__ void __launch_bounds__(256) colsum_partial8_kernel(const bf16* __restrict__ dy, long long M, int C, int chunks,
                                                              float* __restrict__ partial) {
  __shared__ float red[2048];
  const int CV = C >> 3;
  const int rows = 256 / CV;
  const int row = threadIdx.x / CV, cv = threadIdx.x - row * CV;
  const long long per = (M + chunks - 1) / chunks;
  const long long m0 = (long long)blockIdx.x * per, m1 = min(M, m0 + per);
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (row < rows) {
    const bf16* src = dy + cv * 8;
    auto accum = [&](const uint4& q) {
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[2 * j] += __uint_as_float(u[j] << 16); s[2 * j + 1] += __uint_as_float(u[j] & 0xffff0000u); }
    };
    long long m = m0 + row;
    for (; m + 3ll * rows < m1; m += 4ll * rows) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(src + (m + (long long)u * rows) * C));
#pragma unroll
      for (int u = 0; u < 4; ++u) accum(q[u]);
    }
    for (; m < m1; m += rows) accum(__ldg(reinterpret_cast<const uint4*>(src + m * C)));
#pragma unroll
    for (int j = 0; j < 8; ++j) red[row * C + cv * 8 + j] = s[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float t = 0.f;
    for (int r = 0; r < rows; ++r) t += red[r * C + c];
    partial[(long long)blockIdx.x * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int chunks, int C, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[(long long)k * C + c];
  out[c] = s;
}

template <typename TP>
__global__ void pack_weight_kernel(const float* __restrict__ w, TP* __restrict__ packed, int Cout, int Cin, int taps) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Cout * Cin * taps;
  if (i >= total) return;
  int ci = (int)(i % Cin);
  int tap = (int)((i / Cin) % taps);
  int co = (int)(i / ((long long)Cin * taps));
  packed[i] = from_f32<TP>(w[((long long)co * Cin + ci) * taps + tap]);
}

// same re-layout through shared memory: one CTA = one output channel x 64 input channels x all taps; the fp32 source run
// (64 * taps contiguous floats) is read coalesced, each tap row of 64 packed values is written contiguously
template <typename TP>
__global__ void __launch_bounds__(256) pack_weight_tiled_kernel(const float* __restrict__ w, TP* __restrict__ packed, int Cin, int taps) {
  extern __shared__ float wt[];                        // [64 * taps]
  const int co = blockIdx.x, ci0 = blockIdx.y * 64;
  const int nci = min(64, Cin - ci0);
  const float* src = w + ((long long)co * Cin + ci0) * taps;
  for (int i = threadIdx.x; i < nci * taps; i += 256) wt[i] = src[i];
  __syncthreads();
  TP* dst = packed + (long long)co * taps * Cin + ci0;
  for (int i = threadIdx.x; i < nci * taps; i += 256) {
    const int tap = i / nci, ci = i - tap * nci;
    dst[(long long)tap * Cin + ci] = from_f32<TP>(wt[ci * taps + tap]);
  }
}

// ---- host -----------------------------------------------------------------------------------------
int wgrad_splits(const ConvGeom& g) {
  long long tiles = ceil_div(g.Cout, BM) * (long long)g.taps * ceil_div(g.Cin, BN);
  long long want = ceil_div((long long)sm_count() * 2, tiles);
  long long max_splits = ceil_div(g.M, 256);
  long long s = std::min(want, max_splits);
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

// RHO_TRACE_SIMT=1: one stderr line per launch of a CUDA-core conv engine, to find the layers that fall off the tensor-core path
static void trace_simt(const char* what, const ConvGeom& g) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("RHO_TRACE_SIMT"); on = (e && e[0] == '1') ? 1 : 0; }
  if (on)
    fprintf(stderr, "[rho] %s on CUDA cores: N=%d DHW=%dx%dx%d C1=%d C2=%d Cout=%d k=%dx%dx%d s=%d%d%d u=%d%d%d\n", what, g.N, g.D, g.H, g.W,
            g.C1, g.C2, g.Cout, g.kd, g.kh, g.kw, g.sd, g.sh, g.sw, g.ud, g.uh, g.uw);
}

template <typename TA, typename TW, typename TO>
static int launch_fwd_t(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(g.M, BM), (unsigned)ceil_div(g.Cout, BN));
  const bool vec = (g.C1 % 4 == 0) && (g.C2 % 4 == 0);
  if (vec)
    conv_simt_kernel<TA, TW, TO, 0, true><<<grid, 256, 0, st>>>(g, (const TA*)p->x, (const TA*)p->x2, (const TW*)p->w_packed, p->bias,
                                                               (const TA*)p->residual, (TO*)p->y, (TO*)nullptr);
  else
    conv_simt_kernel<TA, TW, TO, 0, false><<<grid, 256, 0, st>>>(g, (const TA*)p->x, (const TA*)p->x2, (const TW*)p->w_packed, p->bias,
                                                                (const TA*)p->residual, (TO*)p->y, (TO*)nullptr);
  RHO_CHECK_LAUNCH("conv_fwd_simt");
  return RHO_OK;
}

int conv_fwd_simt(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  trace_simt("conv fwd", g);
  const int a = p->act_dtype, w = p->w_dtype, f = p->y_f32;
  if (a == RHO_F32 && w == RHO_F32) return launch_fwd_t<float, float, float>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16 && !f) return launch_fwd_t<bf16, bf16, bf16>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16 && f) return launch_fwd_t<bf16, bf16, float>(g, p, st);
  if (a == RHO_BF16 && w == RHO_F32 && !f) return launch_fwd_t<bf16, float, bf16>(g, p, st);
  if (a == RHO_BF16 && w == RHO_F32 && f) return launch_fwd_t<bf16, float, float>(g, p, st);
  if (a == RHO_F32 && w == RHO_BF16) return launch_fwd_t<float, bf16, float>(g, p, st);
  return fail(RHO_ERR_INVALID_ARG, "conv_fwd: bad dtype combination");
}

template <typename TA, typename TW>
static int launch_dgrad_t(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(g.Min, BM), (unsigned)ceil_div(g.Cin, BN));
  const bool vec = (g.Cout % 4 == 0) && (g.Cin % 4 == 0) && (g.C1 % 4 == 0);
  if (vec)
    conv_simt_kernel<TA, TW, TA, 1, true><<<grid, 256, 0, st>>>(g, (const TA*)p->dy, (const TA*)nullptr, (const TW*)p->fwd.w_packed,
                                                               nullptr, (const TA*)nullptr, (TA*)p->dx, (TA*)p->dx2);
  else
    conv_simt_kernel<TA, TW, TA, 1, false><<<grid, 256, 0, st>>>(g, (const TA*)p->dy, (const TA*)nullptr, (const TW*)p->fwd.w_packed,
                                                                nullptr, (const TA*)nullptr, (TA*)p->dx, (TA*)p->dx2);
  RHO_CHECK_LAUNCH("conv_dgrad_simt");
  return RHO_OK;
}

int conv_dgrad_simt(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  trace_simt("conv dgrad", g);
  const int a = p->fwd.act_dtype, w = p->fwd.w_dtype;
  if (a == RHO_F32 && w == RHO_F32) return launch_dgrad_t<float, float>(g, p, st);
  if (a == RHO_BF16 && w == RHO_BF16) return launch_dgrad_t<bf16, bf16>(g, p, st);
  if (a == RHO_BF16 && w == RHO_F32) return launch_dgrad_t<bf16, float>(g, p, st);
  if (a == RHO_F32 && w == RHO_BF16) return launch_dgrad_t<float, bf16>(g, p, st);
  return fail(RHO_ERR_INVALID_ARG, "conv_dgrad: bad dtype combination");
}

size_t conv_wgrad_simt_workspace(const ConvGeom& g) {
  size_t ws = (size_t)wgrad_splits(g) * g.Cout * g.Ktot * sizeof(float);
  size_t bias_ws = (size_t)256 * g.Cout * sizeof(float);
  return align_up(ws, 256) + bias_ws;
}

int conv_wgrad_simt(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  const int splits = wgrad_splits(g);
  size_t need = conv_wgrad_simt_workspace(g);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "conv_wgrad: workspace %zu < %zu", p->workspace_bytes, need);
  float* partial = (float*)p->workspace;
  float* bias_ws = (float*)((char*)p->workspace + align_up((size_t)splits * g.Cout * g.Ktot * sizeof(float), 256));
  if (p->dw) {
    trace_simt("conv wgrad", g);
    RHO_REQUIRE(p->fwd.x, "conv_wgrad: x required");
    long long per_split = ceil_div(ceil_div(g.M, splits), BK) * BK;
    dim3 grid((unsigned)ceil_div(g.Cout, BM), (unsigned)(g.taps * ceil_div(g.Cin, BN)), (unsigned)splits);
    if (p->fwd.act_dtype == RHO_F32)
      conv_wgrad_simt_kernel<float><<<grid, 256, 0, st>>>(g, (const float*)p->fwd.x, (const float*)p->fwd.x2, (const float*)p->dy, splits,
                                                         per_split, partial);
    else
      conv_wgrad_simt_kernel<bf16><<<grid, 256, 0, st>>>(g, (const bf16*)p->fwd.x, (const bf16*)p->fwd.x2, (const bf16*)p->dy, splits,
                                                        per_split, partial);
    RHO_CHECK_LAUNCH("conv_wgrad_simt");
    long long total = (long long)g.Cout * g.Ktot;
    conv_wgrad_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(partial, splits, g.Cout, g.Cin, g.taps, p->dw);
    RHO_CHECK_LAUNCH("conv_wgrad_reduce");
  }
  if (p->dbias) {
    int chunks = (int)std::min<long long>(256, std::max<long long>(1, g.M / 64));
    if (p->fwd.act_dtype == RHO_F32)
      colsum_partial_kernel<float><<<chunks, 256, 0, st>>>((const float*)p->dy, g.M, g.Cout, chunks, bias_ws);
    else if (g.Cout % 8 == 0 && g.Cout <= 2048)
      colsum_partial8_kernel<<<chunks, 256, 0, st>>>((const bf16*)p->dy, g.M, g.Cout, chunks, bias_ws);
    else
      colsum_partial_kernel<bf16><<<chunks, 256, 0, st>>>((const bf16*)p->dy, g.M, g.Cout, chunks, bias_ws);
    RHO_CHECK_LAUNCH("colsum_partial");
    colsum_final_kernel<<<(g.Cout + 127) / 128, 128, 0, st>>>(bias_ws, chunks, g.Cout, p->dbias);
    RHO_CHECK_LAUNCH("colsum_final");
  }
  return RHO_OK;
}

}  // namespace rho

using namespace rho;

extern "C" int rho_conv_pack_weight(const float* w, void* packed, int packed_dtype, int32_t Cout, int32_t Cin, int32_t taps,
                                    rho_stream_t stream) {
  RHO_REQUIRE(w && packed && Cout > 0 && Cin > 0 && taps > 0, "conv_pack_weight: bad args");
  long long total = (long long)Cout * Cin * taps;
  unsigned blocks = (unsigned)ceil_div(total, 256);
  if (taps > 1 && taps <= 64 && Cout <= 65535 * 32) {
    dim3 grid((unsigned)Cout, (unsigned)ceil_div(Cin, 64));
    const size_t smem = (size_t)64 * taps * sizeof(float);
    if (packed_dtype == RHO_F32) pack_weight_tiled_kernel<float><<<grid, 256, smem, as_stream(stream)>>>(w, (float*)packed, Cin, taps);
    else if (packed_dtype == RHO_BF16) pack_weight_tiled_kernel<bf16><<<grid, 256, smem, as_stream(stream)>>>(w, (bf16*)packed, Cin, taps);
    else return fail(RHO_ERR_INVALID_ARG, "conv_pack_weight: bad dtype");
    RHO_CHECK_LAUNCH("pack_weight");
    return RHO_OK;
  }
  if (packed_dtype == RHO_F32) pack_weight_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(w, (float*)packed, Cout, Cin, taps);
  else if (packed_dtype == RHO_BF16) pack_weight_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>(w, (bf16*)packed, Cout, Cin, taps);
  else return fail(RHO_ERR_INVALID_ARG, "conv_pack_weight: bad dtype");
  RHO_CHECK_LAUNCH("pack_weight");
  return RHO_OK;
}
