// Kernels that only the two remaining registered backbones need: the v1 UNet (rho_diffusion/models/unet.py:30-269:
// conv -> activation chains, ConvTranspose with stride 1) and the diffusion VisionTransformer (models/vit.py: patch embedding,
// LayerNorm, token MLPs, kernel = stride transposed conv).  Everything here is HBM-bound element-wise / row-wise work; the
// contractions of both backbones run on the conv_nd and attention engines.
#include "common.cuh"

namespace rho {

// ---- stand-alone activation (unet.py:122-123: activation(conv(x)); vit.py:155-161: Linear -> activation) -----------------
template <typename T>
__global__ void __launch_bounds__(256) activation_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n4, int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4];
    load4(x + i * 4, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = apply_act(v[e], act);
    store4(y + i * 4, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {          // tail (n % 4 elements)
    const int64_t i = n4 * 4 + threadIdx.x;
    y[i] = from_f32<T>(apply_act(to_f32(x[i]), act));
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    activation_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int64_t n4, int64_t n, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4], g[4];
    load4(x + i * 4, v);
    load4(dy + i * 4, g);
#pragma unroll
    for (int e = 0; e < 4; ++e) g[e] *= act_grad(v[e], act);
    store4(dx + i * 4, g);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    dx[i] = from_f32<T>(to_f32(dy[i]) * act_grad(to_f32(x[i]), act));
  }
}

// ---- ConvTranspose (stride 1) as a convolution: dst[b][a][taps-1-t] = src[a][b][t] --------------------------------------
// nn.ConvTranspose{2,3}d(k, stride 1, padding k//2) with weight [Cin, Cout, *k] equals conv_nd with the weight
// [Cout, Cin, *k] flipped along every kernel axis (unet.py:69-75, 94-100).  The map is its own adjoint with A and B swapped,
// so the same kernel turns the gradient of the effective weight into the gradient of the parameter.  taps = 1: plain transpose.
__global__ void convt_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int B, int taps) {
  const int64_t total = (int64_t)A * B * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int a = (int)((i / taps) % A);
    const int b = (int)(i / ((int64_t)taps * A));
    dst[i] = src[((int64_t)a * B + b) * taps + (taps - 1 - t)];
  }
}

// ---- patch embedding layout (vit.py:72-77, 119-126, 358-371) -----------------------------------------------------------
// A conv with kernel = stride = patch is a Linear over non-overlapping patches: `patchify` gathers x [N, C, D, H, W] into tokens
// [N, T, C pd ph pw] (tokens in row-major order of the patch grid = einops "n l h w d -> n (h w d) l"; features in the order of
// weight.view(E, -1): (c, kd, kh, kw)); the inverse scatters tokens back (the kernel = stride ConvTranspose after its Linear) and
// adds the per-channel bias.
template <typename TS, typename TD, bool INVERSE>
__global__ void __launch_bounds__(256) patchify_kernel(const TS* __restrict__ src, TD* __restrict__ dst, const float* __restrict__ bias,
                                                        int N, int C, int D, int H, int W, int pd, int ph, int pw) {
  const int gd = D / pd, gh = H / ph, gw = W / pw;
  const int64_t F = (int64_t)C * pd * ph * pw, T = (int64_t)gd * gh * gw;
  const int64_t total = (int64_t)N * C * D * H * W;
  // one thread per element of the NC tensor (coalesced on that side; the token side is read / written in runs of pw)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int d = (int)((i / ((int64_t)W * H)) % D);
    const int c = (int)((i / ((int64_t)W * H * D)) % C);
    const int n = (int)(i / ((int64_t)W * H * D * C));
    if (d >= gd * pd || h >= gh * ph || w >= gw * pw) {          // positions the strided conv never reads (floor extents)
      if (INVERSE) dst[i] = from_f32<TD>(0.f);
      continue;
    }
    const int64_t tok = ((int64_t)(d / pd) * gh + h / ph) * gw + w / pw;
    const int64_t f = (((int64_t)c * pd + d % pd) * ph + h % ph) * pw + w % pw;
    const int64_t j = ((int64_t)n * T + tok) * F + f;
    if (INVERSE) dst[i] = from_f32<TD>(to_f32(src[j]) + (bias ? bias[c] : 0.f));
    else dst[j] = from_f32<TD>(to_f32(src[i]));
  }
}

// dst[c * reps + r] = src[c] (the ConvTranspose bias seen by its Linear); inverse: dst[c] = sum_r src[c * reps + r]
__global__ void bias_expand_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int reps, int inverse) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (inverse) {
    float s = 0.f;
    for (int r = 0; r < reps; ++r) s += src[(int64_t)c * reps + r];
    dst[c] = s;
  } else {
    for (int r = 0; r < reps; ++r) dst[(int64_t)c * reps + r] = src[c];
  }
}

// y[o][i] = x[o][i] + add[i]  (positional embedding added to every sample, vit.py:349-352)
template <typename T>
__global__ void __launch_bounds__(256) add_bcast_kernel(const T* __restrict__ x, const float* __restrict__ add, T* __restrict__ y,
                                                         int64_t outer, int64_t inner4) {
  const int64_t total = outer * inner4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4], a[4];
    load4(x + i * 4, v);
    load4(add + (i % inner4) * 4, a);
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] += a[e];
    store4(y + i * 4, v);
  }
}

// out[g][c] = sum_{r < rows} src[(g * rows + r) * ld + c], fixed summation order (deterministic): block = 32 columns x 8 row lanes
template <typename T>
__global__ void __launch_bounds__(256) colsum_groups_kernel(const T* __restrict__ src, int64_t ld, int rows, int64_t Ccols, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int g = blockIdx.y;
  float s = 0.f;
  if (c < Ccols)
    for (int r = ty; r < rows; r += 8) s += to_f32(src[((int64_t)g * rows + r) * ld + c]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < Ccols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    out[(int64_t)g * Ccols + c] = t;
  }
}

// ---- LayerNorm over the last dim of tokens [R, C] (vit.py:140-141, 181-186), one warp per row ----------------------------
// x_in = x + pre_add[row / rows_per_sample] (the diffusion-time embedding added before norm_1, vit.py:179-181; optional)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
    layernorm_fwd_kernel(const TI* __restrict__ x, const float* __restrict__ pre_add, int rows_per_sample, const float* __restrict__ gamma,
                         const float* __restrict__ beta, TO* __restrict__ y, float* __restrict__ stats, int64_t R, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const TI* xr = x + row * C;
  const float* pa = pre_add ? pre_add + (row / rows_per_sample) * C : nullptr;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += to_f32(xr[c]) + (pa ? pa[c] : 0.f);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = to_f32(xr[c]) + (pa ? pa[c] : 0.f) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0 && stats) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
  TO* yr = y + row * C;
  for (int c = lane; c < C; c += 32) {
    const float v = to_f32(xr[c]) + (pa ? pa[c] : 0.f);
    yr[c] = from_f32<TO>((v - mean) * rstd * gamma[c] + beta[c]);
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
    layernorm_bwd_dx_kernel(const TI* __restrict__ x, const float* __restrict__ pre_add, int rows_per_sample, const float* __restrict__ gamma,
                            const float* __restrict__ stats, const TO* __restrict__ dy, TI* __restrict__ dx, int64_t R, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const TI* xr = x + row * C;
  const TO* gr = dy + row * C;
  const float* pa = pre_add ? pre_add + (row / rows_per_sample) * C : nullptr;
  const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float xh = (to_f32(xr[c]) + (pa ? pa[c] : 0.f) - mean) * rstd;
    const float g = to_f32(gr[c]) * gamma[c];
    s1 += g;
    s2 += g * xh;
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) / (float)C;
  TI* dr = dx + row * C;
  for (int c = lane; c < C; c += 32) {
    const float xh = (to_f32(xr[c]) + (pa ? pa[c] : 0.f) - mean) * rstd;
    const float g = to_f32(gr[c]) * gamma[c];
    dr[c] = from_f32<TI>(rstd * (g - s1 - xh * s2));
  }
}

// partial[split][0][c] = sum_r dy * xhat, partial[split][1][c] = sum_r dy over the rows of one split (fixed order)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
    layernorm_bwd_param_kernel(const TI* __restrict__ x, const float* __restrict__ pre_add, int rows_per_sample, const float* __restrict__ stats,
                               const TO* __restrict__ dy, float* __restrict__ partial, int64_t R, int C, int rows_per_split) {
  __shared__ float red[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = r0 + rows_per_split < R ? r0 + rows_per_split : R;
  float sg = 0.f, sb = 0.f;
  if (c < C)
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const float pa = pre_add ? pre_add[(r / rows_per_sample) * C + c] : 0.f;
      const float xh = (to_f32(x[r * C + c]) + pa - stats[r * 2]) * stats[r * 2 + 1];
      const float g = to_f32(dy[r * C + c]);
      sg += g * xh;
      sb += g;
    }
  red[0][ty][tx] = sg;
  red[1][ty][tx] = sb;
  __syncthreads();
  if (ty < 2 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[ty][k][tx];
    partial[((int64_t)blockIdx.y * 2 + ty) * C + c] = t;
  }
}

static int ew_blocks(int64_t items) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(items, 256), (int64_t)sm_count() * 8)); }

}  // namespace rho

using namespace rho;

extern "C" {

int rho_activation(const void* x, void* y, int dtype, int64_t n, int act, rho_stream_t stream) {
  RHO_REQUIRE(x && y, "activation: null pointer");
  RHO_REQUIRE(act >= RHO_ACT_NONE && act <= RHO_ACT_SYMLOG, "activation: bad activation code %d", act);
  if (n == 0) return RHO_OK;
  cudaStream_t st = as_stream(stream);
  RHO_REQUIRE((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "activation: pointers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  if (dtype == RHO_F32) activation_kernel<float><<<ew_blocks(n4), 256, 0, st>>>((const float*)x, (float*)y, n4, n, act);
  else if (dtype == RHO_BF16) activation_kernel<bf16><<<ew_blocks(n4), 256, 0, st>>>((const bf16*)x, (bf16*)y, n4, n, act);
  else return fail(RHO_ERR_INVALID_ARG, "activation: bad dtype");
  RHO_CHECK_LAUNCH("activation");
  return RHO_OK;
}

int rho_activation_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t n, int act, rho_stream_t stream) {
  RHO_REQUIRE(x && dy && dx, "activation_bwd: null pointer");
  RHO_REQUIRE(act >= RHO_ACT_NONE && act <= RHO_ACT_SYMLOG, "activation_bwd: bad activation code %d", act);
  if (n == 0) return RHO_OK;
  cudaStream_t st = as_stream(stream);
  RHO_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0, "activation_bwd: pointers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  if (dtype == RHO_F32) activation_bwd_kernel<float><<<ew_blocks(n4), 256, 0, st>>>((const float*)x, (const float*)dy, (float*)dx, n4, n, act);
  else if (dtype == RHO_BF16) activation_bwd_kernel<bf16><<<ew_blocks(n4), 256, 0, st>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, n4, n, act);
  else return fail(RHO_ERR_INVALID_ARG, "activation_bwd: bad dtype");
  RHO_CHECK_LAUNCH("activation_bwd");
  return RHO_OK;
}

int rho_convt_weight(const float* src, float* dst, int32_t A, int32_t B, int32_t taps, rho_stream_t stream) {
  RHO_REQUIRE(src && dst && A > 0 && B > 0 && taps > 0, "convt_weight: bad arguments");
  convt_weight_kernel<<<ew_blocks((int64_t)A * B * taps), 256, 0, as_stream(stream)>>>(src, dst, A, B, taps);
  RHO_CHECK_LAUNCH("convt_weight");
  return RHO_OK;
}

int rho_patchify(const void* src, int src_dtype, void* dst, int dst_dtype, const float* bias, int32_t N, int32_t C, int32_t D, int32_t H,
                 int32_t W, int32_t pd, int32_t ph, int32_t pw, int32_t inverse, rho_stream_t stream) {
  RHO_REQUIRE(src && dst, "patchify: null pointer");
  RHO_REQUIRE(N > 0 && C > 0 && D >= pd && H >= ph && W >= pw && pd > 0 && ph > 0 && pw > 0, "patchify: bad geometry");
  RHO_REQUIRE(inverse || !bias, "patchify: the bias belongs to the inverse (token -> field) direction");
  cudaStream_t st = as_stream(stream);
  const int blocks = ew_blocks((int64_t)N * C * D * H * W);
#define RHO_PF(TS, TD)                                                                                                          \
  do {                                                                                                                          \
    if (inverse) patchify_kernel<TS, TD, true><<<blocks, 256, 0, st>>>((const TS*)src, (TD*)dst, bias, N, C, D, H, W, pd, ph, pw); \
    else patchify_kernel<TS, TD, false><<<blocks, 256, 0, st>>>((const TS*)src, (TD*)dst, bias, N, C, D, H, W, pd, ph, pw);      \
  } while (0)
  if (src_dtype == RHO_F32 && dst_dtype == RHO_F32) RHO_PF(float, float);
  else if (src_dtype == RHO_F32 && dst_dtype == RHO_BF16) RHO_PF(float, bf16);
  else if (src_dtype == RHO_BF16 && dst_dtype == RHO_F32) RHO_PF(bf16, float);
  else if (src_dtype == RHO_BF16 && dst_dtype == RHO_BF16) RHO_PF(bf16, bf16);
  else return fail(RHO_ERR_INVALID_ARG, "patchify: bad dtype");
#undef RHO_PF
  RHO_CHECK_LAUNCH("patchify");
  return RHO_OK;
}

int rho_bias_expand(const float* src, float* dst, int32_t C, int32_t reps, int32_t inverse, rho_stream_t stream) {
  RHO_REQUIRE(src && dst && C > 0 && reps > 0, "bias_expand: bad arguments");
  bias_expand_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, as_stream(stream)>>>(src, dst, C, reps, inverse);
  RHO_CHECK_LAUNCH("bias_expand");
  return RHO_OK;
}

int rho_add_bcast(const void* x, int dtype, const float* add, void* y, int64_t outer, int64_t inner, rho_stream_t stream) {
  RHO_REQUIRE(x && add && y, "add_bcast: null pointer");
  RHO_REQUIRE(inner % 4 == 0 && (((uintptr_t)x | (uintptr_t)add | (uintptr_t)y) & 15) == 0, "add_bcast: inner %% 4 and 16-byte alignment required");
  if (outer * inner == 0) return RHO_OK;
  cudaStream_t st = as_stream(stream);
  const int blocks = ew_blocks(outer * inner / 4);
  if (dtype == RHO_F32) add_bcast_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, add, (float*)y, outer, inner / 4);
  else if (dtype == RHO_BF16) add_bcast_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)x, add, (bf16*)y, outer, inner / 4);
  else return fail(RHO_ERR_INVALID_ARG, "add_bcast: bad dtype");
  RHO_CHECK_LAUNCH("add_bcast");
  return RHO_OK;
}

int rho_colsum(const void* src, int dtype, int64_t ld, int32_t groups, int32_t rows, int64_t cols, float* out, rho_stream_t stream) {
  RHO_REQUIRE(src && out && groups > 0 && rows > 0 && cols > 0 && ld >= cols, "colsum: bad arguments");
  RHO_REQUIRE(groups <= 65535 && ceil_div(cols, 32) <= 2147483647LL, "colsum: dims too large");
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)groups);
  if (dtype == RHO_F32) colsum_groups_kernel<float><<<grid, 256, 0, st>>>((const float*)src, ld, rows, cols, out);
  else if (dtype == RHO_BF16) colsum_groups_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)src, ld, rows, cols, out);
  else return fail(RHO_ERR_INVALID_ARG, "colsum: bad dtype");
  RHO_CHECK_LAUNCH("colsum");
  return RHO_OK;
}

int rho_layernorm_fwd(const rho_layernorm_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p && p->x && p->gamma && p->beta && p->y, "layernorm_fwd: null pointer");
  RHO_REQUIRE(p->rows > 0 && p->C > 0 && (!p->pre_add || p->rows_per_sample > 0), "layernorm_fwd: bad shape");
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = (unsigned)ceil_div(p->rows, 8);
#define RHO_LN(TI, TO)                                                                                                                  \
  layernorm_fwd_kernel<TI, TO><<<blocks, 256, 0, st>>>((const TI*)p->x, p->pre_add, p->rows_per_sample, p->gamma, p->beta, (TO*)p->y, \
                                                        p->stats, p->rows, p->C, p->eps)
  if (p->in_dtype == RHO_F32 && p->out_dtype == RHO_F32) RHO_LN(float, float);
  else if (p->in_dtype == RHO_BF16 && p->out_dtype == RHO_BF16) RHO_LN(bf16, bf16);
  else if (p->in_dtype == RHO_F32 && p->out_dtype == RHO_BF16) RHO_LN(float, bf16);
  else if (p->in_dtype == RHO_BF16 && p->out_dtype == RHO_F32) RHO_LN(bf16, float);
  else return fail(RHO_ERR_INVALID_ARG, "layernorm_fwd: bad dtype");
#undef RHO_LN
  RHO_CHECK_LAUNCH("layernorm_fwd");
  return RHO_OK;
}

static int ln_splits(int64_t rows) { return (int)std::max<int64_t>(1, std::min<int64_t>(64, rows / 64)); }

size_t rho_layernorm_bwd_workspace_bytes(const rho_layernorm_params* p) {
  if (!p) return 0;
  return (size_t)ln_splits(p->rows) * 2 * (size_t)p->C * sizeof(float);
}

int rho_layernorm_bwd(const rho_layernorm_params* p, const void* dy, void* dx, float* dgamma, float* dbeta, float* dpre_add, void* workspace,
                      size_t workspace_bytes, rho_stream_t stream) {
  RHO_REQUIRE(p && p->x && p->gamma && p->stats && dy && dx, "layernorm_bwd: null pointer");
  RHO_REQUIRE(p->in_dtype == p->out_dtype, "layernorm_bwd: dy has the dtype of y, dx that of x; both must agree");
  RHO_REQUIRE(!dpre_add || (p->pre_add && p->rows % p->rows_per_sample == 0), "layernorm_bwd: dpre_add without pre_add");
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = (unsigned)ceil_div(p->rows, 8);
  const int splits = ln_splits(p->rows);
  const int rows_per_split = (int)ceil_div(p->rows, splits);
  const bool params = dgamma || dbeta;
  if (params) {
    RHO_REQUIRE(dgamma && dbeta, "layernorm_bwd: dgamma and dbeta come together");
    if (!workspace || workspace_bytes < rho_layernorm_bwd_workspace_bytes(p))
      return fail(RHO_ERR_WORKSPACE, "layernorm_bwd: workspace %zu < %zu", workspace_bytes, rho_layernorm_bwd_workspace_bytes(p));
  }
  float* partial = (float*)workspace;
  dim3 pgrid((unsigned)ceil_div(p->C, 32), (unsigned)splits);
#define RHO_LNB(T)                                                                                                                       \
  do {                                                                                                                                   \
    layernorm_bwd_dx_kernel<T, T><<<blocks, 256, 0, st>>>((const T*)p->x, p->pre_add, p->rows_per_sample, p->gamma, p->stats, (const T*)dy, \
                                                           (T*)dx, p->rows, p->C);                                                       \
    if (params)                                                                                                                          \
      layernorm_bwd_param_kernel<T, T><<<pgrid, 256, 0, st>>>((const T*)p->x, p->pre_add, p->rows_per_sample, p->stats, (const T*)dy,    \
                                                               partial, p->rows, p->C, rows_per_split);                                  \
  } while (0)
  if (p->in_dtype == RHO_F32) RHO_LNB(float);
  else if (p->in_dtype == RHO_BF16) RHO_LNB(bf16);
  else return fail(RHO_ERR_INVALID_ARG, "layernorm_bwd: bad dtype");
#undef RHO_LNB
  RHO_CHECK_LAUNCH("layernorm_bwd");
  if (params) {
    count_launch();
    // partial rows are [split][dgamma | dbeta][C]: two strided column sums in fixed order
    dim3 g((unsigned)ceil_div(p->C, 32), 1);
    colsum_groups_kernel<float><<<g, 256, 0, st>>>(partial, 2 * (int64_t)p->C, splits, p->C, dgamma);
    colsum_groups_kernel<float><<<g, 256, 0, st>>>(partial + p->C, 2 * (int64_t)p->C, splits, p->C, dbeta);
    RHO_CHECK_LAUNCH("layernorm_bwd_reduce");
    count_launch();
  }
  if (dpre_add) {
    const int samples = (int)(p->rows / p->rows_per_sample);
    int rc = rho_colsum(dx, p->in_dtype, p->C, samples, p->rows_per_sample, p->C, dpre_add, stream);
    if (rc != RHO_OK) return rc;
  }
  return RHO_OK;
}

}  // extern "C"
