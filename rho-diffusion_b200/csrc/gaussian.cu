// GaussianDiffusionPipeline step kernels (rho_diffusion/diffusion/gaussian_diffusion.py): x0-prediction sampling with
// dynamic thresholding.  All HBM-bound elementwise / histogram work:
//   rho_lincomb2     q_sample with host-built coefficient tables (gaussian_diffusion.py:294-312): 12 B / element
//   rho_abs_quantile per-sample torch.quantile(|x|, q) (gaussian_diffusion.py:400-405) without a sort: three histogram
//                    passes over the bit pattern of |x| (11 + 11 + 10 bits) select the two order statistics that the
//                    linear interpolation needs -- exact, deterministic (integer atomics), 3 x 4 B / element
//   rho_x0_step      threshold -> pred_xstart -> DDIM (eta) or ancestral posterior update in one pass
//                    (gaussian_diffusion.py:338-443, 512-556, 654-702): 12-16 B / element
#include "common.cuh"

namespace rho {

// ---- q_sample with explicit coefficient tables -------------------------------------------------------------------
__global__ void __launch_bounds__(256) lincomb2_kernel(const float* __restrict__ a, const float* __restrict__ b, const int64_t* __restrict__ t,
                                                       const float* __restrict__ ca, const float* __restrict__ cb, int T, int64_t per_sample,
                                                       float* __restrict__ out) {
  const int n = blockIdx.y;
  int64_t tt = t[n];
  tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
  const float fa = ca[tt], fb = cb[tt];
  const int64_t base = (int64_t)n * per_sample;
  const bool vec_ok = ((base & 3) == 0) && ((per_sample & 3) == 0);
  if (vec_ok) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < per_sample; i += (int64_t)gridDim.x * blockDim.x) {
      float x[4], e[4], o[4];
      load4(a + base + i * 4, x);
      load4(b + base + i * 4, e);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = __fadd_rn(__fmul_rn(fa, x[j]), __fmul_rn(fb, e[j]));   // mul, mul, add: no FMA contraction
      store4(out + base + i * 4, o);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (int64_t)gridDim.x * blockDim.x)
      out[base + i] = __fadd_rn(__fmul_rn(fa, a[base + i]), __fmul_rn(fb, b[base + i]));
  }
}

// ---- per-sample quantile of |x| by radix selection ------------------------------------------------------------------
// state[n] = {prefix_lo, rank_lo, prefix_hi, rank_hi}: the bits of the two targets decided so far and their ranks inside
// the set of elements that share those bits.
constexpr int QS_BINS = 2048;

struct QState {
  uint32_t prefix[2];
  uint32_t rank[2];
};

template <int SHIFT, int BITS, int PREFIX_SHIFT>
__global__ void __launch_bounds__(256) quantile_hist_kernel(const float* __restrict__ x, int64_t per_sample, const QState* __restrict__ state,
                                                            uint32_t* __restrict__ hist /* [N][2][QS_BINS] */) {
  __shared__ uint32_t sh[2][QS_BINS];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * QS_BINS; i += blockDim.x) (&sh[0][0])[i] = 0u;
  __syncthreads();
  uint32_t p0 = 0u, p1 = 0u;
  if constexpr (PREFIX_SHIFT < 32) { p0 = state[n].prefix[0] >> PREFIX_SHIFT; p1 = state[n].prefix[1] >> PREFIX_SHIFT; }
  const bool same = p0 == p1;
  const float* xs = x + (int64_t)n * per_sample;
  constexpr uint32_t MASK = (1u << BITS) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = __float_as_uint(xs[i]) & 0x7fffffffu;
    const uint32_t bin = (key >> SHIFT) & MASK;
    if constexpr (PREFIX_SHIFT >= 32) {
      atomicAdd(&sh[0][bin], 1u);
    } else {
      const uint32_t hi = key >> PREFIX_SHIFT;
      if (hi == p0) atomicAdd(&sh[0][bin], 1u);
      if (!same && hi == p1) atomicAdd(&sh[1][bin], 1u);
    }
  }
  __syncthreads();
  uint32_t* h = hist + (int64_t)n * 2 * QS_BINS;
  for (int i = threadIdx.x; i < (1 << BITS); i += blockDim.x) {
    if (sh[0][i]) atomicAdd(&h[i], sh[0][i]);
    if (!same && sh[1][i]) atomicAdd(&h[QS_BINS + i], sh[1][i]);
  }
}

// one CTA per sample: locate the bin that holds each target rank, append its bits to the prefix, clear the histograms
template <int SHIFT, int BITS, int PREFIX_SHIFT>
__global__ void __launch_bounds__(256) quantile_select_kernel(QState* __restrict__ state, uint32_t* __restrict__ hist, uint32_t k_lo, uint32_t k_hi) {
  const int n = blockIdx.x;
  uint32_t* h = hist + (int64_t)n * 2 * QS_BINS;
  __shared__ uint32_t cnt[2][QS_BINS];
  __shared__ uint32_t part[2][256];
  QState st;
  bool same = true;
  if constexpr (PREFIX_SHIFT >= 32) { st.prefix[0] = st.prefix[1] = 0u; st.rank[0] = k_lo; st.rank[1] = k_hi; }
  else { st = state[n]; same = (st.prefix[0] >> PREFIX_SHIFT) == (st.prefix[1] >> PREFIX_SHIFT); }
  constexpr int NB = 1 << BITS;
  constexpr int PER = (NB + 255) / 256;
  for (int i = threadIdx.x; i < NB; i += 256) {
    cnt[0][i] = h[i];
    cnt[1][i] = same ? h[i] : h[QS_BINS + i];
  }
  __syncthreads();
  for (int w = 0; w < 2; ++w) {
    uint32_t s = 0;
    for (int j = 0; j < PER; ++j) { const int i = threadIdx.x * PER + j; if (i < NB) s += cnt[w][i]; }
    part[w][threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int w = threadIdx.x;
    uint32_t r = st.rank[w], acc = 0;
    int seg = 0;
    for (; seg < 255; ++seg) { if (acc + part[w][seg] > r) break; acc += part[w][seg]; }
    int bin = seg * PER;
    const int end = min(NB, bin + PER);
    for (; bin < end - 1; ++bin) { if (acc + cnt[w][bin] > r) break; acc += cnt[w][bin]; }
    state[n].prefix[w] = st.prefix[w] | ((uint32_t)bin << SHIFT);
    state[n].rank[w] = r - acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * QS_BINS; i += 256) h[i] = 0u;
}

// torch.lerp(lo, hi, w) (ATen lerp: two-sided form) and the clamp_(min = floor)
__global__ void quantile_finish_kernel(const QState* __restrict__ state, int N, float weight, float floor_value, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float lo = __uint_as_float(state[n].prefix[0]), hi = __uint_as_float(state[n].prefix[1]);
  const float d = hi - lo;
  float q = weight < 0.5f ? lo + weight * d : hi - d * (1.0f - weight);
  out[n] = fmaxf(q, floor_value);
}

// ---- fused x0-prediction step ------------------------------------------------------------------------------------------
template <typename TM>
__global__ void __launch_bounds__(256) x0_step_kernel(rho_x0_step_params p) {
  const int n = blockIdx.y;
  int64_t t = p.t_scalar ? *p.t_scalar : p.t[n];
  if (t < 0 || t >= p.T) return;
  const float s = p.thresh ? p.thresh[n] : 0.f;
  const float mask = t != 0 ? 1.f : 0.f;
  // per-step scalars, evaluated in fp32 in the reference's order (gaussian_diffusion.py:462-466, 681-700 / 314-336, 547-555)
  float c_rx = 0.f, c_rm1 = 1.f, c_ap = 0.f, c_eps = 0.f, c_noise = 0.f, c1 = 0.f, c2 = 0.f;
  if (p.mode == 0) {
    const float ab = p.alphas_cumprod[t], abp = p.alphas_cumprod_prev[t];
    const float sigma = __fmul_rn(__fmul_rn(p.eta, sqrtf((1.f - abp) / (1.f - ab))), sqrtf(1.f - ab / abp));
    c_rx = p.sqrt_recip_alphas_cumprod[t];
    c_rm1 = p.sqrt_recipm1_alphas_cumprod[t];
    c_ap = sqrtf(abp);
    c_eps = sqrtf(__fsub_rn(1.f - abp, __fmul_rn(sigma, sigma)));
    c_noise = mask * sigma;
  } else {
    c1 = p.posterior_mean_coef1[t];
    c2 = p.posterior_mean_coef2[t];
    c_noise = mask * expf(0.5f * p.model_log_variance[t]);
  }
  const int64_t base = (int64_t)n * p.per_sample;
  const TM* mo = (const TM*)p.model_out + base;
  float* x = p.x + base;
  const float* z = p.noise ? p.noise + base : nullptr;
  float* px = p.pred_xstart ? p.pred_xstart + base : nullptr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    float pred = to_f32(mo[i]);
    if (p.thresh) pred = fminf(fmaxf(pred, -s), s) / s;
    const float xv = x[i];
    float out;
    if (p.mode == 0) {
      const float eps = __fsub_rn(__fmul_rn(c_rx, xv), pred) / c_rm1;
      out = __fadd_rn(__fmul_rn(pred, c_ap), __fmul_rn(c_eps, eps));
    } else {
      out = __fadd_rn(__fmul_rn(c1, pred), __fmul_rn(c2, xv));
    }
    if (z && c_noise != 0.f) out = __fadd_rn(out, __fmul_rn(c_noise, z[i]));
    x[i] = out;
    if (px) px[i] = pred;
  }
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_lincomb2(const float* a, const float* b, const int64_t* t, const float* ca, const float* cb, int32_t T, int32_t N, int64_t per_sample,
                 float* out, rho_stream_t stream) {
  RHO_REQUIRE(a && b && t && ca && cb && out && T > 0 && N > 0 && per_sample > 0, "lincomb2: bad args");
  RHO_REQUIRE(N <= 65535, "lincomb2: N too large");
  int bx = (int)std::min<int64_t>(ceil_div(ceil_div(per_sample, 4), 256), std::max<int64_t>(1, (int64_t)sm_count() * 8 / N));
  lincomb2_kernel<<<dim3(bx, N), 256, 0, as_stream(stream)>>>(a, b, t, ca, cb, T, per_sample, out);
  RHO_CHECK_LAUNCH("lincomb2");
  return RHO_OK;
}

size_t rho_abs_quantile_workspace_bytes(int32_t N) { return (size_t)N * (2 * QS_BINS * sizeof(uint32_t) + sizeof(QState)); }

int rho_abs_quantile(const float* x, int32_t N, int64_t per_sample, int64_t k_lo, int64_t k_hi, float weight, float floor_value, float* out,
                     void* workspace, size_t workspace_bytes, rho_stream_t stream) {
  RHO_REQUIRE(x && out && N > 0 && per_sample > 0, "abs_quantile: bad args");
  RHO_REQUIRE(N <= 65535 && per_sample < (1ll << 32), "abs_quantile: size out of range");
  RHO_REQUIRE(k_lo >= 0 && k_lo <= k_hi && k_hi < per_sample, "abs_quantile: ranks out of range");
  if (!workspace || workspace_bytes < rho_abs_quantile_workspace_bytes(N)) return fail(RHO_ERR_WORKSPACE, "abs_quantile: workspace too small");
  cudaStream_t st = as_stream(stream);
  uint32_t* hist = (uint32_t*)workspace;
  QState* state = (QState*)((char*)workspace + (size_t)N * 2 * QS_BINS * sizeof(uint32_t));
  cudaError_t e = cudaMemsetAsync(workspace, 0, rho_abs_quantile_workspace_bytes(N), st);
  if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "abs_quantile: memset: %s", cudaGetErrorString(e));
  int bx = (int)std::min<int64_t>(ceil_div(per_sample, 256 * 8), std::max<int64_t>(1, (int64_t)sm_count() * 4 / N));
  dim3 grid(bx, N);
  // keys are the 31 magnitude bits: 11 (bits 30..20) + 11 (19..9) + 9 (8..0)
  quantile_hist_kernel<20, 11, 32><<<grid, 256, 0, st>>>(x, per_sample, state, hist);
  RHO_CHECK_LAUNCH("quantile_hist_a");
  quantile_select_kernel<20, 11, 32><<<N, 256, 0, st>>>(state, hist, (uint32_t)k_lo, (uint32_t)k_hi);
  RHO_CHECK_LAUNCH("quantile_select_a");
  quantile_hist_kernel<9, 11, 20><<<grid, 256, 0, st>>>(x, per_sample, state, hist);
  RHO_CHECK_LAUNCH("quantile_hist_b");
  quantile_select_kernel<9, 11, 20><<<N, 256, 0, st>>>(state, hist, 0u, 0u);
  RHO_CHECK_LAUNCH("quantile_select_b");
  quantile_hist_kernel<0, 9, 9><<<grid, 256, 0, st>>>(x, per_sample, state, hist);
  RHO_CHECK_LAUNCH("quantile_hist_c");
  quantile_select_kernel<0, 9, 9><<<N, 256, 0, st>>>(state, hist, 0u, 0u);
  RHO_CHECK_LAUNCH("quantile_select_c");
  quantile_finish_kernel<<<(N + 127) / 128, 128, 0, st>>>(state, N, weight, floor_value, out);
  RHO_CHECK_LAUNCH("quantile_finish");
  return RHO_OK;
}

int rho_x0_step(const rho_x0_step_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p && p->model_out && p->x && p->N > 0 && p->per_sample > 0 && p->T > 0, "x0_step: bad args");
  RHO_REQUIRE(p->t || p->t_scalar, "x0_step: no timestep source");
  RHO_REQUIRE(p->N <= 65535, "x0_step: N too large");
  if (p->mode == 0) RHO_REQUIRE(p->alphas_cumprod && p->alphas_cumprod_prev && p->sqrt_recip_alphas_cumprod && p->sqrt_recipm1_alphas_cumprod,
                                "x0_step: DDIM tables missing");
  else if (p->mode == 1) RHO_REQUIRE(p->posterior_mean_coef1 && p->posterior_mean_coef2 && p->model_log_variance,
                                     "x0_step: posterior tables missing");
  else return fail(RHO_ERR_INVALID_ARG, "x0_step: mode must be 0 (DDIM) or 1 (ancestral)");
  int bx = (int)std::min<int64_t>(ceil_div(p->per_sample, 256 * 4), std::max<int64_t>(1, (int64_t)sm_count() * 8 / p->N));
  dim3 grid(bx, p->N);
  if (p->model_dtype == RHO_F32) x0_step_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(*p);
  else if (p->model_dtype == RHO_BF16) x0_step_kernel<bf16><<<grid, 256, 0, as_stream(stream)>>>(*p);
  else return fail(RHO_ERR_INVALID_ARG, "x0_step: bad model dtype");
  RHO_CHECK_LAUNCH("x0_step");
  return RHO_OK;
}

}  // extern "C"
