// DDPM noise-schedule math (rho_diffusion/diffusion/ddpm.py): forward noising, the ancestral update, the
// loss, and the RNG -- each a single fused elementwise (HBM-bound) kernel, with the step index read from
// device memory so that one captured CUDA graph can be replayed for every step of the sampling loop.
//
// Algorithmic bytes per element (fp32): q_sample 12 (rd x0, eps; wr x_t); p_sample_step 16 with external
// z (rd x, eps, z; wr x) or 12 with in-kernel Philox noise; mse 8 (+4 with the gradient).
#include "common.cuh"

namespace rho {

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------------------------------
struct Philox {
  uint32_t c[4];
};
__device__ __forceinline__ Philox philox4x32_10(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox p;
  p.c[0] = c0; p.c[1] = c1; p.c[2] = c2; p.c[3] = c3;
  return p;
}
// 24-bit uniforms in (0,1), Box-Muller; 4 normals per counter
__device__ __forceinline__ void philox_normal4(uint64_t group, uint64_t offset, uint64_t seed, float (&z)[4]) {
  Philox p = philox4x32_10(group, offset, seed);
  const float k = 5.9604644775390625e-08f;  // 2^-24
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float u1 = ((float)(p.c[2 * h] >> 8) + 0.5f) * k;
    float u2 = ((float)(p.c[2 * h + 1] >> 8) + 0.5f) * k;
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    z[2 * h] = r * c;
    z[2 * h + 1] = r * s;
  }
}

__global__ void __launch_bounds__(256) randn_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t offset) {
  const int64_t groups = (n + 3) / 4;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4((uint64_t)g, offset, seed, z);
    if (g * 4 + 3 < n) store4(out + g * 4, z);
    else for (int j = 0; j < 4 && g * 4 + j < n; ++j) out[g * 4 + j] = z[j];
  }
}

__global__ void philox_raw_kernel(uint32_t* __restrict__ out, int64_t groups, uint64_t seed, uint64_t offset) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  Philox p = philox4x32_10((uint64_t)g, offset, seed);
  for (int j = 0; j < 4; ++j) out[g * 4 + j] = p.c[j];
}

// x_t = sqrt(abar[t_n]) * x0 + sqrt(1 - abar[t_n]) * eps      (ddpm.py:122-129)
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       const int64_t* __restrict__ t, const float* __restrict__ alpha_bar, int T,
                                                       int64_t per_sample, int64_t vec_per_sample, float* __restrict__ x_t) {
  const int n = blockIdx.y;
  int64_t tt = t[n];
  tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
  const float ab = alpha_bar[tt];
  const float a = sqrtf(ab), b = sqrtf(1.0f - ab);
  const int64_t base = (int64_t)n * per_sample;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec_per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    if (i * 4 + 3 < per_sample && ((base & 3) == 0)) {
      float x[4], e[4], o[4];
      load4(x0 + base + i * 4, x);
      load4(noise + base + i * 4, e);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = a * x[j] + b * e[j];
      store4(x_t + base + i * 4, o);
    } else {
      for (int j = 0; j < 4 && i * 4 + j < per_sample; ++j) {
        int64_t k = base + i * 4 + j;
        x_t[k] = a * x0[k] + b * noise[k];
      }
    }
  }
}

// one ancestral step (ddpm.py:196-218); see rho_b200.h
__global__ void __launch_bounds__(256) p_sample_step_kernel(rho_p_sample_params p) {
  int64_t t = *p.t;
  if (t <= 0 || t >= p.T) return;  // t == 0: the model output is discarded and x is returned as is
  const float alpha = p.alpha[t], beta = p.beta[t], ab = p.alpha_bar[t];
  const float inv_sqrt_alpha = 1.0f / sqrtf(alpha);
  const float c_eps = beta / sqrtf(1.0f - ab);
  const float c_z = p.noise_scale * sqrtf(beta);
  const bool noisy = t > 1;
  const uint64_t seed = p.seed_dev ? *p.seed_dev : p.seed;
  const int64_t groups = (p.n + 3) / 4;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    const bool full = g * 4 + 3 < p.n;
    if (noisy) {
      if (p.z) {
        if (full) load4(p.z + g * 4, z);
        else for (int j = 0; j < 4 && g * 4 + j < p.n; ++j) z[j] = p.z[g * 4 + j];
      } else {
        philox_normal4((uint64_t)g, (uint64_t)t, seed, z);
      }
    }
    float x[4] = {0, 0, 0, 0}, e[4] = {0, 0, 0, 0};
    if (full) { load4(p.x + g * 4, x); load4(p.eps + g * 4, e); }
    else for (int j = 0; j < 4 && g * 4 + j < p.n; ++j) { x[j] = p.x[g * 4 + j]; e[j] = p.eps[g * 4 + j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = inv_sqrt_alpha * (x[j] - c_eps * e[j]) + c_z * z[j];
      x[j] = fminf(fmaxf(v, p.clamp_lo), p.clamp_hi);
    }
    if (full) store4(p.x + g * 4, x);
    else for (int j = 0; j < 4 && g * 4 + j < p.n; ++j) p.x[g * 4 + j] = x[j];
  }
}

__global__ void advance_timestep_kernel(int64_t* t_scalar, int64_t* t_vec, int B) {
  const int64_t t = *t_scalar - 1;
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) t_vec[i] = t;
  if (threadIdx.x == 0) *t_scalar = t;
}

template <typename TP>
__global__ void __launch_bounds__(256) mse_partial_kernel(const TP* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                          float inv_n, float grad_scale, TP* __restrict__ dpred,
                                                          float* __restrict__ partial) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = to_f32(pred[i]) - target[i];
    s = fmaf(d, d, s);
    if (dpred) dpred[i] = from_f32<TP>(2.0f * d * inv_n * grad_scale);
  }
  __shared__ float ws[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void mse_final_kernel(const float* __restrict__ partial, int blocks, float inv_n, float* __restrict__ loss) {
  // single warp, fixed order
  float s = 0.f;
  for (int i = threadIdx.x; i < blocks; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *loss = s * inv_n;
}

static int mse_blocks(int64_t n) {
  int64_t b = ceil_div(n, 256 * 8);
  if (b > 1024) b = 1024;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, rho_stream_t stream) {
  RHO_REQUIRE(out && n >= 0, "randn: bad args");
  if (n == 0) return RHO_OK;
  int blocks = (int)std::min<int64_t>(ceil_div((n + 3) / 4, 256), (int64_t)sm_count() * 8);
  randn_kernel<<<blocks, 256, 0, as_stream(stream)>>>(out, n, seed, offset);
  RHO_CHECK_LAUNCH("randn");
  return RHO_OK;
}

// test hook: the raw Philox4x32-10 stream (bit-exact against the numpy restatement in oracle/)
int rho_philox_raw(uint32_t* out, int64_t groups, uint64_t seed, uint64_t offset, rho_stream_t stream) {
  RHO_REQUIRE(out && groups > 0, "philox_raw: bad args");
  philox_raw_kernel<<<(unsigned)ceil_div(groups, 128), 128, 0, as_stream(stream)>>>(out, groups, seed, offset);
  RHO_CHECK_LAUNCH("philox_raw");
  return RHO_OK;
}

int rho_q_sample(const float* x0, const float* noise, const int64_t* t, const float* alpha_bar, int32_t T, int32_t N, int64_t per_sample,
                 float* x_t, rho_stream_t stream) {
  RHO_REQUIRE(x0 && noise && t && alpha_bar && x_t && T > 0 && N > 0 && per_sample > 0, "q_sample: bad args");
  RHO_REQUIRE(N <= 65535, "q_sample: N too large");
  int64_t vec = (per_sample + 3) / 4;
  int bx = (int)std::min<int64_t>(ceil_div(vec, 256), std::max<int64_t>(1, (int64_t)sm_count() * 8 / N));
  dim3 grid(bx, N);
  q_sample_kernel<<<grid, 256, 0, as_stream(stream)>>>(x0, noise, t, alpha_bar, T, per_sample, vec, x_t);
  RHO_CHECK_LAUNCH("q_sample");
  return RHO_OK;
}

int rho_p_sample_step(const rho_p_sample_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p && p->x && p->eps && p->t && p->alpha && p->beta && p->alpha_bar && p->T > 0 && p->n > 0, "p_sample_step: bad args");
  int blocks = (int)std::min<int64_t>(ceil_div((p->n + 3) / 4, 256), (int64_t)sm_count() * 8);
  p_sample_step_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*p);
  RHO_CHECK_LAUNCH("p_sample_step");
  return RHO_OK;
}

int rho_advance_timestep(int64_t* t_scalar, int64_t* t_vec, int32_t B, rho_stream_t stream) {
  RHO_REQUIRE(t_scalar && t_vec && B > 0, "advance_timestep: bad args");
  advance_timestep_kernel<<<1, 128, 0, as_stream(stream)>>>(t_scalar, t_vec, B);
  RHO_CHECK_LAUNCH("advance_timestep");
  return RHO_OK;
}

size_t rho_mse_workspace_bytes(int64_t n) { return (size_t)mse_blocks(n) * sizeof(float); }

int rho_mse_loss(const void* pred, int pred_dtype, const float* target, int64_t n, float* loss, void* dpred, float grad_scale,
                 void* workspace, size_t workspace_bytes, rho_stream_t stream) {
  RHO_REQUIRE(pred && target && loss && n > 0, "mse_loss: bad args");
  int blocks = mse_blocks(n);
  if (!workspace || workspace_bytes < (size_t)blocks * sizeof(float)) return fail(RHO_ERR_WORKSPACE, "mse_loss: workspace too small");
  cudaStream_t st = as_stream(stream);
  const float inv_n = 1.0f / (float)n;
  if (pred_dtype == RHO_F32)
    mse_partial_kernel<float><<<blocks, 256, 0, st>>>((const float*)pred, target, n, inv_n, grad_scale, (float*)dpred, (float*)workspace);
  else if (pred_dtype == RHO_BF16)
    mse_partial_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)pred, target, n, inv_n, grad_scale, (bf16*)dpred, (float*)workspace);
  else
    return fail(RHO_ERR_INVALID_ARG, "mse_loss: bad dtype");
  RHO_CHECK_LAUNCH("mse_partial");
  mse_final_kernel<<<1, 32, 0, st>>>((const float*)workspace, blocks, inv_n, loss);
  RHO_CHECK_LAUNCH("mse_final");
  return RHO_OK;
}

}  // extern "C"
