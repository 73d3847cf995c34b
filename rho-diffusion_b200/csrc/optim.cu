// torch.optim.AdamW as configured by AbstractDiffusionPipeline.configure_optimizers
// (rho_diffusion/diffusion/abstract_diffusion.py:86-148) fused with ExponentialMovingAverage.update
// (rho_diffusion/ema.py:41-57) over FLAT fp32 buffers: one launch for all 160 M parameters instead of
// ~310 small ones per tensor.  HBM-bound: reads p, g, m, v (, shadow), writes p, m, v (, shadow):
// 28 B/param without EMA, 36 B/param with.
#include "common.cuh"

namespace rho {

struct AdamScalars {
  float decay, step_size, beta1, omb1, beta2, omb2, bc2_sqrt, eps, one_minus_d, grad_scale;
};

__host__ __device__ inline AdamScalars adam_scalars(double lr, double beta1, double beta2, double eps, double wd, int step, float ema_decay_t,
                                                    float grad_scale) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  AdamScalars k;
  k.decay = (float)(1.0 - lr * wd);
  k.step_size = (float)(lr / bc1);
  k.beta1 = (float)beta1; k.omb1 = (float)(1.0 - beta1);
  k.beta2 = (float)beta2; k.omb2 = (float)(1.0 - beta2);
  k.bc2_sqrt = (float)sqrt(bc2);
  k.eps = (float)eps;
  k.one_minus_d = 1.0f - ema_decay_t;
  k.grad_scale = grad_scale;
  return k;
}

// device-resident step: ++step, scalars for this step -> buf + 16 bytes (one thread; doubles, like the host path)
__global__ void adamw_scalars_kernel(rho_adamw_params a) {
  const int step = a.step_dev[0] + 1;
  a.step_dev[0] = step;
  const float d_t = a.shadow ? (float)((double)a.ema_decay * (1.0 - exp(-(double)step / 2000.0))) : 0.f;
  *reinterpret_cast<AdamScalars*>(a.step_dev + 4) = adam_scalars(a.lr, a.beta1, a.beta2, a.eps, a.weight_decay, step, d_t, a.grad_scale);
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(rho_adamw_params a, AdamScalars k_host) {
  const AdamScalars k = a.step_dev ? *reinterpret_cast<const AdamScalars*>(a.step_dev + 4) : k_host;
  const int64_t n4 = a.n >> 2;
  const float one_minus_d = k.one_minus_d;
  // torch.optim.AdamW (single-tensor form): p *= 1 - lr*wd; m.lerp_(g, 1-b1); v = b2*v + (1-b2)*g*g;
  // p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps).  All scalars are formed in double on the host, like torch.
  auto upd = [&](float& p, float g, float& m, float& v) {
    g *= k.grad_scale;
    p *= k.decay;
    m = m + (g - m) * k.omb1;
    v = v * k.beta2 + k.omb2 * g * g;
    const float denom = sqrtf(v) / k.bc2_sqrt + k.eps;
    p -= k.step_size * (m / denom);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float p[4], g[4], m[4], v[4];
    load4(a.p + i * 4, p); load4(a.g + i * 4, g); load4(a.m + i * 4, m); load4(a.v + i * 4, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) upd(p[j], g[j], m[j], v[j]);
    store4(a.p + i * 4, p); store4(a.m + i * 4, m); store4(a.v + i * 4, v);
    if (a.shadow) {
      float s[4];
      load4(a.shadow + i * 4, s);
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] -= one_minus_d * (s[j] - p[j]);
      store4(a.shadow + i * 4, s);
    }
  }
  // tail
  if (blockIdx.x == 0) {
    for (int64_t i = n4 * 4 + threadIdx.x; i < a.n; i += blockDim.x) {
      float p = a.p[i], m = a.m[i], v = a.v[i];
      upd(p, a.g[i], m, v);
      a.p[i] = p; a.m[i] = m; a.v[i] = v;
      if (a.shadow) a.shadow[i] -= one_minus_d * (a.shadow[i] - p);
    }
  }
}

__global__ void __launch_bounds__(256) ema_update_kernel(float* __restrict__ shadow, const float* __restrict__ p, int64_t n,
                                                         float one_minus_d) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = shadow[i];
    shadow[i] = s - one_minus_d * (s - p[i]);
  }
}

}  // namespace rho

using namespace rho;

extern "C" int rho_ema_update(float* shadow, const float* p, int64_t n, float one_minus_decay, rho_stream_t stream) {
  RHO_REQUIRE(shadow && p && n >= 0, "ema_update: bad args");
  if (n == 0) return RHO_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
  ema_update_kernel<<<blocks, 256, 0, as_stream(stream)>>>(shadow, p, n, one_minus_decay);
  RHO_CHECK_LAUNCH("ema_update");
  return RHO_OK;
}

extern "C" int rho_adamw_ema_step(const rho_adamw_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p && p->p && p->g && p->m && p->v && p->n > 0 && (p->step >= 1 || p->step_dev), "adamw_ema_step: bad args");
  AdamScalars k;
  memset(&k, 0, sizeof(k));
  if (p->step_dev) {
    adamw_scalars_kernel<<<1, 1, 0, as_stream(stream)>>>(*p);
    RHO_CHECK_LAUNCH("adamw_scalars");
  } else {
    k = adam_scalars(p->lr, p->beta1, p->beta2, p->eps, p->weight_decay, p->step, p->ema_decay_t, p->grad_scale);
  }
  int blocks = (int)std::min<int64_t>(ceil_div(p->n / 4 + 1, 256), (int64_t)sm_count() * 8);
  adamw_ema_kernel<<<blocks, 256, 0, as_stream(stream)>>>(*p, k);
  RHO_CHECK_LAUNCH("adamw_ema_step");
  return RHO_OK;
}
