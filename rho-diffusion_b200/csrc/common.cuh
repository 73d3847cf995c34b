// Shared device/host helpers for the rho_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/rho_b200.h"

namespace rho {

// ---- error plumbing -------------------------------------------------------------------------
extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_launch_count;

int fail(int code, const char* fmt, ...);
inline void count_launch(int n = 1) { g_launch_count.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define RHO_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return ::rho::fail(RHO_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define RHO_CHECK_LAUNCH(name)                                                              \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) return ::rho::fail(RHO_ERR_CUDA, "%s: %s", name, cudaGetErrorString(_e)); \
    ::rho::count_launch();                                                                  \
  } while (0)

inline cudaStream_t as_stream(rho_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
int sm_count();

// ---- dtype helpers --------------------------------------------------------------------------
using bf16 = __nv_bfloat16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector load/store with conversion to/from fp32 (pointer must be 4-element aligned)
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// ---- activations (registry "activations": SiLU, ReLU, Tanh, Sigmoid, GELU) --------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case RHO_ACT_SILU: return x * sigmoidf_(x);
    case RHO_ACT_RELU: return x > 0.f ? x : 0.f;
    case RHO_ACT_TANH: return tanhf(x);
    case RHO_ACT_SIGMOID: return sigmoidf_(x);
    case RHO_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    case RHO_ACT_SYMLOG: { float t = tanhf(x); return t + t * logf(x * t + 1.0f); }     // activations.py:60-61
    default: return x;
  }
}
// derivative of the activation at pre-activation x
__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case RHO_ACT_SILU: { float s = sigmoidf_(x); return s * (1.0f + x * (1.0f - s)); }
    case RHO_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case RHO_ACT_TANH: { float t = tanhf(x); return 1.0f - t * t; }
    case RHO_ACT_SIGMOID: { float s = sigmoidf_(x); return s * (1.0f - s); }
    case RHO_ACT_GELU: {
      float c = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
      return c + x * 0.3989422804014327f * expf(-0.5f * x * x);
    }
    case RHO_ACT_SYMLOG: {
      const float t = tanhf(x), dt = 1.0f - t * t, u = x * t + 1.0f;
      return dt * (1.0f + logf(u)) + t * (t + x * dt) / u;
    }
    default: return 1.f;
  }
}

// ---- reductions -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Chan et al. parallel merge of (count, mean, M2) moments
struct Moments {
  float n, mean, m2;
};
__device__ __forceinline__ Moments merge(Moments a, Moments b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  Moments r;
  r.n = a.n + b.n;
  float d = b.mean - a.mean;
  float f = b.n / r.n;
  r.mean = a.mean + d * f;
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}

}  // namespace rho
