// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o mufu mufu.cu   (run on the GPU box: gpurun -- tools/ubench/mufu)
// Throughput of ex2.approx.ftz.f32 vs ex2.approx.ftz.bf16x2 (elements / clk / SM) and of fma.rn.f32x2 vs fma.rn.f32.
#include <cstdio>
#include <cuda_bf16.h>
#include <cstdint>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
  float a[8];
  uint32_t b[8];
  unsigned long long c[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; b[i] = 0x3c003c00u + i + threadIdx.x; c[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); }
  const unsigned long long m2 = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(0.998f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(b[i]));
      if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(0.999f));
      if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(c[i]) : "l"(m2));
      if (MODE == 4) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(b[i]) + __uint_as_float((uint32_t)c[i]) + __uint_as_float((uint32_t)(c[i] >> 32));
  if (s == 123.456f) out[0] = s;
}
template <int MODE>
void run(const char* name, int elems_per_op) {
  float* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096, blocks = 148 * 8;
  k<MODE><<<blocks, 256>>>(d, 16, 0.1f);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, iters, 0.1f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * 256 * iters * 8;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %.3f ms  %.1f lane-ops/clk/SM (at %d MHz nominal)  %.1f elements/clk/SM\n", name, ms, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000,
         ops * elems_per_op / (ms * 1e-3) / 148 / (clk * 1e3));
}
int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("fma.rn.f32", 1);
  run<3>("fma.rn.f32x2", 2);
  run<4>("tanh.approx.f32", 1);
  return 0;
}
