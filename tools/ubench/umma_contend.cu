// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_contend umma_contend.cu
// Experiment: does what the OTHER warps of a flash-attention CTA do slow the tensor pipe down?  One thread streams the MMA mix of
// the attention backward (4 SS N=64 + 4 SS N=64 + 8 TS N=64 per tile, M = 128, K = 16) while 16 other warps
//   mode 0: are parked at a __syncthreads-free exit (nothing)         mode 1: spin on mbarrier.try_wait (what waiting row owners do)
//   mode 2: stream tcgen05.ld / tcgen05.st over the S / dP columns   mode 3: run ex2 + FMA chains (MUFU / FMA pipes, no memory)
#include <cstdio>
#include <cstdlib>
#include "../../rho-diffusion_b200/csrc/tc_common.cuh"
using namespace rho;

template <int MODE>
__global__ void __launch_bounds__(576) k(long long* out, int iters, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, done, dummy[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += 576) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 1) {
    if (threadIdx.x == 32) {
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc2 = idesc1 | (1u << 16);
      uint8_t *sA1 = smem, *sA2 = smem + 16384, *sB1 = smem + 2 * 16384, *sB2 = smem + 3 * 16384;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        const int j = it % 3;
        if (MODE >= 5) {                       // ring of 6 streamed-operand stages (8 KB tiles), as in the kernel
          sB1 = smem + 2 * 16384 + (it % 6) * 8192;
          sB2 = sB1 + 6 * 8192 > smem + 6 * 16384 - 8192 ? smem + 2 * 16384 : sB1 + 8192;
        }
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ts(tmem + 448u, tmem + (uint32_t)(128 * j + 32 * (kk >> 1) + 8 * (kk & 1)), at_desc_mnmajor(smem_u32(sB2 + kk * 2048), 8192), idesc2, 1);
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ts(tmem + 384u, tmem + (uint32_t)(128 * j + 64 + 32 * (kk >> 1) + 8 * (kk & 1)), at_desc_mnmajor(smem_u32(sB1 + kk * 2048), 8192), idesc2, 1);
        if (MODE == 4 || MODE == 6) umma_commit(&dummy[0]);
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem + (uint32_t)(128 * j), at_desc_kmajor<128>(smem_u32(sA1)) + 2 * kk, at_desc_kmajor<128>(smem_u32(sB1)) + 2 * kk, idesc1, kk != 0);
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem + (uint32_t)(128 * j + 64), at_desc_kmajor<128>(smem_u32(sA2)) + 2 * kk, at_desc_kmajor<128>(smem_u32(sB2)) + 2 * kk, idesc1, kk != 0);
        if (MODE == 4 || MODE == 6) umma_commit(&dummy[1]);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      mbar_arrive(&done);
    }
  } else if (warp >= 2) {
    const uint32_t lane_addr = (uint32_t)(32 * (warp & 3)) << 16;
    if (MODE == 1) {
      mbar_wait(&done, 0);
    } else if (MODE == 2) {
      float acc = 0.f;
      while (!*((volatile uint64_t*)&done) || true) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_addr + (uint32_t)(32 * ((warp >> 2) & 3)), v);
        acc += __uint_as_float(v[3]);
        uint32_t w[16];
        for (int e = 0; e < 16; ++e) w[e] = v[e];
        tmem_st16_async(tmem + lane_addr + 256u + (uint32_t)(16 * ((warp >> 2) & 3)), w);
        tmem_wait_st();
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done)) : "memory");
        if (ok) break;
      }
      if (acc == 123.f) sink[0] = acc;
    } else if (MODE == 3) {
      float x = (float)threadIdx.x * 1e-3f, acc = 0.f;
      for (;;) {
#pragma unroll
        for (int e = 0; e < 32; ++e) { x = ex2_approx(x * 0.5f - 1.0f); acc = fmaf(acc, 0.99f, x); }
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done)) : "memory");
        if (ok) break;
      }
      if (acc == 123.f) sink[0] = acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int MODE>
static void run(const char* what) {
  long long* d; float* sink;
  cudaMalloc(&d, 16); cudaMalloc(&sink, 16);
  const size_t smem = 6 * 16384 + 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 512;
  k<MODE><<<148, 576, smem>>>(d, iters, sink);
  k<MODE><<<148, 576, smem>>>(d, iters, sink);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d %-62s: %6.1f clk / MMA (16 MMAs per tile: %7.1f clk / tile)  %s\n", MODE, what, (double)h / iters / 16, (double)h / iters,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  run<0>("other warps idle");
  run<1>("16 warps spin on mbarrier.try_wait");
  run<2>("16 warps stream tcgen05.ld / tcgen05.st");
  run<3>("16 warps run ex2 + FMA chains");
  run<4>("idle warps, two tcgen05.commit per tile (as the kernel)");
  run<5>("idle warps, streamed operands from a 6-stage ring");
  run<6>("idle warps, commits + ring");
  return 0;
}
