// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu   (run on the GPU box: gpurun -- tools/ubench/umma_rate)
// Experiment: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N and of the operand row width
// (128-byte rows / SWIZZLE_128B vs 64-byte rows / SWIZZLE_64B), one CTA per SM, one issuing thread, operands resident in
// shared memory (contents irrelevant).  Answers: are the N = 32 / N = 64 MMAs of the small-channel convs operand-fetch bound?
#include <cstdio>
#include <cstdlib>
#include "../../rho-diffusion_b200/csrc/tc_common.cuh"
using namespace rho;

template <int ROW_BYTES>
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}

template <int N, int ROW_BYTES, bool K_OUTER>
__global__ void __launch_bounds__(128) k(long long* out, int iters, int tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (1152 * ROW_BYTES + 256 * ROW_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 1152 * ROW_BYTES);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // like the halo kernel: `tiles` accumulators share one B tile; A starts at an arbitrary row.  K_OUTER: consecutive MMAs go
      // to different accumulators (k step outermost) instead of walking the K steps of one accumulator back to back
      const uint64_t bd = desc_k<ROW_BYTES>(b0);
      if (K_OUTER) {
#pragma unroll
        for (int kk = 0; kk < ROW_BYTES / 32; ++kk)
          for (int i = 0; i < tiles; ++i) {
            const uint64_t ad = desc_k<ROW_BYTES>(a0 + (uint32_t)((it % 3) + i * 128) * ROW_BYTES);
            umma_f16(tmem + (uint32_t)(i * N), ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, true);
          }
      } else {
        for (int i = 0; i < tiles; ++i) {
          const uint64_t ad = desc_k<ROW_BYTES>(a0 + (uint32_t)((it % 3) + i * 128) * ROW_BYTES);
#pragma unroll
          for (int kk = 0; kk < ROW_BYTES / 32; ++kk) umma_f16(tmem + (uint32_t)(i * N), ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, true);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int N, int ROW_BYTES, bool K_OUTER = false>
static void run(int tiles) {
  long long* d;
  cudaMalloc(&d, 8);
  const size_t smem = (size_t)(1152 + 256) * ROW_BYTES + 1024;
  cudaFuncSetAttribute(k<N, ROW_BYTES, K_OUTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 512;
  k<N, ROW_BYTES, K_OUTER><<<148, 128, smem>>>(d, iters, tiles);
  k<N, ROW_BYTES, K_OUTER><<<148, 128, smem>>>(d, iters, tiles);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double mmas = (double)iters * tiles * (ROW_BYTES / 32);
  printf("%s N=%3d row=%3dB tiles=%d: %7.1f clk / MMA   (%5.1f%% of the 8192 FLOP/clk/SM dense peak)  %s\n", K_OUTER ? "k-outer" : "k-inner", N, ROW_BYTES, tiles, h / mmas,
         100.0 * (2.0 * 128 * N * 16) / (h / mmas) / 8192.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<32, 64>(8); run<32, 128>(8); run<64, 64>(4); run<64, 128>(4); run<128, 128>(2); run<256, 128>(1); run<256, 128>(2);
  run<32, 64>(1); run<64, 128>(1);
  run<32, 64, true>(8); run<32, 128, true>(8); run<64, 64, true>(4); run<64, 128, true>(4); run<64, 128, true>(5); run<128, 128, true>(2);
  run<64, 128, true>(2); run<32, 64, true>(2); run<32, 64, true>(4);
  return 0;
}
