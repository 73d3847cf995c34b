// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_attn umma_attn.cu   (run on the GPU box: gpurun -- tools/ubench/umma_attn)
// Experiment: cycles per tcgen05.mma for the MMA sequences of a flash-attention key tile at head dim 64 (M = 128, K = 16):
//   QK  : SS form, N = 128 (A = Q [128 x 64] K-major, B = K tile [128 keys x 64] K-major), 4 K steps per accumulator
//   PV  : N = 64, B = V tile [128 keys x 64 ch] MN-major, 8 K steps per accumulator; A = P either from TENSOR MEMORY (TS form)
//         or from shared memory (SS form)
// in several issue orders (same accumulator back to back, alternating between the two query tiles' accumulators, QK and PV
// interleaved).  One CTA per SM, one issuing thread, operand contents irrelevant.
#include <cstdio>
#include <cstdlib>
#include "../../rho-diffusion_b200/csrc/tc_common.cuh"
using namespace rho;

__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// TMEM map (as attention_tc2.cu): S_i at 128 i, O_i at 256 + 64 i, P_i at 384 + 64 i.  smem: Q0, Q1, K, V, Psm0, Psm1 (16 KB each)
template <int MODE>
__global__ void __launch_bounds__(128) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 6 * 16384 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_qk64 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint8_t *sQ = smem, *sK = smem + 2 * 16384, *sV = smem + 3 * 16384, *sP = smem + 4 * 16384;
    auto qk = [&](int i, int k) {
      umma_f16(tmem + 128u * i, at_desc_kmajor<128>(smem_u32(sQ + i * 16384)) + 2 * k, at_desc_kmajor<128>(smem_u32(sK)) + 2 * k, idesc_qk, 1);
    };
    auto qk64 = [&](int i, int h, int k) {
      umma_f16(tmem + 128u * i + 64u * h, at_desc_kmajor<128>(smem_u32(sQ + i * 16384)) + 2 * k,
               at_desc_kmajor<128>(smem_u32(sK + h * 8192)) + 2 * k, idesc_qk64, 1);
    };
    auto pv_ts = [&](int i, int k) {
      umma_ts(tmem + 256u + 64u * i, tmem + 384u + 64u * i + 8u * k, at_desc_mnmajor(smem_u32(sV + k * 2048), 16384), idesc_pv, 1);
    };
    auto pv_ss = [&](int i, int k) {   // P tile in smem: [128 rows x 128 keys] = two 64-key K blocks of 16 KB... use one (k < 4) twice
      umma_f16(tmem + 256u + 64u * i, at_desc_kmajor<128>(smem_u32(sP + i * 16384)) + 2 * (k & 3), at_desc_mnmajor(smem_u32(sV + k * 2048), 16384),
               idesc_pv, 1);
    };
    const long long t0 = clock64();
    int n = 0;
    for (int it = 0; it < iters; ++it) {
      if (MODE == 0) { for (int i = 0; i < 2; ++i) for (int kk = 0; kk < 4; ++kk) qk(i, kk); n = 8; }            // QK only, 2 accumulators
      if (MODE == 1) { for (int kk = 0; kk < 8; ++kk) pv_ts(0, kk); n = 8; }                                       // PV TS, one accumulator
      if (MODE == 2) { for (int kk = 0; kk < 8; ++kk) { pv_ts(0, kk); pv_ts(1, kk); } n = 16; }                     // PV TS, alternate every MMA
      if (MODE == 3) { for (int g = 0; g < 2; ++g) for (int i = 0; i < 2; ++i) for (int kk = 0; kk < 4; ++kk) pv_ts(i, 4 * g + kk); n = 16; }   // alternate every 4
      if (MODE == 4) { for (int kk = 0; kk < 8; ++kk) pv_ss(0, kk); n = 8; }                                       // PV SS, one accumulator
      if (MODE == 5) { for (int i = 0; i < 2; ++i) { for (int kk = 0; kk < 4; ++kk) qk(i, kk); for (int kk = 0; kk < 8; ++kk) pv_ts(i, kk); } n = 24; }   // v2 order
      if (MODE == 6) { for (int i = 0; i < 2; ++i) for (int kk = 0; kk < 4; ++kk) { qk(i, kk); pv_ts(i, 2 * kk); pv_ts(i, 2 * kk + 1); } n = 24; }        // QK / PV interleaved
      if (MODE == 7) { for (int i = 0; i < 2; ++i) for (int kk = 0; kk < 4; ++kk) { qk(i, kk); pv_ts(i ^ 1, 2 * kk); pv_ts(i ^ 1, 2 * kk + 1); } n = 24; }  // interleaved, other tile's PV
      if (MODE == 8) { for (int i = 0; i < 2; ++i) for (int h = 0; h < 2; ++h) { for (int kk = 0; kk < 4; ++kk) pv_ts(i, 4 * h + kk); for (int kk = 0; kk < 4; ++kk) qk64(i, h, kk); } n = 32; }   // half-tile order
      if (MODE == 9) { for (int kk = 0; kk < 8; ++kk) { pv_ss(0, kk); pv_ss(1, kk); } n = 16; }                     // PV SS alternate
      if (MODE == 10) { for (int kk = 0; kk < 4; ++kk) { qk(0, kk); qk(1, kk); } n = 8; }                           // QK alternate every MMA
      if (MODE == 11) { for (int i = 0; i < 2; ++i) { for (int kk = 0; kk < 4; ++kk) qk(i, kk); for (int kk = 0; kk < 8; ++kk) pv_ss(i, kk); } n = 24; }   // v1-like order (P in smem)
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = n; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int MODE>
static void run(const char* what, double ideal_per_iter) {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 6 * 16384 + 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 512;
  k<MODE><<<148, 128, smem>>>(d, iters);
  k<MODE><<<148, 128, smem>>>(d, iters);
  long long h[2] = {0, 0};
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %2d %-58s: %7.1f clk / MMA, %7.1f clk / iteration (ideal %6.0f = %5.1f%%)  %s\n", MODE, what, (double)h[0] / iters / h[1], (double)h[0] / iters,
         ideal_per_iter, 100.0 * ideal_per_iter / ((double)h[0] / iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<0>("QK SS N=128, 4 K steps x 2 accumulators", 8 * 64.0);
  run<10>("QK SS N=128, alternate accumulator every MMA", 8 * 64.0);
  run<1>("PV TS N=64, 8 K steps, one accumulator", 8 * 32.0);
  run<2>("PV TS N=64, two accumulators alternating every MMA", 16 * 32.0);
  run<3>("PV TS N=64, two accumulators alternating every 4 MMAs", 16 * 32.0);
  run<4>("PV SS N=64 (P in smem), one accumulator", 8 * 32.0);
  run<9>("PV SS N=64 (P in smem), alternating every MMA", 16 * 32.0);
  run<5>("tile pair: QK0 PV0 QK1 PV1 (TS)", 1024.0);
  run<11>("tile pair: QK0 PV0 QK1 PV1 (SS)", 1024.0);
  run<6>("tile pair: QK / PV interleaved 1:2, same query tile", 1024.0);
  run<7>("tile pair: QK / PV interleaved 1:2, other query tile", 1024.0);
  run<8>("tile pair in halves: PV(4) QK64(4) x 4", 1024.0);
  return 0;
}
