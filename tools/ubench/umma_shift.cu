// build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_shift umma_shift.cu   (run on the GPU box: gpurun -- tools/ubench/umma_shift)
// Experiment: does a K-major SWIZZLE_128B UMMA operand descriptor accept a start address shifted by whole 128-byte rows
// (not 1024-aligned), and which base_offset value does it need?  D[128x64] = A[shift:shift+128, 0:64] * B[64x64]^T.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../rho-diffusion_b200/csrc/tc_common.cuh"
using namespace rho;

__global__ void __launch_bounds__(128) k(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ D, int shift, int bo_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 320 rows x 128 B
  uint8_t* sB = smem + 320 * 128;     // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // absolute-address swizzle: 16-byte chunk index ^= (row & 7), row = byte offset / 128 from a 1024-aligned base
  for (int i = tid; i < 320 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sA + row * 128 + ((ch ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(A + row * 64 + ch * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    *reinterpret_cast<uint4*>(sB + row * 128 + ((ch ^ (row & 7)) * 16)) = *reinterpret_cast<const uint4*>(B + row * 64 + ch * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a_addr = smem_u32(sA) + shift * 128;
      uint64_t ad = at_desc_kmajor<128>(a_addr);
      if (bo_mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t bd = at_desc_kmajor<128>(smem_u32(sB));
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, kk != 0);
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int r = tid;
  for (int c = 0; c < 64; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int e = 0; e < 32; ++e) D[r * 64 + c + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

int main() {
  std::vector<float> hA(320 * 64), hB(64 * 64);
  std::vector<bf16> bA(320 * 64), bB(64 * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { bA[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); hA[i] = __bfloat162float(bA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { bB[i] = __float2bfloat16((rand() % 17 - 8) / 8.f); hB[i] = __bfloat162float(bB[i]); }
  bf16 *dA, *dB; float* dD;
  cudaMalloc(&dA, bA.size() * 2); cudaMalloc(&dB, bB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, bA.data(), bA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, bB.data(), bB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 320 * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hD(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 4, 7, 8, 9, 13, 64, 65, 66, 67, 131};
  for (int bo = 0; bo < 2; ++bo)
    for (int s : shifts) {
      cudaMemset(dD, 0, 128 * 64 * 4);
      k<<<1, 128, smem>>>(dA, dB, dD, s, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("bo_mode %d shift %d: CUDA error %s\n", bo, s, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hD.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 64; ++c) {
          double ref = 0;
          for (int kk = 0; kk < 64; ++kk) ref += (double)hA[(r + s) * 64 + kk] * hB[c * 64 + kk];
          maxerr = fmax(maxerr, fabs(ref - hD[r * 64 + c]));
        }
      printf("bo_mode %d shift %3d: max abs err %.4f %s\n", bo, s, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
