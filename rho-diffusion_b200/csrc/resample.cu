// The optional resampling / regularisation paths of the UNet (HBM-bound elementwise kernels, channels-last):
//   rho_resample  avg_pool_nd (Downsample(use_conv=False), unet_v2.py:163-165; ResBlock down, :224-225), nearest upsample
//                 (Upsample(use_conv=False), unet_v2.py:124-131; ResBlock up, :221-222) and their adjoints
//   rho_dropout   nn.Dropout(p) in ResBlock.out_layers (unet_v2.py:239): counter-based Philox mask (seed, offset), the same
//                 call with the same (seed, offset) on dy is its backward
#include "common.cuh"

namespace rho {

struct ResampleGeom {
  int N, Di, Hi, Wi, Do, Ho, Wo, C;
  int fd, fh, fw;
  int mode;    // 0 avg-pool, 1 nearest upsample, 2 sum of children (adjoint of 1), 3 adjoint of avg-pool
};

template <typename T>
__global__ void __launch_bounds__(256) resample_kernel(ResampleGeom g, const T* __restrict__ src, T* __restrict__ dst) {
  const int cv = g.C >> 2;
  const long long total = (long long)g.N * g.Do * g.Ho * g.Wo * cv;
  const float inv = 1.0f / (float)(g.fd * g.fh * g.fw);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 4;
    long long pos = i / cv;
    const int w = (int)(pos % g.Wo); pos /= g.Wo;
    const int h = (int)(pos % g.Ho); pos /= g.Ho;
    const int d = (int)(pos % g.Do);
    const int n = (int)(pos / g.Do);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (g.mode == 0 || g.mode == 2) {
      // output position <- window of fd x fh x fw input positions (mean or sum)
      for (int a = 0; a < g.fd; ++a)
        for (int b = 0; b < g.fh; ++b)
          for (int e = 0; e < g.fw; ++e) {
            float v[4];
            load4(src + ((((long long)n * g.Di + (d * g.fd + a)) * g.Hi + (h * g.fh + b)) * g.Wi + (w * g.fw + e)) * g.C + c, v);
            acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
          }
      if (g.mode == 0) { acc[0] *= inv; acc[1] *= inv; acc[2] *= inv; acc[3] *= inv; }
    } else {
      // output position <- its parent input position (copy, or 1 / window for the avg-pool adjoint; positions the
      // pooling dropped -- odd extents, floor -- receive zero)
      const int pd = d / g.fd, ph = h / g.fh, pw = w / g.fw;
      if (pd < g.Di && ph < g.Hi && pw < g.Wi) {
        load4(src + ((((long long)n * g.Di + pd) * g.Hi + ph) * g.Wi + pw) * g.C + c, acc);
        if (g.mode == 3) { acc[0] *= inv; acc[1] *= inv; acc[2] *= inv; acc[3] *= inv; }
      }
    }
    store4(dst + ((((long long)n * g.Do + d) * g.Ho + h) * g.Wo + w) * g.C + c, acc);
  }
}

__device__ __forceinline__ void philox4(uint64_t ctr, uint64_t offset, uint64_t key, uint32_t (&out)[4]) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = (uint32_t)offset, c3 = (uint32_t)(offset >> 32);
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// y = x * keep / (1 - p), keep_i = [u_i >= p], u_i = 24 random bits of Philox4x32-10(counter = i / 4, offset; key = seed)
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float p, float scale,
                                                      uint64_t seed, uint64_t offset, const int32_t* __restrict__ step_dev) {
  if (step_dev) offset += (uint64_t)(uint32_t)step_dev[0] << 32;      // a new mask per optimizer step under graph replay
  const long long groups = (n + 3) / 4;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4((uint64_t)g, offset, seed, r);
    if (g * 4 + 3 < n) {
      float v[4];
      load4(x + g * 4, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = ((float)(r[j] >> 8) * 5.9604644775390625e-08f >= p) ? v[j] * scale : 0.f;
      store4(y + g * 4, v);
    } else {
      for (int j = 0; j < 4 && g * 4 + j < n; ++j)
        y[g * 4 + j] = from_f32<T>(((float)(r[j] >> 8) * 5.9604644775390625e-08f >= p) ? to_f32(x[g * 4 + j]) * scale : 0.f);
    }
  }
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_resample(const void* src, void* dst, int dtype, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t fd, int32_t fh,
                 int32_t fw, int32_t mode, int32_t Do, int32_t Ho, int32_t Wo, rho_stream_t stream) {
  RHO_REQUIRE(src && dst && N > 0 && D > 0 && H > 0 && W > 0 && C > 0 && Do > 0 && Ho > 0 && Wo > 0, "resample: bad args");
  RHO_REQUIRE(C % 4 == 0, "resample: channels must be a multiple of 4");
  RHO_REQUIRE(fd >= 1 && fh >= 1 && fw >= 1 && mode >= 0 && mode <= 3, "resample: bad factors / mode");
  if (mode == 0 || mode == 2) RHO_REQUIRE(Do * fd <= D && Ho * fh <= H && Wo * fw <= W, "resample: pooled extent exceeds the input");
  ResampleGeom g{N, D, H, W, Do, Ho, Wo, C, fd, fh, fw, mode};
  const long long total = (long long)N * Do * Ho * Wo * (C >> 2);
  const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 16);
  if (dtype == RHO_F32) resample_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(g, (const float*)src, (float*)dst);
  else if (dtype == RHO_BF16) resample_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>(g, (const bf16*)src, (bf16*)dst);
  else return fail(RHO_ERR_INVALID_ARG, "resample: bad dtype");
  RHO_CHECK_LAUNCH("resample");
  return RHO_OK;
}

int rho_dropout(const void* x, void* y, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, const int32_t* step_dev,
                rho_stream_t stream) {
  RHO_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f, "dropout: bad args");
  const int blocks = (int)std::min<long long>(ceil_div((n + 3) / 4, 256), (long long)sm_count() * 16);
  const float scale = 1.0f / (1.0f - p);
  if (dtype == RHO_F32) dropout_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, n, p, scale, seed, offset, step_dev);
  else if (dtype == RHO_BF16) dropout_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>((const bf16*)x, (bf16*)y, n, p, scale, seed, offset, step_dev);
  else return fail(RHO_ERR_INVALID_ARG, "dropout: bad dtype");
  RHO_CHECK_LAUNCH("dropout");
  return RHO_OK;
}

}  // extern "C"
