// Error plumbing, ABI version and the layout / cast kernels at the UNet boundary.
#include <stdarg.h>

#include "common.cuh"

namespace rho {

thread_local char g_last_error[512] = {0};
std::atomic<uint64_t> g_launch_count{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// NC[S] -> N[S]C through a 32x32 shared-memory tile (coalesced on both sides).
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) transpose_cs_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int64_t R, int64_t Cc) {
  // src: [batch][R][Cc] -> dst: [batch][Cc][R]
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t tiles_c = (Cc + 31) / 32;
  const int64_t c0 = ((int64_t)blockIdx.x % tiles_c) * 32, r0 = ((int64_t)blockIdx.x / tiles_c) * 32;
  const TS* s = src + b * R * Cc;
  TD* d = dst + b * R * Cc;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = r0 + ty + i * 8, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + i * 8][tx] = to_f32(s[r * Cc + c]);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t c = c0 + ty + i * 8, r = r0 + tx;
    if (r < R && c < Cc) d[c * R + r] = from_f32<TD>(tile[tx][ty + i * 8]);
  }
}

template <typename TA, typename TB, typename TD>
__global__ void add_cast_kernel(const TA* __restrict__ a, const TB* __restrict__ b, TD* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float v = to_f32(a[i]);
    if (b) v += to_f32(b[i]);
    d[i] = from_f32<TD>(v);
  }
}

// 8 elements per thread per iteration (16-byte accesses for bf16, 2 x 16 for fp32): the HBM-bound form used when n % 8 == 0 and
// the pointers are 16-byte aligned -- the gradient sums at the UNet's fan-out points (ops.fork) run through it
template <typename TA, typename TB, typename TD>
__global__ void __launch_bounds__(256) add_cast8_kernel(const TA* __restrict__ a, const TB* __restrict__ b, TD* __restrict__ d, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8], w[4];
    load4(a + i * 8, w); v[0] = w[0]; v[1] = w[1]; v[2] = w[2]; v[3] = w[3];
    load4(a + i * 8 + 4, w); v[4] = w[0]; v[5] = w[1]; v[6] = w[2]; v[7] = w[3];
    if (b) {
      load4(b + i * 8, w); v[0] += w[0]; v[1] += w[1]; v[2] += w[2]; v[3] += w[3];
      load4(b + i * 8 + 4, w); v[4] += w[0]; v[5] += w[1]; v[6] += w[2]; v[7] += w[3];
    }
    const float lo[4] = {v[0], v[1], v[2], v[3]}, hi[4] = {v[4], v[5], v[6], v[7]};
    store4(d + i * 8, lo);
    store4(d + i * 8 + 4, hi);
  }
}

// fp32 -> three bf16 terms x = x1 + x2 + x3 (x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2): 24 mantissa bits), written
// as SIX channel blocks [x1 | x1 | x2 | x1 | x2 | x3] of a channels-last tensor [rows, 6 C].  Contracted with the weight blocks
// [w1 | w2 | w1 | w3 | w2 | w1] it gives x1 w1 + x1 w2 + x2 w1 + x1 w3 + x2 w2 + x3 w1 = x w up to 2^-24-relative terms: an
// fp32-accurate convolution on the bf16 tensor cores with fp32 accumulation (the "split-bf16" engine, rho_b200.h).
// x2 (optional): second source of a channel concat (C = C1 + C2).
__global__ void __launch_bounds__(256) split3_concat_kernel(const float* __restrict__ x, int C1, const float* __restrict__ x2, int C2,
                                                            bf16* __restrict__ out, long long rows) {
  const int C = C1 + C2, cv = C >> 2;
  const long long total = rows * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cv;
    const int c = (int)(i % cv) * 4;
    float v[4];
    if (c < C1) load4(x + row * C1 + c, v); else load4(x2 + row * C2 + (c - C1), v);
    float t1[4], t2[4], t3[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      t1[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
      const float r1 = v[e] - t1[e];
      t2[e] = __bfloat162float(__float2bfloat16_rn(r1));
      t3[e] = __bfloat162float(__float2bfloat16_rn(r1 - t2[e]));
    }
    bf16* o = out + row * 6 * C + c;
    store4(o, t1); store4(o + C, t1); store4(o + 2 * C, t2); store4(o + 3 * C, t1); store4(o + 4 * C, t2); store4(o + 5 * C, t3);
  }
}

// weights [Cout, Cin, taps] fp32 -> [Cout, 6 Cin, taps] fp32 holding the bf16-exact terms in the block order [w1 | w2 | w1 | w3 | w2 | w1]
__global__ void __launch_bounds__(256) split3_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int taps) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int ci = (int)((i / taps) % Cin);
    const long long co = i / ((long long)taps * Cin);
    const float v = w[i];
    const float w1 = __bfloat162float(__float2bfloat16_rn(v));
    const float r1 = v - w1;
    const float w2 = __bfloat162float(__float2bfloat16_rn(r1));
    const float w3 = __bfloat162float(__float2bfloat16_rn(r1 - w2));
    float* o = out + (co * 6 * Cin + ci) * taps + t;
    const long long blk = (long long)Cin * taps;
    o[0] = w1; o[blk] = w2; o[2 * blk] = w1; o[3 * blk] = w3; o[4 * blk] = w2; o[5 * blk] = w1;
  }
}

template <typename T>
__global__ void scale_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ s_dev, float s_host) {
  const float s = (s_dev ? *s_dev : 1.0f) * s_host;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = from_f32<T>(to_f32(x[i]) * s);
}

template <typename TS, typename TD>
static int launch_transpose(const void* src, void* dst, int64_t batch, int64_t R, int64_t Cc, cudaStream_t st) {
  if (batch == 0 || R == 0 || Cc == 0) return RHO_OK;
  if (R == 1 || Cc == 1) {  // pure cast/copy
    int64_t n = batch * R * Cc;
    int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
    add_cast_kernel<TS, TS, TD><<<blocks, 256, 0, st>>>((const TS*)src, (const TS*)nullptr, (TD*)dst, n);
    RHO_CHECK_LAUNCH("cast");
    return RHO_OK;
  }
  int64_t tiles = ceil_div(Cc, 32) * ceil_div(R, 32);
  if (tiles > 2147483647LL || batch > 65535) return fail(RHO_ERR_UNSUPPORTED, "transpose: dims too large");
  dim3 grid((unsigned)tiles, 1, (unsigned)batch);
  transpose_cs_kernel<TS, TD><<<grid, 256, 0, st>>>((const TS*)src, (TD*)dst, R, Cc);
  RHO_CHECK_LAUNCH("transpose");
  return RHO_OK;
}

static int transpose_dispatch(const void* src, int sdt, void* dst, int ddt, int64_t batch, int64_t R, int64_t Cc, cudaStream_t st) {
  if (sdt == RHO_F32 && ddt == RHO_F32) return launch_transpose<float, float>(src, dst, batch, R, Cc, st);
  if (sdt == RHO_F32 && ddt == RHO_BF16) return launch_transpose<float, bf16>(src, dst, batch, R, Cc, st);
  if (sdt == RHO_BF16 && ddt == RHO_F32) return launch_transpose<bf16, float>(src, dst, batch, R, Cc, st);
  if (sdt == RHO_BF16 && ddt == RHO_BF16) return launch_transpose<bf16, bf16>(src, dst, batch, R, Cc, st);
  return fail(RHO_ERR_INVALID_ARG, "transpose: bad dtype");
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_abi_version(void) { return 1; }

const char* rho_error_string(int code) {
  switch (code) {
    case RHO_OK: return "ok";
    case RHO_ERR_INVALID_ARG: return "invalid argument";
    case RHO_ERR_UNSUPPORTED: return "unsupported configuration";
    case RHO_ERR_CUDA: return "CUDA error";
    case RHO_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown error";
  }
}

const char* rho_last_error_message(void) { return g_last_error; }

uint64_t rho_launch_count(void) { return g_launch_count.load(); }

int rho_nc_to_cl(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t N, int64_t C, int64_t S, rho_stream_t stream) {
  RHO_REQUIRE(src && dst, "nc_to_cl: null pointer");
  // src [N][C][S] -> dst [N][S][C]
  return transpose_dispatch(src, src_dtype, dst, dst_dtype, N, C, S, as_stream(stream));
}

int rho_cl_to_nc(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t N, int64_t C, int64_t S, rho_stream_t stream) {
  RHO_REQUIRE(src && dst, "cl_to_nc: null pointer");
  // src [N][S][C] -> dst [N][C][S]
  return transpose_dispatch(src, src_dtype, dst, dst_dtype, N, S, C, as_stream(stream));
}

int rho_scale(void* x, int dtype, int64_t n, const float* scale_dev, float scale_host, rho_stream_t stream) {
  RHO_REQUIRE(x, "scale: null pointer");
  if (n == 0) return RHO_OK;
  int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
  if (dtype == RHO_F32) scale_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>((float*)x, n, scale_dev, scale_host);
  else if (dtype == RHO_BF16) scale_kernel<bf16><<<blocks, 256, 0, as_stream(stream)>>>((bf16*)x, n, scale_dev, scale_host);
  else return fail(RHO_ERR_INVALID_ARG, "scale: bad dtype");
  RHO_CHECK_LAUNCH("scale");
  return RHO_OK;
}

int rho_split3_concat(const float* x, int32_t C1, const float* x2, int32_t C2, void* out, int64_t rows, rho_stream_t stream) {
  RHO_REQUIRE(x && out && rows > 0 && C1 > 0 && C1 % 4 == 0 && C2 >= 0 && C2 % 4 == 0 && (C2 == 0 || x2), "split3_concat: bad args");
  const long long total = rows * ((C1 + C2) >> 2);
  const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 16);
  split3_concat_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, C1, x2, C2, (bf16*)out, rows);
  RHO_CHECK_LAUNCH("split3_concat");
  return RHO_OK;
}

int rho_split3_weight(const float* w, float* out, int32_t Cout, int32_t Cin, int32_t taps, rho_stream_t stream) {
  RHO_REQUIRE(w && out && Cout > 0 && Cin > 0 && taps > 0, "split3_weight: bad args");
  const long long total = (long long)Cout * Cin * taps;
  const int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 16);
  split3_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, out, Cout, Cin, taps);
  RHO_CHECK_LAUNCH("split3_weight");
  return RHO_OK;
}

int rho_add_cast(const void* a, int adt, const void* b, int bdt, void* d, int ddt, int64_t n, rho_stream_t stream) {
  RHO_REQUIRE(a && d, "add_cast: null pointer");
  if (n == 0) return RHO_OK;
  cudaStream_t st = as_stream(stream);
  int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
  const bool vec = n % 8 == 0 && n >= 8192 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)d) & 15) == 0;
  if (vec) blocks = (int)std::min<int64_t>(ceil_div(n / 8, 256), (int64_t)sm_count() * 8);
#define RHO_AC(TA, TB, TD)                                                                                      \
  do {                                                                                                          \
    if (vec) add_cast8_kernel<TA, TB, TD><<<blocks, 256, 0, st>>>((const TA*)a, (const TB*)b, (TD*)d, n / 8);    \
    else add_cast_kernel<TA, TB, TD><<<blocks, 256, 0, st>>>((const TA*)a, (const TB*)b, (TD*)d, n);             \
  } while (0)
  int key = adt * 4 + (b ? bdt : adt) * 2 + ddt;
  switch (key) {
    case 0: RHO_AC(float, float, float); break;
    case 1: RHO_AC(float, float, bf16); break;
    case 2: RHO_AC(float, bf16, float); break;
    case 3: RHO_AC(float, bf16, bf16); break;
    case 4: RHO_AC(bf16, float, float); break;
    case 5: RHO_AC(bf16, float, bf16); break;
    case 6: RHO_AC(bf16, bf16, float); break;
    case 7: RHO_AC(bf16, bf16, bf16); break;
    default: return fail(RHO_ERR_INVALID_ARG, "add_cast: bad dtype");
  }
#undef RHO_AC
  RHO_CHECK_LAUNCH("add_cast");
  return RHO_OK;
}

}  // extern "C"
