// QKVAttention / QKVAttentionLegacy (rho_diffusion/models/unet_v2.py:365-436), flash-style, on CUDA cores.
//
// This is the exact-fp32 engine (and the general fallback for head dims the tcgen05 engine does not
// take).  The T x T weight matrix of the reference is never materialised: online softmax in fp32,
// 64x64 (ch <= 128) or 32x32 (ch <= 256) tiles staged in shared memory, 4x4 / 2x2 register tiles.
// qkv is channels-last [N, T, 3*heads*ch], so each head's Q/K/V slice is a row-major [T, ch] matrix
// with row stride 3*heads*ch -- the natural flash layout (the reference's channel-major [N, C, T]
// layout would have needed a transpose).
//
// Backward: recompute P from the saved log-sum-exp; two deterministic passes (dQ over query tiles,
// dK/dV over key tiles), no atomics.
#include "common.cuh"

namespace rho {

struct AttnOff {
  int q, k, v, row;  // channel offsets of this head's q/k/v slices, row stride (3*heads*ch)
};
__device__ __forceinline__ AttnOff attn_off(int heads, int ch, int order, int h) {
  AttnOff o;
  const int C = heads * ch;
  o.row = 3 * C;
  if (order == 0) { o.q = h * 3 * ch; o.k = o.q + ch; o.v = o.q + 2 * ch; }
  else { o.q = h * ch; o.k = C + h * ch; o.v = 2 * C + h * ch; }
  return o;
}

// cooperative tile load: rows [r0, r0+R) x ch of a [T, row_stride] slice -> smem [R][ld] fp32, zero padded
template <typename T, int R>
__device__ __forceinline__ void load_tile(const T* __restrict__ base, int row_stride, int r0, int Tlen, int ch, float* __restrict__ s,
                                          int ld) {
  const int vec_per_row = ch >> 2;
  for (int i = threadIdx.x; i < R * vec_per_row; i += blockDim.x) {
    int r = i / vec_per_row, c = (i % vec_per_row) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r0 + r < Tlen) load4(base + (long long)(r0 + r) * row_stride + c, v);
    *reinterpret_cast<float4*>(s + r * ld + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}

// TS = tile size (64 or 32); R = TS/16 rows per thread; CJ = max float4 channel groups per thread (ch <= 64*CJ)
template <typename T, int TS, int CJ>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int Tlen,
                                                       int heads, int ch, int order, float scale2) {
  constexpr int R = TS / 16;
  extern __shared__ __align__(16) float smem[];
  const int ld = ch + 4, ldp = TS + 4;
  float* Qs = smem;
  float* Ks = Qs + TS * ld;
  float* Vs = Ks + TS * ld;
  float* Ps = Vs + TS * ld;
  const int n = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const AttnOff o = attn_off(heads, ch, order, h);
  const T* base = qkv + (long long)n * Tlen * o.row;

  load_tile<T, TS>(base + o.q, o.row, q0, Tlen, ch, Qs, ld);
  float m_i[R], l_i[R];
  float acc[R][CJ][4];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    m_i[i] = -INFINITY; l_i[i] = 0.f;
#pragma unroll
    for (int j = 0; j < CJ; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  }
  const int c4 = ch >> 2;
  for (int k0 = 0; k0 < Tlen; k0 += TS) {
    __syncthreads();  // previous iteration's readers of Ks/Vs/Ps are done (also orders the Qs load)
    load_tile<T, TS>(base + o.k, o.row, k0, Tlen, ch, Ks, ld);
    load_tile<T, TS>(base + o.v, o.row, k0, Tlen, ch, Vs, ld);
    __syncthreads();
    float s[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) s[i][j] = 0.f;
    for (int c = 0; c < c4; ++c) {
      float4 qa[R], kb[R];
#pragma unroll
      for (int i = 0; i < R; ++i) qa[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * ld + c * 4);
#pragma unroll
      for (int j = 0; j < R; ++j) kb[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * ld + c * 4);
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) s[i][j] = dot4(qa[i], kb[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        s[i][j] = (k0 + tx + 16 * j < Tlen) ? s[i][j] * scale2 : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mnew = fmaxf(m_i[i], mx);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        float p = expf(s[i][j] - mnew);
        rs += p;
        Ps[(ty + 16 * i) * ldp + tx + 16 * j] = p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      const float alpha = expf(m_i[i] - mnew);
      l_i[i] = l_i[i] * alpha + rs;
      m_i[i] = mnew;
#pragma unroll
      for (int j = 0; j < CJ; ++j) { acc[i][j][0] *= alpha; acc[i][j][1] *= alpha; acc[i][j][2] *= alpha; acc[i][j][3] *= alpha; }
    }
    __syncthreads();
    for (int k = 0; k < TS; k += 4) {
      float4 pr[R];
#pragma unroll
      for (int i = 0; i < R; ++i) pr[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * ldp + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
          const int c = (tx + 16 * j) * 4;
          if (c < ch) {
            const float4 v = *reinterpret_cast<const float4*>(Vs + (k + kk) * ld + c);
#pragma unroll
            for (int i = 0; i < R; ++i) {
              const float p = kk == 0 ? pr[i].x : kk == 1 ? pr[i].y : kk == 2 ? pr[i].z : pr[i].w;
              acc[i][j][0] = fmaf(p, v.x, acc[i][j][0]); acc[i][j][1] = fmaf(p, v.y, acc[i][j][1]);
              acc[i][j][2] = fmaf(p, v.z, acc[i][j][2]); acc[i][j][3] = fmaf(p, v.w, acc[i][j][3]);
            }
          }
        }
      }
    }
  }
  const int C = heads * ch;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int q = q0 + ty + 16 * i;
    if (q >= Tlen) continue;
    const float inv = 1.f / l_i[i];
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
      const int c = (tx + 16 * j) * 4;
      if (c < ch) {
        float v[4] = {acc[i][j][0] * inv, acc[i][j][1] * inv, acc[i][j][2] * inv, acc[i][j][3] * inv};
        store4(out + ((long long)n * Tlen + q) * C + h * ch + c, v);
      }
    }
    if (lse && tx == 0) lse[((long long)n * heads + h) * Tlen + q] = m_i[i] + logf(l_i[i]);
  }
}

// D[n][h][q] = sum_c dO[n][q][h*ch+c] * O[n][q][h*ch+c]
template <typename T>
__global__ void attn_bwd_delta_kernel(const T* __restrict__ out, const T* __restrict__ dout, float* __restrict__ delta, int Tlen, int heads,
                                      int ch, long long total) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over n*heads*T
  if (i >= total) return;
  int q = (int)(i % Tlen);
  int h = (int)((i / Tlen) % heads);
  long long n = i / ((long long)Tlen * heads);
  const long long base = (n * Tlen + q) * (long long)(heads * ch) + h * ch;
  float s = 0.f;
  for (int c = 0; c < ch; c += 4) {
    float a[4], b[4];
    load4(out + base + c, a);
    load4(dout + base + c, b);
    s += a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  }
  delta[i] = s;
}

// MODE 0: dQ for one query tile (loop over key tiles);  MODE 1: dK, dV for one key tile (loop over query tiles)
template <typename T, int TS, int CJ, int MODE>
__global__ void __launch_bounds__(256) attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ lse,
                                                       const float* __restrict__ delta, T* __restrict__ dqkv, int Tlen, int heads, int ch,
                                                       int order, float scale2) {
  constexpr int R = TS / 16;
  extern __shared__ __align__(16) float smem[];
  const int ld = ch + 4, ldp = TS + 4;
  float* Qs = smem;
  float* Ks = Qs + TS * ld;
  float* Vs = Ks + TS * ld;
  float* dOs = Vs + TS * ld;
  float* Ps = dOs + TS * ld;    // MODE 0: unused; MODE 1: P^T [k][q]
  float* dSs = Ps + TS * ldp;   // MODE 0: dS [q][k]; MODE 1: dS^T [k][q]
  float* lse_s = dSs + TS * ldp;
  float* del_s = lse_s + TS;
  const int n = blockIdx.z, h = blockIdx.y, t0 = blockIdx.x * TS;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const AttnOff o = attn_off(heads, ch, order, h);
  const int C = heads * ch;
  const T* base = qkv + (long long)n * Tlen * o.row;
  const T* dob = dout + (long long)n * Tlen * C + h * ch;
  const float* lse_b = lse + ((long long)n * heads + h) * Tlen;
  const float* del_b = delta + ((long long)n * heads + h) * Tlen;
  const int c4 = ch >> 2;

  float acc0[R][CJ][4], acc1[R][CJ][4];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < CJ; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc0[i][j][e] = 0.f; acc1[i][j][e] = 0.f; }

  if (MODE == 0) {
    load_tile<T, TS>(base + o.q, o.row, t0, Tlen, ch, Qs, ld);
    load_tile<T, TS>(dob, C, t0, Tlen, ch, dOs, ld);
    for (int i = threadIdx.x; i < TS; i += blockDim.x) {
      lse_s[i] = (t0 + i < Tlen) ? lse_b[t0 + i] : 0.f;
      del_s[i] = (t0 + i < Tlen) ? del_b[t0 + i] : 0.f;
    }
  } else {
    load_tile<T, TS>(base + o.k, o.row, t0, Tlen, ch, Ks, ld);
    load_tile<T, TS>(base + o.v, o.row, t0, Tlen, ch, Vs, ld);
  }

  for (int u0 = 0; u0 < Tlen; u0 += TS) {
    __syncthreads();
    const int q0 = MODE == 0 ? t0 : u0;
    const int k0 = MODE == 0 ? u0 : t0;
    if (MODE == 0) {
      load_tile<T, TS>(base + o.k, o.row, u0, Tlen, ch, Ks, ld);
      load_tile<T, TS>(base + o.v, o.row, u0, Tlen, ch, Vs, ld);
    } else {
      load_tile<T, TS>(base + o.q, o.row, u0, Tlen, ch, Qs, ld);
      load_tile<T, TS>(dob, C, u0, Tlen, ch, dOs, ld);
      for (int i = threadIdx.x; i < TS; i += blockDim.x) {
        lse_s[i] = (u0 + i < Tlen) ? lse_b[u0 + i] : 0.f;
        del_s[i] = (u0 + i < Tlen) ? del_b[u0 + i] : 0.f;
      }
    }
    __syncthreads();
    float s[R][R], dp[R][R];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) { s[i][j] = 0.f; dp[i][j] = 0.f; }
    for (int c = 0; c < c4; ++c) {
      float4 qa[R], da[R], kb[R], vb[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        qa[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * ld + c * 4);
        da[i] = *reinterpret_cast<const float4*>(dOs + (ty + 16 * i) * ld + c * 4);
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        kb[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * ld + c * 4);
        vb[j] = *reinterpret_cast<const float4*>(Vs + (tx + 16 * j) * ld + c * 4);
      }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) { s[i][j] = dot4(qa[i], kb[j], s[i][j]); dp[i][j] = dot4(da[i], vb[j], dp[i][j]); }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int qi = ty + 16 * i;
      const bool qv = q0 + qi < Tlen;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int kj = tx + 16 * j;
        const bool valid = qv && (k0 + kj < Tlen);
        const float p = valid ? expf(s[i][j] * scale2 - lse_s[qi]) : 0.f;
        const float ds = p * (dp[i][j] - del_s[qi]) * scale2;
        if (MODE == 0) dSs[qi * ldp + kj] = ds;
        else { Ps[kj * ldp + qi] = p; dSs[kj * ldp + qi] = ds; }
      }
    }
    __syncthreads();
    // accumulate: MODE 0: dQ[q][c] += dS[q][k] K[k][c];  MODE 1: dV[k][c] += P^T[k][q] dO[q][c], dK[k][c] += dS^T[k][q] Q[q][c]
    for (int u = 0; u < TS; u += 4) {
      float4 a0[R], a1[R];
#pragma unroll
      for (int i = 0; i < R; ++i) {
        a0[i] = *reinterpret_cast<const float4*>(dSs + (ty + 16 * i) * ldp + u);
        if (MODE == 1) a1[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * ldp + u);
      }
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
#pragma unroll
        for (int j = 0; j < CJ; ++j) {
          const int c = (tx + 16 * j) * 4;
          if (c < ch) {
            const float4 b0 = *reinterpret_cast<const float4*>((MODE == 0 ? Ks : Qs) + (u + uu) * ld + c);
            float4 b1;
            if (MODE == 1) b1 = *reinterpret_cast<const float4*>(dOs + (u + uu) * ld + c);
#pragma unroll
            for (int i = 0; i < R; ++i) {
              const float w0 = uu == 0 ? a0[i].x : uu == 1 ? a0[i].y : uu == 2 ? a0[i].z : a0[i].w;
              acc0[i][j][0] = fmaf(w0, b0.x, acc0[i][j][0]); acc0[i][j][1] = fmaf(w0, b0.y, acc0[i][j][1]);
              acc0[i][j][2] = fmaf(w0, b0.z, acc0[i][j][2]); acc0[i][j][3] = fmaf(w0, b0.w, acc0[i][j][3]);
              if (MODE == 1) {
                const float w1 = uu == 0 ? a1[i].x : uu == 1 ? a1[i].y : uu == 2 ? a1[i].z : a1[i].w;
                acc1[i][j][0] = fmaf(w1, b1.x, acc1[i][j][0]); acc1[i][j][1] = fmaf(w1, b1.y, acc1[i][j][1]);
                acc1[i][j][2] = fmaf(w1, b1.z, acc1[i][j][2]); acc1[i][j][3] = fmaf(w1, b1.w, acc1[i][j][3]);
              }
            }
          }
        }
      }
    }
  }
  T* dbase = dqkv + (long long)n * Tlen * o.row;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int r = t0 + ty + 16 * i;
    if (r >= Tlen) continue;
#pragma unroll
    for (int j = 0; j < CJ; ++j) {
      const int c = (tx + 16 * j) * 4;
      if (c < ch) {
        if (MODE == 0) store4(dbase + (long long)r * o.row + o.q + c, acc0[i][j]);
        else {
          store4(dbase + (long long)r * o.row + o.k + c, acc0[i][j]);
          store4(dbase + (long long)r * o.row + o.v + c, acc1[i][j]);
        }
      }
    }
  }
}

static int attn_validate(const rho_attn_params* p, const char* who) {
  RHO_REQUIRE(p, "%s: null params", who);
  RHO_REQUIRE(p->N > 0 && p->T > 0 && p->heads > 0 && p->ch > 0, "%s: bad dims", who);
  RHO_REQUIRE(p->ch % 4 == 0, "%s: head dim %d must be a multiple of 4", who, p->ch);
  if (p->ch > 256) return fail(RHO_ERR_UNSUPPORTED, "%s: head dim %d > 256", who, p->ch);
  RHO_REQUIRE(p->order == 0 || p->order == 1, "%s: bad order", who);
  RHO_REQUIRE(p->qkv && p->out, "%s: null tensor", who);
  RHO_REQUIRE(p->N <= 65535 && p->heads <= 65535, "%s: grid too large", who);
  return RHO_OK;
}

template <typename KernelT>
static int set_smem(KernelT k, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  return RHO_OK;
}

template <typename T>
static int attn_fwd_t(const rho_attn_params* p, cudaStream_t st) {
  const float scale2 = 1.0f / sqrtf((float)p->ch);
  const int ch = p->ch;
#define RHO_AF(TS, CJ)                                                                                       \
  do {                                                                                                       \
    size_t smem = ((size_t)3 * TS * (ch + 4) + (size_t)TS * (TS + 4)) * sizeof(float);                       \
    int rc = set_smem(attn_fwd_kernel<T, TS, CJ>, smem);                                                     \
    if (rc) return rc;                                                                                       \
    dim3 grid((unsigned)ceil_div(p->T, TS), p->heads, p->N);                                                 \
    attn_fwd_kernel<T, TS, CJ><<<grid, 256, smem, st>>>((const T*)p->qkv, (T*)p->out, p->lse, p->T, p->heads, ch, p->order, scale2); \
  } while (0)
  if (ch <= 64) RHO_AF(64, 1);
  else if (ch <= 128) RHO_AF(64, 2);
  else RHO_AF(32, 4);
#undef RHO_AF
  RHO_CHECK_LAUNCH("attention_fwd");
  return RHO_OK;
}

// attention_bwd_tc.cu
bool attention_bwd_tc_supported(const rho_attn_bwd_params* p);
int attention_bwd_tc(const rho_attn_bwd_params* p, cudaStream_t st);
size_t attention_bwd_tc_workspace(const rho_attn_bwd_params* p);

template <typename T>
static int attn_bwd_t(const rho_attn_bwd_params* p, cudaStream_t st) {
  const rho_attn_params& f = p->fwd;
  const float scale2 = 1.0f / sqrtf((float)f.ch);
  const int ch = f.ch;
  if (attention_bwd_tc_supported(p)) return attention_bwd_tc(p, st);   // bf16, head dim 64 / 128: tcgen05
  float* delta = (float*)p->workspace;
  long long total = (long long)f.N * f.heads * f.T;
  attn_bwd_delta_kernel<T><<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const T*)f.out, (const T*)p->dout, delta, f.T, f.heads, ch, total);
  RHO_CHECK_LAUNCH("attention_bwd_delta");
#define RHO_AB(TS, CJ)                                                                                                   \
  do {                                                                                                                   \
    size_t smem = ((size_t)4 * TS * (ch + 4) + (size_t)2 * TS * (TS + 4) + 2 * TS) * sizeof(float);                      \
    int rc = set_smem(attn_bwd_kernel<T, TS, CJ, 0>, smem);                                                              \
    if (rc) return rc;                                                                                                   \
    rc = set_smem(attn_bwd_kernel<T, TS, CJ, 1>, smem);                                                                  \
    if (rc) return rc;                                                                                                   \
    dim3 grid((unsigned)ceil_div(f.T, TS), f.heads, f.N);                                                                \
    attn_bwd_kernel<T, TS, CJ, 0><<<grid, 256, smem, st>>>((const T*)f.qkv, (const T*)p->dout, f.lse, delta, (T*)p->dqkv, f.T, f.heads, ch, f.order, scale2); \
    attn_bwd_kernel<T, TS, CJ, 1><<<grid, 256, smem, st>>>((const T*)f.qkv, (const T*)p->dout, f.lse, delta, (T*)p->dqkv, f.T, f.heads, ch, f.order, scale2); \
  } while (0)
  if (ch <= 64) RHO_AB(64, 1);
  else if (ch <= 128) RHO_AB(64, 2);
  else RHO_AB(32, 4);
#undef RHO_AB
  RHO_CHECK_LAUNCH("attention_bwd");
  count_launch();
  return RHO_OK;
}

// attention_tc.cu
bool attention_fwd_tc_supported(const rho_attn_params* p);
int attention_fwd_tc(const rho_attn_params* p, cudaStream_t st);

int attention_fwd_simt(const rho_attn_params* p, cudaStream_t st) {
  if (p->dtype == RHO_F32) return attn_fwd_t<float>(p, st);
  if (p->dtype == RHO_BF16) return attn_fwd_t<bf16>(p, st);
  return fail(RHO_ERR_INVALID_ARG, "attention_fwd: bad dtype");
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_attention_fwd(const rho_attn_params* p, rho_stream_t stream) {
  int rc = attn_validate(p, "attention_fwd");
  if (rc) return rc;
  if (attention_fwd_tc_supported(p)) return attention_fwd_tc(p, as_stream(stream));   // bf16, head dim 64 / 128: tcgen05
  return attention_fwd_simt(p, as_stream(stream));
}

size_t rho_attention_bwd_workspace_bytes(const rho_attn_bwd_params* p) {
  if (!p) return 0;
  const size_t simt = (size_t)p->fwd.N * p->fwd.heads * p->fwd.T * sizeof(float);
  const size_t tc = attention_bwd_tc_workspace(p);
  return simt > tc ? simt : tc;
}

int rho_attention_bwd(const rho_attn_bwd_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p, "attention_bwd: null params");
  int rc = attn_validate(&p->fwd, "attention_bwd");
  if (rc) return rc;
  RHO_REQUIRE(p->dout && p->dqkv && p->fwd.lse, "attention_bwd: null tensor (lse from the forward is required)");
  size_t need = rho_attention_bwd_workspace_bytes(p);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "attention_bwd: workspace %zu < %zu", p->workspace_bytes, need);
  if (p->fwd.dtype == RHO_F32) return attn_bwd_t<float>(p, as_stream(stream));
  if (p->fwd.dtype == RHO_BF16) return attn_bwd_t<bf16>(p, as_stream(stream));
  return fail(RHO_ERR_INVALID_ARG, "attention_bwd: bad dtype");
}

}  // extern "C"
