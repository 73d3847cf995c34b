// GroupNorm32 (+ embedding add / FiLM) + activation, forward and backward, channels-last.
//
// Replaces `normalization(C)` = GroupNorm32(32, C) (rho_diffusion/layers.py:71-74,122-129) fused with
// the activation that follows it and with the conditioning of ResBlock._forward
// (rho_diffusion/models/unet_v2.py:285-292).  The input may be the (never materialised) channel
// concat of two tensors (unet_v2.py:729).
//
// HBM-bound.  Data layout [N][S][C] (C contiguous): a CTA streams a contiguous slab of
// `positions x C` so every load is a fully coalesced 16-byte (fp32) / 8-byte (bf16) vector; each
// thread owns a fixed 4-channel column, so the per-channel parameters live in registers.
// Statistics use shifted sums per thread and Chan's pairwise merge across threads / CTAs in a fixed
// order (deterministic, no atomics).
#include "common.cuh"

namespace rho {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_C = 2048;

struct GnMap {
  int CV;             // 4-channel vectors per position
  int rows_per_pass;  // positions handled concurrently by one CTA pass
  int cv_blocks;      // passes over the channel dimension (1 unless C > 1024)
};

__host__ __device__ inline GnMap gn_map(int C) {
  GnMap m;
  m.CV = C / 4;
  if (m.CV <= GN_THREADS) {
    m.rows_per_pass = GN_THREADS / m.CV;
    m.cv_blocks = 1;
  } else {
    m.rows_per_pass = 1;
    m.cv_blocks = (m.CV + GN_THREADS - 1) / GN_THREADS;
  }
  return m;
}

template <typename T>
__device__ __forceinline__ void gn_load(const T* __restrict__ x, const T* __restrict__ x2, int C1, int C2, int64_t row, int c,
                                        float (&v)[4]) {
  if (c < C1) load4(x + row * C1 + c, v);
  else load4(x2 + row * C2 + (c - C1), v);
}

// ------------------------------------------------------------------------------------------------
// stats
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const T* __restrict__ x, const T* __restrict__ x2,
                                                              const float* __restrict__ pre_add, int pre_stride, int C1, int C2,
                                                              int64_t S, int G, int chunks, float* __restrict__ partial) {
  const int C = C1 + C2;
  const GnMap mp = gn_map(C);
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per;
  const int64_t p1 = min(S, p0 + per);
  const int tid = threadIdx.x;

  __shared__ float ch_n[GN_MAX_C], ch_mean[GN_MAX_C], ch_m2[GN_MAX_C];
  __shared__ float sc_n[1024], sc_mean[1024], sc_m2[1024];

  const T* xb = x + (int64_t)n * S * C1;
  const T* x2b = x2 ? x2 + (int64_t)n * S * C2 : nullptr;

  for (int cvb = 0; cvb < mp.cv_blocks; ++cvb) {
    int row, cv;
    bool active;
    if (mp.cv_blocks == 1) {
      row = tid / mp.CV;
      cv = tid % mp.CV;
      active = row < mp.rows_per_pass;
    } else {
      row = 0;
      cv = cvb * GN_THREADS + tid;
      active = cv < mp.CV;
    }
    const int c = cv * 4;
    float K[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float cnt = 0.f;
    if (active) {
      float pa[4] = {0, 0, 0, 0};
      if (pre_add) {
#pragma unroll
        for (int j = 0; j < 4; ++j) pa[j] = pre_add[(int64_t)n * pre_stride + c + j];
      }
      int64_t p = p0 + row;
      if (p < p1) {
        float v[4];
        gn_load(xb, x2b, C1, C2, p, c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) K[j] = v[j] + pa[j];
      }
      for (; p < p1; p += mp.rows_per_pass) {
        float v[4];
        gn_load(xb, x2b, C1, C2, p, c, v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d = (v[j] + pa[j]) - K[j];
          s1[j] += d;
          s2[j] = fmaf(d, d, s2[j]);
        }
        cnt += 1.f;
      }
    }
    // per-thread channel moments -> scratch[row][channel within pass]
    if (mp.cv_blocks == 1) {
      if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int idx = row * C + c + j;
          float mean = cnt > 0.f ? K[j] + s1[j] / cnt : 0.f;
          float m2 = cnt > 0.f ? fmaxf(s2[j] - s1[j] * s1[j] / cnt, 0.f) : 0.f;
          sc_n[idx] = cnt; sc_mean[idx] = mean; sc_m2[idx] = m2;
        }
      }
      __syncthreads();
      for (int cc = tid; cc < C; cc += GN_THREADS) {
        Moments a = {sc_n[cc], sc_mean[cc], sc_m2[cc]};
        for (int r = 1; r < mp.rows_per_pass; ++r) {
          Moments b = {sc_n[r * C + cc], sc_mean[r * C + cc], sc_m2[r * C + cc]};
          a = merge(a, b);
        }
        ch_n[cc] = a.n; ch_mean[cc] = a.mean; ch_m2[cc] = a.m2;
      }
    } else if (active) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float mean = cnt > 0.f ? K[j] + s1[j] / cnt : 0.f;
        float m2 = cnt > 0.f ? fmaxf(s2[j] - s1[j] * s1[j] / cnt, 0.f) : 0.f;
        ch_n[c + j] = cnt; ch_mean[c + j] = mean; ch_m2[c + j] = m2;
      }
    }
  }
  __syncthreads();
  const int cpg = C / G;
  for (int g = tid; g < G; g += GN_THREADS) {
    Moments a = {ch_n[g * cpg], ch_mean[g * cpg], ch_m2[g * cpg]};
    for (int j = 1; j < cpg; ++j) {
      Moments b = {ch_n[g * cpg + j], ch_mean[g * cpg + j], ch_m2[g * cpg + j]};
      a = merge(a, b);
    }
    float* o = partial + (((int64_t)n * chunks + chunk) * G + g) * 3;
    o[0] = a.n; o[1] = a.mean; o[2] = a.m2;
  }
}

// one warp per (n, group): lanes merge a strided subset of the chunk partials, then a fixed xor-shuffle tree merges the
// 32 lane results (same order every run: deterministic)
__global__ void __launch_bounds__(128) gn_stats_finalize_kernel(const float* __restrict__ partial, int NG, int G, int chunks, float eps,
                                                                float* __restrict__ stats) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= NG) return;
  const int n = i / G, g = i % G;
  Moments a = {0.f, 0.f, 0.f};
  for (int k = lane; k < chunks; k += 32) {
    const float* o = partial + (((int64_t)n * chunks + k) * G + g) * 3;
    Moments b = {o[0], o[1], o[2]};
    a = merge(a, b);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Moments b;
    b.n = __shfl_xor_sync(0xffffffffu, a.n, off);
    b.mean = __shfl_xor_sync(0xffffffffu, a.mean, off);
    b.m2 = __shfl_xor_sync(0xffffffffu, a.m2, off);
    // merge in a lane-independent order so that every lane ends with bit-identical results
    a = (lane & off) ? merge(b, a) : merge(a, b);
  }
  if (lane == 0) {
    const float var = a.n > 0.f ? a.m2 / a.n : 0.f;
    stats[2 * i] = a.mean;
    stats[2 * i + 1] = rsqrtf(var + eps);
  }
}

// ------------------------------------------------------------------------------------------------
// apply
// ------------------------------------------------------------------------------------------------
struct GnCoef {
  float mean[4], rstd[4], gamma[4], beta[4], pa[4], sc[4], sh[4];
};

__device__ __forceinline__ void gn_load_coef(GnCoef& k, int n, int c, int C, int G, const float* __restrict__ stats,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             const float* __restrict__ pre_add, int pre_stride, const float* __restrict__ scale,
                                             const float* __restrict__ shift, int film_stride) {
  const int cpg = C / G;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int g = (c + j) / cpg;
    k.mean[j] = stats[((int64_t)n * G + g) * 2];
    k.rstd[j] = stats[((int64_t)n * G + g) * 2 + 1];
    k.gamma[j] = gamma[c + j];
    k.beta[j] = beta[c + j];
    k.pa[j] = pre_add ? pre_add[(int64_t)n * pre_stride + c + j] : 0.f;
    k.sc[j] = scale ? scale[(int64_t)n * film_stride + c + j] : 0.f;
    k.sh[j] = shift ? shift[(int64_t)n * film_stride + c + j] : 0.f;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(GN_THREADS) gn_apply_kernel(const TI* __restrict__ x, const TI* __restrict__ x2,
                                                              const float* __restrict__ pre_add, int pre_stride,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ scale, const float* __restrict__ shift,
                                                              int film_stride, const float* __restrict__ stats, int C1, int C2,
                                                              int64_t S, int G, int chunks, int act, TO* __restrict__ y) {
  const int C = C1 + C2;
  const GnMap mp = gn_map(C);
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per;
  const int64_t p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const TI* xb = x + (int64_t)n * S * C1;
  const TI* x2b = x2 ? x2 + (int64_t)n * S * C2 : nullptr;
  TO* yb = y + (int64_t)n * S * C;
  const bool film = scale != nullptr;

  for (int cvb = 0; cvb < mp.cv_blocks; ++cvb) {
    int row, cv;
    bool active;
    if (mp.cv_blocks == 1) {
      row = tid / mp.CV; cv = tid % mp.CV; active = row < mp.rows_per_pass;
    } else {
      row = 0; cv = cvb * GN_THREADS + tid; active = cv < mp.CV;
    }
    if (!active) continue;
    const int c = cv * 4;
    GnCoef k;
    gn_load_coef(k, n, c, C, G, stats, gamma, beta, pre_add, pre_stride, scale, shift, film_stride);
    const int step = mp.rows_per_pass;
    int64_t p = p0 + row;
    // 4 rows in flight per thread
    for (; p + 3 * (int64_t)step < p1; p += 4 * (int64_t)step) {
      float v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) gn_load(xb, x2b, C1, C2, p + (int64_t)u * step, c, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = ((v[u][j] + k.pa[j]) - k.mean[j]) * k.rstd[j] * k.gamma[j] + k.beta[j];
          if (film) t = t * (1.f + k.sc[j]) + k.sh[j];
          v[u][j] = apply_act(t, act);
        }
        store4(yb + (p + (int64_t)u * step) * C + c, v[u]);
      }
    }
    for (; p < p1; p += step) {
      float v[4];
      gn_load(xb, x2b, C1, C2, p, c, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = ((v[j] + k.pa[j]) - k.mean[j]) * k.rstd[j] * k.gamma[j] + k.beta[j];
        if (film) t = t * (1.f + k.sc[j]) + k.sh[j];
        v[j] = apply_act(t, act);
      }
      store4(yb + p * C + c, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
//   xhat = ((x + pa) - mean) * rstd ; u = xhat*gamma + beta ; v = film ? u*(1+sc)+sh : u ; y = act(v)
//   dv = dy * act'(v)
//   pass 1: per (n, c) S1 = sum_p dv, S2 = sum_p dv*xhat, S3 = sum_p xhat            (chunk partials)
//   reduce: dshift = S1 ; dscale = gamma*S2 + beta*S1 ; A = gamma*(1+sc)
//           m1[n,g] = sum_{c in g} A*S1 / (cpg*S) ; m2[n,g] = sum_{c in g} A*S2 / (cpg*S)
//           dgamma[c] = sum_n (1+sc)*S2 ; dbeta[c] = sum_n (1+sc)*S1
//           dpre[n,c] = rstd*(A*S1 - S*m1 - m2*S3)
//   pass 2: dx = rstd * (A*dv - m1 - xhat*m2)
// ------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(GN_THREADS) gn_bwd_pass1_kernel(const TI* __restrict__ x, const TI* __restrict__ x2,
                                                                  const TO* __restrict__ dy, const float* __restrict__ pre_add,
                                                                  int pre_stride, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, int film_stride,
                                                                  const float* __restrict__ stats, int C1, int C2, int64_t S, int G,
                                                                  int chunks, int act, float* __restrict__ partial) {
  const int C = C1 + C2;
  const GnMap mp = gn_map(C);
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per;
  const int64_t p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const TI* xb = x + (int64_t)n * S * C1;
  const TI* x2b = x2 ? x2 + (int64_t)n * S * C2 : nullptr;
  const TO* dyb = dy + (int64_t)n * S * C;
  const bool film = scale != nullptr;
  __shared__ float sc1[1024], sc2[1024], sc3[1024];
  float* out = partial + ((int64_t)n * chunks + chunk) * C * 3;

  for (int cvb = 0; cvb < mp.cv_blocks; ++cvb) {
    int row, cv;
    bool active;
    if (mp.cv_blocks == 1) {
      row = tid / mp.CV; cv = tid % mp.CV; active = row < mp.rows_per_pass;
    } else {
      row = 0; cv = cvb * GN_THREADS + tid; active = cv < mp.CV;
    }
    const int c = cv * 4;
    float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0}, a3[4] = {0, 0, 0, 0};
    if (active) {
      GnCoef k;
      gn_load_coef(k, n, c, C, G, stats, gamma, beta, pre_add, pre_stride, scale, shift, film_stride);
      for (int64_t p = p0 + row; p < p1; p += mp.rows_per_pass) {
        float v[4], d[4];
        gn_load(xb, x2b, C1, C2, p, c, v);
        load4(dyb + p * C + c, d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float xh = ((v[j] + k.pa[j]) - k.mean[j]) * k.rstd[j];
          float t = xh * k.gamma[j] + k.beta[j];
          if (film) t = t * (1.f + k.sc[j]) + k.sh[j];
          float dv = d[j] * act_grad(t, act);
          a1[j] += dv;
          a2[j] = fmaf(dv, xh, a2[j]);
          a3[j] += xh;
        }
      }
    }
    if (mp.cv_blocks == 1) {
      if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int idx = row * C + c + j;
          sc1[idx] = a1[j]; sc2[idx] = a2[j]; sc3[idx] = a3[j];
        }
      }
      __syncthreads();
      for (int cc = tid; cc < C; cc += GN_THREADS) {
        float t1 = 0.f, t2 = 0.f, t3 = 0.f;
        for (int r = 0; r < mp.rows_per_pass; ++r) {
          t1 += sc1[r * C + cc]; t2 += sc2[r * C + cc]; t3 += sc3[r * C + cc];
        }
        out[cc * 3] = t1; out[cc * 3 + 1] = t2; out[cc * 3 + 2] = t3;
      }
    } else if (active) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        out[(c + j) * 3] = a1[j]; out[(c + j) * 3 + 1] = a2[j]; out[(c + j) * 3 + 2] = a3[j];
      }
    }
  }
}

// chunk partials [n][chunk][C][3] -> [n][C][3]: CTA = 8 channels x 32 chunk lanes, fixed summation order (deterministic)
__global__ void __launch_bounds__(256) gn_bwd_colsum_kernel(const float* __restrict__ partial, int chunks, int C, float* __restrict__ out) {
  __shared__ float red[8][8][3];
  const int n = blockIdx.y;
  const int cl = threadIdx.x & 7, kl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  float t1 = 0.f, t2 = 0.f, t3 = 0.f;
  if (c < C) {
    for (int k = kl; k < chunks; k += 32) {
      const float* o = partial + (((int64_t)n * chunks + k) * C + c) * 3;
      t1 += o[0]; t2 += o[1]; t3 += o[2];
    }
  }
#pragma unroll
  for (int off = 8; off < 32; off <<= 1) {
    t1 += __shfl_xor_sync(0xffffffffu, t1, off);
    t2 += __shfl_xor_sync(0xffffffffu, t2, off);
    t3 += __shfl_xor_sync(0xffffffffu, t3, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 8) { red[warp][lane][0] = t1; red[warp][lane][1] = t2; red[warp][lane][2] = t3; }
  __syncthreads();
  if (threadIdx.x < 24) {
    const int ch = threadIdx.x / 3, j = threadIdx.x % 3;
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][ch][j];
    if (blockIdx.x * 8 + ch < C) out[((int64_t)n * C + blockIdx.x * 8 + ch) * 3 + j] = s;
  }
}

// one CTA per sample n: reduce chunk partials -> S[n][c][3], group means, per-sample param grads
__global__ void __launch_bounds__(GN_THREADS) gn_bwd_reduce_kernel(const float* __restrict__ partial, int chunks, int C, int G,
                                                                   int64_t S, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const float* __restrict__ scale,
                                                                   int film_stride, const float* __restrict__ stats,
                                                                   float* __restrict__ sums /*[N][C][3]*/,
                                                                   float* __restrict__ gmean /*[N][G][2]*/,
                                                                   float* __restrict__ dscale, float* __restrict__ dshift,
                                                                   int dfilm_stride, float* __restrict__ dpre_add) {
  const int n = blockIdx.x;
  __shared__ float g1[GN_MAX_C], g2[GN_MAX_C];
  __shared__ float m1s[64], m2s[64];
  const int cpg = C / G;
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    float t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float* o = partial + (((int64_t)n * chunks + k) * C + c) * 3;
      t1 += o[0]; t2 += o[1]; t3 += o[2];
    }
    float* so = sums + ((int64_t)n * C + c) * 3;
    so[0] = t1; so[1] = t2; so[2] = t3;
    float sc = scale ? scale[(int64_t)n * film_stride + c] : 0.f;
    float A = gamma[c] * (1.f + sc);
    g1[c] = A * t1;
    g2[c] = A * t2;
    if (dshift) dshift[(int64_t)n * dfilm_stride + c] = t1;
    if (dscale) dscale[(int64_t)n * dfilm_stride + c] = gamma[c] * t2 + beta[c] * t1;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += GN_THREADS) {
    float t1 = 0.f, t2 = 0.f;
    for (int j = 0; j < cpg; ++j) { t1 += g1[g * cpg + j]; t2 += g2[g * cpg + j]; }
    float inv = 1.f / ((float)cpg * (float)S);
    m1s[g] = t1 * inv; m2s[g] = t2 * inv;
    gmean[((int64_t)n * G + g) * 2] = t1 * inv;
    gmean[((int64_t)n * G + g) * 2 + 1] = t2 * inv;
  }
  __syncthreads();
  if (dpre_add) {
    for (int c = threadIdx.x; c < C; c += GN_THREADS) {
      int g = c / cpg;
      float rstd = stats[((int64_t)n * G + g) * 2 + 1];
      const float* so = sums + ((int64_t)n * C + c) * 3;
      dpre_add[(int64_t)n * C + c] = rstd * (g1[c] - (float)S * m1s[g] - m2s[g] * so[2]);
    }
  }
}

// dgamma / dbeta: sums over the batch of the per-sample channel sums.  CTA = 32 channels x 8 sample lanes (sample n goes to lane
// n % 8, so config C's 512-sample batches are not walked by one thread); the lanes meet in shared memory in a fixed order.
__global__ void __launch_bounds__(256) gn_bwd_param_kernel(const float* __restrict__ sums, int N, int C, const float* __restrict__ scale,
                                                           int film_stride, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[2][8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float dg = 0.f, db = 0.f;
  if (c < C) {
    for (int n = ty; n < N; n += 8) {
      float sc = scale ? scale[(int64_t)n * film_stride + c] : 0.f;
      const float* so = sums + ((int64_t)n * C + c) * 3;
      dg += (1.f + sc) * so[1];
      db += (1.f + sc) * so[0];
    }
  }
  red[0][ty][tx] = dg;
  red[1][ty][tx] = db;
  __syncthreads();
  if (ty == 0 && c < C) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { g += red[0][k][tx]; b += red[1][k][tx]; }
    if (dgamma) dgamma[c] = g;
    if (dbeta) dbeta[c] = b;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(GN_THREADS) gn_bwd_pass2_kernel(const TI* __restrict__ x, const TI* __restrict__ x2,
                                                                  const TO* __restrict__ dy, const float* __restrict__ pre_add,
                                                                  int pre_stride, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, const float* __restrict__ scale,
                                                                  const float* __restrict__ shift, int film_stride,
                                                                  const float* __restrict__ stats, const float* __restrict__ gmean,
                                                                  int C1, int C2, int64_t S, int G, int chunks, int act,
                                                                  TI* __restrict__ dx, TI* __restrict__ dx2) {
  const int C = C1 + C2;
  const GnMap mp = gn_map(C);
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per;
  const int64_t p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const TI* xb = x + (int64_t)n * S * C1;
  const TI* x2b = x2 ? x2 + (int64_t)n * S * C2 : nullptr;
  const TO* dyb = dy + (int64_t)n * S * C;
  TI* dxb = dx + (int64_t)n * S * C1;
  TI* dx2b = dx2 ? dx2 + (int64_t)n * S * C2 : nullptr;
  const bool film = scale != nullptr;
  const int cpg = C / G;

  for (int cvb = 0; cvb < mp.cv_blocks; ++cvb) {
    int row, cv;
    bool active;
    if (mp.cv_blocks == 1) {
      row = tid / mp.CV; cv = tid % mp.CV; active = row < mp.rows_per_pass;
    } else {
      row = 0; cv = cvb * GN_THREADS + tid; active = cv < mp.CV;
    }
    if (!active) continue;
    const int c = cv * 4;
    GnCoef k;
    gn_load_coef(k, n, c, C, G, stats, gamma, beta, pre_add, pre_stride, scale, shift, film_stride);
    float m1[4], m2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int g = (c + j) / cpg;
      m1[j] = gmean[((int64_t)n * G + g) * 2];
      m2[j] = gmean[((int64_t)n * G + g) * 2 + 1];
    }
    for (int64_t p = p0 + row; p < p1; p += mp.rows_per_pass) {
      float v[4], d[4], o[4];
      gn_load(xb, x2b, C1, C2, p, c, v);
      load4(dyb + p * C + c, d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float xh = ((v[j] + k.pa[j]) - k.mean[j]) * k.rstd[j];
        float t = xh * k.gamma[j] + k.beta[j];
        float f = 1.f;
        if (film) { f = 1.f + k.sc[j]; t = t * f + k.sh[j]; }
        float dv = d[j] * act_grad(t, act);
        o[j] = k.rstd[j] * (k.gamma[j] * f * dv - m1[j] - xh * m2[j]);
      }
      if (c < C1) store4(dxb + p * C1 + c, o);
      else store4(dx2b + p * C2 + (c - C1), o);
    }
  }
}


// ------------------------------------------------------------------------------------------------
// bf16 fast path: 16-byte vectors (8 channels per thread), 4 rows in flight per thread
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const bf16* __restrict__ p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[2 * j] = __low2float(h[j]); v[2 * j + 1] = __high2float(h[j]); }
}
__device__ __forceinline__ void store8(bf16* __restrict__ p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ void gn_load8(const bf16* __restrict__ x, const bf16* __restrict__ x2, int C1, int C2, int64_t row, int c,
                                         float (&v)[8]) {
  if (c < C1) load8(x + row * C1 + c, v);
  else load8(x2 + row * C2 + (c - C1), v);
}

// One CTA streams a contiguous slab of positions.  Every channel of a group is shifted by the SAME constant (the value
// of the group's first channel at the slab's first position), so the per-thread shifted sums add up across rows and
// across the channels of a group without Chan merges; the merge across slabs happens in gn_stats_finalize_kernel.
__global__ void __launch_bounds__(GN_THREADS, 4) gn_stats8_kernel(const bf16* __restrict__ x, const bf16* __restrict__ x2,
                                                                  const float* __restrict__ pre_add, int pre_stride, int C1, int C2,
                                                                  int64_t S, int G, int chunks, float* __restrict__ partial) {
  const int C = C1 + C2;
  const int CV = C >> 3;
  const int rows = GN_THREADS / CV;
  const int cpg = C / G;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per, p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const int row = tid / CV, cv = tid - row * CV;
  const bool active = row < rows && p0 < p1;
  const int c = cv * 8;
  __shared__ float sm1[2048], sm2[2048];
  const bf16* xb = x + (int64_t)n * S * C1;
  const bf16* x2b = x2 ? x2 + (int64_t)n * S * C2 : nullptr;
  auto value_at = [&](int64_t p, int ch) -> float {
    const float v = ch < C1 ? __bfloat162float(xb[p * C1 + ch]) : __bfloat162float(x2b[p * C2 + (ch - C1)]);
    return v + (pre_add ? pre_add[(int64_t)n * pre_stride + ch] : 0.f);
  };
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (active) {
    float Kp[8];                      // shift minus the per-channel pre-add: d = (v + pa) - K = v - Kp
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gfirst = ((c + j) / cpg) * cpg;
      Kp[j] = value_at(p0, gfirst) - (pre_add ? pre_add[(int64_t)n * pre_stride + c + j] : 0.f);
    }
    const bf16* src = c < C1 ? xb + c : x2b + (c - C1);
    const int64_t stride = c < C1 ? C1 : C2;
    auto accum = [&](const uint4& q) {
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d0 = __uint_as_float(u[j] << 16) - Kp[2 * j];
        const float d1 = __uint_as_float(u[j] & 0xffff0000u) - Kp[2 * j + 1];
        s1[2 * j] += d0; s2[2 * j] = fmaf(d0, d0, s2[2 * j]);
        s1[2 * j + 1] += d1; s2[2 * j + 1] = fmaf(d1, d1, s2[2 * j + 1]);
      }
    };
    int64_t p = p0 + row;
    for (; p + 3 * (int64_t)rows < p1; p += 4 * (int64_t)rows) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(src + (p + (int64_t)u * rows) * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u) accum(q[u]);
    }
    for (; p < p1; p += rows) accum(__ldg(reinterpret_cast<const uint4*>(src + p * stride)));
  }
  if (row < rows) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sm1[row * C + c + j] = s1[j]; sm2[row * C + c + j] = s2[j]; }
  }
  __syncthreads();
  for (int cc = tid; cc < C; cc += GN_THREADS) {
    float a1 = sm1[cc], a2 = sm2[cc];
    for (int r = 1; r < rows; ++r) { a1 += sm1[r * C + cc]; a2 += sm2[r * C + cc]; }
    sm1[cc] = a1; sm2[cc] = a2;
  }
  __syncthreads();
  for (int g = tid; g < G; g += GN_THREADS) {
    float a1 = 0.f, a2 = 0.f;
    for (int j = 0; j < cpg; ++j) { a1 += sm1[g * cpg + j]; a2 += sm2[g * cpg + j]; }
    float* o = partial + (((int64_t)n * chunks + chunk) * G + g) * 3;
    if (p0 < p1) {
      const float cnt = (float)((p1 - p0) * cpg);
      o[0] = cnt;
      o[1] = value_at(p0, g * cpg) + a1 / cnt;
      o[2] = fmaxf(a2 - a1 * a1 / cnt, 0.f);
    } else {
      o[0] = 0.f; o[1] = 0.f; o[2] = 0.f;
    }
  }
}

// SiLU through one MUFU: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx: 2^-11 relative, below bf16 resolution)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS, 4) gn_apply8_kernel(const bf16* __restrict__ x, const bf16* __restrict__ x2,
                                                                  const float* __restrict__ pre_add, int pre_stride,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                                                  int film_stride, const float* __restrict__ stats, int C1, int C2,
                                                                  int64_t S, int G, int chunks, int act, bf16* __restrict__ y) {
  const int C = C1 + C2;
  const int CV = C >> 3;
  const int rows = GN_THREADS / CV;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per, p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const int row = tid / CV, cv = tid - row * CV;
  if (row >= rows) return;
  const int c = cv * 8;
  const int cpg = C / G;
  // y = act(x * a + b): a = rstd*gamma*(1+sc), b = ((pa - mean)*rstd*gamma + beta)*(1+sc) + sh   (halved for the SiLU form)
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    const float mean = stats[((int64_t)n * G + g) * 2], rstd = stats[((int64_t)n * G + g) * 2 + 1];
    const float ga = gamma[c + j] * rstd;
    const float pa = pre_add ? pre_add[(int64_t)n * pre_stride + c + j] : 0.f;
    const float f = scale ? 1.f + scale[(int64_t)n * film_stride + c + j] : 1.f;
    const float sh = shift ? shift[(int64_t)n * film_stride + c + j] : 0.f;
    a[j] = ga * f;
    b[j] = ((pa - mean) * ga + beta[c + j]) * f + sh;
    if (SILU) { a[j] *= 0.5f; b[j] *= 0.5f; }
  }
  const bf16* src = c < C1 ? x + (int64_t)n * S * C1 + c : x2 + (int64_t)n * S * C2 + (c - C1);
  const int64_t stride = c < C1 ? C1 : C2;
  bf16* dst = y + (int64_t)n * S * C + c;
  auto transform = [&](const uint4& q) -> uint4 {
    const uint32_t u[4] = {q.x, q.y, q.z, q.w};
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v0 = fmaf(__uint_as_float(u[j] << 16), a[2 * j], b[2 * j]);
      float v1 = fmaf(__uint_as_float(u[j] & 0xffff0000u), a[2 * j + 1], b[2 * j + 1]);
      if (SILU) {
        v0 = fmaf(v0, tanh_approx(v0), v0);
        v1 = fmaf(v1, tanh_approx(v1), v1);
      } else {
        v0 = apply_act(v0, act);
        v1 = apply_act(v1, act);
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      ow[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return o;
  };
  int64_t p = p0 + row;
  for (; p + 3 * (int64_t)rows < p1; p += 4 * (int64_t)rows) {
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(src + (p + (int64_t)u * rows) * stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(dst + (p + (int64_t)u * rows) * C) = transform(q[u]);
  }
  for (; p < p1; p += rows) *reinterpret_cast<uint4*>(dst + p * C) = transform(__ldg(reinterpret_cast<const uint4*>(src + p * stride)));
}

// ------------------------------------------------------------------------------------------------
// statistics from the per-tile channel sums a conv epilogue left behind (rho_conv_params.stats_partial): one CTA per
// (sample, group); thread t accumulates rows t, t + 256, ... of the group's channels in double, the CTA reduces in a fixed
// tree: deterministic, and no pass over the activations.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_from_partials_kernel(const float* __restrict__ part1, int rows1, int C1,
                                                                     const float* __restrict__ part2, int rows2, int C2, int64_t S, int G,
                                                                     float eps, float* __restrict__ stats) {
  const int n = blockIdx.y, g = blockIdx.x;
  const int cpg = (C1 + C2) / G;
  const int c_lo = g * cpg, c_hi = c_lo + cpg;
  double s1 = 0.0, s2 = 0.0;
  // the group's channels that live in x ...
  {
    const int a = min(c_lo, C1), b = min(c_hi, C1);
    if (b > a) {
      const int w = b - a;
      const float* base = part1 + ((int64_t)n * rows1 * C1 + a) * 2;
      for (int i = threadIdx.x; i < rows1 * w; i += 256) {
        const int r = i / w, c = i - r * w;
        const float2 v = *reinterpret_cast<const float2*>(base + ((int64_t)r * C1 + c) * 2);
        s1 += (double)v.x; s2 += (double)v.y;
      }
    }
  }
  // ... and in the concat partner x2
  {
    const int a = max(c_lo, C1) - C1, b = max(c_hi, C1) - C1;
    if (b > a && part2) {
      const int w = b - a;
      const float* base = part2 + ((int64_t)n * rows2 * C2 + a) * 2;
      for (int i = threadIdx.x; i < rows2 * w; i += 256) {
        const int r = i / w, c = i - r * w;
        const float2 v = *reinterpret_cast<const float2*>(base + ((int64_t)r * C2 + c) * 2);
        s1 += (double)v.x; s2 += (double)v.y;
      }
    }
  }
  __shared__ double sh1[256], sh2[256];
  sh1[threadIdx.x] = s1; sh2[threadIdx.x] = s2;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off) { sh1[threadIdx.x] += sh1[threadIdx.x + off]; sh2[threadIdx.x] += sh2[threadIdx.x + off]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double cnt = (double)S * cpg;
    const double mean = sh1[0] / cnt;
    const double var = fmax(sh2[0] / cnt - mean * mean, 0.0);
    stats[((int64_t)n * G + g) * 2] = (float)mean;
    stats[((int64_t)n * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// small-work variant (few tiles per sample: many-sample 1-D / 2-D batches): one WARP per (sample, group), eight per CTA; lanes
// stride over the (row, channel) pairs, then a fixed xor-shuffle tree in double
__global__ void __launch_bounds__(256) gn_stats_from_partials_warp_kernel(const float* __restrict__ part1, int rows1, int C1,
                                                                          const float* __restrict__ part2, int rows2, int C2, int64_t S,
                                                                          int G, int NG, float eps, float* __restrict__ stats) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= NG) return;
  const int n = i / G, g = i - n * G;
  const int cpg = (C1 + C2) / G;
  const int c_lo = g * cpg, c_hi = c_lo + cpg;
  double s1 = 0.0, s2 = 0.0;
  {
    const int a = min(c_lo, C1), b = min(c_hi, C1);
    if (b > a) {
      const int w = b - a;
      const float* base = part1 + ((int64_t)n * rows1 * C1 + a) * 2;
      for (int k = lane; k < rows1 * w; k += 32) {
        const int r = k / w, c = k - r * w;
        const float2 v = *reinterpret_cast<const float2*>(base + ((int64_t)r * C1 + c) * 2);
        s1 += (double)v.x; s2 += (double)v.y;
      }
    }
  }
  {
    const int a = max(c_lo, C1) - C1, b = max(c_hi, C1) - C1;
    if (b > a && part2) {
      const int w = b - a;
      const float* base = part2 + ((int64_t)n * rows2 * C2 + a) * 2;
      for (int k = lane; k < rows2 * w; k += 32) {
        const int r = k / w, c = k - r * w;
        const float2 v = *reinterpret_cast<const float2*>(base + ((int64_t)r * C2 + c) * 2);
        s1 += (double)v.x; s2 += (double)v.y;
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if (lane == 0) {
    const double cnt = (double)S * cpg;
    const double mean = s1 / cnt;
    const double var = fmax(s2 / cnt - mean * mean, 0.0);
    stats[(int64_t)i * 2] = (float)mean;
    stats[(int64_t)i * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

static int gn_chunks8(int N, int64_t S, int C) {
  const int rows = GN_THREADS / (C >> 3);
  int64_t target = (int64_t)sm_count() * 4;          // one resident wave: 4 CTAs per SM
  int64_t chunks = ceil_div(target, N);
  int64_t max_chunks = ceil_div(S, (int64_t)rows * 8);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 8192) chunks = 8192;
  return (int)chunks;
}

// ------------------------------------------------------------------------------------------------
// backward, bf16 fast path (same thread <-> 8-channel mapping as the forward fast path)
// ------------------------------------------------------------------------------------------------
// d act(t) / dt; SiLU through one MUFU (sigmoid(t) = 0.5 + 0.5 tanh(t/2))
template <bool SILU>
__device__ __forceinline__ float act_grad8(float t, int act) {
  if (SILU) {
    const float sg = fmaf(0.5f, tanh_approx(0.5f * t), 0.5f);
    return sg * fmaf(t, 1.f - sg, 1.f);
  }
  return act_grad(t, act);
}

struct GnBwdCoef {
  float r[8], m0[8];   // xhat = x * r + m0        (r = rstd, m0 = (pre_add - mean) * rstd)
  float gf[8], bf[8];  // t = xhat * gf + bf        (gf = gamma * (1 + sc), bf = beta * (1 + sc) + sh)
};
__device__ __forceinline__ void gn_bwd_coef8(GnBwdCoef& k, int n, int c, int C, int G, const float* __restrict__ stats,
                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                             const float* __restrict__ pre_add, int pre_stride, const float* __restrict__ scale,
                                             const float* __restrict__ shift, int film_stride) {
  const int cpg = C / G;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    const float mean = stats[((int64_t)n * G + g) * 2], rstd = stats[((int64_t)n * G + g) * 2 + 1];
    const float pa = pre_add ? pre_add[(int64_t)n * pre_stride + c + j] : 0.f;
    const float f = scale ? 1.f + scale[(int64_t)n * film_stride + c + j] : 1.f;
    const float sh = shift ? shift[(int64_t)n * film_stride + c + j] : 0.f;
    k.r[j] = rstd;
    k.m0[j] = (pa - mean) * rstd;
    k.gf[j] = gamma[c + j] * f;
    k.bf[j] = beta[c + j] * f + sh;
  }
}
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(u[j] << 16); v[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u); }
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS, 2) gn_bwd_pass1_8_kernel(const bf16* __restrict__ x, const bf16* __restrict__ x2,
                                                                       const bf16* __restrict__ dy, const float* __restrict__ pre_add,
                                                                       int pre_stride, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, const float* __restrict__ scale,
                                                                       const float* __restrict__ shift, int film_stride,
                                                                       const float* __restrict__ stats, int C1, int C2, int64_t S, int G,
                                                                       int chunks, int act, float* __restrict__ partial) {
  const int C = C1 + C2;
  const int CV = C >> 3;
  const int rows = GN_THREADS / CV;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per, p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const int row = tid / CV, cv = tid - row * CV;
  const int c = cv * 8;
  __shared__ float sc1[2048], sc2[2048], sc3[2048];
  float a1[8], a2[8], a3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a1[j] = 0.f; a2[j] = 0.f; a3[j] = 0.f; }
  if (row < rows) {
    GnBwdCoef k;
    gn_bwd_coef8(k, n, c, C, G, stats, gamma, beta, pre_add, pre_stride, scale, shift, film_stride);
    const bf16* src = c < C1 ? x + (int64_t)n * S * C1 + c : x2 + (int64_t)n * S * C2 + (c - C1);
    const int64_t stride = c < C1 ? C1 : C2;
    const bf16* dsrc = dy + (int64_t)n * S * C + c;
    auto accum = [&](const uint4& qx, const uint4& qd) {
      float v[8], d[8];
      unpack8(qx, v);
      unpack8(qd, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = fmaf(v[j], k.r[j], k.m0[j]);
        const float dv = d[j] * act_grad8<SILU>(fmaf(xh, k.gf[j], k.bf[j]), act);
        a1[j] += dv;
        a2[j] = fmaf(dv, xh, a2[j]);
        a3[j] += xh;
      }
    };
    int64_t p = p0 + row;
    for (; p + rows < p1; p += 2 * (int64_t)rows) {
      const uint4 qx0 = __ldg(reinterpret_cast<const uint4*>(src + p * stride));
      const uint4 qx1 = __ldg(reinterpret_cast<const uint4*>(src + (p + rows) * stride));
      const uint4 qd0 = __ldg(reinterpret_cast<const uint4*>(dsrc + p * C));
      const uint4 qd1 = __ldg(reinterpret_cast<const uint4*>(dsrc + (p + rows) * C));
      accum(qx0, qd0);
      accum(qx1, qd1);
    }
    for (; p < p1; p += rows) accum(__ldg(reinterpret_cast<const uint4*>(src + p * stride)), __ldg(reinterpret_cast<const uint4*>(dsrc + p * C)));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = row * C + c + j;
      sc1[idx] = a1[j]; sc2[idx] = a2[j]; sc3[idx] = a3[j];
    }
  }
  __syncthreads();
  float* out = partial + ((int64_t)n * chunks + chunk) * C * 3;
  for (int cc = tid; cc < C; cc += GN_THREADS) {
    float t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (int r = 0; r < rows; ++r) { t1 += sc1[r * C + cc]; t2 += sc2[r * C + cc]; t3 += sc3[r * C + cc]; }
    out[cc * 3] = t1; out[cc * 3 + 1] = t2; out[cc * 3 + 2] = t3;
  }
}

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS, 2) gn_bwd_pass2_8_kernel(const bf16* __restrict__ x, const bf16* __restrict__ x2,
                                                                       const bf16* __restrict__ dy, const float* __restrict__ pre_add,
                                                                       int pre_stride, const float* __restrict__ gamma,
                                                                       const float* __restrict__ beta, const float* __restrict__ scale,
                                                                       const float* __restrict__ shift, int film_stride,
                                                                       const float* __restrict__ stats, const float* __restrict__ gmean,
                                                                       int C1, int C2, int64_t S, int G, int chunks, int act,
                                                                       bf16* __restrict__ dx, bf16* __restrict__ dx2) {
  const int C = C1 + C2;
  const int CV = C >> 3;
  const int rows = GN_THREADS / CV;
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int64_t per = (S + chunks - 1) / chunks;
  const int64_t p0 = (int64_t)chunk * per, p1 = min(S, p0 + per);
  const int tid = threadIdx.x;
  const int row = tid / CV, cv = tid - row * CV;
  if (row >= rows) return;
  const int c = cv * 8;
  const int cpg = C / G;
  GnBwdCoef k;
  gn_bwd_coef8(k, n, c, C, G, stats, gamma, beta, pre_add, pre_stride, scale, shift, film_stride);
  // dx = rstd * (gf * dv - m1 - xhat * m2) = dv * A - B1 - xhat * B2
  float A[8], B1[8], B2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    A[j] = k.r[j] * k.gf[j];
    B1[j] = k.r[j] * gmean[((int64_t)n * G + g) * 2];
    B2[j] = k.r[j] * gmean[((int64_t)n * G + g) * 2 + 1];
  }
  const bf16* src = c < C1 ? x + (int64_t)n * S * C1 + c : x2 + (int64_t)n * S * C2 + (c - C1);
  bf16* dst = c < C1 ? dx + (int64_t)n * S * C1 + c : dx2 + (int64_t)n * S * C2 + (c - C1);
  const int64_t stride = c < C1 ? C1 : C2;
  const bf16* dsrc = dy + (int64_t)n * S * C + c;
  auto transform = [&](const uint4& qx, const uint4& qd) -> uint4 {
    float v[8], d[8];
    unpack8(qx, v);
    unpack8(qd, d);
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float r2[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float xh = fmaf(v[j + e], k.r[j + e], k.m0[j + e]);
        const float dv = d[j + e] * act_grad8<SILU>(fmaf(xh, k.gf[j + e], k.bf[j + e]), act);
        r2[e] = fmaf(dv, A[j + e], -fmaf(xh, B2[j + e], B1[j + e]));
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(r2[0], r2[1]);
      ow[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return o;
  };
  int64_t p = p0 + row;
  for (; p + rows < p1; p += 2 * (int64_t)rows) {
    const uint4 qx0 = __ldg(reinterpret_cast<const uint4*>(src + p * stride));
    const uint4 qx1 = __ldg(reinterpret_cast<const uint4*>(src + (p + rows) * stride));
    const uint4 qd0 = __ldg(reinterpret_cast<const uint4*>(dsrc + p * C));
    const uint4 qd1 = __ldg(reinterpret_cast<const uint4*>(dsrc + (p + rows) * C));
    *reinterpret_cast<uint4*>(dst + p * stride) = transform(qx0, qd0);
    *reinterpret_cast<uint4*>(dst + (p + rows) * stride) = transform(qx1, qd1);
  }
  for (; p < p1; p += rows)
    *reinterpret_cast<uint4*>(dst + p * stride) = transform(__ldg(reinterpret_cast<const uint4*>(src + p * stride)), __ldg(reinterpret_cast<const uint4*>(dsrc + p * C)));
}

static int gn_chunks_bwd8(int N, int64_t S, int C) {
  const int rows = GN_THREADS / (C >> 3);
  int64_t chunks = ceil_div((int64_t)sm_count() * 2, N);     // one resident wave: 2 CTAs per SM
  int64_t max_chunks = ceil_div(S, (int64_t)rows * 8);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  return (int)chunks;
}
static bool gn_fast8(const rho_gn_params* p) {
  const int C = p->C1 + p->C2;
  return p->in_dtype == RHO_BF16 && (p->C1 % 8 == 0) && (p->C2 % 8 == 0) && C <= 2048 && (C >> 3) <= GN_THREADS;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int gn_chunks(int N, int64_t S, int C) {
  GnMap mp = gn_map(C);
  int64_t target = (int64_t)sm_count() * 4;
  int64_t chunks = ceil_div(target, N);
  int64_t max_chunks = ceil_div(S, (int64_t)mp.rows_per_pass * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  if (chunks > 4096) chunks = 4096;
  return (int)chunks;
}

static int gn_validate(const rho_gn_params* p, const char* who) {
  RHO_REQUIRE(p, "%s: null params", who);
  RHO_REQUIRE(p->N > 0 && p->S > 0 && p->C1 > 0 && p->C2 >= 0, "%s: bad dims", who);
  const int C = p->C1 + p->C2;
  RHO_REQUIRE(p->groups > 0 && p->groups <= 64 && C % p->groups == 0, "%s: C=%d not divisible by groups=%d", who, C, p->groups);
  RHO_REQUIRE(C % 4 == 0 && p->C1 % 4 == 0, "%s: channel counts must be multiples of 4 (C1=%d C2=%d)", who, p->C1, p->C2);
  RHO_REQUIRE(C <= GN_MAX_C, "%s: C=%d exceeds %d", who, C, GN_MAX_C);
  RHO_REQUIRE(p->x && p->stats, "%s: null tensor", who);
  RHO_REQUIRE((p->C2 == 0) == (p->x2 == nullptr), "%s: x2/C2 mismatch", who);
  RHO_REQUIRE(p->N <= 65535, "%s: N too large", who);
  return RHO_OK;
}

}  // namespace rho

using namespace rho;

extern "C" {

size_t rho_gn_workspace_bytes(const rho_gn_params* p) {
  if (!p) return 0;
  int chunks = gn_fast8(p) ? gn_chunks8(p->N, p->S, p->C1 + p->C2) : gn_chunks(p->N, p->S, p->C1 + p->C2);
  return (size_t)p->N * chunks * p->groups * 3 * sizeof(float);
}

int rho_gn_stats(const rho_gn_params* p, rho_stream_t stream) {
  int rc = gn_validate(p, "gn_stats");
  if (rc) return rc;
  const int C = p->C1 + p->C2;
  const bool fast = gn_fast8(p);
  int chunks = fast ? gn_chunks8(p->N, p->S, C) : gn_chunks(p->N, p->S, C);
  size_t need = (size_t)p->N * chunks * p->groups * 3 * sizeof(float);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "gn_stats: workspace %zu < %zu", p->workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  dim3 grid(chunks, p->N);
  float* partial = (float*)p->workspace;
  if (fast)
    gn_stats8_kernel<<<grid, GN_THREADS, 0, st>>>((const bf16*)p->x, (const bf16*)p->x2, p->pre_add, p->pre_add_stride, p->C1, p->C2, p->S,
                                                  p->groups, chunks, partial);
  else if (p->in_dtype == RHO_F32)
    gn_stats_kernel<float><<<grid, GN_THREADS, 0, st>>>((const float*)p->x, (const float*)p->x2, p->pre_add, p->pre_add_stride, p->C1,
                                                         p->C2, p->S, p->groups, chunks, partial);
  else if (p->in_dtype == RHO_BF16)
    gn_stats_kernel<bf16><<<grid, GN_THREADS, 0, st>>>((const bf16*)p->x, (const bf16*)p->x2, p->pre_add, p->pre_add_stride, p->C1,
                                                        p->C2, p->S, p->groups, chunks, partial);
  else
    return fail(RHO_ERR_INVALID_ARG, "gn_stats: bad dtype");
  RHO_CHECK_LAUNCH("gn_stats");
  int NG = p->N * p->groups;
  gn_stats_finalize_kernel<<<(NG * 32 + 127) / 128, 128, 0, st>>>(partial, NG, p->groups, chunks, p->eps, p->stats);
  RHO_CHECK_LAUNCH("gn_stats_finalize");
  return RHO_OK;
}

int rho_gn_stats_from_partials(const rho_gn_params* p, const float* part1, int32_t rows1, const float* part2, int32_t rows2,
                               rho_stream_t stream) {
  int rc = gn_validate(p, "gn_stats_from_partials");
  if (rc) return rc;
  RHO_REQUIRE(part1 && rows1 > 0, "gn_stats_from_partials: no statistics for x");
  RHO_REQUIRE((p->C2 == 0) || (part2 && rows2 > 0), "gn_stats_from_partials: no statistics for x2");
  RHO_REQUIRE(p->pre_add == nullptr, "gn_stats_from_partials: pre_add changes the per-group sums; use rho_gn_stats");
  RHO_REQUIRE(p->N <= 65535, "gn_stats_from_partials: N too large");
  const int cpg = (p->C1 + p->C2) / p->groups;
  const long long work = (long long)std::max(rows1, rows2) * cpg;        // (row, channel) pairs per (sample, group)
  // one warp per (sample, group) walks its pairs serially at one L2 round trip (~0.4 us) each -- ptxas pairs every load with its use,
  // also in an unrolled loop -- so it only takes the tiny cases (<= 4 trips per lane); the CTA kernel spreads the rest over 256 threads
  // (measured: 28 launches of the 1024-pair case cost 12.9 us each on one warp per group)
  const int NG = p->N * p->groups;
  if (work <= 128 || (work <= 1024 && NG > 16 * sm_count())) {      // (many-sample batches: enough warps to hide the trips)
    gn_stats_from_partials_warp_kernel<<<(NG * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(part1, rows1, p->C1, part2, rows2, p->C2, p->S,
                                                                                           p->groups, NG, p->eps, p->stats);
  } else {
    gn_stats_from_partials_kernel<<<dim3(p->groups, p->N), 256, 0, as_stream(stream)>>>(part1, rows1, p->C1, part2, rows2, p->C2, p->S,
                                                                                        p->groups, p->eps, p->stats);
  }
  RHO_CHECK_LAUNCH("gn_stats_from_partials");
  return RHO_OK;
}

int rho_gn_apply(const rho_gn_params* p, rho_stream_t stream) {
  int rc = gn_validate(p, "gn_apply");
  if (rc) return rc;
  RHO_REQUIRE(p->y && p->gamma && p->beta, "gn_apply: null tensor");
  RHO_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), "gn_apply: scale/shift must both be set");
  const int C = p->C1 + p->C2;
  cudaStream_t st = as_stream(stream);
  if (gn_fast8(p) && p->out_dtype == RHO_BF16) {
    int chunks8 = gn_chunks8(p->N, p->S, C);
    dim3 grid8(chunks8, p->N);
    if (p->act == RHO_ACT_SILU)
      gn_apply8_kernel<true><<<grid8, GN_THREADS, 0, st>>>((const bf16*)p->x, (const bf16*)p->x2, p->pre_add, p->pre_add_stride, p->gamma,
                                                           p->beta, p->scale, p->shift, p->film_stride, p->stats, p->C1, p->C2, p->S,
                                                           p->groups, chunks8, p->act, (bf16*)p->y);
    else
      gn_apply8_kernel<false><<<grid8, GN_THREADS, 0, st>>>((const bf16*)p->x, (const bf16*)p->x2, p->pre_add, p->pre_add_stride, p->gamma,
                                                            p->beta, p->scale, p->shift, p->film_stride, p->stats, p->C1, p->C2, p->S,
                                                            p->groups, chunks8, p->act, (bf16*)p->y);
    RHO_CHECK_LAUNCH("gn_apply8");
    return RHO_OK;
  }
  int chunks = gn_chunks(p->N, p->S, C);
  dim3 grid(chunks, p->N);
#define RHO_GNA(TI, TO)                                                                                                        \
  gn_apply_kernel<TI, TO><<<grid, GN_THREADS, 0, st>>>((const TI*)p->x, (const TI*)p->x2, p->pre_add, p->pre_add_stride, p->gamma, \
                                                       p->beta, p->scale, p->shift, p->film_stride, p->stats, p->C1, p->C2, p->S,  \
                                                       p->groups, chunks, p->act, (TO*)p->y)
  if (p->in_dtype == RHO_F32 && p->out_dtype == RHO_F32) RHO_GNA(float, float);
  else if (p->in_dtype == RHO_F32 && p->out_dtype == RHO_BF16) RHO_GNA(float, bf16);
  else if (p->in_dtype == RHO_BF16 && p->out_dtype == RHO_BF16) RHO_GNA(bf16, bf16);
  else if (p->in_dtype == RHO_BF16 && p->out_dtype == RHO_F32) RHO_GNA(bf16, float);
  else return fail(RHO_ERR_INVALID_ARG, "gn_apply: bad dtype");
#undef RHO_GNA
  RHO_CHECK_LAUNCH("gn_apply");
  return RHO_OK;
}

size_t rho_gn_bwd_workspace_bytes(const rho_gn_bwd_params* p) {
  if (!p) return 0;
  const rho_gn_params& f = p->fwd;
  const int C = f.C1 + f.C2;
  int chunks = std::max(gn_chunks(f.N, f.S, C), gn_fast8(&f) ? gn_chunks_bwd8(f.N, f.S, C) : 0);
  size_t partial = (size_t)f.N * chunks * C * 3;
  size_t sums = (size_t)f.N * C * 3;
  size_t gmean = (size_t)f.N * f.groups * 2;
  return (partial + 2 * sums + gmean) * sizeof(float);
}

int rho_gn_bwd(const rho_gn_bwd_params* p, rho_stream_t stream) {
  RHO_REQUIRE(p, "gn_bwd: null params");
  const rho_gn_params& f = p->fwd;
  int rc = gn_validate(&f, "gn_bwd");
  if (rc) return rc;
  RHO_REQUIRE(p->dy && p->dx && f.gamma && f.beta, "gn_bwd: null tensor");
  RHO_REQUIRE((f.C2 == 0) || p->dx2, "gn_bwd: dx2 required for concat input");
  const int C = f.C1 + f.C2;
  const bool fast = gn_fast8(&f) && f.out_dtype == RHO_BF16;
  int chunks = fast ? gn_chunks_bwd8(f.N, f.S, C) : gn_chunks(f.N, f.S, C);
  size_t need = rho_gn_bwd_workspace_bytes(p);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "gn_bwd: workspace %zu < %zu", p->workspace_bytes, need);
  float* partial = (float*)p->workspace;
  float* sums = partial + (size_t)f.N * chunks * C * 3;
  float* colsums = sums + (size_t)f.N * C * 3;
  float* gmean = colsums + (size_t)f.N * C * 3;
  cudaStream_t st = as_stream(stream);
  dim3 grid(chunks, f.N);
#define RHO_GNB1(TI, TO)                                                                                                          \
  gn_bwd_pass1_kernel<TI, TO><<<grid, GN_THREADS, 0, st>>>((const TI*)f.x, (const TI*)f.x2, (const TO*)p->dy, f.pre_add,           \
                                                           f.pre_add_stride, f.gamma, f.beta, f.scale, f.shift, f.film_stride,    \
                                                           f.stats, f.C1, f.C2, f.S, f.groups, chunks, f.act, partial)
#define RHO_GNB2(TI, TO)                                                                                                          \
  gn_bwd_pass2_kernel<TI, TO><<<grid, GN_THREADS, 0, st>>>((const TI*)f.x, (const TI*)f.x2, (const TO*)p->dy, f.pre_add,           \
                                                           f.pre_add_stride, f.gamma, f.beta, f.scale, f.shift, f.film_stride,    \
                                                           f.stats, gmean, f.C1, f.C2, f.S, f.groups, chunks, f.act, (TI*)p->dx,   \
                                                           (TI*)p->dx2)
  const int key = f.in_dtype * 2 + f.out_dtype;
#define RHO_GNB8(KERN, SILU, ...)                                                                                              \
  KERN<SILU><<<grid, GN_THREADS, 0, st>>>((const bf16*)f.x, (const bf16*)f.x2, (const bf16*)p->dy, f.pre_add, f.pre_add_stride, \
                                          f.gamma, f.beta, f.scale, f.shift, f.film_stride, f.stats, __VA_ARGS__)
  if (fast) {
    if (f.act == RHO_ACT_SILU) RHO_GNB8(gn_bwd_pass1_8_kernel, true, f.C1, f.C2, f.S, f.groups, chunks, f.act, partial);
    else RHO_GNB8(gn_bwd_pass1_8_kernel, false, f.C1, f.C2, f.S, f.groups, chunks, f.act, partial);
  } else
  switch (key) {
    case 0: RHO_GNB1(float, float); break;
    case 1: RHO_GNB1(float, bf16); break;
    case 2: RHO_GNB1(bf16, float); break;
    case 3: RHO_GNB1(bf16, bf16); break;
    default: return fail(RHO_ERR_INVALID_ARG, "gn_bwd: bad dtype");
  }
  RHO_CHECK_LAUNCH("gn_bwd_pass1");
  gn_bwd_colsum_kernel<<<dim3((C + 7) / 8, f.N), 256, 0, st>>>(partial, chunks, C, colsums);
  RHO_CHECK_LAUNCH("gn_bwd_colsum");
  gn_bwd_reduce_kernel<<<f.N, GN_THREADS, 0, st>>>(colsums, 1, C, f.groups, f.S, f.gamma, f.beta, f.scale, f.film_stride, f.stats,
                                                   sums, gmean, p->dscale, p->dshift, p->dfilm_stride, p->dpre_add);
  RHO_CHECK_LAUNCH("gn_bwd_reduce");
  if (p->dgamma || p->dbeta) {
    gn_bwd_param_kernel<<<(C + 31) / 32, 256, 0, st>>>(sums, f.N, C, f.scale, f.film_stride, p->dgamma, p->dbeta);
    RHO_CHECK_LAUNCH("gn_bwd_param");
  }
  if (fast) {
    if (f.act == RHO_ACT_SILU) RHO_GNB8(gn_bwd_pass2_8_kernel, true, gmean, f.C1, f.C2, f.S, f.groups, chunks, f.act, (bf16*)p->dx, (bf16*)p->dx2);
    else RHO_GNB8(gn_bwd_pass2_8_kernel, false, gmean, f.C1, f.C2, f.S, f.groups, chunks, f.act, (bf16*)p->dx, (bf16*)p->dx2);
  } else
  switch (key) {
    case 0: RHO_GNB2(float, float); break;
    case 1: RHO_GNB2(float, bf16); break;
    case 2: RHO_GNB2(bf16, float); break;
    case 3: RHO_GNB2(bf16, bf16); break;
  }
  RHO_CHECK_LAUNCH("gn_bwd_pass2");
#undef RHO_GNB1
#undef RHO_GNB2
#undef RHO_GNB8
  return RHO_OK;
}

}  // extern "C"
