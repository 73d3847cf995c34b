// Synthetic density-field generators ON THE DEVICE: the training input pipeline of the synthetic
// configurations never touches the host, so it stays off the critical path at 8 GPUs.
//
//   rho_sph_harm_density  the spherical-harmonic volumes of rho_diffusion/data/synthetic.py:45-124 (the data of
//                         examples/config_spherical_harmonics.json), restated with the reference's quirks:
//                         grid = linspace(lo, hi, G) per axis, meshgrid(indexing="xy") (volume[i, j, k] <-> x[j], y[i], z[k]),
//                         theta = atan(sqrt(x^2 + y^2) / z) is passed to scipy's sph_harm as the AZIMUTH and
//                         phi = atan(y / x) as the POLAR angle; solution = Y_l^|m|(polar, azimuth) * r is min-max normalised
//                         as a COMPLEX array (numpy's min / max order complex numbers lexicographically) and the density is
//                         |solution|.  Double precision throughout, like numpy; output fp32.  The lexicographic extremes are
//                         ILL-CONDITIONED: symmetric grid points have real parts that agree to the last bits, so which of them
//                         (with opposite imaginary parts) wins depends on rounding noise of the special-function library.  The
//                         caller therefore normally passes the extremes of the host restatement (`minmax`, cached per (l, |m|)),
//                         which makes the device field equal the reference's to ~1e-15; with `minmax = NULL` the kernel uses its
//                         own lexicographic extremes (equal to the reference's only where no such near-tie exists).
//   rho_gaussian_blobs    n-D (n = 1, 2, 3) Gaussian-blob density fields: rho(x) = sum_k exp(-|x - c_k|^2 / (2 w^2)) with K
//                         centres c_k uniform in [0.2, 0.8]^n drawn from Philox4x32-10 (key = seed, counter = (sample index, k)),
//                         x in [0, 1]^n at cell centres; output = 2 min(rho, 1) - 1 in [-1, 1].  Labels are (K, w).
#include "common.cuh"

namespace rho {

// ---- Philox4x32-10 (same generator as diffusion.cu) --------------------------------------------------------
__device__ __forceinline__ void dg_philox(uint64_t ctr_lo, uint64_t ctr_hi, uint64_t key, uint32_t (&out)[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr int DG_MAX_BLOBS = 8;

__global__ void __launch_bounds__(256) gaussian_blobs_kernel(float* __restrict__ out, int N, int D, int H, int W, const int32_t* __restrict__ n_blobs,
                                                             const float* __restrict__ width, uint64_t seed, uint64_t first_index) {
  const int n = blockIdx.y;
  __shared__ float sc[DG_MAX_BLOBS][3];
  const int K = min(max(n_blobs[n], 0), DG_MAX_BLOBS);
  if (threadIdx.x < K) {
    uint32_t r[4];
    dg_philox(first_index + (uint64_t)n, (uint64_t)threadIdx.x, seed, r);
#pragma unroll
    for (int a = 0; a < 3; ++a) sc[threadIdx.x][a] = 0.2f + 0.6f * ((float)(r[a] >> 8) + 0.5f) * 5.9604644775390625e-08f;
  }
  __syncthreads();
  const float w = width[n];
  const float inv2w2 = 0.5f / (w * w);
  const long long S = (long long)D * H * W;
  const float fd = 1.0f / (float)D, fh = 1.0f / (float)H, fw = 1.0f / (float)W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < S; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), z = (int)(i / ((long long)W * H));
    // coordinate order: the LAST spatial axis is axis 0 of the centre, so 1-D / 2-D fields use (c0) / (c0, c1)
    const float px = ((float)x + 0.5f) * fw, py = ((float)y + 0.5f) * fh, pz = ((float)z + 0.5f) * fd;
    float rho = 0.f;
    for (int k = 0; k < K; ++k) {
      float d2 = (px - sc[k][0]) * (px - sc[k][0]);
      if (H > 1) d2 += (py - sc[k][1]) * (py - sc[k][1]);
      if (D > 1) d2 += (pz - sc[k][2]) * (pz - sc[k][2]);
      rho += expf(-d2 * inv2w2);
    }
    out[(long long)n * S + i] = 2.0f * fminf(rho, 1.0f) - 1.0f;
  }
}

// ---- rotational-spectrum line profiles (rho_diffusion/data/spectroscopy.py:142-189) --------------------------------------
// profile[f] = sum_k I_k exp(-(nu_f - c_k)^2 / (2 w^2)) over the transitions whose centre lies inside the grid, with numpy's
// promotion rules (fp32 difference and square, fp64 division / exp / sum), then divided by its maximum (spectroscopy.py:120-121).
// Ragged batch: sample b owns transitions [offsets[b], offsets[b + 1]).
__global__ void __launch_bounds__(256) lineprofile_kernel(double* __restrict__ prof, double* __restrict__ partial_max, int G,
                                                          const float* __restrict__ grid, const float* __restrict__ centers,
                                                          const float* __restrict__ intens, const int32_t* __restrict__ offsets,
                                                          const double* __restrict__ widths) {
  const int b = blockIdx.y;
  const int k0 = offsets[b], k1 = offsets[b + 1];
  const float lo = grid[0], hi = grid[G - 1];
  const double w = widths[b];
  const double denom = 2.0 * (w * w);
  double mx = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < G; f += gridDim.x * blockDim.x) {
    const float nu = grid[f];
    double acc = 0.0;
    for (int k = k0; k < k1; ++k) {
      const float c = centers[k];
      if (c > hi || c < lo) continue;
      const float d = nu - c;
      acc += (double)intens[k] * exp(-((double)(d * d)) / denom);
    }
    prof[(size_t)b * G + f] = acc;
    mx = fmax(mx, acc);
  }
  __shared__ double red[256];
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial_max[(size_t)b * gridDim.x + blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) lineprofile_normalize_kernel(const double* __restrict__ prof, const double* __restrict__ partial_max, int G,
                                                                    float* __restrict__ out) {
  const int b = blockIdx.y;
  double mx = 0.0;
  for (unsigned i = 0; i < gridDim.x; ++i) mx = fmax(mx, partial_max[(size_t)b * gridDim.x + i]);
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < G; f += gridDim.x * blockDim.x)
    out[(size_t)b * G + f] = (float)(prof[(size_t)b * G + f] / mx);
}

// ---- spherical harmonics -------------------------------------------------------------------------------------
struct Cplx {
  double re, im;
};
__device__ __forceinline__ bool lex_less(const Cplx& a, const Cplx& b) { return a.re < b.re || (a.re == b.re && a.im < b.im); }

// Y_l^m(polar, azimuth) * r at one grid point (m >= 0), scipy.special.sph_harm_y convention (Condon-Shortley phase,
// (1 - x^2)^(m/2) = |sin polar|^m)
__device__ Cplx sph_point(int l, int m, double xv, double yv, double zv, double norm) {
  const double azimuth = atan(sqrt(xv * xv + yv * yv) / zv);     // the reference's "theta"
  const double polar = atan(yv / xv);                            // the reference's "phi"
  const double x = cos(polar), s = fabs(sin(polar));
  double pmm = 1.0;
  for (int i = 1; i <= m; ++i) pmm *= -(double)(2 * i - 1) * s;
  double plm = pmm;
  if (l > m) {
    double a = pmm, b = x * (double)(2 * m + 1) * pmm;
    for (int ll = m + 2; ll <= l; ++ll) {
      const double c = ((double)(2 * ll - 1) * x * b - (double)(ll + m - 1) * a) / (double)(ll - m);
      a = b; b = c;
    }
    plm = b;
  }
  const double r = sqrt(xv * xv + yv * yv + zv * zv);
  double sn, cs;
  sincos((double)m * azimuth, &sn, &cs);
  Cplx o;
  o.re = norm * plm * cs * r;
  o.im = norm * plm * sn * r;
  return o;
}

__global__ void __launch_bounds__(256) sph_harm_solution_kernel(Cplx* __restrict__ sol, Cplx* __restrict__ partial, int G, const int32_t* __restrict__ lv,
                                                                const int32_t* __restrict__ mv, double lo, double hi) {
  const int n = blockIdx.y;
  const int l = lv[n], m = abs(mv[n]);
  double norm = (double)(2 * l + 1) / (4.0 * 3.14159265358979323846);
  for (int i = l - m + 1; i <= l + m; ++i) norm /= (double)i;    // (l - m)! / (l + m)!
  norm = sqrt(norm);
  const double step = (hi - lo) / (double)(G - 1);
  const long long S = (long long)G * G * G;
  Cplx mn = {INFINITY, INFINITY}, mx = {-INFINITY, -INFINITY};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < S; idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % G), j = (int)((idx / G) % G), i = (int)(idx / ((long long)G * G));
    const double xv = j == G - 1 ? hi : lo + j * step, yv = i == G - 1 ? hi : lo + i * step, zv = k == G - 1 ? hi : lo + k * step;
    const Cplx v = sph_point(l, m, xv, yv, zv, norm);
    sol[(long long)n * S + idx] = v;
    if (lex_less(v, mn)) mn = v;
    if (lex_less(mx, v)) mx = v;
  }
  __shared__ Cplx smn[256], smx[256];
  smn[threadIdx.x] = mn; smx[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      if (lex_less(smn[threadIdx.x + o], smn[threadIdx.x])) smn[threadIdx.x] = smn[threadIdx.x + o];
      if (lex_less(smx[threadIdx.x], smx[threadIdx.x + o])) smx[threadIdx.x] = smx[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[((long long)n * gridDim.x + blockIdx.x) * 2] = smn[0];
    partial[((long long)n * gridDim.x + blockIdx.x) * 2 + 1] = smx[0];
  }
}

__global__ void __launch_bounds__(256) sph_harm_normalize_kernel(const Cplx* __restrict__ sol, const Cplx* __restrict__ partial, int G, float* __restrict__ out,
                                                                 const double* __restrict__ minmax) {
  const int n = blockIdx.y;
  __shared__ Cplx s_mn, s_mx;
  if (threadIdx.x == 0 && minmax) {
    s_mn.re = minmax[4 * n]; s_mn.im = minmax[4 * n + 1]; s_mx.re = minmax[4 * n + 2]; s_mx.im = minmax[4 * n + 3];
  } else if (threadIdx.x == 0) {
    Cplx mn = partial[(long long)n * gridDim.x * 2], mx = partial[(long long)n * gridDim.x * 2 + 1];
    for (int b = 1; b < (int)gridDim.x; ++b) {
      const Cplx a = partial[((long long)n * gridDim.x + b) * 2], c = partial[((long long)n * gridDim.x + b) * 2 + 1];
      if (lex_less(a, mn)) mn = a;
      if (lex_less(mx, c)) mx = c;
    }
    s_mn = mn; s_mx = mx;
  }
  __syncthreads();
  const Cplx mn = s_mn;
  const double dr = s_mx.re - mn.re, di = s_mx.im - mn.im;
  const double den = dr * dr + di * di;
  const long long S = (long long)G * G * G;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < S; idx += (long long)gridDim.x * blockDim.x) {
    const Cplx v = sol[(long long)n * S + idx];
    const double ar = v.re - mn.re, ai = v.im - mn.im;
    // (ar + i ai) / (dr + i di)
    const double qr = (ar * dr + ai * di) / den, qi = (ai * dr - ar * di) / den;
    out[(long long)n * S + idx] = (float)sqrt(qr * qr + qi * qi);
  }
}

}  // namespace rho

using namespace rho;

extern "C" {

int rho_gaussian_blobs(float* out, int32_t N, int32_t D, int32_t H, int32_t W, const int32_t* n_blobs, const float* width, uint64_t seed,
                       uint64_t first_index, rho_stream_t stream) {
  RHO_REQUIRE(out && n_blobs && width && N > 0 && N <= 65535 && D > 0 && H > 0 && W > 0, "gaussian_blobs: bad args");
  const long long S = (long long)D * H * W;
  const unsigned bx = (unsigned)std::min<long long>(ceil_div(S, 256), 1024);
  gaussian_blobs_kernel<<<dim3(bx, (unsigned)N), 256, 0, as_stream(stream)>>>(out, N, D, H, W, n_blobs, width, seed, first_index);
  RHO_CHECK_LAUNCH("gaussian_blobs");
  return RHO_OK;
}

static unsigned sph_blocks(int G) { return (unsigned)std::min<long long>(ceil_div((long long)G * G * G, 256), 128); }

size_t rho_sph_harm_workspace_bytes(int32_t N, int32_t G) {
  return (size_t)N * G * G * G * sizeof(Cplx) + (size_t)N * sph_blocks(G) * 2 * sizeof(Cplx);
}

int rho_sph_harm_density(float* out, int32_t N, int32_t G, const int32_t* l, const int32_t* m, double lo, double hi, const double* minmax,
                         void* workspace, size_t workspace_bytes, rho_stream_t stream) {
  RHO_REQUIRE(out && l && m && N > 0 && N <= 65535 && G > 1, "sph_harm_density: bad args");
  if (!workspace || workspace_bytes < rho_sph_harm_workspace_bytes(N, G)) return fail(RHO_ERR_WORKSPACE, "sph_harm_density: workspace too small");
  Cplx* sol = (Cplx*)workspace;
  Cplx* partial = sol + (size_t)N * G * G * G;
  const dim3 grid(sph_blocks(G), (unsigned)N);
  sph_harm_solution_kernel<<<grid, 256, 0, as_stream(stream)>>>(sol, partial, G, l, m, lo, hi);
  RHO_CHECK_LAUNCH("sph_harm_solution");
  sph_harm_normalize_kernel<<<grid, 256, 0, as_stream(stream)>>>(sol, partial, G, out, minmax);
  RHO_CHECK_LAUNCH("sph_harm_normalize");
  return RHO_OK;
}

static unsigned lp_blocks(int G) { return (unsigned)std::min<long long>(ceil_div((long long)G, 256), 256); }

size_t rho_lineprofile_workspace_bytes(int32_t B, int32_t G) { return ((size_t)B * G + (size_t)B * lp_blocks(G)) * sizeof(double); }

int rho_lineprofile(float* out, int32_t B, int32_t G, const float* grid, const float* centers, const float* intensities, const int32_t* offsets,
                    const double* widths, void* workspace, size_t workspace_bytes, rho_stream_t stream) {
  RHO_REQUIRE(out && grid && centers && intensities && offsets && widths && B > 0 && B <= 65535 && G > 1, "lineprofile: bad args");
  if (!workspace || workspace_bytes < rho_lineprofile_workspace_bytes(B, G)) return fail(RHO_ERR_WORKSPACE, "lineprofile: workspace too small");
  double* prof = (double*)workspace;
  double* partial = prof + (size_t)B * G;
  const dim3 gridDim3(lp_blocks(G), (unsigned)B);
  lineprofile_kernel<<<gridDim3, 256, 0, as_stream(stream)>>>(prof, partial, G, grid, centers, intensities, offsets, widths);
  RHO_CHECK_LAUNCH("lineprofile");
  lineprofile_normalize_kernel<<<gridDim3, 256, 0, as_stream(stream)>>>(prof, partial, G, out);
  RHO_CHECK_LAUNCH("lineprofile_normalize");
  return RHO_OK;
}

}  // extern "C"
