// conv_nd weight gradient on the 5th-generation tensor cores (bf16 activations, fp32 accumulation, sm_100a).
//
//   dW[co][tap][ci] = sum over output positions  dY[pos][co] * X[pos + tap offset][ci]
//
// Both operands are channels-last, i.e. the reduction index (position) is the *row* and the channels are contiguous:
// exactly the MN-major canonical layout of tcgen05 (64 channels = one 128-byte swizzled row, 8 rows per atom).  The same
// 5-D TMA boxes as the forward pass are used -- a box of 128 output positions of dY, and the box of X shifted by the
// tap offset (out-of-bounds = zero = padding) -- and fed to the MMA without any transpose:
//   D[m = (tap-or-ci-block atom, ci)][n = co] += A^T B,  A = X box(es) [128 pos x 64 ci] x 2 atoms,  B = dY box(es).
// M = 128 is always two 64-channel atoms: two ci blocks of one tap when Cin >= 128, or the same ci block of TWO taps
// when Cin = 64 (so 64-channel layers still fill the 128-row MMA).  Split-K over position tiles; the fp32 partials
// [split][tap][ci][co] are reduced in a fixed order (deterministic) into the PyTorch layout [co][ci][tap].
// Strided (Downsample, unet_v2.py:153) and fused-upsample (Upsample, unet_v2.py:124-131) convs use the same kernel:
//   stride 2 : the X box of tap t comes from the parity view (coordinate j <-> input 2j + parity) that tap reads;
//   upsample : the reduction runs over the 4 output-parity classes; class (a, b) pairs the parity view of dY with
//              the X box shifted by floor((a + dh) / 2), floor((b + dw) / 2) on the un-upsampled grid.
#include "conv_geom.cuh"
#include "tc_common.cuh"

namespace rho {

struct WgParams {
  int bn, bd, bh, bw, lbw, lbh, lbd;
  int tiles_w, tiles_h, tiles_d, tiles_n, tiles_total;
  int Cout, Cin, taps, C1;
  int n_co_blocks;       // Cout / BN
  int m_units;           // number of (atom pair) units = CTAs along M
  int splits, tiles_per_split;
  int ncls;              // reduction classes per position tile (output parities of a fused upsample; else 1)
  int mode;              // 0 plain (x map = concat source), 1 strided (x map = parity view of the tap), 2 fused upsample
  int8_t off_d[28][4], off_h[28][4], off_w[28][4];   // X box shift per (tap, class)
  int8_t tap_xmap[28];
  // unit u -> two atoms: (tap, ci block) each; -1 tap = unused (duplicate)
  int16_t unit_tap[2][256];
  int16_t unit_cib[2][256];
};

struct WgMaps {
  CUtensorMap x[4];   // plain: x, x2 (concat partner);  strided: parity views (ph * 2 + pw)
  CUtensorMap dy[4];  // fused upsample: output-parity views (a * uw + b);  else dy[0]
};

__device__ __forceinline__ uint64_t wg_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) conv_wgrad_tc_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgParams p,
                                                            float* __restrict__ partial) {
  constexpr int BOX = 128 * 128;                 // bytes of one [128 pos x 64 ch] bf16 box
  constexpr int A_BYTES = 2 * BOX, B_BYTES = (BN / 64) * BOX, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int cob = bid % p.n_co_blocks;
  const int unit = bid / p.n_co_blocks;
  const int t0 = split * p.tiles_per_split;                      // steps = (position tile, class), class fastest
  const int t1 = min(p.tiles_total * p.ncls, t0 + p.tiles_per_split);
  const int nsteps = max(t1 - t0, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // whole warp walks the loop (uniform control flow keeps coordinates in uniform registers); one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t0; t < t1; ++t) {
      const int cls = t % p.ncls;
      int r = t / p.ncls;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; r /= p.tiles_h;
      const int td = r % p.tiles_d; r /= p.tiles_d;
      const int w0 = tw * p.bw, h0 = th * p.bh, d0 = td * p.bd, n0 = r * p.bn;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int tap = p.unit_tap[a][unit], c = p.unit_cib[a][unit] * 64;
          const int src = c >= p.C1 ? 1 : 0;
          const int xm = p.mode == 0 ? src : (p.mode == 1 ? p.tap_xmap[tap] : 0);
          tma_load_5d(sa + a * BOX, &maps.x[xm], &full_bar[stage], src ? c - p.C1 : c, w0 + p.off_w[tap][cls], h0 + p.off_h[tap][cls],
                      d0 + p.off_d[tap][cls], n0);
        }
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_load_5d(sa + A_BYTES + b * BOX, &maps.dy[cls], &full_bar[stage], cob * BN + b * 64, w0, h0, d0, n0);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // D fp32, A/B bf16, both MN-major, M = 128, N = BN
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    for (int s = 0; s < nsteps; ++s) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 16 positions per MMA: two 8-row atoms = 2048 bytes
          const uint64_t a = wg_desc_mn(sa + k * 2048, BOX);
          const uint64_t b = wg_desc_mn(sa + A_BYTES + k * 2048, BOX);
          umma_f16(tmem_base, a, b, idesc, (s | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&accum_bar);
    __syncwarp();
  } else {
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;            // accumulator row: atom = r / 64, ci within block = r % 64
    const int a = r >> 6;
    const int tap = p.unit_tap[a][unit];
    const int ci = p.unit_cib[a][unit] * 64 + (r & 63);
    const bool dup = a == 1 && p.unit_tap[0][unit] == p.unit_tap[1][unit] && p.unit_cib[0][unit] == p.unit_cib[1][unit];
    float* dst = partial + ((((long long)split * p.taps + tap) * p.Cin + ci) * p.Cout) + cob * BN;
    if (nsteps > 0) {
      mbar_wait(&accum_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      if (nsteps > 0) tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);
      else {
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = 0u;
      }
      if (!dup) {
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(dst + c0 + e) =
              make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN));
}

// dw[co][ci][tap] = sum_split partial[split][tap][ci][co]: a full index reversal.  One CTA owns 32 output channels x CIT
// input channels x all taps: the partials are read as 128-byte rows (co fastest), summed over the splits in a fixed order,
// staged in shared memory, and written as runs of CIT * taps contiguous floats per output channel.
__global__ void __launch_bounds__(256) wgrad_tc_reduce_kernel(const float* __restrict__ partial, int splits, int Cout, int Cin, int taps,
                                                              int cit, float* __restrict__ dw) {
  extern __shared__ float tile[];                      // [cit * taps][33]
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * cit;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long total = (long long)Cout * Cin * taps;
  const int nci = min(cit, Cin - ci0);
  const int rows = nci * taps;                          // row r <-> (ci = r / taps, tap = r % taps) in OUTPUT order
  const bool co_ok = co0 + lane < Cout;
  for (int r0 = wid; r0 < rows; r0 += 64) {             // 8 independent rows in flight per warp
    const float* src[8];
    float s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = min(r0 + 8 * u, rows - 1);
      const int ci = ci0 + r / taps, tap = r % taps;
      src[u] = partial + ((long long)tap * Cin + ci) * Cout + co0 + lane;
      s[u] = 0.f;
    }
    if (co_ok) {
      for (int k = 0; k < splits; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] += __ldg(src[u] + (long long)k * total);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + 8 * u;
      if (r < rows) tile[r * 33 + lane] = s[u];
    }
  }
  __syncthreads();
  for (int co = wid; co < 32 && co0 + co < Cout; co += 8) {
    float* dst = dw + ((long long)(co0 + co) * Cin + ci0) * taps;
    for (int r = lane; r < rows; r += 32) dst[r] = tile[r * 33 + co];
  }
}

// small layers (few tiles, many splits): one thread per output element keeps the grid wide
__global__ void wgrad_tc_reduce_flat_kernel(const float* __restrict__ partial, int splits, int Cout, int Cin, int taps, float* __restrict__ dw) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const int ci = (int)((i / Cout) % Cin);
    const int tap = (int)(i / ((long long)Cout * Cin));
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += partial[(long long)k * total + i];
    dw[((long long)co * Cin + ci) * taps + tap] = s;
  }
}

}  // namespace rho
#include "conv_wgrad_halo.cuh"
namespace rho {

static int wg_pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int wg_ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int wg_block_n(int Cout) {
  for (int bn : {256, 128, 64}) if (Cout % bn == 0) return bn;
  return 0;
}

bool conv_wgrad_tc_supported(const ConvGeom& g, const rho_conv_bwd_params* p) {
  const char* e = getenv("RHO_DISABLE_TCGEN05");
  if (e && e[0] == '1') return false;
  if (p->fwd.act_dtype != RHO_BF16) return false;
  const bool plain = g.sd == 1 && g.sh == 1 && g.sw == 1 && g.ud == 1 && g.uh == 1 && g.uw == 1;
  const bool strided = g.sd == 1 && (g.sh == 2 || g.sw == 2) && g.ud == 1 && g.uh == 1 && g.uw == 1;
  const bool upsampled = g.sd == 1 && g.sh == 1 && g.sw == 1 && g.ud == 1 && (g.uh == 2 || g.uw == 2);
  if (!(plain || strided || upsampled)) return false;
  if (plain && g.taps <= 28 && get_encode() != nullptr && wg_halo_plan(g).ok) return true;     // halo kernel (takes 32-channel blocks)
  if (!plain && g.C2 != 0) return false;
  if (strided && ((g.sh == 2 && g.H < 2) || (g.sw == 2 && g.W < 2))) return false;   // an empty parity view cannot be encoded
  if (g.Cin % 64 != 0 || g.C1 % 64 != 0 || wg_block_n(g.Cout) == 0) return false;
  if (g.taps > 28) return false;
  const int units = (g.taps * (g.Cin / 64) + 1) / 2;
  if (units > 256) return false;
  if (g.N > 32767 || g.D > 32767 || g.H > 32767 || g.W > 32767) return false;
  return get_encode() != nullptr;
}

static void wg_plan(const ConvGeom& g, WgParams& wp) {
  memset(&wp, 0, sizeof(wp));
  const bool up = g.uh == 2 || g.uw == 2;
  const bool strided = g.sh == 2 || g.sw == 2;
  wp.mode = up ? 2 : (strided ? 1 : 0);
  wp.ncls = up ? g.uh * g.uw : 1;
  // position tiles walk the output grid; for a fused upsample, the un-upsampled grid (= each output-parity class)
  const int Gw = up ? g.W : g.Wo, Gh = up ? g.H : g.Ho, Gd = up ? g.D : g.Do;
  int rem = 128;
  wp.bw = std::min(rem, wg_pow2_ceil(Gw)); rem /= wp.bw;
  wp.bh = std::min(rem, wg_pow2_ceil(Gh)); rem /= wp.bh;
  wp.bd = std::min(rem, wg_pow2_ceil(Gd)); rem /= wp.bd;
  wp.bn = rem;
  wp.lbw = wg_ilog2(wp.bw); wp.lbh = wg_ilog2(wp.bh); wp.lbd = wg_ilog2(wp.bd);
  wp.tiles_w = (int)ceil_div(Gw, wp.bw); wp.tiles_h = (int)ceil_div(Gh, wp.bh); wp.tiles_d = (int)ceil_div(Gd, wp.bd);
  wp.tiles_n = (int)ceil_div(g.N, wp.bn);
  wp.tiles_total = wp.tiles_w * wp.tiles_h * wp.tiles_d * wp.tiles_n;
  wp.Cout = g.Cout; wp.Cin = g.Cin; wp.taps = g.taps; wp.C1 = g.C1;
  const int bn = wg_block_n(g.Cout);
  wp.n_co_blocks = g.Cout / bn;
  for (int t = 0; t < g.taps; ++t) {
    const int kw = t % g.kw, kh = (t / g.kw) % g.kh, kd = t / (g.kw * g.kh);
    const int ed = kd - g.pd, eh = kh - g.ph, ew = kw - g.pw;
    wp.tap_xmap[t] = 0;
    for (int cls = 0; cls < wp.ncls; ++cls) {
      int dh = eh, dw = ew;
      if (up) {
        const int a = g.uw == 2 ? cls / 2 : cls, b = cls % 2;      // class = a * uw + b
        if (g.uh == 2) dh = (a + eh) >> 1;                          // floor((a + eh) / 2)
        if (g.uw == 2) dw = ((g.uh == 2 ? b : cls) + ew) >> 1;
      } else if (strided) {
        int ph = 0, pw = 0;
        if (g.sh == 2) { ph = eh & 1; dh = eh < 0 ? -1 : 0; }
        if (g.sw == 2) { pw = ew & 1; dw = ew < 0 ? -1 : 0; }
        wp.tap_xmap[t] = (int8_t)(ph * 2 + pw);
      }
      wp.off_d[t][cls] = (int8_t)ed; wp.off_h[t][cls] = (int8_t)dh; wp.off_w[t][cls] = (int8_t)dw;
    }
  }
  // units = consecutive pairs of the (tap, 64-channel block) list: two ci blocks of one tap when the block count is even,
  // the same block of two taps when Cin = 64, and pairs that straddle taps when the block count is odd (Cin = 192, 320, ...)
  int u = 0;
  {
    const int nblk = g.Cin / 64, total = g.taps * nblk;
    for (int i = 0; i < total; i += 2, ++u) {
      const int j = std::min(i + 1, total - 1);        // odd total: the last unit duplicates its atom (skipped by the epilogue)
      wp.unit_tap[0][u] = (int16_t)(i / nblk); wp.unit_cib[0][u] = (int16_t)(i % nblk);
      wp.unit_tap[1][u] = (int16_t)(j / nblk); wp.unit_cib[1][u] = (int16_t)(j % nblk);
    }
  }
  wp.m_units = u;
  long long ctas = (long long)wp.m_units * wp.n_co_blocks;
  long long want = ceil_div((long long)sm_count() * 2, ctas);
  const long long steps = (long long)wp.tiles_total * wp.ncls;
  long long max_splits = std::max<long long>(1, steps / 4);
  long long s = std::max<long long>(1, std::min(want, max_splits));
  if (s > 64) s = 64;
  wp.tiles_per_split = (int)ceil_div(steps, s);
  wp.splits = (int)ceil_div(steps, wp.tiles_per_split);
}

size_t conv_wgrad_tc_workspace(const ConvGeom& g, const rho_conv_bwd_params* p) {
  if (!conv_wgrad_tc_supported(g, p)) return 0;
  {
    const WgHaloPlan hp = wg_halo_plan(g);
    if (hp.ok) return (size_t)hp.p.splits * g.Cout * g.Cin * g.taps * sizeof(float);
  }
  WgParams wp;
  wg_plan(g, wp);
  return (size_t)wp.splits * g.Cout * g.Cin * g.taps * sizeof(float);
}

// view of a channels-last [N][D][H][W][C] bf16 tensor keeping, in H and W, the elements fh * j + ph resp. fw * j + pw
static int wg_encode(CUtensorMap* m, const void* base, int C, int W, int H, int D, int N, const WgParams& wp, int fh = 1, int ph = 0,
                     int fw = 1, int pw = 0) {
  auto enc = get_encode();
  const int Wv = (W - pw + fw - 1) / fw, Hv = (H - ph + fh - 1) / fh;
  base = (const char*)base + ((long long)ph * W + pw) * C * 2;
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2 * fw, (cuuint64_t)C * 2 * W * fh, (cuuint64_t)C * 2 * W * H, (cuuint64_t)C * 2 * W * H * D};
  cuuint32_t bdim[5] = {64, (cuuint32_t)wp.bw, (cuuint32_t)wp.bh, (cuuint32_t)wp.bd, (cuuint32_t)wp.bn};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad) failed: %d", (int)r);
  return RHO_OK;
}

template <int BN, int STAGES>
static int wg_launch(const WgMaps& maps, const WgParams& wp, float* partial, cudaStream_t st) {
  constexpr size_t SMEM = (size_t)STAGES * (2 + BN / 64) * 128 * 128 + 1024;
  auto kern = conv_wgrad_tc_kernel<BN, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "conv_wgrad_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  long long ctas = (long long)wp.m_units * wp.n_co_blocks * wp.splits;
  kern<<<(unsigned)ctas, 192, SMEM, st>>>(maps, wp, partial);
  RHO_CHECK_LAUNCH("conv_wgrad_tcgen05");
  return RHO_OK;
}

static int wg_reduce(const ConvGeom& g, const float* partial, int splits, float* dw, cudaStream_t st);

int conv_wgrad_tc(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  {
    const WgHaloPlan hp = wg_halo_plan(g);
    if (hp.ok) {
      const size_t need_h = (size_t)hp.p.splits * g.Cout * g.Cin * g.taps * sizeof(float);
      if (!p->workspace || p->workspace_bytes < need_h) return fail(RHO_ERR_WORKSPACE, "conv_wgrad_halo: workspace %zu < %zu", p->workspace_bytes, need_h);
      int rc = conv_wgrad_halo_run(hp, g, p, (float*)p->workspace, st);
      if (rc) return rc;
      return wg_reduce(g, (const float*)p->workspace, hp.p.splits, p->dw, st);
    }
  }
  WgParams wp;
  wg_plan(g, wp);
  const size_t need = (size_t)wp.splits * g.Cout * g.Cin * g.taps * sizeof(float);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "conv_wgrad_tc: workspace %zu < %zu", p->workspace_bytes, need);
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = RHO_OK;
  if (wp.mode == 1) {
    for (int m = 0; m < 4 && !rc; ++m) {
      const int ph = m >> 1, pw = m & 1;
      if ((ph && g.sh != 2) || (pw && g.sw != 2)) continue;
      rc = wg_encode(&maps.x[m], p->fwd.x, g.C1, g.W, g.H, g.D, g.N, wp, g.sh, ph, g.sw, pw);
    }
  } else {
    rc = wg_encode(&maps.x[0], p->fwd.x, g.C1, g.W, g.H, g.D, g.N, wp);
    if (!rc && g.C2) rc = wg_encode(&maps.x[1], p->fwd.x2, g.C2, g.W, g.H, g.D, g.N, wp);
  }
  if (rc) return rc;
  if (wp.mode == 2) {
    for (int cls = 0; cls < wp.ncls && !rc; ++cls) {
      const int a = g.uw == 2 ? cls / 2 : cls, b = g.uh == 2 ? cls % 2 : cls;
      rc = wg_encode(&maps.dy[cls], p->dy, g.Cout, g.Wo, g.Ho, g.Do, g.N, wp, g.uh, g.uh == 2 ? a : 0, g.uw, g.uw == 2 ? b : 0);
    }
  } else {
    rc = wg_encode(&maps.dy[0], p->dy, g.Cout, g.Wo, g.Ho, g.Do, g.N, wp);
  }
  if (rc) return rc;
  float* partial = (float*)p->workspace;
  const int bn = wg_block_n(g.Cout);
  if (bn == 256) rc = wg_launch<256, 2>(maps, wp, partial, st);
  else if (bn == 128) rc = wg_launch<128, 3>(maps, wp, partial, st);
  else rc = wg_launch<64, 4>(maps, wp, partial, st);
  if (rc) return rc;
  return wg_reduce(g, partial, wp.splits, p->dw, st);
}

// dw[co][ci][tap] = sum over splits of partial[split][tap][ci][co], fixed order
static int wg_reduce(const ConvGeom& g, const float* partial, int splits, float* dw, cudaStream_t st) {
  const int cit = std::max(1, std::min(g.Cin, 256 / g.taps));
  dim3 rgrid((unsigned)ceil_div(g.Cout, 32), (unsigned)ceil_div(g.Cin, cit));
  if ((long long)rgrid.x * rgrid.y >= 2ll * sm_count()) {
    wgrad_tc_reduce_kernel<<<rgrid, 256, (size_t)cit * g.taps * 33 * sizeof(float), st>>>(partial, splits, g.Cout, g.Cin, g.taps, cit, dw);
  } else {
    const long long total = (long long)g.Cout * g.Cin * g.taps;
    wgrad_tc_reduce_flat_kernel<<<(unsigned)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 8), 256, 0, st>>>(
        partial, splits, g.Cout, g.Cin, g.taps, dw);
  }
  RHO_CHECK_LAUNCH("wgrad_tc_reduce");
  return RHO_OK;
}

}  // namespace rho
