// Halo-reuse weight gradient (included by conv_wgrad_tc.cu) for stride-1 3x3(x3) convs with 64 / 128 output channels on
// wide planes -- the layers where the per-tap kernel is L2 -> shared-memory bound (two X boxes + one dY box = 48 KB of TMA
// traffic per 8 N=64 MMAs, 3.7x what the TMA path sustains).
//
// As in conv_halo.cuh a plane is indexed by the padded-flat position q = h * (W + 1) + w + 1.  One TMA slab of X (rows
// [q0 - Wp - 1, q0 + CH + Wp + 1), 64 input channels) and one slab of dY (rows [q0, q0 + CH), 64 output channels; the w = -1
// column is zero-filled, so padding positions contribute nothing) serve ALL NINE in-plane taps of one depth offset:
//     dW[dd, dh, dw][ci][co] += sum_q X[q + dh * Wp + dw][ci] * dY[q][co]
// Both operands are MN-major (the reduction index q is the row; the swizzle is a function of the absolute shared-memory
// address, so an operand may start at any 128-byte row).  An M = 128 MMA pairs two taps: its second 64-channel atom is the
// same slab (shift_b - shift_a) rows further (leading-dimension byte offset of the descriptor).  Five tap pairs x 64
// columns of TMEM accumulate over a contiguous range of position chunks (split-K over chunks); the fp32 partials use the
// layout of the per-tap kernel ([split][tap][ci][co]) and are reduced by the same deterministic kernels.
#pragma once

namespace rho {

constexpr int WH_CH = 256;        // flat positions per chunk
constexpr int WH_PAIRS = 5;       // 9 in-plane taps -> 5 M = 128 pairs (the last one half used)

struct WgHaloParams {
  int N, D, H, W, Wp;
  int C1, C2, Cin, Cout;
  int kd, pd;
  int k1;                      // 1x1(x1) conv: one tap, shift 0 (first atom of group 0; the rest of the MMA rows are unused)
  int ck, cn;                  // channels per input / output block: 64 or 32 (32: 64-byte rows, SWIZZLE_64B)
  int cibs, cobs;              // blocks of the input / output channels
  int units;                   // kd * cibs * cobs
  int splits;                  // CTAs per unit
  int chunks_per_plane;
  long long chunks_total, chunks_per_split;
  int Rx, Ry;                  // slab rows of the X / dY boxes
  int x_bytes, y_bytes;        // stage sizes (1024-aligned)
  int issuers;                 // MMA issuer threads: 2 (warps 1 and 6 take the even / odd accumulator groups) or 1 (RHO_WGRAD_HALO_ISSUERS)
};

struct WgHaloMaps {
  CUtensorMap x[2];            // x, x2 (concat partner)
  CUtensorMap dy;
};

// MN-major operand, rows of ROW_BYTES (128: SWIZZLE_128B, 64: SWIZZLE_64B), eight-row atoms contiguous
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t wgh_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61;
  return d;
}

// CK input channels per unit: 64 -> an M = 128 MMA pairs two taps (WH_PAIRS groups); 32 -> it takes the three dw taps of one
// dh (atoms one row apart) plus an unused fourth (3 groups).  CN output channels per unit (N of the MMA).
template <int CK, int CN>
__global__ void __launch_bounds__(224, 1) conv_wgrad_halo_kernel(const __grid_constant__ WgHaloMaps maps, const __grid_constant__ WgHaloParams p,
                                                                 float* __restrict__ partial) {
  constexpr int XROW = CK * 2, YROW = CN * 2;
  constexpr int GROUPS = CK == 64 ? WH_PAIRS : 3;
  static_assert(GROUPS * CN <= 512, "accumulators exceed TMEM");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                          // [2 stages]
  uint8_t* sY = smem + 2 * (size_t)p.x_bytes;  // [2 stages]
  __shared__ uint64_t full_bar[2], empty_bar[2], accum_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int u = blockIdx.x / p.splits;
  const int split = blockIdx.x - u * p.splits;
  const int cob = u % p.cobs; u /= p.cobs;
  const int cib = u % p.cibs;
  const int dd = u / p.cibs;
  const long long c_begin = (long long)split * p.chunks_per_split;
  const long long c_end = min(p.chunks_total, c_begin + p.chunks_per_split);
  const int xc = cib * CK;
  const int src = xc >= p.C1 ? 1 : 0;
  const int xcc = src ? xc - p.C1 : xc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (uint32_t)p.issuers); }
    mbar_init(&accum_bar, (uint32_t)p.issuers);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // chunk -> (sample, plane, first flat position); chunks whose X plane (d + dd - pd) is outside the volume contribute nothing
  auto decode = [&](long long c, int& n, int& d, int& q0, bool& live) {
    const int cp = (int)(c % p.chunks_per_plane);
    const long long plane = c / p.chunks_per_plane;
    d = (int)(plane % p.D);
    n = (int)(plane / p.D);
    q0 = cp * WH_CH;
    const int xd = d + dd - p.pd;
    live = xd >= 0 && xd < p.D;
  };

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (long long c = c_begin; c < c_end; ++c) {
      int n, d, q0; bool live;
      decode(c, n, d, q0, live);
      if (!live) continue;
      const int hx = (q0 + p.Wp - 1) / p.Wp - 2;          // floor((q0 - Wp - 1) / Wp)
      const int hy = q0 / p.Wp;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], (uint32_t)(p.Rx * p.Wp * XROW + p.Ry * p.Wp * YROW));
        tma_load_5d(sX + (size_t)stage * p.x_bytes, &maps.x[src], &full_bar[stage], xcc, -1, hx, d + dd - p.pd, n);
        tma_load_5d(sY + (size_t)stage * p.y_bytes, &maps.dy, &full_bar[stage], cob * CN, -1, hy, d, n);
      }
      __syncwarp();
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 || (warp == 6 && p.issuers == 2)) {
    // ===== MMA issuer(s), lean form (see conv_halo.cuh: the first version rebuilt both descriptors with masks and shifts for each of
    // the 80 MMAs of a chunk and divided by 3 per tap pair -- far more scalar work per MMA than the 45 clk the tensor pipe needs).
    // Two issuers (warps 1 and 6) take the even / odd accumulator groups; both commit on the stage and accumulator barriers. =====
    if (elect_one()) {
      const int g_first = warp == 1 ? 0 : 1, g_step = p.issuers;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(CN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint64_t HI_X = ((uint64_t)(((8u * XROW) >> 4) | (1u << 14) | ((XROW == 128 ? 2u : 4u) << 29))) << 32;
      constexpr uint64_t HI_Y = ((uint64_t)(((8u * YROW) >> 4) | (1u << 14) | ((YROW == 128 ? 2u : 4u) << 29))) << 32;
      constexpr uint32_t XR16 = XROW / 16, YR16 = YROW / 16;
      // per group: row shift of the first atom (in descriptor units) and the leading-dimension field (distance to the second atom)
      int g_off[GROUPS];
      uint32_t g_lbo[GROUPS];
#pragma unroll
      for (int gi = 0; gi < GROUPS; ++gi) {
        int sa;
        uint32_t lbo;
        if (p.k1) { sa = 0; lbo = (uint32_t)XROW; }
        else if (CK == 64) {
          const int ta = 2 * gi, tb = 2 * gi + 1;
          sa = (ta / 3 - 1) * p.Wp + (ta % 3 - 1);
          const int sb = tb < 9 ? (tb / 3 - 1) * p.Wp + (tb % 3 - 1) : sa + 1;
          lbo = (uint32_t)(sb - sa) * (uint32_t)XROW;
        } else { sa = (gi - 1) * p.Wp - 1; lbo = (uint32_t)XROW; }
        g_off[gi] = sa * (int)XR16;
        g_lbo[gi] = ((lbo >> 4) & 0x3FFFu) << 16;
      }
      const uint32_t x_lo0 = (smem_u32(sX) & 0x3FFFFu) >> 4, y_lo0 = ((smem_u32(sY) & 0x3FFFFu) >> 4) | ((1024u >> 4) << 16);
      const uint32_t x_stage = (uint32_t)p.x_bytes >> 4, y_stage = (uint32_t)p.y_bytes >> 4;
      int stage = 0;
      uint32_t phase = 0, acc = 0;
      for (long long c = c_begin; c < c_end; ++c) {
        int n, d, q0; bool live;
        decode(c, n, d, q0, live);
        if (!live) continue;
        const int hx = (q0 + p.Wp - 1) / p.Wp - 2;
        const int hy = q0 / p.Wp;
        const uint32_t x_c = x_lo0 + (uint32_t)stage * x_stage + (uint32_t)(q0 - hx * p.Wp) * XR16;
        const uint64_t yb = HI_Y | (uint64_t)(y_lo0 + (uint32_t)stage * y_stage + (uint32_t)(q0 - hy * p.Wp) * YR16);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
#pragma unroll
        for (int gi = 0; gi < GROUPS; ++gi) {
          if ((gi == 0 || !p.k1) && (g_step == 1 || (gi & 1) == g_first)) {
            const uint64_t xb = HI_X | (uint64_t)((x_c + (uint32_t)g_off[gi]) | g_lbo[gi]);
#pragma unroll
            for (int k = 0; k < WH_CH / 16; ++k)
              umma_f16(tmem_base + (uint32_t)(gi * CN), xb + (uint64_t)(k * 16 * XR16), yb + (uint64_t)(k * 16 * YR16), idesc, k == 0 ? acc : 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        acc = 1;
        if (++stage == 2) { stage = 0; phase ^= 1; }
      }
      umma_commit(&accum_bar);
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 6) {
    // epilogue: accumulator row r = (atom a = r / 64 -> tap 2 * pi + a, ci = r % 64), 64 output channels per pair
    bool any = false;
    for (long long c = c_begin; c < c_end; ++c) {
      int n, d, q0; bool live;
      decode(c, n, d, q0, live);
      any = any || live;
    }
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const int a = r / CK, ci = xc + (r % CK);     // atom (tap within the group) and input channel of this accumulator row
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int taps = p.k1 ? 1 : p.kd * 9;
    const int ngroups = p.k1 ? 1 : GROUPS;
#pragma unroll 1
    for (int gi = 0; gi < ngroups; ++gi) {
      const int t9 = p.k1 ? 0 : (CK == 64 ? 2 * gi + a : 3 * gi + a);
      const bool used = p.k1 ? a == 0 : (CK == 64 ? t9 < 9 : a < 3);   // warp-uniform (a is constant per warp)
      float* dst = partial + ((((long long)split * taps + (dd * 9 + (used ? t9 : 0))) * p.Cin + ci) * p.Cout) + cob * CN;
#pragma unroll 1
      for (int c0 = 0; c0 < CN; c0 += 32) {
        uint32_t v[32];
        if (any) tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(gi * CN + c0), v);
        else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        if (used) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(dst + c0 + e) =
                make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ---- host -----------------------------------------------------------------------------------------------------------
struct WgHaloPlan {
  bool ok;
  WgHaloParams p;
  size_t smem;
};

static WgHaloPlan wg_halo_plan(const ConvGeom& g) {
  WgHaloPlan pl;
  memset(&pl, 0, sizeof(pl));
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("RHO_WGRAD_HALO"); mode = e ? atoi(e) : 1; }
  if (!mode) return pl;
  const bool plain = g.sd == 1 && g.sh == 1 && g.sw == 1 && g.ud == 1 && g.uh == 1 && g.uw == 1;
  const bool k1 = g.kd == 1 && g.kh == 1 && g.kw == 1;
  if (!plain || !(k1 || (g.kh == 3 && g.kw == 3 && (g.kd == 1 || g.kd == 3)))) return pl;
  // 1x1 layers with 64-channel blocks stay on the per-tap kernel (no halo to reuse); the 32-channel ones only fit here
  if (k1 && g.C1 % 64 == 0 && g.C2 % 64 == 0 && g.Cout % 64 == 0) return pl;
  const int ck = (g.C1 % 64 == 0 && g.C2 % 64 == 0) ? 64 : ((g.C1 % 32 == 0 && g.C2 % 32 == 0) ? 32 : 0);
  if (ck == 0 || (g.Cout != 32 && g.Cout != 64 && g.Cout != 128)) return pl;
  const int cn = g.Cout == 32 ? 32 : 64;
  if (g.W < 16 || g.W > 255 || (long long)g.H * (g.W + 1) < 2 * WH_CH) return pl;
  WgHaloParams& p = pl.p;
  p.N = g.N; p.D = g.D; p.H = g.H; p.W = g.W; p.Wp = g.W + 1;
  p.C1 = g.C1; p.C2 = g.C2; p.Cin = g.Cin; p.Cout = g.Cout;
  p.kd = g.kd; p.pd = g.kd / 2;
  p.k1 = k1 ? 1 : 0;
  p.ck = ck; p.cn = cn;
  p.cibs = g.Cin / ck; p.cobs = g.Cout / cn;
  p.units = p.kd * p.cibs * p.cobs;
  p.chunks_per_plane = (int)ceil_div((long long)g.H * p.Wp, WH_CH);
  p.chunks_total = (long long)g.N * g.D * p.chunks_per_plane;
  int splits = std::max(1, sm_count() / p.units);
  splits = (int)std::min<long long>(splits, std::max<long long>(1, p.chunks_total / 4));
  splits = std::min(splits, 64);
  p.chunks_per_split = ceil_div(p.chunks_total, (long long)splits);
  p.splits = (int)ceil_div(p.chunks_total, p.chunks_per_split);
  int Rx = 0, Ry = 0;
  for (int c = 0; c < p.chunks_per_plane; ++c) {
    const int q0 = c * WH_CH;
    const int hx = (q0 + p.Wp - 1) / p.Wp - 2;
    Rx = std::max(Rx, (q0 + WH_CH - 1 + p.Wp + 1) / p.Wp - hx + 1);
    Ry = std::max(Ry, (q0 + WH_CH - 1) / p.Wp - q0 / p.Wp + 1);
  }
  if (Rx > 256 || Ry > 256) return pl;
  p.Rx = Rx; p.Ry = Ry;
  p.x_bytes = (int)align_up((size_t)Rx * p.Wp * (size_t)(2 * ck), 1024);
  p.y_bytes = (int)align_up((size_t)Ry * p.Wp * (size_t)(2 * cn), 1024);
  pl.smem = 2 * ((size_t)p.x_bytes + p.y_bytes) + 1024;
  if (pl.smem > 227 * 1024 - 2048) return pl;
  pl.ok = true;
  return pl;
}

static int wg_halo_encode(CUtensorMap* m, const void* base, int C, const WgHaloParams& p, int rows, int cb) {
  auto enc = get_encode();
  cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.N};
  cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * p.W, (cuuint64_t)C * 2 * p.W * p.H, (cuuint64_t)C * 2 * p.W * p.H * p.D};
  cuuint32_t bdim[5] = {(cuuint32_t)cb, (cuuint32_t)p.Wp, (cuuint32_t)rows, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad halo) failed: %d", (int)r);
  return RHO_OK;
}

static int conv_wgrad_halo_run(const WgHaloPlan& pl, const ConvGeom& g, const rho_conv_bwd_params* bp, float* partial, cudaStream_t st) {
  WgHaloParams p = pl.p;
  static int issuers = -1;
  if (issuers < 0) { const char* e = getenv("RHO_WGRAD_HALO_ISSUERS"); issuers = (e && e[0] == '1') ? 1 : 2; }
  p.issuers = p.k1 ? 1 : issuers;
  WgHaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc = wg_halo_encode(&maps.x[0], bp->fwd.x, g.C1, p, p.Rx, p.ck);
  if (!rc && g.C2) rc = wg_halo_encode(&maps.x[1], bp->fwd.x2, g.C2, p, p.Rx, p.ck);
  if (!rc) rc = wg_halo_encode(&maps.dy, bp->dy, g.Cout, p, p.Ry, p.cn);
  if (rc) return rc;
#define RHO_WGH(CK, CN)                                                                                                       \
  if (p.ck == CK && p.cn == CN) {                                                                                             \
    static size_t attr_bytes = 0;                                                                                             \
    if (pl.smem > attr_bytes) {                                                                                               \
      cudaError_t e = cudaFuncSetAttribute(conv_wgrad_halo_kernel<CK, CN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem); \
      if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "conv_wgrad_halo: cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e)); \
      attr_bytes = pl.smem;                                                                                                   \
    }                                                                                                                         \
    conv_wgrad_halo_kernel<CK, CN><<<(unsigned)(p.units * p.splits), 224, pl.smem, st>>>(maps, p, partial);                   \
    RHO_CHECK_LAUNCH("conv_wgrad_halo_tcgen05");                                                                              \
    return RHO_OK;                                                                                                            \
  }
  RHO_WGH(64, 64) RHO_WGH(64, 32) RHO_WGH(32, 64) RHO_WGH(32, 32)
#undef RHO_WGH
  return fail(RHO_ERR_UNSUPPORTED, "conv_wgrad_halo: no kernel for ck=%d cn=%d", p.ck, p.cn);
}

}  // namespace rho
