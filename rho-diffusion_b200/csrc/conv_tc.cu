// conv_nd on the 5th-generation tensor cores: TMA-fed tcgen05 implicit GEMM with TMEM accumulators (sm_100a).
//
//   D[128 output positions x BLOCK_N out-channels] += A[128 x BLOCK_K] * B[BLOCK_N x BLOCK_K]^T   per (tap, channel block)
//
// * activations are channels-last bf16; an M tile is a box of 128 output positions (bn x bd x bh x bw, all powers of
//   two) of the 5-D tensor (N, D, H, W, C).  For every filter tap the A tile is ONE 5-D TMA box load at the box origin
//   shifted by the tap offset; out-of-bounds coordinates are zero-filled by the TMA unit, which is exactly "same"
//   padding -- no im2col buffer, no halo code, no predication.  The box lands in shared memory row-major (one
//   128-byte row of 64 channels per position) with the 128-byte swizzle the UMMA K-major descriptor expects.
// * the skip concat (unet_v2.py:729) is a second tensor map: the K loop walks the channels of x, then of x2.
// * B tiles come from the packed weights [Cout][taps*Cin] through a 2-D tensor map.
// * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> +bias (+residual) -> bf16 -> 16-byte stores).  smem ring of STAGES (A, B) pairs with
//   full/empty mbarriers; tcgen05.commit releases a stage when the MMAs that read it have retired.
// * two CTAs per SM (<= 96 KB smem, <= 256 TMEM columns each) so that one CTA's epilogue overlaps the other's MMAs.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "conv_geom.cuh"
#include "tc_common.cuh"

namespace rho {

constexpr int TC_BM = 128;
constexpr int TC_MAX_TAPS = 28;
constexpr int TC_MAX_MAPS = 4;

struct TcTap {
  int16_t dn, dd, dh, dw;  // coordinate offsets of this tap (in the coordinate system of its tensor map)
  int32_t map;             // which tensor-map set (parity class for strided inputs)
  int32_t kofs;            // K offset (elements) of channel 0 of this tap in a packed weight row
};

struct TcParams {
  int bn, bd, bh, bw;      // tile box (product 128)
  int lbw, lbh, lbd;       // log2 of bw, bh, bd
  int tiles_w, tiles_h, tiles_d, tiles_n;
  int N, Gd, Gh, Gw;       // extents of the output grid this launch covers
  int os_d, os_h, os_w, oo_d, oo_h, oo_w;   // output position = grid * os + oo
  int Do, Ho, Wo;
  int Cout;                // row pitch (channels) of the output tensor(s)
  int C1, C2;              // K channels per source
  int n_blocks;            // Cout / BLOCK_N
  int out_split;           // > 0: output channels >= out_split go to y2 (pitch Cout - out_split), else y (pitch out_split)
  int tma_store;           // epilogue stages 64-channel chunks in shared memory and writes them with TMA tensor stores
  float* stat_part;        // optional per-tile channel sums of the output (fused GroupNorm statistics), see rho_b200.h
  int stat_rows, stat_row_off;   // rows per sample of stat_part; first row of this launch (parity class of a fused upsample)
  int ntaps;
  TcTap taps[TC_MAX_TAPS];
};

struct TcMaps {
  CUtensorMap a[TC_MAX_MAPS][2];
  CUtensorMap w;
  CUtensorMap w_half;      // the same weights boxed BLOCK_N / 2 rows at a time (cluster launches: each CTA multicasts its half)
  CUtensorMap y[2];        // output view(s) of this launch (TMA-store epilogue): y, and y2 for a split output
};

// UMMA shared-memory matrix descriptor, K-major operand, swizzle = row bytes (128 / 64 / 32)
template <int ROW_BYTES>
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  constexpr uint64_t layout = ROW_BYTES == 128 ? 2 : (ROW_BYTES == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address
  d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * ROW_BYTES) >> 4) << 32;             // stride byte offset: 8 rows
  d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

template <int BLOCK_N>
__device__ __forceinline__ uint32_t umma_idesc() {
  return (1u << 4)                      // D format: F32
         | (1u << 7) | (1u << 10)      // A, B format: BF16
         | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------
// CL = 2: launched as clusters of two CTAs that own neighbouring M tiles of the same output-channel block.  Each CTA loads half of
// the weight tile and multicasts it to both, so the weight bytes cross the L2 -> shared-memory path once per pair; a stage is free
// again when BOTH CTAs' MMAs have retired it (their commits arrive on both CTAs' empty barriers).
template <int BLOCK_N, int BLOCK_K, int STAGES, typename TOut, int CL = 1>
__global__ void __launch_bounds__(192, 2) conv_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p,
                                                      const float* __restrict__ bias, const bf16* __restrict__ residual,
                                                      TOut* __restrict__ y, TOut* __restrict__ y2) {
  constexpr int ROW_BYTES = BLOCK_K * 2;
  constexpr int A_BYTES = TC_BM * ROW_BYTES;
  constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to 1024 (the launch adds 1 KB of slack)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[BLOCK_N];
  __shared__ float sstat[2][4][64][2];   // fused GroupNorm statistics: [chunk parity][epilogue warp][channel][sum, sum of squares]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
  const int bx = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nb = bx % p.n_blocks;
  int mt = CL == 2 ? (bx / p.n_blocks) * 2 + (int)crank : bx / p.n_blocks;      // an odd tile count leaves one all-out-of-range tile
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  const int td = mt % p.tiles_d; mt /= p.tiles_d;
  const int tn = mt;
  const int w0 = tw * p.bw, h0 = th * p.bh, d0 = td * p.bd, n0 = tn * p.bn;
  const int kb_per_tap = (p.C1 + p.C2) / BLOCK_K;
  const int num_kb = p.ntaps * kb_per_tap;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL); }
    mbar_init(&accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)BLOCK_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();     // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer: the whole warp walks the loop (uniform control flow), one elected lane issues =====
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < p.ntaps; ++t) {
      const TcTap tap = p.taps[t];
      for (int kb = 0; kb < kb_per_tap; ++kb) {
        const int c = kb * BLOCK_K;
        const int src = c >= p.C1 ? 1 : 0;
        const int cc = src ? c - p.C1 : c;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_5d(sa, &maps.a[tap.map][src], &full_bar[stage], cc, w0 + tap.dw, h0 + tap.dh, d0 + tap.dd, n0 + tap.dn);
          if (CL == 2)     // this CTA's half of the weight tile, to both CTAs
            tma_load_2d_multicast(sa + A_BYTES + crank * (B_BYTES / 2), &maps.w_half, &full_bar[stage], tap.kofs + c,
                                  nb * BLOCK_N + (int)crank * (BLOCK_N / 2), (uint16_t)3);
          else
            tma_load_2d(sa + A_BYTES, &maps.w, &full_bar[stage], tap.kofs + c, nb * BLOCK_N);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: uniform loop, one elected lane issues =====
    const uint32_t idesc = umma_idesc<BLOCK_N>();
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
      const uint64_t adesc = umma_desc<ROW_BYTES>(sa);
      const uint64_t bdesc = umma_desc<ROW_BYTES>(sa + A_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < BLOCK_K / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
        }
        if (CL == 2) umma_commit_multicast(&empty_bar[stage], (uint16_t)3);
        else umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(&accum_bar);
    __syncwarp();
  } else {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w % 4) .. +31 =====
    const int et = threadIdx.x - 64;   // 0..127
    for (int i = et; i < BLOCK_N; i += 128) sbias[i] = bias ? bias[nb * BLOCK_N + i] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const int iw = r & (p.bw - 1), ih = (r >> p.lbw) & (p.bh - 1), id = (r >> (p.lbw + p.lbh)) & (p.bd - 1),
              in = r >> (p.lbw + p.lbh + p.lbd);
    const int gw = w0 + iw, gh = h0 + ih, gd = d0 + id, gn = n0 + in;
    const bool valid = gw < p.Gw && gh < p.Gh && gd < p.Gd && gn < p.N;
    const long long pos = (((long long)gn * p.Do + (gd * p.os_d + p.oo_d)) * p.Ho + (gh * p.os_h + p.oo_h)) * p.Wo + (gw * p.os_w + p.oo_w);
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if constexpr (BLOCK_N >= 64 && sizeof(TOut) == 2) {
      if (p.tma_store) {
        // Every MMA has completed, so the operand ring is free: stage [128 rows x 64 ch] bf16 chunks there (128-byte
        // swizzle, conflict-free 16-byte stores) and hand each chunk to one TMA tensor store (coalesced, clips the tile
        // overhang, asynchronous).
#pragma unroll 1
        for (int cc = 0; cc < BLOCK_N / 64; ++cc) {
          uint8_t* rowp = smem + cc * (TC_BM * 128) + r * 128;
          const int ch = nb * BLOCK_N + cc * 64;                       // first output channel of the chunk
          const bf16* res = (residual && valid) ? residual + pos * p.Cout + ch : nullptr;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(cc * 64 + hh * 32), v);
            float xs[32];                       // the bf16-rounded outputs of this row (0 outside the tensor): statistics input
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[j + e]) + sbias[cc * 64 + hh * 32 + j + e];
              if (res) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(res + hh * 32 + j));
                const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { o[2 * e] += __uint_as_float(u[e] << 16); o[2 * e + 1] += __uint_as_float(u[e] & 0xffff0000u); }
              }
              uint4 w4;
              __nv_bfloat162 h2;
              h2 = __floats2bfloat162_rn(o[0], o[1]); w4.x = *reinterpret_cast<uint32_t*>(&h2);
              h2 = __floats2bfloat162_rn(o[2], o[3]); w4.y = *reinterpret_cast<uint32_t*>(&h2);
              h2 = __floats2bfloat162_rn(o[4], o[5]); w4.z = *reinterpret_cast<uint32_t*>(&h2);
              h2 = __floats2bfloat162_rn(o[6], o[7]); w4.w = *reinterpret_cast<uint32_t*>(&h2);
              if (p.stat_part) {
                if (!valid) w4 = make_uint4(0u, 0u, 0u, 0u);          // rows clipped by the TMA store must not count
                const uint32_t u[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { xs[j + 2 * e] = __uint_as_float(u[e] << 16); xs[j + 2 * e + 1] = __uint_as_float(u[e] & 0xffff0000u); }
              }
              const int chunk = (hh * 4 + (j >> 3)) ^ (r & 7);
              *reinterpret_cast<uint4*>(rowp + chunk * 16) = w4;
            }
            if (p.stat_part) {
              float s1, s2;
              warp_colsum32(xs, lane, s1, s2);
              sstat[cc & 1][warp & 3][hh * 32 + lane][0] = s1;
              sstat[cc & 1][warp & 3][hh * 32 + lane][1] = s2;
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            const bool second = p.out_split > 0 && ch >= p.out_split;
            tma_store_5d(&maps.y[second ? 1 : 0], smem + cc * (TC_BM * 128), second ? ch - p.out_split : ch, w0, h0, d0, n0);
            bulk_commit_group();
          }
          if (p.stat_part && et < 64 && tn < p.tiles_n) {
            // the four warps' partial sums in a fixed order; one row of stat_part per (sample, tile)
            const float a1 = (sstat[cc & 1][0][et][0] + sstat[cc & 1][1][et][0]) + (sstat[cc & 1][2][et][0] + sstat[cc & 1][3][et][0]);
            const float a2 = (sstat[cc & 1][0][et][1] + sstat[cc & 1][1][et][1]) + (sstat[cc & 1][2][et][1] + sstat[cc & 1][3][et][1]);
            const long long row = (long long)tn * p.stat_rows + p.stat_row_off + ((long long)td * p.tiles_h + th) * p.tiles_w + tw;
            *reinterpret_cast<float2*>(p.stat_part + (row * p.Cout + ch + et) * 2) = make_float2(a1, a2);
          }
        }
        if (et == 0) bulk_wait_group_read0();     // shared memory must stay intact until the stores have read it
        goto epilogue_done;
      }
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0, v);   // warp-collective: all lanes execute
      if (valid) {
        const int ch = nb * BLOCK_N + c0;   // first global output channel of this chunk
        TOut* dst;
        const bf16* res = nullptr;
        if (p.out_split > 0 && ch >= p.out_split) dst = y2 + pos * (p.Cout - p.out_split) + (ch - p.out_split);
        else dst = y + pos * (p.out_split > 0 ? p.out_split : p.Cout) + ch;
        if (residual) res = residual + pos * p.Cout + ch;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = __uint_as_float(v[j + e]) + sbias[c0 + j + e];
          if (res) {
            float rr[4];
            load4(res + j, rr);
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] += rr[e];
          }
          store4(dst + j, o);
        }
      }
    }
  epilogue_done:;
  }
  tc_fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();     // no CTA leaves while its peer may still multicast into it or arrive on its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N));
  }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

static int pick_block_k(int C1, int C2) {
  for (int bk : {64, 32, 16}) {
    if (C1 % bk == 0 && (C2 == 0 || C2 % bk == 0)) return bk;
  }
  return 0;
}
// N tile for a launch with `tiles` M tiles: the widest tile that still gives the grid ~2/3 of the SMs (deep UNet levels at
// batch 1-2 have only 32-128 M tiles), never narrower than 64 when a 64-multiple divides Cout
static int pick_block_n_for_grid(int Cout, long long tiles);

static int pick_block_n(int Cout) {
  for (int bn : {256, 128, 64, 32}) {
    if (Cout % bn == 0) return bn;
  }
  return 0;
}

static int pick_block_n_for_grid(int Cout, long long tiles) {
  const int widest = pick_block_n(Cout);
  if (widest <= 64) return widest;
  const long long want = (long long)sm_count() * 2 / 3;
  int best = widest;
  for (int bn : {256, 128, 64}) {
    if (bn > widest || Cout % bn != 0) continue;
    best = bn;
    if (tiles * (Cout / bn) >= want) break;
  }
  return best;
}

struct ActView {
  const void* base;
  long long dim[5];      // C, W, H, D, N
  long long stride[4];   // bytes, for W, H, D, N
};

static int encode_view(CUtensorMap* m, const ActView& v, const int box[5], int bk) {
  auto enc = get_encode();
  if (!enc) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 5; ++i) { gdim[i] = (cuuint64_t)v.dim[i]; bdim[i] = (cuuint32_t)box[i]; }
  for (int i = 0; i < 4; ++i) gstr[i] = (cuuint64_t)v.stride[i];
  CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(v.base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
  return RHO_OK;
}

static int encode_weight_map(CUtensorMap* m, const void* w, long long Ktot, int rows, int bk, int block_n) {
  auto enc = get_encode();
  if (!enc) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)Ktot, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t bdim[2] = {(cuuint32_t)bk, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return RHO_OK;
}

static void pick_tile(TcParams& p) {
  int rem = TC_BM;
  p.bw = std::min(rem, pow2_ceil(p.Gw)); rem /= p.bw;
  p.bh = std::min(rem, pow2_ceil(p.Gh)); rem /= p.bh;
  p.bd = std::min(rem, pow2_ceil(p.Gd)); rem /= p.bd;
  p.bn = rem;
  p.lbw = ilog2(p.bw); p.lbh = ilog2(p.bh); p.lbd = ilog2(p.bd);
  p.tiles_w = (int)ceil_div(p.Gw, p.bw); p.tiles_h = (int)ceil_div(p.Gh, p.bh); p.tiles_d = (int)ceil_div(p.Gd, p.bd);
  p.tiles_n = (int)ceil_div(p.N, p.bn);
}

// RHO_CONV_CLUSTER=1: CTA pairs share the weight tile of BLOCK_N >= 128 launches through TMA multicast (default off)
static int cluster_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RHO_CONV_CLUSTER"); v = e ? atoi(e) : 0; }
  return v;
}

static int encode_weight_map(CUtensorMap* m, const void* w, long long Ktot, int rows, int bk, int block_n);
static int encode_weight_maps(TcMaps& maps, const void* w, long long Ktot, int rows, int bk, int bn) {
  int rc = encode_weight_map(&maps.w, w, Ktot, rows, bk, bn);
  if (!rc && cluster_mode() && bn >= 128) rc = encode_weight_map(&maps.w_half, w, Ktot, rows, bk, bn / 2);
  return rc;
}

// RHO_CONV_DEEP_RING=0: launches with no more CTAs than SMs keep the two-CTAs-per-SM ring depth (default: they take the deep ring)
static int deep_ring_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RHO_CONV_DEEP_RING"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

template <int BLOCK_N, int BLOCK_K, int STAGES, typename TOut, int CL>
static int launch_tc_st(const TcMaps& maps, const TcParams& p, const float* bias, const bf16* residual, TOut* y, TOut* y2, cudaStream_t st) {
  constexpr int STAGE_BYTES = (TC_BM + BLOCK_N) * BLOCK_K * 2;
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
  auto kern = conv_tc_kernel<BLOCK_N, BLOCK_K, STAGES, TOut, CL>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n;
  long long tiles = (CL == 2 ? 2 * ((m_tiles + 1) / 2) : m_tiles) * p.n_blocks;
  if (tiles > 2147483647LL) return fail(RHO_ERR_UNSUPPORTED, "conv_tc: grid too large");
  if (CL == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)tiles); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps, p, bias, residual, y, y2);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "conv_tc: cluster launch: %s", cudaGetErrorString(e));
    return RHO_OK;
  }
  kern<<<(unsigned)tiles, 192, SMEM, st>>>(maps, p, bias, residual, y, y2);
  RHO_CHECK_LAUNCH("conv_fwd_tcgen05");
  return RHO_OK;
}

template <int BLOCK_N, int BLOCK_K, typename TOut, int CL = 1>
static int launch_tc(const TcMaps& maps, const TcParams& p, const float* bias, const bf16* residual, TOut* y, TOut* y2, cudaStream_t st) {
  constexpr int STAGE_BYTES = (TC_BM + BLOCK_N) * BLOCK_K * 2;
  // ring depth: fill ~96 KB so that two CTAs fit in one SM's 227 KB
  constexpr int STAGES = (96 * 1024 / STAGE_BYTES) < 2 ? 2 : ((96 * 1024 / STAGE_BYTES) > 8 ? 8 : (96 * 1024 / STAGE_BYTES));
  // A launch with no more CTAs than SMs (the deep UNet levels at batch 1-2) has one CTA per SM whatever its footprint, and the loads
  // of a lone CTA are latency-bound (stages in flight x stage bytes / TMA round trip): such a launch takes ~200 KB for its ring.
  constexpr int DEEP = (200 * 1024 / STAGE_BYTES) > 8 ? 8 : (200 * 1024 / STAGE_BYTES);
  if constexpr (CL == 1 && DEEP > STAGES) {
    const long long ctas = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * p.n_blocks;
    if (ctas <= sm_count() && deep_ring_enabled()) return launch_tc_st<BLOCK_N, BLOCK_K, DEEP, TOut, 1>(maps, p, bias, residual, y, y2, st);
  }
  return launch_tc_st<BLOCK_N, BLOCK_K, STAGES, TOut, CL>(maps, p, bias, residual, y, y2, st);
}

// view of the output positions this launch writes: grid coordinate j <-> output index j * os + oo in every dim
static int encode_out_map(CUtensorMap* m, void* base, int C, const TcParams& p) {
  ActView v;
  v.dim[0] = C; v.dim[1] = p.Gw; v.dim[2] = p.Gh; v.dim[3] = p.Gd; v.dim[4] = p.N;
  const long long e = (long long)C * 2;
  v.stride[0] = e * p.os_w; v.stride[1] = e * p.Wo * p.os_h; v.stride[2] = e * p.Wo * p.Ho * p.os_d; v.stride[3] = e * p.Wo * p.Ho * p.Do;
  v.base = (char*)base + (((long long)p.oo_d * p.Ho + p.oo_h) * p.Wo + p.oo_w) * e;
  const int box[5] = {64, p.bw, p.bh, p.bd, p.bn};
  return encode_view(m, v, box, 64);
}

static int tma_store_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RHO_CONV_TMA_STORE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

template <typename TOut>
static int dispatch_tc(int bn, int bk, const TcMaps& maps_in, const TcParams& p_in, const float* bias, const bf16* residual, TOut* y, TOut* y2,
                       cudaStream_t st) {
  TcMaps maps = maps_in;
  TcParams p = p_in;
  p.tma_store = 0;
  const int tma_epi = tma_store_enabled();
  const int c_first = p.out_split > 0 ? p.out_split : p.Cout;
  if (tma_epi && sizeof(TOut) == 2 && bn >= 64 && c_first % 64 == 0 && (p.Cout - c_first) % 64 == 0 && (p.out_split == 0 || y2 != nullptr)) {
    int rc = encode_out_map(&maps.y[0], y, c_first, p);
    if (!rc && p.out_split > 0) rc = encode_out_map(&maps.y[1], y2, p.Cout - p.out_split, p);
    if (rc) return rc;
    p.tma_store = 1;
  }
  if (!p.tma_store) p.stat_part = nullptr;       // (conv_fwd_tc_stats_rows only offers statistics when this path is taken)
  if constexpr (sizeof(TOut) == 2) {
    if (cluster_mode() && bn >= 128) {
#define RHO_TC2(BN, BK) \
  if (bn == BN && bk == BK) return launch_tc<BN, BK, TOut, 2>(maps, p, bias, residual, y, y2, st)
      RHO_TC2(256, 64); RHO_TC2(128, 64); RHO_TC2(256, 32); RHO_TC2(128, 32);
#undef RHO_TC2
    }
  }
#define RHO_TC(BN, BK) \
  if (bn == BN && bk == BK) return launch_tc<BN, BK, TOut>(maps, p, bias, residual, y, y2, st)
  RHO_TC(256, 64); RHO_TC(128, 64); RHO_TC(64, 64); RHO_TC(32, 64);
  RHO_TC(256, 32); RHO_TC(128, 32); RHO_TC(64, 32); RHO_TC(32, 32);
#undef RHO_TC
  return fail(RHO_ERR_UNSUPPORTED, "conv_tc: no kernel for BLOCK_N=%d BLOCK_K=%d", bn, bk);
}


// ------------------------------------------------------------------------------------------------------
// fp32-output variant for the split-bf16 engine (six bf16 products per fp32 product, rho_b200.h: rho_split3_concat).
// The tensor core's fp32 accumulation TRUNCATES: measured on the plain kernel above, the error of a split-bf16 conv grows
// linearly with the contraction length (5e-6 at K = 864 ... 1.5e-4 at K = 27648, tools/fp32_split_error.py) -- each of the K / 16
// accumulating MMAs loses up to one ulp of the running sum, always toward zero.  Here a chain is at most TCF_FLUSH K blocks
// (TCF_FLUSH * BLOCK_K / 16 MMAs) long: the issuer alternates between TWO TMEM accumulators, and the epilogue warps add each
// finished partial sum into fp32 REGISTERS (round-to-nearest FADD) while the tensor core fills the other buffer.
constexpr int TCF_FLUSH = 8;

template <int BLOCK_N, int BLOCK_K, int STAGES>
__global__ void __launch_bounds__(192, 1) conv_tc_f32_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams p,
                                                          const float* __restrict__ bias, float* __restrict__ y) {
  constexpr int ROW_BYTES = BLOCK_K * 2;
  constexpr int A_BYTES = TC_BM * ROW_BYTES;
  constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], part_ready[2], part_drained[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[BLOCK_N];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = (int)blockIdx.x % p.n_blocks;
  int mt = (int)blockIdx.x / p.n_blocks;
  const int tw = mt % p.tiles_w; mt /= p.tiles_w;
  const int th = mt % p.tiles_h; mt /= p.tiles_h;
  const int td = mt % p.tiles_d; mt /= p.tiles_d;
  const int tn = mt;
  const int w0 = tw * p.bw, h0 = th * p.bh, d0 = td * p.bd, n0 = tn * p.bn;
  const int kb_per_tap = (p.C1 + p.C2) / BLOCK_K;
  const int num_kb = p.ntaps * kb_per_tap;
  const int num_parts = (num_kb + TCF_FLUSH - 1) / TCF_FLUSH;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&part_ready[b], 1); mbar_init(&part_drained[b], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N < 32 ? 32u : (uint32_t)(2 * BLOCK_N);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < p.ntaps; ++t) {
      const TcTap tap = p.taps[t];
      for (int kb = 0; kb < kb_per_tap; ++kb) {
        const int c = kb * BLOCK_K;
        const int src = c >= p.C1 ? 1 : 0;
        const int cc = src ? c - p.C1 : c;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_5d(sa, &maps.a[tap.map][src], &full_bar[stage], cc, w0 + tap.dw, h0 + tap.dh, d0 + tap.dd, n0 + tap.dn);
          tma_load_2d(sa + A_BYTES, &maps.w, &full_bar[stage], tap.kofs + c, nb * BLOCK_N);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc<BLOCK_N>();
    int stage = 0;
    uint32_t phase = 0;
    for (int part = 0; part < num_parts; ++part) {
      const int b = part & 1;
      if (part >= 2) {                              // the epilogue warps have drained this buffer's previous partial sum
        mbar_wait(&part_drained[b], (uint32_t)((part >> 1) - 1) & 1);
        tc_fence_after();
      }
      const int kb_end = min(num_kb, (part + 1) * TCF_FLUSH);
      for (int kb = part * TCF_FLUSH; kb < kb_end; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t adesc = umma_desc<ROW_BYTES>(sa);
        const uint64_t bdesc = umma_desc<ROW_BYTES>(sa + A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k)
            umma_f16(tmem_base + (uint32_t)(b * BLOCK_N), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                     (kb != part * TCF_FLUSH) || k != 0);
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&part_ready[b]);
      __syncwarp();
    }
  } else {
    const int et = threadIdx.x - 64;   // 0..127
    for (int i = et; i < BLOCK_N; i += 128) sbias[i] = bias ? bias[nb * BLOCK_N + i] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const int iw = r & (p.bw - 1), ih = (r >> p.lbw) & (p.bh - 1), id = (r >> (p.lbw + p.lbh)) & (p.bd - 1),
              in = r >> (p.lbw + p.lbh + p.lbd);
    const int gw = w0 + iw, gh = h0 + ih, gd = d0 + id, gn = n0 + in;
    const bool valid = gw < p.Gw && gh < p.Gh && gd < p.Gd && gn < p.N;
    const long long pos = (((long long)gn * p.Do + (gd * p.os_d + p.oo_d)) * p.Ho + (gh * p.os_h + p.oo_h)) * p.Wo + (gw * p.os_w + p.oo_w);
    float acc[BLOCK_N];
#pragma unroll
    for (int c = 0; c < BLOCK_N; ++c) acc[c] = sbias[c];
    for (int part = 0; part < num_parts; ++part) {
      const int b = part & 1;
      mbar_wait(&part_ready[b], (uint32_t)(part >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(b * BLOCK_N + c0), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) acc[c0 + e] += __uint_as_float(v[e]);
      }
      tc_fence_before();
      mbar_arrive(&part_drained[b]);
    }
    if (valid) {
      float* dst = y + pos * p.Cout + nb * BLOCK_N;
#pragma unroll
      for (int c = 0; c < BLOCK_N; c += 4) {
        const float o[4] = {acc[c], acc[c + 1], acc[c + 2], acc[c + 3]};
        store4(dst + c, o);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

template <int BLOCK_N, int BLOCK_K>
static int launch_tc_f32(const TcMaps& maps, const TcParams& p, const float* bias, float* y, cudaStream_t st) {
  constexpr int STAGE_BYTES = (TC_BM + BLOCK_N) * BLOCK_K * 2;
  constexpr int STAGES = (160 * 1024 / STAGE_BYTES) > 8 ? 8 : (160 * 1024 / STAGE_BYTES);       // one CTA per SM
  constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 1024;
  auto kern = conv_tc_f32_kernel<BLOCK_N, BLOCK_K, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "conv_tc_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const long long tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_d * p.tiles_n * p.n_blocks;
  if (tiles > 2147483647LL) return fail(RHO_ERR_UNSUPPORTED, "conv_tc_f32: grid too large");
  kern<<<(unsigned)tiles, 192, SMEM, st>>>(maps, p, bias, y);
  RHO_CHECK_LAUNCH("conv_fwd_tcgen05_f32");
  return RHO_OK;
}

static int dispatch_tc_f32(int bn, int bk, const TcMaps& maps, const TcParams& p_in, const float* bias, float* y, cudaStream_t st) {
  TcParams p = p_in;
  p.tma_store = 0;
  p.stat_part = nullptr;
#define RHO_TCF(BN, BK) \
  if (bn == BN && bk == BK) return launch_tc_f32<BN, BK>(maps, p, bias, y, st)
  RHO_TCF(128, 64); RHO_TCF(64, 64); RHO_TCF(32, 64);
  RHO_TCF(128, 32); RHO_TCF(64, 32); RHO_TCF(32, 32);
#undef RHO_TCF
  return fail(RHO_ERR_UNSUPPORTED, "conv_tc_f32: no kernel for BLOCK_N=%d BLOCK_K=%d", bn, bk);
}

static bool tc_disabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RHO_DISABLE_TCGEN05");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

}  // namespace rho
#include "conv_halo.cuh"
namespace rho {

// ---- decomposition plans --------------------------------------------------------------------------------
// A fused nearest-x2 upsample turns the 3-tap filter of an upsampled dim into, per OUTPUT parity p, a 2-tap filter on
// the un-upsampled input:  p = 0: {x[i-1]*w0, x[i]*(w1+w2)},  p = 1: {x[i]*(w0+w1), x[i+1]*w2}.  One launch per parity
// class with pre-summed weights ("aux" pack): 12 instead of 27 taps in 3-D (H, W upsampled), 4 instead of 9 in 2-D.
struct DimTaps {
  int n;            // reduced taps
  int off[3];       // input offset of each
  int cnt[3];       // how many original taps are summed
  int src[3][2];    // which
};
static DimTaps dim_taps(int k, int up, int parity) {
  DimTaps d;
  memset(&d, 0, sizeof(d));
  if (k == 1) { d.n = 1; d.off[0] = 0; d.cnt[0] = 1; d.src[0][0] = 0; return d; }
  if (up == 1) {
    d.n = 3;
    for (int i = 0; i < 3; ++i) { d.off[i] = i - 1; d.cnt[i] = 1; d.src[i][0] = i; }
    return d;
  }
  d.n = 2;
  if (parity == 0) { d.off[0] = -1; d.cnt[0] = 1; d.src[0][0] = 0; d.off[1] = 0; d.cnt[1] = 2; d.src[1][0] = 1; d.src[1][1] = 2; }
  else { d.off[0] = 0; d.cnt[0] = 2; d.src[0][0] = 0; d.src[0][1] = 1; d.off[1] = 1; d.cnt[1] = 1; d.src[1][0] = 2; }
  return d;
}

struct AuxTable {
  int nclass, rt, taps, Cin, Cout;
  int cnt[4][12];
  int src[4][12][4];
  int16_t off[4][12][3];   // (dd, dh, dw)
  int par[4][3];           // output parity per class (d, h, w)
};

static bool build_up_table(const ConvGeom& g, AuxTable& t) {
  memset(&t, 0, sizeof(t));
  if (g.ud != 1) return false;               // the UNet never upsamples depth (unet_v2.py:124-129)
  const int cls_h = g.uh == 2 ? 2 : 1, cls_w = g.uw == 2 ? 2 : 1;
  t.nclass = cls_h * cls_w;
  t.taps = g.taps; t.Cin = g.Cin; t.Cout = g.Cout;
  for (int c = 0; c < t.nclass; ++c) {
    const int ph = g.uh == 2 ? c / cls_w : 0, pw = g.uw == 2 ? c % cls_w : 0;
    t.par[c][0] = 0; t.par[c][1] = ph; t.par[c][2] = pw;
    const DimTaps dd = dim_taps(g.kd, 1, 0), dh = dim_taps(g.kh, g.uh, ph), dw = dim_taps(g.kw, g.uw, pw);
    const int rt = dd.n * dh.n * dw.n;
    if (rt > 12) return false;
    t.rt = rt;
    int r = 0;
    for (int a = 0; a < dd.n; ++a)
      for (int b = 0; b < dh.n; ++b)
        for (int e = 0; e < dw.n; ++e, ++r) {
          t.off[c][r][0] = (int16_t)dd.off[a]; t.off[c][r][1] = (int16_t)dh.off[b]; t.off[c][r][2] = (int16_t)dw.off[e];
          int n = 0;
          for (int ia = 0; ia < dd.cnt[a]; ++ia)
            for (int ib = 0; ib < dh.cnt[b]; ++ib)
              for (int ie = 0; ie < dw.cnt[e]; ++ie) {
                if (n >= 4) return false;
                t.src[c][r][n++] = (dd.src[a][ia] * g.kh + dh.src[b][ib]) * g.kw + dw.src[e][ie];
              }
          t.cnt[c][r] = n;
        }
  }
  return true;
}

// aux[class][co][r*Cin + ci] = sum_{k in src(class, r)} w[co][ci][k]   (w: fp32 state-dict layout)
__global__ void pack_aux_kernel(AuxTable t, const float* __restrict__ w, bf16* __restrict__ aux) {
  const long long per_class = (long long)t.Cout * t.rt * t.Cin;
  const long long total = per_class * t.nclass;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / per_class);
    long long rem = i % per_class;
    const int co = (int)(rem / ((long long)t.rt * t.Cin));
    rem %= (long long)t.rt * t.Cin;
    const int r = (int)(rem / t.Cin), ci = (int)(rem % t.Cin);
    float s = 0.f;
    for (int n = 0; n < t.cnt[c][r]; ++n) s += w[((long long)co * t.Cin + ci) * t.taps + t.src[c][r][n]];
    aux[i] = __float2bfloat16_rn(s);
  }
}

static bool tc_common_ok(const ConvGeom& g, const rho_conv_params* p) {
  if (tc_disabled()) return false;
  if (p->act_dtype != RHO_BF16 || p->w_dtype != RHO_BF16) return false;
  // fp32 output (the split-bf16 "exact on tensor cores" engine and nothing else): per-tap kernel, per-thread fp32 stores, no
  // residual (bf16-only in the epilogue), no fused upsample
  if (p->y_f32 && (p->residual != nullptr || !(g.ud == 1 && g.uh == 1 && g.uw == 1))) return false;
  if (pick_block_k(g.C1, g.C2) < 32 || pick_block_n(g.Cout) == 0) return false;
  if (g.taps > TC_MAX_TAPS) return false;
  if (g.N > 32767 || g.D > 32767 || g.H > 32767 || g.W > 32767) return false;
  return get_encode() != nullptr;
}

static bool is_plain(const ConvGeom& g) { return g.sd == 1 && g.sh == 1 && g.sw == 1 && g.ud == 1 && g.uh == 1 && g.uw == 1; }
static bool is_strided(const ConvGeom& g) { return g.sd == 1 && (g.sh == 2 || g.sw == 2) && g.ud == 1 && g.uh == 1 && g.uw == 1; }
static bool is_upsampled(const ConvGeom& g) { return g.sd == 1 && g.sh == 1 && g.sw == 1 && g.ud == 1 && (g.uh == 2 || g.uw == 2); }

int conv_fwd_tc_stats_rows(const ConvGeom& g, const rho_conv_params* p);

bool conv_fwd_tc_supported(const ConvGeom& g, const rho_conv_params* p) {
  if (!tc_common_ok(g, p)) return false;
  if (is_plain(g) || is_strided(g)) return true;
  if (is_upsampled(g)) {
    AuxTable t;
    return p->w_aux != nullptr && build_up_table(g, t);
  }
  return false;
}

size_t conv_aux_weight_bytes(const ConvGeom& g, const rho_conv_params* p) {
  if (!tc_common_ok(g, p) || !is_upsampled(g)) return 0;
  AuxTable t;
  if (!build_up_table(g, t)) return 0;
  return (size_t)t.nclass * t.Cout * t.rt * t.Cin * sizeof(bf16);
}

int conv_pack_weight_aux(const ConvGeom& g, const rho_conv_params* p, const float* w, void* out, cudaStream_t st) {
  AuxTable t;
  if (!is_upsampled(g) || !build_up_table(g, t)) return fail(RHO_ERR_UNSUPPORTED, "conv_pack_weight_aux: no aux form for this geometry");
  long long total = (long long)t.nclass * t.Cout * t.rt * t.Cin;
  int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 8);
  pack_aux_kernel<<<blocks, 256, 0, st>>>(t, w, (bf16*)out);
  RHO_CHECK_LAUNCH("conv_pack_weight_aux");
  return RHO_OK;
}

static ActView plain_view(const void* base, const ConvGeom& g, int C) {
  ActView v;
  v.base = base;
  v.dim[0] = C; v.dim[1] = g.W; v.dim[2] = g.H; v.dim[3] = g.D; v.dim[4] = g.N;
  v.stride[0] = (long long)C * 2; v.stride[1] = v.stride[0] * g.W; v.stride[2] = v.stride[1] * g.H; v.stride[3] = v.stride[2] * g.D;
  return v;
}

// view of the elements with parity (ph, pw) in the strided dims: coordinate j <-> input index 2*j + parity
static ActView parity_view(const void* base, const ConvGeom& g, int C, int ph, int pw) {
  ActView v = plain_view(base, g, C);
  long long off = 0;
  if (g.sw == 2) { v.dim[1] = (g.W - pw + 1) / 2; off += (long long)pw * C * 2; v.stride[0] *= 2; }
  if (g.sh == 2) { v.dim[2] = (g.H - ph + 1) / 2; off += (long long)ph * g.W * C * 2; v.stride[1] *= 2; }
  v.base = (const char*)base + off;
  return v;
}

// rows per sample of rho_conv_params.stats_partial (0: this launch cannot produce the fused GroupNorm statistics)
int conv_fwd_tc_stats_rows(const ConvGeom& g, const rho_conv_params* p) {
  if (!conv_fwd_tc_supported(g, p) || p->y_f32) return 0;
  static int fuse = -1;
  if (fuse < 0) { const char* e = getenv("RHO_FUSE_GN_STATS"); fuse = (e && e[0] == '0') ? 0 : 1; }
  if (!fuse) return 0;
  if (is_plain(g)) {
    const HaloPlan hp = halo_plan(g.N, g.D, g.H, g.W, g.C1, g.C2, g.Cout, g.kd, g.kh, g.kw);
    if (hp.ok) return hp.p.D * hp.p.groups_per_plane;
  }
  if (!tma_store_enabled() || pick_block_n(g.Cout) < 64 || g.Cout % 64 != 0) return 0;
  TcParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.N = g.N;
  int classes = 1;
  if (is_upsampled(g)) {
    AuxTable t;
    if (!build_up_table(g, t)) return 0;
    classes = t.nclass;
    tp.Gd = g.D; tp.Gh = g.H; tp.Gw = g.W;
  } else {
    tp.Gd = g.Do; tp.Gh = g.Ho; tp.Gw = g.Wo;
  }
  pick_tile(tp);
  if (tp.bn != 1) return 0;                       // a tile must not straddle samples
  return classes * tp.tiles_w * tp.tiles_h * tp.tiles_d;
}

int conv_fwd_tc(const ConvGeom& g, const rho_conv_params* p, cudaStream_t st) {
  const int bk = pick_block_k(g.C1, g.C2);
  int bn = pick_block_n(g.Cout);
  TcParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.N = g.N;
  tp.os_d = tp.os_h = tp.os_w = 1;
  tp.Do = g.Do; tp.Ho = g.Ho; tp.Wo = g.Wo;
  tp.Cout = g.Cout; tp.C1 = g.C1; tp.C2 = g.C2;
  {
    // grid-aware N tile (same tile shape for every launch of this conv: the class grid of a fused upsample is g.D/H/W)
    TcParams probe = tp;
    if (is_upsampled(g)) { probe.Gd = g.D; probe.Gh = g.H; probe.Gw = g.W; } else { probe.Gd = g.Do; probe.Gh = g.Ho; probe.Gw = g.Wo; }
    pick_tile(probe);
    bn = pick_block_n_for_grid(g.Cout, (long long)probe.tiles_w * probe.tiles_h * probe.tiles_d * probe.tiles_n);
    if (p->y_f32 && bn > 128) bn = 128;          // the fp32-output kernel keeps BLOCK_N partial sums per thread in registers
  }
  tp.n_blocks = g.Cout / bn;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  float* stat_part = (p->stats_partial && conv_fwd_tc_stats_rows(g, p) > 0) ? p->stats_partial : nullptr;
  tp.stat_part = stat_part;
  tp.stat_rows = stat_part ? conv_fwd_tc_stats_rows(g, p) : 0;

  if (is_upsampled(g)) {
    AuxTable t;
    if (!build_up_table(g, t)) return fail(RHO_ERR_UNSUPPORTED, "conv_tc: upsample plan");
    // one launch per output parity class; the class grid is the un-upsampled input grid
    tp.Gd = g.D; tp.Gh = g.H; tp.Gw = g.W;
    pick_tile(tp);
    const int box[5] = {bk, tp.bw, tp.bh, tp.bd, tp.bn};
    rc = encode_view(&maps.a[0][0], plain_view(p->x, g, g.C1), box, bk);
    if (rc) return rc;
    if (g.C2) { rc = encode_view(&maps.a[0][1], plain_view(p->x2, g, g.C2), box, bk); if (rc) return rc; }
    const long long cls_stride = (long long)t.Cout * t.rt * t.Cin;
    for (int c = 0; c < t.nclass; ++c) {
      tp.os_d = 1; tp.os_h = g.uh; tp.os_w = g.uw;
      tp.oo_d = 0; tp.oo_h = t.par[c][1]; tp.oo_w = t.par[c][2];
      tp.ntaps = t.rt;
      for (int r = 0; r < t.rt; ++r) {
        tp.taps[r].dn = 0; tp.taps[r].dd = t.off[c][r][0]; tp.taps[r].dh = t.off[c][r][1]; tp.taps[r].dw = t.off[c][r][2];
        tp.taps[r].map = 0; tp.taps[r].kofs = r * g.Cin;
      }
      rc = encode_weight_maps(maps, (const bf16*)p->w_aux + c * cls_stride, (long long)t.rt * g.Cin, g.Cout, bk, bn);
      if (rc) return rc;
      tp.stat_row_off = c * tp.tiles_w * tp.tiles_h * tp.tiles_d;
      rc = dispatch_tc<bf16>(bn, bk, maps, tp, p->bias, (const bf16*)p->residual, (bf16*)p->y, (bf16*)nullptr, st);
      if (rc) return rc;
    }
    return RHO_OK;
  }

  if (is_plain(g) && !p->y_f32) {
    const HaloPlan hp = halo_plan(g.N, g.D, g.H, g.W, g.C1, g.C2, g.Cout, g.kd, g.kh, g.kw);
    if (hp.ok) return conv_halo_run(hp, p->x, p->x2, p->w_packed, p->bias, (const bf16*)p->residual, (bf16*)p->y, nullptr, 0, stat_part, st);
  }
  tp.Gd = g.Do; tp.Gh = g.Ho; tp.Gw = g.Wo;
  pick_tile(tp);
  const int box[5] = {bk, tp.bw, tp.bh, tp.bd, tp.bn};
  if (is_plain(g)) {
    rc = encode_view(&maps.a[0][0], plain_view(p->x, g, g.C1), box, bk);
    if (rc) return rc;
    if (g.C2) { rc = encode_view(&maps.a[0][1], plain_view(p->x2, g, g.C2), box, bk); if (rc) return rc; }
    tp.ntaps = g.taps;
    for (int t = 0; t < g.taps; ++t) {
      const int kw = t % g.kw, kh = (t / g.kw) % g.kh, kd = t / (g.kw * g.kh);
      tp.taps[t].dn = 0;
      tp.taps[t].dd = (int16_t)(kd - g.pd); tp.taps[t].dh = (int16_t)(kh - g.ph); tp.taps[t].dw = (int16_t)(kw - g.pw);
      tp.taps[t].map = 0;
      tp.taps[t].kofs = t * g.Cin;
    }
  } else {
    // stride 2 in H and/or W: input index 2*o + k - pad  ->  (parity, coordinate) of a strided view
    //   k - pad = -1: parity 1, coordinate o - 1;   0: parity 0, coordinate o;   +1: parity 1, coordinate o
    bool used[4] = {false, false, false, false};
    int nt = 0;
    for (int t = 0; t < g.taps; ++t) {
      const int kw = t % g.kw, kh = (t / g.kw) % g.kh, kd = t / (g.kw * g.kh);
      const int eh = kh - g.ph, ew = kw - g.pw;
      int ph = 0, pw = 0, dh = eh, dw = ew;
      if (g.sh == 2) { ph = eh & 1; dh = eh < 0 ? -1 : 0; }
      if (g.sw == 2) { pw = ew & 1; dw = ew < 0 ? -1 : 0; }
      // a parity class with no elements (dim of size 1) contributes nothing
      if ((g.sh == 2 && (g.H - ph + 1) / 2 == 0) || (g.sw == 2 && (g.W - pw + 1) / 2 == 0)) continue;
      const int m = ph * 2 + pw;
      used[m] = true;
      tp.taps[nt].dn = 0; tp.taps[nt].dd = (int16_t)(kd - g.pd); tp.taps[nt].dh = (int16_t)dh; tp.taps[nt].dw = (int16_t)dw;
      tp.taps[nt].map = m; tp.taps[nt].kofs = t * g.Cin;
      ++nt;
    }
    tp.ntaps = nt;
    for (int m = 0; m < 4; ++m) {
      if (!used[m]) continue;
      rc = encode_view(&maps.a[m][0], parity_view(p->x, g, g.C1, m >> 1, m & 1), box, bk);
      if (rc) return rc;
      if (g.C2) { rc = encode_view(&maps.a[m][1], parity_view(p->x2, g, g.C2, m >> 1, m & 1), box, bk); if (rc) return rc; }
    }
  }
  rc = encode_weight_maps(maps, p->w_packed, g.Ktot, g.Cout, bk, bn);
  if (rc) return rc;
  if (p->y_f32) return dispatch_tc_f32(bn, bk, maps, tp, p->bias, (float*)p->y, st);
  return dispatch_tc<bf16>(bn, bk, maps, tp, p->bias, (const bf16*)p->residual, (bf16*)p->y, (bf16*)nullptr, st);
}

// ---- dgrad on the tensor cores ---------------------------------------------------------------------------
// dX = "same" convolution of dY with the spatially flipped, channel-transposed filter: the forward kernel is reused with
//   source = dY (K channels = Cout), output channels = Cin (split into dx | dx2 for a concat input),
//   weights = [Cin][taps][Cout] with tap index reversed (a pure permutation of the packed bf16 forward weights, so it is
//   rebuilt in the workspace on every call).
// Fused-upsample convs: dgrad at the upsampled resolution into a temporary, then a 2x2 (H, W) sum-pool -- the adjoint of
// the nearest-neighbour gather.  Strided (Downsample) convs: one launch per INPUT parity class (a_h, a_w); class a = 0 of a
// stride-2 dim receives only tap k = 1 (from output i), class a = 1 taps k = 0 (from output i + 1) and k = 2 (from i);
// the class grid is written back with output stride 2 / offset a.
// tiled variant: one CTA transposes a [32 co][32 ci] block of one tap through shared memory (coalesced both ways)
__global__ void __launch_bounds__(256) pack_dgrad_tiled_kernel(const bf16* __restrict__ wf, bf16* __restrict__ wd, int Cout, int Cin, int taps) {
  __shared__ bf16 tile[32][34];
  const int t = blockIdx.z, co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int co = co0 + r, ci = ci0 + tx;
    if (co < Cout && ci < Cin) tile[r][tx] = wf[((long long)co * taps + (taps - 1 - t)) * Cin + ci];
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r, co = co0 + tx;
    if (co < Cout && ci < Cin) wd[((long long)ci * taps + t) * Cout + co] = tile[tx][r];
  }
}
__global__ void pack_dgrad_kernel(const bf16* __restrict__ wf, bf16* __restrict__ wd, int Cout, int Cin, int taps) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes wd [ci][t][co]
    const int co = (int)(i % Cout);
    const int t = (int)((i / Cout) % taps);
    const int ci = (int)(i / ((long long)Cout * taps));
    wd[i] = wf[((long long)co * taps + (taps - 1 - t)) * Cin + ci];
  }
}

// dx[n,d,h,w,c] = sum_{a<uh, b<uw} t[n,d,h*uh+a,w*uw+b,c]; channels [0,C1) -> dx, [C1,C) -> dx2
__global__ void sum_pool_split_kernel(const bf16* __restrict__ t, bf16* __restrict__ dx, bf16* __restrict__ dx2, long long NDin, int H, int W,
                                      int uh, int uw, int C, int C1) {
  const int cv = C >> 2;
  const long long total = NDin * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * 4;
    long long pos = i / cv;
    const int w = (int)(pos % W); pos /= W;
    const int h = (int)(pos % H);
    const long long nd = pos / H;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < uh; ++a)
      for (int b = 0; b < uw; ++b) {
        float v[4];
        load4(t + (((nd * (H * uh) + (h * uh + a)) * (long long)(W * uw)) + (w * uw + b)) * C + c, v);
        acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
      }
    const long long o = (nd * H + h) * W + w;
    if (c < C1) store4(dx + o * C1 + c, acc);
    else store4(dx2 + o * (C - C1) + (c - C1), acc);
  }
}

static bool dgrad_tc_shape_ok(const ConvGeom& g, const rho_conv_bwd_params* p) {
  if (tc_disabled()) return false;
  if (p->fwd.act_dtype != RHO_BF16 || p->fwd.w_dtype != RHO_BF16) return false;
  if (!(is_plain(g) || is_upsampled(g) || is_strided(g))) return false;
  if (is_strided(g) && (g.C2 != 0 || (g.sh == 2 && g.kh != 3) || (g.sw == 2 && g.kw != 3))) return false;
  if (pick_block_k(g.Cout, 0) < 32 || pick_block_n(g.Cin) == 0) return false;
  if (g.C2 > 0 && (g.C1 % 32 != 0)) return false;
  if (g.taps > TC_MAX_TAPS) return false;
  if (g.N > 32767 || g.Du > 32767 || g.Hu > 32767 || g.Wu > 32767) return false;
  return get_encode() != nullptr;
}

bool conv_dgrad_tc_supported(const ConvGeom& g, const rho_conv_bwd_params* p) { return dgrad_tc_shape_ok(g, p); }

size_t conv_dgrad_tc_workspace(const ConvGeom& g, const rho_conv_bwd_params* p) {
  if (!dgrad_tc_shape_ok(g, p)) return 0;
  size_t w = align_up((size_t)g.Cout * g.Cin * g.taps * sizeof(bf16), 1024);
  size_t t = is_upsampled(g) ? align_up((size_t)g.N * g.Du * g.Hu * g.Wu * g.Cin * sizeof(bf16), 1024) : 0;
  return w + t;
}

int conv_dgrad_tc(const ConvGeom& g, const rho_conv_bwd_params* p, cudaStream_t st) {
  const size_t need = conv_dgrad_tc_workspace(g, p);
  if (!p->workspace || p->workspace_bytes < need) return fail(RHO_ERR_WORKSPACE, "conv_dgrad_tc: workspace %zu < %zu", p->workspace_bytes, need);
  bf16* wd = (bf16*)p->workspace;
  bf16* tmp = (bf16*)((char*)p->workspace + align_up((size_t)g.Cout * g.Cin * g.taps * sizeof(bf16), 1024));
  {
    dim3 pgrid((unsigned)ceil_div(g.Cout, 32), (unsigned)ceil_div(g.Cin, 32), (unsigned)g.taps);
    pack_dgrad_tiled_kernel<<<pgrid, 256, 0, st>>>((const bf16*)p->fwd.w_packed, wd, g.Cout, g.Cin, g.taps);
    RHO_CHECK_LAUNCH("pack_dgrad");
  }
  const bool up = is_upsampled(g);
  const int bk = pick_block_k(g.Cout, 0);
  int bn = pick_block_n(g.Cin);
  TcParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.N = g.N;
  tp.Gd = g.Du; tp.Gh = g.Hu; tp.Gw = g.Wu;       // dgrad output grid = (upsampled) input grid = dY grid for stride 1
  tp.os_d = tp.os_h = tp.os_w = 1;
  tp.Do = g.Du; tp.Ho = g.Hu; tp.Wo = g.Wu;
  tp.Cout = g.Cin; tp.C1 = g.Cout; tp.C2 = 0;
  if (!is_strided(g)) {
    TcParams probe = tp;
    pick_tile(probe);
    bn = pick_block_n_for_grid(g.Cin, (long long)probe.tiles_w * probe.tiles_h * probe.tiles_d * probe.tiles_n);
  }
  tp.n_blocks = g.Cin / bn;
  tp.out_split = (!up && g.C2 > 0) ? g.C1 : 0;
  if (is_strided(g)) {
    TcMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = encode_weight_maps(maps, wd, (long long)g.taps * g.Cout, g.Cin, bk, bn);
    if (rc) return rc;
    ActView v;
    v.base = p->dy;
    v.dim[0] = g.Cout; v.dim[1] = g.Wo; v.dim[2] = g.Ho; v.dim[3] = g.Do; v.dim[4] = g.N;
    v.stride[0] = (long long)g.Cout * 2; v.stride[1] = v.stride[0] * g.Wo; v.stride[2] = v.stride[1] * g.Ho; v.stride[3] = v.stride[2] * g.Do;
    tp.Do = g.D; tp.Ho = g.H; tp.Wo = g.W;
    for (int ah = 0; ah < g.sh; ++ah)
      for (int aw = 0; aw < g.sw; ++aw) {
        tp.Gd = g.D; tp.Gh = (g.H - ah + g.sh - 1) / g.sh; tp.Gw = (g.W - aw + g.sw - 1) / g.sw;
        if (tp.Gh <= 0 || tp.Gw <= 0) continue;
        tp.os_d = 1; tp.os_h = g.sh; tp.os_w = g.sw;
        tp.oo_d = 0; tp.oo_h = ah; tp.oo_w = aw;
        pick_tile(tp);
        int nt = 0;
        for (int t = 0; t < g.taps; ++t) {
          const int kw = t % g.kw, kh = (t / g.kw) % g.kh, kd = t / (g.kw * g.kh);
          int dh, dw;
          if (g.sh == 2) { if (((kh + ah) & 1) == 0) continue; dh = (ah == 1 && kh == 0) ? 1 : 0; } else dh = g.ph - kh;
          if (g.sw == 2) { if (((kw + aw) & 1) == 0) continue; dw = (aw == 1 && kw == 0) ? 1 : 0; } else dw = g.pw - kw;
          tp.taps[nt].dn = 0; tp.taps[nt].dd = (int16_t)(g.pd - kd); tp.taps[nt].dh = (int16_t)dh; tp.taps[nt].dw = (int16_t)dw;
          tp.taps[nt].map = 0;
          tp.taps[nt].kofs = (g.taps - 1 - t) * g.Cout;     // pack_dgrad_kernel stores the taps reversed
          ++nt;
        }
        tp.ntaps = nt;
        const int box[5] = {bk, tp.bw, tp.bh, tp.bd, tp.bn};
        rc = encode_view(&maps.a[0][0], v, box, bk);
        if (rc) return rc;
        rc = dispatch_tc<bf16>(bn, bk, maps, tp, nullptr, nullptr, (bf16*)p->dx, (bf16*)nullptr, st);
        if (rc) return rc;
      }
    return RHO_OK;
  }
  if (!up) {
    // plain stride-1 conv of dY with the flipped filter: the halo-reuse kernel takes it when the shape qualifies
    const HaloPlan hp = halo_plan(g.N, g.Do, g.Ho, g.Wo, g.Cout, 0, g.Cin, g.kd, g.kh, g.kw);
    if (hp.ok) return conv_halo_run(hp, p->dy, nullptr, wd, nullptr, nullptr, (bf16*)p->dx, (bf16*)p->dx2, tp.out_split, nullptr, st);
  }
  pick_tile(tp);
  tp.ntaps = g.taps;
  for (int t = 0; t < g.taps; ++t) {
    const int kw = t % g.kw, kh = (t / g.kw) % g.kh, kd = t / (g.kw * g.kh);
    tp.taps[t].dn = 0;
    tp.taps[t].dd = (int16_t)(kd - g.pd); tp.taps[t].dh = (int16_t)(kh - g.ph); tp.taps[t].dw = (int16_t)(kw - g.pw);
    tp.taps[t].map = 0;
    tp.taps[t].kofs = t * g.Cout;
  }
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int box[5] = {bk, tp.bw, tp.bh, tp.bd, tp.bn};
  ActView v;
  v.base = p->dy;
  v.dim[0] = g.Cout; v.dim[1] = g.Wo; v.dim[2] = g.Ho; v.dim[3] = g.Do; v.dim[4] = g.N;
  v.stride[0] = (long long)g.Cout * 2; v.stride[1] = v.stride[0] * g.Wo; v.stride[2] = v.stride[1] * g.Ho; v.stride[3] = v.stride[2] * g.Do;
  int rc = encode_view(&maps.a[0][0], v, box, bk);
  if (rc) return rc;
  rc = encode_weight_maps(maps, wd, (long long)g.taps * g.Cout, g.Cin, bk, bn);
  if (rc) return rc;
  if (!up) return dispatch_tc<bf16>(bn, bk, maps, tp, nullptr, nullptr, (bf16*)p->dx, (bf16*)p->dx2, st);
  rc = dispatch_tc<bf16>(bn, bk, maps, tp, nullptr, nullptr, tmp, (bf16*)nullptr, st);
  if (rc) return rc;
  long long total = (long long)g.N * g.D * g.H * g.W * (g.Cin >> 2);
  int blocks = (int)std::min<long long>(ceil_div(total, 256), (long long)sm_count() * 16);
  sum_pool_split_kernel<<<blocks, 256, 0, st>>>(tmp, (bf16*)p->dx, (bf16*)p->dx2, (long long)g.N * g.D, g.H, g.W, g.uh, g.uw, g.Cin, g.C1);
  RHO_CHECK_LAUNCH("sum_pool_split");
  return RHO_OK;
}

}  // namespace rho
