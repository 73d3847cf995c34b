// Backward of QKVAttention / QKVAttentionLegacy (rho_diffusion/models/unet_v2.py:365-436; reached through the
// always-on checkpoint of AttentionBlock, unet_v2.py:334) on the 5th-generation tensor cores (sm_100a).
//
// With S = scale * Q K^T, P = exp(S - lse) (lse from the forward), delta = rowsum(dO o O):
//     dV = P^T dO        dP = dO V^T        dS = P o (dP - delta)        dQ = scale * dS K        dK = scale * dS^T Q
// Two deterministic kernels (no atomics), both recomputing P from the saved log-sum-exp:
//   attn_bwd_dq_kernel : one CTA = 128 queries, loops over 32-key tiles.
//        S  = Q K_j^T  (128x32)  and  dP = dO V_j^T (128x32) -> TMEM, DOUBLE-BUFFERED ([64b, 64b+32) / [64b+32, 64b+64)):
//        the issuer launches tile j+1's pair before it waits for dS_j, so the tensor core works under the exponentials
//        thread r <-> query row r: dS row -> bf16 -> swizzled smem;   dQ += dS K_j  -> TMEM [128,128+ch)
//   attn_bwd_dkv_kernel: one CTA = 128 keys, loops over 32-query tiles (the transposed problem, so that the thread
//        <-> TMEM-lane mapping is again one thread per output row):
//        S^T = K Q_i^T (128x32), dP^T = V dO_i^T -> TMEM, double-buffered the same way
//        P^T, dS^T rows -> bf16 -> smem;   dV += P^T dO_i -> TMEM,   dK += dS^T Q_i -> TMEM
//        (lse / delta are per COLUMN here: staged per tile in shared memory and read as broadcasts)
// Every operand is a TMA box of the channels-last qkv / dout tensors in the 128-byte-swizzled layout: a [rows x 64 ch]
// tile serves as a K-major operand (Q K^T, dO V^T) and, unchanged, as an MN-major B operand (dS K, P^T dO, dS^T Q).
// Out-of-range rows are zero-filled by TMA, which makes every padded contribution exactly 0 without masks.
// Warp roles as in the forward: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = the 128
// row-owner threads.  ch = 64: 256 TMEM columns and <= 100 KB smem per CTA -> two CTAs per SM, so one CTA's
// exponentials (MUFU) overlap the other's MMAs.
#include "tc_common.cuh"

namespace rho {

struct AttnBwdTcParams {
  int T, heads, C;
  int q_col, k_col, v_col, head_stride;
  float scale, scale_log2;
};

constexpr int AB_ROWS = 128;   // rows (queries resp. keys) owned by a CTA
constexpr int AB_COLS = 32;    // columns (keys resp. queries) per inner tile
constexpr int AB_STAGES = 4;   // TMA ring depth of the 32-row operand tiles
constexpr int AB_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2-9: two row-owner warps per TMEM lane quarter (16 columns each)
constexpr float AB_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Per-row softmax statistics in the layout the backward kernels stream: ld[(n, h)][tile of 32 rows][64] with
// [0,32) = lse * log2(e) and [32,64) = delta = rowsum(dO o O); rows >= T are zero (256-byte aligned tiles for bulk copies)
__global__ void attn_bwd_stats_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, const float* __restrict__ lse,
                                      float* __restrict__ ld, int T, int heads, int ch, int tiles, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over (n, h, tile, 32 rows)
  if (i >= total) return;
  const int rr = (int)(i & 31);
  const long long tile_idx = i >> 5;
  const int tile = (int)(tile_idx % tiles);
  const long long nh = tile_idx / tiles;
  const int h = (int)(nh % heads);
  const long long n = nh / heads;
  const int q = tile * 32 + rr;
  float l2 = 0.f, dl = 0.f;
  if (q < T) {
    const long long base = (n * T + q) * (long long)(heads * ch) + h * ch;
    for (int c = 0; c < ch; c += 8) {
      const uint4 a = *reinterpret_cast<const uint4*>(out + base + c);
      const uint4 b = *reinterpret_cast<const uint4*>(dout + base + c);
      const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dl += __uint_as_float(ua[j] << 16) * __uint_as_float(ub[j] << 16) + __uint_as_float(ua[j] & 0xffff0000u) * __uint_as_float(ub[j] & 0xffff0000u);
    }
    l2 = lse[nh * T + q] * AB_LOG2E;
  }
  float* o = ld + tile_idx * 64;
  o[rr] = l2;
  o[32 + rr] = dl;
}

// store 32 consecutive bf16 columns [c, c+32) (c = 0 or 32) of row r into a [128 rows][64 cols] 128B-swizzled K-major tile
__device__ __forceinline__ void store_row_chunk32(uint8_t* tile, int r, int c, const uint32_t (&packed)[16]) {
  uint8_t* rowp = tile + r * 128;
  const int ch0 = c >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int chunk = (ch0 + i) ^ (r & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
  }
}
// 16 consecutive bf16 columns [c, c+16) (c multiple of 16)
__device__ __forceinline__ void store_row_chunk16(uint8_t* tile, int r, int c, const uint32_t (&packed)[8]) {
  uint8_t* rowp = tile + r * 128;
  const int ch0 = c >> 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int chunk = (ch0 + i) ^ (r & 7);
    *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------------------------
// dQ
// ------------------------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(AB_THREADS, CH == 64 ? 2 : 1)
    attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap qkv128, const __grid_constant__ CUtensorMap qkv32,
                       const __grid_constant__ CUtensorMap do128, const AttnBwdTcParams p, const float* __restrict__ ld,
                       bf16* __restrict__ dqkv) {
  constexpr int KB = CH / 64;
  constexpr int ROW_TILE = AB_ROWS * 128;       // [128 rows x 64 ch] bf16
  constexpr int COL_TILE = AB_COLS * 128;       // [32 rows x 64 ch] bf16
  constexpr int Q_BYTES = KB * ROW_TILE, KV_BYTES = KB * COL_TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + Q_BYTES;
  uint8_t* sK = sdO + Q_BYTES;                  // [AB_STAGES][KB][32 x 64]
  uint8_t* sV = sK + AB_STAGES * KV_BYTES;
  uint8_t* sdS = sV + AB_STAGES * KV_BYTES;     // [128][64]: tile j uses the 32-column half j & 1
  __shared__ uint64_t q_full, kv_full[AB_STAGES], kv_empty[AB_STAGES], sdp_ready[2], ds_ready[2], dq_done;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AB_ROWS, h = blockIdx.y, n = blockIdx.z;
  const int nt = (p.T + AB_COLS - 1) / AB_COLS;
  const int qc = p.q_col + h * p.head_stride, kc = p.k_col + h * p.head_stride, vc = p.v_col + h * p.head_stride;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&sdp_ready[s], 1); mbar_init(&ds_ready[s], 256); }
    mbar_init(&dq_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // buffer b: S at [64 b, 64 b + 32), dP at [64 b + 32, 64 b + 64);  dQ at [128, 128 + CH)
  const uint32_t tmem_sdp = tmem_base_s;
  const uint32_t tmem_dq = tmem_base_s + 128;

  // Producer and issuer warps keep warp-uniform control flow; only the asynchronous instructions are issued by one lane.
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&q_full, 2 * Q_BYTES);
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_3d(sQ + kb * ROW_TILE, &qkv128, &q_full, qc + kb * 64, q0, n);
        tma_load_3d(sdO + kb * ROW_TILE, &do128, &q_full, h * CH + kb * 64, q0, n);
      }
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % AB_STAGES;
      mbar_wait(&kv_empty[s], ((uint32_t)(j / AB_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&kv_full[s], 2 * KV_BYTES);
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(sK + s * KV_BYTES + kb * COL_TILE, &qkv32, &kv_full[s], kc + kb * 64, j * AB_COLS, n);
          tma_load_3d(sV + s * KV_BYTES + kb * COL_TILE, &qkv32, &kv_full[s], vc + kb * 64, j * AB_COLS, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AB_COLS >> 3) << 17) | ((uint32_t)(AB_ROWS >> 4) << 24);
    const uint32_t idesc_dq = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B MN-major */ | ((uint32_t)(CH >> 3) << 17) |
                              ((uint32_t)(AB_ROWS >> 4) << 24);
    auto issue_sdp = [&](int j) {
      const int s = j % AB_STAGES, b = j & 1;
      mbar_wait(&kv_full[s], (uint32_t)(j / AB_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ts = tmem_sdp + (uint32_t)(b * 64), tdp = ts + 32;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a = at_desc_kmajor<128>(smem_u32(sQ + kb * ROW_TILE));
          const uint64_t bd = at_desc_kmajor<128>(smem_u32(sK + s * KV_BYTES + kb * COL_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(ts, a + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_s, (kb | k) != 0);
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a = at_desc_kmajor<128>(smem_u32(sdO + kb * ROW_TILE));
          const uint64_t bd = at_desc_kmajor<128>(smem_u32(sV + s * KV_BYTES + kb * COL_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tdp, a + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_s, (kb | k) != 0);
        }
        umma_commit(&sdp_ready[b]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    issue_sdp(0);
    for (int j = 0; j < nt; ++j) {
      const int s = j % AB_STAGES, b = j & 1;
      // TMEM buffer (j+1)&1 was last read for tile j-1, whose ds_ready was awaited in the previous iteration
      if (j + 1 < nt) issue_sdp(j + 1);
      mbar_wait(&ds_ready[b], (uint32_t)(j >> 1) & 1);   // dS_j is in smem; the row owners are done reading S_j / dP_j
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a = at_desc_kmajor<128>(smem_u32(sdS)) + (uint64_t)(4 * b);   // 32-column half b = +64 bytes
#pragma unroll
        for (int k = 0; k < 2; ++k) {             // 16 keys per MMA = two 8-row atoms of the K tile
          const uint64_t bd = at_desc_mnmajor(smem_u32(sK + s * KV_BYTES + k * 16 * 128), COL_TILE);
          umma_f16(tmem_dq, a + (uint64_t)(2 * k), bd, idesc_dq, (j | k) != 0);
        }
        umma_commit(&kv_empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&dq_done);
    __syncwarp();
  } else {
    const int lane_base = 32 * (warp & 3);
    const int half = (warp - 2) >> 2;             // which 16 of the tile's 32 columns this warp owns
    const int r = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const int q = q0 + r;
    const int ntile = (p.T + 31) / 32;
    const float* ldrow = ld + ((((long long)n * p.heads + h) * ntile + (q >> 5)) << 6) + (q & 31);
    const float lse2 = q < p.T ? ldrow[0] : 0.f;
    const float dl = q < p.T ? ldrow[32] : 0.f;
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      mbar_wait(&sdp_ready[b], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      uint32_t vs[16], vd[16];
      tmem_ld16_async(tmem_sdp + lane_addr + (uint32_t)(b * 64 + half * 16), vs);
      tmem_ld16_async(tmem_sdp + lane_addr + (uint32_t)(b * 64 + 32 + half * 16), vd);
      tmem_wait_ld();
      uint32_t packed[8];
#pragma unroll
      for (int e = 0; e < 16; e += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(vs[e]), p.scale_log2, -lse2));
        const float p1 = ex2_approx(fmaf(__uint_as_float(vs[e + 1]), p.scale_log2, -lse2));
        packed[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(vd[e]) - dl), p1 * (__uint_as_float(vd[e + 1]) - dl));
      }
      store_row_chunk16(sdS, r, b * 32 + half * 16, packed);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&ds_ready[b]);
    }
    mbar_wait(&dq_done, 0);
    tc_fence_after();
    bf16* orow = dqkv + ((long long)n * p.T + q) * (3 * p.C) + qc;
#pragma unroll 1
    for (int c = half * (CH / 2); c < (half + 1) * (CH / 2); c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_dq + lane_addr + (uint32_t)c, v);
      if (q < p.T) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[e]) * p.scale, __uint_as_float(v[e + 1]) * p.scale);
          o.y = pack_bf16x2(__uint_as_float(v[e + 2]) * p.scale, __uint_as_float(v[e + 3]) * p.scale);
          o.z = pack_bf16x2(__uint_as_float(v[e + 4]) * p.scale, __uint_as_float(v[e + 5]) * p.scale);
          o.w = pack_bf16x2(__uint_as_float(v[e + 6]) * p.scale, __uint_as_float(v[e + 7]) * p.scale);
          *reinterpret_cast<uint4*>(orow + c + e) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(256u));
}

// ------------------------------------------------------------------------------------------------------------------
// dK, dV
// ------------------------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(AB_THREADS, CH == 64 ? 2 : 1)
    attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap qkv128, const __grid_constant__ CUtensorMap qkv32,
                        const __grid_constant__ CUtensorMap do32, const AttnBwdTcParams p, const float* __restrict__ ld,
                        bf16* __restrict__ dqkv) {
  constexpr int KB = CH / 64;
  constexpr int ROW_TILE = AB_ROWS * 128;
  constexpr int COL_TILE = AB_COLS * 128;
  constexpr int K_BYTES = KB * ROW_TILE, QD_BYTES = KB * COL_TILE;
  constexpr uint32_t TMEM_COLS = CH == 64 ? 256u : 512u;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + K_BYTES;
  uint8_t* sQ = sV + K_BYTES;                   // [AB_STAGES][KB][32 x 64]
  uint8_t* sdO = sQ + AB_STAGES * QD_BYTES;
  uint8_t* sP = sdO + AB_STAGES * QD_BYTES;     // P^T  [128 keys][64]: tile i uses the 32-column half i & 1
  uint8_t* sdS = sP + ROW_TILE;                 // dS^T [128 keys][64]
  __shared__ uint64_t kv_full, qd_full[AB_STAGES], qd_empty[AB_STAGES], st_ready[2], pds_ready[2], acc_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(128) float sLD[AB_STAGES][64];   // per query tile: [0,32) lse*log2e, [32,64) delta (bulk-copied)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * AB_ROWS, h = blockIdx.y, n = blockIdx.z;
  const int ni = (p.T + AB_COLS - 1) / AB_COLS;
  const int qc = p.q_col + h * p.head_stride, kc = p.k_col + h * p.head_stride, vc = p.v_col + h * p.head_stride;

  if (threadIdx.x == 0) {
    mbar_init(&kv_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) { mbar_init(&qd_full[s], 1); mbar_init(&qd_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&st_ready[s], 1); mbar_init(&pds_ready[s], 256); }
    mbar_init(&acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // buffer b: S^T at [64 b, 64 b + 32), dP^T at [64 b + 32, 64 b + 64);  dK at [128, 128 + CH), dV at [128 + CH, 128 + 2 CH)
  const uint32_t tmem_st = tmem_base_s;
  const uint32_t tmem_dk = tmem_base_s + 128;
  const uint32_t tmem_dv = tmem_base_s + 128 + CH;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&kv_full, 2 * K_BYTES);
      for (int kb = 0; kb < KB; ++kb) {
        tma_load_3d(sK + kb * ROW_TILE, &qkv128, &kv_full, kc + kb * 64, k0, n);
        tma_load_3d(sV + kb * ROW_TILE, &qkv128, &kv_full, vc + kb * 64, k0, n);
      }
    }
    __syncwarp();
    for (int i = 0; i < ni; ++i) {
      const int s = i % AB_STAGES;
      mbar_wait(&qd_empty[s], ((uint32_t)(i / AB_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&qd_full[s], 2 * QD_BYTES + 256);
        bulk_load(sLD[s], ld + ((((long long)n * p.heads + h) * ni + i) << 6), 256, &qd_full[s]);
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(sQ + s * QD_BYTES + kb * COL_TILE, &qkv32, &qd_full[s], qc + kb * 64, i * AB_COLS, n);
          tma_load_3d(sdO + s * QD_BYTES + kb * COL_TILE, &do32, &qd_full[s], h * CH + kb * 64, i * AB_COLS, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc_st = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AB_COLS >> 3) << 17) | ((uint32_t)(AB_ROWS >> 4) << 24);
    const uint32_t idesc_acc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B MN-major */ | ((uint32_t)(CH >> 3) << 17) |
                               ((uint32_t)(AB_ROWS >> 4) << 24);
    auto issue_st = [&](int i) {
      const int s = i % AB_STAGES, b = i & 1;
      mbar_wait(&qd_full[s], (uint32_t)(i / AB_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ts = tmem_st + (uint32_t)(b * 64), tdp = ts + 32;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a = at_desc_kmajor<128>(smem_u32(sK + kb * ROW_TILE));
          const uint64_t bd = at_desc_kmajor<128>(smem_u32(sQ + s * QD_BYTES + kb * COL_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(ts, a + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_st, (kb | k) != 0);
        }
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a = at_desc_kmajor<128>(smem_u32(sV + kb * ROW_TILE));
          const uint64_t bd = at_desc_kmajor<128>(smem_u32(sdO + s * QD_BYTES + kb * COL_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tdp, a + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_st, (kb | k) != 0);
        }
        umma_commit(&st_ready[b]);
      }
      __syncwarp();
    };
    mbar_wait(&kv_full, 0);
    issue_st(0);
    for (int i = 0; i < ni; ++i) {
      const int s = i % AB_STAGES, b = i & 1;
      if (i + 1 < ni) issue_st(i + 1);
      mbar_wait(&pds_ready[b], (uint32_t)(i >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ap = at_desc_kmajor<128>(smem_u32(sP)) + (uint64_t)(4 * b);
        const uint64_t ads = at_desc_kmajor<128>(smem_u32(sdS)) + (uint64_t)(4 * b);
#pragma unroll
        for (int k = 0; k < 2; ++k) {             // 16 queries per MMA
          const uint64_t bdo = at_desc_mnmajor(smem_u32(sdO + s * QD_BYTES + k * 16 * 128), COL_TILE);
          umma_f16(tmem_dv, ap + (uint64_t)(2 * k), bdo, idesc_acc, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t bq = at_desc_mnmajor(smem_u32(sQ + s * QD_BYTES + k * 16 * 128), COL_TILE);
          umma_f16(tmem_dk, ads + (uint64_t)(2 * k), bq, idesc_acc, (i | k) != 0);
        }
        umma_commit(&qd_empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&acc_done);
    __syncwarp();
  } else {
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;               // key row
    const int half = (warp - 2) >> 2;             // which 16 of the tile's 32 query columns this warp owns
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    for (int i = 0; i < ni; ++i) {
      const int b = i & 1, s = i % AB_STAGES;
      // the statistics of this query tile landed with its Q / dO tiles (same barrier): pull them into registers
      // before waiting for the MMAs
      mbar_wait(&qd_full[s], (uint32_t)(i / AB_STAGES) & 1);
      float4 l4[4], d4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        l4[e] = *reinterpret_cast<const float4*>(&sLD[s][half * 16 + 4 * e]);
        d4[e] = *reinterpret_cast<const float4*>(&sLD[s][32 + half * 16 + 4 * e]);
      }
      mbar_wait(&st_ready[b], (uint32_t)(i >> 1) & 1);
      tc_fence_after();
      uint32_t vs[16], vd[16];
      tmem_ld16_async(tmem_st + lane_addr + (uint32_t)(b * 64 + half * 16), vs);
      tmem_ld16_async(tmem_st + lane_addr + (uint32_t)(b * 64 + 32 + half * 16), vd);
      tmem_wait_ld();
      uint32_t pp[8], pd[8];
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const float4 lq = l4[e >> 2], dq = d4[e >> 2];
        const float p0 = ex2_approx(fmaf(__uint_as_float(vs[e]), p.scale_log2, -lq.x));
        const float p1 = ex2_approx(fmaf(__uint_as_float(vs[e + 1]), p.scale_log2, -lq.y));
        const float p2 = ex2_approx(fmaf(__uint_as_float(vs[e + 2]), p.scale_log2, -lq.z));
        const float p3 = ex2_approx(fmaf(__uint_as_float(vs[e + 3]), p.scale_log2, -lq.w));
        pp[e >> 1] = pack_bf16x2(p0, p1);
        pp[(e >> 1) + 1] = pack_bf16x2(p2, p3);
        pd[e >> 1] = pack_bf16x2(p0 * (__uint_as_float(vd[e]) - dq.x), p1 * (__uint_as_float(vd[e + 1]) - dq.y));
        pd[(e >> 1) + 1] = pack_bf16x2(p2 * (__uint_as_float(vd[e + 2]) - dq.z), p3 * (__uint_as_float(vd[e + 3]) - dq.w));
      }
      store_row_chunk16(sP, r, b * 32 + half * 16, pp);
      store_row_chunk16(sdS, r, b * 32 + half * 16, pd);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&pds_ready[b]);
    }
    mbar_wait(&acc_done, 0);
    tc_fence_after();
    const int key = k0 + r;
    bf16* krow = dqkv + ((long long)n * p.T + key) * (3 * p.C) + kc;
    bf16* vrow = dqkv + ((long long)n * p.T + key) * (3 * p.C) + vc;
#pragma unroll 1
    for (int c = half * (CH / 2); c < (half + 1) * (CH / 2); c += 32) {
      uint32_t v[32], w[32];
      tmem_ld32_async(tmem_dk + lane_addr + (uint32_t)c, v);
      tmem_ld32_async(tmem_dv + lane_addr + (uint32_t)c, w);
      tmem_wait_ld();
      if (key < p.T) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[e]) * p.scale, __uint_as_float(v[e + 1]) * p.scale);
          o.y = pack_bf16x2(__uint_as_float(v[e + 2]) * p.scale, __uint_as_float(v[e + 3]) * p.scale);
          o.z = pack_bf16x2(__uint_as_float(v[e + 4]) * p.scale, __uint_as_float(v[e + 5]) * p.scale);
          o.w = pack_bf16x2(__uint_as_float(v[e + 6]) * p.scale, __uint_as_float(v[e + 7]) * p.scale);
          *reinterpret_cast<uint4*>(krow + c + e) = o;
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(w[e]), __uint_as_float(w[e + 1]));
          u.y = pack_bf16x2(__uint_as_float(w[e + 2]), __uint_as_float(w[e + 3]));
          u.z = pack_bf16x2(__uint_as_float(w[e + 4]), __uint_as_float(w[e + 5]));
          u.w = pack_bf16x2(__uint_as_float(w[e + 6]), __uint_as_float(w[e + 7]));
          *reinterpret_cast<uint4*>(vrow + c + e) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(TMEM_COLS));
}

// ---- host ---------------------------------------------------------------------------------------------
bool attention_bwd_tc_supported(const rho_attn_bwd_params* p) {
  const char* e = getenv("RHO_DISABLE_TCGEN05");
  if (e && e[0] == '1') return false;
  const rho_attn_params& f = p->fwd;
  if (f.dtype != RHO_BF16) return false;
  if (f.ch != 64 && f.ch != 128) return false;
  if (f.T < 1 || f.N > 65535 || f.heads > 65535) return false;
  return get_encode() != nullptr;
}

static int make_map(CUtensorMap* map, const void* base, int channels, int T, int N, int box_rows) {
  cuuint64_t gdim[3] = {(cuuint64_t)channels, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t gstr[2] = {(cuuint64_t)channels * 2, (cuuint64_t)channels * 2 * (cuuint64_t)T};
  cuuint32_t bdim[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, bdim, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "attention_bwd_tc: cuTensorMapEncodeTiled failed: %d", (int)r);
  return RHO_OK;
}

template <int CH>
static int launch_attn_bwd_tc(const rho_attn_bwd_params* p, const AttnBwdTcParams& ap, const float* ld, cudaStream_t st) {
  constexpr int KB = CH / 64;
  constexpr size_t SMEM_DQ = (size_t)(2 * KB * 16384 + 2 * AB_STAGES * KB * 4096 + 16384) + 1024;
  constexpr size_t SMEM_DKV = (size_t)(2 * KB * 16384 + 2 * AB_STAGES * KB * 4096 + 2 * 16384) + 1024;
  const rho_attn_params& f = p->fwd;
  auto kq = attn_bwd_dq_kernel<CH>;
  auto kkv = attn_bwd_dkv_kernel<CH>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DQ);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_DKV);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "attention_bwd_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  CUtensorMap qkv128, qkv32, do128, do32;
  int rc = make_map(&qkv128, f.qkv, 3 * ap.C, f.T, f.N, AB_ROWS);
  if (!rc) rc = make_map(&qkv32, f.qkv, 3 * ap.C, f.T, f.N, AB_COLS);
  if (!rc) rc = make_map(&do128, p->dout, ap.C, f.T, f.N, AB_ROWS);
  if (!rc) rc = make_map(&do32, p->dout, ap.C, f.T, f.N, AB_COLS);
  if (rc) return rc;
  dim3 grid((unsigned)ceil_div(f.T, AB_ROWS), f.heads, f.N);
  kq<<<grid, AB_THREADS, SMEM_DQ, st>>>(qkv128, qkv32, do128, ap, ld, (bf16*)p->dqkv);
  RHO_CHECK_LAUNCH("attention_bwd_dq_tcgen05");
  kkv<<<grid, AB_THREADS, SMEM_DKV, st>>>(qkv128, qkv32, do32, ap, ld, (bf16*)p->dqkv);
  RHO_CHECK_LAUNCH("attention_bwd_dkv_tcgen05");
  return RHO_OK;
}

// attention_bwd_tc2.cu: 64-column tiles, P / dS in tensor memory, two row-owner groups on alternate tiles (head dim 64)
bool attention_bwd_tc2_supported(const rho_attn_bwd_params* p);
size_t attention_bwd_tc2_workspace(const rho_attn_bwd_params* p);
int attention_bwd_tc2(const rho_attn_bwd_params* p, cudaStream_t st);

size_t attention_bwd_tc_workspace(const rho_attn_bwd_params* p) {
  if (attention_bwd_tc2_supported(p)) return attention_bwd_tc2_workspace(p);
  const rho_attn_params& f = p->fwd;
  return (size_t)f.N * f.heads * ((f.T + 31) / 32) * 64 * sizeof(float);
}

// per-row statistics for `tiles` 32-row tiles per (sample, head) (rows >= T are zero)
int attn_bwd_stats(const rho_attn_bwd_params* p, float* ld, int tiles, cudaStream_t st) {
  const rho_attn_params& f = p->fwd;
  const long long total = (long long)f.N * f.heads * tiles * 32;
  attn_bwd_stats_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>((const bf16*)f.out, (const bf16*)p->dout, f.lse, ld, f.T, f.heads, f.ch,
                                                                       tiles, total);
  RHO_CHECK_LAUNCH("attention_bwd_stats");
  return RHO_OK;
}

int attention_bwd_tc(const rho_attn_bwd_params* p, cudaStream_t st) {
  if (attention_bwd_tc2_supported(p)) return attention_bwd_tc2(p, st);
  const rho_attn_params& f = p->fwd;
  float* ld = (float*)p->workspace;
  {
    int rc = attn_bwd_stats(p, ld, (f.T + 31) / 32, st);
    if (rc) return rc;
  }
  AttnBwdTcParams ap;
  ap.T = f.T; ap.heads = f.heads; ap.C = f.heads * f.ch;
  if (f.order == 0) { ap.q_col = 0; ap.k_col = f.ch; ap.v_col = 2 * f.ch; ap.head_stride = 3 * f.ch; }
  else { ap.q_col = 0; ap.k_col = ap.C; ap.v_col = 2 * ap.C; ap.head_stride = f.ch; }
  ap.scale = 1.0f / sqrtf((float)f.ch);
  ap.scale_log2 = ap.scale * AB_LOG2E;
  if (f.ch == 64) return launch_attn_bwd_tc<64>(p, ap, ld, st);
  return launch_attn_bwd_tc<128>(p, ap, ld, st);
}

}  // namespace rho
