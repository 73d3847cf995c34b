// Backward of QKVAttention / QKVAttentionLegacy (rho_diffusion/models/unet_v2.py:365-436), second-generation tcgen05 kernels
// for head dim 64 (the first generation, attention_bwd_tc.cu, stays for head dim 128 and as the switchable fallback).
//
// Same mathematics and the same two deterministic passes as attention_bwd_tc.cu (no atomics):
//     P = exp(S - lse),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  dQ = scale dS K,  dK = scale dS^T Q
//   DKV pass: one CTA = 128 keys (thread <-> key row), loops over 64-query tiles;   DQ pass: one CTA = 128 queries, 64-key tiles.
// What changed, and why (profiles/r01g_ncu_full_attn_bwd_dkv_summary.json, profiles/r02_umma_attn.txt, tools/attn_trace.py):
//   * inner tiles of 64 columns instead of 32: an M = 128, K = 16 tcgen05.mma costs ~45-52 clk whether N is 32 or 64, so the
//     S / dP MMAs of the 32-column kernels ran at 30 % of the tensor peak and the kernels were tensor-unit-bound;
//   * P^T / dS^T (dS) never touch shared memory: the row owners store them as bf16 pairs into TENSOR MEMORY (tcgen05.st) and the
//     dV / dK (dQ) MMAs read their A operand from there (TS form) -- no swizzled st.shared, no async-proxy fence on the chain;
//   * one CTA per SM with all 512 TMEM columns and THREE S / dP buffers in flight.  Measured with two buffers (ncu source page,
//     profiles/r02_attn_bwd2_*): the row owners sat 59 % of their time in the st_ready wait -- the fixed latency of one tile's
//     round trip (MMA start-up, tcgen05.commit -> mbarrier, LDTM ~250 clk, STTM + wait::st, mbarrier -> issuer) is ~1400 clk,
//     twice the tile's MMA time, so throughput = round trip / buffers.  P (bf16) overwrites the S columns its thread has just
//     read and dS the dP columns (in place, no separate P / dS buffers), which is what makes room for the third buffer;
//   * two row-owner groups of eight warps (TWO threads per row: P and dS are pure element-wise functions of S, dP and the saved
//     statistics, so splitting a row is free) take alternate tiles; a group's next tile was issued one event earlier, so in
//     steady state it never waits for an MMA;
//   * the issuer waits for the streamed operands BEFORE it waits for the row owners, and issues everything of an event
//     straight-line from one elected region (the accumulating MMAs of tile i, then S / dP of tile i + 3 into the buffer they free).
// TMEM map: buffer j at 128 j: S at +0..63, dP at +64..127; thread (row, c) overwrites S columns [32 c, 32 c + 16) with its 32 P
// values and dP columns [32 c, 32 c + 16) with its dS values; dK | dQ at 384, dV at 448.
// Out-of-range rows are zero-filled by TMA (and their statistics are zero), which makes every padded contribution exactly 0.
#include "tc_common.cuh"

namespace rho {

struct AttnBwd2Params {
  int T, heads, C;
  int q_col, k_col, v_col, head_stride;
  float scale, scale_log2;
  int lean;                 // MMA issue loop: 1 = one thread, compile-time ring / buffer indices, descriptors advanced by adds; 0 = first version
};

constexpr int B2_ROWS = 128;          // rows owned by a CTA
constexpr int B2_COLS = 64;           // columns per inner tile
constexpr int B2_ST = 6;              // ring stages of the streamed operand tiles (4 tiles are in flight)
constexpr int B2_NB = 3;              // S / dP buffers in tensor memory
constexpr int B2_ROW_TILE = 128 * 128;   // [128 rows x 64 ch] bf16
constexpr int B2_COL_TILE = 64 * 128;    // [64 rows x 64 ch] bf16

__device__ __forceinline__ void b2_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t b2_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// DKV = true : resident A1 = K, A2 = V (rows = keys);   streamed B1 = Q_i, B2 = dO_i;  accumulators dK (384), dV (448)
// DKV = false: resident A1 = Q, A2 = dO (rows = queries); streamed B1 = K_j, B2 = V_j;  accumulator dQ (384)
template <bool DKV>
__global__ void __launch_bounds__(576, 1)
    attn_bwd2_kernel(const __grid_constant__ CUtensorMap qkv128, const __grid_constant__ CUtensorMap qkv64, const __grid_constant__ CUtensorMap do128,
                     const __grid_constant__ CUtensorMap do64, const AttnBwd2Params p, const float* __restrict__ ld, bf16* __restrict__ dqkv) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA1 = smem;
  uint8_t* sA2 = sA1 + B2_ROW_TILE;
  uint8_t* sB1 = sA2 + B2_ROW_TILE;                  // [B2_ST][64 x 64]
  uint8_t* sB2 = sB1 + B2_ST * B2_COL_TILE;
  __shared__ uint64_t a_full, b_full[B2_ST], b_empty[B2_ST], st_ready[B2_NB], pds_ready[B2_NB], acc_final;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(128) float sLD[B2_ST][128];   // DKV: per query tile [lse2 32 | delta 32 | lse2 32 | delta 32] (bulk-copied)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * B2_ROWS, h = blockIdx.y, n = blockIdx.z;
  const int ni = (p.T + B2_COLS - 1) / B2_COLS;
  const int qc = p.q_col + h * p.head_stride, kc = p.k_col + h * p.head_stride, vc = p.v_col + h * p.head_stride;

  if (threadIdx.x == 0) {
    mbar_init(&a_full, 1);
    for (int s = 0; s < B2_ST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int b = 0; b < B2_NB; ++b) { mbar_init(&st_ready[b], 1); mbar_init(&pds_ready[b], 256); }
    mbar_init(&acc_final, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(&a_full, 2 * B2_ROW_TILE);
      if (DKV) {
        tma_load_3d(sA1, &qkv128, &a_full, kc, r0, n);
        tma_load_3d(sA2, &qkv128, &a_full, vc, r0, n);
      } else {
        tma_load_3d(sA1, &qkv128, &a_full, qc, r0, n);
        tma_load_3d(sA2, &do128, &a_full, h * 64, r0, n);
      }
    }
    __syncwarp();
    for (int i = 0; i < ni; ++i) {
      const int s = i % B2_ST;
      mbar_wait(&b_empty[s], ((uint32_t)(i / B2_ST) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&b_full[s], 2 * B2_COL_TILE + (DKV ? 512 : 0));
        if (DKV) {
          b2_bulk_load(sLD[s], ld + ((((long long)n * p.heads + h) * (2 * ni) + 2 * i) << 6), 512, &b_full[s]);
          tma_load_3d(sB1 + s * B2_COL_TILE, &qkv64, &b_full[s], qc, i * B2_COLS, n);
          tma_load_3d(sB2 + s * B2_COL_TILE, &do64, &b_full[s], h * 64, i * B2_COLS, n);
        } else {
          tma_load_3d(sB1 + s * B2_COL_TILE, &qkv64, &b_full[s], kc, i * B2_COLS, n);
          tma_load_3d(sB2 + s * B2_COL_TILE, &qkv64, &b_full[s], vc, i * B2_COLS, n);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1 && p.lean) {
    // ===== MMA issuer, lean form =====
    // The first version spent ~150 dependent uniform-datapath instructions per tile on 16 MMAs (descriptor assembly from rolled ring /
    // buffer indices): ~85 clk to issue an MMA the tensor pipe executes in ~45, and the row owners never made the issuer wait -- the
    // kernel was issue-bound (the same finding as in the halo conv kernel, profiles/r02_halo_lean_issue.txt).  Here one elected thread
    // runs the whole loop, unrolled by 6 tiles so that the ring stage (i % 6) and the S / dP buffer (i % 3) are compile-time constants:
    // every descriptor is a base register pair plus an immediate.  All operands share one descriptor high word (128-byte rows,
    // SWIZZLE_128B, 1024-byte atoms); the low word is the address in 16-byte units (+ the leading-dimension field of the MN-major form).
    if (elect_one()) {
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(B2_COLS >> 3) << 17) | ((uint32_t)(B2_ROWS >> 4) << 24);
      const uint32_t idesc2 = idesc1 | (1u << 16);   // B MN-major
      constexpr uint64_t HI = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;
      const uint64_t a1 = HI | (uint64_t)(((smem_u32(sA1) & 0x3FFFFu) >> 4) | (1u << 16));
      const uint64_t a2 = HI | (uint64_t)(((smem_u32(sA2) & 0x3FFFFu) >> 4) | (1u << 16));
      const uint64_t b1k = HI | (uint64_t)(((smem_u32(sB1) & 0x3FFFFu) >> 4) | (1u << 16));                              // K-major view of Q_i / K_j
      const uint64_t b2k = HI | (uint64_t)(((smem_u32(sB2) & 0x3FFFFu) >> 4) | (1u << 16));
      const uint64_t b1m = HI | (uint64_t)(((smem_u32(sB1) & 0x3FFFFu) >> 4) | ((uint32_t)(B2_COL_TILE >> 4) << 16));    // MN-major view
      const uint64_t b2m = HI | (uint64_t)(((smem_u32(sB2) & 0x3FFFFu) >> 4) | ((uint32_t)(B2_COL_TILE >> 4) << 16));
      constexpr uint32_t STAGE = B2_COL_TILE >> 4;          // descriptor units per ring stage
      auto mma1 = [&](const int s, const int j) {           // S[j] = A1 B1^T, dP[j] = A2 B2^T of the tile in stage s
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(128 * j), a1 + (uint64_t)(2 * k), b1k + (uint64_t)(s * STAGE + 2 * k), idesc1, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(128 * j + 64), a2 + (uint64_t)(2 * k), b2k + (uint64_t)(s * STAGE + 2 * k), idesc1, k != 0);
        umma_commit(&st_ready[j]);
      };
      mbar_wait(&a_full, 0);
#pragma unroll
      for (int u = 0; u < B2_NB; ++u) {
        if (u < ni) {
          mbar_wait(&b_full[u], 0);
          tc_fence_after();
          mma1(u, u);
        }
      }
      uint32_t acc = 0;                                     // 0 only for the first accumulating MMAs of the kernel
      for (int i0 = 0; i0 < ni; i0 += B2_ST) {
        const uint32_t ph6 = (uint32_t)(i0 / B2_ST) & 1;    // ring parity of tiles i0 .. i0 + 5
#pragma unroll
        for (int u = 0; u < B2_ST; ++u) {
          const int i = i0 + u;
          if (i < ni) {
            const int s = u, j = u % B2_NB, sn = (u + B2_NB) % B2_ST;
            if (i + B2_NB < ni) mbar_wait(&b_full[sn], u + B2_NB < B2_ST ? ph6 : ph6 ^ 1);    // not on the critical loop
            mbar_wait(&pds_ready[j], (uint32_t)(i / B2_NB) & 1);   // P / dS of tile i are in TMEM (in place of S / dP[j])
            tc_fence_after();
            // accumulating MMAs of tile i first: the in-order tensor pipe then guarantees they have read P / dS before S / dP of
            // tile i + 3 overwrite the buffer.  K step k covers 16 columns = 8 TMEM columns at 32 (k / 2) + 8 (k % 2).
            if (DKV) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ts(tmem + 448u, tmem + (uint32_t)(128 * j + 32 * (k >> 1) + 8 * (k & 1)), b2m + (uint64_t)(s * STAGE + k * 128), idesc2,
                            k == 0 ? acc : 1u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ts(tmem + 384u, tmem + (uint32_t)(128 * j + 64 + 32 * (k >> 1) + 8 * (k & 1)), b1m + (uint64_t)(s * STAGE + k * 128), idesc2,
                          k == 0 ? acc : 1u);
            acc = 1;
            umma_commit(&b_empty[s]);
            if (i + B2_NB < ni) mma1(sn, j);
            if (i == ni - 1) umma_commit(&acc_final);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (first version, RHO_ATTN_BWD_LEAN=0) =====
    const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(B2_COLS >> 3) << 17) | ((uint32_t)(B2_ROWS >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B MN-major */ | ((uint32_t)(64 >> 3) << 17) |
                            ((uint32_t)(B2_ROWS >> 4) << 24);
    auto mma1 = [&](int i) {                 // S[j] = A1 B1_i^T, dP[j] = A2 B2_i^T into buffer j = i % 3   (called by the elected lane)
      const int s = i % B2_ST, j = i % B2_NB;
      const uint64_t a1 = at_desc_kmajor<128>(smem_u32(sA1)), a2 = at_desc_kmajor<128>(smem_u32(sA2));
      const uint64_t b1 = at_desc_kmajor<128>(smem_u32(sB1 + s * B2_COL_TILE)), b2 = at_desc_kmajor<128>(smem_u32(sB2 + s * B2_COL_TILE));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(128 * j), a1 + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), idesc1, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(128 * j + 64), a2 + (uint64_t)(2 * k), b2 + (uint64_t)(2 * k), idesc1, k != 0);
      umma_commit(&st_ready[j]);
    };
    mbar_wait(&a_full, 0);
    for (int i = 0; i < B2_NB && i < ni; ++i) {
      mbar_wait(&b_full[i % B2_ST], 0);
      tc_fence_after();
      if (elect_one()) mma1(i);
      __syncwarp();
    }
    for (int i = 0; i < ni; ++i) {
      const int s = i % B2_ST, j = i % B2_NB;
      if (i + B2_NB < ni) mbar_wait(&b_full[(i + B2_NB) % B2_ST], (uint32_t)((i + B2_NB) / B2_ST) & 1);   // not on the critical loop
      mbar_wait(&pds_ready[j], (uint32_t)(i / B2_NB) & 1);   // P / dS of tile i are in TMEM (in place of S / dP[j])
      tc_fence_after();
      if (elect_one()) {
        // accumulating MMAs of tile i first: the in-order tensor pipe then guarantees they have read P / dS before S / dP of
        // tile i + 3 overwrite the buffer.  K step k covers 16 columns = 8 TMEM columns at 32 (k / 2) + 8 (k % 2).
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (DKV) {
            const uint64_t bdo = at_desc_mnmajor(smem_u32(sB2 + s * B2_COL_TILE + k * 16 * 128), B2_COL_TILE);
            umma_f16_ts(tmem + 448u, tmem + (uint32_t)(128 * j + 32 * (k >> 1) + 8 * (k & 1)), bdo, idesc2, (i | k) != 0);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t b1 = at_desc_mnmajor(smem_u32(sB1 + s * B2_COL_TILE + k * 16 * 128), B2_COL_TILE);
          umma_f16_ts(tmem + 384u, tmem + (uint32_t)(128 * j + 64 + 32 * (k >> 1) + 8 * (k & 1)), b1, idesc2, (i | k) != 0);
        }
        umma_commit(&b_empty[s]);
        if (i + B2_NB < ni) mma1(i + B2_NB);
        if (i == ni - 1) umma_commit(&acc_final);
      }
      __syncwarp();
    }
  } else {
    // ===== row owners: 16 warps = (group g: tiles i = g, g + 2, ...) x (column half c) x (TMEM lane quarter); two threads per row =====
    const int wid = warp - 2;
    const int g = wid >> 3, c = (wid >> 2) & 1;
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const int row = r0 + r;
    float lse2 = 0.f, dl = 0.f;
    if (!DKV) {                                              // per-ROW statistics (rows = queries)
      const float* ldrow = ld + ((((long long)n * p.heads + h) * (2 * ni) + (row >> 5)) << 6) + (row & 31);
      if (row < p.T) { lse2 = ldrow[0]; dl = ldrow[32]; }
    }
    for (int i = g; i < ni; i += 2) {
      const int s = i % B2_ST, j = i % B2_NB;
      const uint32_t t_s = tmem + lane_addr + (uint32_t)(128 * j + 32 * c), t_dp = t_s + 64u;
      if (DKV) mbar_wait(&b_full[s], (uint32_t)(i / B2_ST) & 1);      // this tile's per-column statistics landed with Q_i / dO_i
      mbar_wait(&st_ready[j], (uint32_t)(i / B2_NB) & 1);
      tc_fence_after();
      uint32_t vs[32], vd[32], pp[16], pd[16];
      tmem_ld32_async(t_s, vs);
      tmem_ld32_async(t_dp, vd);
      tmem_wait_ld();
      const float4* L4 = reinterpret_cast<const float4*>(&sLD[s][64 * c]);
      const float4* D4 = reinterpret_cast<const float4*>(&sLD[s][64 * c + 32]);
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        float4 lq, dq;
        if (DKV) { lq = L4[e >> 2]; dq = D4[e >> 2]; }
        else { lq = make_float4(lse2, lse2, lse2, lse2); dq = make_float4(dl, dl, dl, dl); }
        const float p0 = ex2_approx(fmaf(__uint_as_float(vs[e]), p.scale_log2, -lq.x));
        const float p1 = ex2_approx(fmaf(__uint_as_float(vs[e + 1]), p.scale_log2, -lq.y));
        const float p2 = ex2_approx(fmaf(__uint_as_float(vs[e + 2]), p.scale_log2, -lq.z));
        const float p3 = ex2_approx(fmaf(__uint_as_float(vs[e + 3]), p.scale_log2, -lq.w));
        if (DKV) {
          pp[e >> 1] = b2_pack(p0, p1);
          pp[(e >> 1) + 1] = b2_pack(p2, p3);
        }
        pd[e >> 1] = b2_pack(p0 * (__uint_as_float(vd[e]) - dq.x), p1 * (__uint_as_float(vd[e + 1]) - dq.y));
        pd[(e >> 1) + 1] = b2_pack(p2 * (__uint_as_float(vd[e + 2]) - dq.z), p3 * (__uint_as_float(vd[e + 3]) - dq.w));
      }
      // in place: this thread's 32 S (dP) values are in registers; their first 16 columns receive its P (dS) pairs
      if (DKV) tmem_st16_async(t_s, pp);
      tmem_st16_async(t_dp, pd);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&pds_ready[j]);
    }
    // ---- epilogue: warp (part = wid / 4) writes channels [16 part, 16 part + 16) of its row
    mbar_wait(&acc_final, 0);
    tc_fence_after();
    const int part = wid >> 2;
    bf16* base = dqkv + ((long long)n * p.T + row) * (3 * p.C);
    uint32_t v[16], w[16];
    tmem_ld16_async(tmem + lane_addr + 384u + (uint32_t)(16 * part), v);
    if (DKV) tmem_ld16_async(tmem + lane_addr + 448u + (uint32_t)(16 * part), w);
    tmem_wait_ld();
    if (row < p.T) {
      bf16* o0 = base + (DKV ? kc : qc) + 16 * part;
#pragma unroll
      for (int e = 0; e < 16; e += 8) {
        uint4 o;
        o.x = b2_pack(__uint_as_float(v[e]) * p.scale, __uint_as_float(v[e + 1]) * p.scale);
        o.y = b2_pack(__uint_as_float(v[e + 2]) * p.scale, __uint_as_float(v[e + 3]) * p.scale);
        o.z = b2_pack(__uint_as_float(v[e + 4]) * p.scale, __uint_as_float(v[e + 5]) * p.scale);
        o.w = b2_pack(__uint_as_float(v[e + 6]) * p.scale, __uint_as_float(v[e + 7]) * p.scale);
        *reinterpret_cast<uint4*>(o0 + e) = o;
        if (DKV) {
          uint4 u;
          u.x = b2_pack(__uint_as_float(w[e]), __uint_as_float(w[e + 1]));
          u.y = b2_pack(__uint_as_float(w[e + 2]), __uint_as_float(w[e + 3]));
          u.z = b2_pack(__uint_as_float(w[e + 4]), __uint_as_float(w[e + 5]));
          u.w = b2_pack(__uint_as_float(w[e + 6]), __uint_as_float(w[e + 7]));
          *reinterpret_cast<uint4*>(base + vc + 16 * part + e) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u));
}

// ---- host ---------------------------------------------------------------------------------------------
int attn_bwd_stats(const rho_attn_bwd_params* p, float* ld, int tiles, cudaStream_t st);      // attention_bwd_tc.cu

bool attention_bwd_tc2_supported(const rho_attn_bwd_params* p) {
  static int on = -1, min_t = 0;
  if (on < 0) {
    const char* e = getenv("RHO_ATTN_BWD_V2");
    on = (e && e[0] == '0') ? 0 : 1;
    const char* t = getenv("RHO_ATTN_BWD_V2_MIN_T");
    min_t = t ? atoi(t) : 512;
  }
  const rho_attn_params& f = p->fwd;
  return on && f.dtype == RHO_BF16 && f.ch == 64 && f.T >= min_t && f.N <= 65535 && f.heads <= 65535 && get_encode() != nullptr;
}

size_t attention_bwd_tc2_workspace(const rho_attn_bwd_params* p) {
  const rho_attn_params& f = p->fwd;
  return (size_t)f.N * f.heads * (2 * ((f.T + 63) / 64)) * 64 * sizeof(float);
}

static int b2_make_map(CUtensorMap* map, const void* base, int channels, int T, int N, int box_rows) {
  cuuint64_t gdim[3] = {(cuuint64_t)channels, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t gstr[2] = {(cuuint64_t)channels * 2, (cuuint64_t)channels * 2 * (cuuint64_t)T};
  cuuint32_t bdim[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, bdim, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "attention_bwd_tc2: cuTensorMapEncodeTiled failed: %d", (int)r);
  return RHO_OK;
}

int attention_bwd_tc2(const rho_attn_bwd_params* p, cudaStream_t st) {
  const rho_attn_params& f = p->fwd;
  float* ld = (float*)p->workspace;
  const int tiles = 2 * ((f.T + 63) / 64);
  int rc = attn_bwd_stats(p, ld, tiles, st);
  if (rc) return rc;
  AttnBwd2Params ap;
  ap.T = f.T; ap.heads = f.heads; ap.C = f.heads * f.ch;
  if (f.order == 0) { ap.q_col = 0; ap.k_col = f.ch; ap.v_col = 2 * f.ch; ap.head_stride = 3 * f.ch; }
  else { ap.q_col = 0; ap.k_col = ap.C; ap.v_col = 2 * ap.C; ap.head_stride = f.ch; }
  ap.scale = 1.0f / sqrtf((float)f.ch);
  ap.scale_log2 = ap.scale * 1.4426950408889634f;
  static int lean = -1;
  if (lean < 0) { const char* e = getenv("RHO_ATTN_BWD_LEAN"); lean = (e && e[0] == '0') ? 0 : 1; }
  ap.lean = lean;
  constexpr size_t SMEM = (size_t)(2 * B2_ROW_TILE + 2 * B2_ST * B2_COL_TILE) + 1024;   // 129 KB
  auto kq = attn_bwd2_kernel<false>;
  auto kkv = attn_bwd2_kernel<true>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "attention_bwd_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  CUtensorMap qkv128, qkv64, do128, do64;
  rc = b2_make_map(&qkv128, f.qkv, 3 * ap.C, f.T, f.N, B2_ROWS);
  if (!rc) rc = b2_make_map(&qkv64, f.qkv, 3 * ap.C, f.T, f.N, B2_COLS);
  if (!rc) rc = b2_make_map(&do128, p->dout, ap.C, f.T, f.N, B2_ROWS);
  if (!rc) rc = b2_make_map(&do64, p->dout, ap.C, f.T, f.N, B2_COLS);
  if (rc) return rc;
  dim3 grid((unsigned)ceil_div(f.T, B2_ROWS), f.heads, f.N);
  kq<<<grid, 576, SMEM, st>>>(qkv128, qkv64, do128, do64, ap, ld, (bf16*)p->dqkv);
  RHO_CHECK_LAUNCH("attention_bwd_dq_tcgen05_v2");
  kkv<<<grid, 576, SMEM, st>>>(qkv128, qkv64, do128, do64, ap, ld, (bf16*)p->dqkv);
  RHO_CHECK_LAUNCH("attention_bwd_dkv_tcgen05_v2");
  return RHO_OK;
}

}  // namespace rho
