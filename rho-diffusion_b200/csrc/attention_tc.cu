// QKVAttention / QKVAttentionLegacy (rho_diffusion/models/unet_v2.py:365-436) on the 5th-generation tensor cores:
// flash-style forward with tcgen05 MMAs, TMEM accumulators and TMA-fed operands (sm_100a).
//
// One CTA = 128 query rows of one (sample, head); key tiles of 64.
//   S_j = Q K_j^T  : tcgen05.mma 128x64x(ch)  -> TMEM columns [64*(j&1), +64)  -- S is DOUBLE-BUFFERED: the issuer launches
//                    S_{j+1} before it waits for P_j, so the tensor core computes the next logits while the row owners
//                    exponentiate the current ones (no MMA latency on the softmax critical path)
//   softmax        : thread r owns TMEM lane r = query row r, so row max / row sum need NO shuffles.  One sweep per tile:
//                    p = exp2(s*c - m*c) against the running max known BEFORE the tile (lazy max: O is rescaled and the
//                    tile redone only when a row outgrows it by 2^8; the stale max cancels in O / l and in the lse);
//                    P goes to shared memory as bf16 in the 128-byte-swizzled K-major layout (double-buffered)
//   O += P_j V_j   : tcgen05.mma 128x(ch)x64  -> TMEM columns [128,128+ch)  (A = P, B = V as an MN-major operand:
//                    V is [keys x ch] row-major, exactly what channels-last qkv gives)
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = softmax / correction /
// epilogue.  ch = 64: two CTAs per SM (81 KB smem, 256 TMEM columns each).  The T x T weight matrix never exists in HBM
// (the reference materialises 4.3 GB per sample per block at T = 16384).
#include <type_traits>

#include "tc_common.cuh"

namespace rho {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 64;    // keys per tile

struct AttnTcParams {
  int T, heads, ch, order;
  int q_col, k_col, v_col;   // channel offsets of head 0's q / k / v slices (per-head stride below)
  int head_stride;           // channels between consecutive heads' slices
  int C;                     // heads * ch
  float scale_log2;          // ch^-1/2 * log2(e)
  float scale;               // ch^-1/2
};

template <int CH, int POLY_MASK, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 320 : 192, CH == 64 ? 2 : 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap q_map, const __grid_constant__ CUtensorMap kv_map, const AttnTcParams p,
                   bf16* __restrict__ out, float* __restrict__ lse) {
  constexpr int KB = CH / 64;                   // 64-channel K blocks of Q / K, N atoms of V
  constexpr int Q_TILE = AT_BM * 128;           // [128 rows x 64 ch] bf16
  constexpr int KV_TILE = AT_BN * 128;          // [64 rows x 64 ch] bf16
  constexpr int Q_BYTES = KB * Q_TILE, KV_BYTES = KB * KV_TILE;
  constexpr int P_TILE = AT_BM * 128;           // [128 rows x 64 keys] bf16
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;                   // [2 stages]
  uint8_t* sV = sK + 2 * KV_BYTES;              // [2 stages]
  uint8_t* sP = sV + 2 * KV_BYTES;              // [2 buffers]
  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_ready[2], p_ready[2], o_done, o_final;
  __shared__ uint32_t tmem_base_s;
  __shared__ float sL[2][AT_BM];                // SPLIT: the two column halves' partial row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM, h = blockIdx.y, n = blockIdx.z;
  const int nt = (p.T + AT_BN - 1) / AT_BN;
  const int qc = p.q_col + h * p.head_stride, kc = p.k_col + h * p.head_stride, vc = p.v_col + h * p.head_stride;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_ready[s], 1);
      mbar_init(&p_ready[s], SPLIT ? 256 : 128);
    }
    mbar_init(&o_done, 1);
    mbar_init(&o_final, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_base_s;          // S: columns [0, 64) and [64, 128)
  const uint32_t tmem_o = tmem_base_s + 128;    // O: columns [128, 128 + CH)

  // Producer and issuer warps keep warp-uniform control flow (all lanes wait on the barriers); only the asynchronous
  // instructions are issued by one elected lane.
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(&q_full, Q_BYTES);
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(sQ + kb * Q_TILE, &q_map, &q_full, qc + kb * 64, q0, n);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1;
      mbar_wait(&k_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[s], KV_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(sK + s * KV_BYTES + kb * KV_TILE, &kv_map, &k_full[s], kc + kb * 64, j * AT_BN, n);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&v_full[s], KV_BYTES);
        for (int kb = 0; kb < KB; ++kb) tma_load_3d(sV + s * KV_BYTES + kb * KV_TILE, &kv_map, &v_full[s], vc + kb * 64, j * AT_BN, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // instruction descriptors: D fp32, A/B bf16, M = 128
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_BN >> 3) << 17) | ((uint32_t)(AT_BM >> 4) << 24);
    const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B is MN-major */ | ((uint32_t)(CH >> 3) << 17) |
                              ((uint32_t)(AT_BM >> 4) << 24);
    auto issue_qk = [&](int j) {
      const int s = j & 1;
      mbar_wait(&k_full[s], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t a = at_desc_kmajor<128>(smem_u32(sQ + kb * Q_TILE));
          const uint64_t b = at_desc_kmajor<128>(smem_u32(sK + s * KV_BYTES + kb * KV_TILE));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_s + (uint32_t)(s * AT_BN), a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), idesc_qk, (kb | k) != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_ready[s]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    issue_qk(0);
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1;
      // S buffer (j+1)&1 was last read for tile j-1, whose p_ready was awaited in the previous iteration
      if (j + 1 < nt) issue_qk(j + 1);
      mbar_wait(&p_ready[s], ph);               // P_j is in smem and the row owners are done reading S_j
      mbar_wait(&v_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a = at_desc_kmajor<128>(smem_u32(sP + s * P_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) {             // 16 keys per MMA: two 8-row K atoms = 2048 bytes of V
          const uint64_t b = at_desc_mnmajor(smem_u32(sV + s * KV_BYTES + k * 16 * 128), KV_TILE);
          umma_f16(tmem_o, a + (uint64_t)(2 * k), b, idesc_pv, (j | k) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(&o_done);
        if (j == nt - 1) umma_commit(&o_final);
      }
      __syncwarp();
    }
  } else if constexpr (SPLIT) {
    // ===== softmax / correction / epilogue, TWO threads per query row (warps 2-5: S columns [0, 32), warps 6-9: [32, 64)):
    // twice the warps keep the MUFU busy while partner warps sit in TMEM loads, fences and barrier waits.  Both threads of
    // a row reduce the FULL row maximum (the other half of S is loaded for that alone), so they take identical rescale
    // decisions without talking to each other; the row sum is kept in two halves and joined once, in the epilogue.
    const int half = (warp - 2) >> 2;
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    constexpr float AT_TAU = 8.0f;
    constexpr int OC = CH / 2;                    // O columns owned by this thread: [half * OC, +OC)
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      mbar_wait(&s_ready[s], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_s + (uint32_t)(s * AT_BN);
      uint8_t* pbuf = sP + s * P_TILE;
      const int kvalid = p.T - j * AT_BN;
      const bool full = kvalid >= AT_BN;
      uint32_t v[32], u[32];
      tmem_ld32_async(ts + lane_addr + (uint32_t)(half * 32), v);
      tmem_ld32_async(ts + lane_addr + (uint32_t)((half ^ 1) * 32), u);
      tmem_wait_ld();
      const int c = half * 32, co = (half ^ 1) * 32;
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (full) {
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[e]), __uint_as_float(u[e])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[e + 1]), __uint_as_float(u[e + 1])));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          mx0 = fmaxf(mx0, c + e < kvalid ? __uint_as_float(v[e]) : -INFINITY);
          mx1 = fmaxf(mx1, co + e < kvalid ? __uint_as_float(u[e]) : -INFINITY);
        }
      }
      const float mx = fmaxf(mx0, mx1);
      if (j == 0) {
        m = mx;
      } else if (__any_sync(0xffffffffu, (mx - m) * p.scale_log2 > AT_TAU)) {
        // a row of this warp outgrew the stale max by more than 2^AT_TAU: rescale this thread's half of the O row once
        // every PV issued so far (tiles < j) has completed (the partner warp takes the same branch for the other half)
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2_approx((m - m_new) * p.scale_log2);
        mbar_wait(&o_done, (uint32_t)(j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < OC; cc += 32) {
          uint32_t w[32];
          tmem_ld32(tmem_o + lane_addr + (uint32_t)(half * OC + cc), w);
#pragma unroll
          for (int e = 0; e < 32; ++e) w[e] = __float_as_uint(__uint_as_float(w[e]) * alpha);
          tmem_st32(tmem_o + lane_addr + (uint32_t)(half * OC + cc), w);
        }
        l *= alpha;
        m = m_new;
      }
      const float mb = m * p.scale_log2;
      const uint64_t negmb2 = pack2(-mb, -mb);
      uint64_t ls01 = 0ull, ls23 = 0ull;
      uint32_t packed[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const uint64_t x2 = fma2(pack2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), scale2, negmb2);
        float x0, x1;
        unpack2(x2, x0, x1);
        float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        if (!full) {
          p0 = c + e < kvalid ? p0 : 0.f;
          p1 = c + e + 1 < kvalid ? p1 : 0.f;
        }
        const uint64_t p2 = pack2(p0, p1);
        if (e & 2) ls23 = add2(ls23, p2); else ls01 = add2(ls01, p2);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
        packed[e >> 1] = *reinterpret_cast<uint32_t*>(&b2);
      }
      uint8_t* rowp = pbuf + r * 128;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int chunk = (half * 4 + i) ^ (r & 7);
        *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
      }
      float a0, a1, a2, a3;
      unpack2(ls01, a0, a1);
      unpack2(ls23, a2, a3);
      l += (a0 + a1) + (a2 + a3);
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[s]);
    }
    // ---- epilogue: join the two halves of the row sum, then O / l -> bf16 -> global (each thread its OC columns)
    sL[half][r] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float ltot = sL[0][r] + sL[1][r];
    mbar_wait(&o_final, 0);
    tc_fence_after();
    const int q = q0 + r;
    const float inv = 1.0f / ltot;
    bf16* orow = out + ((long long)n * p.T + q) * p.C + h * CH + half * OC;
#pragma unroll 1
    for (int cc = 0; cc < OC; cc += 32) {
      uint32_t w[32];
      tmem_ld32(tmem_o + lane_addr + (uint32_t)(half * OC + cc), w);
      if (q < p.T) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          __nv_bfloat162 h2;
          h2 = __floats2bfloat162_rn(__uint_as_float(w[e]) * inv, __uint_as_float(w[e + 1]) * inv); o.x = *reinterpret_cast<uint32_t*>(&h2);
          h2 = __floats2bfloat162_rn(__uint_as_float(w[e + 2]) * inv, __uint_as_float(w[e + 3]) * inv); o.y = *reinterpret_cast<uint32_t*>(&h2);
          h2 = __floats2bfloat162_rn(__uint_as_float(w[e + 4]) * inv, __uint_as_float(w[e + 5]) * inv); o.z = *reinterpret_cast<uint32_t*>(&h2);
          h2 = __floats2bfloat162_rn(__uint_as_float(w[e + 6]) * inv, __uint_as_float(w[e + 7]) * inv); o.w = *reinterpret_cast<uint32_t*>(&h2);
          *reinterpret_cast<uint4*>(orow + cc + e) = o;
        }
      }
    }
    if (half == 0 && lse && q < p.T) lse[((long long)n * p.heads + h) * p.T + q] = m * p.scale + logf(ltot);
  } else {
    // ===== softmax / correction / epilogue: thread <-> query row =====
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    float m = -INFINITY, l = 0.f;
    // One sweep over the 64 logits of this row: p = exp2(s*c - mb) (one FFMA + one MUFU per element) -> bf16 P in swizzled
    // smem; returns the row sum and the row max of the raw logits.
    // Exponentials: the MUFU unit delivers 16 ex2 / clk / SM, half of what the tensor core can consume at ch = 64, so a
    // compile-time subset of the element PAIRS (POLY_MASK, period 8 pairs) is evaluated on the FMA pipes instead, two
    // elements per instruction (f32x2): x clamped to >= -126, n = round(x) by the magic-number add, f = x - n in
    // [-0.5, 0.5], 2^f by a degree-3 minimax polynomial (7.5e-5 relative, far below bf16 resolution), 2^n by an integer
    // add into the exponent field.
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);
    auto exp_pass = [&](auto full_c, uint32_t ts, uint8_t* pbuf, int kvalid, float mb, float& lsum, float& mx_out) {
      constexpr bool full = decltype(full_c)::value;
      uint32_t va[32], vb[32];
      tmem_ld32_async(ts + lane_addr, va);
      tmem_ld32_async(ts + lane_addr + 32u, vb);
      tmem_wait_ld();
      const uint64_t negmb2 = pack2(-mb, -mb);
      uint64_t ls01 = 0ull, ls23 = 0ull;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t(&v)[32] = half ? vb : va;
        const int c = half * 32;
        uint32_t packed[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float s0 = __uint_as_float(v[e]), s1 = __uint_as_float(v[e + 1]);
          const uint64_t x2 = fma2(pack2(s0, s1), scale2, negmb2);
          uint64_t p2;
          if ((POLY_MASK >> ((e >> 1) & 7)) & 1) {
            float x0, x1;
            unpack2(x2, x0, x1);
            const uint64_t xc = pack2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
            const uint64_t t2 = add2(xc, pack2(12582912.f, 12582912.f));
            const uint64_t n2 = add2(t2, pack2(-12582912.f, -12582912.f));
            const uint64_t f2 = fma2(n2, pack2(-1.f, -1.f), xc);
            uint64_t q2 = fma2(f2, pack2(0.0551716648042202f, 0.0551716648042202f), pack2(0.2426111251115799f, 0.2426111251115799f));
            q2 = fma2(q2, f2, pack2(0.6932609677314758f, 0.6932609677314758f));
            q2 = fma2(q2, f2, pack2(0.9999280571937561f, 0.9999280571937561f));
            uint32_t tl, th, ql, qh;
            unpack2u(t2, tl, th);
            unpack2u(q2, ql, qh);
            p2 = pack2u(ql + (tl << 23), qh + (th << 23));     // 2^n: add n into the exponent field
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            p2 = pack2(ex2_approx(x0), ex2_approx(x1));
          }
          if (full) {
            mx0 = fmaxf(mx0, s0);
            mx1 = fmaxf(mx1, s1);
          } else {
            const bool ok0 = c + e < kvalid, ok1 = c + e + 1 < kvalid;
            float p0, p1;
            unpack2(p2, p0, p1);
            p2 = pack2(ok0 ? p0 : 0.f, ok1 ? p1 : 0.f);
            mx0 = fmaxf(mx0, ok0 ? s0 : -INFINITY);
            mx1 = fmaxf(mx1, ok1 ? s1 : -INFINITY);
          }
          if (e & 2) ls23 = add2(ls23, p2); else ls01 = add2(ls01, p2);
          float p0, p1;
          unpack2(p2, p0, p1);
          __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
          packed[e >> 1] = *reinterpret_cast<uint32_t*>(&b2);
        }
        // 32 keys = 4 chunks of 16 bytes; chunk index within the 128-byte row = c / 8 + i
        uint8_t* rowp = pbuf + r * 128;
        const int ch0 = c >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int chunk = (ch0 + i) ^ (r & 7);
          *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
        }
      }
      float a0, a1, a2, a3;
      unpack2(ls01, a0, a1);
      unpack2(ls23, a2, a3);
      lsum = (a0 + a1) + (a2 + a3);
      mx_out = fmaxf(mx0, mx1);
    };
    constexpr float AT_TAU = 8.0f;
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      mbar_wait(&s_ready[s], (uint32_t)(j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_s + (uint32_t)(s * AT_BN);
      uint8_t* pbuf = sP + s * P_TILE;
      const int kvalid = p.T - j * AT_BN;         // keys >= kvalid are padding (only in the last tile)
      float lsum, mx;
      if (j == 0) {
        // first tile: the row max is not known yet -> one extra sweep for it
        uint32_t va[32], vb[32];
        tmem_ld32_async(ts + lane_addr, va);
        tmem_ld32_async(ts + lane_addr + 32u, vb);
        tmem_wait_ld();
        float mx0 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          mx0 = fmaxf(mx0, e < kvalid ? __uint_as_float(va[e]) : -INFINITY);
          mx0 = fmaxf(mx0, 32 + e < kvalid ? __uint_as_float(vb[e]) : -INFINITY);
        }
        m = mx0;
      }
      // one sweep against the running max known before the tile; redone (once) after a rescale of O when a row of this
      // warp outgrew the stale max by more than 2^AT_TAU.  Single call site per variant: the sweep is ~1000 instructions.
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        if (kvalid >= AT_BN) exp_pass(std::true_type{}, ts, pbuf, kvalid, m * p.scale_log2, lsum, mx);
        else exp_pass(std::false_type{}, ts, pbuf, kvalid, m * p.scale_log2, lsum, mx);
        if (attempt == 1 || j == 0 || !__any_sync(0xffffffffu, (mx - m) * p.scale_log2 > AT_TAU)) break;
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2_approx((m - m_new) * p.scale_log2);
        // rescale this row of O once every PV issued so far (tiles < j) has completed
        mbar_wait(&o_done, (uint32_t)(j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < CH; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_o + lane_addr + (uint32_t)c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st32(tmem_o + lane_addr + (uint32_t)c, v);
        }
        l *= alpha;
        m = m_new;
      }
      l += lsum;
      tc_fence_before();          // order our tcgen05.ld/st before the MMAs the issuer launches after the barrier
      fence_proxy_async_smem();   // make the P stores visible to the tensor core (async proxy)
      mbar_arrive(&p_ready[s]);
    }
    // ---- epilogue: O / l -> bf16 -> global.  A dedicated single-phase barrier: a parity wait on o_done would pass spuriously
    // if PV(nt-2) had not completed yet (the barrier two phases behind has the same parity as the awaited one).
    mbar_wait(&o_final, 0);
    tc_fence_after();
    const int q = q0 + r;
    const float inv = 1.0f / l;
    bf16* orow = out + ((long long)n * p.T + q) * p.C + h * CH;
#pragma unroll 1
    for (int c = 0; c < CH; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_o + lane_addr + (uint32_t)c, v);
      if (q < p.T) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          uint4 o;
          __nv_bfloat162 h2;
          h2 = __floats2bfloat162_rn(__uint_as_float(v[e]) * inv, __uint_as_float(v[e + 1]) * inv); o.x = *reinterpret_cast<uint32_t*>(&h2);
          h2 = __floats2bfloat162_rn(__uint_as_float(v[e + 2]) * inv, __uint_as_float(v[e + 3]) * inv); o.y = *reinterpret_cast<uint32_t*>(&h2);
          h2 = __floats2bfloat162_rn(__uint_as_float(v[e + 4]) * inv, __uint_as_float(v[e + 5]) * inv); o.z = *reinterpret_cast<uint32_t*>(&h2);
          h2 = __floats2bfloat162_rn(__uint_as_float(v[e + 6]) * inv, __uint_as_float(v[e + 7]) * inv); o.w = *reinterpret_cast<uint32_t*>(&h2);
          *reinterpret_cast<uint4*>(orow + c + e) = o;
        }
      }
    }
    if (lse && q < p.T) lse[((long long)n * p.heads + h) * p.T + q] = m * p.scale + logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(256u));
}

// ---- host ---------------------------------------------------------------------------------------------
bool attention_fwd_tc_supported(const rho_attn_params* p) {
  const char* e = getenv("RHO_DISABLE_TCGEN05");
  if (e && e[0] == '1') return false;
  if (p->dtype != RHO_BF16) return false;
  if (p->ch != 64 && p->ch != 128) return false;
  if (p->T < 1 || p->N > 65535 || p->heads > 65535) return false;
  return get_encode() != nullptr;
}

template <int CH, int POLY_MASK, bool SPLIT>
static int launch_attn_tc(const CUtensorMap& qmap, const CUtensorMap& kvmap, const AttnTcParams& ap, const rho_attn_params* p, cudaStream_t st) {
  constexpr int KB = CH / 64;
  constexpr size_t SMEM = (size_t)KB * 16384 + 4 * (size_t)KB * 8192 + 2 * 16384 + 1024;
  auto kern = attn_tc_kernel<CH, POLY_MASK, SPLIT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid((unsigned)ceil_div(p->T, AT_BM), p->heads, p->N);
  kern<<<grid, SPLIT ? 320 : 192, SMEM, st>>>(qmap, kvmap, ap, (bf16*)p->out, p->lse);
  RHO_CHECK_LAUNCH("attention_fwd_tcgen05");
  return RHO_OK;
}

// attention_tc2.cu: 128-key tiles, P in tensor memory, two query tiles per CTA (head dim 64, long sequences)
bool attention_fwd_tc2_supported(const rho_attn_params* p);
int attention_fwd_tc2(const rho_attn_params* p, cudaStream_t st);

int attention_fwd_tc(const rho_attn_params* p, cudaStream_t st) {
  if (attention_fwd_tc2_supported(p)) return attention_fwd_tc2(p, st);
  auto enc = get_encode();
  if (!enc) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  const int C = p->heads * p->ch;
  CUtensorMap map, kvmap;
  cuuint64_t gdim[3] = {(cuuint64_t)(3 * C), (cuuint64_t)p->T, (cuuint64_t)p->N};
  cuuint64_t gstr[2] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)(3 * C) * 2 * (cuuint64_t)p->T};
  cuuint32_t bdim[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p->qkv), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed: %d", (int)r);
  bdim[1] = AT_BN;
  r = enc(&kvmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p->qkv), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(kv) failed: %d", (int)r);
  AttnTcParams ap;
  ap.T = p->T; ap.heads = p->heads; ap.ch = p->ch; ap.order = p->order; ap.C = C;
  if (p->order == 0) { ap.q_col = 0; ap.k_col = p->ch; ap.v_col = 2 * p->ch; ap.head_stride = 3 * p->ch; }
  else { ap.q_col = 0; ap.k_col = C; ap.v_col = 2 * C; ap.head_stride = p->ch; }
  ap.scale = 1.0f / sqrtf((float)p->ch);
  ap.scale_log2 = ap.scale * 1.4426950408889634f;
  // pairs (of every 8) whose exponentials run on the FMA pipes: RHO_ATTN_POLY = 0 | 2 | 3 | 4.  Measured on B200 (T = 16384,
  // ch = 64, N = 8): 742 / 735 / 739 / 696 TFLOP/s -- the sweep is issue-bound as soon as the MUFU is relieved, so the default
  // stays 0 (all exponentials on the MUFU); the hook is kept for re-tuning.
  static int poly = -1;
  if (poly < 0) { const char* e = getenv("RHO_ATTN_POLY"); poly = e ? atoi(e) : 0; }
  // two threads per query row (320-thread CTAs): RHO_ATTN_SPLIT = 1 | 0.  Measured on B200 (T = 16384, ch = 64, N = 8): 687 vs
  // 724 TFLOP/s for the one-thread-per-row kernel, 153.6 vs 156.9 evals/s end to end -- the sweep is MUFU-bound, more warps
  // only add the second row-max reduction and TMEM reads.  Default off; kept as a tuning hook.
  static int split = -1;
  if (split < 0) { const char* e = getenv("RHO_ATTN_SPLIT"); split = e ? atoi(e) : 0; }
  if (split && poly == 0) {
    if (p->ch == 64) return launch_attn_tc<64, 0x00, true>(map, kvmap, ap, p, st);
    return launch_attn_tc<128, 0x00, true>(map, kvmap, ap, p, st);
  }
  if (p->ch == 64) {
    if (poly == 0) return launch_attn_tc<64, 0x00, false>(map, kvmap, ap, p, st);
    if (poly == 2) return launch_attn_tc<64, 0x22, false>(map, kvmap, ap, p, st);
    if (poly == 4) return launch_attn_tc<64, 0xAA, false>(map, kvmap, ap, p, st);
    if (poly == 3) return launch_attn_tc<64, 0x4A, false>(map, kvmap, ap, p, st);
    return launch_attn_tc<64, 0x00, false>(map, kvmap, ap, p, st);
  }
  if (poly != 0) return launch_attn_tc<128, 0x22, false>(map, kvmap, ap, p, st);
  return launch_attn_tc<128, 0x00, false>(map, kvmap, ap, p, st);
}

}  // namespace rho
