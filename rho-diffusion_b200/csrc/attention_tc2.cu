// QKVAttention / QKVAttentionLegacy forward (rho_diffusion/models/unet_v2.py:365-436), second-generation tcgen05 kernel for
// head dim 64: 128-key tiles, P kept in TENSOR MEMORY, two query tiles per CTA in ping-pong.
//
// Why (round-1 profile, profiles/r01h_ncu_full_attn_tc_kernel_summary.json): with 64-key tiles both MMAs of a tile are N = 64
// shared-memory-operand MMAs (57 % of the tensor peak at best, tools/ubench/umma_rate.cu), P went registers -> swizzled shared
// memory -> async-proxy fence -> MMA, and one softmax warp-group per CTA left the MUFU idle 34 % of the time.
//
// One CTA = 256 query rows of one (sample, head) = two 128-row tiles, one CTA per SM (all 512 TMEM columns):
//   S_i = Q_i K_j^T   tcgen05.mma 128 x 128 x 64 (SS)      -> TMEM columns [128 i, +128)        i = 0, 1
//   P_i = exp2(...)   softmax warp-group i (thread <-> query row, no shuffles), bf16 pairs stored with tcgen05.st
//                                                           -> TMEM columns [384 + 64 i, +64)
//   O_i += P_i V_j    tcgen05.mma 128 x 64 x 128, A operand read FROM TMEM (TS form), B = V as an MN-major smem operand
//                                                           -> TMEM columns [256 + 64 i, +64)
// While warp-group 0 exponentiates S_0 the tensor core runs QK_1 / PV_1 and vice versa, so the MUFU (16 ex2 / clk / SM: the
// bound at head dim 64) never waits for an MMA and the N = 64 PV MMA fetches only 2 KB of V per K step from shared memory.
// The issuer launches QK_i(j+1) BEFORE PV_i(j) (both only need P_i(j) complete), so the next logits arrive 256 clk earlier
// (PV first: 680 vs 759 TFLOP/s); the row owners therefore wait for PV_i(j) (o_done) before their first P store of tile j+1.
// Measured variants that did not pay (tools/attn_trace.py timelines, profiles/r02_attn_*): S produced / consumed in 64-key halves
// with their own barriers (N = 64 QK MMAs cost 45 clk each instead of 64 per two: 720 TFLOP/s); two threads per query row
// (16 softmax warps, one 64-thread barrier.red.or per tile for the redo vote: 802 TFLOP/s).
// Softmax bookkeeping: p = exp2(s c - m c) against the running max known BEFORE the tile; no per-element max is taken on
// the common path -- an outgrown max shows up as a row sum > 2^16 (some p > 2^9), in which case the warp takes the exact
// max, rescales its O rows in TMEM and redoes the tile (P is not aliased onto S, so S is still there).  The stale max
// cancels in O / l and in the lse.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = softmax / correction /
// epilogue of query tile 0, warps 6-9 = of query tile 1 (a warp touches TMEM lanes [32 (warp % 4), +32)).
#include <type_traits>

#include "tc_common.cuh"

namespace rho {

constexpr int A2_BM = 128;            // rows per query tile (two tiles per CTA)
constexpr int A2_BN = 128;            // keys per tile
constexpr int A2_ST = 3;              // K / V ring stages
constexpr int A2_TILE = 128 * 128;    // bytes of a [128 rows x 64 channels] bf16 tile (128-byte swizzled rows)

struct AttnTc2Params {
  int T, heads, C;
  int q_col, k_col, v_col, head_stride;
  float scale_log2, scale;
  int ld_split;               // tuning hook: start the exponentials after the first 32-column TMEM load
  long long* trace;           // debug: clock64 stamps of CTA (0,0,0)'s hand-offs, [tile][8] (tools/attn_trace.py), or null
};
__device__ __forceinline__ void a2_stamp(const AttnTc2Params& p, int j, int slot) {
  if (p.trace && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && j < 64 && (threadIdx.x & 31) == 0) p.trace[j * 16 + slot] = clock64();
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// 32 logits of one row -> 32 probabilities (bf16 pairs) + partial row sums.  POLY_MASK selects the element pairs (period 8)
// whose exponential runs on the FMA pipes (degree-3 minimax on [-0.5, 0.5], 7.5e-5 relative) instead of the MUFU.
template <bool FULL, int POLY_MASK>
__device__ __forceinline__ void a2_exp32(const uint32_t (&v)[32], int c0, int kvalid, uint64_t scale2, uint64_t negmb2, uint64_t& ls0,
                                         uint64_t& ls1, uint32_t (&packed)[16]) {
#pragma unroll
  for (int e = 0; e < 32; e += 2) {
    const uint64_t x2 = fma2(pack2(__uint_as_float(v[e]), __uint_as_float(v[e + 1])), scale2, negmb2);
    float p0, p1;
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
      p0 = __uint_as_float(ql + (tl << 23));      // 2^n: add n into the exponent field
      p1 = __uint_as_float(qh + (th << 23));
    } else {
      float x0, x1;
      unpack2(x2, x0, x1);
      p0 = ex2_approx(x0);
      p1 = ex2_approx(x1);
    }
    if (!FULL) {
      p0 = c0 + e < kvalid ? p0 : 0.f;
      p1 = c0 + e + 1 < kvalid ? p1 : 0.f;
    }
    if (e & 2) ls1 = add2(ls1, pack2(p0, p1)); else ls0 = add2(ls0, pack2(p0, p1));
    __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);          // low half = the lower key index
    packed[e >> 1] = *reinterpret_cast<uint32_t*>(&b2);
  }
}

// exact max of the (valid) raw logits of this thread's row
__device__ __forceinline__ float a2_rowmax(uint32_t ts, int kvalid) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < A2_BN; c += 32) {
    uint32_t v[32];
    tmem_ld32(ts + (uint32_t)c, v);
    if (kvalid >= c + 32) {
#pragma unroll
      for (int e = 0; e < 32; e += 2) mx = fmax3(mx, __uint_as_float(v[e]), __uint_as_float(v[e + 1]));
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) mx = fmaxf(mx, c + e < kvalid ? __uint_as_float(v[e]) : -INFINITY);
    }
  }
  return mx;
}

// One sweep over the 128 logits of this row: P -> TMEM, returns the row sum.  `wait_bar` (PV of the previous tile, which
// still reads the P buffer) is awaited right before the first P store.
template <bool FULL, int POLY_MASK>
__device__ __forceinline__ float a2_exp_pass(uint32_t ts, uint32_t tp, int kvalid, float mb, float scale_log2, uint64_t* wait_bar,
                                             uint32_t wait_parity, int ld_split) {
  const uint64_t scale2 = pack2(scale_log2, scale_log2);
  const uint64_t negmb2 = pack2(-mb, -mb);
  uint64_t ls0 = 0ull, ls1 = 0ull;
  uint32_t a[32], b[32], packed[16];
  tmem_ld32_async(ts, a);
  if (ld_split) tmem_wait_ld();
  tmem_ld32_async(ts + 32u, b);
  if (!ld_split) tmem_wait_ld();
  a2_exp32<FULL, POLY_MASK>(a, 0, kvalid, scale2, negmb2, ls0, ls1, packed);
  tmem_wait_ld();
  tmem_ld32_async(ts + 64u, a);
  if (wait_bar) {
    mbar_wait(wait_bar, wait_parity);
    tc_fence_after();
  }
  tmem_st16_async(tp, packed);
  a2_exp32<FULL, POLY_MASK>(b, 32, kvalid, scale2, negmb2, ls0, ls1, packed);
  tmem_wait_ld();
  tmem_ld32_async(ts + 96u, b);
  tmem_st16_async(tp + 16u, packed);
  a2_exp32<FULL, POLY_MASK>(a, 64, kvalid, scale2, negmb2, ls0, ls1, packed);
  tmem_wait_ld();
  tmem_st16_async(tp + 32u, packed);
  a2_exp32<FULL, POLY_MASK>(b, 96, kvalid, scale2, negmb2, ls0, ls1, packed);
  tmem_st16_async(tp + 48u, packed);
  tmem_wait_st();
  float s0, s1, s2, s3;
  unpack2(ls0, s0, s1);
  unpack2(ls1, s2, s3);
  return (s0 + s1) + (s2 + s3);
}

template <int POLY_MASK>
__global__ void __launch_bounds__(320, 1)
    attn_tc2_kernel(const __grid_constant__ CUtensorMap qkv_map, const AttnTc2Params p, bf16* __restrict__ out, float* __restrict__ lse) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                          // [2 query tiles]
  uint8_t* sK = sQ + 2 * A2_TILE;              // [A2_ST stages]
  uint8_t* sV = sK + A2_ST * A2_TILE;          // [A2_ST stages]
  __shared__ uint64_t q_full, k_full[A2_ST], k_empty[A2_ST], v_full[A2_ST], v_empty[A2_ST];
  __shared__ uint64_t s_ready[2], p_ready[2], o_done[2], o_final[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * A2_BM), h = blockIdx.y, n = blockIdx.z;
  const int nt = (p.T + A2_BN - 1) / A2_BN;
  const int qc = p.q_col + h * p.head_stride, kc = p.k_col + h * p.head_stride, vc = p.v_col + h * p.head_stride;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < A2_ST; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], A2_BM);
      mbar_init(&o_done[i], 1);
      mbar_init(&o_final[i], 1);
    }
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
    // ===== TMA producer (warp-uniform control flow; one elected lane issues) =====
    if (elect_one()) {
      mbar_expect_tx(&q_full, 2 * A2_TILE);
      tma_load_3d(sQ, &qkv_map, &q_full, qc, q0, n);
      tma_load_3d(sQ + A2_TILE, &qkv_map, &q_full, qc, q0 + A2_BM, n);
    }
    __syncwarp();
    for (int j = 0; j < nt; ++j) {
      const int s = j % A2_ST;
      const uint32_t ph = (uint32_t)(j / A2_ST) & 1;
      mbar_wait(&k_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&k_full[s], A2_TILE);
        tma_load_3d(sK + s * A2_TILE, &qkv_map, &k_full[s], kc, j * A2_BN, n);
      }
      __syncwarp();
      mbar_wait(&v_empty[s], ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&v_full[s], A2_TILE);
        tma_load_3d(sV + s * A2_TILE, &qkv_map, &v_full[s], vc, j * A2_BN, n);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptors: D fp32, A / B bf16, M = 128
    const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(A2_BN >> 3) << 17) | ((uint32_t)(A2_BM >> 4) << 24);
    const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /* B is MN-major */ | ((uint32_t)(64 >> 3) << 17) |
                              ((uint32_t)(A2_BM >> 4) << 24);
    auto issue_qk = [&](int i, int j) {        // S_i = Q_i K_j^T  (K_j's stage must be full)
      if (elect_one()) {
        const uint64_t a = at_desc_kmajor<128>(smem_u32(sQ + i * A2_TILE));
        const uint64_t b = at_desc_kmajor<128>(smem_u32(sK + (j % A2_ST) * A2_TILE));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(128 * i), a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), idesc_qk, k != 0);
        umma_commit(&s_ready[i]);
        if (i == 1) umma_commit(&k_empty[j % A2_ST]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_qk(0, 0);
    issue_qk(1, 0);
    // The issuer's own instruction latency is on the row owners' critical loop (p_ready -> QK -> s_ready): everything that does
    // not depend on P (waiting for the K / V stages, which the producer fills tiles ahead) happens BEFORE the p_ready wait, and
    // the MMAs of a hand-off are issued straight-line from one elected region.  Measured dead ends (tools/attn_trace.py): a
    // rolled K loop (240 clk per MMA of scalar latency), polling the other tile's p_ready between PV groups so that its QK
    // cuts in line (648 vs 737 TFLOP/s: three elected regions + polls cost more than the queueing they remove).
    for (int j = 0; j < nt; ++j) {
      const int sj = j % A2_ST;
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        if (i == 0) {
          if (j + 1 < nt) mbar_wait(&k_full[(j + 1) % A2_ST], (uint32_t)((j + 1) / A2_ST) & 1);
          mbar_wait(&v_full[sj], (uint32_t)(j / A2_ST) & 1);
        }
        mbar_wait(&p_ready[i], (uint32_t)j & 1);      // P_i(j) is in TMEM and the row owners are done reading S_i(j)
        tc_fence_after();
        a2_stamp(p, j, 4 + 2 * i);
        if (elect_one()) {
          if (j + 1 < nt) {                           // S_i = Q_i K_{j+1}^T first: the row owners wait for it
            const uint64_t a = at_desc_kmajor<128>(smem_u32(sQ + i * A2_TILE));
            const uint64_t b = at_desc_kmajor<128>(smem_u32(sK + ((j + 1) % A2_ST) * A2_TILE));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(tmem + (uint32_t)(128 * i), a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), idesc_qk, k != 0);
            umma_commit(&s_ready[i]);
            if (i == 1) umma_commit(&k_empty[(j + 1) % A2_ST]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {               // 16 keys per MMA: 8 TMEM columns of P, two 8-row K atoms (2048 B) of V
            const uint64_t b = at_desc_mnmajor(smem_u32(sV + sj * A2_TILE + k * 16 * 128), A2_TILE);
            umma_f16_ts(tmem + 256u + (uint32_t)(64 * i), tmem + 384u + (uint32_t)(64 * i + 8 * k), b, idesc_pv, (j | k) != 0);
          }
          umma_commit(&o_done[i]);
          if (i == 1) umma_commit(&v_empty[sj]);
          if (j == nt - 1) umma_commit(&o_final[i]);
        }
        __syncwarp();
        a2_stamp(p, j, 5 + 2 * i);
      }
    }
  } else {
    // ===== softmax / correction / epilogue: warp-group i <-> query tile i, thread <-> query row =====
    const int i = (warp - 2) >> 2;
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const uint32_t ts = tmem + lane_addr + (uint32_t)(128 * i);
    const uint32_t to = tmem + lane_addr + 256u + (uint32_t)(64 * i);
    const uint32_t tp = tmem + lane_addr + 384u + (uint32_t)(64 * i);
    constexpr float A2_SUM_LIMIT = 65536.f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < nt; ++j) {
      mbar_wait(&s_ready[i], (uint32_t)j & 1);
      tc_fence_after();
      if ((warp & 3) == 0) a2_stamp(p, j, 2 * i);
      const int kvalid = p.T - j * A2_BN;           // keys >= kvalid are padding (only in the last tile)
      if (j == 0) m = a2_rowmax(ts, kvalid);
      float lsum;
#pragma unroll 1
      for (int attempt = 0; attempt < 2; ++attempt) {
        uint64_t* wb = (j > 0 && attempt == 0) ? &o_done[i] : nullptr;
        if (kvalid >= A2_BN) lsum = a2_exp_pass<true, POLY_MASK>(ts, tp, kvalid, m * p.scale_log2, p.scale_log2, wb, (uint32_t)(j - 1) & 1, p.ld_split);
        else lsum = a2_exp_pass<false, POLY_MASK>(ts, tp, kvalid, m * p.scale_log2, p.scale_log2, wb, (uint32_t)(j - 1) & 1, p.ld_split);
        if (attempt == 1 || j == 0 || !__any_sync(0xffffffffu, !(lsum <= A2_SUM_LIMIT))) break;
        // a row of this warp outgrew the stale max: exact max, rescale this row of O (PV of tile j-1 has completed: o_done was
        // awaited above), redo the tile
        const float m_new = fmaxf(m, a2_rowmax(ts, kvalid));
        const float alpha = ex2_approx((m - m_new) * p.scale_log2);
#pragma unroll 1
        for (int c = 0; c < 64; c += 32) {
          uint32_t v[32];
          tmem_ld32(to + (uint32_t)c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
          tmem_st32(to + (uint32_t)c, v);
        }
        l *= alpha;
        m = m_new;
      }
      l += lsum;
      tc_fence_before();          // order our tcgen05.ld / st before the MMAs the issuer launches after the barrier
      mbar_arrive(&p_ready[i]);
      if ((warp & 3) == 0) a2_stamp(p, j, 2 * i + 1);
    }
    // ---- epilogue: O / l -> bf16 -> global
    mbar_wait(&o_final[i], 0);
    tc_fence_after();
    const int q = q0 + i * A2_BM + r;
    const float inv = 1.0f / l;
    bf16* orow = out + ((long long)n * p.T + q) * p.C + h * 64;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld32(to + (uint32_t)c, v);
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
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512u));
}


// ---- host ---------------------------------------------------------------------------------------------
static long long* g_attn_trace = nullptr;
void attention_tc2_set_trace(long long* buf) { g_attn_trace = buf; }

bool attention_fwd_tc2_supported(const rho_attn_params* p) {
  static int on = -1, min_t = 0;
  if (on < 0) {
    const char* e = getenv("RHO_ATTN_V2");
    on = (e && e[0] == '0') ? 0 : 1;
    const char* t = getenv("RHO_ATTN_V2_MIN_T");
    min_t = t ? atoi(t) : 512;
  }
  return on && p->dtype == RHO_BF16 && p->ch == 64 && p->T >= min_t && p->N <= 65535 && p->heads <= 65535;
}

template <int POLY_MASK>
static int launch_attn_tc2(const CUtensorMap& map, const AttnTc2Params& ap, const rho_attn_params* p, cudaStream_t st) {
  constexpr size_t SMEM = (size_t)(2 + 2 * A2_ST) * A2_TILE + 1024;
  auto kern = attn_tc2_kernel<POLY_MASK>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "attention_tc2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid((unsigned)ceil_div(p->T, 2 * A2_BM), p->heads, p->N);
  kern<<<grid, 320, SMEM, st>>>(map, ap, (bf16*)p->out, p->lse);
  RHO_CHECK_LAUNCH("attention_fwd_tcgen05_v2");
  return RHO_OK;
}

int attention_fwd_tc2(const rho_attn_params* p, cudaStream_t st) {
  auto enc = get_encode();
  if (!enc) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  const int C = p->heads * p->ch;
  CUtensorMap map;
  cuuint64_t gdim[3] = {(cuuint64_t)(3 * C), (cuuint64_t)p->T, (cuuint64_t)p->N};
  cuuint64_t gstr[2] = {(cuuint64_t)(3 * C) * 2, (cuuint64_t)(3 * C) * 2 * (cuuint64_t)p->T};
  cuuint32_t bdim[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p->qkv), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(RHO_ERR_CUDA, "cuTensorMapEncodeTiled(qkv) failed: %d", (int)r);
  AttnTc2Params ap;
  ap.T = p->T; ap.heads = p->heads; ap.C = C;
  if (p->order == 0) { ap.q_col = 0; ap.k_col = p->ch; ap.v_col = 2 * p->ch; ap.head_stride = 3 * p->ch; }
  else { ap.q_col = 0; ap.k_col = C; ap.v_col = 2 * C; ap.head_stride = p->ch; }
  ap.scale = 1.0f / sqrtf((float)p->ch);
  ap.scale_log2 = ap.scale * 1.4426950408889634f;
  // element pairs (of every 8) whose exponentials run on the FMA pipes instead of the MUFU: RHO_ATTN_POLY = 0 | 1 | 2 | 3 | 4.
  // Measured on B200 (T = 16384, ch = 64, N = 8): 760 / 845 / 890 / 897 / 848 TFLOP/s -- unlike the first-generation kernel this one
  // is MUFU-limited, so the default is 3 of every 8 pairs on the FMA pipes.
  static int poly = -1, cut = 1;
  if (poly < 0) {
    const char* e = getenv("RHO_ATTN_POLY"); poly = e ? atoi(e) : 3;
    const char* f = getenv("RHO_ATTN_LDSPLIT"); cut = f ? atoi(f) : 0;
  }
  ap.ld_split = cut;
  ap.trace = g_attn_trace;
  switch (poly) {
    case 0: return launch_attn_tc2<0x00>(map, ap, p, st);
    case 1: return launch_attn_tc2<0x02>(map, ap, p, st);
    case 2: return launch_attn_tc2<0x22>(map, ap, p, st);
    case 4: return launch_attn_tc2<0xAA>(map, ap, p, st);
    default: return launch_attn_tc2<0x4A>(map, ap, p, st);
  }
}

}  // namespace rho

extern "C" void rho_debug_attention_trace(void* buf) { rho::attention_tc2_set_trace(reinterpret_cast<long long*>(buf)); }
