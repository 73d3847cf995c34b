// Halo-reuse variant of the tcgen05 implicit-GEMM conv (included by conv_tc.cu) for the layers whose per-tap operand
// reloads make the generic kernel L2 -> shared-memory bound: stride-1 3x3(x3) convs with Cout <= 128 on wide planes
// (the 64^3 and 32x32x64 levels of the 3-D UNet: 128 output positions x 64 channels need 27 x 24 KB of TMA traffic per
// 28 MFLOP tile there, which caps the generic kernel at ~635 TFLOP/s).
//
// Idea: index one (n, d) plane by the PADDED-FLAT position q = h * Wp + (w + 1), Wp = W + 1 (one zero column shared by the
// right edge of row h and the left edge of row h + 1).  In that index space every in-plane filter tap (dh, dw) is a pure
// row shift by dh * Wp + dw.  So ONE TMA box load of rows [h_lo, h_lo + R) x columns [-1, W) x 64 channels per depth
// slice (out-of-bounds rows / the w = -1 column are zero-filled by the TMA unit = "same" padding) serves all 9 in-plane
// taps: the UMMA A descriptor simply starts (q0 + dh * Wp + dw - h_lo * Wp) rows into the slab.  A 128-byte-swizzled
// K-major operand may start at any 128-byte row because the swizzle is a function of the absolute shared-memory address
// (verified on sm_100a with tools/ubench/umma_shift.cu; base_offset stays 0).
// One CTA owns MT consecutive 128-position tiles (MT accumulators in TMEM), so a weight tile is loaded once per MT MMAs
// groups and the slab halo (2 Wp + 2 rows) is amortised over MT * 128 rows: 4-5x less TMA traffic than per-tap loads.
// The kernel is persistent (one CTA per SM, static round-robin over work items) with double-buffered TMEM accumulators,
// so the epilogue of item i overlaps the MMAs of item i + 1.  Rows with q % Wp == 0 are the padding column: computed and
// dropped (1 / Wp waste).
// Warps: 0 = TMA producer of the activation slabs, 10 = TMA producer of the weight tiles, 1 = TMEM allocator + MMA issuer, 11 = second MMA issuer,
// 2..9 = epilogue (two warps per scheduler, alternating tiles of an
// item; a single epilogue warp per scheduler ran at 0.36 IPC and bounded the 32-channel layers).
// 1x1 convs (k1) run through the same pipeline as a streaming GEMM: one tap, shift 0, slabs without halo rows.  In the
// per-tap kernel their short K loops leave every tile latency-bound (3.0 TB/s); here they reach 4.9 - 5.2 TB/s.
#pragma once

namespace rho {

constexpr int HALO_MAX_STAGES = 8;

struct HaloParams {
  int N, D, H, W, Wp;
  int C1, C2, Cin, Cout;
  int kd, pd;                 // depth taps (1 or 3) and padding
  int k1;                     // 1x1(x1) conv: one tap, no halo rows (a streaming GEMM through the same persistent pipeline)
  int n_blocks;               // Cout / BLOCK_N
  int tiles_per_plane, groups_per_plane;
  long long total_items;
  int R;                      // slab rows (h extent of the A box)
  int a_stage_bytes;          // R * Wp * 128 rounded up to 1024
  int SA, SB;                 // ring depths
  int out_split;              // > 0: channels >= out_split go to y2 (dgrad of a concat input)
  int issuers;                // MMA issuer threads: 2 (warps 1 and 11 take the even / odd tiles of an item; MT >= 2) or 1 (RHO_HALO_ISSUERS)
  float* stat_part;           // optional fused GroupNorm statistics: one row of channel sums per work item (see rho_b200.h)
};

struct HaloMaps {
  CUtensorMap a[2];
  CUtensorMap w;
};

template <int BLOCK_N, int MT, int BLOCK_K>
__global__ void __launch_bounds__(384, 1) conv_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloParams p,
                                                           const float* __restrict__ bias, const bf16* __restrict__ residual,
                                                           bf16* __restrict__ y, bf16* __restrict__ y2) {
  constexpr int ROW_BYTES = BLOCK_K * 2;          // one position's channels of a K block: 128 B (64 ch) or 64 B (32 ch)
  constexpr int B_BYTES = BLOCK_N * ROW_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * MT * BLOCK_N;
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two >= 32");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.SA * p.a_stage_bytes;
  __shared__ uint64_t a_full[HALO_MAX_STAGES], a_empty[HALO_MAX_STAGES], b_full[HALO_MAX_STAGES], b_empty[HALO_MAX_STAGES];
  __shared__ uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sbias[128];
  __shared__ float sstat[8][BLOCK_N][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.Cin / BLOCK_K;

  if (threadIdx.x == 0) {
    const uint32_t ni = (uint32_t)p.issuers;      // every issuer commits on the barriers that release operands / publish accumulators
    for (int s = 0; s < HALO_MAX_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], ni); mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], ni); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], ni); mbar_init(&acc_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // work item -> (output-channel block, tile group, plane)
  auto decode = [&](long long item, int& nb, int& q0, int& mt_here, int& d, int& n, int& h_lo) {
    nb = (int)(item % p.n_blocks);
    long long g = item / p.n_blocks;
    const int group = (int)(g % p.groups_per_plane);
    const long long plane = g / p.groups_per_plane;
    d = (int)(plane % p.D);
    n = (int)(plane / p.D);
    q0 = group * MT * 128;
    mt_here = min(MT, p.tiles_per_plane - group * MT);
    h_lo = p.k1 ? q0 / p.Wp : (q0 + p.Wp - 1) / p.Wp - 2;         // 3x3: floor((q0 - Wp - 1) / Wp) for q0 >= 0
  };

  if (warp == 0) {
    // ===== TMA producer of the activation slabs =====
    // (its own warp: with one producer for both operands the next slab could only be requested once all nine weight tiles of the
    // current one were in flight, i.e. at most SB taps ahead of the MMAs -- less than a slab's HBM latency)
    if (elect_one()) {
      int sa = 0;
      uint32_t pha = 0;
      for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int nb, q0, mt_here, d, n, h_lo;
        decode(item, nb, q0, mt_here, d, n, h_lo);
        for (int kb = 0; kb < KB; ++kb) {
          const int c = kb * BLOCK_K;
          const int src = c >= p.C1 ? 1 : 0;
          const int cc = src ? c - p.C1 : c;
          for (int dd = 0; dd < p.kd; ++dd) {
            mbar_wait(&a_empty[sa], pha ^ 1);
            mbar_expect_tx(&a_full[sa], (uint32_t)(p.R * p.Wp * ROW_BYTES));
            tma_load_5d(sA + (size_t)sa * p.a_stage_bytes, &maps.a[src], &a_full[sa], cc, -1, h_lo, d + dd - p.pd, n);
            if (++sa == p.SA) { sa = 0; pha ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // ===== TMA producer of the weight tiles =====
    if (elect_one()) {
      const int ntap = p.k1 ? 1 : 9;
      int sb = 0;
      uint32_t phb = 0;
      for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int nb = (int)(item % p.n_blocks);
        for (int kb = 0; kb < KB; ++kb) {
          const int c = kb * BLOCK_K;
          for (int dd = 0; dd < p.kd; ++dd) {
            for (int t9 = 0; t9 < ntap; ++t9) {
              mbar_wait(&b_empty[sb], phb ^ 1);
              mbar_expect_tx(&b_full[sb], B_BYTES);
              tma_load_2d(sB + sb * B_BYTES, &maps.w, &b_full[sb], (dd * 9 + t9) * p.Cin + c, nb * BLOCK_N);
              if (++sb == p.SB) { sb = 0; phb ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || (warp == 11 && p.issuers == 2)) {
    // ===== MMA issuer(s) =====
    // Two issuers (warps 1 and 11, on different schedulers) split the tiles of a work item: even tiles / odd tiles.  tcgen05.mma
    // may be issued by any thread; the two streams target different accumulators, each thread's tcgen05.commit covers its own
    // MMAs, and every operand / accumulator barrier counts both commits.  One thread could not issue N = 64 MMAs as fast as the
    // tensor pipe executes them even in the lean form below (~70 clk per MMA against 45).
    // Lean issue loop: ncu's warp-state samples of the first version (one warp, descriptors assembled with masks and shifts per MMA, a
    // division per tap, elect / reconverge per tap) put the issuer in scalar code ~70 % of the time and almost never on a full MMA
    // queue -- an MMA cost 85 - 98 clk to ISSUE while the tensor pipe needs 45 (N = 64) / 64 (N = 128) clk to execute it.  Here one
    // thread per issuer runs the whole loop, the nine tap shifts are registers, and a descriptor is a per-tap base plus an immediate
    // (its low word is the shared-memory address in 16-byte units, < 2^14 for any address below 256 KB).
    if (elect_one()) {
      const int first_tile = warp == 1 ? 0 : 1, tile_step = p.issuers;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t RB16 = ROW_BYTES / 16;                     // descriptor address units per operand row
      constexpr uint32_t DESC_HI = (uint32_t)((8 * ROW_BYTES) >> 4) | (1u << 14) | ((ROW_BYTES == 128 ? 2u : (ROW_BYTES == 64 ? 4u : 6u)) << 29);
      const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(sB) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t a_stage_lo = (uint32_t)p.a_stage_bytes >> 4;
      int tap_off[9];
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9) tap_off[t9] = p.k1 ? 0 : ((t9 / 3 - 1) * p.Wp + (t9 % 3 - 1)) * (int)RB16;
      int sa = 0, sb = 0, ab = 0;
      uint32_t pha = 0, phb = 0, phacc = 0;
      for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        int nb, q0, mt_here, d, n, h_lo;
        decode(item, nb, q0, mt_here, d, n, h_lo);
        mbar_wait(&acc_empty[ab], phacc ^ 1);       // the epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * MT * BLOCK_N);
        const uint32_t item_off = (uint32_t)(q0 - h_lo * p.Wp) * RB16;
        uint32_t acc = 0;                           // accumulate flag of each tile's k = 0 MMA: 0 on the first tap of the item only
        auto tap = [&](uint32_t a_t) {
          mbar_wait(&b_full[sb], phb);
          tc_fence_after();
          const uint64_t bd = ((uint64_t)DESC_HI << 32) | (uint64_t)(b_lo0 + (uint32_t)sb * (uint32_t)(B_BYTES >> 4));
          const uint64_t ad = ((uint64_t)DESC_HI << 32) | (uint64_t)a_t;
          if (tile_step == 1) {
#pragma unroll
            for (int i = 0; i < MT; ++i) {
              if (i < mt_here) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)
                  umma_f16(tacc + (uint32_t)(i * BLOCK_N), ad + (uint64_t)((uint32_t)(i * 128) * RB16 + (uint32_t)(2 * k)), bd + (uint64_t)(2 * k), idesc,
                           k == 0 ? acc : 1u);
              }
            }
          } else {
            const uint64_t ad1 = ad + (uint64_t)((uint32_t)(first_tile * 128) * RB16);
            const uint32_t tacc1 = tacc + (uint32_t)(first_tile * BLOCK_N);
#pragma unroll
            for (int h = 0; h < (MT + 1) / 2; ++h) {                // this issuer's tiles: first_tile + 2 h
              if (first_tile + 2 * h < mt_here) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / 16; ++k)
                  umma_f16(tacc1 + (uint32_t)(2 * h * BLOCK_N), ad1 + (uint64_t)((uint32_t)(2 * h * 128) * RB16 + (uint32_t)(2 * k)), bd + (uint64_t)(2 * k),
                           idesc, k == 0 ? acc : 1u);
              }
            }
          }
          umma_commit(&b_empty[sb]);
          acc = 1;
          if (++sb == p.SB) { sb = 0; phb ^= 1; }
        };
        for (int kb = 0; kb < KB; ++kb) {
          for (int dd = 0; dd < p.kd; ++dd) {
            mbar_wait(&a_full[sa], pha);
            const uint32_t a_base = a_lo0 + (uint32_t)sa * a_stage_lo + item_off;
            if (p.k1) {
              tap(a_base);
            } else {
#pragma unroll
              for (int t9 = 0; t9 < 9; ++t9) tap(a_base + (uint32_t)tap_off[t9]);
            }
            umma_commit(&a_empty[sa]);
            if (++sa == p.SA) { sa = 0; pha ^= 1; }
          }
        }
        umma_commit(&acc_full[ab]);
        if (++ab == 2) { ab = 0; phacc ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 10) {
    // ===== epilogue: 8 warps; warp w owns TMEM lanes 32 * (w % 4) .. + 31, warps 2..5 take the even tiles of an item and
    // warps 6..9 the odd ones (two warps per scheduler: one hides the other's tcgen05.ld / store latency).  The fused
    // GroupNorm statistics are accumulated per lane (= per row) over the warp's tiles and transposed across the warp once
    // per item and 32-channel chunk. =====
    const int et = threadIdx.x - 64;            // 0 .. 255
    const int ew = warp - 2;                    // 0 .. 7
    const int set = ew >> 2;
    const int lane_base = 32 * (warp & 3);
    const int r = lane_base + lane;
    int ab = 0;
    uint32_t phacc = 0;
    int nb_loaded = -1;
    for (long long item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int nb, q0, mt_here, d, n, h_lo;
      decode(item, nb, q0, mt_here, d, n, h_lo);
      if (nb != nb_loaded) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < BLOCK_N; i += 256) sbias[i] = bias ? bias[nb * BLOCK_N + i] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        nb_loaded = nb;
      }
      mbar_wait(&acc_full[ab], phacc);
      tc_fence_after();
      const long long plane_pos = ((long long)n * p.D + d) * p.H * p.W;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        float a1[32], a2[32];                   // this lane's row sums / sums of squares of channels c0 .. c0 + 31
#pragma unroll
        for (int j = 0; j < 32; ++j) { a1[j] = 0.f; a2[j] = 0.f; }
        const int ch = nb * BLOCK_N + c0;
#pragma unroll 1
        for (int i = set; i < mt_here; i += 2) {
          const int q = q0 + i * 128 + r;
          const int h = q / p.Wp, wq = q - h * p.Wp;
          const bool valid = wq != 0 && h < p.H;
          const long long pos = plane_pos + (long long)h * p.W + (wq - 1);
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)((ab * MT + i) * BLOCK_N + c0), v);
          bf16* dst;
          if (p.out_split > 0 && ch >= p.out_split) dst = y2 + pos * (p.Cout - p.out_split) + (ch - p.out_split);
          else dst = y + pos * (p.out_split > 0 ? p.out_split : p.Cout) + ch;
          const bf16* res = (residual && valid) ? residual + pos * p.Cout + ch : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 16) {
            uint4 w4[2];
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int jj = j + hf * 8;
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[jj + e]) + sbias[c0 + jj + e];
              if (res) {
                const uint4 qv = __ldg(reinterpret_cast<const uint4*>(res + jj));
                const uint32_t u[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { o[2 * e] += __uint_as_float(u[e] << 16); o[2 * e + 1] += __uint_as_float(u[e] & 0xffff0000u); }
              }
              __nv_bfloat162 h2;
              h2 = __floats2bfloat162_rn(o[0], o[1]); w4[hf].x = *reinterpret_cast<uint32_t*>(&h2);
              h2 = __floats2bfloat162_rn(o[2], o[3]); w4[hf].y = *reinterpret_cast<uint32_t*>(&h2);
              h2 = __floats2bfloat162_rn(o[4], o[5]); w4[hf].z = *reinterpret_cast<uint32_t*>(&h2);
              h2 = __floats2bfloat162_rn(o[6], o[7]); w4[hf].w = *reinterpret_cast<uint32_t*>(&h2);
            }
            if (valid) {
              st_global_256(dst + j, w4[0], w4[1]);       // a full 32-byte sector per lane (rows are >= 64 bytes apart)
              if (p.stat_part) {                // statistics of the bf16-rounded outputs, exactly what the next GroupNorm reads
                const uint32_t u[8] = {w4[0].x, w4[0].y, w4[0].z, w4[0].w, w4[1].x, w4[1].y, w4[1].z, w4[1].w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float x0 = __uint_as_float(u[e] << 16), x1 = __uint_as_float(u[e] & 0xffff0000u);
                  a1[j + 2 * e] += x0; a2[j + 2 * e] = fmaf(x0, x0, a2[j + 2 * e]);
                  a1[j + 2 * e + 1] += x1; a2[j + 2 * e + 1] = fmaf(x1, x1, a2[j + 2 * e + 1]);
                }
              }
            }
          }
        }
        if (p.stat_part) {
          sstat[ew][c0 + lane][0] = warp_colsum32_pass<false>(a1, lane);
          sstat[ew][c0 + lane][1] = warp_colsum32_pass<false>(a2, lane);
        }
      }
      if (p.stat_part) {
        // the eight warps' sums in a fixed order -> one row of stat_part per item
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et < BLOCK_N) {
          float t1 = 0.f, t2 = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) { t1 += sstat[k][et][0]; t2 += sstat[k][et][1]; }
          const long long row = item / p.n_blocks;       // (sample, plane, group) in launch order = n * rows_per_sample + ...
          *reinterpret_cast<float2*>(p.stat_part + (row * p.Cout + nb * BLOCK_N + et) * 2) = make_float2(t1, t2);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
      if (++ab == 2) { ab = 0; phacc ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// ---- host ---------------------------------------------------------------------------------------------------------
struct HaloPlan {
  bool ok;
  int bn, mt, bk;
  HaloParams p;
  size_t smem;
};

static int halo_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RHO_CONV_HALO"); v = e ? atoi(e) : 1; }
  return v;
}

// Cin = K channels (C1 + C2, each a multiple of 64), Cout = output channels of this launch
static HaloPlan halo_plan(int N, int D, int H, int W, int C1, int C2, int Cout, int kd, int kh, int kw) {
  HaloPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (halo_mode() == 0 || tc_disabled()) return pl;
  const bool k1 = kd == 1 && kh == 1 && kw == 1;
  if (!k1 && (kh != 3 || kw != 3 || (kd != 1 && kd != 3))) return pl;
  const int bk = (C1 % 64 == 0 && C2 % 64 == 0) ? 64 : ((C1 % 32 == 0 && C2 % 32 == 0) ? 32 : 0);
  // RHO_HALO_WIDE_K1=1 (opt-in): wide 1x1 convs (the attention blocks' qkv / proj_out: 256 -> 768 / 256 over T = 16384 positions)
  // stream through this pipeline in 128-channel output blocks, 1-D sequences re-factored into rows of <= 255 positions.  Parity-
  // green (tests/test_gpu_switches.py) but measured SLOWER than the per-tap kernel (256 -> 768: 170 vs 146 us, 1.6 vs 1.8 TB/s:
  // every output block reloads the A slab, 128 B/clk/SM of L2 -> shared-memory traffic) and the power-capped sampling step is
  // unchanged (48.18 ms both ways, profiles/r02_wide_k1_streaming.txt), so it stays off; the A-stationary loop order is the fix.
  static int wide_k1 = -1;
  if (wide_k1 < 0) { const char* e = getenv("RHO_HALO_WIDE_K1"); wide_k1 = (e && e[0] == '1') ? 1 : 0; }
  const int n_blocks = (k1 && wide_k1 && Cout > 128 && Cout % 128 == 0 && Cout <= 4096) ? Cout / 128 : 1;
  const int bn = n_blocks > 1 ? 128 : Cout;
  if (bk == 0 || (bn != 32 && bn != 64 && bn != 128) || (bn == 128 && bk != 64)) return pl;
  if (k1 && wide_k1 && (W < 16 || W > 255 || (long long)H * (W + 1) < 256)) {
    // a 1x1 conv is position-independent: any factorisation of the D H W positions of a sample into rows of W' columns serves
    // (1-D sequences arrive as D = H = 1, W = T)
    const long long S = (long long)D * H * W;
    int wf = 0;
    for (int c = 255; c >= 16 && !wf; --c)
      if (S % c == 0 && (S / c) * (c + 1) >= 256 && S / c <= 2147483647LL) wf = c;
    if (!wf) return pl;
    D = 1; H = (int)(S / wf); W = wf;
  }
  if (W < 16 || W > 255 || (long long)H * (W + 1) < 256) return pl;
  pl.bn = bn;
  pl.bk = bk;
  HaloParams& p = pl.p;
  p.N = N; p.D = D; p.H = H; p.W = W; p.Wp = W + 1;
  p.C1 = C1; p.C2 = C2; p.Cin = C1 + C2; p.Cout = Cout;
  p.kd = kd; p.pd = kd / 2;
  p.k1 = k1 ? 1 : 0;
  p.n_blocks = n_blocks;
  p.tiles_per_plane = (int)ceil_div((long long)H * p.Wp, 128);
  // static shared memory: statistics scratch (8 warps x BLOCK_N channels x 2 floats), bias, barriers; 1 KB alignment slack, 1 KB reserve
  const size_t budget = 227 * 1024 - ((size_t)64 * pl.bn + 1536) - 1024 - 1024;
  const size_t b_bytes = (size_t)pl.bn * (size_t)(2 * bk);
  // tiles per work item: as many as TMEM (2 x MT x BLOCK_N <= 512 columns) and two slab stages in shared memory allow
  bool fits = false;
  for (int mt = 256 / pl.bn; mt >= 1 && !fits; mt >>= 1) {
    pl.mt = mt;
    p.groups_per_plane = (int)ceil_div(p.tiles_per_plane, mt);
    int R = 0;
    for (int g = 0; g < p.groups_per_plane; ++g) {
      const int q0 = g * mt * 128;
      const int h_lo = k1 ? q0 / p.Wp : (q0 + p.Wp - 1) / p.Wp - 2;
      const int last = q0 + mt * 128 - 1 + (k1 ? 0 : p.Wp + 1);          // highest flat row any tap of the group touches
      R = std::max(R, last / p.Wp - h_lo + 1);
    }
    if (R > 256) continue;
    p.R = R;
    p.a_stage_bytes = (int)align_up((size_t)R * p.Wp * (size_t)(2 * bk), 1024);
    // weight ring: 5 (4) stages preferred, 3 at least, before giving up tiles per item
    for (int sb = pl.bn <= 64 ? 5 : 4; sb >= 3 && !fits; --sb) {
      const long long sa = ((long long)budget - (long long)sb * (long long)b_bytes) / p.a_stage_bytes;
      if (sa < 2) continue;
      p.SB = sb;
      p.SA = (int)std::min<long long>(sa, 3);
      fits = true;
    }
  }
  if (!fits) return pl;
  p.total_items = (long long)N * D * p.groups_per_plane * p.n_blocks;
  // spend what is left on a deeper weight ring
  long long extra = ((long long)budget - (long long)p.SA * p.a_stage_bytes) / (long long)b_bytes;
  p.SB = (int)std::min<long long>(HALO_MAX_STAGES, std::max<long long>(p.SB, extra));
  pl.smem = (size_t)p.SA * p.a_stage_bytes + (size_t)p.SB * b_bytes + 1024;
  pl.ok = true;
  return pl;
}

template <int BLOCK_N, int MT, int BLOCK_K>
static int launch_halo(const HaloMaps& maps, const HaloPlan& pl, const float* bias, const bf16* residual, bf16* y, bf16* y2, cudaStream_t st) {
  auto kern = conv_halo_kernel<BLOCK_N, MT, BLOCK_K>;
  static size_t attr_bytes = 0;
  if (pl.smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return fail(RHO_ERR_CUDA, "conv_halo: cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e));
    attr_bytes = pl.smem;
  }
  const unsigned grid = (unsigned)std::min<long long>(pl.p.total_items, (long long)sm_count());
  kern<<<grid, 384, pl.smem, st>>>(maps, pl.p, bias, residual, y, y2);
  RHO_CHECK_LAUNCH("conv_halo_tcgen05");
  return RHO_OK;
}

// x / x2: channels-last [N, D, H, W, C1 / C2]; w: packed [Cout][taps][Cin] bf16 (taps in (kd, kh, kw) order)
static int conv_halo_run(const HaloPlan& pl_in, const void* x, const void* x2, const void* w, const float* bias, const bf16* residual, bf16* y,
                         bf16* y2, int out_split, float* stat_part, cudaStream_t st) {
  HaloPlan pl = pl_in;
  HaloParams& p = pl.p;
  p.out_split = out_split;
  p.stat_part = stat_part;
  static int issuers = -1;
  if (issuers < 0) { const char* e = getenv("RHO_HALO_ISSUERS"); issuers = (e && e[0] == '1') ? 1 : 2; }
  p.issuers = pl.mt >= 2 ? issuers : 1;
  HaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int box[5] = {pl.bk, p.Wp, p.R, 1, 1};
  for (int s = 0; s < 2; ++s) {
    const int C = s ? p.C2 : p.C1;
    if (C == 0) continue;
    ActView v;
    v.base = s ? x2 : x;
    v.dim[0] = C; v.dim[1] = p.W; v.dim[2] = p.H; v.dim[3] = p.D; v.dim[4] = p.N;
    v.stride[0] = (long long)C * 2; v.stride[1] = v.stride[0] * p.W; v.stride[2] = v.stride[1] * p.H; v.stride[3] = v.stride[2] * p.D;
    int rc = encode_view(&maps.a[s], v, box, pl.bk);
    if (rc) return rc;
  }
  int rc = encode_weight_map(&maps.w, w, (long long)(p.k1 ? 1 : p.kd * 9) * p.Cin, p.Cout, pl.bk, pl.bn);
  if (rc) return rc;
#define RHO_HALO(BN, MT, BK) \
  if (pl.bn == BN && pl.mt == MT && pl.bk == BK) return launch_halo<BN, MT, BK>(maps, pl, bias, residual, y, y2, st)
  RHO_HALO(32, 8, 32); RHO_HALO(32, 4, 32); RHO_HALO(32, 2, 32); RHO_HALO(32, 1, 32);
  RHO_HALO(32, 8, 64); RHO_HALO(32, 4, 64); RHO_HALO(32, 2, 64); RHO_HALO(32, 1, 64);
  RHO_HALO(64, 4, 32); RHO_HALO(64, 2, 32); RHO_HALO(64, 1, 32);
  RHO_HALO(64, 4, 64); RHO_HALO(64, 2, 64); RHO_HALO(64, 1, 64);
  RHO_HALO(128, 2, 64); RHO_HALO(128, 1, 64);
#undef RHO_HALO
  return fail(RHO_ERR_UNSUPPORTED, "conv_halo: no kernel for BLOCK_N=%d MT=%d BLOCK_K=%d", pl.bn, pl.mt, pl.bk);
}

}  // namespace rho
