/*
 * rho_b200.h -- C ABI of the B200-native hot path of rho-diffusion (DDPM training step +
 * ancestral sampling loop over the conditional n-D UNet).
 *
 * This is the drop-in boundary.  The reference (intel/rho-diffusion) is 100 % Python and has no
 * native FFI; backend acceleration enters it through `ipex.optimize(model, optimizer=...)`
 * (scripts/training_ddp.py:182-183, scripts/inference.py:161-163), the Lightning XPU
 * accelerator/strategy registration (rho_diffusion/xpu.py:242, rho_diffusion/__init__.py:27-35)
 * and `dist.init_process_group("ccl"|"nccl")` (rho_diffusion/utils.py:45-81).  Those hooks hand a
 * torch module to a vendor library whose kernels then execute the ATen ops listed below.  This
 * header is what replaces that vendor library: one entry point per ATen op class on the hot path,
 * each citing the reference call site it stands in for.  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, POD parameter structs; no torch / C++ types.
 *   - every call is asynchronous on `stream`, never allocates, never synchronises and is therefore
 *     CUDA-graph capturable.  Scratch memory is caller-allocated: query `*_workspace_bytes` first.
 *   - return value: RHO_OK (0) or a negative RHO_ERR_* code; `rho_error_string` names it and
 *     `rho_last_error_message` (thread-local) gives detail.  Nothing throws.
 *   - activations are CHANNELS-LAST: [N, D, H, W, C] with D = 1 (and H = 1) for 2-D (1-D) data, C
 *     contiguous.  Parameters keep the PyTorch layouts of the reference state_dict unless a
 *     `rho_*_pack_*` call says otherwise.
 *   - dtype enums select the storage type of activations (RHO_F32 | RHO_BF16); all accumulation,
 *     normalisation statistics and softmax are fp32.
 *   - one process per GPU; calls are made from the thread that owns the stream; the library keeps
 *     no mutable global state except a mutex-guarded cache of kernel attributes.
 */
#ifndef RHO_B200_H_
#define RHO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* rho_stream_t; /* == cudaStream_t */

enum {
  RHO_OK = 0,
  RHO_ERR_INVALID_ARG = -1,
  RHO_ERR_UNSUPPORTED = -2,
  RHO_ERR_CUDA = -3,
  RHO_ERR_WORKSPACE = -4
};
enum { RHO_F32 = 0, RHO_BF16 = 1 };
enum { RHO_ACT_NONE = 0, RHO_ACT_SILU = 1, RHO_ACT_RELU = 2, RHO_ACT_TANH = 3, RHO_ACT_SIGMOID = 4, RHO_ACT_GELU = 5,
       RHO_ACT_SYMLOG = 6 /* SymmetricLog, rho_diffusion/activations.py:35-61: tanh(x) (1 + log(x tanh(x) + 1)) */ };
/* conv engine selection */
enum { RHO_CONV_AUTO = 0, RHO_CONV_SIMT = 1, RHO_CONV_TCGEN05 = 2 };

int rho_abi_version(void);
const char* rho_error_string(int code);
const char* rho_last_error_message(void);
/* number of kernels this library has launched in the calling process (all threads) since load */
uint64_t rho_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * layout: NC[D]HW <-> channels-last, with dtype conversion.  Replaces the implicit layout of
 * `x.type(self.dtype)` at the UNet boundary (rho_diffusion/models/unet_v2.py:722, 731).
 * S = product of spatial dims. */
int rho_nc_to_cl(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t N, int64_t C, int64_t S, rho_stream_t stream);
int rho_cl_to_nc(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t N, int64_t C, int64_t S, rho_stream_t stream);
/* dst = a (+ b) elementwise with dtype conversion; b may be NULL */
int rho_add_cast(const void* a, int a_dtype, const void* b, int b_dtype, void* dst, int dst_dtype, int64_t n, rho_stream_t stream);
/* Split-bf16 engine ("fp32-accurate on the tensor cores", the exact path the reference's TF32-off arithmetic asks for,
 * scripts/training_ddp.py:42-43): rho_split3_concat writes x = x1 + x2 + x3 (three bf16 terms, 24 mantissa bits) as the six
 * channel blocks [x1 | x1 | x2 | x1 | x2 | x3] of a channels-last bf16 tensor [rows, 6 (C1 + C2)] (x2 = optional second source
 * of a channel concat); rho_split3_weight expands fp32 weights [Cout, Cin, taps] to [Cout, 6 Cin, taps] with the bf16-exact terms
 * [w1 | w2 | w1 | w3 | w2 | w1].  rho_conv_fwd on those operands with y_f32 = 1 accumulates the six products in fp32 (TMEM). */
int rho_split3_concat(const float* x, int32_t C1, const float* x2, int32_t C2, void* out, int64_t rows, rho_stream_t stream);
int rho_split3_weight(const float* w, float* out, int32_t Cout, int32_t Cin, int32_t taps, rho_stream_t stream);
/* x *= (*scale_dev if scale_dev else 1) * scale_host, in place (upstream-gradient scaling without a host sync) */
int rho_scale(void* x, int dtype, int64_t n, const float* scale_dev, float scale_host, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * conv_nd: `conv_nd(dims, Cin, Cout, k, stride, padding)` -> nn.Conv{1,2,3}d
 * (rho_diffusion/layers.py:77-88; call sites unet_v2.py:120,155-162,215,241,248-256,323,331,535,682),
 * with the surrounding data movement fused in:
 *   - channel concat of two inputs, replacing `torch.cat([h, hs.pop()], 1)` (unet_v2.py:729)
 *   - nearest x2 upsample of the input, replacing `F.interpolate` (unet_v2.py:124-131)
 *   - bias, and residual add `skip_connection(x) + h` / `x + h` (unet_v2.py:293, 342)
 * k in {1,3} per dim ("same" padding k/2), stride in {1,2} per dim, upsample in {1,2} per dim.
 * Output dims: Do = (D*ud + 2*(kd/2) - kd)/sd + 1, etc. */
typedef struct rho_conv_params {
  int32_t N, D, H, W;        /* input dims BEFORE the optional upsample */
  int32_t C1, C2;            /* channels of x and of x2 (0 when x2 == NULL); Cin = C1 + C2 */
  int32_t Cout;
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t ud, uh, uw;
  int32_t act_dtype;         /* dtype of x, x2, residual, y (and dy, dx) */
  int32_t w_dtype;           /* dtype of the packed weights (RHO_F32 or RHO_BF16) */
  int32_t engine;            /* RHO_CONV_AUTO | _SIMT | _TCGEN05 */
  int32_t y_f32;             /* 1: y is written as fp32 regardless of act_dtype (final conv -> eps) */
  const void* x;
  const void* x2;
  const void* w_packed;      /* from rho_conv_pack_weight */
  const void* w_aux;         /* optional: from rho_conv_pack_weight_aux (pre-summed weights of the fused-upsample
                                decomposition used by the tcgen05 engine); NULL selects the CUDA-core gather instead */
  const float* bias;         /* [Cout] or NULL */
  const void* residual;      /* [N,Do,Ho,Wo,Cout] or NULL */
  void* y;                   /* [N,Do,Ho,Wo,Cout] */
  void* workspace;
  size_t workspace_bytes;
  /* optional fused GroupNorm statistics of the OUTPUT (the next layer of every conv in the UNet is a GroupNorm32,
   * layers.py:71-74): when non-NULL and rho_conv_stats_rows(p) = R > 0, the epilogue also writes, for every sample n, row
   * r < R and output channel c, the sum and the sum of squares of the bf16-rounded outputs of one tile:
   * stats_partial[((n * R + r) * Cout + c) * 2 + {0, 1}] (fp32, every slot written, fixed order: deterministic).
   * rho_gn_stats_from_partials turns them into the (mean, rstd) that rho_gn_apply / rho_gn_bwd consume. */
  float* stats_partial;
} rho_conv_params;
/* rows per sample of stats_partial for this geometry / engine, or 0 when the launch cannot produce statistics */
int32_t rho_conv_stats_rows(const rho_conv_params* p);

/* packed forward weights: [Cout][taps][Cin] (K-major rows, K = tap*Cin + ci, tap = (kd*KH + kh)*KW + kw).
 * `w` is the reference state_dict tensor, fp32 [Cout, Cin, kd, kh, kw]. */
int rho_conv_pack_weight(const float* w, void* packed, int packed_dtype, int32_t Cout, int32_t Cin, int32_t taps, rho_stream_t stream);
/* engine-specific second packing (see w_aux): bytes needed for this geometry (0: none exists), and the packing itself.
 * `w` is the fp32 state_dict tensor; sums are formed in fp32 and rounded once. */
size_t rho_conv_aux_weight_bytes(const rho_conv_params* p);
int rho_conv_pack_weight_aux(const rho_conv_params* p, const float* w, void* aux, rho_stream_t stream);
size_t rho_conv_workspace_bytes(const rho_conv_params* p);
int rho_conv_fwd(const rho_conv_params* p, rho_stream_t stream);
/* which engine rho_conv_fwd would use for p (RHO_CONV_SIMT or RHO_CONV_TCGEN05) */
int rho_conv_engine_for(const rho_conv_params* p);

/* backward of the above (autograd of nn.ConvNd in `loss.backward()`, scripts/training_ddp.py:197).
 * dgrad: dx (and dx2 for the concat partner) from dy, using the same packed forward weights.
 * wgrad: dw fp32 in the PyTorch layout [Cout, Cin, kd, kh, kw] and dbias [Cout] (either may be NULL). */
typedef struct rho_conv_bwd_params {
  rho_conv_params fwd;       /* geometry + x/x2 (wgrad) + w_packed (dgrad); y/bias/residual ignored */
  const void* dy;            /* [N,Do,Ho,Wo,Cout], act_dtype */
  void* dx;                  /* [N,D,H,W,C1] act_dtype, or NULL */
  void* dx2;                 /* [N,D,H,W,C2] act_dtype, or NULL */
  float* dw;                 /* or NULL */
  float* dbias;              /* or NULL */
  void* workspace;
  size_t workspace_bytes;
} rho_conv_bwd_params;
size_t rho_conv_bwd_workspace_bytes(const rho_conv_bwd_params* p);
int rho_conv_dgrad(const rho_conv_bwd_params* p, rho_stream_t stream);
int rho_conv_wgrad(const rho_conv_bwd_params* p, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm32 (+ embedding add / FiLM) + activation:
 *   `normalization(C)` = GroupNorm32(32, C) (layers.py:71-74, 122-129), the activation that follows it
 *   (unet_v2.py:214, 238, 681) and the conditioning in ResBlock._forward (unet_v2.py:285-292):
 *     FiLM   : act( GN(h) * (1 + scale) + shift )
 *     add    : act( GN(h + emb) )                      (pre_add)
 *   The input may be the channel concat of x and x2 (up-path ResBlocks), never materialised.
 * Two launches: rho_gn_stats (mean, rstd per (n, group); deterministic two-stage Welford merge) and
 * rho_gn_apply.  HBM-bound: 1 read for stats, 1 read + 1 write for apply. */
typedef struct rho_gn_params {
  int32_t N;
  int64_t S;                 /* spatial positions per sample */
  int32_t C1, C2;            /* C = C1 + C2 */
  int32_t groups;            /* 32 */
  int32_t in_dtype, out_dtype;
  int32_t act;
  float eps;
  const void* x;
  const void* x2;
  const float* pre_add;      /* [N, pre_add_stride] or NULL */
  int32_t pre_add_stride;
  const float* gamma;        /* [C] */
  const float* beta;         /* [C] */
  const float* scale;        /* FiLM scale [N, film_stride] or NULL */
  const float* shift;        /* FiLM shift, same stride */
  int32_t film_stride;
  float* stats;              /* [N, groups, 2] = (mean, rstd) */
  void* y;                   /* [N, S, C] out_dtype */
  void* workspace;
  size_t workspace_bytes;
} rho_gn_params;
size_t rho_gn_workspace_bytes(const rho_gn_params* p);
int rho_gn_stats(const rho_gn_params* p, rho_stream_t stream);
/* same result as rho_gn_stats (p->stats), but from the per-tile channel sums a conv epilogue left behind
 * (rho_conv_params.stats_partial): part1 / rows1 for x (C1 channels), part2 / rows2 for x2 (C2 channels, or NULL / 0).
 * Sums are merged in double precision in a fixed order; no pass over the activations.  Not usable with pre_add. */
int rho_gn_stats_from_partials(const rho_gn_params* p, const float* part1, int32_t rows1, const float* part2, int32_t rows2,
                               rho_stream_t stream);
int rho_gn_apply(const rho_gn_params* p, rho_stream_t stream);

typedef struct rho_gn_bwd_params {
  rho_gn_params fwd;         /* same tensors as the forward call (x, stats, gamma, FiLM...) ; y unused */
  const void* dy;            /* [N,S,C] out_dtype */
  void* dx;                  /* [N,S,C1] in_dtype */
  void* dx2;                 /* [N,S,C2] in_dtype or NULL */
  float* dgamma;             /* [C] */
  float* dbeta;              /* [C] */
  float* dscale;             /* [N, dfilm_stride] or NULL */
  float* dshift;
  int32_t dfilm_stride;
  float* dpre_add;           /* [N, C] or NULL */
  void* workspace;
  size_t workspace_bytes;
} rho_gn_bwd_params;
size_t rho_gn_bwd_workspace_bytes(const rho_gn_bwd_params* p);
int rho_gn_bwd(const rho_gn_bwd_params* p, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * QKVAttentionLegacy / QKVAttention (unet_v2.py:365-436): softmax_fp32((q*s)^T (k*s)) v, s = ch^-1/4,
 * flash-style (the T x T weight matrix is never materialised).  qkv is channels-last [N, T, 3*heads*ch];
 * order 0 (legacy): channel = head*3ch + {q,k,v}*ch + c;  order 1 (new): channel = {q,k,v}*heads*ch + head*ch + c. */
typedef struct rho_attn_params {
  int32_t N, T, heads, ch;
  int32_t order;
  int32_t dtype;
  const void* qkv;
  void* out;                 /* [N, T, heads*ch] */
  float* lse;                /* [N, heads, T] log-sum-exp of the scaled logits, or NULL (inference) */
} rho_attn_params;
int rho_attention_fwd(const rho_attn_params* p, rho_stream_t stream);
/* Debug hook (tools/attn_trace.py): `buf` = device array of 64*16 int64 (or NULL to switch off).  While set, CTA (0,0,0) of the
 * head-dim-64 tcgen05 forward writes clock64() stamps of its warp hand-offs, [key tile][slot]: slots 0/1 (2/3) = row owners of
 * query tile 0 (1) saw S ready / published P; slots 4/5 (6/7) = the MMA issuer saw P of tile 0 (1) / finished issuing for it. */
void rho_debug_attention_trace(void* buf);
typedef struct rho_attn_bwd_params {
  rho_attn_params fwd;       /* qkv, out, lse from the forward */
  const void* dout;          /* [N, T, heads*ch] */
  void* dqkv;                /* [N, T, 3*heads*ch] */
  void* workspace;
  size_t workspace_bytes;
} rho_attn_bwd_params;
size_t rho_attention_bwd_workspace_bytes(const rho_attn_bwd_params* p);
int rho_attention_bwd(const rho_attn_bwd_params* p, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * embedding path (all fp32; tiny):
 *   sinosoidal_position_embedding (models/common.py:27-43), nn.Linear of time_embed / emb_layers
 *   (unet_v2.py:521-525, 229-235), MultiEmbeddings.forward (models/conditioning.py:115-139). */
/* omega: device table [dim/2] fp32, omega[i] = wavelength^(2i/dim) (built once on the host like the schedule tables);
 * out[b, 2i] = sin(t[b]/omega[i]), out[b, 2i+1] = cos(t[b]/omega[i]) */
int rho_timestep_embedding(const int64_t* t, int32_t B, int32_t dim, const float* omega, float* out, rho_stream_t stream);
typedef struct rho_linear_params {
  int32_t B, I, O;
  int32_t act_in;            /* y = W act_in(x) + b (+ add) */
  const float* x;  int32_t x_stride;
  const float* w;            /* [O, I] */
  const float* b;            /* [O] or NULL */
  const float* add; int32_t add_stride; /* optional [B, O] */
  float* y; int32_t y_stride;
} rho_linear_params;
int rho_linear_fwd(const rho_linear_params* p, rho_stream_t stream);
/* dx = act_in'(x) * (W^T dy) ; dw = dy^T act_in(x) ; db = sum_b dy.  Any output may be NULL. */
int rho_linear_bwd(const rho_linear_params* p, const float* dy, int32_t dy_stride, float* dx, int32_t dx_stride,
                   float* dw, float* db, rho_stream_t stream);
/* out[b,:] (+)= table[index of y[b*y_stride] in values[0..V)] ; idx_out[b] receives the index (-1: no match) */
int rho_embedding_match(const float* y, int32_t y_stride, int32_t B, const float* values, int32_t V,
                        const float* table, int32_t dim, float* out, int32_t accumulate, int32_t* idx_out,
                        rho_stream_t stream);
/* dtable[idx[b], :] += dout[b, :]  (dtable must be zero-initialised by the caller; deterministic order) */
int rho_embedding_match_bwd(const int32_t* idx, int32_t B, const float* dout, int32_t dim, float* dtable, int32_t V,
                            rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * diffusion math (rho_diffusion/diffusion/ddpm.py):
 *   rho_randn            torch.randn_like (ddpm.py:101-102, 171, 196-199): Philox4x32-10 + Box-Muller,
 *                        element i of call (seed, offset) is counter (i/4, offset), lane i%4.
 *   rho_q_sample         forward_process (ddpm.py:104-130): x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) eps
 *   rho_p_sample_step    one iteration of reverse_process (ddpm.py:196-218), step index read from device
 *                        memory so that a captured CUDA graph can be replayed T times:
 *                          t == 0 : x unchanged;  t <= 1 : z = 0
 *                          x <- clamp((x - beta_t/sqrt(1-abar_t) eps)/sqrt(alpha_t) + noise_scale sqrt(beta_t) z, lo, hi)
 *                        z is read from `z` if non-NULL, else generated in-kernel from (seed, t).
 *   rho_advance_timestep t_vec[b] = *t - 1 ; *t = *t - 1   (the host loop `for t in steps`, ddpm.py:195-205)
 *   rho_mse_loss         nn.MSELoss (ddpm.py:280) and its gradient, deterministic two-stage reduction. */
int rho_randn(float* out, int64_t n, uint64_t seed, uint64_t offset, rho_stream_t stream);
/* test hook: raw Philox4x32-10 words, out[4g + j] = philox(counter = (g, offset), key = seed)[j] */
int rho_philox_raw(uint32_t* out, int64_t groups, uint64_t seed, uint64_t offset, rho_stream_t stream);
int rho_q_sample(const float* x0, const float* noise, const int64_t* t, const float* alpha_bar, int32_t T,
                 int32_t N, int64_t per_sample, float* x_t, rho_stream_t stream);
typedef struct rho_p_sample_params {
  float* x;                  /* in/out, fp32 */
  const float* eps;          /* model output, fp32 */
  const float* z;            /* optional externally supplied noise for this step */
  const int64_t* t;          /* device scalar: current step index */
  const float* alpha; const float* beta; const float* alpha_bar; /* [T] fp32 tables */
  int32_t T;
  int64_t n;                 /* total elements */
  float noise_scale;         /* 0.8 (ddpm.py:215) */
  float clamp_lo, clamp_hi;  /* -1, 1 (ddpm.py:218) */
  uint64_t seed;
  const uint64_t* seed_dev;  /* optional device scalar: when non-NULL it replaces `seed` (one captured graph then serves every
                              * reverse_process call: the host writes the run's seed into it before replaying) */
} rho_p_sample_params;
int rho_p_sample_step(const rho_p_sample_params* p, rho_stream_t stream);
int rho_advance_timestep(int64_t* t_scalar, int64_t* t_vec, int32_t B, rho_stream_t stream);
size_t rho_mse_workspace_bytes(int64_t n);
int rho_mse_loss(const void* pred, int pred_dtype, const float* target, int64_t n, float* loss,
                 void* dpred /* same dtype as pred, or NULL */, float grad_scale, void* workspace, size_t workspace_bytes,
                 rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * optimizer: torch.optim.AdamW as configured by configure_optimizers (abstract_diffusion.py:86-148)
 * fused with ExponentialMovingAverage.update (ema.py:41-57), over flat fp32 buffers:
 *   g' = g * grad_scale ; p *= 1 - lr*wd ; m,v update ; p -= lr/bc1 * m/(sqrt(v)/sqrt(bc2) + eps)
 *   shadow -= (1 - ema_decay_t) * (shadow - p)      (skipped when shadow == NULL)
 *   p_bf16 (optional) receives the bf16 copy used by the next forward. */
typedef struct rho_adamw_params {
  float* p; const float* g; float* m; float* v; float* shadow;
  int64_t n;
  double lr, beta1, beta2, eps, weight_decay; /* doubles: 1 - beta2 must be formed in double, as torch does */
  int32_t step;              /* 1-based */
  float ema_decay_t;         /* d * (1 - exp(-step/2000)) */
  float grad_scale;
  /* optional device-resident step counter (so that a CUDA graph holding the whole training step can be replayed):
   * a 64-byte device buffer whose first int32 is the number of steps taken so far.  When non-NULL the call increments it
   * on the device and derives the bias corrections and the EMA warm-up factor (ema_decay * (1 - exp(-step / 2000)),
   * ema.py:33) from it; `step` and `ema_decay_t` are ignored. */
  int32_t* step_dev;
  float ema_decay;
} rho_adamw_params;
int rho_adamw_ema_step(const rho_adamw_params* p, rho_stream_t stream);
/* stand-alone ExponentialMovingAverage.update for one tensor (ema.py:52-57): shadow -= one_minus_decay * (shadow - p) */
int rho_ema_update(float* shadow, const float* p, int64_t n, float one_minus_decay, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * GaussianDiffusionPipeline (rho_diffusion/diffusion/gaussian_diffusion.py): the x0-prediction pipeline that
 * scripts/inference.py:122-176 drives (cosine schedule, FIXED_LARGE variance, dynamic thresholding, DDIM eta = 0).
 *   rho_lincomb2      out = ca[t_n] * a + cb[t_n] * b per sample: q_sample (gaussian_diffusion.py:294-312) with the
 *                     reference's float64-built sqrt_alphas_cumprod / sqrt_one_minus_alphas_cumprod tables cast to fp32
 *                     (_extract_into_tensor, :91-104)
 *   rho_abs_quantile  out[n] = max(floor_value, torch.quantile(|x[n, :]|, q)) with torch's linear interpolation
 *                     between the order statistics k_lo = floor(q (P-1)) and k_hi = ceil(q (P-1)), weight = frac
 *                     (process_xstart, :400-408).  Exact radix selection on the bit patterns, no sort.
 *   rho_x0_step       one sampling update for a model that predicts x0 (:338-443):
 *                       pred = thresh ? clamp(model_out, -s_n, s_n) / s_n : model_out
 *                       mode 0, ddim_sample (:654-702): eps = (sqrt_recip_abar x - pred) / sqrt_recipm1_abar,
 *                               sigma = eta sqrt((1-abar_prev)/(1-abar)) sqrt(1 - abar/abar_prev),
 *                               x <- pred sqrt(abar_prev) + sqrt(1 - abar_prev - sigma^2) eps + [t != 0] sigma noise
 *                       mode 1, p_sample (:512-556):  x <- coef1 pred + coef2 x + [t != 0] exp(0.5 logvar) noise
 *                     t comes per sample from `t` or, for graph replay, from the device scalar `t_scalar`. */
int rho_lincomb2(const float* a, const float* b, const int64_t* t, const float* ca, const float* cb, int32_t T,
                 int32_t N, int64_t per_sample, float* out, rho_stream_t stream);
size_t rho_abs_quantile_workspace_bytes(int32_t N);
int rho_abs_quantile(const float* x, int32_t N, int64_t per_sample, int64_t k_lo, int64_t k_hi, float weight,
                     float floor_value, float* out, void* workspace, size_t workspace_bytes, rho_stream_t stream);
typedef struct rho_x0_step_params {
  const void* model_out;     /* [N, per_sample] model prediction of x0, NC[D]HW order */
  int32_t model_dtype;       /* RHO_F32 | RHO_BF16 */
  int32_t mode;              /* 0 = DDIM, 1 = ancestral posterior */
  float* x;                  /* in/out x_t -> x_{t-1}, fp32 */
  const float* noise;        /* optional: NULL drops the noise term (mode 0 with eta == 0; mode 1 => the posterior mean) */
  float* pred_xstart;        /* optional output */
  const float* thresh;       /* [N] dynamic threshold s_n >= 1, or NULL for no clipping */
  const int64_t* t;          /* [N] per-sample step, or NULL */
  const int64_t* t_scalar;   /* device scalar step (takes precedence), or NULL */
  int32_t T, N;
  int64_t per_sample;
  float eta;
  const float* alphas_cumprod; const float* alphas_cumprod_prev;
  const float* sqrt_recip_alphas_cumprod; const float* sqrt_recipm1_alphas_cumprod;       /* mode 0 */
  const float* posterior_mean_coef1; const float* posterior_mean_coef2; const float* model_log_variance; /* mode 1 */
} rho_x0_step_params;
int rho_x0_step(const rho_x0_step_params* p, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * optional resampling / regularisation paths of the UNet (channels-last [N, D, H, W, C], C % 4 == 0):
 *   rho_resample  mode 0: avg_pool_nd with kernel = stride = (fd, fh, fw), floor extents (layers.py:90-102;
 *                         Downsample(use_conv=False), unet_v2.py:163-165; ResBlock(down=True), :224-225)
 *                 mode 1: nearest upsample by (fd, fh, fw) (F.interpolate, unet_v2.py:124-131, without the conv)
 *                 mode 2: sum over the (fd, fh, fw) children = adjoint of mode 1
 *                 mode 3: adjoint of mode 0 (parent / window; positions the pooling dropped get 0)
 *                 (D, H, W) are the extents of `src`, (Do, Ho, Wo) those of `dst`.
 *   rho_dropout   nn.Dropout(p) of ResBlock.out_layers (unet_v2.py:239): y = x * [u >= p] / (1 - p) with u from
 *                 Philox4x32-10(counter = i / 4, offset; key = seed); the same call on dy is the backward.
 *                 `step_dev` (optional): device int32 step counter (rho_adamw_params.step_dev); its value is folded into the
 *                 high word of the Philox offset, so a training step replayed from a CUDA graph draws a new mask every step. */
int rho_resample(const void* src, void* dst, int dtype, int32_t N, int32_t D, int32_t H, int32_t W, int32_t C, int32_t fd, int32_t fh,
                 int32_t fw, int32_t mode, int32_t Do, int32_t Ho, int32_t Wo, rho_stream_t stream);
int rho_dropout(const void* x, void* y, int dtype, int64_t n, float p, uint64_t seed, uint64_t offset, const int32_t* step_dev,
                rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * synthetic density-field generators on the device (keeps the input pipeline off the host):
 *   rho_sph_harm_density  out[n, i, j, k] (fp32, [N, G, G, G]) = the normalised density of compute_spherical_harmonic
 *                         (rho_diffusion/data/synthetic.py:84-124) for (l[n], m[n]) on the grid linspace(lo, hi, G)^3 with the
 *                         reference's meshgrid(indexing="xy") / swapped-angle / complex min-max conventions; double precision
 *                         inside.  The lexicographic complex extremes are decided by rounding noise at symmetric grid points, so
 *                         callers pass the host restatement's extremes per sample (`minmax`); NULL = the kernel's own.
 *                         workspace: rho_sph_harm_workspace_bytes(N, G).
 *   rho_gaussian_blobs    out[n] (fp32, [N, D, H, W], D = 1 / H = 1 for lower n) = 2 min(sum_k exp(-|x - c_k|^2 / 2 w_n^2), 1) - 1
 *                         with n_blobs[n] (<= 8) centres from Philox4x32-10(counter = (first_index + n, k), key = seed), each
 *                         coordinate 0.2 + 0.6 u, x at cell centres of the unit box (the last axis is coordinate 0).
 *   rho_lineprofile       out[b, f] (fp32, [B, G]) = the rotational spectrum of SpectroscopyDataset.simulate_lineprofile
 *                         (rho_diffusion/data/spectroscopy.py:142-189) divided by its maximum (:120-121): sum over the transitions
 *                         k in [offsets[b], offsets[b + 1]) whose centre lies inside [grid[0], grid[G - 1]] of
 *                         intensities[k] exp(-(grid[f] - centers[k])^2 / (2 widths[b]^2)), with numpy's promotion (fp32 difference and
 *                         square, fp64 division, exp and sum).  workspace: rho_lineprofile_workspace_bytes(B, G). */
size_t rho_lineprofile_workspace_bytes(int32_t B, int32_t G);
int rho_lineprofile(float* out, int32_t B, int32_t G, const float* grid, const float* centers, const float* intensities, const int32_t* offsets,
                    const double* widths, void* workspace, size_t workspace_bytes, rho_stream_t stream);
size_t rho_sph_harm_workspace_bytes(int32_t N, int32_t G);
int rho_sph_harm_density(float* out, int32_t N, int32_t G, const int32_t* l, const int32_t* m, double lo, double hi,
                         const double* minmax /* optional [N][4]: re, im of the complex min and max to normalise with */, void* workspace,
                         size_t workspace_bytes, rho_stream_t stream);
int rho_gaussian_blobs(float* out, int32_t N, int32_t D, int32_t H, int32_t W, const int32_t* n_blobs, const float* width, uint64_t seed,
                       uint64_t first_index, rho_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * the two remaining registered backbones: v1 UNet (rho_diffusion/models/unet.py:30-269) and the diffusion
 * VisionTransformer (rho_diffusion/models/vit.py:31-372).  Their contractions run on rho_conv_* / rho_attention_*; these are the
 * element-wise / row-wise pieces around them:
 *   rho_activation[_bwd]  y = act(x) / dx = dy act'(x): `self.activation(self.conv1(x))` (unet.py:122-123), Linear -> activation
 *                         (vit.py:155-161, 166-170, 311-315); 16-byte aligned pointers.
 *   rho_convt_weight      dst[b][a][taps-1-t] = src[a][b][t] (fp32): nn.ConvTranspose{2,3}d(k, stride 1, padding k//2) of the v1 up
 *                         path (unet.py:69-75, 94-100) IS conv_nd with the [Cin, Cout, *k] weight transposed and flipped; the map is
 *                         its own adjoint with A and B swapped (weight gradient); taps = 1 is a plain transpose.
 *   rho_patchify          inverse = 0: x [N, C, D, H, W] -> tokens [N, T, C pd ph pw], tokens in row-major order of the patch grid,
 *                         features in the order of weight.view(E, -1): the kernel = stride = patch conv of PatchEmbedding
 *                         (vit.py:72-77, 119-126) becomes a Linear over tokens; inverse = 1: tokens -> field (+ bias[c]): the
 *                         kernel = stride ConvTranspose `output_conv` (vit.py:297-303, 358-371) after its Linear.  Each direction
 *                         is the other's adjoint (without the bias).
 *   rho_bias_expand       dst[c reps + r] = src[c]; inverse: dst[c] = sum_r src[c reps + r] (that ConvTranspose's bias as a Linear bias)
 *   rho_add_bcast         y[o][i] = x[o][i] + add[i] (positional embedding added to every sample, vit.py:349-352); inner % 4 == 0
 *   rho_colsum            out[g][c] = sum_{r < rows} src[(g rows + r) ld + c], fixed order (gradients of the broadcast operands)
 *   rho_layernorm_*       nn.LayerNorm(C) over tokens [rows, C] (vit.py:140-141), optionally of x + pre_add[row / rows_per_sample]
 *                         (the diffusion-time embedding added before norm_1, vit.py:179-181); stats [rows, 2] = (mean, rstd);
 *                         backward: dx, dgamma / dbeta (deterministic two-stage sums over a workspace), dpre_add [samples, C]. */
typedef struct {
  int64_t rows; int32_t C; int32_t rows_per_sample;
  int32_t in_dtype, out_dtype; float eps; int32_t _pad;
  const void* x; const float* pre_add /* optional [rows / rows_per_sample, C] */; const float* gamma; const float* beta;
  void* y; float* stats;
} rho_layernorm_params;
int rho_activation(const void* x, void* y, int dtype, int64_t n, int act, rho_stream_t stream);
int rho_activation_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t n, int act, rho_stream_t stream);
int rho_convt_weight(const float* src, float* dst, int32_t A, int32_t B, int32_t taps, rho_stream_t stream);
int rho_patchify(const void* src, int src_dtype, void* dst, int dst_dtype, const float* bias, int32_t N, int32_t C, int32_t D, int32_t H,
                 int32_t W, int32_t pd, int32_t ph, int32_t pw, int32_t inverse, rho_stream_t stream);
int rho_bias_expand(const float* src, float* dst, int32_t C, int32_t reps, int32_t inverse, rho_stream_t stream);
int rho_add_bcast(const void* x, int dtype, const float* add, void* y, int64_t outer, int64_t inner, rho_stream_t stream);
int rho_colsum(const void* src, int dtype, int64_t ld, int32_t groups, int32_t rows, int64_t cols, float* out, rho_stream_t stream);
int rho_layernorm_fwd(const rho_layernorm_params* p, rho_stream_t stream);
size_t rho_layernorm_bwd_workspace_bytes(const rho_layernorm_params* p);
int rho_layernorm_bwd(const rho_layernorm_params* p, const void* dy, void* dx, float* dgamma, float* dbeta, float* dpre_add, void* workspace,
                      size_t workspace_bytes, rho_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RHO_B200_H_ */
