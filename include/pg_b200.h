/* pg_b200.h — C ABI of libpg_b200.so, the sm_100a kernel library behind the drop-in
 * `pytorch_generative.nn.*` / `models.*` Module API.
 *
 * The reference (EugenHotaj/pytorch-generative) ships no native interface: its hot path is Python
 * nn.Modules whose arithmetic is delegated to torch.  The entry points below are
 * therefore what a maintainer's ctypes binding would call in place of those torch ops; each one
 * cites the reference call site it replaces.  Conventions:
 *   - every pointer is a raw device pointer unless stated otherwise; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value 0 = success, non-zero = failure with a message available from pg_last_error();
 *   - no allocation, no global stream state, no torch/pybind types;
 *   - activations are "pixel-major": a [P, C] row-major matrix with P = N*H*W pixels (NHWC), which is
 *     the layout every GEMM-shaped op wants; NCHW<->pixel-major converters are provided for the
 *     module boundary.
 */
#ifndef PG_B200_H_
#define PG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_ABI_VERSION 1

int pg_abi_version(void);
const char* pg_last_error(void);
/* Number of SMs of the current device (148 on B200); used by callers to size split-K. */
int pg_sm_count(void);
/* Leaves `n` SMs out of every persistent grid from now on (returns the previous setting; 0 = use them all): room for the
 * NCCL kernels of the gradient all-reduce that run concurrently with the backward pass under data parallelism. */
int pg_reserve_sms(int n);
/* Number of kernels this library has launched in this process (all threads); bench.py's gpu_launches. */
unsigned long long pg_launch_count(void);

/* Activation ids (shared by the GEMM epilogue and the elementwise kernels). */
enum { PG_ACT_NONE = 0, PG_ACT_RELU = 1, PG_ACT_GELU = 2, PG_ACT_ELU = 3, PG_ACT_TANH = 4,
       /* pg_gemm_epilogue.dact only: `aux` already holds the derivative (see PG_ACT_STORE_DERIV) */
       PG_ACT_GIVEN = 5,
       /* pg_gemm_epilogue.dact only: `aux` holds the ACTIVATED value (relu(pre) / elu(pre)), from which the derivative
        * follows without the pre-activation: relu' = [a > 0], elu' = a > 0 ? 1 : a + 1 */
       PG_ACT_RELU_OUT = 6, PG_ACT_ELU_OUT = 7 };
/* OR-ed into pg_gemm_epilogue.act: out_pre receives act'(pre) instead of pre, so that the matching dgrad epilogue
 * (dact = PG_ACT_GIVEN) is a single multiply. */
#define PG_ACT_STORE_DERIV 0x100
/* OR-ed into pg_gemm_epilogue.act: res0 / res1 point to bf16 [M,N] matrices (pitch ld_res in bf16 elements, a multiple of 8)
 * instead of fp32 ones — short-lived sums (GatedPixelCNN's vertical-to-horizontal link) that are not a residual stream. */
#define PG_ACT_RES_BF16 0x200

/* ---------------------------------------------------------------------------------------------
 * Channel contraction (every nn.Conv2d 1x1 on the path and, per live tap, every masked conv):
 *   reference: torch.nn.Conv2d.forward at nn/attention.py:105-118,140-144,161;
 *   models/autoregressive/image_gpt.py:40-48,101-103; pixel_cnn.py:35-49,95-103;
 *   gated_pixel_cnn.py:79-99,176-182; pixel_snail.py:86-87,173-180 — and their autograd
 *   (dgrad / wgrad).
 *
 *   acc[m,n] = sum_k A(m,k) * B(n,k)            bf16 inputs, fp32 accumulation on tcgen05
 *   t        = alpha * acc + bias[n]
 *   t       *= act'(aux[m,n])                   if dact != PG_ACT_NONE   (backward through an activation)
 *   pre      = t + res0[m,n] + res1[m,n]
 *   out_f32[m,n]  = pre  (or += pre when accumulate=1; bias/res only added by split 0)
 *   out_pre[m,n]  = bf16(pre)
 *   out_bf16[m,n] = bf16(act(pre))
 *
 * Operand layouts: a_mn_major=0 -> A is [M,K] row-major with pitch lda (K contiguous);
 *                  a_mn_major=1 -> A is [K,M] row-major with pitch lda (M contiguous).
 *                  b_mn_major=0 -> B is [N,K] row-major (a conv weight [Cout,Cin]);
 *                  b_mn_major=1 -> B is [K,N] row-major.
 * So forward = (0,0) with B = W; dgrad = (0,1) with B = W; wgrad = (1,1) with A = dY, B = X.
 * Pitches must be multiples of 8 elements and bases 16-byte aligned (TMA requirement).
 * ------------------------------------------------------------------------------------------- */
typedef struct pg_gemm_epilogue {
  const float* bias;   /* [N] or NULL */
  const void* aux;     /* bf16 [M,N] pre-activation (or the derivative itself, dact = PG_ACT_GIVEN), used when dact != 0 */
  const float* res0;   /* fp32 [M,N] or NULL */
  const float* res1;   /* fp32 [M,N] or NULL */
  void* out_bf16;      /* bf16 [M,N] or NULL */
  void* out_pre;       /* bf16 [M,N] or NULL: pre-activation (act'(pre) with PG_ACT_STORE_DERIV) */
  float* out_f32;      /* fp32 [M,N] or NULL */
  int64_t ld_aux, ld_res, ld_out_bf16, ld_out_pre, ld_out_f32; /* row pitches, elements */
  int32_t act;         /* activation applied to out_bf16 */
  int32_t dact;        /* activation whose derivative (at aux) scales the accumulator */
  int32_t accumulate;  /* 1: out_f32 is accumulated with fp32 atomics (split-K / grad accumulation) */
  float alpha;
  float* bias_grad;    /* weight-gradient GEMMs only (a_mn_major = 1, impl 0), or NULL: fp32 [M] += sum_k A(m,k), i.e. the
                        * bias gradient sum_p dY[p, cout] of the same layer, reduced from the staged A tiles */
} pg_gemm_epilogue;

/* impl: 0 = tcgen05/TMA kernel (the product); 1 = plain SIMT kernel kept as an on-device cross-check
 * for the tests (same epilogue code); 2 = skinny-rows kernel for M <= 32 (the per-pixel step of incremental
 * sampling: one warp per output column, weights streamed once).  split_k >= 1 (only with accumulate=1 and no
 * activation). */
int pg_gemm_bf16(const void* A, int a_mn_major, int64_t lda, const void* B, int b_mn_major, int64_t ldb,
                 int M, int N, int K, int split_k, const pg_gemm_epilogue* epi, int impl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tap-loop convolution, im2col-free (CausalConv2d with wide channels — reference nn/convolution.py:41-43;
 * GatedPixelCNN 1xN / Nx1 stacks — gated_pixel_cnn.py:63-99 with the crops at 115,121; PixelSNAIL 2x2 —
 * pixel_snail.py:41-56), forward and both gradients on the pg_gemm_bf16 kernel:
 *   conv(x)[p] = sum_t W_t . x[p + (dy_t, dx_t)], zero outside the image (the reference's pad + front crop).
 * The shifted operand is never materialised: for every tap the TMA unit loads the [C, W, H, N] box displaced by
 * (dy_t, dx_t) from the pixel-major tensor (out-of-image elements arrive as zeros) and the tensor core accumulates
 * one K slab per (tap, 64 channels).  Epilogue semantics are pg_gemm_bf16's.
 *   PG_CONV_FWD:   A = x  [P, >=C] bf16, C = Cin;  B = packed weight [Cout, T*Cin] (tap-major columns);
 *                  M = P, N = Cout, K = T*Cin.
 *   PG_CONV_DGRAD: A = dy [P, >=C] bf16, C = Cout, taps negated by the caller;  B = the same packed weight;
 *                  M = P, N = Cin (= the per-tap column stride of the packed weight), K = T*Cout.
 *   PG_CONV_WGRAD: A = dy [P, Cout] bf16;  B = x [P, >=C] bf16, C = Cin;  M = Cout, N = T*Cin, K = P;
 *                  out_f32 [Cout, T*Cin] accumulated (split_k as pg_gemm_bf16).
 * Geometry limits (else use pg_tap_gather): C % 64 == 0, W | 64, H*W % 128 == 0, <= 32 taps, |offset| <= 64.
 * ------------------------------------------------------------------------------------------- */
enum { PG_CONV_FWD = 1, PG_CONV_DGRAD = 2, PG_CONV_WGRAD = 3 };
typedef struct pg_conv_geom {
  int32_t mode;        /* PG_CONV_* */
  int32_t N, H, W;     /* images, rows, columns: P = N*H*W pixels */
  int32_t C;           /* channels of the shifted tensor (per-tap K extent for fwd / dgrad, per-tap N extent for wgrad) */
  int32_t n_taps;
  int32_t dy[32], dx[32];
} pg_conv_geom;
int pg_gemm_bf16_conv(const void* A, int64_t lda, const void* B, int64_t ldb, int M, int N, int K, int split_k,
                      const pg_gemm_epilogue* epi, const pg_conv_geom* geom, void* stream);

/* Column sums of a bf16 [P, C] matrix into fp32 out[C] (bias gradients; accumulate=1 adds). */
int pg_colsum_bf16(const void* x, int64_t ld, int P, int C, float* out, int accumulate, void* stream);
int pg_colsum_f32(const float* x, int64_t ld, int P, int C, float* out, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * NCHWLayerNorm — reference nn/convolution.py:69-75 (permute -> nn.LayerNorm(C) -> permute).
 * Pixel-major x [P, C] fp32 (the residual stream) -> y bf16 and/or fp32; eps as nn.LayerNorm (1e-5),
 * biased variance; mean/rstd [P] fp32 are saved for backward.
 * Backward: dx = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat)), g = dy * gamma;
 *   dx_out_f32 = dx + dres0 + dres1 (fused residual-gradient adds), optional bf16 copy for the next
 *   dgrad GEMM; dgamma/dbeta are accumulated (atomics) into fp32 [C] buffers that the caller zeroed;
 *   dx_colsum (optional, [C], caller-zeroed) receives the column sums of dx_out = the bias gradient of the
 *   layer whose output x is (saves a separate pass over the gradient).
 * ------------------------------------------------------------------------------------------- */
int pg_layernorm_fwd(const float* x, const float* gamma, const float* beta, int P, int C, float eps,
                     void* y_bf16, float* y_f32, float* mean, float* rstd, void* stream);
int pg_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                     const float* mean, const float* rstd, int P, int C, const float* dres0,
                     const float* dres1, float* dx_f32, void* dx_bf16, float* dgamma, float* dbeta,
                     float* dx_colsum, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GatedActivation — reference nn/convolution.py:46-66: act(x[:, :C]) * sigmoid(x[:, C:]).
 * Pixel-major x [P, 2C] (bf16 or fp32) -> y [P, C].  act is PG_ACT_TANH (GatedPixelCNN) or
 * PG_ACT_NONE (PixelSNAIL's nn.Identity).  Backward writes dx [P, 2C].
 * ------------------------------------------------------------------------------------------- */
int pg_gated_act_fwd(const void* x, int x_is_f32, int P, int C, int act, void* y, int y_is_f32, void* stream);
int pg_gated_act_bwd(const void* x, int x_is_f32, const void* dy, int dy_is_f32, int P, int C, int act,
                     void* dx, int dx_is_f32, void* stream);
/* y = res + act(x[:, :C]) * sigmoid(x[:, C:]) in one pass: the residual sum around PixelSNAIL's gated residual block
 * (reference pixel_snail.py:52-56), res / y fp32 [P, C]. */
int pg_gated_res_fwd(const void* x, int x_is_f32, const float* res, int P, int C, int act, float* y, void* stream);
/* out = bf16(dy * act'(pre)) given the ACTIVATED value ya = act(pre) (relu / elu; elu'(pre) = ya + 1 for pre <= 0): the
 * backward of an activation whose output, not input, was kept (elu(conv(.)) outputs, pixel_snail.py:27-28,115-119). */
int pg_dact_from_out(const void* dy, int dy_is_f32, const void* ya_bf16, int64_t numel, int act, void* out_bf16, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Recipe loss — reference models/autoregressive/image_gpt.py:158-162 (identical in the other three):
 * BCEWithLogits(preds, x, reduction="none").sum(1).mean().  logits/target are [N, D] fp32 in any
 * common memory order (elementwise); loss_sum receives sum over all elements (caller divides by N);
 * dlogits = (sigmoid(l) - t) * scale.
 * ------------------------------------------------------------------------------------------- */
int pg_bce_logits_fwd_bwd(const float* logits, const float* target, int64_t numel, float grad_scale,
                          float* loss_sum /* 1 float, accumulated */, float* dlogits /* or NULL */,
                          void* stream);

/* Layout converters for the Module boundary (NCHW fp32 <-> pixel-major). */
int pg_nchw_to_pm(const float* x_nchw, int N, int C, int HW, void* out, int out_is_f32, int64_t ld_out,
                  void* stream);
/* act (PG_ACT_*) is applied on the way out: used for activations that follow a convolution. */
int pg_pm_to_nchw(const void* x_pm, int x_is_f32, int64_t ld_x, int N, int C, int HW, int act, float* out_nchw,
                  void* stream);
/* out = bf16(dy * act'(pre)): gradient through such an output activation. */
int pg_dact_mul(const void* dy_bf16, int64_t ld_dy, const float* pre_f32, int64_t ld_pre, int P, int C, int act,
                void* out_bf16, int64_t ld_out, void* stream);
/* out = bf16(act(x)) over a pitched pixel-major [P, C] matrix (fp32 or bf16 in): builds the tensor-core operand of a
 * convolution whose input activation (ReLU / ELU in front of the conv: pixel_cnn.py:35-49, pixel_snail.py:27-28) was not
 * already emitted by the producing GEMM's epilogue.  C % 8 == 0. */
int pg_act_cast_bf16(const void* x, int x_is_f32, int64_t ld_x, int P, int C, int act, void* out_bf16, int64_t ld_out,
                     void* stream);
/* fp32 -> bf16 cast of a dense buffer (weights packing; masked taps already zeroed by the caller). */
int pg_cast_f32_to_bf16(const float* x, void* y, int64_t numel, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CausalAttention core — reference nn/attention.py:147-160 (mask, q@k^T/sqrt(dk), masked softmax,
 * re-zero, attn@v, head concat).  q/k/v/o are pixel-major bf16 with per-image sequences of length S
 * (seq index = row*W+col), heads are contiguous channel blocks of dk (q,k) / dv (v,o) channels.
 * strict=1 is mask_center=True (position i attends j<i; row 0 yields zeros), strict=0 attends j<=i.
 * `scale` multiplies q.k (the reference uses 1/sqrt(embed_channels/n_heads); it is passed explicitly so
 * that head slots may be zero-padded: the tcgen05 kernels require dk == 64 and dv in {64, 128}, narrower
 * heads are laid out in 64-wide slots whose extra columns are zero).
 * lse [N, H, S] fp32 (log-sum-exp of scaled scores; rows without keys store 0 and o = 0).
 * impl: 0 = tcgen05 kernel, 1 = SIMT cross-check (any dk, dv <= 128, S <= 1024); backward only: 2 = experimental
 * split-phase tcgen05 kernel (dv slot 64), opt-in, never selected by the product.
 * ------------------------------------------------------------------------------------------- */
int pg_causal_attn_fwd(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v, int64_t ld_v,
                       void* o, int64_t ld_o, float* lse, int N, int S, int H, int dk, int dv, float scale,
                       int strict, int impl, void* stream);
/* Scratch: delta [N, H, S] fp32; dq_accum [N*S, H*dk] fp32, contiguous, contents ignored on entry (the tcgen05
 * kernels accumulate dQ across key tiles with fp32 bulk reduce-adds, then round it to bf16; the library clears the buffer
 * itself, inside the delta pass; NULL for impl 1).  dq/dk/dv are bf16 pixel-major. */
int pg_causal_attn_bwd(const void* q, int64_t ld_q, const void* k, int64_t ld_k, const void* v, int64_t ld_v,
                       const void* o, int64_t ld_o, const void* d_o, int64_t ld_do, const float* lse,
                       float* delta, float* dq_accum, void* dq, int64_t ld_dq, void* dk_, int64_t ld_dk,
                       void* dv_, int64_t ld_dv, int N, int S, int H, int dk, int dv, float scale, int strict,
                       int impl, void* stream);

/* Incremental (KV-cached) attention for AutoregressiveModel.sample (reference models/base.py:97-120 recomputes the full
 * forward per pixel; causality makes the per-position update exact): appends the current position's k / v rows
 * ([N, H*d]) to the caches ([N*S, H*d]) at row *pos_dev and attends over cache rows [0, pos] ([0, pos) if strict).
 * pos_dev is a device int so that one captured CUDA graph serves the whole raster scan. */
int pg_attn_decode(const void* q, int64_t ld_q, const void* k_new, int64_t ld_kn, const void* v_new, int64_t ld_vn,
                   void* k_cache, int64_t ld_kc, void* v_cache, int64_t ld_vc, void* o, int64_t ld_o,
                   const int* pos_dev, int N, int S, int H, int dk, int dv, float scale, int strict, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tap-list convolution for small channel counts (CausalConv2d input layers, Cin in {1,3}):
 * reference nn/convolution.py:41-43 (weight.data *= mask; F.conv2d).  x is NCHW fp32 (the model
 * input), w is the masked OIHW fp32 weight, output pixel-major.  taps are all kh*kw positions; masked
 * taps contribute zero because the caller zeroes the weight in place exactly as the reference does.
 * wgrad is dense over kh*kw (masked taps receive gradient, as autograd does in the reference).
 * ------------------------------------------------------------------------------------------- */
int pg_conv_small_fwd(const float* x_nchw, const float* w_oihw, const float* bias, int N, int Cin, int H, int W,
                      int Cout, int kh, int kw, int pad_h, int pad_w, int pre_act /* applied to x */, float* out_f32,
                      void* out_bf16, int act_bf16, void* stream);
int pg_conv_small_bwd(const float* x_nchw, const float* w_oihw, const float* dy_pm /* [P,Cout] fp32 */, int N,
                      int Cin, int H, int W, int Cout, int kh, int kw, int pad_h, int pad_w, int pre_act,
                      float* dw_oihw /* accumulated */, float* dbias /* accumulated */,
                      float* dx_nchw /* or NULL; overwritten */, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Wide-channel tap-list convolutions (CausalConv2d with Cin >= 8; GatedPixelCNN 1xN / Nx1 — reference
 * gated_pixel_cnn.py:63-99 with the crops at 115,121; PixelSNAIL 2x2 — pixel_snail.py:41-56).
 * conv(x)[p] = sum_t W_t . x[p + (dy_t, dx_t)], zero outside the image (= the reference's pad + front crop).
 * pg_tap_gather builds X_cat[p, t*C + c] = act(x[p + off_t, c]) in bf16; the contraction over K = T*C is
 * pg_gemm_bf16 (forward, dgrad to dX_cat, wgrad from X_cat); pg_tap_scatter folds dX_cat back:
 * dx[p, c] = act'(x_pre[p, c]) * sum_t dX_cat[p - off_t, t*C + c].  C % 8 == 0, T <= 32.
 * ------------------------------------------------------------------------------------------- */
int pg_tap_gather(const void* x_pm, int64_t ld_x, int N, int H, int W, int C, int T, const int* dy /* host */,
                  const int* dx /* host */, int act, void* out /* bf16 [P, T*C] */, void* stream);
int pg_tap_scatter(const void* dxcat /* bf16 [P, T*C] */, int N, int H, int W, int C, int T, const int* dy,
                   const int* dx, int act, const void* x_pre /* bf16 [P, ld_pre] or NULL */, int64_t ld_pre,
                   float* dx_f32, void* dx_bf16, int64_t ld_dx, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LinearCausalAttention numerator — reference nn/attention.py:168-200 (`_UnnormalizedLinearCausalAttention`: a Python loop
 * over the sequence, forward and backward).  q, k: [B, L, d] fp32, v / g / out: [B, L, dv] fp32, B = images x heads,
 * contiguous.  out_i = q_i . S_i,  S_i = sum_{j <= i} k_j^T v_j.  Backward: dq_i = g_i S_i^T; with R_i = sum_{j >= i}
 * q_j^T g_j: dv_i = k_i R_i, dk_i = v_i R_i^T.  One CTA per (image, head), state in registers, O(L (d + dv)) memory.
 * d <= 64, dv <= 128.
 * ------------------------------------------------------------------------------------------- */
int pg_linear_attn_fwd(const float* q, const float* k, const float* v, float* out, int B, int L, int d, int dv, void* stream);
int pg_linear_attn_bwd(const float* q, const float* k, const float* v, const float* g, float* dq, float* dk, float* dv_out,
                       int B, int L, int d, int dv, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer part of the training step — reference trainer.py:182-191 (`clip_grad_norm_(model.parameters(), max_norm)`
 * then `optimizer.step()` with torch.optim.Adam as every recipe builds it, e.g. image_gpt.py:155) over ALL parameters in
 * two launches.  Tensors are given as device arrays of device pointers (one entry per parameter, fp32, contiguous);
 * `numel` [n_tensors] int64 (device); `chunks` [n_chunks] pairs of int32 (tensor index, chunk index) (device): block b
 * handles elements [chunk * chunk_elems, +chunk_elems) of its tensor.
 *   pg_grad_sqnorm: partials[b] = sum of g^2 over block b's chunk.
 *   pg_adam_step:   norm = sqrt(sum partials) (same order in every block: deterministic); norm_out[0] = norm;
 *                   if skip_above > 0 and norm > skip_above: nothing is updated and norm_out[1] = 0 (the trainer's
 *                   skip_grad_norm rule), else norm_out[1] = 1 and, with c = min(1, max_norm / (norm + 1e-6)):
 *                   g *= c (written back only when c < 1), m = b1 m + (1-b1) g, v = b2 v + (1-b2) g^2,
 *                   p -= lr / (1-b1^step) * m / (sqrt(v) / sqrt(1-b2^step) + eps)      (torch.optim.Adam, no amsgrad).
 * ------------------------------------------------------------------------------------------- */
/* dst[t][i] = bf16(src[t][i]) for many fp32 tensors in one launch (same pointer-array / chunk-table convention): the
 * per-step refresh of the bf16 tensor-core copies of the fp32 master weights. */
int pg_cast_multi_bf16(const void* src_ptrs, const void* dst_ptrs, const int64_t* numel, const void* chunks, int n_chunks,
                       int chunk_elems, void* stream);
int pg_grad_sqnorm(const void* grad_ptrs, const int64_t* numel, const void* chunks, int n_chunks, int chunk_elems,
                   float* partials, void* stream);
int pg_adam_step(const void* param_ptrs, const void* grad_ptrs, const void* exp_avg_ptrs, const void* exp_avg_sq_ptrs,
                 const int64_t* numel, const void* chunks, int n_chunks, int chunk_elems, const float* partials,
                 float max_norm, float skip_above, double lr, double beta1, double beta2, double eps, int step,
                 float* norm_out /* [2] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PG_B200_H_ */
