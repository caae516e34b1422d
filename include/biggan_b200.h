/* biggan_b200.h -- C ABI of libbiggan_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for the BigGAN G+D training step of david-jk/BigGAN-Tensorflow.  The reference
 * has no FFI: its operator surface is the Python module ops.py / DiffAugment_tf.py, whose math is
 * executed by stock TensorFlow kernels.  Each entry point below replaces the TF kernels behind one
 * reference function (cited as file:line into /root/reference); biggan_tensorflow_b200/ops.py is the
 * Python mirror that keeps the reference signatures and calls these through ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (bg_status); bg_last_error() gives the text
 *  - all pointers are DEVICE pointers owned by the caller; the library allocates nothing, keeps no
 *    pointer after return, and only enqueues work on `stream` (a cudaStream_t passed as void*)
 *  - activations are NHWC; `dtype` selects their storage: BG_F32 (parity mode) or BG_BF16
 *    (performance mode, fp32 accumulate).  Master weights / gradients of weights are always fp32.
 *  - workspaces are caller-allocated; query the *_workspace_bytes function first
 */
#ifndef BIGGAN_B200_H
#define BIGGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { BG_OK = 0, BG_ERR_SHAPE = -1, BG_ERR_UNSUPPORTED = -2, BG_ERR_CUDA = -3,
               BG_ERR_WORKSPACE = -4 } bg_status;
typedef enum { BG_F32 = 0, BG_BF16 = 1 } bg_dtype;
typedef enum { BG_ACT_NONE = 0, BG_ACT_RELU = 1, BG_ACT_TANH = 2 } bg_act;
typedef enum { BG_PAD_ZERO = 0, BG_PAD_REFLECT = 1 } bg_pad_mode;

const char* bg_last_error(void);
int bg_abi_version(void);
/* 1 if the tcgen05/TMA tensor-core path is usable on the current device (sm_100), else 0 */
int bg_has_tensor_core_path(void);
/* number of conv launches served by the tcgen05 forward-type / wgrad kernels so far (tests, bench) */
int bg_debug_counters(int* tc_fwd, int* tc_wgrad);

/* ---------------------------------------------------------------------------------------------
 * Spectral norm -- ops.py:718-747 (spectral_norm).  W is the [K, C] view `reshape(w, [-1, C])`.
 * bg_sn_power_iter: v_hat = l2n(u W^T), u_hat = l2n(v_hat W), sigma = v_hat W u_hat^T; writes
 *   v_hat -> v[K], u_hat -> u_out[C] (may alias u), sigma[0].   One HBM read of W (+ L2 re-read).
 * bg_sn_scale: out = W / sigma in `dtype`; if out_t != NULL also writes the per-tap transposed copy
 *   out_t[t][c][r] for W viewed as [taps][rows][C] (K = taps*rows) -- the K-major operand layout of
 *   the tcgen05 conv kernels.  c_pad >= C: output channel count of the (pre-zeroed) destination
 *   buffers, r_pad >= rows: input channel count per tap there -- out is [taps][r_pad][c_pad], out_t is
 *   [taps][c_pad][r_pad] (zero-padded channels make the narrow layers -- C/8 attention projections, the RGB
 *   convs -- tensor-core eligible).
 * bg_sn_bwd: dW = dWn/sigma - (<dWn,W>/sigma^2) v_hat^T u_hat  (gradient with u_hat, v_hat frozen,
 *   ops.py:738-739).  dWn is laid out [K/rows][r_pad][g_ld] (the padded layout of the forward).
 * ------------------------------------------------------------------------------------------- */
size_t bg_sn_workspace_bytes(int K, int C);
int bg_sn_power_iter(const float* W, const float* u, float* u_out, float* v, float* sigma,
                     int K, int C, void* ws, size_t ws_bytes, void* stream);
int bg_sn_scale(const float* W, const float* sigma, void* out, void* out_t, int dtype,
                int taps, int rows, int C, int r_pad, int c_pad, void* stream);
int bg_sn_bwd(const float* W, const float* dWn, int rows, int r_pad, int g_ld, const float* u_hat,
              const float* v_hat, const float* sigma, float* dW, int K, int C, void* ws, size_t ws_bytes,
              void* stream);

/* Grouped spectral norm: every weight matrix of a network in 4 launches (forward) / 2 launches (backward)
 * driven by a device-resident plan built once.  All pointers in bg_sn_item are device pointers that must stay
 * valid (persistent buffers): W master weights, u (updated in place), v (K floats), sigma (1 float), out / out_t
 * normalised weights (layouts of bg_sn_scale; out_t may be NULL), dWn gradient w.r.t. the normalised weights
 * ([taps][r_pad][g_ld] fp32), dW destination of the weight gradient ([K, C] fp32). */
typedef struct {
  const float* W; float* u; float* v; float* sigma; void* out; void* out_t; const float* dWn; float* dW;
  int K, C, taps, rows, r_pad, c_pad, dtype, g_ld;
} bg_sn_item;
size_t bg_sn_group_plan_bytes(const bg_sn_item* items, int n);
/* host-side planning + synchronous upload into plan_dev; totals[8] (host ints) is passed to the launches */
int bg_sn_group_plan(const bg_sn_item* items, int n, void* plan_dev, size_t plan_bytes, int* totals);
int bg_sn_group_forward(const void* plan_dev, const int* totals, void* stream);
int bg_sn_group_backward(const void* plan_dev, const int* totals, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolution family -- ops.py:49-113 (conv), ops.py:116-139 (deconv), ops.py:148-175
 * (fully_connected == 1x1 conv on a 1x1 map).
 *   transposed == 0: y = conv2d(pad(x), w[kh,kw,Ci,Co], stride), pad = explicit (pad_t,pad_l,pad_b,
 *                    pad_r) with REFLECT (tf.pad, ops.py:82) or ZERO fill; w is HWIO.
 *   transposed == 1: y = tf.nn.conv2d_transpose(x, w[kh,kw,Co,Ci], stride, 'SAME'), output
 *                    Ho = Hi*stride; pad_t/pad_l = pad_before of the adjoint SAME conv.
 * (Hi,Wi,Ci) always describe x and (Ho,Wo,Co) describe y, for fwd, dgrad and wgrad alike.
 * w   : normalised weights in the TF-native layout above, fp32 (BG_F32) or bf16 (BG_BF16)
 * w_t : per-tap transposed copy (see bg_sn_scale), may be NULL (then only the SIMT path is used)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int B, Hi, Wi, Ci;
  int Ho, Wo, Co;
  int kh, kw, stride;
  int pad_t, pad_l, pad_b, pad_r;
  int pad_mode;      /* bg_pad_mode */
  int transposed;    /* 0 conv, 1 conv2d_transpose */
} bg_conv_desc;

/* y = act(conv(x) + bias + residual);  bias [Co] fp32 or NULL; residual same shape as y or NULL
 * (residual may alias y).  ws: scratch for the reflect-padded bf16 input of the tensor-core path.
 * force_simt != 0 disables the tensor-core path (testing). */
size_t bg_conv2d_fwd_workspace_bytes(const bg_conv_desc* d, int dtype);
int bg_conv2d_fwd(const bg_conv_desc* d, const void* x, const void* w, const void* w_t,
                  const float* bias, const void* residual, void* y, int act, int dtype,
                  void* ws, size_t ws_bytes, int force_simt, void* stream);
/* dx = conv_dgrad(dy) (+ residual).  For pad_mode REFLECT with padding the caller gets the gradient
 * w.r.t. the *unpadded* x (the reflect-pad adjoint is folded inside; ws holds the padded gradient).
 * relu_mask (may be NULL; reflect-padded convs, or UNPADDED plain convs on the bf16 tensor-core path -- anything else returns
 * BG_ERR_UNSUPPORTED; same shape and dtype as dx): dx = residual + (relu_mask > 0 ?
 * conv_dgrad(dy) : 0) -- the backward of a ReLU sitting on the conv's input (ops.py:297-301: act -> conv in the
 * residual blocks) is applied where the gradient is produced instead of in a pass of its own. */
size_t bg_conv2d_dgrad_workspace_bytes(const bg_conv_desc* d, int dtype);
int bg_conv2d_dgrad(const bg_conv_desc* d, const void* dy, const void* w, const void* w_t,
                    const void* residual, const void* relu_mask, void* dx, int dtype, void* ws,
                    size_t ws_bytes, int force_simt, void* stream);
/* dw (fp32, TF-native layout of w) = wgrad(x, dy); deterministic split-K through ws. */
size_t bg_conv2d_wgrad_workspace_bytes(const bg_conv_desc* d, int dtype);
int bg_conv2d_wgrad(const bg_conv_desc* d, const void* x, const void* dy, float* dw, int dtype,
                    void* ws, size_t ws_bytes, int force_simt, void* stream);
/* db[C] = sum over rows of dy[rows, C] (bias gradient; tf.nn.bias_add adjoint) */
size_t bg_colsum_workspace_bytes(long rows, int C);
int bg_colsum(const void* dy, float* db, long rows, int C, int dtype, void* ws, size_t ws_bytes,
              void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batch norm family -- ops.py:611-643 (condition_batch_norm), ops.py:580-585 (batch_norm).
 * x [B, HW, C].  Per-sample partial statistics so that the same kernels serve the local batch and,
 * after an all-reduce of (sum, sumsq) across ranks, the cross-replica (global batch) statistics.
 *   bg_bn_stats   : sums[0:C] = sum x, sums[C:2C] = sum x^2 over (B,HW)        (double precision)
 *   bg_bn_finalize: mean, var(biased) from sums / count; EMA update of the running stats
 *                   run_mean = run_mean*momentum + mean*(1-momentum); run_var likewise with
 *                   var*var_correction (1 for ops.py:631-632, n/(n-1) for tf.layers fused BN)
 *   bg_bn_apply   : y = act((x-mean)*rsqrt(var+eps)*gamma[b] + beta[b]); gamma/beta [B,C] when
 *                   gb_bstride == C (conditional) or [C] when gb_bstride == 0
 *   bg_bn_bwd_reduce: dgamma[b,c] = sum_hw g*xhat, dbeta[b,c] = sum_hw g, g = dy*act'(y)
 *   bg_bn_bwd_sums  : s[0:C] = sum_b gamma*dbeta, s[C:2C] = sum_b gamma*dgamma   (double; all-reduce
 *                     these across ranks for sync-BN)
 *   bg_bn_bwd_apply : dx = rstd*(gamma*g - s1/count - xhat*s2/count)
 * ------------------------------------------------------------------------------------------- */
size_t bg_bn_workspace_bytes(int B, long HW, int C);
int bg_bn_stats(const void* x, double* sums, int B, long HW, int C, int dtype, void* ws,
                size_t ws_bytes, void* stream);
int bg_bn_finalize(const double* sums, double count, float* mean, float* var, float* run_mean,
                   float* run_var, float momentum, float var_correction, int C, void* stream);
int bg_bn_apply(const void* x, const float* mean, const float* var, const float* gamma,
                const float* beta, int gb_bstride, void* y, int act, float eps, int B, long HW,
                int C, int dtype, void* stream);
int bg_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* var,
                     const float* gamma, const float* beta, int gb_bstride, float* dgamma,
                     float* dbeta, int act, float eps, int B, long HW, int C, int dtype, void* ws,
                     size_t ws_bytes, void* stream);
int bg_bn_bwd_sums(const float* dgamma, const float* dbeta, const float* gamma, int gb_bstride,
                   double* s, int B, int C, void* stream);
int bg_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* var,
                    const float* gamma, const float* beta, int gb_bstride, const double* s,
                    double count, void* dx, int act, float eps, int B, long HW, int C, int dtype,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Self-attention pieces -- ops.py:467-492 (self_attention_2), ops.py:445-465, ops.py:508-510.
 * bg_maxpool2_fwd/bwd : 2x2/2 max pooling on [B,H,W,C] (H,W even); bwd routes to the first maximum
 * bg_bgemm            : C[b] = alpha * op(A[b]) op(B[b]) (+ C[b] if accumulate); generic strides in
 *                       elements (row stride, col stride, batch stride); fp32 accumulate
 * bg_softmax_fwd/bwd  : rows of [rows, n]: y = softmax(x);  dx = y*(dy - sum(dy*y))
 * bg_attn2_fwd/bwd    : fused flash-style softmax(g f^T) h without materialising the scores (bf16 mode)
 * ------------------------------------------------------------------------------------------- */
/* q [B,Nq,dq], k [B,Nk,dq], v [B,Nk,dv] bf16 (dq a multiple of 16 <= 64: zero-pad the channels; dv a
 * multiple of 16 <= 128).  o [B,Nq,dv] bf16 = softmax(q k^T) v, lse [B,Nq] fp32 = log sum exp of the scores
 * (saved for the backward).  tcgen05 kernels only (no fallback): BG_ERR_UNSUPPORTED elsewhere. */
int bg_attn2_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Nq, int Nk,
                 int dq, int dv, void* stream);
/* the same with a workspace of bg_attn2_fwd_workspace_bytes(B) bytes (B floats): a pre-kernel writes the per-image bound
 * max_i |q_i| * max_j |k_j| >= |q_i . k_j| there; images whose bound is <= 30 skip the row-maximum pass (softmax is invariant
 * to the shift; nothing underflows for shifts that small) and use the bound as the shift -- o and lse are the same numbers. */
size_t bg_attn2_fwd_workspace_bytes(int B);
int bg_attn2_fwd_ws(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Nq, int Nk,
                    int dq, int dv, void* ws, size_t ws_bytes, void* stream);
/* backward of bg_attn2_fwd: d_o [B,Nq,dv] -> dq [B,Nq,dq], dk [B,Nk,dq], dv [B,Nk,dv] (bf16);
 * delta [B,Nq] fp32 scratch (row-wise <d_o, o>).  Recomputes the probabilities from lse.
 * ws (bg_attn2_bwd_workspace_bytes, may be NULL): fp32 per-key-block partials of dQ -- with it the dK/dV kernel also
 * forms dQ and a fixed-order reduce finishes it; without it a second kernel recomputes the probabilities for dQ. */
size_t bg_attn2_bwd_workspace_bytes(int B, int Nq, int Nk, int dq);
int bg_attn2_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                 const float* lse, float* delta, void* dq_out, void* dk_out, void* dv_out, int B, int Nq,
                 int Nk, int dq, int dv, void* ws, size_t ws_bytes, void* stream);
int bg_maxpool2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
int bg_maxpool2_bwd(const void* x, const void* dy, void* dx, int B, int H, int W, int C, int dtype,
                    void* stream);
int bg_bgemm(const void* A, const void* Bm, void* C, int batch, int M, int N, int K,
             long a_rs, long a_cs, long a_bs, long b_rs, long b_cs, long b_bs, long c_rs, long c_bs,
             float alpha, int accumulate, int dtype, void* stream);
int bg_softmax_fwd(const void* x, void* y, long rows, int n, int dtype, void* stream);
int bg_softmax_bwd(const void* y, const void* dy, void* dx, long rows, int n, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Elementwise / small reductions -- ops.py:529-530 (relu), :539-540 (tanh), :503-506
 * (global_sum_pooling), :490 (gamma*o + x), residual adds of ops.py:198/266/313.
 * ------------------------------------------------------------------------------------------- */
int bg_relu_fwd(const void* x, void* y, long n, int dtype, void* stream);
int bg_relu_bwd(const void* x, const void* dy, void* dx, long n, int dtype, void* stream);   /* dx = dy*(x>0) */
int bg_tanh_bwd(const void* y, const void* dy, void* dx, long n, int dtype, void* stream);   /* dx = dy*(1-y^2) */
int bg_add(const void* a, const void* b, void* out, long n, int dtype, void* stream);
int bg_axpby(const void* x, const void* y, const float* alpha_dev, float alpha, float beta,
             void* out, long n, int dtype, void* stream);   /* out = (alpha_dev?*alpha_dev:alpha)*x + beta*y */
int bg_dot(const void* a, const void* b, float* out, long n, int dtype, void* ws, size_t ws_bytes,
           void* stream);                                    /* out[0] = <a,b> (deterministic) */
size_t bg_dot_workspace_bytes(long n);
int bg_sumpool_fwd(const void* x, void* y, int B, long HW, int C, int relu_first, int dtype, void* stream);
int bg_sumpool_bwd(const void* x, const void* dy, void* dx, int B, long HW, int C, int relu_first,
                   int dtype, void* stream);
int bg_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long n, void* stream);
/* 2x2/2 max pooling (ops.py:508-510) of a CHANNEL SLICE of a wider NHWC tensor: x points at the slice's first channel,
 * pixels are ldx elements apart; y / dy are contiguous [B,H/2,W/2,C]; dx is a slice with pixel stride ld_dx, every position
 * of which is written.  bf16, C and strides multiples of 8. */
int bg_maxpool2_slice_fwd(const void* x, long ldx, void* y, int B, int H, int W, int C, int dtype, void* stream);
int bg_maxpool2_slice_bwd(const void* x, long ldx, const void* dy, void* dx, long ld_dx, int B, int H, int W, int C,
                          int dtype, void* stream);
/* channel-slice copy between NHWC tensors: dst[p*ld_dst + c] = src[p*ld_src + c] for p < npix, c < width (strides in
 * elements; the pointers already point at the first channel of the slice) */
int bg_copy_channels(const void* src, long ld_src, void* dst, long ld_dst, long npix, int width, int dtype, void* stream);
/* [npix, Cs] -> [npix, Cd]: copies min(Cs, Cd) channels, zero-fills the rest (channel pad / slice + cast) */
int bg_pad_channels(const void* src, int src_dtype, void* dst, int dst_dtype, long npix, int Cs, int Cd,
                    void* stream);
int bg_reflect_pad(const void* x, void* y, int B, int H, int W, int C, int pt, int pl, int pb, int pr,
                   int dtype, void* stream);
/* y = relu(reflect_pad(x)) in one pass (bf16, C a multiple of 8): the activation in front of the first conv of
 * resblock_down (ops.py:297-301) rides in the padding copy */
int bg_reflect_pad_relu(const void* x, void* y, int B, int H, int W, int C, int pt, int pl, int pb, int pr,
                        int dtype, void* stream);
/* bf16 image tensors stored zero-padded to Cx >= 3 channels (ops.py:49-113 with Cin = 3, the first D convs):
 * z[b,oy,ox,(ky*kw+kx)*3+c] = x[b, R(oy*stride+ky-pad_t), R(ox*stride+kx-pad_l), c] (R: reflect / zero), channels
 * kh*kw*3..Cz-1 zero -- the conv then runs as a 1x1 GEMM with K = Cz.  col2im is the adjoint (gather, deterministic):
 * dx[..., 0..2] = sum of the dz entries that read the pixel, channels 3..Cx-1 zero. */
int bg_im2col_rgb(const void* x, void* z, int B, int H, int W, int Cx, int Ho, int Wo, int kh, int kw, int stride,
                  int pad_t, int pad_l, int pad_mode, int Cz, void* stream);
int bg_col2im_rgb(const void* dz, void* dx, int B, int H, int W, int Cx, int Ho, int Wo, int kh, int kw, int stride,
                  int pad_t, int pad_l, int pad_b, int pad_r, int pad_mode, int Cz, void* stream);
/* RGB OUTPUT conv in tensor-core mode (ops.py:49-113 with channels = 3, kernel 3, stride 1, pad 1 -- `G_logit`,
 * BigGAN.py:570): the nine taps are stacked along the GEMM's N.  bg_rgb_stack_weights builds, from the normalised
 * HWIO weights wn [3,3,Ci,c_pad] (bf16), the 1x1 operand wall [Ci,32] (column (a*3+c)*3+n = wn[a,c,:,n], columns
 * 27..31 zero) and its K-major copy wall_t [32,Ci]; Y = conv1x1(x, wall) is a plain bg_conv2d_fwd on the UNPADDED
 * input; bg_rgb_tapsum_fwd finishes  out[b,i,j,n] = act(bias[n] + sum_{a,c} Y[b,R(i+a-1),R(j+c-1),(a*3+c)*3+n])
 * (R: reflect or zero padding of ONE pixel; out has Co channels, 3..Co-1 zero).  Backward: bg_rgb_tapsum_bwd is the
 * adjoint of the tap sum and of the activation (y = the forward output, d_out its gradient) -> dY [B,H,W,32];
 * dx = bg_conv2d_dgrad of the 1x1 conv (relu_mask allowed), d wall [Ci,32] = bg_conv2d_wgrad of the 1x1 conv, and
 * bg_rgb_unstack_wgrad scatters it back to d wn [3,3,Ci,c_pad] (fp32; columns 3..c_pad-1 zero). */
int bg_rgb_stack_weights(const void* wn, void* wall, void* wall_t, int Ci, int c_pad, void* stream);
int bg_rgb_unstack_wgrad(const float* dwall, float* dwn, int Ci, int c_pad, void* stream);
int bg_rgb_tapsum_fwd(const void* Y, const float* bias, void* out, int B, int H, int W, int Co, int pad_mode,
                      int act, void* stream);
int bg_rgb_tapsum_bwd(const void* y, const void* d_out, void* dY, int B, int H, int W, int Co, int pad_mode,
                      int act, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses -- ops.py:753-797 (discriminator_loss 'hinge'), :799-840 (generator_loss 'hinge'),
 * :847-848 (flood_loss), utils.py:366-369 (cls_loss_fn 'logistic').  Each writes the scalar loss
 * (loss[0], fp32) and the gradient w.r.t. its logits; `inv_count` = 1/global batch (data parallel:
 * local sums scaled by the global count).
 * ------------------------------------------------------------------------------------------- */
int bg_hinge_d_fwd_bwd(const void* real, const void* fake, int n_real, int n_fake, float inv_count,
                       float flood, float* loss, void* dreal, void* dfake, int dtype, void* stream);
int bg_hinge_g_fwd_bwd(const void* fake, int n, float inv_count, float flood, float* loss,
                       void* dfake, int dtype, void* stream);
size_t bg_logistic_cls_workspace_bytes(int B, int n_labels);   /* ws may be NULL (single-block kernel) */
int bg_logistic_cls_fwd_bwd(const void* logits, const float* labels, const float* cls_w, int B,
                            int n_labels, float weight, float inv_count, float* loss, void* dlogits,
                            int accumulate_loss, int dtype, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DiffAugment -- DiffAugment_tf.py:8-73.  x [B,H,W,C]; per-sample randoms are INPUTS:
 * rnd_f [3,B] = U[0,1) for brightness, saturation, contrast; rnd_i [4,B] = translation_x,
 * translation_y, cutout_x, cutout_y (int32, ranges of DiffAugment_tf.py:44-45,:57-58).
 * flags: bit0 color, bit1 translation, bit2 cutout (applied in that order = policy
 * 'color,translation,cutout').  mean_ws [B] fp32 scratch for the contrast mean.
 * ------------------------------------------------------------------------------------------- */
int bg_diffaug_fwd(const void* x, void* y, const float* rnd_f, const int* rnd_i, int flags, int B,
                   int H, int W, int C, float* mean_ws, int dtype, void* stream);
int bg_diffaug_bwd(const void* dy, void* dx, const float* rnd_f, const int* rnd_i, int flags, int B,
                   int H, int W, int C, float* mean_ws, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer -- tf.train.AdamOptimizer + contrib MovingAverageOptimizer (BigGAN.py:923-926).
 * One launch over a flat parameter arena: m,v,p(,ema) updated in place from g.
 * lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) is computed by the caller; lr_t_dev (device scalar, optional) overrides
 * lr_t so that a captured CUDA graph can be replayed with a new step size.  grad_scale multiplies g first.
 * ------------------------------------------------------------------------------------------- */
int bg_adam_ema_step(float* p, const float* g, float* m, float* v, float* ema, long n, float lr_t,
                     const float* lr_t_dev, float beta1, float beta2, float eps, float ema_decay,
                     float grad_scale, void* stream);
/* state[0] = t (step count, as float) += 1; state[1] = lr*sqrt(1-beta2^t)/(1-beta1^t): device-side Adam schedule */
int bg_adam_lr_step(float* state, float lr, float beta1, float beta2, void* stream);
int bg_fill(void* p, float value, long n, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SURVEY 8(f) "next" rows (next_ops.cu): the fork's CLI-default extras and layer variants.
 *   lrelu ops.py:525-526, prelu :532-537 (alpha [C] fp32), avg_pooling :512-514, up_sample :516-519,
 *   tf.nn.depth_to_space of subpixel_conv :23-27, orthogonal regularisers utils.py:185-235.
 * ------------------------------------------------------------------------------------------- */
int bg_lrelu_fwd(const void* x, void* y, long n, float alpha, int dtype, void* stream);
int bg_lrelu_bwd(const void* x, const void* dy, void* dx, long n, float alpha, int dtype, void* stream);
size_t bg_prelu_workspace_bytes(long npix, int C);
int bg_prelu_fwd(const void* x, const float* alpha, void* y, long npix, int C, int dtype, void* stream);
/* dx may be NULL; dalpha [C] fp32 is overwritten (accumulate = 0) or added to (accumulate = 1) */
int bg_prelu_bwd(const void* x, const float* alpha, const void* dy, void* dx, float* dalpha, int accumulate,
                 long npix, int C, int dtype, void* ws, size_t ws_bytes, void* stream);
/* x [B,H,W,C] -> y [B,H/2,W/2,C] (even H, W); the backward takes the INPUT size H, W */
int bg_avgpool2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
int bg_avgpool2_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream);
/* x [B,H,W,C] -> y [B,2H,2W,C] nearest neighbour; the backward sums the 2x2 windows of dy into dx [B,H,W,C] */
int bg_upsample2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);
int bg_upsample2_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream);
/* deep [B,H,W,C] <-> shallow [B,H*r,W*r,C/r^2]; inverse = 0: src deep -> dst shallow (forward), 1: the adjoint */
int bg_depth_to_space(const void* src, void* dst, int B, int H, int W, int C, int r, int inverse, int dtype,
                      void* stream);
/* G [C,C] fp32 = W^T W.  kind 0 'ortho' / 1 'ortho_cosine'.  Writes dLoss/dG, per-row loss terms (rowloss [C]),
 * and loss[0] (= or += scale * l2_loss(reg)).  dLoss/dW = W (dG + dG^T) is left to the caller (bg_bgemm). */
int bg_ortho_reg_rows(const float* G, float* dG, float* rowloss, float* loss, int accumulate, int C,
                      float scale, int kind, void* stream);

/* conditional batch renorm, ops.py:645-715.  bg_renorm_factors: r, d [C] from the batch statistics (mean, var) and the
 * running ones (rmean, rvar) blended by the fade-in weight (weight NULL = 1: shared population statistics); with
 * update_state the running statistics move by `decay` in place and weight_out[0] = weight * fade + one_minus_fade.
 * bg_renorm_affine_fwd: gamma2 = r gamma, beta2 = beta + d gamma over [B,C]; _bwd: dgamma = r dgamma2 + d dbeta2. */
int bg_renorm_factors(const float* mean, const float* var, float* rmean, float* rvar, const float* weight,
                      float* weight_out, float* r, float* d, int C, float eps, float rmin, float rmax, float dmax,
                      float decay, float one_minus_decay, float fade, float one_minus_fade, int update_state,
                      void* stream);
int bg_renorm_affine_fwd(const float* gamma, const float* beta, const float* r, const float* d, float* gamma2,
                         float* beta2, int B, int C, void* stream);
int bg_renorm_affine_bwd(const float* dgamma2, const float* dbeta2, const float* r, const float* d, float* dgamma,
                         int B, int C, void* stream);

/* device half of ImageData.image_processing (utils.py:28-32): src uint8 [B,h,w,c] -> dst [B,S,S,Cd] in `dtype`:
 * TF-1 bilinear resize (align_corners = False), left-right flip where flip[b] != 0 (flip may be NULL), x / 127.5 - 1;
 * channels c..Cd-1 are zero.  The PNG decode half is host code: include/biggan_io.h. */
int bg_image_prepare(const void* src_u8, const int* flip, void* dst, int B, int h, int w, int c, int S, int Cd,
                     int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Grouped fully-connected layers (fc_group.cu) -- ops.py:148-175 as called by condition_batch_norm (ops.py:623-627):
 * all gamma / beta projections of one generator pass in ONE launch (their inputs, z chunk + class vector, exist before
 * the first block), and all their weight / bias gradients in ONE launch at the end of the backward pass.
 *   forward : y_i[M, N_i] (fp32) = x_i[M, K_i] (fp32) . w_i[N_i, ldw_i]^T + bias_i;  w = K-major normalised weights
 *             (the transposed copy bg_sn_scale writes), w_dtype BG_BF16 or BG_F32; K_i, ldw_i multiples of 4.
 *   backward: w_i[K_i, ldw_i] (fp32, OUT) = x_i^T . y_i  (y_i = d loss / d y_i, fp32 [M, N_i]);  bias_i[N_i] (OUT, may be
 *             NULL) = column sums of y_i.  Fixed summation order.  At most 32 problems per call.
 * ------------------------------------------------------------------------------------------- */
typedef struct { const void* x; void* w; void* bias; void* y; int K, N, ldw, reserved; } bg_fc_group_item;
int bg_fc_group_fwd(const bg_fc_group_item* items, int n, int M, int w_dtype, void* stream);
int bg_fc_group_bwd(const bg_fc_group_item* items, int n, int M, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel small exchanges (comm.cu; SURVEY 8e collective C2 -- new design, the reference is single-device):
 * one-shot all-reduce over NVLink peer memory for the cross-replica BatchNorm statistics (2C doubles per BN layer,
 * forward and backward; the sums ops.py:630 takes over the local batch become sums over the global batch).
 *   bg_comm_local_alloc: allocates this rank's exchange buffer on the current device (the one exception to "the library
 *     allocates nothing": the buffer must be a whole cudaMalloc allocation to be exportable) and returns its 64-byte
 *     CUDA IPC handle in handle64_out (host memory).  The caller exchanges the handles of all ranks (host side).
 *   bg_comm_open: handles = world x 64 bytes in rank order (host memory) -> opaque communicator in *comm_out.
 *   bg_comm_allreduce_f64: out[i] = sum over ranks of in[i], n <= 4096 doubles, out may alias in; one kernel on `stream`,
 *     graph-capturable, result bit-identical on every rank (fixed rank order).  Every rank must issue the same
 *     sequence of calls.
 * ------------------------------------------------------------------------------------------- */
size_t bg_comm_buffer_bytes(void);
int bg_comm_local_alloc(void** local_out, void* handle64_out);
int bg_comm_open(int rank, int world, const void* handles, void* local, void** comm_out);
int bg_comm_allreduce_f64(void* comm, const double* in, double* out, int n, void* stream);
int bg_comm_destroy(void* comm);

/* ---------------------------------------------------------------------------------------------
 * Host utility (no device work): CRC-32C of `n` bytes continuing from `init` (0 = fresh); the 32 result bits
 * are returned in the int.  The checksum of TensorFlow's checkpoint format (checkpoint.py; saver at BigGAN.py:1018,
 * :1255-1284).
 * ------------------------------------------------------------------------------------------- */
int bg_crc32c(const void* data, size_t n, int init);

#ifdef __cplusplus
}
#endif
#endif /* BIGGAN_B200_H */
