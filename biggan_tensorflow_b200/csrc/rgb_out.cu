// Tap-stacked RGB output conv (G_logit, ops.py:49-113 with channels = 3, kernel 3, stride 1, pad 1; BigGAN.py:570).
// With 3 (16-padded) output channels a 3x3 conv costs 54 tensor-core instructions per 128-pixel tile whose N is 16: the tensor pipe's
// operand-read floor, 7x above the HBM time of the layer.  Instead the nine taps are stacked along N,
//   Y[b,i,j,(a*3+c)*3+n] = sum_ci x[b,i,j,ci] W[a,c,ci,n]        (ONE 1x1 GEMM Ci -> 27 -> 32 on the UNPADDED input),
//   out[b,i,j,n] = act(bias[n] + sum_{a,c} Y[b, R(i+a-1), R(j+c-1), (a*3+c)*3+n])   (R: reflect / zero; rgb_tapsum_fwd_kernel),
// and the backward pass needs only the adjoint of the tap sum,
//   dY[b,i,j,(a*3+c)*3+n] = sum over output pixels p with R(p + (a-1,c-1)) = (i,j) of d pre-activation[b,p,n],
// after which dx = relu-mask(dY Wall^T) is a 1x1 GEMM 32 -> Ci and dWall = x^T dY a 1x1 weight gradient: no padded copy of x,
// no padded-domain gradient and no fold pass.  All of it HBM-bound elementwise / gather work here.
#include "bg_common.cuh"

#include "rgb_out_kernels.cuh"

extern "C" int bg_rgb_stack_weights(const void* wn, void* wall, void* wall_t, int Ci, int c_pad, void* stream) {
  BG_REQUIRE(Ci > 0 && c_pad >= 3, BG_ERR_SHAPE, "rgb_stack_weights: bad shape");
  BG_LAUNCH(rgb_stack_weights_kernel, (int)bg_cdiv((long)Ci * 32, 256), 256, stream, (const bf16*)wn, (bf16*)wall, (bf16*)wall_t, Ci, c_pad);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_rgb_unstack_wgrad(const float* dwall, float* dwn, int Ci, int c_pad, void* stream) {
  BG_REQUIRE(Ci > 0 && c_pad >= 3, BG_ERR_SHAPE, "rgb_unstack_wgrad: bad shape");
  BG_LAUNCH(rgb_unstack_wgrad_kernel, (int)bg_cdiv((long)9 * Ci * c_pad, 256), 256, stream, dwall, dwn, Ci, c_pad);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_rgb_tapsum_fwd(const void* Y, const float* bias, void* out, int B, int H, int W, int Co, int pad_mode, int act,
                                 void* stream) {
  BG_REQUIRE(B > 0 && H >= 2 && W >= 2 && Co >= 8 && Co % 8 == 0, BG_ERR_SHAPE, "rgb_tapsum_fwd: bad shape");
  BG_REQUIRE((((uintptr_t)Y | (uintptr_t)out) & 15) == 0, BG_ERR_SHAPE, "rgb_tapsum_fwd: buffers must be 16-byte aligned");
  const long nwarps = (long)B * H * ((W + 29) / 30);
  long blocks = bg_cdiv(nwarps, 8);
  if (blocks > (long)bg_num_sms() * 16) blocks = (long)bg_num_sms() * 16;
  BG_LAUNCH(rgb_tapsum_fwd_kernel, (int)blocks, 256, stream, (const bf16*)Y, bias, (bf16*)out, B, H, W, Co, pad_mode == BG_PAD_REFLECT ? 1 : 0,
            act);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_rgb_tapsum_bwd(const void* y, const void* d_out, void* dY, int B, int H, int W, int Co, int pad_mode, int act,
                                 void* stream) {
  BG_REQUIRE(B > 0 && H >= 2 && W >= 2 && Co >= 4 && Co % 4 == 0, BG_ERR_SHAPE, "rgb_tapsum_bwd: bad shape");
  BG_REQUIRE((((uintptr_t)y | (uintptr_t)d_out) & 7) == 0 && ((uintptr_t)dY & 15) == 0, BG_ERR_SHAPE, "rgb_tapsum_bwd: misaligned buffers");
  const long total = (long)B * H * W;
  long blocks = bg_cdiv(total, 256);
  if (blocks > (long)bg_num_sms() * 16) blocks = (long)bg_num_sms() * 16;
  BG_LAUNCH(rgb_tapsum_bwd_kernel, (int)blocks, 256, stream, (const bf16*)y, (const bf16*)d_out, (bf16*)dY, B, H, W, Co,
            pad_mode == BG_PAD_REFLECT ? 1 : 0, act);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
