// SURVEY 8(f) "next" rows: the fork's CLI-default extras and layer variants around the hot path.
//   lrelu / prelu (ops.py:525-537), avg_pooling (:512-514), up_sample (:516-519), depth_to_space of subpixel_conv
//   (:23-27), the orthogonal regularisers of utils.py:180-235 on a Gram matrix.
// All of it is HBM-bound elementwise / gather / row-reduction work: coalesced channel-fastest indexing, grids capped
// at 16 CTAs per SM, deterministic fixed-order reductions (replica-consistent gradients).
#include "bg_common.cuh"

#include "next_ops_kernels.cuh"

static inline int nx_blocks(long n) {
  long b = bg_cdiv(n, 256L);
  long mx = (long)bg_num_sms() * 16;
  if (b > mx) b = mx;
  if (b < 1) b = 1;
  return (int)b;
}
extern "C" int bg_lrelu_fwd(const void* x, void* y, long n, float alpha, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(lrelu_fwd_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)x, (T*)y, n, alpha));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_lrelu_bwd(const void* x, const void* dy, void* dx, long n, float alpha, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(lrelu_bwd_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)x, (const T*)dy, (T*)dx, n, alpha));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

static inline int prelu_slabs(long npix) {
  long s = bg_cdiv(npix, 64);
  long mx = (long)bg_num_sms() * 8;
  if (s > mx) s = mx;
  if (s < 1) s = 1;
  return (int)s;
}
extern "C" size_t bg_prelu_workspace_bytes(long npix, int C) { return (size_t)prelu_slabs(npix) * C * sizeof(float); }
extern "C" int bg_prelu_fwd(const void* x, const float* alpha, void* y, long npix, int C, int dtype, void* stream) {
  const long n = npix * C;
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(prelu_fwd_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)x, alpha, (T*)y, n, C));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_prelu_bwd(const void* x, const float* alpha, const void* dy, void* dx, float* dalpha, int accumulate, long npix, int C,
                            int dtype, void* ws, size_t ws_bytes, void* stream) {
  const long n = npix * C;
  const int slabs = prelu_slabs(npix);
  BG_REQUIRE(ws_bytes >= (size_t)slabs * C * sizeof(float), BG_ERR_WORKSPACE, "prelu_bwd: workspace too small");
  const long rows = bg_cdiv(npix, slabs);
  cudaStream_t st = (cudaStream_t)stream;
  if (dx) {
    BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(prelu_bwd_dx_kernel<T>, nx_blocks(n), 256, st, (const T*)x, alpha, (const T*)dy, (T*)dx, n, C));
    BG_LAUNCH_CHECK();
  }
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(prelu_bwd_dalpha_partial_kernel<T>, slabs, 256, st, (const T*)x, (const T*)dy, (float*)ws, npix, C, rows));
  BG_LAUNCH_CHECK();
  BG_LAUNCH(prelu_bwd_dalpha_final_kernel, (unsigned)bg_cdiv(C, 128), 128, st, (const float*)ws, dalpha, slabs, C, accumulate);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_avgpool2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  BG_REQUIRE(H % 2 == 0 && W % 2 == 0, BG_ERR_UNSUPPORTED, "avgpool2: odd size");
  const long n = (long)B * (H / 2) * (W / 2) * C;
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(pool2_sum_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)x, (T*)y, H, W, C, n, 0.25f));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_avgpool2_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream) {
  BG_REQUIRE(H % 2 == 0 && W % 2 == 0, BG_ERR_UNSUPPORTED, "avgpool2: odd size");     // H, W: size of the pooled INPUT (= of dx)
  const long n = (long)B * H * W * C;
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(repeat2_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)dy, (T*)dx, H / 2, W / 2, C, n, 0.25f));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_upsample2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  const long n = (long)B * (2 * H) * (2 * W) * C;                                      // H, W: size of x
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(repeat2_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)x, (T*)y, H, W, C, n, 1.0f));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_upsample2_bwd(const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream) {
  const long n = (long)B * H * W * C;                                                  // H, W: size of x (= of dx); dy is [B,2H,2W,C]
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(pool2_sum_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)dy, (T*)dx, 2 * H, 2 * W, C, n, 1.0f));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_depth_to_space(const void* src, void* dst, int B, int H, int W, int C, int r, int inverse, int dtype, void* stream) {
  BG_REQUIRE(r >= 1 && C % (r * r) == 0, BG_ERR_SHAPE, "depth_to_space: C=%d is not a multiple of r^2 (r=%d)", C, r);
  const long n = (long)B * H * W * C;      // H, W, C: the DEEP tensor [B,H,W,C]
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(depth_to_space_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const T*)src, (T*)dst, H, W, C, r, n, inverse));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_ortho_reg_rows(const float* G, float* dG, float* rowloss, float* loss, int accumulate, int C, float scale, int kind,
                                 void* stream) {
  BG_REQUIRE(kind == 0 || kind == 1, BG_ERR_UNSUPPORTED, "Unknown regularization method.");
  BG_LAUNCH(ortho_rows_kernel, C, 256, (cudaStream_t)stream, G, dG, rowloss, C, scale, kind);
  BG_LAUNCH_CHECK();
  BG_LAUNCH(ortho_sum_rows_kernel, 1, 32, (cudaStream_t)stream, rowloss, C, loss, accumulate);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_renorm_factors(const float* mean, const float* var, float* rmean, float* rvar, const float* weight, float* weight_out,
                                 float* r, float* d, int C, float eps, float rmin, float rmax, float dmax, float decay, float one_minus_decay,
                                 float fade, float one_minus_fade, int update_state, void* stream) {
  BG_LAUNCH(renorm_factors_kernel, (unsigned)bg_cdiv(C, 128), 128, (cudaStream_t)stream, mean, var, rmean, rvar, weight, weight_out, r, d, C,
            eps, rmin, rmax, dmax, decay, one_minus_decay, fade, one_minus_fade, update_state);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_renorm_affine_fwd(const float* gamma, const float* beta, const float* r, const float* d, float* gamma2, float* beta2,
                                    int B, int C, void* stream) {
  const long n = (long)B * C;
  BG_LAUNCH(renorm_affine_fwd_kernel, nx_blocks(n), 256, (cudaStream_t)stream, gamma, beta, r, d, gamma2, beta2, n, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_renorm_affine_bwd(const float* dgamma2, const float* dbeta2, const float* r, const float* d, float* dgamma, int B, int C,
                                    void* stream) {
  const long n = (long)B * C;
  BG_LAUNCH(renorm_affine_bwd_kernel, nx_blocks(n), 256, (cudaStream_t)stream, dgamma2, dbeta2, r, d, dgamma, n, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_image_prepare(const void* src_u8, const int* flip, void* dst, int B, int h, int w, int c, int S, int Cd, int dtype,
                                void* stream) {
  BG_REQUIRE(B > 0 && h > 0 && w > 0 && c > 0 && S > 0 && Cd >= c, BG_ERR_SHAPE, "image_prepare: bad shape");
  const long n = (long)B * S * S * Cd;
  BG_DISPATCH_DTYPE(dtype, T, BG_LAUNCH(image_prepare_kernel<T>, nx_blocks(n), 256, (cudaStream_t)stream, (const unsigned char*)src_u8, flip,
                                        (T*)dst, h, w, c, S, Cd, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
