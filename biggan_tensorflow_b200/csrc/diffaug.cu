// DiffAugment (DiffAugment_tf.py:8-73): color (brightness :20-23, saturation :26-30, contrast :33-37),
// translation (:40-50), cutout (:53-66) fused into one stats pass + one apply pass, forward and
// backward.  HBM-bound gather/elementwise work; the integer index math follows the reference exactly
// (bit-exact masks / shifts for identical sampled offsets, which are inputs).
#include "bg_common.cuh"

#define DA_COLOR 1
#define DA_TRANS 2
#define DA_CUTOUT 4

__device__ __forceinline__ int da_cutout_size(int S) { return (int)((float)S * 0.5f + 0.5f); }   // :56 int(S*ratio+0.5)

// 1 if output pixel (i,j) is zeroed by rand_cutout (:57-66); rows [ox-cs/2, ox-cs/2+cs-1] clipped
__device__ __forceinline__ bool da_cut(int i, int j, int ox, int oy, int H, int W) {
  const int cs0 = da_cutout_size(H), cs1 = da_cutout_size(W);
  const int lo_i = ox - cs0 / 2, lo_j = oy - cs1 / 2;
  return i >= lo_i && i <= lo_i + cs0 - 1 && j >= lo_j && j <= lo_j + cs1 - 1;
}

// per-sample mean over (H,W,C):  MODE 0 of x (contrast mean, forward)
//                                MODE 1 of the gradient arriving at the colour stage (backward)
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
da_mean_kernel(const T* __restrict__ t, const int* __restrict__ rnd_i, int flags, int B, int H, int W, int C,
               float* __restrict__ mean_out) {
  __shared__ float sh[33];
  const int b = blockIdx.x;
  const long n = (long)H * W * C;
  const T* tb = t + (long)b * n;
  float s = 0.f;
  if (MODE == 0) {
    for (long i = threadIdx.x; i < n; i += blockDim.x) s += bg_ld(tb, i);
  } else {
    const int tx = (flags & DA_TRANS) ? rnd_i[0 * B + b] : 0, ty = (flags & DA_TRANS) ? rnd_i[1 * B + b] : 0;
    const int ox = rnd_i[2 * B + b], oy = rnd_i[3 * B + b];
    // gradient at colour-stage pixel (y,x) = dout[y-tx, x-ty] if that output pixel exists and is not cut
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
      const int c = (int)(i % C); const long p = i / C;
      const int x = (int)(p % W), y = (int)(p / W);
      const int oi = y - tx, oj = x - ty;
      if (oi < 0 || oi >= H || oj < 0 || oj >= W) continue;
      if ((flags & DA_CUTOUT) && da_cut(oi, oj, ox, oy, H, W)) continue;
      s += bg_ld(tb, ((long)oi * W + oj) * C + c);
    }
  }
  s = bg_block_sum(s, sh);
  if (threadIdx.x == 0) mean_out[b] = s / (float)n;
}

template <typename T>
__global__ void __launch_bounds__(256)
da_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ rnd_f, const int* __restrict__ rnd_i,
              int flags, int B, int H, int W, int C, const float* __restrict__ mean_ws) {
  const long npix = (long)B * H * W;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long)gridDim.x * blockDim.x) {
    const int j = (int)(p % W); const long r = p / W;
    const int i = (int)(r % H); const int b = (int)(r / H);
    T* yo = y + p * C;
    bool zero = false;
    int si = i, sj = j;
    if (flags & DA_CUTOUT) zero = da_cut(i, j, rnd_i[2 * B + b], rnd_i[3 * B + b], H, W);
    if (flags & DA_TRANS) {
      si = i + rnd_i[0 * B + b];      // translation_x shifts axis 1 (H)   (:46,:48)
      sj = j + rnd_i[1 * B + b];      // translation_y shifts axis 2 (W)   (:47,:49)
      if (si < 0 || si >= H || sj < 0 || sj >= W) zero = true;
    }
    if (zero) { for (int c = 0; c < C; ++c) bg_st(yo, c, 0.f); continue; }
    const T* xs = x + (((long)b * H + si) * W + sj) * C;
    if (flags & DA_COLOR) {
      const float br = rnd_f[0 * B + b] - 0.5f;            // :21
      const float sat = rnd_f[1 * B + b] * 2.f;            // :27
      const float con = rnd_f[2 * B + b] + 0.5f;           // :34
      const float m_all = mean_ws[b] + br;                 // mean_hwc after brightness (saturation keeps it)
      float mc = 0.f;
      for (int c = 0; c < C; ++c) mc += bg_ld(xs, c) + br;
      mc /= (float)C;                                      // :28 reduce_mean(axis=3)
      for (int c = 0; c < C; ++c) {
        float v = bg_ld(xs, c) + br;
        v = (v - mc) * sat + mc;                           // :29
        v = (v - m_all) * con + m_all;                     // :36
        bg_st(yo, c, v);
      }
    } else {
      for (int c = 0; c < C; ++c) yo[c] = xs[c];
    }
  }
}

// backward: g3(y,x) = mask*dout[y-tx, x-ty]; contrast: g2 = con*g3 + (1-con)*mean_hwc(g3);
// saturation: g1 = sat*g2 + (1-sat)*mean_c(g2); brightness: identity.
template <typename T>
__global__ void __launch_bounds__(256)
da_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, const float* __restrict__ rnd_f, const int* __restrict__ rnd_i,
              int flags, int B, int H, int W, int C, const float* __restrict__ mean_ws) {
  const long npix = (long)B * H * W;
  for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long)gridDim.x * blockDim.x) {
    const int x = (int)(p % W); const long r = p / W;
    const int y = (int)(r % H); const int b = (int)(r / H);
    int oi = y, oj = x;
    if (flags & DA_TRANS) { oi = y - rnd_i[0 * B + b]; oj = x - rnd_i[1 * B + b]; }
    bool live = oi >= 0 && oi < H && oj >= 0 && oj < W;
    if (live && (flags & DA_CUTOUT)) live = !da_cut(oi, oj, rnd_i[2 * B + b], rnd_i[3 * B + b], H, W);
    const T* src = dy + (((long)b * H + (live ? oi : 0)) * W + (live ? oj : 0)) * C;
    T* dst = dx + p * C;
    if (flags & DA_COLOR) {
      const float sat = rnd_f[1 * B + b] * 2.f, con = rnd_f[2 * B + b] + 0.5f;
      const float gm = mean_ws[b];
      float mc = 0.f;
      for (int c = 0; c < C; ++c) mc += con * (live ? bg_ld(src, c) : 0.f) + (1.f - con) * gm;
      mc /= (float)C;
      for (int c = 0; c < C; ++c) {
        const float g2 = con * (live ? bg_ld(src, c) : 0.f) + (1.f - con) * gm;
        bg_st(dst, c, sat * g2 + (1.f - sat) * mc);
      }
    } else {
      for (int c = 0; c < C; ++c) bg_st(dst, c, live ? bg_ld(src, c) : 0.f);
    }
  }
}

extern "C" int bg_diffaug_fwd(const void* x, void* y, const float* rnd_f, const int* rnd_i, int flags, int B, int H, int W,
                              int C, float* mean_ws, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BG_REQUIRE(x != y, BG_ERR_UNSUPPORTED, "diffaug: in-place not supported");
  if (flags & DA_COLOR) {
    BG_REQUIRE(mean_ws, BG_ERR_WORKSPACE, "diffaug: mean_ws required for color");
    BG_DISPATCH_DTYPE(dtype, T, da_mean_kernel<T, 0><<<B, 256, 0, st>>>((const T*)x, rnd_i, flags, B, H, W, C, mean_ws));
    BG_LAUNCH_CHECK();
  }
  const long npix = (long)B * H * W;
  const int blocks = (int)min(bg_cdiv(npix, 256), (long)bg_num_sms() * 16);
  BG_DISPATCH_DTYPE(dtype, T, da_fwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, (T*)y, rnd_f, rnd_i, flags, B, H, W, C, mean_ws));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_diffaug_bwd(const void* dy, void* dx, const float* rnd_f, const int* rnd_i, int flags, int B, int H, int W,
                              int C, float* mean_ws, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BG_REQUIRE(dx != dy, BG_ERR_UNSUPPORTED, "diffaug: in-place not supported");
  if (flags & DA_COLOR) {
    BG_REQUIRE(mean_ws, BG_ERR_WORKSPACE, "diffaug: mean_ws required for color");
    BG_DISPATCH_DTYPE(dtype, T, da_mean_kernel<T, 1><<<B, 256, 0, st>>>((const T*)dy, rnd_i, flags, B, H, W, C, mean_ws));
    BG_LAUNCH_CHECK();
  }
  const long npix = (long)B * H * W;
  const int blocks = (int)min(bg_cdiv(npix, 256), (long)bg_num_sms() * 16);
  BG_DISPATCH_DTYPE(dtype, T, da_bwd_kernel<T><<<blocks, 256, 0, st>>>((const T*)dy, (T*)dx, rnd_f, rnd_i, flags, B, H, W, C, mean_ws));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
