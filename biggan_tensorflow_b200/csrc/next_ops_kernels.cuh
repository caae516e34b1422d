// Device kernels of next_ops.cu, kept free of includes and launch syntax so that the same source also
// compiles for the host under tests/host_emul/cuda_host_emul.h (the CPU execution check of tests/test_next_ops_host_emul.py).
// Include bg_common.cuh (or the host emulation prelude) first.
#pragma once

#define NX_LOOP(i, n) \
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x, _s = (long)gridDim.x * blockDim.x; i < (n); i += _s)

// ------------------------------------------------------------------------------------------------------------
// leaky relu: y = max(x, alpha x)  (tf.nn.leaky_relu);  dx = dy * (x > 0 ? 1 : alpha)  (LeakyReluGrad)
// ------------------------------------------------------------------------------------------------------------
template <typename T> __global__ void lrelu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long n, float alpha) {
  NX_LOOP(i, n) { const float v = bg_ld(x, i); bg_st(y, i, fmaxf(v, alpha * v)); }
}
template <typename T> __global__ void lrelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long n, float alpha) {
  NX_LOOP(i, n) { const float g = bg_ld(dy, i); bg_st(dx, i, bg_ld(x, i) > 0.f ? g : alpha * g); }
}
// ------------------------------------------------------------------------------------------------------------
// prelu (ops.py:532-537): y = relu(x) + alpha_c (x - |x|) / 2, alpha per channel (fp32 parameter).
//   dx = dy * (x > 0 ? 1 : x < 0 ? alpha_c : alpha_c / 2);   dalpha_c = sum_pixels dy * min(x, 0)
// The channel reduction is two fixed-order stages: each CTA sums a slab of pixels for every channel (threads along
// the channel axis, coalesced), a second kernel adds the slabs in order.
// ------------------------------------------------------------------------------------------------------------
template <typename T> __global__ void prelu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ alpha, T* __restrict__ y, long n, int C) {
  NX_LOOP(i, n) {
    const float v = bg_ld(x, i);
    bg_st(y, i, v > 0.f ? v : alpha[i % C] * v);
  }
}
template <typename T> __global__ void prelu_bwd_dx_kernel(const T* __restrict__ x, const float* __restrict__ alpha, const T* __restrict__ dy,
                                                          T* __restrict__ dx, long n, int C) {
  NX_LOOP(i, n) {
    const float v = bg_ld(x, i), g = bg_ld(dy, i), a = alpha[i % C];
    bg_st(dx, i, v > 0.f ? g : (v < 0.f ? a * g : 0.5f * a * g));
  }
}
template <typename T> __global__ void __launch_bounds__(256) prelu_bwd_dalpha_partial_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                                          float* __restrict__ partial, long npix, int C, long rows_per_cta) {
  const long r0 = (long)blockIdx.x * rows_per_cta;
  const long r1 = r0 + rows_per_cta < npix ? r0 + rows_per_cta : npix;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (long r = r0; r < r1; ++r) acc += bg_ld(dy, r * C + c) * fminf(bg_ld(x, r * C + c), 0.f);
    partial[(long)blockIdx.x * C + c] = acc;
  }
}
__global__ void prelu_bwd_dalpha_final_kernel(const float* __restrict__ partial, float* __restrict__ dalpha, int nslab, int C, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int s = 0; s < nslab; ++s) acc += partial[(long)s * C + c];
  dalpha[c] = accumulate ? dalpha[c] + acc : acc;
}
// ------------------------------------------------------------------------------------------------------------
// 2x2/2 average pooling (tf.layers.average_pooling2d(2, 2, 'SAME') on even maps), nearest-neighbour x2 up-sampling
// (tf.image.resize_nearest_neighbor to twice the size): each is the other's adjoint up to the factor 1/4.
// ------------------------------------------------------------------------------------------------------------
template <typename T> __global__ void pool2_sum_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, long n, float scale) {
  const int Ho = H / 2, Wo = W / 2;       // x [B,H,W,C] -> y [B,H/2,W/2,C] = scale * (sum of the 2x2 window)
  NX_LOOP(i, n) {
    const int c = (int)(i % C); long r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long b = r / Ho;
    const long base = ((b * H + 2 * oy) * W + 2 * ox) * C + c;
    const float s = bg_ld(x, base) + bg_ld(x, base + C) + bg_ld(x, base + (long)W * C) + bg_ld(x, base + (long)W * C + C);
    bg_st(y, i, s * scale);
  }
}
template <typename T> __global__ void repeat2_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, long n, float scale) {
  const int Ho = 2 * H, Wo = 2 * W;       // x [B,H,W,C] -> y [B,2H,2W,C], y[.., oy, ox, ..] = scale * x[.., oy/2, ox/2, ..]
  NX_LOOP(i, n) {
    const int c = (int)(i % C); long r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long b = r / Ho;
    bg_st(y, i, scale * bg_ld(x, ((b * H + oy / 2) * W + ox / 2) * C + c));
  }
}
// ------------------------------------------------------------------------------------------------------------
// tf.nn.depth_to_space (NHWC, block r): y[b, h r + i, w r + j, c] = x[b, h, w, (i r + j) Co + c], Co = C / r^2.
// inverse != 0 runs the adjoint (space_to_depth) with the same index map: x <- y.
// ------------------------------------------------------------------------------------------------------------
template <typename T> __global__ void depth_to_space_kernel(const T* __restrict__ src, T* __restrict__ dst, int H, int W, int C, int r, long n, int inverse) {
  const int Co = C / (r * r);
  NX_LOOP(i, n) {                          // i enumerates the [B, H r, W r, Co] tensor
    const int c = (int)(i % Co); long t = i / Co;
    const int ox = (int)(t % ((long)W * r)); t /= (long)W * r;
    const int oy = (int)(t % ((long)H * r)); const long b = t / ((long)H * r);
    const long deep = ((b * H + oy / r) * W + ox / r) * C + ((oy % r) * r + (ox % r)) * Co + c;
    if (inverse) bg_st(dst, deep, bg_ld(src, i)); else bg_st(dst, i, bg_ld(src, deep));
  }
}
// ------------------------------------------------------------------------------------------------------------
// Orthogonal regularisers (utils.py:185-235) on the Gram matrix G = W^T W [C, C] (fp32):
//   kind 0 'ortho':        reg = G - I
//   kind 1 'ortho_cosine': reg = cosine_similarity(G, 1 - I):  na_i = G_i * rsqrt(max(|G_i|^2, 1e-12)),
//                          reg_ij = (sum_k na_ik - na_ij) / sqrt(C - 1)       (every row of 1 - I has norm sqrt(C-1))
//   loss = scale * sum(reg^2) / 2.
// Every row of reg depends on one row of G only: one CTA per row writes dLoss/dG for that row and the row's loss term;
// bg_ortho_reg_rows then adds the C row terms in order.  dLoss/dW = W (dG + dG^T) is two GEMMs on the caller's side.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ortho_rows_kernel(const float* __restrict__ G, float* __restrict__ dG, float* __restrict__ rowloss,
                                                          int C, float scale, int kind) {
  __shared__ float sh[40];
  const int i = blockIdx.x;
  const float* g = G + (long)i * C;
  float* d = dG + (long)i * C;
  if (kind == 0) {
    float acc = 0.f;
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
      const float r = g[k] - (k == i ? 1.f : 0.f);
      d[k] = scale * r;
      acc += r * r;
    }
    acc = bg_block_sum(acc, sh);
    if (threadIdx.x == 0) rowloss[i] = 0.5f * scale * acc;
    return;
  }
  float ss = 0.f, rs = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) { const float v = g[k]; ss += v * v; rs += v; }
  ss = bg_block_sum(ss, sh);
  rs = bg_block_sum(rs, sh);
  const bool clamped = ss < 1e-12f;
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  const float rsn = rs * inv;                       // sum_k na_ik
  const float q = C > 1 ? rsqrtf((float)(C - 1)) : 0.f;
  // S_ik = (rsn - na_ik) q;  loss_i = scale/2 sum_k S_ik^2;  dS = scale S
  // dna_ik = q (sum_j dS_ij - dS_ik);  dG_ik = inv (dna_ik - na_ik <dna_i, na_i>)
  float l = 0.f, sds = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) { const float s = (rsn - g[k] * inv) * q; l += s * s; sds += s; }
  l = bg_block_sum(l, sh);
  sds = bg_block_sum(sds, sh) * scale;              // sum_j dS_ij
  float dot = 0.f;
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float na = g[k] * inv;
    const float dna = q * (sds - scale * (rsn - na) * q);
    dot += dna * na;
  }
  dot = bg_block_sum(dot, sh);
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float na = g[k] * inv;
    const float dna = q * (sds - scale * (rsn - na) * q);
    d[k] = clamped ? inv * dna : inv * (dna - na * dot);      // under the eps clamp the normaliser is a constant
  }
  if (threadIdx.x == 0) rowloss[i] = 0.5f * scale * l;
}
__global__ void ortho_sum_rows_kernel(const float* __restrict__ rowloss, int C, float* __restrict__ loss, int accumulate) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int i = 0; i < C; ++i) acc += rowloss[i];
    loss[0] = accumulate ? loss[0] + acc : acc;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Conditional batch RE-normalisation (ops.py:645-715).  Per channel, from the batch statistics and the training-time
// running statistics (blended with the batch ones by the fade-in weight w):
//   sigma = sqrt(var + eps), rsigma = sqrt(rvar + eps), wsig = w rsigma + (1-w) sigma, wmean = w rmean + (1-w) mean
//   r = clip(sigma / wsig, rmin, rmax),   d = clip((mean - wmean) / wsig, -dmax, dmax)        (constants in the backward)
// then, unless the statistics are the shared population ones (update_state = 0), the running statistics move by
// `decay` and the fade-in weight by `fade` (the complements 1 - decay, 1 - fade are passed in: the reference forms them
// in double precision before they become float32 constants).  The normalisation itself is a conditional batch norm with
//   gamma' = r gamma,  beta' = beta + d gamma      (per sample and channel)
// whose adjoint is  dgamma = r dgamma' + d dbeta',  dbeta = dbeta'.
// ------------------------------------------------------------------------------------------------------------
__global__ void renorm_factors_kernel(const float* __restrict__ mean, const float* __restrict__ var, float* __restrict__ rmean,
                                      float* __restrict__ rvar, const float* __restrict__ weight, float* __restrict__ weight_out,
                                      float* __restrict__ r, float* __restrict__ d, int C, float eps, float rmin, float rmax, float dmax,
                                      float decay, float one_minus_decay, float fade, float one_minus_fade, int update_state) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const float w = weight ? weight[0] : 1.f;
  if (c == 0 && weight_out) weight_out[0] = update_state ? w * fade + one_minus_fade : w;
  if (c >= C) return;
  const float m = mean[c], v = var[c], rm = rmean[c], rv = rvar[c];
  const float sigma = sqrtf(v + eps), rsigma = sqrtf(rv + eps);
  const float wsig = w * rsigma + (1.f - w) * sigma;
  const float wmean = w * rm + (1.f - w) * m;
  r[c] = fminf(fmaxf(sigma / wsig, rmin), rmax);
  d[c] = fminf(fmaxf((m - wmean) / wsig, -dmax), dmax);
  if (update_state) {
    rmean[c] = rm * decay + m * one_minus_decay;
    rvar[c] = rv * decay + v * one_minus_decay;
  }
}
__global__ void renorm_affine_fwd_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ r,
                                         const float* __restrict__ d, float* __restrict__ gamma2, float* __restrict__ beta2, long n, int C) {
  NX_LOOP(i, n) {
    const int c = (int)(i % C);
    const float g = gamma[i];
    gamma2[i] = r[c] * g;
    beta2[i] = beta[i] + d[c] * g;
  }
}
__global__ void renorm_affine_bwd_kernel(const float* __restrict__ dgamma2, const float* __restrict__ dbeta2, const float* __restrict__ r,
                                         const float* __restrict__ d, float* __restrict__ dgamma, long n, int C) {
  NX_LOOP(i, n) {           // dbeta = dbeta2 needs no kernel
    const int c = (int)(i % C);
    dgamma[i] = r[c] * dgamma2[i] + d[c] * dbeta2[i];
  }
}

// ------------------------------------------------------------------------------------------------------------
// Device half of ImageData.image_processing (utils.py:28-32): tf.image.resize_images(x, [S, S]) (TF-1 bilinear,
// align_corners = False: source coordinate = index * in / out, neighbours floor and min(ceil, in - 1)),
// tf.image.random_flip_left_right (the coin is an input, one int per image), cast / 127.5 - 1.
// src uint8 [B, h, w, c] -> dst [B, S, S, Cd] (channels c..Cd-1 zero: the 16-channel image layout of the bf16 mode).
// ------------------------------------------------------------------------------------------------------------
template <typename T> __global__ void image_prepare_kernel(const unsigned char* __restrict__ src, const int* __restrict__ flip, T* __restrict__ dst,
                                                           int h, int w, int c, int S, int Cd, long n) {
  const float sy = (float)h / (float)S, sx = (float)w / (float)S;
  NX_LOOP(i, n) {
    const int ch = (int)(i % Cd); long t = i / Cd;
    const int ox = (int)(t % S); t /= S;
    const int oy = (int)(t % S); const long b = t / S;
    if (ch >= c) { bg_st(dst, i, 0.f); continue; }
    const int rx = (flip && flip[b]) ? S - 1 - ox : ox;          // the flip acts on the resized image
    const float fy = oy * sy, fx = rx * sx;
    const int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    const int y1 = min((int)ceilf(fy), h - 1), x1 = min((int)ceilf(fx), w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const unsigned char* img = src + b * (long)h * w * c + ch;
    const float tl = img[((long)y0 * w + x0) * c], tr = img[((long)y0 * w + x1) * c];
    const float bl = img[((long)y1 * w + x0) * c], br = img[((long)y1 * w + x1) * c];
    const float top = tl + (tr - tl) * lx, bot = bl + (br - bl) * lx;
    bg_st(dst, i, (top + (bot - top) * ly) / 127.5f - 1.f);
  }
}
