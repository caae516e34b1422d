// Device kernels of rgb_out.cu (tap-stacked RGB output conv), kept free of includes and launch syntax so that the same source
// also compiles for the host under tests/host_emul/cuda_host_emul.h (CPU execution check: tests/test_rgb_out_host_emul.py).
// Include bg_common.cuh (or the host emulation prelude) first.  bf16 tensors are addressed as raw 16-bit patterns.
#pragma once

struct alignas(16) RgbU4 { uint32_t x, y, z, w; };
struct alignas(8) RgbU2 { uint32_t x, y; };
__device__ __forceinline__ float rgb_bits_to_float(uint32_t h) {
  const uint32_t u = h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
__device__ __forceinline__ uint32_t rgb_float_to_bits(float f) {          // round to nearest even
#ifdef BG_HOST_EMULATION
  return bg_emul_float_to_bf16(f).v;
#else
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16(f));
#endif
}
__device__ __forceinline__ float rgb_act(float v, int act) {
  return act == BG_ACT_RELU ? fmaxf(v, 0.f) : (act == BG_ACT_TANH ? tanhf(v) : v);
}

__global__ void __launch_bounds__(256)
rgb_stack_weights_kernel(const bf16* __restrict__ wn, bf16* __restrict__ wall, bf16* __restrict__ wall_t, int Ci, int c_pad) {
  const uint16_t* w = reinterpret_cast<const uint16_t*>(wn);
  uint16_t* a = reinterpret_cast<uint16_t*>(wall);
  uint16_t* t = reinterpret_cast<uint16_t*>(wall_t);
  const int total = Ci * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c = idx >> 5, col = idx & 31;
    uint16_t v = 0;
    if (col < 27) v = w[((long)(col / 3) * Ci + c) * c_pad + col % 3];
    a[idx] = v;                        // HWIO of the 1x1 conv: [Ci][32]
    t[(long)col * Ci + c] = v;         // K-major copy: [32][Ci]
  }
}
__global__ void __launch_bounds__(256)
rgb_unstack_wgrad_kernel(const float* __restrict__ dwall, float* __restrict__ dwn, int Ci, int c_pad) {
  const int total = 9 * Ci * c_pad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx % c_pad;
    const int r = idx / c_pad;
    const int c = r % Ci, tap = r / Ci;
    dwn[idx] = n < 3 ? dwall[(long)c * 32 + tap * 3 + n] : 0.f;
  }
}

// One warp per 30 output pixels of an image row: lane l owns SOURCE column j0 - 1 + l, reads the 32-byte window of each of the
// three source rows that holds that row's nine columns (two 16-byte loads, full sectors), sums over the rows, and the dx = -1 / +1
// taps arrive from the neighbouring lanes by shuffle.
__global__ void __launch_bounds__(256)
rgb_tapsum_fwd_kernel(const bf16* __restrict__ Y, const float* __restrict__ bias, bf16* __restrict__ out, int B, int H, int W, int Co,
                      int reflect, int act) {
  const int segs = (W + 29) / 30;
  const long nwarps = (long)B * H * segs;
  const int lane = threadIdx.x & 31;
  const long wstep = ((long)gridDim.x * blockDim.x) >> 5;
  for (long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < nwarps; wid += wstep) {
    const int seg = (int)(wid % segs);
    const long r = wid / segs;
    const int i = (int)(r % H);
    const int b = (int)(r / H);
    const int jj = seg * 30 - 1 + lane;
    int sj = jj;
    bool cok = jj >= 0 && jj < W;
    if (reflect) {
      if (jj == -1) { sj = 1; cok = true; }
      else if (jj == W) { sj = W - 2; cok = true; }
    }
    float s[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int n = 0; n < 3; ++n) s[c][n] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int si = i + a - 1;
      bool rok = si >= 0 && si < H;
      if (reflect) {
        if (si < 0) { si = 1; rok = true; }
        else if (si >= H) { si = H - 2; rok = true; }
      }
      if (rok && cok) {
        // columns 9a .. 9a+8 = bytes 18a .. 18a+17, inside the 32-byte window that starts at byte 16a (element offset a)
        const RgbU4* yp = reinterpret_cast<const RgbU4*>(Y + (((long)b * H + si) * W + sj) * 32) + a;
        const RgbU4 v0 = yp[0], v1 = yp[1];
        const uint32_t wd[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int n = 0; n < 3; ++n) {
            const int e = a + c * 3 + n;                      // compile-time after unrolling
            const uint32_t word = wd[e >> 1];
            s[c][n] += rgb_bits_to_float((e & 1) ? (word >> 16) : (word & 0xffffu));
          }
      }
    }
    float o[3];
#pragma unroll
    for (int n = 0; n < 3; ++n) {
      const float lft = __shfl_up_sync(0xffffffffu, s[0][n], 1);
      const float rgt = __shfl_down_sync(0xffffffffu, s[2][n], 1);
      o[n] = lft + s[1][n] + rgt;
    }
    if (lane >= 1 && lane <= 30 && jj < W) {
#pragma unroll
      for (int n = 0; n < 3; ++n) o[n] = rgb_act(o[n] + (bias ? bias[n] : 0.f), act);
      RgbU4* o4 = reinterpret_cast<RgbU4*>(out + (((long)b * H + i) * W + jj) * Co);
      RgbU4 first;
      first.x = rgb_float_to_bits(o[0]) | (rgb_float_to_bits(o[1]) << 16);
      first.y = rgb_float_to_bits(o[2]);
      first.z = 0u; first.w = 0u;
      o4[0] = first;
      RgbU4 zero;
      zero.x = zero.y = zero.z = zero.w = 0u;
      for (int q = 1; q < Co / 8; ++q) o4[q] = zero;
    }
  }
}

// Adjoint of the tap sum (and of the activation): one thread per pixel of dY.
__global__ void __launch_bounds__(256)
rgb_tapsum_bwd_kernel(const bf16* __restrict__ y, const bf16* __restrict__ dout, bf16* __restrict__ dY, int B, int H, int W, int Co,
                      int reflect, int act) {
  const long total = (long)B * H * W;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ix = (int)(idx % W);
    const long r = idx / W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    // padded coordinates R maps onto this pixel
    int py[3], px[3], ny = 0, nx = 0;
    py[ny++] = iy;
    px[nx++] = ix;
    if (reflect) {
      if (iy == 1) py[ny++] = -1;
      if (iy == H - 2) py[ny++] = H;
      if (ix == 1) px[nx++] = -1;
      if (ix == W - 2) px[nx++] = W;
    }
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    for (int u = 0; u < ny; ++u)
      for (int v = 0; v < nx; ++v) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const int sy = py[u] - (a - 1);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int sx = px[v] - (c - 1);
            if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
              const long off = (((long)b * H + sy) * W + sx) * Co;
              const RgbU2 g = *reinterpret_cast<const RgbU2*>(dout + off);
              float g0 = rgb_bits_to_float(g.x & 0xffffu), g1 = rgb_bits_to_float(g.x >> 16), g2 = rgb_bits_to_float(g.y & 0xffffu);
              if (act != BG_ACT_NONE) {
                const RgbU2 yy = *reinterpret_cast<const RgbU2*>(y + off);
                const float y0 = rgb_bits_to_float(yy.x & 0xffffu), y1 = rgb_bits_to_float(yy.x >> 16), y2 = rgb_bits_to_float(yy.y & 0xffffu);
                if (act == BG_ACT_TANH) { g0 *= 1.f - y0 * y0; g1 *= 1.f - y1 * y1; g2 *= 1.f - y2 * y2; }
                else { g0 = y0 > 0.f ? g0 : 0.f; g1 = y1 > 0.f ? g1 : 0.f; g2 = y2 > 0.f ? g2 : 0.f; }
              }
              acc[(a * 3 + c) * 3] += g0; acc[(a * 3 + c) * 3 + 1] += g1; acc[(a * 3 + c) * 3 + 2] += g2;
            }
          }
        }
      }
    RgbU4* zp = reinterpret_cast<RgbU4*>(dY + idx * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t wd[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k0 = 8 * q + 2 * e;
        wd[e] = rgb_float_to_bits(k0 < 27 ? acc[k0] : 0.f) | (rgb_float_to_bits(k0 + 1 < 27 ? acc[k0 + 1] : 0.f) << 16);
      }
      RgbU4 o;
      o.x = wd[0]; o.y = wd[1]; o.z = wd[2]; o.w = wd[3];
      zp[q] = o;
    }
  }
}
