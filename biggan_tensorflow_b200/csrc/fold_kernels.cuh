// Device kernels of the small-map dgrad path of conv_simt.cu, kept free of includes and launch syntax so that the same source also
// compiles for the host under tests/host_emul/cuda_host_emul.h (CPU execution check: tests/test_fold_kernels_host_emul.py).
// Include bg_common.cuh (or the host emulation prelude) first.
#pragma once

// Small reflect-padded maps (3x3 / stride 1 / pad 1, H <= 4): the padded-domain dgrad computes (H+2)^2 / H^2 = 2.25x the pixels at 4x4.
// The fold commutes with the per-pixel weight contraction, so it is applied to dy instead, once per tap:
//   dyf[b, t*H + i, j, :] = sum over padded coordinates (py, px) that reflect onto (i, j) of dy[b, py - (ky-1), px - (kx-1), :]   (t = ky*3 + kx)
// and dx = sum_t dyf_t W_t^T is a plain 9-"tap" GEMM over the UNPADDED domain (tap t reads rows t*H .. t*H+H-1 of the stacked tensor; bulk
// stores, ReLU mask and gradient accumulation in the ordinary epilogue, no pad ring, no fold pass).  dyf costs 9x dy, which at 4x4 is small.
__global__ void __launch_bounds__(256)
reflect_dyfold_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dyf, int B, int H, int W, int C) {
  const int CV = C / 8;
  const long total = (long)B * 9 * H * W * CV;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV); long k = idx / CV;
    const int j = (int)(k % W); k /= W;
    const int i = (int)(k % H); k /= H;
    const int t = (int)(k % 9); const int b = (int)(k / 9);
    const int oy = t / 3 - 1, ox = t % 3 - 1;
    int py[3], px[3], ny = 0, nx = 0;
    py[ny++] = i;
    if (i == 1) py[ny++] = -1;
    if (i == H - 2) py[ny++] = H;
    px[nx++] = j;
    if (j == 1) px[nx++] = -1;
    if (j == W - 2) px[nx++] = W;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int u = 0; u < ny; ++u) {
      const int sy = py[u] - oy;
      if (sy < 0 || sy >= H) continue;
      for (int v = 0; v < nx; ++v) {
        const int sx = px[v] - ox;
        if (sx < 0 || sx >= W) continue;
        float r[8];
        Vec<bf16>::ld(dy, ((((long)b * H + sy) * W + sx) * C) + (long)cv * 8, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += r[e];
      }
    }
    Vec<bf16>::st(dyf, ((((long)b * 9 * H + (long)t * H + i) * W + j) * C) + (long)cv * 8, acc);
  }
}
// 4x4 maps (the only size on the BigGAN path): one thread owns 8 channels of one image -- the 16 pixels are loaded once into registers
// and all 9 x 16 folded pixels are formed with compile-time index arithmetic (the generic kernel spent its time in runtime divisions:
// 123 us per launch for 226 MB in the ncu launch list)
__global__ void __launch_bounds__(128)
reflect_dyfold4_kernel(const bf16* __restrict__ dy, bf16* __restrict__ dyf, int B, int C) {
  const int CV = C / 8;
  const long total = (long)B * CV;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV); const int b = (int)(idx / CV);
    float v[16][8];
#pragma unroll
    for (int p = 0; p < 16; ++p) Vec<bf16>::ld(dy, ((long)b * 16 + p) * C + (long)cv * 8, v[p]);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int oy = t / 3 - 1, ox = t % 3 - 1;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
          // padded rows that reflect onto i: i itself, -1 when i == 1, 4 when i == 2 (same for columns)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int py = u == 0 ? i : (i == 1 ? -1 : (i == 2 ? 4 : -9));
            const int sy = py - oy;
            if (py == -9 || sy < 0 || sy > 3) continue;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const int px = w == 0 ? j : (j == 1 ? -1 : (j == 2 ? 4 : -9));
              const int sx = px - ox;
              if (px == -9 || sx < 0 || sx > 3) continue;
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[e] += v[sy * 4 + sx][e];
            }
          }
          Vec<bf16>::st(dyf, (((long)b * 36 + t * 4 + i) * 4 + j) * C + (long)cv * 8, acc);
        }
    }
  }
}
