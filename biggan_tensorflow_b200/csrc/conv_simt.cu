// SIMT (CUDA-core, fp32 accumulate) gather-GEMM kernels: the fp32 parity-mode implementation of
// conv / conv2d_transpose / their dgrad + wgrad (ops.py:49-139), and the fallback for shapes the
// tcgen05 path does not take (Cin = 3, Cout = 3, ...).  Same geometry abstraction as conv_tc.cu.
#include "bg_common.cuh"

#define GF_BM 128
#define GF_BN 64
#define GF_BK 16

// ------------------------------------------------------------------------------------------------
// forward-type: out[b, oy, ox, n] = act( sum_t sum_k in[b, iy_t, ix_t, k] * w[t, k, n] + bias + res )
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gather_gemm_fwd_kernel(const __grid_constant__ BgGeom g, const T* __restrict__ in,
                       const T* __restrict__ w, const float* __restrict__ bias,
                       const T* residual, T* out, int act, int tiles_per_phase, int kc, float* __restrict__ partial) {
  // kc > 0: split-K (skinny GEMMs, e.g. fully_connected with M = batch rows): this CTA contracts over
  // channels [blockIdx.z*kc, +kc) and writes an fp32 partial tile; splitk_fwd_reduce_kernel finishes.
  __shared__ float As[GF_BK][GF_BM + 4];
  __shared__ float Bs[GF_BK][GF_BN + 4];
  const int tid = threadIdx.x;
  const int phase = blockIdx.x / tiles_per_phase;
  const int mt = blockIdx.x % tiles_per_phase;
  const int n0 = blockIdx.y * GF_BN;
  const BgPhase& ph = g.ph[phase];
  const int HWp = ph.Hp * ph.Wp;
  const int Mph = g.B * HWp;
  if (mt * GF_BM >= Mph) return;

  // A-load role: pixel pa, 8 consecutive k
  const int pa = tid >> 1, ka = (tid & 1) * 8;
  const int ma = mt * GF_BM + pa;
  const bool mvalid = ma < Mph;
  int ab = 0, ai = 0, aj = 0;
  if (mvalid) { ab = ma / HWp; int r = ma - ab * HWp; ai = r / ph.Wp; aj = r - ai * ph.Wp; }
  const bool vecA = (g.Ci % 4 == 0);
  // B-load role
  const bool n_contig = (g.wn == 1);
  int kb, nb;
  if (n_contig) { kb = tid >> 4; nb = (tid & 15) * 4; } else { nb = tid >> 2; kb = (tid & 3) * 4; }

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < ph.ntaps; ++t) {
    const BgTap tap = ph.taps[t];
    int iy = ai * g.istep + tap.dy, ix = aj * g.istep + tap.dx;
    bool ok = mvalid;
    if (g.pad_mode == BG_PAD_REFLECT) {
      if (iy < 0) iy = -iy;
      if (iy >= g.Hi) iy = 2 * (g.Hi - 1) - iy;
      if (ix < 0) ix = -ix;
      if (ix >= g.Wi) ix = 2 * (g.Wi - 1) - ix;
    }
    ok = ok && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
    const long abase = (((long)ab * g.Hi + iy) * g.Wi + ix) * g.Ci;
    const T* wt = w + (long)tap.widx * g.wtap;
    const int k_lo = kc ? blockIdx.z * kc : 0, k_hi = kc ? min(g.Ci, k_lo + kc) : g.Ci;
    for (int k0 = k_lo; k0 < k_hi; k0 += GF_BK) {
      float ra[8], rb[4];
      if (vecA) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int k = k0 + ka + h * 4;
          f4 v = {0.f, 0.f, 0.f, 0.f};
          if (ok && k < k_hi) v = bg_ld4(in, abase + k);
          ra[h * 4 + 0] = v.x; ra[h * 4 + 1] = v.y; ra[h * 4 + 2] = v.z; ra[h * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          int k = k0 + ka + e;
          ra[e] = (ok && k < k_hi) ? bg_ld(in, abase + k) : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int k = k0 + kb + (n_contig ? 0 : e);
        int n = n0 + nb + (n_contig ? e : 0);
        rb[e] = (k < k_hi && n < g.Co) ? bg_ld(wt, (long)k * g.wk + (long)n * g.wn) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 8; ++e) As[ka + e][pa] = ra[e];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (n_contig) Bs[kb][nb + e] = rb[e]; else Bs[kb + e][nb] = rb[e];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < GF_BK; ++kk) {
        float a[8], b[4];
        float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
  // epilogue
  if (partial) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = mt * GF_BM + ty * 8 + i;
      if (m >= Mph) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < g.Co) partial[((long)blockIdx.z * Mph + m) * g.Co + n] = acc[i][j];
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = mt * GF_BM + ty * 8 + i;
    if (m >= Mph) continue;
    const int b = m / HWp; const int r = m - b * HWp; const int pi = r / ph.Wp; const int pj = r - pi * ph.Wp;
    const int oy = ph.oy0 + pi * g.ostep, ox = ph.ox0 + pj * g.ostep;
    const long obase = (((long)b * g.Ho + oy) * g.Wo + ox) * g.Co;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.Co) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (residual) v += bg_ld(residual, obase + n);
      bg_st(out, obase + n, bg_apply_act(v, act));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad-type: dw[t, k, n] = sum_pixels in[b, iy_t, ix_t, k] * og[b, oy, ox, n]   (one phase)
// partial results per split z go to part + z*zstride with the final (strided) dw layout.
// ------------------------------------------------------------------------------------------------
#define GW_BK 64
#define GW_BN 64
#define GW_BP 16

template <typename T>
__global__ void __launch_bounds__(256)
gather_gemm_wgrad_kernel(const __grid_constant__ BgGeom g, const T* __restrict__ in,
                         const T* __restrict__ og, float* part, long zstride, int nsplit, int ntile_n) {
  __shared__ float As[GW_BP][GW_BK + 4];
  __shared__ float Bs[GW_BP][GW_BN + 4];
  const int tid = threadIdx.x;
  const BgPhase& ph = g.ph[0];
  const BgTap tap = ph.taps[blockIdx.y];
  const int k0 = (blockIdx.x / ntile_n) * GW_BK, n0 = (blockIdx.x % ntile_n) * GW_BN;
  const int HWp = ph.Hp * ph.Wp;
  const int Mph = g.B * HWp;
  const int chunk = (int)(((long)Mph + nsplit - 1) / nsplit);
  const int mbeg = blockIdx.z * chunk;
  const int mend = min(Mph, mbeg + chunk);

  const int pl = tid >> 4, c4 = (tid & 15) * 4;     // load role: pixel row, 4 channels
  const int tx = tid & 15, ty = tid >> 4;           // compute role: 4 n x 4 k
  const bool vecA = (g.Ci % 4 == 0), vecB = (g.Co % 4 == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int m0 = mbeg; m0 < mend; m0 += GW_BP) {
    const int m = m0 + pl;
    float ra[4] = {0.f, 0.f, 0.f, 0.f}, rb[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      const int b = m / HWp; const int r = m - b * HWp; const int pi = r / ph.Wp; const int pj = r - pi * ph.Wp;
      int iy = pi * g.istep + tap.dy, ix = pj * g.istep + tap.dx;
      if (g.pad_mode == BG_PAD_REFLECT) {
        if (iy < 0) iy = -iy;
        if (iy >= g.Hi) iy = 2 * (g.Hi - 1) - iy;
        if (ix < 0) ix = -ix;
        if (ix >= g.Wi) ix = 2 * (g.Wi - 1) - ix;
      }
      if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) {
        const long abase = (((long)b * g.Hi + iy) * g.Wi + ix) * g.Ci;
        const int k = k0 + c4;
        if (vecA) { if (k < g.Ci) { f4 v = bg_ld4(in, abase + k); ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w; } }
        else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (k + e < g.Ci) ra[e] = bg_ld(in, abase + k + e);
        }
      }
      const int oy = ph.oy0 + pi * g.ostep, ox = ph.ox0 + pj * g.ostep;
      const long obase = (((long)b * g.Ho + oy) * g.Wo + ox) * g.Co;
      const int n = n0 + c4;
      if (vecB) { if (n < g.Co) { f4 v = bg_ld4(og, obase + n); rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w; } }
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (n + e < g.Co) rb[e] = bg_ld(og, obase + n + e);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[pl][c4]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
    *reinterpret_cast<float4*>(&Bs[pl][c4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    __syncthreads();
#pragma unroll
    for (int p = 0; p < GW_BP; ++p) {
      float4 a = *reinterpret_cast<const float4*>(&As[p][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[p][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* dst = part + (long)blockIdx.z * zstride + (long)tap.widx * g.wtap;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= g.Ci) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.Co) continue;
      dst[(long)k * g.wk + (long)n * g.wn] = acc[i][j];
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, long n, long zstride, int nsplit) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += part[(long)z * zstride + i];
    dw[i] = s;
  }
}

// adjoint of tf.pad(mode='REFLECT') (ops.py:82): fold the padded gradient back (+ residual).
// One thread per (pixel, VEC channels): 16-byte accesses when C % VEC == 0 (VEC = 8 bf16 / 4 fp32).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
reflect_fold_kernel(const T* __restrict__ dxp, const T* residual, const T* __restrict__ mask, T* dx, int B, int H, int W, int C, int pt, int pl,
                    int pb, int pr) {
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  const int CV = C / VEC;
  const long total = (long)B * H * W * CV;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV); const long pix = idx / CV;
    const int x = (int)(pix % W); const long r = pix / W;
    const int y = (int)(r % H); const int b = (int)(r / H);
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = y + pt;
    if (y >= 1 && y <= pt) ys[ny++] = pt - y;
    if (y >= H - 1 - pb && y <= H - 2) ys[ny++] = 2 * (H - 1) - y + pt;
    xs[nx++] = x + pl;
    if (x >= 1 && x <= pl) xs[nx++] = pl - x;
    if (x >= W - 1 - pr && x <= W - 2) xs[nx++] = 2 * (W - 1) - x + pl;
    float acc[VEC];
    const long o = pix * C + (long)cv * VEC;
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int q = 0; q < nx; ++q) {
        const long src = (((long)b * Hp + ys[a]) * Wp + xs[q]) * C + (long)cv * VEC;
#pragma unroll
        for (int e = 0; e < VEC; e += 4) { f4 v = bg_ld4(dxp, src + e); acc[e] += v.x; acc[e + 1] += v.y; acc[e + 2] += v.z; acc[e + 3] += v.w; }
      }
    if (mask) {                                      // ReLU mask of the conv's input: applies to the gradient, not to the residual
#pragma unroll
      for (int e = 0; e < VEC; e += 4) {
        f4 m = bg_ld4(mask, o + e);
        if (!(m.x > 0.f)) acc[e] = 0.f;
        if (!(m.y > 0.f)) acc[e + 1] = 0.f;
        if (!(m.z > 0.f)) acc[e + 2] = 0.f;
        if (!(m.w > 0.f)) acc[e + 3] = 0.f;
      }
    }
    if (residual) {
#pragma unroll
      for (int e = 0; e < VEC; e += 4) { f4 v = bg_ld4(residual, o + e); acc[e] += v.x; acc[e + 1] += v.y; acc[e + 2] += v.z; acc[e + 3] += v.w; }
    }
#pragma unroll
    for (int e = 0; e < VEC; e += 4) bg_st4(dx, o + e, f4{acc[e], acc[e + 1], acc[e + 2], acc[e + 3]});
  }
}

// Fold restricted to the band of target pixels that have pad-ring preimages (rows 1..pt, H-1-pb..H-2 and the matching
// columns): the tensor-core dgrad already wrote every interior preimage (+ residual) straight into dx; this adds the ring
// preimages kept in the padded-domain buffer.  Requires H >= pt + pb + 2 and W >= pl + pr + 2 (disjoint bands).
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
reflect_fold_band_kernel(const T* __restrict__ ring, const T* __restrict__ mask, T* dx, int B, int H, int W, int C, int pt, int pl, int pb,
                         int pr) {
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  const int CV = C / VEC;
  const int nR = pt + pb, nC = pl + pr;
  const long band = (long)nR * W + (long)(H - nR) * nC;
  const long total = (long)B * band * CV;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % CV); long k = idx / CV;
    const int b = (int)(k / band); k -= (long)b * band;
    int y, x;
    if (k < (long)nR * W) {
      const int ri = (int)(k / W);
      x = (int)(k - (long)ri * W);
      y = ri < pt ? 1 + ri : H - 1 - pb + (ri - pt);
    } else {
      k -= (long)nR * W;
      const int jj = (int)(k / nC), ci = (int)(k - (long)jj * nC);
      x = ci < pl ? 1 + ci : W - 1 - pr + (ci - pl);
      y = jj == 0 ? 0 : (jj == H - nR - 1 ? H - 1 : pt + jj);
    }
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = y + pt;
    if (y >= 1 && y <= pt) ys[ny++] = pt - y;
    if (y >= H - 1 - pb && y <= H - 2) ys[ny++] = 2 * (H - 1) - y + pt;
    xs[nx++] = x + pl;
    if (x >= 1 && x <= pl) xs[nx++] = pl - x;
    if (x >= W - 1 - pr && x <= W - 2) xs[nx++] = 2 * (W - 1) - x + pl;
    const long o = (((long)b * H + y) * W + x) * C + (long)cv * VEC;
    float acc[VEC], add[VEC];
#pragma unroll
    for (int e = 0; e < VEC; e += 4) { f4 v = bg_ld4(dx, o + e); acc[e] = v.x; acc[e + 1] = v.y; acc[e + 2] = v.z; acc[e + 3] = v.w; }
#pragma unroll
    for (int e = 0; e < VEC; ++e) add[e] = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int q = 0; q < nx; ++q) {
        if (a == 0 && q == 0) continue;             // the interior preimage is already in dx
        const long src = (((long)b * Hp + ys[a]) * Wp + xs[q]) * C + (long)cv * VEC;
#pragma unroll
        for (int e = 0; e < VEC; e += 4) { f4 v = bg_ld4(ring, src + e); add[e] += v.x; add[e + 1] += v.y; add[e + 2] += v.z; add[e + 3] += v.w; }
      }
    if (mask) {                                      // the interior preimage was masked by the dgrad epilogue; mask the ring's share too
#pragma unroll
      for (int e = 0; e < VEC; e += 4) {
        f4 m = bg_ld4(mask, o + e);
        if (!(m.x > 0.f)) add[e] = 0.f;
        if (!(m.y > 0.f)) add[e + 1] = 0.f;
        if (!(m.z > 0.f)) add[e + 2] = 0.f;
        if (!(m.w > 0.f)) add[e + 3] = 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] += add[e];
#pragma unroll
    for (int e = 0; e < VEC; e += 4) bg_st4(dx, o + e, f4{acc[e], acc[e + 1], acc[e + 2], acc[e + 3]});
  }
}

template <typename T>
__global__ void reflect_fold_scalar_kernel(const T* __restrict__ dxp, const T* residual, const T* __restrict__ mask, T* dx, int B, int H, int W,
                                           int C, int pt, int pl, int pb, int pr) {
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  const long total = (long)B * H * W * C;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C); long r = idx / C;
    const int x = (int)(r % W); r /= W;
    const int y = (int)(r % H); const int b = (int)(r / H);
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = y + pt;
    if (y >= 1 && y <= pt) ys[ny++] = pt - y;
    if (y >= H - 1 - pb && y <= H - 2) ys[ny++] = 2 * (H - 1) - y + pt;
    xs[nx++] = x + pl;
    if (x >= 1 && x <= pl) xs[nx++] = pl - x;
    if (x >= W - 1 - pr && x <= W - 2) xs[nx++] = 2 * (W - 1) - x + pl;
    float s = 0.f;
    for (int a = 0; a < ny; ++a)
      for (int e = 0; e < nx; ++e) s += bg_ld(dxp, (((long)b * Hp + ys[a]) * Wp + xs[e]) * C + c);
    if (mask && !(bg_ld(mask, idx) > 0.f)) s = 0.f;
    if (residual) s += bg_ld(residual, idx);
    bg_st(dx, idx, s);
  }
}

// One block per padded row (b, yp): the source row is resolved once, threads move 16-byte vectors (raw copy, no conversion).
template <int ACT>
__global__ void __launch_bounds__(256)
reflect_pad_rows_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int CV, int pt, int pl, int pb, int pr) {
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  const int yp = blockIdx.x % Hp, b = blockIdx.x / Hp;
  int iy = yp - pt;
  if (iy < 0) iy = -iy;
  if (iy >= H) iy = 2 * (H - 1) - iy;
  const uint4* src = x + ((long)b * H + iy) * W * CV;
  uint4* dst = y + (long)blockIdx.x * Wp * CV;
  const int n = Wp * CV;
  for (int v = threadIdx.x; v < n; v += 256) {
    const int xp = v / CV, cv = v - xp * CV;
    int ix = xp - pl;
    if (ix < 0) ix = -ix;
    if (ix >= W) ix = 2 * (W - 1) - ix;
    uint4 t = __ldg(src + (long)ix * CV + cv);
    if (ACT == BG_ACT_RELU) {                        // bf16 only: clear every negative element (sign bit set)
      uint32_t* w = reinterpret_cast<uint32_t*>(&t);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (w[e] & 0x00008000u) w[e] &= 0xFFFF0000u;
        if (w[e] & 0x80000000u) w[e] &= 0x0000FFFFu;
      }
    }
    dst[v] = t;
  }
}

template <typename T>
__global__ void reflect_pad_scalar_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int pt, int pl,
                                          int pb, int pr) {
  const int Hp = H + pt + pb, Wp = W + pl + pr;
  const long total = (long)B * Hp * Wp * C;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C); long r = idx / C;
    const int xp = (int)(r % Wp); r /= Wp;
    const int yp = (int)(r % Hp); const int b = (int)(r / Hp);
    int iy = yp - pt, ix = xp - pl;
    if (iy < 0) iy = -iy;
    if (iy >= H) iy = 2 * (H - 1) - iy;
    if (ix < 0) ix = -ix;
    if (ix >= W) ix = 2 * (W - 1) - ix;
    y[idx] = x[(((long)b * H + iy) * W + ix) * C + c];
  }
}

// column sums of [rows, C] -> partial[chunk][C] -> out[C]   (bias gradient)
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, float* __restrict__ part, long rows, int C, long rows_per_chunk) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  const long r0 = (long)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < C) for (long r = r0 + ty; r < r1; r += 8) s += bg_ld(x, r * C + c);
  __shared__ float sh[8][33];
  sh[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x & 31];
    part[(long)blockIdx.y * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, float* __restrict__ out, int nchunk, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  int i = 0;
  for (; i + 8 <= nchunk; i += 8) {          // 8 independent loads in flight, added in the same fixed order
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(part + (long)(i + k) * C + c);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
  }
  for (; i < nchunk; ++i) s += part[(long)i * C + c];
  out[c] = s;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int bg_tc_fwd_launch(const BgGeom& g, const void* in, int in_H, int in_W, int shift_y, int shift_x, const void* w_kmajor,
                     const float* bias, const void* residual, void* out, int act, void* sk_ws, size_t sk_bytes, const struct BgTcFold* fold,
                     cudaStream_t st, int* taken);
struct BgTcFold { int pt, pl, Hi, Wi; void* ring; const void* mask; };
size_t bg_tc_fwd_splitk_bytes(const BgGeom& g);
static inline size_t align256(size_t n) { return (n + 255) / 256 * 256; }
int bg_tc_conv_wgrad(const BgGeom& g, const bg_conv_desc* d, const void* in, const void* og, float* dw,
                     void* ws, size_t ws_bytes, cudaStream_t st, int* taken);
size_t bg_tc_conv_wgrad_ws(const BgGeom& g, const bg_conv_desc* d);

static bool reflect_padded(const bg_conv_desc* d) {
  return !d->transposed && d->pad_mode == BG_PAD_REFLECT && (d->pad_t | d->pad_b | d->pad_l | d->pad_r);
}
static size_t padded_bytes(const bg_conv_desc* d, size_t es) {
  return (size_t)d->B * (d->Hi + d->pad_t + d->pad_b) * (d->Wi + d->pad_l + d->pad_r) * d->Ci * es + 256;
}

int bg_small_fwd(const BgGeom& g, const void* in, const void* w, const float* bias, const void* residual, void* out, int act, int dtype,
                 cudaStream_t st, int* taken);
size_t bg_small_wgrad_ws(const BgGeom& g, long numel);
int bg_small_wgrad(const BgGeom& g, const void* in, const void* og, float* dw, long numel, int dtype, void* ws, size_t ws_bytes,
                   cudaStream_t st, int* taken);

// split-K factor for skinny single-tap GEMMs (fully_connected: M = batch rows): enough CTAs to fill the GPU
static int fwd_ksplit(const BgGeom& g) {
  if (g.nphase != 1 || g.ph[0].ntaps != 1 || g.ostep != 1 || g.ph[0].oy0 || g.ph[0].ox0) return 1;
  const long M = (long)g.B * g.ph[0].Hp * g.ph[0].Wp;
  if (M > 1024 || g.Ci < 256) return 1;
  const long tiles = bg_cdiv(M, GF_BM) * bg_cdiv(g.Co, GF_BN);
  long ks = bg_cdiv(2L * bg_num_sms(), tiles);
  ks = min(ks, (long)g.Ci / 64);
  ks = min(ks, 64L);
  return (int)max(ks, 1L);
}
static size_t fwd_ksplit_bytes(const BgGeom& g) {
  const int ks = fwd_ksplit(g);
  return ks > 1 ? (size_t)ks * g.B * g.ph[0].Hp * g.ph[0].Wp * g.Co * sizeof(float) : 0;
}

__global__ void splitk_fwd_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, const float* residual,
                                         float* out, long n, int Co, int nz, int act) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < nz; ++z) s += part[(long)z * n + i];
    if (bias) s += bias[i % Co];
    if (residual) s += residual[i];
    out[i] = bg_apply_act(s, act);
  }
}

template <typename T>
static int launch_fwd_type(const BgGeom& g, const T* in, const T* w, const float* bias, const T* residual, T* out,
                           int act, cudaStream_t st, void* ws = nullptr, size_t ws_bytes = 0) {
  if (g.Co <= 4 || g.Ci <= 4) {      // RGB layers: memory-bound direct kernels (conv_small.cu)
    int taken = 0;
    int rc = bg_small_fwd(g, in, w, bias, residual, out, act, sizeof(T) == 4 ? BG_F32 : BG_BF16, st, &taken);
    if (rc) return rc;
    if (taken) return BG_OK;
  }
  long maxM = 0;
  for (int p = 0; p < g.nphase; ++p) maxM = max(maxM, (long)g.B * g.ph[p].Hp * g.ph[p].Wp);
  const int tiles = (int)bg_cdiv(maxM, GF_BM);
  const int ks = fwd_ksplit(g);
  if (ks > 1 && sizeof(T) == 4 && ws && ws_bytes >= (size_t)ks * maxM * g.Co * sizeof(float)) {
    const int kc = (int)(bg_cdiv(bg_cdiv(g.Ci, ks), 16) * 16);
    const int nz = (int)bg_cdiv(g.Ci, kc);
    dim3 grid(tiles, (unsigned)bg_cdiv(g.Co, GF_BN), nz);
    gather_gemm_fwd_kernel<T><<<grid, 256, 0, st>>>(g, in, w, nullptr, nullptr, out, 0, tiles, kc, (float*)ws);
    BG_LAUNCH_CHECK();
    const long n = maxM * g.Co;
    splitk_fwd_reduce_kernel<<<(unsigned)min(bg_cdiv(n, 256), (long)bg_num_sms() * 8), 256, 0, st>>>(
        (const float*)ws, bias, (const float*)residual, (float*)out, n, g.Co, nz, act);
    BG_LAUNCH_CHECK();
    return BG_OK;
  }
  dim3 grid(tiles * g.nphase, (unsigned)bg_cdiv(g.Co, GF_BN));
  gather_gemm_fwd_kernel<T><<<grid, 256, 0, st>>>(g, in, w, bias, residual, out, act, tiles, 0, nullptr);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// fc.cu
int bg_fc_fwd(const float* x, const float* w, const float* bias, float* y, int M, int K, int N, int act, cudaStream_t st);
int bg_fc_wgrad(const float* x, const float* dy, float* dw, int M, int K, int N, cudaStream_t st);
static inline bool fc_shape(const bg_conv_desc* d) {
  return !d->transposed && d->Hi == 1 && d->Wi == 1 && d->Ho == 1 && d->Wo == 1 && d->kh == 1 && d->kw == 1;
}

extern "C" size_t bg_conv2d_fwd_workspace_bytes(const bg_conv_desc* d, int dtype) {
  BgGeom g;
  if (bg_build_geom(d, BG_OP_FWD, &g) != BG_OK) return 0;
  if (dtype == BG_F32) return fwd_ksplit_bytes(g);
  // bf16: [reflect-padded copy (only when the caller did not pre-pad)][tensor-core split-K partials]
  return (reflect_padded(d) ? align256(padded_bytes(d, 2)) : 0) + bg_tc_fwd_splitk_bytes(g);
}

extern "C" int bg_conv2d_fwd(const bg_conv_desc* d, const void* x, const void* w, const void* w_t, const float* bias,
                             const void* residual, void* y, int act, int dtype, void* ws, size_t ws_bytes, int force_simt,
                             void* stream) {
  BgGeom g;
  int rc = bg_build_geom(d, BG_OP_FWD, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!force_simt && dtype == BG_BF16 && bg_has_tensor_core_path()) {
    // K-major B operand [tap][n][k]: transposed conv has it natively, plain conv needs w_t
    const void* wk = d->transposed ? w : w_t;
    const bool pad = reflect_padded(d);
    if (wk && (!pad || (ws && ws_bytes >= padded_bytes(d, 2)))) {
      const void* in = x;
      int inH = d->Hi, inW = d->Wi, sy = 0, sx = 0;
      char* sk = (char*)ws;
      size_t sk_bytes = ws ? ws_bytes : 0;
      if (pad && g.Ci % 8 == 0 && g.Ci >= 16 && g.Co % 16 == 0) {
        rc = bg_reflect_pad(x, ws, d->B, d->Hi, d->Wi, d->Ci, d->pad_t, d->pad_l, d->pad_b, d->pad_r, dtype, stream);
        if (rc) return rc;
        in = ws; inH += d->pad_t + d->pad_b; inW += d->pad_l + d->pad_r; sy = d->pad_t; sx = d->pad_l;
      }
      if (pad) {
        const size_t used = align256(padded_bytes(d, 2));
        sk = (sk_bytes > used) ? sk + used : nullptr;
        sk_bytes = (sk_bytes > used) ? sk_bytes - used : 0;
      }
      if (!pad || in == ws) {
        int taken = 0;
        rc = bg_tc_fwd_launch(g, in, inH, inW, sy, sx, wk, bias, residual, y, act, sk, sk_bytes, nullptr, st, &taken);
        if (rc) return rc;
        if (taken) return BG_OK;
      }
    }
  }
  // fully_connected: dedicated skinny-GEMM kernel (fc.cu) when its grid fills the GPU and the (serial) contraction is short
  if (dtype == BG_F32 && fc_shape(d) && !residual && d->Co >= 64 && d->Ci <= 512 && bg_cdiv(d->Co, 128) * bg_cdiv(d->B, 16) >= 48)
    return bg_fc_fwd((const float*)x, (const float*)w, bias, (float*)y, d->B, d->Ci, d->Co, act, st);
  BG_DISPATCH_DTYPE(dtype, T, return launch_fwd_type<T>(g, (const T*)x, (const T*)w, bias, (const T*)residual, (T*)y, act, st, ws, ws_bytes));
}

#include "fold_kernels.cuh"     // reflect_dyfold_kernel, reflect_dyfold4_kernel (per-tap folded dy of the small reflect-padded maps)
static bool dgrad_dyfold(const bg_conv_desc* d, int dtype) {
  static int mh = -1;                               // A/B knob: largest map side taking this path (0 = off)
  if (mh < 0) { const char* e = getenv("BG_TC_DYFOLD_MAXH"); mh = e ? atoi(e) : 4; }
  return dtype == BG_BF16 && !d->transposed && d->pad_mode == BG_PAD_REFLECT && d->kh == 3 && d->kw == 3 && d->stride == 1 && d->pad_t == 1 &&
         d->pad_b == 1 && d->pad_l == 1 && d->pad_r == 1 && d->Hi >= 2 && d->Wi >= 2 && d->Hi <= mh && d->Wi <= mh && d->Co % 8 == 0 &&
         d->Ci % 16 == 0 && d->Co >= 16 && 9 * d->Hi < 32768 && bg_has_tensor_core_path();
}
static size_t dgrad_dyfold_bytes(const bg_conv_desc* d) { return align256((size_t)d->B * 9 * d->Ho * d->Wo * d->Co * 2 + 256); }

static bool dgrad_padded(const bg_conv_desc* d) {
  return !d->transposed && d->pad_mode == BG_PAD_REFLECT && (d->pad_t | d->pad_b | d->pad_l | d->pad_r);
}

static size_t dgrad_padded_bytes(const bg_conv_desc* d, int dtype) {
  size_t es = dtype == BG_F32 ? 4 : 2;
  return align256((size_t)d->B * (d->Hi + d->pad_t + d->pad_b) * (d->Wi + d->pad_l + d->pad_r) * d->Ci * es + 256);
}
extern "C" size_t bg_conv2d_dgrad_workspace_bytes(const bg_conv_desc* d, int dtype) {
  BgGeom g;
  const bool geom_ok = bg_build_geom(d, BG_OP_DGRAD, &g) == BG_OK;
  // [padded-domain gradient (reflect convs)][tensor-core split-K partials (bf16) | SIMT split-K partials (fp32, unpadded)]
  size_t n = dgrad_padded(d) ? dgrad_padded_bytes(d, dtype) : 0;
  if (dgrad_dyfold(d, dtype) && dgrad_dyfold_bytes(d) > n) n = dgrad_dyfold_bytes(d);     // (the two layouts share the front of the workspace)
  if (geom_ok && dtype == BG_BF16) n += bg_tc_fwd_splitk_bytes(g);
  if (geom_ok && dtype == BG_F32 && !dgrad_padded(d)) n += fwd_ksplit_bytes(g);
  return n;
}

extern "C" int bg_conv2d_dgrad(const bg_conv_desc* d, const void* dy, const void* w, const void* w_t, const void* residual,
                               const void* relu_mask, void* dx, int dtype, void* ws, size_t ws_bytes, int force_simt, void* stream) {
  BgGeom g;
  int rc = bg_build_geom(d, BG_OP_DGRAD, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool padded = dgrad_padded(d);
  // relu_mask: reflect-padded convs (the fold pass applies it) or, without padding, the tensor-core kernel's epilogue (bf16)
  BG_REQUIRE(!relu_mask || padded || (dtype == BG_BF16 && !force_simt && !d->transposed && !(d->pad_t | d->pad_l | d->pad_b | d->pad_r)),
             BG_ERR_UNSUPPORTED, "bg_conv2d_dgrad: relu_mask needs a reflect-padded conv, or an unpadded conv on the bf16 tensor-core path");
  if (padded) BG_REQUIRE(ws && ws_bytes >= dgrad_padded_bytes(d, dtype), BG_ERR_WORKSPACE, "dgrad workspace too small");
  void* gout = padded ? ws : dx;
  const void* gres = padded ? nullptr : residual;
  char* sk = (char*)ws;
  size_t sk_bytes = ws ? ws_bytes : 0;
  if (padded) {
    const size_t used = dgrad_padded_bytes(d, dtype);
    sk = (sk_bytes > used) ? sk + used : nullptr;
    sk_bytes = (sk_bytes > used) ? sk_bytes - used : 0;
  }
  bool done = false;
  if (!force_simt && dtype == BG_BF16) {
    const void* wk = d->transposed ? w_t : w;     // dgrad contracts over Co: HWIO [tap][ci][co] is K-major natively
    if (wk) {
      int taken = 0;
      if (dgrad_dyfold(d, dtype) && ws && ws_bytes >= dgrad_dyfold_bytes(d) && ((((uintptr_t)ws) | ((uintptr_t)dy)) & 15) == 0) {
        const long tot = (long)d->B * 9 * d->Ho * d->Wo * (d->Co / 8);
        if (d->Ho == 4 && d->Wo == 4)
          reflect_dyfold4_kernel<<<(int)min(bg_cdiv((long)d->B * (d->Co / 8), 128), (long)bg_num_sms() * 16), 128, 0, st>>>((const bf16*)dy, (bf16*)ws, d->B,
                                                                                                                     d->Co);
        else
          reflect_dyfold_kernel<<<(int)min(bg_cdiv(tot, 256), (long)bg_num_sms() * 16), 256, 0, st>>>((const bf16*)dy, (bf16*)ws, d->B, d->Ho, d->Wo, d->Co);
        BG_LAUNCH_CHECK();
        BgGeom g2;
        memset(&g2, 0, sizeof(g2));
        g2.B = d->B; g2.Hi = 9 * d->Ho; g2.Wi = d->Wo; g2.Ci = d->Co;
        g2.Ho = d->Hi; g2.Wo = d->Wi; g2.Co = d->Ci;
        g2.ostep = 1; g2.istep = 1; g2.pad_mode = BG_PAD_ZERO; g2.nphase = 1;
        g2.wtap = g.wtap; g2.wk = g.wk; g2.wn = g.wn;
        g2.ph[0].oy0 = 0; g2.ph[0].ox0 = 0; g2.ph[0].Hp = d->Hi; g2.ph[0].Wp = d->Wi; g2.ph[0].ntaps = 9;
        for (int t = 0; t < 9; ++t) { g2.ph[0].taps[t].dy = (short)(t * d->Ho); g2.ph[0].taps[t].dx = 0; g2.ph[0].taps[t].widx = t; }
        BgTcFold mo = {0, 0, 0, 0, nullptr, relu_mask};
        rc = bg_tc_fwd_launch(g2, ws, g2.Hi, g2.Wi, 0, 0, wk, nullptr, residual, dx, BG_ACT_NONE, nullptr, 0, relu_mask ? &mo : nullptr, st, &taken);
        if (rc) return rc;
        if (taken) return BG_OK;
      }
      static int fold_env = -1;
      if (fold_env < 0) { const char* e = getenv("BG_TC_FOLD"); fold_env = e ? atoi(e) : 1; }
      // reflect-padded conv: interior straight into dx (+ residual), pad ring into the padded buffer, band fold afterwards
      const bool band = fold_env && padded && d->Ci % 8 == 0 && d->Hi >= d->pad_t + d->pad_b + 2 && d->Wi >= d->pad_l + d->pad_r + 2 &&
                        bg_tc_fwd_splitk_bytes(g) == 0;
      if (band) {
        BgTcFold fold = {d->pad_t, d->pad_l, d->Hi, d->Wi, ws};
        fold.mask = relu_mask;
        rc = bg_tc_fwd_launch(g, dy, g.Hi, g.Wi, 0, 0, wk, nullptr, residual, dx, BG_ACT_NONE, nullptr, 0, &fold, st, &taken);
        if (rc) return rc;
        if (taken) {
          const long total = (long)d->B * ((long)(d->pad_t + d->pad_b) * d->Wi + (long)(d->Hi - d->pad_t - d->pad_b) * (d->pad_l + d->pad_r)) *
                             (d->Ci / 8);
          const int blocks = (int)min(bg_cdiv(total, 256), (long)bg_num_sms() * 16);
          reflect_fold_band_kernel<bf16, 8><<<blocks, 256, 0, st>>>((const bf16*)ws, (const bf16*)relu_mask, (bf16*)dx, d->B, d->Hi, d->Wi, d->Ci,
                                                                  d->pad_t, d->pad_l, d->pad_b, d->pad_r);
          BG_LAUNCH_CHECK();
          return BG_OK;
        }
      }
      BgTcFold mask_only = {0, 0, 0, 0, nullptr, relu_mask};
      rc = bg_tc_fwd_launch(g, dy, g.Hi, g.Wi, 0, 0, wk, nullptr, gres, gout, BG_ACT_NONE, sk, sk_bytes, (relu_mask && !padded) ? &mask_only : nullptr,
                            st, &taken);
      if (rc) return rc;
      done = taken != 0;
    }
  }
  BG_REQUIRE(done || !relu_mask || padded, BG_ERR_UNSUPPORTED, "bg_conv2d_dgrad: relu_mask on an unpadded conv was not taken by the tensor-core kernel");
  if (!done) {
    BG_DISPATCH_DTYPE(dtype, T, rc = launch_fwd_type<T>(g, (const T*)dy, (const T*)w, nullptr, (const T*)gres, (T*)gout, BG_ACT_NONE, st,
                                                        padded ? nullptr : ws, padded ? 0 : ws_bytes));
    if (rc) return rc;
  }
  if (padded) {
    const long total = (long)d->B * d->Hi * d->Wi * d->Ci;
    const int vec = dtype == BG_F32 ? 4 : 8;
    if (d->Ci % vec == 0) {
      const int blocks = (int)min(bg_cdiv(total / vec, 256), (long)bg_num_sms() * 32);
      if (dtype == BG_F32)
        reflect_fold_kernel<float, 4><<<blocks, 256, 0, st>>>((const float*)ws, (const float*)residual, (const float*)relu_mask, (float*)dx, d->B,
                                                            d->Hi, d->Wi, d->Ci, d->pad_t, d->pad_l, d->pad_b, d->pad_r);
      else
        reflect_fold_kernel<bf16, 8><<<blocks, 256, 0, st>>>((const bf16*)ws, (const bf16*)residual, (const bf16*)relu_mask, (bf16*)dx, d->B,
                                                           d->Hi, d->Wi, d->Ci, d->pad_t, d->pad_l, d->pad_b, d->pad_r);
    } else {
      const int blocks = (int)min(bg_cdiv(total, 256), (long)bg_num_sms() * 16);
      BG_DISPATCH_DTYPE(dtype, T, reflect_fold_scalar_kernel<T><<<blocks, 256, 0, st>>>((const T*)ws, (const T*)residual, (const T*)relu_mask, (T*)dx,
                                                                                     d->B, d->Hi, d->Wi, d->Ci, d->pad_t, d->pad_l, d->pad_b,
                                                                                     d->pad_r));
    }
    BG_LAUNCH_CHECK();
  }
  return BG_OK;
}

static int wgrad_nsplit(const BgGeom& g) {
  const BgPhase& ph = g.ph[0];
  const long M = (long)g.B * ph.Hp * ph.Wp;
  const long tiles = (long)ph.ntaps * bg_cdiv(g.Ci, GW_BK) * bg_cdiv(g.Co, GW_BN);
  long ns = bg_cdiv(4L * bg_num_sms(), tiles);
  ns = min(ns, bg_cdiv(M, 256));
  ns = max(1L, min(ns, 64L));
  return (int)ns;
}

static long wgrad_numel(const bg_conv_desc* d) { return (long)d->kh * d->kw * d->Ci * d->Co; }

extern "C" size_t bg_conv2d_wgrad_workspace_bytes(const bg_conv_desc* d, int dtype) {
  BgGeom g;
  if (bg_build_geom(d, BG_OP_WGRAD, &g)) return 0;
  size_t simt = (size_t)wgrad_nsplit(g) * wgrad_numel(d) * sizeof(float);
  size_t tc = dtype == BG_BF16 ? bg_tc_conv_wgrad_ws(g, d) : 0;
  size_t sm = bg_small_wgrad_ws(g, wgrad_numel(d));
  if (sm > simt) simt = sm;
  return (simt > tc ? simt : tc) + 256;
}

extern "C" int bg_conv2d_wgrad(const bg_conv_desc* d, const void* x, const void* dy, float* dw, int dtype, void* ws,
                               size_t ws_bytes, int force_simt, void* stream) {
  BgGeom g;
  int rc = bg_build_geom(d, BG_OP_WGRAD, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const void* in = d->transposed ? dy : x;
  const void* og = d->transposed ? x : dy;
  if (!force_simt && dtype == BG_BF16) {
    int taken = 0;
    rc = bg_tc_conv_wgrad(g, d, in, og, dw, ws, ws_bytes, st, &taken);
    if (rc) return rc;
    if (taken) return BG_OK;
  }
  const long numel = wgrad_numel(d);
  if (dtype == BG_F32 && fc_shape(d) && d->Co >= 64 && d->B <= 512 && bg_cdiv(d->Co, 128) * bg_cdiv(d->Ci, 16) >= 48)
    return bg_fc_wgrad((const float*)x, (const float*)dy, dw, d->B, d->Ci, d->Co, st);
  if (g.Co <= 4 || g.Ci <= 4) {
    int taken = 0;
    rc = bg_small_wgrad(g, in, og, dw, numel, dtype, ws, ws_bytes, st, &taken);
    if (rc) return rc;
    if (taken) return BG_OK;
  }
  const int ns = wgrad_nsplit(g);
  float* part = dw;
  if (ns > 1) {
    BG_REQUIRE(ws && ws_bytes >= (size_t)ns * numel * sizeof(float), BG_ERR_WORKSPACE, "wgrad workspace too small");
    part = (float*)ws;
  }
  const int ntn = (int)bg_cdiv(g.Co, GW_BN);
  dim3 grid((unsigned)(bg_cdiv(g.Ci, GW_BK) * ntn), g.ph[0].ntaps, ns);
  BG_DISPATCH_DTYPE(dtype, T, gather_gemm_wgrad_kernel<T><<<grid, 256, 0, st>>>(g, (const T*)in, (const T*)og, part, numel, ns, ntn));
  BG_LAUNCH_CHECK();
  if (ns > 1) {
    const int blocks = (int)min(bg_cdiv(numel, 256), (long)bg_num_sms() * 8);
    splitk_reduce_kernel<<<blocks, 256, 0, st>>>(part, dw, numel, numel, ns);
    BG_LAUNCH_CHECK();
  }
  return BG_OK;
}

size_t bg_bn_colsum_ws(long rows, int C);            // norm.cu: column sums on the vectorised statistics kernels
bool bg_bn_colsum_applicable(int C, int dtype);
int bg_bn_colsum(const void* x, float* out, long rows, int C, int dtype, void* ws, cudaStream_t st);

extern "C" size_t bg_colsum_workspace_bytes(long rows, int C) {
  long nchunk = min(bg_cdiv(rows, 64), 1024L);
  const size_t a = (size_t)nchunk * C * sizeof(float) + 256, b = (C % 4 == 0) ? bg_bn_colsum_ws(rows, C) : 0;
  return a > b ? a : b;
}

extern "C" int bg_colsum(const void* dy, float* db, long rows, int C, int dtype, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (rows >= 256 && bg_bn_colsum_applicable(C, dtype) && ws && ws_bytes >= bg_bn_colsum_ws(rows, C) && (((uintptr_t)dy) & 15) == 0)
    return bg_bn_colsum(dy, db, rows, C, dtype, ws, st);
  long nchunk = min(bg_cdiv(rows, 64), 1024L);
  BG_REQUIRE(ws && ws_bytes >= (size_t)nchunk * C * sizeof(float), BG_ERR_WORKSPACE, "colsum workspace too small");
  const long rpc = bg_cdiv(rows, nchunk);
  dim3 grid((unsigned)bg_cdiv(C, 32), (unsigned)nchunk);
  BG_DISPATCH_DTYPE(dtype, T, colsum_partial_kernel<T><<<grid, 256, 0, st>>>((const T*)dy, (float*)ws, rows, C, rpc));
  BG_LAUNCH_CHECK();
  colsum_final_kernel<<<(unsigned)bg_cdiv(C, 128), 128, 0, st>>>((const float*)ws, db, (int)nchunk, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// y = relu(reflect_pad(x)) in one pass (bf16, 16-byte vectors): the `act(x)` in front of the first conv of resblock_down
// (ops.py:297-301) rides in the padding copy that conv needs anyway
extern "C" int bg_reflect_pad_relu(const void* x, void* y, int B, int H, int W, int C, int pt, int pl, int pb, int pr, int dtype,
                                   void* stream) {
  BG_REQUIRE(dtype == BG_BF16 && C % 8 == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && (long)B * (H + pt + pb) < 2147483647L,
             BG_ERR_UNSUPPORTED, "bg_reflect_pad_relu: bf16, C a multiple of 8, 16-byte aligned");
  reflect_pad_rows_kernel<BG_ACT_RELU><<<(unsigned)(B * (H + pt + pb)), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, H, W, C / 8, pt, pl,
                                                                                                    pb, pr);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_reflect_pad(const void* x, void* y, int B, int H, int W, int C, int pt, int pl, int pb, int pr, int dtype,
                              void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long total = (long)B * (H + pt + pb) * (W + pl + pr) * C;
  const int vec = dtype == BG_F32 ? 4 : 8;
  if (C % vec == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0 && (long)B * (H + pt + pb) < 2147483647L) {
    reflect_pad_rows_kernel<BG_ACT_NONE><<<(unsigned)(B * (H + pt + pb)), 256, 0, st>>>((const uint4*)x, (uint4*)y, H, W, C / vec, pt, pl, pb, pr);
  } else {
    const int blocks = (int)min(bg_cdiv(total, 256), (long)bg_num_sms() * 16);
    BG_DISPATCH_DTYPE(dtype, T, reflect_pad_scalar_kernel<T><<<blocks, 256, 0, st>>>((const T*)x, (T*)y, B, H, W, C, pt, pl, pb, pr));
  }
  BG_LAUNCH_CHECK();
  return BG_OK;
}
