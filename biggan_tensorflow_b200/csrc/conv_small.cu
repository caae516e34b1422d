// Memory-bound convolutions with a tiny channel count on one side (RGB layers): the first
// discriminator convs (3 -> ch, ops.py:302/310 via BigGAN.py:633), `G_logit` (ch -> 3, BigGAN.py:570)
// and their dgrad / wgrad.  These are not GEMM-shaped (N or K = 3): one thread per output pixel (or per
// pixel x 8 channels), weights in shared memory, activations streamed once with 16-byte loads.
// Same phase/tap geometry as the other conv kernels (bg_common.cuh).
#include "bg_common.cuh"

#define SMALL_MAX_W 10240   // floats of shared memory for the weights

__device__ __forceinline__ float small_act(float v, int act) {
  if (act == BG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == BG_ACT_TANH) return tanhf(v);
  return v;
}

__device__ __forceinline__ bool small_src(const BgGeom& g, int i, int j, const BgTap& tap, int* iy, int* ix) {
  int y = i * g.istep + tap.dy, x = j * g.istep + tap.dx;
  if (g.pad_mode == BG_PAD_REFLECT) {
    if (y < 0) y = -y;
    if (y >= g.Hi) y = 2 * (g.Hi - 1) - y;
    if (x < 0) x = -x;
    if (x >= g.Wi) x = 2 * (g.Wi - 1) - x;
  }
  *iy = y; *ix = x;
  return y >= 0 && y < g.Hi && x >= 0 && x < g.Wi;
}

// ---- N = g.Co <= 4: one thread per output pixel, all (<=4) output channels ----------------------
template <typename T>
__global__ void __launch_bounds__(256)
small_n_fwd_kernel(const __grid_constant__ BgGeom g, const T* __restrict__ in, const T* __restrict__ w,
                   const float* __restrict__ bias, const T* residual, T* out, int act, int tiles_per_phase) {
  extern __shared__ float wsm[];     // [tap][k][4]
  const int phase = blockIdx.x / tiles_per_phase;
  const int tile = blockIdx.x % tiles_per_phase;
  const BgPhase& ph = g.ph[phase];
  for (int idx = threadIdx.x; idx < ph.ntaps * g.Ci * 4; idx += blockDim.x) {
    const int n = idx & 3, k = (idx >> 2) % g.Ci, t = (idx >> 2) / g.Ci;
    wsm[idx] = (n < g.Co) ? bg_ld(w, (long)ph.taps[t].widx * g.wtap + (long)k * g.wk + (long)n * g.wn) : 0.f;
  }
  __syncthreads();
  const int HWp = ph.Hp * ph.Wp;
  const long m = (long)tile * blockDim.x + threadIdx.x;
  if (m >= (long)g.B * HWp) return;
  const int b = (int)(m / HWp); const int r = (int)(m - (long)b * HWp); const int i = r / ph.Wp, j = r - i * ph.Wp;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = (g.Ci % 8 == 0);
  for (int t = 0; t < ph.ntaps; ++t) {
    int iy, ix;
    if (!small_src(g, i, j, ph.taps[t], &iy, &ix)) continue;
    const long base = (((long)b * g.Hi + iy) * g.Wi + ix) * g.Ci;
    const float* wt = wsm + (long)t * g.Ci * 4;
    if (vec) {
      for (int k = 0; k < g.Ci; k += 8) {
        f4 a = bg_ld4(in, base + k), c = bg_ld4(in, base + k + 4);
        const float xv[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 w4 = *reinterpret_cast<const float4*>(wt + (k + e) * 4);
          acc[0] = fmaf(xv[e], w4.x, acc[0]); acc[1] = fmaf(xv[e], w4.y, acc[1]);
          acc[2] = fmaf(xv[e], w4.z, acc[2]); acc[3] = fmaf(xv[e], w4.w, acc[3]);
        }
      }
    } else {
      for (int k = 0; k < g.Ci; ++k) {
        const float xv = bg_ld(in, base + k);
        const float4 w4 = *reinterpret_cast<const float4*>(wt + k * 4);
        acc[0] = fmaf(xv, w4.x, acc[0]); acc[1] = fmaf(xv, w4.y, acc[1]);
        acc[2] = fmaf(xv, w4.z, acc[2]); acc[3] = fmaf(xv, w4.w, acc[3]);
      }
    }
  }
  const int oy = ph.oy0 + i * g.ostep, ox = ph.ox0 + j * g.ostep;
  const long obase = (((long)b * g.Ho + oy) * g.Wo + ox) * g.Co;
  for (int n = 0; n < g.Co; ++n) {
    float v = acc[n];
    if (bias) v += bias[n];
    if (residual) v += bg_ld(residual, obase + n);
    bg_st(out, obase + n, small_act(v, act));
  }
}

// ---- N <= 4 on a 1x1 map (fully_connected to <= 4 units: D_logit, BigGAN.py:687): one WARP per sample -------------------
// The thread-per-pixel kernel above put a whole 1536-long dot product into one thread with 6 KB-strided loads: 2 CTAs, 150 us
// for 3 MB (ncu launch list).  Here the lanes stride over K (coalesced), then a butterfly sum (fixed order: deterministic).
template <typename T>
__global__ void __launch_bounds__(256)
small_n_fc_kernel(const __grid_constant__ BgGeom g, const T* __restrict__ in, const T* __restrict__ w, const float* __restrict__ bias,
                  const T* residual, T* out, int act) {
  const int row = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= g.B) return;                          // warp-uniform
  const long w0 = (long)g.ph[0].taps[0].widx * g.wtap;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = lane; k < g.Ci; k += 32) {
    const float xv = bg_ld(in, (long)row * g.Ci + k);
#pragma unroll
    for (int n = 0; n < 4; ++n)
      if (n < g.Co) acc[n] = fmaf(xv, bg_ld(w, w0 + (long)k * g.wk + (long)n * g.wn), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 4; ++n)
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  if (lane == 0) {
    const long obase = (long)row * g.Co;
    for (int n = 0; n < g.Co; ++n) {
      float v = acc[n];
      if (bias) v += bias[n];
      if (residual) v += bg_ld(residual, obase + n);
      bg_st(out, obase + n, small_act(v, act));
    }
  }
}

// ---- K = g.Ci <= 4: one thread per (output pixel, 8 output channels) ------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
small_k_fwd_kernel(const __grid_constant__ BgGeom g, const T* __restrict__ in, const T* __restrict__ w,
                   const float* __restrict__ bias, const T* residual, T* out, int act, int tiles_per_phase, int ngrp) {
  extern __shared__ float wsm[];     // [tap][k][Co]
  const int phase = blockIdx.x / tiles_per_phase;
  const int tile = blockIdx.x % tiles_per_phase;
  const BgPhase& ph = g.ph[phase];
  for (int idx = threadIdx.x; idx < ph.ntaps * g.Ci * g.Co; idx += blockDim.x) {
    const int n = idx % g.Co, k = (idx / g.Co) % g.Ci, t = idx / (g.Co * g.Ci);
    wsm[idx] = bg_ld(w, (long)ph.taps[t].widx * g.wtap + (long)k * g.wk + (long)n * g.wn);
  }
  __syncthreads();
  const int HWp = ph.Hp * ph.Wp;
  const long gid = (long)tile * blockDim.x + threadIdx.x;
  const long m = gid / ngrp;
  const int n0 = (int)(gid % ngrp) * 8;
  if (m >= (long)g.B * HWp) return;
  const int b = (int)(m / HWp); const int r = (int)(m - (long)b * HWp); const int i = r / ph.Wp, j = r - i * ph.Wp;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int t = 0; t < ph.ntaps; ++t) {
    int iy, ix;
    if (!small_src(g, i, j, ph.taps[t], &iy, &ix)) continue;
    const long base = (((long)b * g.Hi + iy) * g.Wi + ix) * g.Ci;
    for (int k = 0; k < g.Ci; ++k) {
      const float xv = bg_ld(in, base + k);
      const float* wr = wsm + ((long)t * g.Ci + k) * g.Co + n0;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (n0 + e < g.Co) acc[e] = fmaf(xv, wr[e], acc[e]);
    }
  }
  const int oy = ph.oy0 + i * g.ostep, ox = ph.ox0 + j * g.ostep;
  const long obase = (((long)b * g.Ho + oy) * g.Wo + ox) * g.Co + n0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (n0 + e >= g.Co) break;
    float v = acc[e];
    if (bias) v += bias[n0 + e];
    if (residual) v += bg_ld(residual, obase + e);
    acc[e] = small_act(v, act);
  }
  if (g.Co % 8 == 0) {
    bg_st4(out, obase, f4{acc[0], acc[1], acc[2], acc[3]});
    bg_st4(out, obase + 4, f4{acc[4], acc[5], acc[6], acc[7]});
  } else {
    for (int e = 0; e < 8 && n0 + e < g.Co; ++e) bg_st(out, obase + e, acc[e]);
  }
}

// ---- wgrad with min(Ci, Co) <= 4: one thread per (tap, channel of the big side); the block walks a
// chunk of pixel rows; partial sums per chunk are reduced afterwards (deterministic) ----------------
template <typename T, bool SMALL_IS_N>
__global__ void __launch_bounds__(256)
small_wgrad_kernel(const __grid_constant__ BgGeom g, const T* __restrict__ in, const T* __restrict__ og, float* __restrict__ part,
                   long zstride, int rows_per_chunk) {
  const BgPhase& ph = g.ph[0];
  const int big = SMALL_IS_N ? g.Ci : g.Co;
  const int small = SMALL_IS_N ? g.Co : g.Ci;
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;       // (tap, big index)
  if (pair >= ph.ntaps * big) return;
  const int t = pair / big, c = pair - t * big;
  const BgTap tap = ph.taps[t];
  const int nrows = g.B * ph.Hp;
  const int r0 = blockIdx.y * rows_per_chunk, r1 = min(nrows, r0 + rows_per_chunk);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; ++r) {
    const int b = r / ph.Hp, i = r - b * ph.Hp;
    int iy = i * g.istep + tap.dy;
    if (g.pad_mode == BG_PAD_REFLECT) { if (iy < 0) iy = -iy; if (iy >= g.Hi) iy = 2 * (g.Hi - 1) - iy; }
    if (iy < 0 || iy >= g.Hi) continue;
    const long in_row = ((long)b * g.Hi + iy) * g.Wi;
    const long og_row = ((long)b * g.Ho + (ph.oy0 + i * g.ostep)) * g.Wo;
    for (int j = 0; j < ph.Wp; ++j) {
      int ix = j * g.istep + tap.dx;
      if (g.pad_mode == BG_PAD_REFLECT) { if (ix < 0) ix = -ix; if (ix >= g.Wi) ix = 2 * (g.Wi - 1) - ix; }
      if (ix < 0 || ix >= g.Wi) continue;
      const long ib = (in_row + ix) * g.Ci, ob = (og_row + ph.ox0 + j * g.ostep) * g.Co;
      if (SMALL_IS_N) {
        const float v = bg_ld(in, ib + c);
        for (int n = 0; n < small; ++n) acc[n] = fmaf(v, bg_ld(og, ob + n), acc[n]);
      } else {
        const float v = bg_ld(og, ob + c);
        for (int k = 0; k < small; ++k) acc[k] = fmaf(bg_ld(in, ib + k), v, acc[k]);
      }
    }
  }
  float* dst = part + (long)blockIdx.y * zstride + (long)tap.widx * g.wtap;
  for (int s = 0; s < small; ++s) {
    const int k = SMALL_IS_N ? c : s, n = SMALL_IS_N ? s : c;
    dst[(long)k * g.wk + (long)n * g.wn] = acc[s];
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T>
static int small_fwd_launch(const BgGeom& g, const T* in, const T* w, const float* bias, const T* residual, T* out, int act,
                            cudaStream_t st, int* taken) {
  *taken = 0;
  int maxtaps = 0;
  long maxM = 0;
  for (int p = 0; p < g.nphase; ++p) {
    maxtaps = max(maxtaps, g.ph[p].ntaps);
    maxM = max(maxM, (long)g.B * g.ph[p].Hp * g.ph[p].Wp);
  }
  if (g.Co <= 4 && g.nphase == 1 && g.Hi == 1 && g.Wi == 1 && g.Ho == 1 && g.Wo == 1 && g.ph[0].ntaps == 1 && g.ph[0].taps[0].dy == 0 &&
      g.ph[0].taps[0].dx == 0 && g.Ci >= 64) {
    small_n_fc_kernel<T><<<(int)bg_cdiv((long)g.B * 32, 256), 256, 0, st>>>(g, in, w, bias, residual, out, act);
    BG_LAUNCH_CHECK();
    *taken = 1;
  } else if (g.Co <= 4 && (long)maxtaps * g.Ci * 4 <= SMALL_MAX_W) {
    const int tiles = (int)bg_cdiv(maxM, 256);
    small_n_fwd_kernel<T><<<tiles * g.nphase, 256, (size_t)maxtaps * g.Ci * 4 * sizeof(float), st>>>(g, in, w, bias, residual, out, act, tiles);
    BG_LAUNCH_CHECK();
    *taken = 1;
  } else if (g.Ci <= 4 && (long)maxtaps * g.Ci * g.Co <= SMALL_MAX_W) {
    const int ngrp = (int)bg_cdiv(g.Co, 8);
    const int tiles = (int)bg_cdiv(maxM * ngrp, 256);
    small_k_fwd_kernel<T><<<tiles * g.nphase, 256, (size_t)maxtaps * g.Ci * g.Co * sizeof(float), st>>>(g, in, w, bias, residual, out, act,
                                                                                                   tiles, ngrp);
    BG_LAUNCH_CHECK();
    *taken = 1;
  }
  return BG_OK;
}

int bg_small_fwd(const BgGeom& g, const void* in, const void* w, const float* bias, const void* residual, void* out, int act, int dtype,
                 cudaStream_t st, int* taken) {
  BG_DISPATCH_DTYPE(dtype, T, return small_fwd_launch<T>(g, (const T*)in, (const T*)w, bias, (const T*)residual, (T*)out, act, st, taken));
}

static int small_wgrad_chunks(const BgGeom& g) {
  const int nrows = g.B * g.ph[0].Hp;
  const int big = (g.Co <= 4) ? g.Ci : g.Co;
  const long xblocks = bg_cdiv((long)g.ph[0].ntaps * big, 256);
  long want = bg_cdiv(8L * bg_num_sms(), xblocks);
  if (want > nrows) want = nrows;
  if (want < 1) want = 1;
  return (int)want;
}

bool bg_small_wgrad_applicable(const BgGeom& g) { return g.nphase == 1 && (g.Co <= 4 || g.Ci <= 4); }

size_t bg_small_wgrad_ws(const BgGeom& g, long numel) {
  return bg_small_wgrad_applicable(g) ? (size_t)small_wgrad_chunks(g) * numel * sizeof(float) : 0;
}

__global__ void small_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, long n, long zstride, int nsplit) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < nsplit; ++z) s += part[(long)z * zstride + i];
  dw[i] = s;
}

int bg_small_wgrad(const BgGeom& g, const void* in, const void* og, float* dw, long numel, int dtype, void* ws, size_t ws_bytes,
                   cudaStream_t st, int* taken) {
  *taken = 0;
  if (!bg_small_wgrad_applicable(g)) return BG_OK;
  const int nchunk = small_wgrad_chunks(g);
  BG_REQUIRE(ws && ws_bytes >= (size_t)nchunk * numel * sizeof(float), BG_ERR_WORKSPACE, "small wgrad workspace too small");
  const int nrows = g.B * g.ph[0].Hp;
  const int rpc = (int)bg_cdiv(nrows, nchunk);
  const int big = (g.Co <= 4) ? g.Ci : g.Co;
  dim3 grid((unsigned)bg_cdiv((long)g.ph[0].ntaps * big, 256), (unsigned)nchunk);
  float* part = (float*)ws;
  if (g.Co <= 4) {
    BG_DISPATCH_DTYPE(dtype, T, small_wgrad_kernel<T, true><<<grid, 256, 0, st>>>(g, (const T*)in, (const T*)og, part, numel, rpc));
  } else {
    BG_DISPATCH_DTYPE(dtype, T, small_wgrad_kernel<T, false><<<grid, 256, 0, st>>>(g, (const T*)in, (const T*)og, part, numel, rpc));
  }
  BG_LAUNCH_CHECK();
  small_reduce_kernel<<<(unsigned)bg_cdiv(numel, 256), 256, 0, st>>>(part, dw, numel, numel, nchunk);
  BG_LAUNCH_CHECK();
  *taken = 1;
  return BG_OK;
}
