// Grouped small-M fully-connected layers (ops.py:148-175 as used by condition_batch_norm, ops.py:623-627):
// the conditional-BatchNorm gamma / beta projections of a generator pass -- 2 per CBN layer, 20 at 128x128 -- all read
// inputs that exist before the first block runs (z chunk + class vector), so they are hoisted to the start of the pass and
// computed by ONE launch (forward) and ONE launch (weight + bias gradients, at the end of the backward pass) instead of
// ~180 tiny launches per iteration.  The work is weight-bandwidth / latency bound (M = per-GPU batch <= 256 rows,
// K ~ 1032, N = 96..1536): fp32 SIMT FMA on bf16 (or fp32) spectrally-normalised weights, fp32 activations and
// fp32 outputs -- no operand rounding of z, gamma, beta or their gradients.
//
//   forward : y_i[M, N_i]  = x_i[M, K_i] . Wt_i[N_i, K_i]^T + bias_i          (Wt = the K-major copy written by spectral norm)
//   backward: dW_i[K_i, N_i] = x_i^T . dy_i ,  dbias_i[n] = sum_m dy_i[m, n]   (fixed summation order: deterministic)
#include "bg_common.cuh"

#define FCG_MAX 32
struct FcgFwdItem { const float* x; const void* wt; const float* bias; float* y; int K, N, ldw, tile0; };
struct FcgBwdItem { const float* x; const float* dy; float* dW; float* dbias; int K, N, ld_dw, tile0; };
struct FcgFwdParams { FcgFwdItem it[FCG_MAX]; int n, M, ntiles; };
struct FcgBwdParams { FcgBwdItem it[FCG_MAX]; int n, M, ntiles; };

__device__ __forceinline__ void fcg_ld4(const float* p, float v[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void fcg_ld4(const bf16* p, float v[4]) {
  const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  v[0] = __low2float(h[0]); v[1] = __high2float(h[0]); v[2] = __low2float(h[1]); v[3] = __high2float(h[1]);
}

// CTA tile 64 (rows of x) x 64 (output columns), 256 threads, 4 x 4 outputs per thread, K in chunks of 16.
template <typename WT>
__global__ void __launch_bounds__(256) fcg_fwd_kernel(const __grid_constant__ FcgFwdParams p) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  int pi = 0;
  while (pi + 1 < p.n && (int)blockIdx.x >= p.it[pi + 1].tile0) ++pi;
  const FcgFwdItem& it = p.it[pi];
  const int t = blockIdx.x - it.tile0;
  const int ntn = (it.N + 63) / 64;
  const int m0 = (t / ntn) * 64, n0 = (t % ntn) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;          // loader: row lr of the tile, 4 consecutive k
  const WT* wt = (const WT*)it.wt;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < it.K; k0 += 16) {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + lr < p.M && k0 + lk < it.K) fcg_ld4(it.x + (long)(m0 + lr) * it.K + k0 + lk, a);      // K % 4 == 0
    if (n0 + lr < it.N && k0 + lk < it.K) fcg_ld4(wt + (long)(n0 + lr) * it.ldw + k0 + lk, b);
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) { As[lk + e][lr] = a[e]; Bs[lk + e][lr] = b[e]; }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < it.N) it.y[(long)m * it.N + n] = acc[i][j] + (it.bias ? it.bias[n] : 0.f);
    }
  }
}

// CTA tile 64 (k) x 64 (n); the contraction runs over the M rows in chunks of 16.
__global__ void __launch_bounds__(256) fcg_bwd_kernel(const __grid_constant__ FcgBwdParams p) {
  __shared__ float Xs[16][68];
  __shared__ float Gs[16][68];
  int pi = 0;
  while (pi + 1 < p.n && (int)blockIdx.x >= p.it[pi + 1].tile0) ++pi;
  const FcgBwdItem& it = p.it[pi];
  const int t = blockIdx.x - it.tile0;
  const int ntn = (it.N + 63) / 64;
  const int k0 = (t / ntn) * 64, n0 = (t % ntn) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lm = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;         // loader: row lm of the chunk, 4 consecutive columns
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;                                                     // threads 0..63 of the k-tile-0 CTAs: column sums of dy
  for (int m0 = 0; m0 < p.M; m0 += 16) {
    float x[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + lm < p.M) {
      if (k0 + lc < it.K) fcg_ld4(it.x + (long)(m0 + lm) * it.K + k0 + lc, x);                     // K % 4 == 0
      if (n0 + lc < it.N) fcg_ld4(it.dy + (long)(m0 + lm) * it.N + n0 + lc, g);                    // N % 4 == 0
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Xs[lm][lc]) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(&Gs[lm][lc]) = make_float4(g[0], g[1], g[2], g[3]);
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[mm][ty * 4]);
      const float4 gv = *reinterpret_cast<const float4*>(&Gs[mm][tx * 4]);
      const float xr[4] = {xv.x, xv.y, xv.z, xv.w}, gr[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xr[i], gr[j], acc[i][j]);
    }
    if (k0 == 0 && threadIdx.x < 64) {
#pragma unroll
      for (int mm = 0; mm < 16; ++mm) bsum += Gs[mm][threadIdx.x];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + ty * 4 + i;
    if (k >= it.K) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < it.N) it.dW[(long)k * it.ld_dw + n] = acc[i][j];
    }
  }
  if (k0 == 0 && threadIdx.x < 64 && it.dbias && n0 + (int)threadIdx.x < it.N) it.dbias[n0 + threadIdx.x] = bsum;
}

extern "C" int bg_fc_group_fwd(const bg_fc_group_item* items, int n, int M, int w_dtype, void* stream) {
  BG_REQUIRE(n >= 1 && n <= FCG_MAX, BG_ERR_SHAPE, "bg_fc_group_fwd: %d problems (1..%d)", n, FCG_MAX);
  static FcgFwdParams P;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    const bg_fc_group_item& s = items[i];
    BG_REQUIRE(s.K % 4 == 0 && s.ldw % 4 == 0 && s.N >= 1, BG_ERR_UNSUPPORTED, "bg_fc_group_fwd: K = %d, ldw = %d must be multiples of 4", s.K, s.ldw);
    P.it[i] = {(const float*)s.x, (const void*)s.w, (const float*)s.bias, (float*)s.y, s.K, s.N, s.ldw, tiles};
    tiles += (int)(bg_cdiv(M, 64) * bg_cdiv(s.N, 64));
  }
  P.n = n; P.M = M; P.ntiles = tiles;
  if (w_dtype == BG_BF16) fcg_fwd_kernel<bf16><<<tiles, 256, 0, (cudaStream_t)stream>>>(P);
  else fcg_fwd_kernel<float><<<tiles, 256, 0, (cudaStream_t)stream>>>(P);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_fc_group_bwd(const bg_fc_group_item* items, int n, int M, void* stream) {
  BG_REQUIRE(n >= 1 && n <= FCG_MAX, BG_ERR_SHAPE, "bg_fc_group_bwd: %d problems (1..%d)", n, FCG_MAX);
  static FcgBwdParams P;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    const bg_fc_group_item& s = items[i];
    BG_REQUIRE(s.K % 4 == 0 && s.N % 4 == 0 && s.ldw >= s.N, BG_ERR_UNSUPPORTED, "bg_fc_group_bwd: K = %d, N = %d must be multiples of 4", s.K, s.N);
    P.it[i] = {(const float*)s.x, (const float*)s.y, (float*)s.w, (float*)s.bias, s.K, s.N, s.ldw, tiles};
    tiles += (int)(bg_cdiv(s.K, 64) * bg_cdiv(s.N, 64));
  }
  P.n = n; P.M = M; P.ntiles = tiles;
  fcg_bwd_kernel<<<tiles, 256, 0, (cudaStream_t)stream>>>(P);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
