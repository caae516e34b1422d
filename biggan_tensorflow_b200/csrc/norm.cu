// (Conditional) batch norm: ops.py:611-643 (condition_batch_norm), ops.py:580-585 (batch_norm).
// HBM-bound channel reductions over NHWC with per-sample gamma/beta and fused ReLU.  Statistics
// are produced as double-precision (sum, sumsq) so that data-parallel ranks can all-reduce them
// (cross-replica BN) before bg_bn_finalize.
#include "bg_common.cuh"

#define BN_CPB 128   // channels per block (32 lanes x 4)

static inline int bn_nchunk(int B, long rows, int C) {
  // chunks of one sample's HW rows; B * nchunk * channel-blocks ~ 16 CTAs per SM
  long want = bg_cdiv(16L * bg_num_sms(), (long)B);
  long mx = bg_cdiv(rows, 32);
  long n = want < mx ? want : mx;
  if (n < 1) n = 1;
  if (n > 4096) n = 4096;
  return (int)n;
}

// Thread layout of the streaming kernels: a pixel row of C channels is covered by TPR = C / V threads holding one
// 16-byte vector each (V = 8 bf16 / 4 fp32), R = blockDim / TPRB consecutive rows per block iteration, so a warp
// always reads/writes full contiguous lines whatever C is.  Channel blocks (gridDim.x) only when TPR > 256.
struct BnLayout { int V, TPR, nxb, TPRB, R, threads; };
static BnLayout bn_layout(int C, int dtype) {
  BnLayout L;
  L.V = dtype == BG_F32 ? 4 : 8;
  L.TPR = C / L.V;
  L.nxb = (int)bg_cdiv(L.TPR, 256);
  while (L.TPR % L.nxb) ++L.nxb;
  L.TPRB = L.TPR / L.nxb;
  L.R = 256 / L.TPRB;
  if (L.R < 1) L.R = 1;
  L.threads = L.TPRB * L.R;
  return L;
}
static inline bool bn_vec_ok(int C, int dtype) { return C % (dtype == BG_F32 ? 4 : 8) == 0; }

// MODE 0: q1 = x, q2 = x^2.   MODE 1: g = dy * act'(y); q1 = g, q2 = g * xhat
template <typename T, int MODE>
__global__ void __launch_bounds__(256, 4)
bn_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, const float* __restrict__ mean,
                 const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int gb_bstride, int act, float eps, float* __restrict__ part, long HW, int C, int nchunk, int TPRB, int R) {
  constexpr int V = Vec<T>::N;
  extern __shared__ float bn_sh[];                    // [2][R][TPRB*V]
  const int r_in = threadIdx.x / TPRB, cv = threadIdx.x - r_in * TPRB;
  const int c = (blockIdx.x * TPRB + cv) * V;         // first channel of this thread
  const int b = blockIdx.z;
  const long rpc = (HW + nchunk - 1) / nchunk;
  const long r0 = (long)blockIdx.y * rpc, r1 = min(HW, r0 + rpc);
  float s1[V], s2[V];
#pragma unroll
  for (int e = 0; e < V; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
  // bf16 storage: accumulate sum(g) and sum(g * x) and form sum(g * xhat) = rs * (sum(g x) - mu sum(g)) once at the end --
  // the per-thread mean / rstd vectors would cost 16 registers and a third of the occupancy; the fp32 (parity) path keeps
  // the term-by-term form
  constexpr bool LEAN = (MODE == 1) && (sizeof(T) == 2);
  float inv[V], shift[V], mu[LEAN ? 1 : V], rs[LEAN ? 1 : V];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float rse = rsqrtf(var[c + e] + eps), mue = mean[c + e];
      if constexpr (!LEAN) { rs[e] = rse; mu[e] = mue; }
      inv[e] = rse * gamma[(long)b * gb_bstride + c + e];
      shift[e] = beta[(long)b * gb_bstride + c + e] - mue * inv[e];
    }
  }
  const long base = (long)b * HW * C + c;
#pragma unroll (MODE == 0 ? 4 : 2)
  for (long r = r0 + r_in; r < r1; r += R) {
    float xe[V];
    Vec<T>::ld(x, base + r * C, xe);
    if (MODE == 0) {
#pragma unroll
      for (int e = 0; e < V; ++e) { s1[e] += xe[e]; s2[e] = fmaf(xe[e], xe[e], s2[e]); }
    } else {
      float de[V];
      Vec<T>::ld(dy, base + r * C, de);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float g = de[e];
        if (act == BG_ACT_RELU && !(xe[e] * inv[e] + shift[e] > 0.f)) g = 0.f;
        s1[e] += g;
        if constexpr (LEAN) s2[e] = fmaf(g, xe[e], s2[e]);
        else s2[e] = fmaf(g, (xe[e] - mu[e]) * rs[e], s2[e]);
      }
    }
  }
  if constexpr (LEAN) {
#pragma unroll
    for (int e = 0; e < V; ++e) s2[e] = rsqrtf(var[c + e] + eps) * (s2[e] - mean[c + e] * s1[e]);
  }
  const int CB = TPRB * V;                            // channels of this block
  float* sh1 = bn_sh;
  float* sh2 = bn_sh + R * CB;
#pragma unroll
  for (int e = 0; e < V; ++e) { sh1[r_in * CB + cv * V + e] = s1[e]; sh2[r_in * CB + cv * V + e] = s2[e]; }
  __syncthreads();
  for (int i = threadIdx.x; i < CB; i += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int r = 0; r < R; ++r) { a += sh1[r * CB + i]; q += sh2[r * CB + i]; }
    float* dst = part + (((long)b * nchunk + blockIdx.y) * 2) * C + blockIdx.x * CB;
    dst[i] = a;
    dst[C + i] = q;
  }
}

// sums over (sample, chunk) partials -> double [2C].  Block = 32 channels x 32 row groups (fixed-order tree).
__global__ void __launch_bounds__(1024) bn_stats_final_kernel(const float* __restrict__ part, double* __restrict__ sums, int nrow, int C) {
  __shared__ double sa[32][33], sq[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double a = 0.0, q = 0.0;
  if (c < C) {
    // the partial rows are L2 hits whose latency dominated this kernel (one dependent load pair per iteration): issue 8
    // row pairs at once, then add them in the same fixed order
    int i = ty;
    for (; i + 7 * 32 < nrow; i += 8 * 32) {
      float va[8], vq[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        va[k] = __ldg(part + ((long)(i + k * 32) * 2) * C + c);
        vq[k] = __ldg(part + ((long)(i + k * 32) * 2 + 1) * C + c);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { a += (double)va[k]; q += (double)vq[k]; }
    }
    for (; i < nrow; i += 32) { a += (double)part[((long)i * 2) * C + c]; q += (double)part[((long)i * 2 + 1) * C + c]; }
  }
  sa[ty][tx] = a; sq[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    double ta = 0.0, tq = 0.0;
    for (int i = 0; i < 32; ++i) { ta += sa[i][tx]; tq += sq[i][tx]; }
    sums[c] = ta;
    sums[C + c] = tq;
  }
}

// per-sample: dbeta[b,c], dgamma[b,c] = sum over chunks
__global__ void bn_bwd_final_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    int nchunk, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  float a = 0.f, q = 0.f;
  for (int i = 0; i < nchunk; ++i) {
    a += part[(((long)b * nchunk + i) * 2) * C + c];
    q += part[(((long)b * nchunk + i) * 2 + 1) * C + c];
  }
  dbeta[(long)b * C + c] = a;
  dgamma[(long)b * C + c] = q;
}

extern "C" size_t bg_bn_workspace_bytes(int B, long HW, int C) {
  return (size_t)B * bn_nchunk(B, HW, C) * 2 * C * sizeof(float) + 256;
}

extern "C" int bg_bn_stats(const void* x, double* sums, int B, long HW, int C, int dtype, void* ws, size_t ws_bytes,
                           void* stream) {
  BG_REQUIRE(C % 4 == 0, BG_ERR_UNSUPPORTED, "bn: C=%d must be a multiple of 4", C);
  BG_REQUIRE(ws && ws_bytes >= bg_bn_workspace_bytes(B, HW, C), BG_ERR_WORKSPACE, "bn workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  BG_REQUIRE(bn_vec_ok(C, dtype), BG_ERR_UNSUPPORTED, "bn: C=%d must be a multiple of %d", C, dtype == BG_F32 ? 4 : 8);
  const int nchunk = bn_nchunk(B, HW, C);
  const BnLayout L = bn_layout(C, dtype);
  dim3 grid(L.nxb, nchunk, B);
  const size_t shb = (size_t)2 * L.R * L.TPRB * L.V * sizeof(float);
  BG_DISPATCH_DTYPE(dtype, T, bn_reduce_kernel<T, 0><<<grid, L.threads, shb, st>>>((const T*)x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0,
                                                                                0.f, (float*)ws, HW, C, nchunk, L.TPRB, L.R));
  BG_LAUNCH_CHECK();
  bn_stats_final_kernel<<<(unsigned)bg_cdiv(C, 32), 1024, 0, st>>>((const float*)ws, sums, B * nchunk, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// Column sums of a [rows, C] matrix (bias gradients) on the statistics kernels above: the vectorised reduce pass
// (16-byte loads, full lines whatever C is) + a fixed-order tree over its partial rows.
__global__ void __launch_bounds__(1024) colsum_from_part_kernel(const float* __restrict__ part, float* __restrict__ out, int nrow, int C) {
  __shared__ double sa[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double a = 0.0;
  if (c < C) {
    int i = ty;
    for (; i + 7 * 32 < nrow; i += 8 * 32) {
      float va[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) va[k] = __ldg(part + ((long)(i + k * 32) * 2) * C + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) a += (double)va[k];
    }
    for (; i < nrow; i += 32) a += (double)part[((long)i * 2) * C + c];
  }
  sa[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && c < C) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += sa[i][tx];
    out[c] = (float)t;
  }
}
size_t bg_bn_colsum_ws(long rows, int C) { return bg_bn_workspace_bytes(1, rows, C); }
bool bg_bn_colsum_applicable(int C, int dtype) { return bn_vec_ok(C, dtype); }
int bg_bn_colsum(const void* x, float* out, long rows, int C, int dtype, void* ws, cudaStream_t st) {
  const int nchunk = bn_nchunk(1, rows, C);
  const BnLayout L = bn_layout(C, dtype);
  dim3 grid(L.nxb, nchunk, 1);
  const size_t shb = (size_t)2 * L.R * L.TPRB * L.V * sizeof(float);
  BG_DISPATCH_DTYPE(dtype, T, bn_reduce_kernel<T, 0><<<grid, L.threads, shb, st>>>((const T*)x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0,
                                                                                0.f, (float*)ws, rows, C, nchunk, L.TPRB, L.R));
  BG_LAUNCH_CHECK();
  colsum_from_part_kernel<<<(unsigned)bg_cdiv(C, 32), 1024, 0, st>>>((const float*)ws, out, nchunk, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, float* __restrict__ mean,
                                   float* __restrict__ var, float* run_mean, float* run_var, float momentum,
                                   float var_correction, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] / count;
  double v = sums[C + c] / count - m * m;
  if (v < 0.0) v = 0.0;
  mean[c] = (float)m;
  var[c] = (float)v;
  if (run_mean) run_mean[c] = run_mean[c] * momentum + (float)m * (1.f - momentum);
  if (run_var) run_var[c] = run_var[c] * momentum + (float)v * var_correction * (1.f - momentum);
}

extern "C" int bg_bn_finalize(const double* sums, double count, float* mean, float* var, float* run_mean, float* run_var,
                              float momentum, float var_correction, int C, void* stream) {
  bn_finalize_kernel<<<(unsigned)bg_cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, count, mean, var, run_mean, run_var,
                                                                                momentum, var_correction, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// y = act(x*inv + (beta - mean*inv)), inv = rsqrt(var+eps)*gamma   (tf.nn.batch_normalization form)
// grid: x = 128-channel block, y = row chunk of one sample, z = sample; a thread keeps its 4 channels'
// scale/shift in registers and streams rows (coalesced 16-byte accesses, no per-element index math).
template <typename T>
__global__ void __launch_bounds__(256)
bn_apply_kernel(const T* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                const float* __restrict__ gamma, const float* __restrict__ beta, int gb_bstride, T* __restrict__ y, int act,
                float eps, long HW, int C, int nchunk, int TPRB, int R) {
  constexpr int V = Vec<T>::N;
  const int r_in = threadIdx.x / TPRB, cv = threadIdx.x - r_in * TPRB;
  const int c = (blockIdx.x * TPRB + cv) * V;
  const int b = blockIdx.z;
  const long rpc = (HW + nchunk - 1) / nchunk;
  const long r0 = (long)blockIdx.y * rpc, r1 = min(HW, r0 + rpc);
  float inv[V], shift[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    inv[e] = rsqrtf(var[c + e] + eps) * gamma[(long)b * gb_bstride + c + e];
    shift[e] = beta[(long)b * gb_bstride + c + e] - mean[c + e] * inv[e];
  }
  const long base = (long)b * HW * C + c;
#pragma unroll 2
  for (long r = r0 + r_in; r < r1; r += R) {
    float v[V];
    Vec<T>::ld(x, base + r * C, v);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      v[e] = v[e] * inv[e] + shift[e];
      if (act == BG_ACT_RELU) v[e] = fmaxf(v[e], 0.f);
    }
    Vec<T>::st(y, base + r * C, v);
  }
}

static inline int bn_apply_chunks(int B, long HW, int C) {
  long want = bg_cdiv(16L * bg_num_sms(), (long)B);
  long mx = bg_cdiv(HW, 32);
  long n = want < mx ? want : mx;
  return (int)(n < 1 ? 1 : n);
}

extern "C" int bg_bn_apply(const void* x, const float* mean, const float* var, const float* gamma, const float* beta,
                           int gb_bstride, void* y, int act, float eps, int B, long HW, int C, int dtype, void* stream) {
  BG_REQUIRE(bn_vec_ok(C, dtype), BG_ERR_UNSUPPORTED, "bn: C=%d must be a multiple of %d", C, dtype == BG_F32 ? 4 : 8);
  const int nchunk = bn_apply_chunks(B, HW, C);
  const BnLayout L = bn_layout(C, dtype);
  dim3 grid(L.nxb, nchunk, B);
  BG_DISPATCH_DTYPE(dtype, T, bn_apply_kernel<T><<<grid, L.threads, 0, (cudaStream_t)stream>>>((const T*)x, mean, var, gamma, beta, gb_bstride,
                                                                                           (T*)y, act, eps, HW, C, nchunk, L.TPRB, L.R));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* var, const float* gamma,
                                const float* beta, int gb_bstride, float* dgamma, float* dbeta, int act, float eps, int B,
                                long HW, int C, int dtype, void* ws, size_t ws_bytes, void* stream) {
  BG_REQUIRE(C % 4 == 0, BG_ERR_UNSUPPORTED, "bn: C=%d must be a multiple of 4", C);
  BG_REQUIRE(ws && ws_bytes >= bg_bn_workspace_bytes(B, HW, C), BG_ERR_WORKSPACE, "bn workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  BG_REQUIRE(bn_vec_ok(C, dtype), BG_ERR_UNSUPPORTED, "bn: C=%d must be a multiple of %d", C, dtype == BG_F32 ? 4 : 8);
  const int nchunk = bn_nchunk(B, HW, C);
  const BnLayout L = bn_layout(C, dtype);
  dim3 grid(L.nxb, nchunk, B);
  const size_t shb = (size_t)2 * L.R * L.TPRB * L.V * sizeof(float);
  BG_DISPATCH_DTYPE(dtype, T, bn_reduce_kernel<T, 1><<<grid, L.threads, shb, st>>>((const T*)x, (const T*)dy, mean, var, gamma, beta, gb_bstride,
                                                                                act, eps, (float*)ws, HW, C, nchunk, L.TPRB, L.R));
  BG_LAUNCH_CHECK();
  dim3 g2((unsigned)bg_cdiv(C, 128), B);
  bn_bwd_final_kernel<<<g2, 128, 0, st>>>((const float*)ws, dgamma, dbeta, nchunk, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

__global__ void __launch_bounds__(1024)
bn_bwd_sums_kernel(const float* __restrict__ dgamma, const float* __restrict__ dbeta, const float* __restrict__ gamma, int gb_bstride,
                   double* __restrict__ s, int B, int C) {
  __shared__ double sa[32][33], sq[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double a = 0.0, q = 0.0;
  if (c < C)
    for (int b = ty; b < B; b += 32) {
      const double gm = (double)gamma[(long)b * gb_bstride + c];
      a += gm * (double)dbeta[(long)b * C + c];
      q += gm * (double)dgamma[(long)b * C + c];
    }
  sa[ty][tx] = a; sq[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < C) {
    double ta = 0.0, tq = 0.0;
    for (int i = 0; i < 32; ++i) { ta += sa[i][tx]; tq += sq[i][tx]; }
    s[c] = ta;
    s[C + c] = tq;
  }
}

extern "C" int bg_bn_bwd_sums(const float* dgamma, const float* dbeta, const float* gamma, int gb_bstride, double* s, int B,
                              int C, void* stream) {
  bn_bwd_sums_kernel<<<(unsigned)bg_cdiv(C, 32), 1024, 0, (cudaStream_t)stream>>>(dgamma, dbeta, gamma, gb_bstride, s, B, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// dx = rstd * (gamma*g - s1/count - xhat * s2/count)      (same thread layout as bn_apply_kernel)
template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                    int gb_bstride, const double* __restrict__ s, double count, T* __restrict__ dx, int act, float eps,
                    long HW, int C, int nchunk, int TPRB, int R) {
  constexpr int V = Vec<T>::N;
  const int r_in = threadIdx.x / TPRB, cv = threadIdx.x - r_in * TPRB;
  const int c = (blockIdx.x * TPRB + cv) * V;
  const int b = blockIdx.z;
  const long rpc = (HW + nchunk - 1) / nchunk;
  const long r0 = (long)blockIdx.y * rpc, r1 = min(HW, r0 + rpc);
  // dx = inv*g + kb + kx*x with inv = rstd*gamma, kx = -rstd^2*m2, kb = -rstd*m1 + mean*rstd^2*m2: four per-channel constants
  // instead of seven, which leaves registers for two rows in flight per thread
  float inv[V], shift[V], kb[V], kx[V];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    const float rs = rsqrtf(var[c + e] + eps), mu = mean[c + e];
    const float m1 = (float)(s[c + e] / count), m2 = (float)(s[C + c + e] / count);
    inv[e] = rs * gamma[(long)b * gb_bstride + c + e];
    shift[e] = beta[(long)b * gb_bstride + c + e] - mu * inv[e];
    kx[e] = -rs * rs * m2;
    kb[e] = -rs * m1 - mu * kx[e];
  }
  const long base = (long)b * HW * C + c;
#pragma unroll 2
  for (long r = r0 + r_in; r < r1; r += R) {
    float xe[V], de[V];
    Vec<T>::ld(x, base + r * C, xe);
    Vec<T>::ld(dy, base + r * C, de);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float g = de[e];
      if (act == BG_ACT_RELU && !(xe[e] * inv[e] + shift[e] > 0.f)) g = 0.f;
      de[e] = fmaf(inv[e], g, fmaf(kx[e], xe[e], kb[e]));
    }
    Vec<T>::st(dx, base + r * C, de);
  }
}

extern "C" int bg_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* var, const float* gamma,
                               const float* beta, int gb_bstride, const double* s, double count, void* dx, int act, float eps,
                               int B, long HW, int C, int dtype, void* stream) {
  BG_REQUIRE(bn_vec_ok(C, dtype), BG_ERR_UNSUPPORTED, "bn: C=%d must be a multiple of %d", C, dtype == BG_F32 ? 4 : 8);
  const int nchunk = bn_apply_chunks(B, HW, C);
  const BnLayout L = bn_layout(C, dtype);
  dim3 grid(L.nxb, nchunk, B);
  BG_DISPATCH_DTYPE(dtype, T, bn_bwd_apply_kernel<T><<<grid, L.threads, 0, (cudaStream_t)stream>>>((const T*)dy, (const T*)x, mean, var, gamma,
                                                                                               beta, gb_bstride, s, count, (T*)dx, act,
                                                                                               eps, HW, C, nchunk, L.TPRB, L.R));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
