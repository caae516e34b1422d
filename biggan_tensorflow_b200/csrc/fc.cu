// fully_connected (ops.py:148-175) in fp32: skinny GEMMs whose M is the batch (the conditional-BatchNorm gamma/beta
// projections, class embeddings, G's first dense layer).  One kernel serves forward and weight gradient:
//   out[r, n] = sum_c a(r, c) * S[c, n]
//   forward : r = sample, c = input feature,  a = x[r, c],  S = W [K, N]        (+ bias)
//   wgrad   : r = input feature, c = sample,  a = x[c, r],  S = dy [M, N]
// S is streamed once per 16-row block with perfectly coalesced reads (thread = output column); the 16 x 256 slice of `a`
// sits in shared memory and is read as broadcast float4.  These launches are latency-, not bandwidth-bound: the point is
// one short kernel per layer instead of a split-K GEMM plus reduction.
#include "bg_common.cuh"

#define FC_ROWS 16
#define FC_COLS 128
#define FC_CCH 256

__global__ void __launch_bounds__(FC_COLS)
fc_rows_kernel(const float* __restrict__ A, long a_rs, long a_cs, const float* __restrict__ S, const float* __restrict__ bias,
               float* __restrict__ out, int R, int C, int N, int act) {
  __shared__ __align__(16) float as[FC_CCH][FC_ROWS];
  const int n = blockIdx.x * FC_COLS + threadIdx.x;
  const int r0 = blockIdx.y * FC_ROWS;
  float acc[FC_ROWS];
#pragma unroll
  for (int i = 0; i < FC_ROWS; ++i) acc[i] = 0.f;
  for (int c0 = 0; c0 < C; c0 += FC_CCH) {
    const int cn = min(FC_CCH, C - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * FC_ROWS; i += FC_COLS) {
      // consecutive threads walk the contiguous direction of `a`
      int r, c;
      if (a_cs == 1) { c = i % cn; r = i / cn; } else { r = i % FC_ROWS; c = i / FC_ROWS; }
      as[c][r] = (r0 + r < R) ? A[(long)(r0 + r) * a_rs + (long)(c0 + c) * a_cs] : 0.f;
    }
    __syncthreads();
    if (n < N) {
      const float* sp = S + (long)c0 * N + n;
#pragma unroll 8
      for (int c = 0; c < cn; ++c) {
        const float sv = __ldg(sp + (long)c * N);
        const float4* a4 = reinterpret_cast<const float4*>(&as[c][0]);
#pragma unroll
        for (int q = 0; q < FC_ROWS / 4; ++q) {
          const float4 a = a4[q];
          acc[4 * q] = fmaf(a.x, sv, acc[4 * q]); acc[4 * q + 1] = fmaf(a.y, sv, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a.z, sv, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(a.w, sv, acc[4 * q + 3]);
        }
      }
    }
  }
  if (n >= N) return;
  const float bv = bias ? bias[n] : 0.f;
#pragma unroll
  for (int i = 0; i < FC_ROWS; ++i)
    if (r0 + i < R) out[(long)(r0 + i) * N + n] = bg_apply_act(acc[i] + bv, act);
}

// y[M, N] = x[M, K] W[K, N] + bias
int bg_fc_fwd(const float* x, const float* w, const float* bias, float* y, int M, int K, int N, int act, cudaStream_t st) {
  dim3 grid((unsigned)bg_cdiv(N, FC_COLS), (unsigned)bg_cdiv(M, FC_ROWS));
  fc_rows_kernel<<<grid, FC_COLS, 0, st>>>(x, K, 1, w, bias, y, M, K, N, act);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
// dW[K, N] = x[M, K]^T dy[M, N]
int bg_fc_wgrad(const float* x, const float* dy, float* dw, int M, int K, int N, cudaStream_t st) {
  dim3 grid((unsigned)bg_cdiv(N, FC_COLS), (unsigned)bg_cdiv(K, FC_ROWS));
  fc_rows_kernel<<<grid, FC_COLS, 0, st>>>(x, 1, K, dy, nullptr, dw, K, M, N, BG_ACT_NONE);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
