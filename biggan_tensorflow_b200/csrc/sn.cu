// Spectral norm (ops.py:718-747): fused one-step power iteration, scale/cast writer, backward.
// HBM-bound: W [K,C] fp32 is streamed once from HBM for both matvecs (the second use of each row
// tile hits L1/L2), once more for the scale pass.  All reductions are fixed-order (deterministic,
// so data-parallel replicas keep bit-identical u).
#include "bg_common.cuh"

#define SN_MAX_RT 64

static inline int sn_rows_per_block(int K) {
  long rt = bg_cdiv(K, 4L * bg_num_sms());
  if (rt < 1) rt = 1;
  if (rt > SN_MAX_RT) rt = SN_MAX_RT;
  return (int)rt;
}

// v[r] = W[r,:].u for the block's rows, then pcol[blk][c] = sum_r v[r] * W[r,c]
__global__ void __launch_bounds__(256)
sn_matvec_kernel(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ vraw,
                 float* __restrict__ pcol, int K, int C, int RT) {
  __shared__ float vs[SN_MAX_RT];
  const int r0 = blockIdx.x * RT;
  const int nr = min(RT, K - r0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = (C % 4 == 0);
  for (int r = warp; r < nr; r += 8) {
    const float* row = W + (long)(r0 + r) * C;
    float s = 0.f;
    if (vec) {
      for (int c = lane * 4; c < C; c += 128) {
        float4 w4 = *reinterpret_cast<const float4*>(row + c);
        float4 u4 = *reinterpret_cast<const float4*>(u + c);
        s += w4.x * u4.x + w4.y * u4.y + w4.z * u4.z + w4.w * u4.w;
      }
    } else {
      for (int c = lane; c < C; c += 32) s += row[c] * u[c];
    }
    s = bg_warp_sum(s);
    if (lane == 0) { vs[r] = s; vraw[r0 + r] = s; }
  }
  __syncthreads();
  float* dst = pcol + (long)blockIdx.x * C;
  if (vec) {
    for (int c = threadIdx.x * 4; c < C; c += 1024) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < nr; ++r) {
        float4 w4 = *reinterpret_cast<const float4*>(W + (long)(r0 + r) * C + c);
        const float vr = vs[r];
        a.x = fmaf(vr, w4.x, a.x); a.y = fmaf(vr, w4.y, a.y); a.z = fmaf(vr, w4.z, a.z); a.w = fmaf(vr, w4.w, a.w);
      }
      *reinterpret_cast<float4*>(dst + c) = a;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += 256) {
      float a = 0.f;
      for (int r = 0; r < nr; ++r) a = fmaf(vs[r], W[(long)(r0 + r) * C + c], a);
      dst[c] = a;
    }
  }
}

__global__ void sn_colreduce_kernel(const float* __restrict__ pcol, float* __restrict__ uraw, int nblk, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += pcol[(long)b * C + c];
  uraw[c] = s;
}

// single block: normalise v, form u' = uraw * inv|v|, u_hat, sigma = u'.u_hat
__global__ void __launch_bounds__(1024)
sn_finalize_kernel(float* __restrict__ v, const float* __restrict__ uraw, float* __restrict__ u_out,
                   float* __restrict__ sigma, int K, int C) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) { float t = v[k]; s += t * t; }
  const float ssv = bg_block_sum(s, sh);
  const float inv_v = rsqrtf(fmaxf(ssv, 1e-12f));          // tf.nn.l2_normalize: x*rsqrt(max(sum x^2, 1e-12))
  for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] = v[k] * inv_v;
  s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { float t = uraw[c] * inv_v; s += t * t; }
  const float ssu = bg_block_sum(s, sh);
  const float inv_u = rsqrtf(fmaxf(ssu, 1e-12f));
  for (int c = threadIdx.x; c < C; c += blockDim.x) u_out[c] = uraw[c] * inv_v * inv_u;
  if (threadIdx.x == 0) sigma[0] = ssu * inv_u;            // (v_hat W) . u_hat
}

extern "C" size_t bg_sn_workspace_bytes(int K, int C) {
  const int RT = sn_rows_per_block(K);
  const long nblk = bg_cdiv(K, RT);
  size_t a = (size_t)(nblk + 1) * C * sizeof(float);     // power iteration: pcol + uraw
  size_t b = 1024 * sizeof(float);                       // backward: dot partials
  return (a > b ? a : b) + 256;
}

extern "C" int bg_sn_power_iter(const float* W, const float* u, float* u_out, float* v, float* sigma, int K, int C,
                                void* ws, size_t ws_bytes, void* stream) {
  BG_REQUIRE(K > 0 && C > 0, BG_ERR_SHAPE, "sn: bad shape %dx%d", K, C);
  BG_REQUIRE(ws && ws_bytes >= bg_sn_workspace_bytes(K, C), BG_ERR_WORKSPACE, "sn workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int RT = sn_rows_per_block(K);
  const int nblk = (int)bg_cdiv(K, RT);
  float* pcol = (float*)ws;
  float* uraw = pcol + (long)nblk * C;
  sn_matvec_kernel<<<nblk, 256, 0, st>>>(W, u, v, pcol, K, C, RT);
  BG_LAUNCH_CHECK();
  sn_colreduce_kernel<<<(unsigned)bg_cdiv(C, 256), 256, 0, st>>>(pcol, uraw, nblk, C);
  BG_LAUNCH_CHECK();
  sn_finalize_kernel<<<1, 1024, 0, st>>>(v, uraw, u_out, sigma, K, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// out = W / sigma (+ per-tap transposed copy out_t[t][c][r])
template <typename T>
__global__ void sn_scale_kernel(const float* __restrict__ W, const float* __restrict__ sigma, T* __restrict__ out, long n) {
  const float sg = sigma[0];
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 w4 = *reinterpret_cast<const float4*>(W + i);
    bg_st4(out, i, f4{w4.x / sg, w4.y / sg, w4.z / sg, w4.w / sg});
  }
  // tail (n % 4) handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    long j = (n & ~3L) + threadIdx.x;
    bg_st(out, j, W[j] / sg);
  }
}

// generic writer: out[(t*r_pad + r)*c_pad + c] and out_t[t][c][r] (tap stride c_pad*r_pad, row stride r_pad);
// destinations with r_pad > rows or c_pad > C are pre-zeroed by the caller (zero-padded channels)
template <typename T>
__global__ void sn_scale_t_kernel(const float* __restrict__ W, const float* __restrict__ sigma, T* __restrict__ out,
                                  T* __restrict__ out_t, int rows, int C, int r_pad, int c_pad) {
  __shared__ float tile[32][33];
  const float sg = sigma[0];
  const long tapoff = (long)blockIdx.z * rows * C, tapoff_p = (long)blockIdx.z * r_pad * c_pad;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < C) {
      const float v = W[tapoff + (long)r * C + c] / sg;
      tile[i][threadIdx.x] = v;
      bg_st(out, tapoff_p + (long)r * c_pad + c, v);
    }
  }
  if (!out_t) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < C) bg_st(out_t, tapoff_p + (long)c * r_pad + r, tile[threadIdx.x][i]);
  }
}

extern "C" int bg_sn_scale(const float* W, const float* sigma, void* out, void* out_t, int dtype, int taps, int rows,
                           int C, int r_pad, int c_pad, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long n = (long)taps * rows * C;
  BG_REQUIRE(n > 0 && c_pad >= C && r_pad >= rows, BG_ERR_SHAPE, "sn_scale: bad shape");
  if (!out_t && c_pad == C && r_pad == rows) {
    const int blocks = (int)min(bg_cdiv(n, 1024), (long)bg_num_sms() * 8);
    BG_DISPATCH_DTYPE(dtype, T, sn_scale_kernel<T><<<blocks, 256, 0, st>>>(W, sigma, (T*)out, n));
  } else {
    dim3 grid((unsigned)bg_cdiv(C, 32), (unsigned)bg_cdiv(rows, 32), taps), block(32, 8);
    BG_DISPATCH_DTYPE(dtype, T, sn_scale_t_kernel<T><<<grid, block, 0, st>>>(W, sigma, (T*)out, (T*)out_t, rows, C, r_pad, c_pad));
  }
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// backward: dW = G/sigma - (<G,W>/sigma^2) v_hat[k] u_hat[c]
__global__ void __launch_bounds__(256)
sn_dot_partial_kernel(const float* __restrict__ G, const float* __restrict__ W, float* __restrict__ part, long n, int C, int g_ld, int rows, int r_pad) {
  __shared__ float sh[33];
  float s = 0.f;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const long k = i / C;
    const long gi = ((k / rows) * r_pad + (k % rows)) * g_ld + (i % C);
    s = fmaf(G[gi], W[i], s);
  }
  s = bg_block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
sn_bwd_apply_kernel(const float* __restrict__ G, const float* __restrict__ part, int npart, const float* __restrict__ u_hat,
                    const float* __restrict__ v_hat, const float* __restrict__ sigma, float* __restrict__ dW, int K, int C, int g_ld, int rows, int r_pad) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < npart; i += blockDim.x) s += part[i];
  const float dot = bg_block_sum(s, sh);       // identical in every block: fixed order
  const float sg = sigma[0];
  const float coef = dot / (sg * sg);
  const long n = (long)K * C;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int k = (int)(i / C), c = (int)(i - (long)k * C);
    dW[i] = G[((long)(k / rows) * r_pad + (k % rows)) * g_ld + c] / sg - coef * v_hat[k] * u_hat[c];
  }
}

extern "C" int bg_sn_bwd(const float* W, const float* dWn, int rows, int r_pad, int g_ld, const float* u_hat, const float* v_hat,
                         const float* sigma, float* dW, int K, int C, void* ws, size_t ws_bytes, void* stream) {
  BG_REQUIRE(g_ld >= C && rows > 0 && r_pad >= rows && K % rows == 0, BG_ERR_SHAPE, "sn_bwd: bad gradient layout");
  BG_REQUIRE(ws && ws_bytes >= 1024 * sizeof(float), BG_ERR_WORKSPACE, "sn_bwd workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const long n = (long)K * C;
  const int npart = (int)min(bg_cdiv(n, 1024), 1024L);
  sn_dot_partial_kernel<<<npart, 256, 0, st>>>(dWn, W, (float*)ws, n, C, g_ld, rows, r_pad);
  BG_LAUNCH_CHECK();
  const int blocks = (int)min(bg_cdiv(n, 256), (long)bg_num_sms() * 8);
  sn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(dWn, (const float*)ws, npart, u_hat, v_hat, sigma, dW, K, C, g_ld, rows, r_pad);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
