// Elementwise / small-reduction kernels of the BigGAN step: relu (ops.py:529-530), tanh (:539-540),
// global_sum_pooling (:503-506), max_pooling (:508-510), softmax + batched matmuls of
// self_attention_2 (:481-485), gamma*o + x (:490), residual adds, casts, fills.
#include "bg_common.cuh"

static inline int ew_blocks(long n, int per_thread = 1) {
  long b = bg_cdiv(n, 256L * per_thread);
  long mx = (long)bg_num_sms() * 16;
  if (b > mx) b = mx;
  if (b < 1) b = 1;
  return (int)b;
}

#define EW_LOOP(i, n)                                                     \
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x, _s = (long)gridDim.x * blockDim.x; i < (n); i += _s)

// vector main loop over [0, nv*N) + scalar tail
#define EW_VLOOP(T, i, n) \
  const long _nv = vok ? (n) / Vec<T>::N : 0; \
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x, _s = (long)gridDim.x * blockDim.x; i < _nv; i += _s)
#define EW_TAIL(T, i, n) \
  for (long i = _nv * Vec<T>::N + (long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long)gridDim.x * blockDim.x)
// vok: every pointer of the launch is 16-byte aligned (else the scalar tail loop covers everything)
static inline int ew_vok(const void* a, const void* b = nullptr, const void* c = nullptr) {
  return ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)c)) & 15) == 0 ? 1 : 0;
}

template <typename T> __global__ void relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long n, int vok) {
  EW_VLOOP(T, i, n) {
    float v[Vec<T>::N];
    Vec<T>::ld(x, i * Vec<T>::N, v);
#pragma unroll
    for (int e = 0; e < Vec<T>::N; ++e) v[e] = fmaxf(v[e], 0.f);
    Vec<T>::st(y, i * Vec<T>::N, v);
  }
  EW_TAIL(T, i, n) bg_st(y, i, fmaxf(bg_ld(x, i), 0.f));
}
template <typename T> __global__ void relu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long n, int vok) {
  EW_VLOOP(T, i, n) {
    float v[Vec<T>::N], g[Vec<T>::N];
    Vec<T>::ld(x, i * Vec<T>::N, v);
    Vec<T>::ld(dy, i * Vec<T>::N, g);
#pragma unroll
    for (int e = 0; e < Vec<T>::N; ++e) g[e] = v[e] > 0.f ? g[e] : 0.f;
    Vec<T>::st(dx, i * Vec<T>::N, g);
  }
  EW_TAIL(T, i, n) bg_st(dx, i, bg_ld(x, i) > 0.f ? bg_ld(dy, i) : 0.f);
}
template <typename T> __global__ void tanh_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long n, int vok) {
  EW_VLOOP(T, i, n) {
    float v[Vec<T>::N], g[Vec<T>::N];
    Vec<T>::ld(y, i * Vec<T>::N, v);
    Vec<T>::ld(dy, i * Vec<T>::N, g);
#pragma unroll
    for (int e = 0; e < Vec<T>::N; ++e) g[e] = g[e] * (1.f - v[e] * v[e]);
    Vec<T>::st(dx, i * Vec<T>::N, g);
  }
  EW_TAIL(T, i, n) { float t = bg_ld(y, i); bg_st(dx, i, bg_ld(dy, i) * (1.f - t * t)); }
}
template <typename T> __global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long n, int vok) {
  EW_VLOOP(T, i, n) {
    float v[Vec<T>::N], g[Vec<T>::N];
    Vec<T>::ld(a, i * Vec<T>::N, v);
    Vec<T>::ld(b, i * Vec<T>::N, g);
#pragma unroll
    for (int e = 0; e < Vec<T>::N; ++e) v[e] += g[e];
    Vec<T>::st(o, i * Vec<T>::N, v);
  }
  EW_TAIL(T, i, n) bg_st(o, i, bg_ld(a, i) + bg_ld(b, i));
}
template <typename T>
__global__ void axpby_kernel(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ alpha_dev, float alpha,
                             float beta, T* __restrict__ o, long n, int vok) {
  const float a = alpha_dev ? alpha_dev[0] : alpha;
  EW_VLOOP(T, i, n) {
    float v[Vec<T>::N], g[Vec<T>::N];
    Vec<T>::ld(x, i * Vec<T>::N, v);
    if (y) {
      Vec<T>::ld(y, i * Vec<T>::N, g);
#pragma unroll
      for (int e = 0; e < Vec<T>::N; ++e) v[e] = a * v[e] + beta * g[e];
    } else {
#pragma unroll
      for (int e = 0; e < Vec<T>::N; ++e) v[e] = a * v[e];
    }
    Vec<T>::st(o, i * Vec<T>::N, v);
  }
  EW_TAIL(T, i, n) bg_st(o, i, a * bg_ld(x, i) + (y ? beta * bg_ld(y, i) : 0.f));
}
template <typename T> __global__ void fill_kernel(T* __restrict__ p, float v, long n) { EW_LOOP(i, n) bg_st(p, i, v); }
template <typename S, typename D> __global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, long n) {
  EW_LOOP(i, n) bg_st(d, i, bg_ld(s, i));
}

extern "C" int bg_relu_fwd(const void* x, void* y, long n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, relu_fwd_kernel<T><<<ew_blocks(n, 4), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, n, ew_vok(x, y)));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_relu_bwd(const void* x, const void* dy, void* dx, long n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, relu_bwd_kernel<T><<<ew_blocks(n, 4), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)dy, (T*)dx, n, ew_vok(x, dy, dx)));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_tanh_bwd(const void* y, const void* dy, void* dx, long n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, tanh_bwd_kernel<T><<<ew_blocks(n, 4), 256, 0, (cudaStream_t)stream>>>((const T*)y, (const T*)dy, (T*)dx, n, ew_vok(y, dy, dx)));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_add(const void* a, const void* b, void* out, long n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, add_kernel<T><<<ew_blocks(n, 4), 256, 0, (cudaStream_t)stream>>>((const T*)a, (const T*)b, (T*)out, n, ew_vok(a, b, out)));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_axpby(const void* x, const void* y, const float* alpha_dev, float alpha, float beta, void* out, long n,
                        int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, axpby_kernel<T><<<ew_blocks(n, 4), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)y, alpha_dev, alpha,
                                                                                          beta, (T*)out, n, ew_vok(x, y, out)));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_fill(void* p, float value, long n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, fill_kernel<T><<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>((T*)p, value, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_cast(const void* src, int sdt, void* dst, int ddt, long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (sdt == BG_F32 && ddt == BG_BF16) cast_kernel<float, bf16><<<ew_blocks(n), 256, 0, st>>>((const float*)src, (bf16*)dst, n);
  else if (sdt == BG_BF16 && ddt == BG_F32) cast_kernel<bf16, float><<<ew_blocks(n), 256, 0, st>>>((const bf16*)src, (float*)dst, n);
  else if (sdt == BG_F32 && ddt == BG_F32) cast_kernel<float, float><<<ew_blocks(n), 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (sdt == BG_BF16 && ddt == BG_BF16) cast_kernel<bf16, bf16><<<ew_blocks(n), 256, 0, st>>>((const bf16*)src, (bf16*)dst, n);
  else { bg_set_error("bad cast dtypes"); return BG_ERR_UNSUPPORTED; }
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// deterministic dot product: per-block partials, then a single block sums them in order
template <typename T>
__global__ void __launch_bounds__(256) dot_partial_kernel(const T* __restrict__ a, const T* __restrict__ b, float* __restrict__ part, long n) {
  __shared__ float sh[33];
  float s = 0.f;
  EW_LOOP(i, n) s = fmaf(bg_ld(a, i), bg_ld(b, i), s);
  s = bg_block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) dot_final_kernel(const float* __restrict__ part, int np, float* __restrict__ out) {
  __shared__ float sh[33];
  float s = 0.f;
  for (int i = threadIdx.x; i < np; i += blockDim.x) s += part[i];
  s = bg_block_sum(s, sh);
  if (threadIdx.x == 0) out[0] = s;
}
extern "C" size_t bg_dot_workspace_bytes(long n) { (void)n; return 1024 * sizeof(float) + 256; }
extern "C" int bg_dot(const void* a, const void* b, float* out, long n, int dtype, void* ws, size_t ws_bytes, void* stream) {
  BG_REQUIRE(ws && ws_bytes >= 1024 * sizeof(float), BG_ERR_WORKSPACE, "dot workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int np = (int)min(bg_cdiv(n, 2048), 1024L);
  BG_DISPATCH_DTYPE(dtype, T, dot_partial_kernel<T><<<np, 256, 0, st>>>((const T*)a, (const T*)b, (float*)ws, n));
  BG_LAUNCH_CHECK();
  dot_final_kernel<<<1, 256, 0, st>>>((const float*)ws, np, out);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// global sum pooling with optional fused ReLU (BigGAN.py:669-671): y[b,c] = sum_hw relu(x[b,hw,c])
template <typename T>
__global__ void sumpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long HW, int C, int relu_first) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  float s = 0.f;
  for (long r = 0; r < HW; ++r) {
    float v = bg_ld(x, ((long)b * HW + r) * C + c);
    s += relu_first ? fmaxf(v, 0.f) : v;
  }
  bg_st(y, (long)b * C + c, s);
}
template <typename T>
__global__ void sumpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long HW, int C, int relu_first, long n) {
  EW_LOOP(i, n) {
    const int c = (int)(i % C);
    const long b = i / ((long)C * HW);
    float g = bg_ld(dy, b * C + c);
    if (relu_first && !(bg_ld(x, i) > 0.f)) g = 0.f;
    bg_st(dx, i, g);
  }
}
extern "C" int bg_sumpool_fwd(const void* x, void* y, int B, long HW, int C, int relu_first, int dtype, void* stream) {
  dim3 grid((unsigned)bg_cdiv(C, 128), B);
  BG_DISPATCH_DTYPE(dtype, T, sumpool_fwd_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, HW, C, relu_first));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_sumpool_bwd(const void* x, const void* dy, void* dx, int B, long HW, int C, int relu_first, int dtype, void* stream) {
  const long n = (long)B * HW * C;
  BG_DISPATCH_DTYPE(dtype, T, sumpool_bwd_kernel<T><<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)dy, (T*)dx, HW, C,
                                                                                                relu_first, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// 2x2/2 max pooling, NHWC (tf.layers.max_pooling2d(2,2,'SAME') on even sizes)
template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, long n) {
  const int Ho = H / 2, Wo = W / 2;
  EW_LOOP(i, n) {
    const int c = (int)(i % C); long r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long b = r / Ho;
    const long base = ((b * H + 2 * oy) * W + 2 * ox) * C + c;
    float m = bg_ld(x, base);
    m = fmaxf(m, bg_ld(x, base + C));
    m = fmaxf(m, bg_ld(x, base + (long)W * C));
    m = fmaxf(m, bg_ld(x, base + (long)W * C + C));
    bg_st(y, i, m);
  }
}
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int H, int W, int C, long n) {
  const int Ho = H / 2, Wo = W / 2;
  EW_LOOP(i, n) {   // one thread per pooled element: writes its 4 inputs
    const int c = (int)(i % C); long r = i / C;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long b = r / Ho;
    const long base = ((b * H + 2 * oy) * W + 2 * ox) * C + c;
    const long off[4] = {0, (long)C, (long)W * C, (long)W * C + C};
    float v[4];
    int best = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = bg_ld(x, base + off[k]); if (v[k] > v[best]) best = k; }
    const float g = bg_ld(dy, i);
#pragma unroll
    for (int k = 0; k < 4; ++k) bg_st(dx, base + off[k], k == best ? g : 0.f);
  }
}
extern "C" int bg_maxpool2_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  BG_REQUIRE(H % 2 == 0 && W % 2 == 0, BG_ERR_UNSUPPORTED, "maxpool2: odd size");
  const long n = (long)B * (H / 2) * (W / 2) * C;
  BG_DISPATCH_DTYPE(dtype, T, maxpool2_fwd_kernel<T><<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, H, W, C, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_maxpool2_bwd(const void* x, const void* dy, void* dx, int B, int H, int W, int C, int dtype, void* stream) {
  BG_REQUIRE(H % 2 == 0 && W % 2 == 0, BG_ERR_UNSUPPORTED, "maxpool2: odd size");
  const long n = (long)B * (H / 2) * (W / 2) * C;
  BG_DISPATCH_DTYPE(dtype, T, maxpool2_bwd_kernel<T><<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, (const T*)dy, (T*)dx, H, W, C, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// the same pooling on a CHANNEL SLICE of a wider NHWC tensor (the f / h thirds of the fused attention projection):
// x points at the first channel of the slice, pixels are ldx elements apart; y is contiguous [B, H/2, W/2, C].
// backward: dx is a slice too (pixel stride ld_dx); every position of the slice is written.  8 channels (16 B) per thread.
__global__ void maxpool2_slice_fwd_kernel(const bf16* __restrict__ x, long ldx, bf16* __restrict__ y, int H, int W, int C8, long n) {
  const int Ho = H / 2, Wo = W / 2;
  EW_LOOP(i, n) {
    const int c = (int)(i % C8); long r = i / C8;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long b = r / Ho;
    const long p00 = (b * H + 2 * oy) * W + 2 * ox;
    const long off[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
    float m[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      Vec<bf16>::ld(x, off[k] * ldx + (long)c * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = k == 0 ? v[e] : fmaxf(m[e], v[e]);
    }
    Vec<bf16>::st(y, i * 8, m);
  }
}
__global__ void maxpool2_slice_bwd_kernel(const bf16* __restrict__ x, long ldx, const bf16* __restrict__ dy, bf16* __restrict__ dx, long ld_dx,
                                          int H, int W, int C8, long n) {
  const int Ho = H / 2, Wo = W / 2;
  EW_LOOP(i, n) {
    const int c = (int)(i % C8); long r = i / C8;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho); const long b = r / Ho;
    const long p00 = (b * H + 2 * oy) * W + 2 * ox;
    const long off[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
    float v[4][8], g[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) Vec<bf16>::ld(x, off[k] * ldx + (long)c * 8, v[k]);
    Vec<bf16>::ld(dy, i * 8, g);
    int best[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {                    // first maximum wins (same tie rule as maxpool2_bwd_kernel)
      best[e] = 0;
#pragma unroll
      for (int k = 1; k < 4; ++k) if (v[k][e] > v[best[e]][e]) best[e] = k;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = best[e] == k ? g[e] : 0.f;
      Vec<bf16>::st(dx, off[k] * ld_dx + (long)c * 8, o);
    }
  }
}
static bool slice_vec_ok(const void* p, long ld, int C) { return C % 8 == 0 && ld % 8 == 0 && (((uintptr_t)p) & 15) == 0; }
extern "C" int bg_maxpool2_slice_fwd(const void* x, long ldx, void* y, int B, int H, int W, int C, int dtype, void* stream) {
  BG_REQUIRE(H % 2 == 0 && W % 2 == 0, BG_ERR_UNSUPPORTED, "maxpool2: odd size");
  BG_REQUIRE(dtype == BG_BF16 && slice_vec_ok(x, ldx, C) && slice_vec_ok(y, C, C), BG_ERR_UNSUPPORTED,
             "maxpool2_slice: bf16, C and strides multiples of 8, 16-byte aligned slices");
  const long n = (long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2_slice_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, H, W, C / 8, n);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_maxpool2_slice_bwd(const void* x, long ldx, const void* dy, void* dx, long ld_dx, int B, int H, int W, int C, int dtype,
                                     void* stream) {
  BG_REQUIRE(H % 2 == 0 && W % 2 == 0, BG_ERR_UNSUPPORTED, "maxpool2: odd size");
  BG_REQUIRE(dtype == BG_BF16 && slice_vec_ok(x, ldx, C) && slice_vec_ok(dx, ld_dx, C) && slice_vec_ok(dy, C, C), BG_ERR_UNSUPPORTED,
             "maxpool2_slice: bf16, C and strides multiples of 8, 16-byte aligned slices");
  const long n = (long)B * (H / 2) * (W / 2) * (C / 8);
  maxpool2_slice_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)dy, (bf16*)dx, ld_dx, H, W, C / 8, n);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// softmax over the last axis: one warp per row (n up to a few thousand)
template <typename T>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long rows, int n) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * n;
  T* yr = y + row * n;
  float m = -INFINITY;
  for (int i = lane; i < n; i += 32) m = fmaxf(m, bg_ld(xr, i));
  m = bg_warp_max(m);
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += expf(bg_ld(xr, i) - m);
  s = bg_warp_sum(s);
  const float inv = 1.f / s;
  for (int i = lane; i < n; i += 32) bg_st(yr, i, expf(bg_ld(xr, i) - m) * inv);
}
template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, long rows, int n) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* yr = y + row * n;
  const T* dr = dy + row * n;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s = fmaf(bg_ld(yr, i), bg_ld(dr, i), s);
  s = bg_warp_sum(s);
  for (int i = lane; i < n; i += 32) bg_st(dx + row * n, i, bg_ld(yr, i) * (bg_ld(dr, i) - s));
}
extern "C" int bg_softmax_fwd(const void* x, void* y, long rows, int n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, softmax_fwd_kernel<T><<<(unsigned)bg_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, rows, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_softmax_bwd(const void* y, const void* dy, void* dx, long rows, int n, int dtype, void* stream) {
  BG_DISPATCH_DTYPE(dtype, T, softmax_bwd_kernel<T><<<(unsigned)bg_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>((const T*)y, (const T*)dy, (T*)dx, rows, n));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// generic strided batched GEMM (SIMT, fp32 accumulate): C[b][m][n] = alpha * sum_k A[b][m,k] B[b][k,n] (+C)
template <typename T>
__global__ void __launch_bounds__(256)
bgemm_kernel(const T* __restrict__ A, const T* __restrict__ Bm, T* __restrict__ Cm, int M, int N, int K, long a_rs, long a_cs,
             long a_bs, long b_rs, long b_cs, long b_bs, long c_rs, long c_bs, float alpha, int accumulate) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const T* Ab = A + (long)blockIdx.z * a_bs;
  const T* Bb = Bm + (long)blockIdx.z * b_bs;
  T* Cb = Cm + (long)blockIdx.z * c_bs;
  // load roles: A tile 64(m) x 16(k); B tile 16(k) x 64(n); 4 elements per thread each
  const bool a_kcontig = (a_cs == 1), b_ncontig = (b_cs == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    float ra[4], rb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int m, k;
      if (a_kcontig) { m = tid >> 2; k = (tid & 3) * 4 + e; } else { k = tid >> 4; m = (tid & 15) * 4 + e; }
      ra[e] = (m0 + m < M && k0 + k < K) ? bg_ld(Ab, (long)(m0 + m) * a_rs + (long)(k0 + k) * a_cs) : 0.f;
      int kk, n;
      if (b_ncontig) { kk = tid >> 4; n = (tid & 15) * 4 + e; } else { n = tid >> 2; kk = (tid & 3) * 4 + e; }
      rb[e] = (k0 + kk < K && n0 + n < N) ? bg_ld(Bb, (long)(k0 + kk) * b_rs + (long)(n0 + n) * b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (a_kcontig) As[(tid & 3) * 4 + e][tid >> 2] = ra[e]; else As[tid >> 4][(tid & 15) * 4 + e] = ra[e];
      if (b_ncontig) Bs[tid >> 4][(tid & 15) * 4 + e] = rb[e]; else Bs[(tid & 3) * 4 + e][tid >> 2] = rb[e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = alpha * acc[i][j];
      if (accumulate) v += bg_ld(Cb, (long)m * c_rs + n);
      bg_st(Cb, (long)m * c_rs + n, v);
    }
  }
}
extern "C" int bg_bgemm(const void* A, const void* Bm, void* C, int batch, int M, int N, int K, long a_rs, long a_cs, long a_bs,
                        long b_rs, long b_cs, long b_bs, long c_rs, long c_bs, float alpha, int accumulate, int dtype,
                        void* stream) {
  BG_REQUIRE(batch > 0 && M > 0 && N > 0 && K > 0, BG_ERR_SHAPE, "bgemm: bad shape");
  dim3 grid((unsigned)bg_cdiv(N, 64), (unsigned)bg_cdiv(M, 64), batch);
  BG_DISPATCH_DTYPE(dtype, T, bgemm_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)A, (const T*)Bm, (T*)C, M, N, K, a_rs, a_cs, a_bs,
                                                                                  b_rs, b_cs, b_bs, c_rs, c_bs, alpha, accumulate));
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// channel pad / slice of an NHWC tensor viewed as [npix, C]: dst[p, c] = c < min(Cs, Cd) ? src[p, c] : 0
template <typename S, typename D>
__global__ void pad_channels_kernel(const S* __restrict__ src, D* __restrict__ dst, long npix, int Cs, int Cd) {
  const long n = npix * Cd;
  EW_LOOP(i, n) {
    const int c = (int)(i % Cd);
    const long p = i / Cd;
    bg_st(dst, i, c < Cs ? bg_ld(src, p * Cs + c) : 0.f);
  }
}
extern "C" int bg_pad_channels(const void* src, int src_dtype, void* dst, int dst_dtype, long npix, int Cs, int Cd, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long n = npix * Cd;
  if (src_dtype == BG_F32 && dst_dtype == BG_BF16) pad_channels_kernel<float, bf16><<<ew_blocks(n), 256, 0, st>>>((const float*)src, (bf16*)dst, npix, Cs, Cd);
  else if (src_dtype == BG_BF16 && dst_dtype == BG_F32) pad_channels_kernel<bf16, float><<<ew_blocks(n), 256, 0, st>>>((const bf16*)src, (float*)dst, npix, Cs, Cd);
  else if (src_dtype == BG_F32 && dst_dtype == BG_F32) pad_channels_kernel<float, float><<<ew_blocks(n), 256, 0, st>>>((const float*)src, (float*)dst, npix, Cs, Cd);
  else if (src_dtype == BG_BF16 && dst_dtype == BG_BF16) pad_channels_kernel<bf16, bf16><<<ew_blocks(n), 256, 0, st>>>((const bf16*)src, (bf16*)dst, npix, Cs, Cd);
  else { bg_set_error("bad dtypes"); return BG_ERR_UNSUPPORTED; }
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// channel-slice copy between NHWC tensors of different widths: dst[p * ld_dst + c] = src[p * ld_src + c], c < width
// (the fused f / g / h attention projection, ops.py:473-479, produces one wide tensor; its slices feed max-pool / attention).
// 16-byte vectors when width, both strides and both base addresses allow it.
template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, long ld_src, T* __restrict__ dst, long ld_dst, long npix, int width) {
  const long n = npix * width;
  EW_LOOP(i, n) {
    const long p = i / width;
    const int c = (int)(i - p * width);
    dst[p * ld_dst + c] = src[p * ld_src + c];
  }
}
__global__ void copy_channels_vec_kernel(const uint4* __restrict__ src, long ld_src, uint4* __restrict__ dst, long ld_dst, long npix, int width) {
  const long n = npix * width;                       // everything in 16-byte units
  EW_LOOP(i, n) {
    const long p = i / width;
    const int c = (int)(i - p * width);
    dst[p * ld_dst + c] = __ldg(src + p * ld_src + c);
  }
}
extern "C" int bg_copy_channels(const void* src, long ld_src, void* dst, long ld_dst, long npix, int width, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int es = dtype == BG_F32 ? 4 : 2, per = 16 / es;
  if (npix <= 0 || width <= 0) return BG_OK;
  if (width % per == 0 && ld_src % per == 0 && ld_dst % per == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const long n = npix * (width / per);
    copy_channels_vec_kernel<<<ew_blocks(n), 256, 0, st>>>((const uint4*)src, ld_src / per, (uint4*)dst, ld_dst / per, npix, width / per);
  } else {
    const long n = npix * width;
    BG_DISPATCH_DTYPE(dtype, T, copy_channels_kernel<T><<<ew_blocks(n), 256, 0, st>>>((const T*)src, ld_src, (T*)dst, ld_dst, npix, width));
  }
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// im2col of the 3-channel image tensors (bf16, stored zero-padded to Cx channels): the first D convs
// (ops.py:49-113 with Cin = 3) become a 1x1 tensor-core GEMM with K = kh*kw*3 (padded to Cz) instead
// of kh*kw passes over 16 padded channels.
//   z[b, oy, ox, (ky*kw + kx)*3 + c] = x[b, R(oy*s + ky - pt), R(ox*s + kx - pl), c],  R = reflect / zero
// col2im is its adjoint (gather form: deterministic, no atomics).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int rgb_src_index(int p, int n, int reflect) {     // padded coordinate -> source index or -1
  if (p >= 0 && p < n) return p;
  if (!reflect) return -1;
  if (p < 0) p = -p;
  if (p >= n) p = 2 * (n - 1) - p;
  return (p >= 0 && p < n) ? p : -1;
}

__global__ void __launch_bounds__(256)
im2col_rgb_kernel(const bf16* __restrict__ x, bf16* __restrict__ z, int B, int H, int W, int Cx, int Ho, int Wo, int kh, int kw, int s,
                  int pt, int pl, int reflect, int Cz) {
  const long total = (long)B * Ho * Wo;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo); const long r = idx / Wo;
    const int oy = (int)(r % Ho); const int b = (int)(r / Ho);
    bf16* zp = z + idx * Cz;
    int k = 0;
    for (int ky = 0; ky < kh; ++ky) {
      const int iy = rgb_src_index(oy * s + ky - pt, H, reflect);
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = rgb_src_index(ox * s + kx - pl, W, reflect);
        bf16 v0 = __float2bfloat16(0.f), v1 = v0, v2 = v0;
        if (iy >= 0 && ix >= 0) {
          const bf16* xp = x + (((long)b * H + iy) * W + ix) * Cx;
          v0 = xp[0]; v1 = xp[1]; v2 = xp[2];
        }
        zp[k] = v0; zp[k + 1] = v1; zp[k + 2] = v2;
        k += 3;
      }
    }
    for (; k < Cz; ++k) zp[k] = __float2bfloat16(0.f);
  }
}

__global__ void __launch_bounds__(256)
col2im_rgb_kernel(const bf16* __restrict__ dz, bf16* __restrict__ dx, int B, int H, int W, int Cx, int Ho, int Wo, int kh, int kw, int s,
                  int pt, int pl, int pb, int pr, int reflect, int Cz) {
  const long total = (long)B * H * W;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W); const long r = idx / W;
    const int y = (int)(r % H); const int b = (int)(r / H);
    // padded coordinates that read this pixel
    int py[3], px[3], ny = 0, nx = 0;
    py[ny++] = y;
    px[nx++] = x;
    if (reflect) {
      if (y >= 1 && y <= pt) py[ny++] = -y;
      if (y >= H - 1 - pb && y <= H - 2) py[ny++] = 2 * (H - 1) - y;
      if (x >= 1 && x <= pl) px[nx++] = -x;
      if (x >= W - 1 - pr && x <= W - 2) px[nx++] = 2 * (W - 1) - x;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = 0; i < ny; ++i)
      for (int ky = 0; ky < kh; ++ky) {
        const int ty = py[i] + pt - ky;
        if (ty < 0 || ty % s) continue;
        const int oy = ty / s;
        if (oy >= Ho) continue;
        for (int j = 0; j < nx; ++j)
          for (int kx = 0; kx < kw; ++kx) {
            const int tx = px[j] + pl - kx;
            if (tx < 0 || tx % s) continue;
            const int ox = tx / s;
            if (ox >= Wo) continue;
            const bf16* zp = dz + (((long)b * Ho + oy) * Wo + ox) * Cz + (ky * kw + kx) * 3;
            a0 += __bfloat162float(zp[0]); a1 += __bfloat162float(zp[1]); a2 += __bfloat162float(zp[2]);
          }
      }
    bf16* xp = dx + idx * Cx;
    if (Cx == 16 && (((uintptr_t)dx) & 15) == 0) {
      __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h2 = __floats2bfloat162_rn(a2, 0.f);
      uint4 o0 = make_uint4(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h2), 0u, 0u);
      reinterpret_cast<uint4*>(xp)[0] = o0;
      reinterpret_cast<uint4*>(xp)[1] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      xp[0] = __float2bfloat16(a0); xp[1] = __float2bfloat16(a1); xp[2] = __float2bfloat16(a2);
      for (int c = 3; c < Cx; ++c) xp[c] = __float2bfloat16(0.f);
    }
  }
}

// 3x3 fast paths: everything in registers, 8-byte pixel loads, 16-byte stores
__global__ void __launch_bounds__(256)
im2col_rgb3x3_kernel(const bf16* __restrict__ x, bf16* __restrict__ z, int B, int H, int W, int Cx, int Ho, int Wo, int s, int pt, int pl,
                     int reflect) {       // Cz == 32, Cx % 4 == 0
  const long total = (long)B * Ho * Wo;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % Wo); const long r = idx / Wo;
    const int oy = (int)(r % Ho); const int b = (int)(r / Ho);
    unsigned short v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = rgb_src_index(oy * s + ky - pt, H, reflect);
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = rgb_src_index(ox * s + kx - pl, W, reflect);
        if (iy >= 0 && ix >= 0) {
          const uint2 px = __ldg(reinterpret_cast<const uint2*>(x + (((long)b * H + iy) * W + ix) * Cx));
          v[(ky * 3 + kx) * 3] = (unsigned short)(px.x & 0xffffu);
          v[(ky * 3 + kx) * 3 + 1] = (unsigned short)(px.x >> 16);
          v[(ky * 3 + kx) * 3 + 2] = (unsigned short)(px.y & 0xffffu);
        }
      }
    }
    uint4* zp = reinterpret_cast<uint4*>(z + idx * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 o;
      o.x = (uint32_t)v[8 * q] | ((uint32_t)v[8 * q + 1] << 16);
      o.y = (uint32_t)v[8 * q + 2] | ((uint32_t)v[8 * q + 3] << 16);
      o.z = (uint32_t)v[8 * q + 4] | ((uint32_t)v[8 * q + 5] << 16);
      o.w = (uint32_t)v[8 * q + 6] | ((uint32_t)v[8 * q + 7] << 16);
      zp[q] = o;
    }
  }
}

extern "C" int bg_im2col_rgb(const void* x, void* z, int B, int H, int W, int Cx, int Ho, int Wo, int kh, int kw, int stride, int pad_t,
                             int pad_l, int pad_mode, int Cz, void* stream) {
  BG_REQUIRE(Cx >= 3 && Cz >= kh * kw * 3 && B > 0, BG_ERR_SHAPE, "im2col_rgb: bad shape");
  const long total = (long)B * Ho * Wo;
  const int blocks = (int)min(bg_cdiv(total, 256), (long)bg_num_sms() * 16);
  if (kh == 3 && kw == 3 && Cz == 32 && Cx % 4 == 0 && (((uintptr_t)x | (uintptr_t)z) & 15) == 0)
    im2col_rgb3x3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)z, B, H, W, Cx, Ho, Wo, stride, pad_t, pad_l,
                                                                  pad_mode == BG_PAD_REFLECT ? 1 : 0);
  else
    im2col_rgb_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)z, B, H, W, Cx, Ho, Wo, kh, kw, stride, pad_t, pad_l,
                                                               pad_mode == BG_PAD_REFLECT ? 1 : 0, Cz);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
extern "C" int bg_col2im_rgb(const void* dz, void* dx, int B, int H, int W, int Cx, int Ho, int Wo, int kh, int kw, int stride, int pad_t,
                             int pad_l, int pad_b, int pad_r, int pad_mode, int Cz, void* stream) {
  BG_REQUIRE(Cx >= 3 && Cz >= kh * kw * 3 && B > 0, BG_ERR_SHAPE, "col2im_rgb: bad shape");
  const long total = (long)B * H * W;
  const int blocks = (int)min(bg_cdiv(total, 256), (long)bg_num_sms() * 16);
  col2im_rgb_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dz, (bf16*)dx, B, H, W, Cx, Ho, Wo, kh, kw, stride, pad_t, pad_l,
                                                             pad_b, pad_r, pad_mode == BG_PAD_REFLECT ? 1 : 0, Cz);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

