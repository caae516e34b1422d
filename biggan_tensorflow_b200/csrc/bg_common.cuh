// Shared helpers for libbiggan_b200 (sm_100a only).
#pragma once
#ifdef BG_HOST_EMULATION
// CPU execution check of the simple kernels (tests/host_emul/cuda_host_emul.h, test infrastructure): never defined in the
// product build (csrc/Makefile).
#include "../../tests/host_emul/cuda_host_emul.h"
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/biggan_b200.h"

typedef __nv_bfloat16 bf16;

void bg_set_error(const char* fmt, ...);

#define BG_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      bg_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(_e));  \
      return BG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

#define BG_LAUNCH_CHECK() BG_CHECK_CUDA(cudaGetLastError())

#define BG_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      bg_set_error(__VA_ARGS__);     \
      return (code);                 \
    }                                \
  } while (0)

static inline int bg_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static inline long bg_cdiv(long a, long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float bg_ld(const float* p, long i) { return p[i]; }
__device__ __forceinline__ float bg_ld(const bf16* p, long i) { return __bfloat162float(p[i]); }
__device__ __forceinline__ void bg_st(float* p, long i, float v) { p[i] = v; }
__device__ __forceinline__ void bg_st(bf16* p, long i, float v) { p[i] = __float2bfloat16(v); }

// 4-wide vector access (16 B for float, 8 B for bf16); pointer must be suitably aligned
struct f4 { float x, y, z, w; };
__device__ __forceinline__ f4 bg_ld4(const float* p, long i) {
  float4 v = *reinterpret_cast<const float4*>(p + i);
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ f4 bg_ld4(const bf16* p, long i) {
  uint2 r = *reinterpret_cast<const uint2*>(p + i);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  return {__low2float(a), __high2float(a), __low2float(b), __high2float(b)};
}
__device__ __forceinline__ void bg_st4(float* p, long i, f4 v) {
  *reinterpret_cast<float4*>(p + i) = make_float4(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void bg_st4(bf16* p, long i, f4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p + i) = r;
}

__device__ __forceinline__ float bg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double bg_warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float bg_warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum (blockDim.x multiple of 32, <= 1024); all threads get the result
__device__ __forceinline__ float bg_block_sum(float v, float* sh /* >= 33 floats */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = bg_warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[lane] : 0.f;
    t = bg_warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ double bg_block_sum_d(double v, double* sh /* >= 33 doubles */) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = bg_warp_sum_d(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[lane] : 0.0;
    t = bg_warp_sum_d(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

__device__ __forceinline__ float bg_apply_act(float v, int act) {
  if (act == BG_ACT_RELU) return fmaxf(v, 0.f);
  if (act == BG_ACT_TANH) return tanhf(v);
  return v;
}

// 16-byte vector access: 4 floats or 8 bf16 per thread (n % VEC == 0 and 16-byte aligned base required)
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void ld(const float* p, long i, float* v) { float4 t = *reinterpret_cast<const float4*>(p + i); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void st(float* p, long i, const float* v) { *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void ld(const bf16* p, long i, float* v) {
    uint4 t = *reinterpret_cast<const uint4*>(p + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] = __low2float(h[e]); v[2 * e + 1] = __high2float(h[e]); }
  }
  static __device__ __forceinline__ void st(bf16* p, long i, const float* v) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p + i) = t;
  }
};
// dispatch on activation storage dtype
#define BG_DISPATCH_DTYPE(dtype, T, ...)                                        \
  do {                                                                          \
    if ((dtype) == BG_F32) { typedef float T; __VA_ARGS__; }                    \
    else if ((dtype) == BG_BF16) { typedef bf16 T; __VA_ARGS__; }               \
    else { bg_set_error("bad dtype %d", (int)(dtype)); return BG_ERR_UNSUPPORTED; } \
  } while (0)

// ---------------- gather-GEMM geometry shared by the SIMT and tcgen05 conv kernels -------------
struct BgTap { short dy, dx; int widx; };
#define BG_MAX_TAPS 36
struct BgPhase { int oy0, ox0, Hp, Wp, ntaps; BgTap taps[BG_MAX_TAPS]; };
struct BgGeom {
  int B, Hi, Wi, Ci;      // tensor the gather reads (k runs over Ci)
  int Ho, Wo, Co;         // tensor the kernel writes / pairs with (n runs over Co)
  int ostep;              // output pixel of phase p: (oy0 + i*ostep, ox0 + j*ostep), i<Hp, j<Wp
  int istep;              // gathered pixel for tap t: (i*istep + dy, j*istep + dx)
  int pad_mode;           // 0 zero fill, 1 reflect
  int nphase;
  long wtap, wk, wn;      // weight element (tap, k, n) lives at widx*wtap + k*wk + n*wn
  BgPhase ph[4];
};

enum BgConvOp { BG_OP_FWD = 0, BG_OP_DGRAD = 1, BG_OP_WGRAD = 2 };
// Builds the gather geometry of `op` for conv descriptor d.  For DGRAD with reflect padding the
// output is the PADDED gradient (Ho/Wo = padded size); the caller folds it afterwards.
int bg_build_geom(const bg_conv_desc* d, int op, BgGeom* g);

// kernel launch on a stream (one spelling for nvcc and for the host emulation build of the simple kernels)
#define BG_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif  // BG_HOST_EMULATION
