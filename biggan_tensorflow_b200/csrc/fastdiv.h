// Exact division by a launch-time constant (host + device; plain C++ so that tests/test_fastdiv.py can compile it with g++).
#pragma once
#include <stdint.h>
#if !defined(__CUDACC__)
#define BG_FD_HD inline
#else
#define BG_FD_HD __host__ __device__ __forceinline__
#endif

// Division by a launch-time constant in the per-tile index decode of the persistent kernels: the epilogue warps and the two
// single-thread issue loops are latency-bound dependent instruction streams, and one runtime `/` is ~25 instructions (MUFU.RCP
// + fix-ups) -- ncu on the short-main-loop layers showed ~200 of the ~510 instructions per warp and tile in these divisions.
// q = umulhi(n, mul) >> shr is exact for 0 <= n < 2^31 (round-up multiplier with p = 31 + ceil(log2 d) bits).
struct FastDiv { uint32_t mul, shr; int d; };
static inline FastDiv make_fastdiv(int d) {
  FastDiv f; f.d = d; f.mul = 0; f.shr = 0;
  if (d > 1) {
    uint32_t lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;                // ceil(log2 d)
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + (uint32_t)d - 1) / (uint32_t)d);
    f.shr = p - 32;
  }
  return f;
}
BG_FD_HD int fast_div(int n, const FastDiv& f) {
#ifdef __CUDA_ARCH__
  return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
#else
  return f.d == 1 ? n : (int)((((unsigned long long)(uint32_t)n * f.mul) >> 32) >> f.shr);
#endif
}

