// TEST INFRASTRUCTURE ONLY.  A minimal host execution model for the simple (non-tensor-core) kernels of
// libbiggan_b200: with -DBG_HOST_EMULATION, csrc/bg_common.cuh includes this file instead of the CUDA headers, so a
// .cu file written against bg_common.cuh + BG_LAUNCH compiles with g++ and its kernels run on the CPU:
// blocks one after another, the threads of a block as real std::threads (so __syncthreads and warp shuffles work).
// It checks index arithmetic, launch geometry and reductions of the actual kernel source when no GPU is available;
// it says nothing about performance or about anything hardware specific.  Used by tests/test_next_ops_host_emul.py.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include "../../include/biggan_b200.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static          // blocks run one at a time, so one static instance per kernel is one block's memory

typedef void* cudaStream_t;
struct bg_emul_dim3 { unsigned x = 1, y = 1, z = 1; bg_emul_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef bg_emul_dim3 dim3;
inline thread_local bg_emul_dim3 threadIdx, blockIdx;
inline bg_emul_dim3 blockDim, gridDim;
inline std::barrier<>* bg_emul_block_barrier = nullptr;
inline std::vector<std::unique_ptr<std::barrier<>>> bg_emul_warp_barriers;
inline float bg_emul_shfl[1024];

inline void __syncthreads() { bg_emul_block_barrier->arrive_and_wait(); }
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned t = threadIdx.x, w = t >> 5;
  bg_emul_shfl[t] = v;
  bg_emul_warp_barriers[w]->arrive_and_wait();
  const unsigned src = (t & ~31u) | ((t ^ (unsigned)lane_mask) & 31u);
  const float r = src < blockDim.x ? bg_emul_shfl[src] : v;
  bg_emul_warp_barriers[w]->arrive_and_wait();
  return r;
}
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
using std::min;
using std::max;

template <typename F> inline void bg_emul_launch(bg_emul_dim3 grid, bg_emul_dim3 block, F&& body) {
  gridDim = grid; blockDim = block;
  const unsigned nthreads = block.x;
  for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
    std::barrier<> bar((std::ptrdiff_t)nthreads);
    bg_emul_block_barrier = &bar;
    bg_emul_warp_barriers.clear();
    for (unsigned w = 0; w * 32 < nthreads; ++w)
      bg_emul_warp_barriers.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, nthreads - w * 32)));
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < nthreads; ++t)
      ts.emplace_back([&, t] { threadIdx = bg_emul_dim3(t, 0, 0); blockIdx = bg_emul_dim3(bx, by, bz); body(); });
    for (auto& th : ts) th.join();
  }
}
#define BG_LAUNCH(kernel, grid, block, stream, ...) bg_emul_launch(bg_emul_dim3(grid), bg_emul_dim3(block), [&] { kernel(__VA_ARGS__); })

// ---- the slice of bg_common.cuh the simple kernels use ----------------------------------------------------------
struct bf16 { uint16_t v; };
inline float bg_emul_bf16_to_float(bf16 h) { uint32_t u = (uint32_t)h.v << 16; float f; std::memcpy(&f, &u, 4); return f; }
inline bf16 bg_emul_float_to_bf16(float f) {          // round to nearest even, like __float2bfloat16
  uint32_t u; std::memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return bf16{(uint16_t)(u >> 16)};
}
inline char bg_emul_error[512];
inline void bg_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(bg_emul_error, sizeof(bg_emul_error), fmt, ap); va_end(ap); }
#define BG_LAUNCH_CHECK() do { } while (0)
#define BG_REQUIRE(cond, code, ...) do { if (!(cond)) { bg_set_error(__VA_ARGS__); return (code); } } while (0)
static inline int bg_num_sms() { return 2; }
static inline long bg_cdiv(long a, long b) { return (a + b - 1) / b; }
inline float bg_ld(const float* p, long i) { return p[i]; }
inline float bg_ld(const bf16* p, long i) { return bg_emul_bf16_to_float(p[i]); }
inline void bg_st(float* p, long i, float v) { p[i] = v; }
inline void bg_st(bf16* p, long i, float v) { p[i] = bg_emul_float_to_bf16(v); }
inline float bg_warp_sum(float v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
inline float bg_block_sum(float v, float* sh) {
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
#define BG_DISPATCH_DTYPE(dtype, T, ...)                                        \
  do {                                                                          \
    if ((dtype) == BG_F32) { typedef float T; __VA_ARGS__; }                    \
    else if ((dtype) == BG_BF16) { typedef bf16 T; __VA_ARGS__; }               \
    else { bg_set_error("bad dtype %d", (int)(dtype)); return BG_ERR_UNSUPPORTED; } \
  } while (0)
