// TEST INFRASTRUCTURE ONLY.  A minimal host execution model for the simple (non-tensor-core) kernels of
// libbiggan_b200: with -DBG_HOST_EMULATION, csrc/bg_common.cuh includes this file instead of the CUDA headers, so a
// .cu file written against bg_common.cuh + BG_LAUNCH compiles with g++ and its kernels run on the CPU:
// blocks one after another; the threads of a block in turn, or as round-robin fibers when the kernel synchronises
// (so __syncthreads and warp shuffles work).
// It checks index arithmetic, launch geometry and reductions of the actual kernel source when no GPU is available;
// it says nothing about performance or about anything hardware specific.  Used by tests/test_next_ops_host_emul.py.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <ucontext.h>
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
inline bg_emul_dim3 threadIdx, blockIdx, blockDim, gridDim;

// Execution model.  A kernel whose threads never synchronise runs one CUDA thread after the other.  The first
// __syncthreads / warp shuffle it executes raises bg_emul_needs_threads (the kernels here reach their first barrier
// before any global write); the kernel is then remembered as cooperative and the launch restarts with one FIBER
// (ucontext coroutine) per CUDA thread of a block, scheduled round-robin in the calling host thread: a fiber that waits
// at a barrier yields until all parties have arrived.  Blocks run one after another (a kernel's `static` shared memory is
// one block's memory).
struct bg_emul_needs_threads {};
struct bg_emul_barrier { unsigned count = 0, gen = 0, parties = 0; };
#if defined(__x86_64__)
// swapcontext() costs two signal-mask system calls per switch; a fiber switch only has to exchange the callee-saved
// registers and the stack pointer (System V x86-64 ABI).
extern "C" void bg_emul_switch(void** save_sp, void* load_sp);
__asm__(".text\n.globl bg_emul_switch\n.type bg_emul_switch,@function\nbg_emul_switch:\n"
        "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
        "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
        "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
        ".size bg_emul_switch,.-bg_emul_switch\n");
struct bg_emul_fiber { void* sp = nullptr; bool done = false; };
inline void* bg_emul_main_sp = nullptr;
#define BG_EMUL_TO_MAIN(f) bg_emul_switch(&(f).sp, bg_emul_main_sp)
#define BG_EMUL_TO_FIBER(f) bg_emul_switch(&bg_emul_main_sp, (f).sp)
#else
struct bg_emul_fiber { ucontext_t ctx; bool done = false; };
inline ucontext_t bg_emul_main_ctx;
#define BG_EMUL_TO_MAIN(f) swapcontext(&(f).ctx, &bg_emul_main_ctx)
#define BG_EMUL_TO_FIBER(f) swapcontext(&bg_emul_main_ctx, &(f).ctx)
#endif
inline bool bg_emul_sequential = false;
inline std::vector<bg_emul_fiber> bg_emul_fibers;
inline std::vector<char> bg_emul_stacks;
inline unsigned bg_emul_cur = 0;
inline std::function<void()> bg_emul_body;
inline bg_emul_barrier bg_emul_block_bar;
inline std::vector<bg_emul_barrier> bg_emul_warp_bars;
inline float bg_emul_shfl[1024];
constexpr size_t BG_EMUL_STACK = 128 * 1024;

inline void bg_emul_wait(bg_emul_barrier& b) {
  const unsigned g = b.gen;
  if (++b.count == b.parties) { b.count = 0; ++b.gen; return; }
  while (b.gen == g) BG_EMUL_TO_MAIN(bg_emul_fibers[bg_emul_cur]);
}
inline void __syncthreads() {
  if (bg_emul_sequential) throw bg_emul_needs_threads();
  bg_emul_wait(bg_emul_block_bar);
}
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  if (bg_emul_sequential) throw bg_emul_needs_threads();
  const unsigned t = bg_emul_cur, w = t >> 5;
  bg_emul_shfl[t] = v;
  bg_emul_wait(bg_emul_warp_bars[w]);
  const unsigned src = (t & ~31u) | ((t ^ (unsigned)lane_mask) & 31u);
  const float r = src < blockDim.x ? bg_emul_shfl[src] : v;
  bg_emul_wait(bg_emul_warp_bars[w]);
  return r;
}
// lane l receives lane l - delta's value (lanes below delta keep their own) / lane l + delta's (lanes past the end keep their own)
inline float bg_emul_shfl_rel(float v, int rel) {
  if (bg_emul_sequential) throw bg_emul_needs_threads();
  const unsigned t = bg_emul_cur, w = t >> 5;
  bg_emul_shfl[t] = v;
  bg_emul_wait(bg_emul_warp_bars[w]);
  const int lane = (int)(t & 31u) + rel;
  const unsigned src = (t & ~31u) + (unsigned)lane;
  const float r = (lane >= 0 && lane < 32 && src < blockDim.x) ? bg_emul_shfl[src] : v;
  bg_emul_wait(bg_emul_warp_bars[w]);
  return r;
}
inline float __shfl_up_sync(unsigned, float v, unsigned delta) { return bg_emul_shfl_rel(v, -(int)delta); }
inline float __shfl_down_sync(unsigned, float v, unsigned delta) { return bg_emul_shfl_rel(v, (int)delta); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
using std::min;
using std::max;

inline void bg_emul_fiber_entry() {
  bg_emul_body();
  bg_emul_fibers[bg_emul_cur].done = true;
  BG_EMUL_TO_MAIN(bg_emul_fibers[bg_emul_cur]);
  __builtin_trap();                        // a finished fiber is never resumed
}
inline std::vector<const void*>& bg_emul_cooperative_kernels() { static std::vector<const void*> v; return v; }

template <typename F> inline void bg_emul_launch(const void* kernel, bg_emul_dim3 grid, bg_emul_dim3 block, F&& body) {
  gridDim = grid; blockDim = block;
  const unsigned nthreads = block.x;
  auto& coop = bg_emul_cooperative_kernels();
  bool fibers = std::find(coop.begin(), coop.end(), kernel) != coop.end();
  if (!fibers) {
    try {
      bg_emul_sequential = true;
      for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned t = 0; t < nthreads; ++t) { threadIdx = bg_emul_dim3(t, 0, 0); blockIdx = bg_emul_dim3(bx, by, bz); body(); }
      bg_emul_sequential = false;
      return;
    } catch (const bg_emul_needs_threads&) {
      bg_emul_sequential = false;
      coop.push_back(kernel);
    }
  }
  bg_emul_body = body;
  bg_emul_fibers.assign(nthreads, bg_emul_fiber());
  if (bg_emul_stacks.size() < nthreads * BG_EMUL_STACK) bg_emul_stacks.resize(nthreads * BG_EMUL_STACK);
  for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx) {
    bg_emul_block_bar = bg_emul_barrier();
    bg_emul_block_bar.parties = nthreads;
    bg_emul_warp_bars.assign((nthreads + 31) / 32, bg_emul_barrier());
    for (unsigned w = 0; w < bg_emul_warp_bars.size(); ++w) bg_emul_warp_bars[w].parties = std::min(32u, nthreads - w * 32);
    for (unsigned t = 0; t < nthreads; ++t) {
      bg_emul_fiber& f = bg_emul_fibers[t];
      f.done = false;
#if defined(__x86_64__)
      // initial frame: six callee-saved registers (zero) below the entry address; after `ret` the stack pointer is
      // 8 mod 16, as at any function entry
      uintptr_t top = (uintptr_t)(bg_emul_stacks.data() + (size_t)(t + 1) * BG_EMUL_STACK) & ~(uintptr_t)15;
      void** sp = (void**)(top - 8);
      *--sp = (void*)&bg_emul_fiber_entry;
      for (int r = 0; r < 6; ++r) *--sp = nullptr;
      f.sp = sp;
#else
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = bg_emul_stacks.data() + (size_t)t * BG_EMUL_STACK;
      f.ctx.uc_stack.ss_size = BG_EMUL_STACK;
      f.ctx.uc_link = &bg_emul_main_ctx;
      makecontext(&f.ctx, bg_emul_fiber_entry, 0);
#endif
    }
    blockIdx = bg_emul_dim3(bx, by, bz);
    for (bool any = true; any;) {
      any = false;
      for (unsigned t = 0; t < nthreads; ++t) {
        if (bg_emul_fibers[t].done) continue;
        bg_emul_cur = t;
        threadIdx = bg_emul_dim3(t, 0, 0);
        BG_EMUL_TO_FIBER(bg_emul_fibers[t]);
        any = any || !bg_emul_fibers[t].done;
      }
    }
  }
}
#define BG_LAUNCH(kernel, grid, block, stream, ...) \
  bg_emul_launch(reinterpret_cast<const void*>(&kernel), bg_emul_dim3(grid), bg_emul_dim3(block), [&] { kernel(__VA_ARGS__); })

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
// 4- / 8-wide vector access of bg_common.cuh (element loops here)
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static void ld(const float* p, long i, float* v) { for (int e = 0; e < 4; ++e) v[e] = p[i + e]; }
  static void st(float* p, long i, const float* v) { for (int e = 0; e < 4; ++e) p[i + e] = v[e]; }
};
template <> struct Vec<bf16> {
  static constexpr int N = 8;
  static void ld(const bf16* p, long i, float* v) { for (int e = 0; e < 8; ++e) v[e] = bg_emul_bf16_to_float(p[i + e]); }
  static void st(bf16* p, long i, const float* v) { for (int e = 0; e < 8; ++e) p[i + e] = bg_emul_float_to_bf16(v[e]); }
};
#define BG_DISPATCH_DTYPE(dtype, T, ...)                                        \
  do {                                                                          \
    if ((dtype) == BG_F32) { typedef float T; __VA_ARGS__; }                    \
    else if ((dtype) == BG_BF16) { typedef bf16 T; __VA_ARGS__; }               \
    else { bg_set_error("bad dtype %d", (int)(dtype)); return BG_ERR_UNSUPPORTED; } \
  } while (0)
