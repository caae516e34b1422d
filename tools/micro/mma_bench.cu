// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per M=128 x N x K=16 bf16 MMA issued back to back by one thread,
// as a function of N and of how many independent TMEM accumulators the stream alternates between.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../biggan_tensorflow_b200/csrc mma_bench.cu -o mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_ptx.cuh"

template <int mode>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, int nacc, int mnmajor, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && (mode ? elect_one_sync() : (uint32_t)(lane == 0))) {
    const uint32_t idesc = umma_idesc_bf16(128, N, mnmajor, mnmajor);
    const uint32_t sa = base, sb = base + 32768u;
    // warm-up
    for (int i = 0; i < 8; ++i) umma_bf16(tmem, umma_desc_kmajor_sw128(sa), umma_desc_kmajor_sw128(sb), idesc, 1u);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
      const uint32_t k = (uint32_t)(i & 3);
      const uint64_t ad = mnmajor ? umma_desc_mnmajor_sw128(sa + k * 2048, 8192, 1024) : umma_desc_kmajor_sw128(sa + k * 32);
      const uint64_t bd = mnmajor ? umma_desc_mnmajor_sw128(sb + k * 2048, 8192, 1024) : umma_desc_kmajor_sw128(sb + k * 32);
      umma_bf16(d, ad, bd, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 512;
  printf("M=128 K=16 bf16 tcgen05.mma, %d back-to-back issues by one thread; cycles per MMA: issue / complete\n", iters);
  for (int mode = 0; mode < 2; ++mode)
  for (int mn = 0; mn < 2; ++mn)
    for (int N : {16, 48, 96, 128, 256})
      for (int nacc : {1, 2}) {
        if (nacc * N > 512) continue;
        for (int grid : {148}) {
          if (mode) mma_rate_kernel<1><<<grid, 128, 97 * 1024>>>(N, iters, nacc, mn, out);
          else mma_rate_kernel<0><<<grid, 128, 97 * 1024>>>(N, iters, nacc, mn, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          printf("%s %s N=%3d nacc=%d grid=%3d: issue %.1f  complete %.1f  (floor N/2 = %.0f)\n", mode ? "elect " : "lane==0", mn ? "MN-major" : "K-major ", N, nacc, grid,
                 (double)out[0] / iters, (double)out[1] / iters, N / 2.0);
        }
      }
  return 0;
}
