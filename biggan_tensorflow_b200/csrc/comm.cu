// One-shot all-reduce over NVLink peer memory for the SMALL exchanges of the data-parallel step (collective C2):
// the cross-replica BatchNorm statistics -- [sum x, sum x^2] forward, [sum dy, sum dy xhat] backward, 2C doubles <= 24 KB,
// 33 of them per training iteration.  An NCCL all-reduce of that size is pure launch + protocol latency; here every rank
// stores its vector straight into a slot of every peer's buffer (P2P stores through NVSwitch), raises a flag, waits for
// the flags of all peers in its OWN buffer and sums the slots in rank order.  One kernel, one CTA, no host involvement,
// CUDA-graph capturable, and -- because every rank adds the same values in the same order -- the result is bit-identical
// on all ranks (replica-consistent BN statistics without a broadcast).
//
// Buffers are plain cudaMalloc allocations exported with cudaIpcGetMemHandle; the handles travel through the caller's
// process group (torch.distributed all_gather, host side) and are opened with cudaIpcOpenMemHandle.
//
// Protocol (per rank; `seq` lives in device memory so that graph replays advance it):
//   seq += 1; set = seq & 1
//   data[set][rank][0..n) of EVERY peer <- my vector;  fence.sys;  flag[set][rank] of every peer <- seq  (release, sys scope)
//   wait until flag[set][r] of my own buffer >= seq for all r (acquire);  out[i] = sum_r data[set][r][i]  (rank order)
// Two slot sets suffice: a rank can only start call seq+1 after it saw every peer's flag of call seq, and a peer raises that
// flag after it finished reading call seq-1 -- so nobody still reads set (seq+1)&1 when it is overwritten.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "bg_common.cuh"

#define BGC_MAX_WORLD 8
#define BGC_MAX_N 4096                       // doubles per vector (2C, C <= 2048)

struct BgCommBuf {                           // layout of every rank's exported buffer
  double data[2][BGC_MAX_WORLD][BGC_MAX_N];
  unsigned long long flag[2][BGC_MAX_WORLD];
};

struct BgComm {
  int rank, world;
  BgCommBuf* local;
  BgCommBuf* peer[BGC_MAX_WORLD];            // peer[rank] == local
  unsigned long long* seq;                   // device counter
  BgCommBuf** peer_dev;                      // device copy of peer[]
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// out may alias in.  1 CTA of 1024 threads.
__global__ void __launch_bounds__(1024) comm_allreduce_f64_kernel(BgCommBuf* const* __restrict__ peers, BgCommBuf* local,
                                                                  unsigned long long* seq_p, const double* in, double* out,
                                                                  int n, int rank, int world) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = *seq_p + 1ull;
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int set = (int)(seq & 1ull);
  // 1. push my vector into slot [set][rank] of every rank's buffer (own buffer included: one code path)
  for (int idx = threadIdx.x; idx < n * world; idx += blockDim.x) {
    const int p = idx / n, i = idx - p * n;
    peers[p]->data[set][rank][i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. signal, 3. wait
  if (threadIdx.x < world) {
    st_release_sys(&peers[threadIdx.x]->flag[set][rank], seq);
    const unsigned long long* f = &local->flag[set][threadIdx.x];
    while (ld_acquire_sys(f) < seq) { __nanosleep(40); }
  }
  __syncthreads();
  // 4. reduce in rank order (volatile loads: the slots were written by other devices, never serve them from L1)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += ld_volatile_f64(&local->data[set][r][i]);
    out[i] = s;
  }
  if (threadIdx.x == 0) *seq_p = seq;
}

extern "C" size_t bg_comm_buffer_bytes(void) { return sizeof(BgCommBuf); }

// Allocates this rank's buffer on the current device, zeroes it and returns its IPC handle (64 bytes).
extern "C" int bg_comm_local_alloc(void** local_out, void* handle64_out) {
  void* p = nullptr;
  BG_CHECK_CUDA(cudaMalloc(&p, sizeof(BgCommBuf)));
  BG_CHECK_CUDA(cudaMemset(p, 0, sizeof(BgCommBuf)));
  BG_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  BG_CHECK_CUDA(cudaIpcGetMemHandle(&h, p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64_out, &h, 64);
  *local_out = p;
  return BG_OK;
}

// handles: world x 64 bytes in rank order (host memory).  Opens every peer's buffer; returns an opaque communicator.
extern "C" int bg_comm_open(int rank, int world, const void* handles, void* local, void** comm_out) {
  BG_REQUIRE(world >= 1 && world <= BGC_MAX_WORLD && rank >= 0 && rank < world, BG_ERR_SHAPE, "bg_comm_open: rank %d of %d (max %d ranks)",
             rank, world, BGC_MAX_WORLD);
  BgComm* c = (BgComm*)calloc(1, sizeof(BgComm));
  c->rank = rank; c->world = world; c->local = (BgCommBuf*)local;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { c->peer[r] = c->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    BG_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = (BgCommBuf*)p;
  }
  BG_CHECK_CUDA(cudaMalloc((void**)&c->seq, sizeof(unsigned long long)));
  BG_CHECK_CUDA(cudaMemset(c->seq, 0, sizeof(unsigned long long)));
  BG_CHECK_CUDA(cudaMalloc((void**)&c->peer_dev, sizeof(BgCommBuf*) * BGC_MAX_WORLD));
  BG_CHECK_CUDA(cudaMemcpy(c->peer_dev, c->peer, sizeof(BgCommBuf*) * BGC_MAX_WORLD, cudaMemcpyHostToDevice));
  BG_CHECK_CUDA(cudaDeviceSynchronize());
  *comm_out = c;
  return BG_OK;
}

// In-place capable sum over ranks of n doubles (n <= 4096).  Every rank must issue the same sequence of calls.
extern "C" int bg_comm_allreduce_f64(void* comm, const double* in, double* out, int n, void* stream) {
  BgComm* c = (BgComm*)comm;
  BG_REQUIRE(c != nullptr, BG_ERR_SHAPE, "bg_comm_allreduce_f64: null communicator");
  BG_REQUIRE(n >= 1 && n <= BGC_MAX_N, BG_ERR_SHAPE, "bg_comm_allreduce_f64: n = %d (max %d)", n, BGC_MAX_N);
  comm_allreduce_f64_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(c->peer_dev, c->local, c->seq, in, out, n, c->rank, c->world);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_comm_destroy(void* comm) {
  BgComm* c = (BgComm*)comm;
  if (!c) return BG_OK;
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->seq);
  cudaFree(c->peer_dev);
  cudaFree(c->local);
  free(c);
  return BG_OK;
}
