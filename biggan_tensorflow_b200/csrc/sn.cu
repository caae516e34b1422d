// Spectral norm (ops.py:718-747): fused one-step power iteration, scale/cast writer, backward.
// HBM-bound: W [K,C] fp32 is streamed once from HBM for both matvecs (the second use of each row
// tile hits L1/L2), once more for the scale pass.  All reductions are fixed-order (deterministic,
// so data-parallel replicas keep bit-identical u).
//
// Two front ends over the same device bodies: per-matrix entry points (bg_sn_power_iter / bg_sn_scale /
// bg_sn_bwd) and GROUPED entry points (bg_sn_group_*) that process every weight matrix of a network in
// 4 (forward) / 2 (backward) launches from a device-resident plan -- the per-step cost of ~60 small
// matrices is otherwise launch latency, which does not shrink with data parallelism.
#include "bg_common.cuh"

#define SN_MAX_RT 64

#define SN_TILE_FLOATS 8192           // 32 KB row chunk; two buffers (cp.async double buffering), 3 blocks per SM

// the staged path needs whole rows in the tile and 16-byte rows
#define SN_STAGED_MAX_C 4096          // column partials live in registers: 4 float4 per thread
static inline bool sn_staged(int C) { return (C % 4 == 0) && C <= SN_STAGED_MAX_C; }
static inline int sn_rows_per_block(int K, int C) {
  long rt;
  if (sn_staged(C)) {
    rt = SN_TILE_FLOATS / C;                         // one 32 KB chunk; the CTA count follows from sn_rows_per_cta
    if (rt > K) rt = K;
  } else {
    rt = bg_cdiv(K, 4L * bg_num_sms());
  }
  if (rt < 1) rt = 1;
  if (rt > SN_MAX_RT) rt = SN_MAX_RT;
  return (int)rt;
}
// rows per CTA: a staged CTA walks ~256 KB of W in RT-row chunks, keeping its column partials in registers, so a large
// matrix leaves a few hundred partial rows for the column reduction instead of thousands
static inline int sn_rows_per_cta(int K, int C, int RT) {
  if (!sn_staged(C)) return RT;
  long chunks = (256L * 1024 + (long)RT * C * 2) / ((long)RT * C * 4);
  if (chunks < 1) chunks = 1;
  return (int)(RT * chunks);
}
static inline size_t sn_matvec_smem(int C, int RT) { return sn_staged(C) ? ((size_t)2 * RT * C + C) * sizeof(float) : 0; }

// ---------------------------------------------------------------------------------------------
// device bodies (blk = block index within the matrix)
// ---------------------------------------------------------------------------------------------
// v[r] = W[r,:].u for the block's rows, then pcol[blk][c] = sum_r v[r] * W[r,c]
__device__ __forceinline__ void sn_matvec_body(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ vraw,
                                               float* __restrict__ pcol, int K, int C, int RT, int blk, float* vs) {
  const int r0 = blk * RT;
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
  float* dst = pcol + (long)blk * C;
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

// Same contraction with the CTA's rows streamed through shared memory in RT-row chunks, double-buffered with cp.async:
// the rows of a chunk are one contiguous piece of W, so the copy is a perfectly coalesced 16-byte stream that stays in
// flight while the previous chunk is contracted; both matvecs read shared memory, HBM sees W exactly once, and the
// column partials stay in registers across chunks.  Summation order per row is identical to sn_matvec_body.
__device__ __forceinline__ void sn_cp16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void sn_matvec_staged_body(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ vraw,
                                                      float* __restrict__ pcol, int K, int C, int RT, int RB, int blk, float* vs, float* tile) {
  const int rb0 = blk * RB;
  const int rb1 = min(K, rb0 + RB);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = (rb1 - rb0 + RT - 1) / RT;
  float4 acc[SN_STAGED_MAX_C / 1024];
#pragma unroll
  for (int g = 0; g < SN_STAGED_MAX_C / 1024; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto issue = [&](int ci) {
    const int r0 = rb0 + ci * RT;
    const int n4 = min(RT, rb1 - r0) * (C >> 2);
    const float* src = W + (long)r0 * C;
    float* dst = tile + (ci & 1) * (RT * C);
    for (int i = threadIdx.x; i < n4; i += 256) sn_cp16(dst + 4 * i, src + 4 * i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  float* us = tile + 2 * RT * C;                       // u staged once per CTA (read by every row of every chunk)
  for (int c = threadIdx.x * 4; c < C; c += 1024) *reinterpret_cast<float4*>(us + c) = __ldg(reinterpret_cast<const float4*>(u + c));
  for (int ci = 0; ci < nchunk; ++ci) {
    if (ci + 1 < nchunk) {
      issue(ci + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int r0 = rb0 + ci * RT;
    const int nr = min(RT, rb1 - r0);
    const float* buf = tile + (ci & 1) * (RT * C);
    for (int r = warp; r < nr; r += 8) {
      const float* row = buf + r * C;
      float s = 0.f;
      for (int c = lane * 4; c < C; c += 128) {
        float4 w4 = *reinterpret_cast<const float4*>(row + c);
        float4 u4 = *reinterpret_cast<const float4*>(us + c);
        s += w4.x * u4.x + w4.y * u4.y + w4.z * u4.z + w4.w * u4.w;
      }
      s = bg_warp_sum(s);
      if (lane == 0) { vs[r] = s; vraw[r0 + r] = s; }
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < SN_STAGED_MAX_C / 1024; ++g) {
      const int c = g * 1024 + threadIdx.x * 4;
      if (c < C) {
        float4 a = acc[g];
        for (int r = 0; r < nr; ++r) {
          float4 w4 = *reinterpret_cast<const float4*>(buf + r * C + c);
          const float vr = vs[r];
          a.x = fmaf(vr, w4.x, a.x); a.y = fmaf(vr, w4.y, a.y); a.z = fmaf(vr, w4.z, a.z); a.w = fmaf(vr, w4.w, a.w);
        }
        acc[g] = a;
      }
    }
    __syncthreads();
  }
  float* dst = pcol + (long)blk * C;
#pragma unroll
  for (int g = 0; g < SN_STAGED_MAX_C / 1024; ++g) {
    const int c = g * 1024 + threadIdx.x * 4;
    if (c < C) *reinterpret_cast<float4*>(dst + c) = acc[g];
  }
}

// uraw[c] = sum over the nblk partial rows; a block owns 32 columns, its 8 warps take every 8th row (128-byte coalesced
// reads) and are combined in fixed order
__device__ __forceinline__ void sn_colreduce_body(const float* __restrict__ pcol, float* __restrict__ uraw, int nblk, int C, int blk,
                                                  float (*sh)[32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blk * 32 + lane;
  float s = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int b = warp; b < nblk; b += 8) s += pcol[(long)b * C + c];
  }
  sh[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && c < C) {
    float t = sh[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) t += sh[w][lane];
    uraw[c] = t;
  }
}

// one block: normalise v, form u' = uraw * inv|v|, u_hat, sigma = u'.u_hat
__device__ __forceinline__ void sn_finalize_body(float* __restrict__ v, const float* __restrict__ uraw, float* __restrict__ u_out,
                                                 float* __restrict__ sigma, int K, int C, float* sh) {
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

// 32x32 tile (xt, yt) of tap z: out[(z*r_pad + r)*c_pad + c] and out_t[z][c][r]; 256 threads
template <typename T>
__device__ __forceinline__ void sn_scale_tile_body(const float* __restrict__ W, float sg, T* __restrict__ out, T* __restrict__ out_t,
                                                   int rows, int C, int r_pad, int c_pad, int z, int yt, int xt, float (*tile)[33]) {
  const long tapoff = (long)z * rows * C, tapoff_p = (long)z * r_pad * c_pad;
  const int r0 = yt * 32, c0 = xt * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < rows && c < C) {
      const float v = W[tapoff + (long)r * C + c] / sg;
      tile[i][tx] = v;
      bg_st(out, tapoff_p + (long)r * c_pad + c, v);
    }
  }
  if (!out_t) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < rows && c < C) bg_st(out_t, tapoff_p + (long)c * r_pad + r, tile[tx][i]);
  }
}

// bf16 fast path: 64x64 tile, 16-byte loads and 16-byte stores for both the [r][c] and the per-tap transposed [c][r] copy.
// Needs C % 4 == 0, c_pad % 8 == 0 and (with out_t) r_pad % 8 == 0; padded destinations are pre-zeroed by the caller and
// the tile holds zeros outside the matrix, so full 8-element groups may be stored.
static inline bool sn_scale_fast(int dtype, int C, int r_pad, int c_pad, const void* out, const void* out_t) {
  return dtype == BG_BF16 && C % 4 == 0 && c_pad % 8 == 0 && (!out_t || r_pad % 8 == 0) && (((uintptr_t)out | (uintptr_t)out_t) & 15) == 0;
}
__device__ __forceinline__ uint4 sn_pack8(const float* v) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
  return t;
}
#define SN_STRIP 1      // 64x64 tiles per CTA along C (a register-prefetching strip of 4 measured slower: 80 registers, 3 CTAs/SM)
__device__ __forceinline__ void sn_scale_tile64_body(const float* __restrict__ W, float sg, bf16* __restrict__ out, bf16* __restrict__ out_t,
                                                     int rows, int C, int r_pad, int c_pad, int z, int yt, int xs, float (*tile)[65]) {
  const long tapoff = (long)z * rows * C, tapoff_p = (long)z * r_pad * c_pad;
  const int r0 = yt * 64, c0 = xs * 64;
  const int t = threadIdx.x;
  // W * (1/sigma): one fp32 ulp from W / sigma, far below the bf16 rounding of the result (bf16 path only)
  const float inv = 1.f / sg;
  float4 w[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int r = r0 + p * 16 + (t >> 4), c = c0 + (t & 15) * 4;
    w[p] = (r < rows && c < C) ? __ldg(reinterpret_cast<const float4*>(W + tapoff + (long)r * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int i = p * 16 + (t >> 4), cc = (t & 15) * 4;
    tile[i][cc] = w[p].x * inv; tile[i][cc + 1] = w[p].y * inv; tile[i][cc + 2] = w[p].z * inv; tile[i][cc + 3] = w[p].w * inv;
  }
  __syncthreads();
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int idx = t + p * 256;
    const int i = idx >> 3, cc = (idx & 7) * 8;
    const int r = r0 + i, c = c0 + cc;
    if (r < rows && c < C) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tile[i][cc + e];
      *reinterpret_cast<uint4*>(out + tapoff_p + (long)r * c_pad + c) = sn_pack8(v);
    }
  }
  if (!out_t) return;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int idx = t + p * 256;
    const int rr = (idx & 7) * 8, ci = idx >> 3;
    const int r = r0 + rr, c = c0 + ci;
    if (r < rows && c < C) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = tile[rr + e][ci];
      *reinterpret_cast<uint4*>(out_t + tapoff_p + (long)c * r_pad + r) = sn_pack8(v);
    }
  }
}

__device__ __forceinline__ long sn_g_index(long i, int C, int rows, int r_pad, int g_ld) {
  const long k = i / C;
  return ((k / rows) * r_pad + (k % rows)) * g_ld + (i % C);
}

// part[blk] = sum over this block's strided slice of G.W   (nblk blocks of 256 threads)
__device__ __forceinline__ void sn_dot_body(const float* __restrict__ G, const float* __restrict__ W, float* __restrict__ part, long n,
                                            int C, int g_ld, int rows, int r_pad, int blk, int nblk, float* sh) {
  float s = 0.f;
  if (g_ld == C && r_pad == rows && (n & 3) == 0) {        // unpadded gradient layout: 16-byte streaming
    const long n4 = n >> 2;
    for (long i = (long)blk * blockDim.x + threadIdx.x; i < n4; i += (long)nblk * blockDim.x) {
      const float4 g = *reinterpret_cast<const float4*>(G + 4 * i), w = *reinterpret_cast<const float4*>(W + 4 * i);
      s += g.x * w.x + g.y * w.y + g.z * w.z + g.w * w.w;
    }
  } else {
    for (long i = (long)blk * blockDim.x + threadIdx.x; i < n; i += (long)nblk * blockDim.x)
      s = fmaf(G[sn_g_index(i, C, rows, r_pad, g_ld)], W[i], s);
  }
  s = bg_block_sum(s, sh);
  if (threadIdx.x == 0) part[blk] = s;
}

// dW = G/sigma - (<G,W>/sigma^2) v_hat[k] u_hat[c]
__device__ __forceinline__ void sn_bwd_apply_body(const float* __restrict__ G, const float* __restrict__ part, int npart,
                                                  const float* __restrict__ u_hat, const float* __restrict__ v_hat, float sg,
                                                  float* __restrict__ dW, int K, int C, int g_ld, int rows, int r_pad, int blk, int nblk,
                                                  float* sh) {
  float s = 0.f;
  for (int i = threadIdx.x; i < npart; i += blockDim.x) s += part[i];
  const float dot = bg_block_sum(s, sh);       // identical in every block: fixed order
  const float coef = dot / (sg * sg);
  const long n = (long)K * C;
  if (g_ld == C && r_pad == rows && (C & 3) == 0) {
    const long n4 = n >> 2;
    const int C4 = C >> 2;
#pragma unroll 4
    for (long i = (long)blk * blockDim.x + threadIdx.x; i < n4; i += (long)nblk * blockDim.x) {
      const int k = (int)(i / C4), c = (int)(i - (long)k * C4) * 4;
      const float4 g = *reinterpret_cast<const float4*>(G + 4 * i), u4 = *reinterpret_cast<const float4*>(u_hat + c);
      const float cv = coef * v_hat[k];
      *reinterpret_cast<float4*>(dW + 4 * i) = make_float4(g.x / sg - cv * u4.x, g.y / sg - cv * u4.y, g.z / sg - cv * u4.z, g.w / sg - cv * u4.w);
    }
    return;
  }
  for (long i = (long)blk * blockDim.x + threadIdx.x; i < n; i += (long)nblk * blockDim.x) {
    const int k = (int)(i / C), c = (int)(i - (long)k * C);
    dW[i] = G[sn_g_index(i, C, rows, r_pad, g_ld)] / sg - coef * v_hat[k] * u_hat[c];
  }
}

// ---------------------------------------------------------------------------------------------
// per-matrix kernels / entry points
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sn_matvec_kernel(const float* W, const float* u, float* vraw, float* pcol, int K, int C, int RT, int RB, int staged) {
  extern __shared__ float4 sn_dyn4[];
  __shared__ float vs[SN_MAX_RT];
  if (staged) sn_matvec_staged_body(W, u, vraw, pcol, K, C, RT, RB, blockIdx.x, vs, reinterpret_cast<float*>(sn_dyn4));
  else sn_matvec_body(W, u, vraw, pcol, K, C, RT, blockIdx.x, vs);
}
static int sn_matvec_attr() {
  static bool done = false;
  if (!done) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(sn_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * SN_TILE_FLOATS + SN_STAGED_MAX_C) * (int)sizeof(float)));
    done = true;
  }
  return BG_OK;
}
__global__ void __launch_bounds__(256) sn_colreduce_kernel(const float* pcol, float* uraw, int nblk, int C) {
  __shared__ float sh[8][32];
  sn_colreduce_body(pcol, uraw, nblk, C, blockIdx.x, sh);
}
__global__ void __launch_bounds__(1024) sn_finalize_kernel(float* v, const float* uraw, float* u_out, float* sigma, int K, int C) {
  __shared__ float sh[33];
  sn_finalize_body(v, uraw, u_out, sigma, K, C, sh);
}

extern "C" size_t bg_sn_workspace_bytes(int K, int C) {
  const int RT = sn_rows_per_block(K, C);
  const long nblk = bg_cdiv(K, sn_rows_per_cta(K, C, RT));
  size_t a = (size_t)(nblk + 1) * C * sizeof(float);     // power iteration: pcol + uraw
  size_t b = 1024 * sizeof(float);                       // backward: dot partials
  return (a > b ? a : b) + 256;
}

extern "C" int bg_sn_power_iter(const float* W, const float* u, float* u_out, float* v, float* sigma, int K, int C,
                                void* ws, size_t ws_bytes, void* stream) {
  BG_REQUIRE(K > 0 && C > 0, BG_ERR_SHAPE, "sn: bad shape %dx%d", K, C);
  BG_REQUIRE(ws && ws_bytes >= bg_sn_workspace_bytes(K, C), BG_ERR_WORKSPACE, "sn workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int RT = sn_rows_per_block(K, C);
  const int RB = sn_rows_per_cta(K, C, RT);
  const int nblk = (int)bg_cdiv(K, RB);
  float* pcol = (float*)ws;
  float* uraw = pcol + (long)nblk * C;
  int rc = sn_matvec_attr();
  if (rc) return rc;
  sn_matvec_kernel<<<nblk, 256, sn_matvec_smem(C, RT), st>>>(W, u, v, pcol, K, C, RT, RB, sn_staged(C) ? 1 : 0);
  BG_LAUNCH_CHECK();
  sn_colreduce_kernel<<<(unsigned)bg_cdiv(C, 32), 256, 0, st>>>(pcol, uraw, nblk, C);
  BG_LAUNCH_CHECK();
  sn_finalize_kernel<<<1, 1024, 0, st>>>(v, uraw, u_out, sigma, K, C);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

// out = W / sigma, plain contiguous case
template <typename T>
__global__ void sn_scale_kernel(const float* __restrict__ W, const float* __restrict__ sigma, T* __restrict__ out, long n) {
  const float sg = sigma[0];
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long stride = (long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 w4 = *reinterpret_cast<const float4*>(W + i);
    bg_st4(out, i, f4{w4.x / sg, w4.y / sg, w4.z / sg, w4.w / sg});
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {      // tail (n % 4)
    long j = (n & ~3L) + threadIdx.x;
    bg_st(out, j, W[j] / sg);
  }
}

// generic writer (padding and/or per-tap transposed copy); destinations with r_pad > rows or c_pad > C are
// pre-zeroed by the caller (zero-padded channels)
template <typename T>
__global__ void __launch_bounds__(256)
sn_scale_t_kernel(const float* W, const float* sigma, T* out, T* out_t, int rows, int C, int r_pad, int c_pad) {
  __shared__ float tile[32][33];
  sn_scale_tile_body<T>(W, sigma[0], out, out_t, rows, C, r_pad, c_pad, blockIdx.z, blockIdx.y, blockIdx.x, tile);
}

__global__ void __launch_bounds__(256)
sn_scale_t64_kernel(const float* W, const float* sigma, bf16* out, bf16* out_t, int rows, int C, int r_pad, int c_pad) {
  __shared__ float tile[64][65];
  sn_scale_tile64_body(W, sigma[0], out, out_t, rows, C, r_pad, c_pad, blockIdx.z, blockIdx.y, blockIdx.x, tile);
}

extern "C" int bg_sn_scale(const float* W, const float* sigma, void* out, void* out_t, int dtype, int taps, int rows,
                           int C, int r_pad, int c_pad, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long n = (long)taps * rows * C;
  BG_REQUIRE(n > 0 && c_pad >= C && r_pad >= rows, BG_ERR_SHAPE, "sn_scale: bad shape");
  if (!out_t && c_pad == C && r_pad == rows) {
    const int blocks = (int)min(bg_cdiv(n, 1024), (long)bg_num_sms() * 8);
    BG_DISPATCH_DTYPE(dtype, T, sn_scale_kernel<T><<<blocks, 256, 0, st>>>(W, sigma, (T*)out, n));
  } else if (sn_scale_fast(dtype, C, r_pad, c_pad, out, out_t)) {
    dim3 grid((unsigned)bg_cdiv(bg_cdiv(C, 64), SN_STRIP), (unsigned)bg_cdiv(rows, 64), taps);
    sn_scale_t64_kernel<<<grid, 256, 0, st>>>(W, sigma, (bf16*)out, (bf16*)out_t, rows, C, r_pad, c_pad);
  } else {
    dim3 grid((unsigned)bg_cdiv(C, 32), (unsigned)bg_cdiv(rows, 32), taps);
    BG_DISPATCH_DTYPE(dtype, T, sn_scale_t_kernel<T><<<grid, 256, 0, st>>>(W, sigma, (T*)out, (T*)out_t, rows, C, r_pad, c_pad));
  }
  BG_LAUNCH_CHECK();
  return BG_OK;
}

__global__ void __launch_bounds__(256)
sn_dot_partial_kernel(const float* G, const float* W, float* part, long n, int C, int g_ld, int rows, int r_pad) {
  __shared__ float sh[33];
  sn_dot_body(G, W, part, n, C, g_ld, rows, r_pad, blockIdx.x, gridDim.x, sh);
}
__global__ void __launch_bounds__(256)
sn_bwd_apply_kernel(const float* G, const float* part, int npart, const float* u_hat, const float* v_hat, const float* sigma, float* dW,
                    int K, int C, int g_ld, int rows, int r_pad) {
  __shared__ float sh[33];
  sn_bwd_apply_body(G, part, npart, u_hat, v_hat, sigma[0], dW, K, C, g_ld, rows, r_pad, blockIdx.x, gridDim.x, sh);
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

// ---------------------------------------------------------------------------------------------
// grouped front end: every matrix of a network per launch
// ---------------------------------------------------------------------------------------------
struct SnItemDev {
  bg_sn_item it;
  float* pcol; float* uraw; float* dotpart;
  int RT, RB, nblk, npart, napply;
  int sx, sy;            // scale tiles per tap (x over C, y over rows)
  int tile;              // scale tile edge: 64 (bf16 fast path) or 32
  int staged;            // matvec stages its row tile in shared memory
};
struct SnMap { int item, local; };

// plan layout (device): [SnItemDev x n][SnMap matvec][SnMap colred][SnMap scale][SnMap dot][SnMap apply][float workspace]
// totals (host, int[8]): n, T_matvec, T_colred, T_scale, T_dot, T_apply, off_maps (bytes), matvec dynamic smem (bytes)
struct SnPlanSizes { long t[5]; size_t off_maps, off_ws, total; };

static void sn_item_sizes(const bg_sn_item& it, SnItemDev* d) {
  d->it = it;
  d->RT = sn_rows_per_block(it.K, it.C);
  d->RB = sn_rows_per_cta(it.K, it.C, d->RT);
  d->nblk = (int)bg_cdiv(it.K, d->RB);
  d->staged = sn_staged(it.C) ? 1 : 0;
  d->tile = sn_scale_fast(it.dtype, it.C, it.r_pad, it.c_pad, it.out, it.out_t) ? 64 : 32;
  const long n = (long)it.K * it.C;
  d->npart = (int)min(bg_cdiv(n, 4096), 512L);
  d->napply = (int)min(bg_cdiv(n, 256 * 16), 2048L);
  if (d->napply < 1) d->napply = 1;
  d->sx = d->tile == 64 ? (int)bg_cdiv(bg_cdiv(it.C, 64), SN_STRIP) : (int)bg_cdiv(it.C, 32);
  d->sy = (int)bg_cdiv(it.rows, d->tile);
}

static SnPlanSizes sn_plan_sizes(const bg_sn_item* items, int n) {
  SnPlanSizes s;
  memset(&s, 0, sizeof(s));
  size_t wsf = 0;
  for (int i = 0; i < n; ++i) {
    SnItemDev d;
    sn_item_sizes(items[i], &d);
    s.t[0] += d.nblk;
    s.t[1] += bg_cdiv(items[i].C, 32);
    s.t[2] += (long)d.sx * d.sy * items[i].taps;
    s.t[3] += d.npart;
    s.t[4] += d.napply;
    wsf += (size_t)(d.nblk + 1) * items[i].C + d.npart + 8;
    wsf = (wsf + 3) / 4 * 4;
  }
  s.off_maps = (sizeof(SnItemDev) * n + 255) / 256 * 256;
  s.off_ws = (s.off_maps + sizeof(SnMap) * (s.t[0] + s.t[1] + s.t[2] + s.t[3] + s.t[4]) + 255) / 256 * 256;
  s.total = s.off_ws + wsf * sizeof(float) + 256;
  return s;
}

extern "C" size_t bg_sn_group_plan_bytes(const bg_sn_item* items, int n) { return sn_plan_sizes(items, n).total; }

// builds the plan on the host and copies it to plan_dev (synchronous; done once at build time)
extern "C" int bg_sn_group_plan(const bg_sn_item* items, int n, void* plan_dev, size_t plan_bytes, int* totals) {
  BG_REQUIRE(n > 0, BG_ERR_SHAPE, "sn_group: empty");
  const SnPlanSizes sz = sn_plan_sizes(items, n);
  BG_REQUIRE(plan_dev && plan_bytes >= sz.total, BG_ERR_WORKSPACE, "sn_group: plan buffer too small");
  char* host = (char*)malloc(sz.off_ws);
  BG_REQUIRE(host != nullptr, BG_ERR_CUDA, "sn_group: out of host memory");
  memset(host, 0, sz.off_ws);
  SnItemDev* its = (SnItemDev*)host;
  SnMap* maps = (SnMap*)(host + sz.off_maps);
  SnMap* m[5];
  m[0] = maps;
  for (int k = 1; k < 5; ++k) m[k] = m[k - 1] + sz.t[k - 1];
  long cur[5] = {0, 0, 0, 0, 0};
  float* ws = (float*)((char*)plan_dev + sz.off_ws);
  size_t wsf = 0;
  for (int i = 0; i < n; ++i) {
    SnItemDev* d = &its[i];
    sn_item_sizes(items[i], d);
    if (!(items[i].K > 0 && items[i].C > 0 && items[i].taps > 0 && items[i].K == items[i].taps * items[i].rows &&
          items[i].r_pad >= items[i].rows && items[i].c_pad >= items[i].C)) {
      free(host);
      bg_set_error("sn_group: bad item %d", i);
      return BG_ERR_SHAPE;
    }
    d->pcol = ws + wsf;
    d->uraw = d->pcol + (size_t)d->nblk * items[i].C;
    d->dotpart = d->uraw + items[i].C;
    wsf += (size_t)(d->nblk + 1) * items[i].C + d->npart + 8;
    wsf = (wsf + 3) / 4 * 4;
    const long cnt[5] = {d->nblk, bg_cdiv(items[i].C, 32), (long)d->sx * d->sy * items[i].taps, d->npart, d->napply};
    for (int k = 0; k < 5; ++k)
      for (long j = 0; j < cnt[k]; ++j) { m[k][cur[k]].item = i; m[k][cur[k]].local = (int)j; cur[k]++; }
  }
  cudaError_t e = cudaMemcpy(plan_dev, host, sz.off_ws, cudaMemcpyHostToDevice);
  free(host);
  BG_CHECK_CUDA(e);
  totals[0] = n;
  for (int k = 0; k < 5; ++k) totals[1 + k] = (int)sz.t[k];
  totals[6] = (int)sz.off_maps;
  size_t dyn = 0;
  for (int i = 0; i < n; ++i) dyn = max(dyn, sn_matvec_smem(items[i].C, sn_rows_per_block(items[i].K, items[i].C)));
  totals[7] = (int)dyn;                // dynamic shared memory of the grouped matvec launch
  return BG_OK;
}

__global__ void __launch_bounds__(256, 3) sng_matvec_kernel(const SnItemDev* its, const SnMap* map) {
  extern __shared__ float4 sn_dyn4[];
  __shared__ float vs[SN_MAX_RT];
  const SnMap mp = map[blockIdx.x];
  const SnItemDev& d = its[mp.item];
  if (d.staged) sn_matvec_staged_body(d.it.W, d.it.u, d.it.v, d.pcol, d.it.K, d.it.C, d.RT, d.RB, mp.local, vs, reinterpret_cast<float*>(sn_dyn4));
  else sn_matvec_body(d.it.W, d.it.u, d.it.v, d.pcol, d.it.K, d.it.C, d.RT, mp.local, vs);
}
__global__ void __launch_bounds__(256) sng_colreduce_kernel(const SnItemDev* its, const SnMap* map) {
  __shared__ float sh[8][32];
  const SnMap mp = map[blockIdx.x];
  const SnItemDev& d = its[mp.item];
  sn_colreduce_body(d.pcol, d.uraw, d.nblk, d.it.C, mp.local, sh);
}
__global__ void __launch_bounds__(1024) sng_finalize_kernel(const SnItemDev* its) {
  __shared__ float sh[33];
  const SnItemDev& d = its[blockIdx.x];
  sn_finalize_body(d.it.v, d.uraw, d.it.u, d.it.sigma, d.it.K, d.it.C, sh);
}
__global__ void __launch_bounds__(256) sng_scale_kernel(const SnItemDev* its, const SnMap* map) {
  __shared__ float tile_raw[64 * 65];
  float (*tile)[33] = reinterpret_cast<float (*)[33]>(tile_raw);
  const SnMap mp = map[blockIdx.x];
  const SnItemDev& d = its[mp.item];
  const int per_tap = d.sx * d.sy;
  const int z = mp.local / per_tap, r = mp.local - z * per_tap;
  const float sg = d.it.sigma[0];
  if (d.tile == 64)
    sn_scale_tile64_body(d.it.W, sg, (bf16*)d.it.out, (bf16*)d.it.out_t, d.it.rows, d.it.C, d.it.r_pad, d.it.c_pad, z, r / d.sx, r % d.sx,
                         reinterpret_cast<float (*)[65]>(tile_raw));
  else if (d.it.dtype == BG_F32)
    sn_scale_tile_body<float>(d.it.W, sg, (float*)d.it.out, (float*)d.it.out_t, d.it.rows, d.it.C, d.it.r_pad, d.it.c_pad, z, r / d.sx, r % d.sx, tile);
  else
    sn_scale_tile_body<bf16>(d.it.W, sg, (bf16*)d.it.out, (bf16*)d.it.out_t, d.it.rows, d.it.C, d.it.r_pad, d.it.c_pad, z, r / d.sx, r % d.sx, tile);
}
__global__ void __launch_bounds__(256) sng_dot_kernel(const SnItemDev* its, const SnMap* map) {
  __shared__ float sh[33];
  const SnMap mp = map[blockIdx.x];
  const SnItemDev& d = its[mp.item];
  sn_dot_body(d.it.dWn, d.it.W, d.dotpart, (long)d.it.K * d.it.C, d.it.C, d.it.g_ld, d.it.rows, d.it.r_pad, mp.local, d.npart, sh);
}
__global__ void __launch_bounds__(256) sng_apply_kernel(const SnItemDev* its, const SnMap* map) {
  __shared__ float sh[33];
  const SnMap mp = map[blockIdx.x];
  const SnItemDev& d = its[mp.item];
  sn_bwd_apply_body(d.it.dWn, d.dotpart, d.npart, d.it.u, d.it.v, d.it.sigma[0], d.it.dW, d.it.K, d.it.C, d.it.g_ld, d.it.rows, d.it.r_pad,
                    mp.local, d.napply, sh);
}

extern "C" int bg_sn_group_forward(const void* plan_dev, const int* totals, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const SnItemDev* its = (const SnItemDev*)plan_dev;
  const SnMap* m0 = (const SnMap*)((const char*)plan_dev + totals[6]);
  const SnMap* m1 = m0 + totals[1];
  const SnMap* m2 = m1 + totals[2];
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(sng_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * SN_TILE_FLOATS + SN_STAGED_MAX_C) * (int)sizeof(float)));
    attr_set = true;
  }
  sng_matvec_kernel<<<totals[1], 256, (size_t)totals[7], st>>>(its, m0);
  BG_LAUNCH_CHECK();
  sng_colreduce_kernel<<<totals[2], 256, 0, st>>>(its, m1);
  BG_LAUNCH_CHECK();
  sng_finalize_kernel<<<totals[0], 1024, 0, st>>>(its);
  BG_LAUNCH_CHECK();
  sng_scale_kernel<<<totals[3], 256, 0, st>>>(its, m2);
  BG_LAUNCH_CHECK();
  return BG_OK;
}

extern "C" int bg_sn_group_backward(const void* plan_dev, const int* totals, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const SnItemDev* its = (const SnItemDev*)plan_dev;
  const SnMap* m0 = (const SnMap*)((const char*)plan_dev + totals[6]);
  const SnMap* m3 = m0 + totals[1] + totals[2] + totals[3];
  const SnMap* m4 = m3 + totals[4];
  sng_dot_kernel<<<totals[4], 256, 0, st>>>(its, m3);
  BG_LAUNCH_CHECK();
  sng_apply_kernel<<<totals[5], 256, 0, st>>>(its, m4);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
