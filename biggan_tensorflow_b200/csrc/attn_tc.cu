// Flash-style fused attention core of self_attention_2 (ops.py:481-485): O = softmax(Q K^T) V without
// materialising the [N, N/4] score / probability tensors, on tcgen05 tensor cores with TMEM accumulators.
//   Q = hw_flatten(g) [B, Nq, dq], K = hw_flatten(maxpool(f)) [B, Nk, dq], V = hw_flatten(maxpool(h)) [B, Nk, dv]
// (bf16; dq zero-padded to a multiple of 16).  No 1/sqrt(d) scale (ops.py:481).
//
// Forward, one CTA per (image, 128 queries), two passes over the key blocks (Nk <= a few thousand):
//   pass 1: S_j = Q K_j^T (TMEM) -> row max;   pass 2: S_j again, P_j = exp(S_j - max) -> bf16 tile in
//   shared memory (128B-swizzled K-major) -> O += P_j V_j (TMEM), row sums.  No accumulator rescaling.
// Single-pass mode (bg_attn2_fwd_ws): softmax is invariant to the shift, so any m >= max_j s_ij works as long as nothing
// underflows.  A tiny pre-kernel gives the per-image Cauchy-Schwarz bound  m_b = max_i |q_i| * max_j |k_j| >= |s_ij|; when
// m_b <= 30 every row keeps exp(s_max - m_b) >= e^-60 and every element within e^-27 of its row maximum stays normal
// (fp32 / bf16 exponent range: e^-87), so pass 1 -- half of the Q K^T
// products and a quarter of the softmax warps' instructions -- is skipped and lse = m_b + log(sum).  Larger bounds (never seen
// with spectrally normalised projections of batch-normalised activations) take the two-pass path.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax / epilogue (one row per thread).
#include "tc_ptx.cuh"

#define AT_KB 128          // keys per block
struct AttnFwdParams {
  CUtensorMap qmap, kmap, vmap;    // (C, rows, B) bf16, box (64, 128, 1)
  int B, Nq, Nk, dq, dv;
  int nkb;                         // key blocks
  int nvblk;                       // 64-channel blocks of V
  bf16* o;
  float* lse;
  const float* bound;              // [B] upper bound of |s| per image, or nullptr (always two passes)
  float bound_max;                 // single-pass threshold
};

// 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 softmax / epilogue.  Two softmax warps
// share each TMEM lane quarter (one query row per lane) and split the 128 key columns of a block in halves;
// row max / row sums are combined through shared memory once per pass.  112 KB of shared memory and 256
// TMEM columns per CTA, so two CTAs are resident per SM and one CTA's softmax overlaps the other's MMAs.
__global__ void __launch_bounds__(320, 2) attn_fwd_kernel(const __grid_constant__ AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full, k_empty, v_full, v_empty, s_full, s_empty, p_full, p_empty, o_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float xch[2][2][128];          // [quantity][half][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t vbytes = (uint32_t)p.nvblk * 16384u;
  const uint32_t Qs = base, Ks = base + 16384u, Vs = Ks + 16384u, Ps = Vs + vbytes;
  const int b = blockIdx.y, q0 = blockIdx.x * 128;
  const float mb = p.bound ? __ldg(p.bound + b) : INFINITY;
  const bool single = mb <= p.bound_max;          // CTA-uniform
  const int first2 = single ? 0 : p.nkb;          // first iteration that forms P and P V
  const int T = first2 + p.nkb;

  if (warp == 0 && elect_one_sync()) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap);
    mbar_init(smem_u32(&q_full), 1);
    mbar_init(smem_u32(&k_full), 1); mbar_init(smem_u32(&k_empty), 1);
    mbar_init(smem_u32(&v_full), 1); mbar_init(smem_u32(&v_empty), 1);
    mbar_init(smem_u32(&s_full), 1); mbar_init(smem_u32(&s_empty), 8);
    mbar_init(smem_u32(&p_full), 8); mbar_init(smem_u32(&p_empty), 1); mbar_init(smem_u32(&o_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128u;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(smem_u32(&q_full), 16384u);
      tma_load_3d(Qs, &p.qmap, smem_u32(&q_full), 0, q0, b);
      for (int it = 0; it < T; ++it) {
        const int j = it >= p.nkb ? it - p.nkb : it;
        mbar_wait(smem_u32(&k_empty), (uint32_t)((it & 1) ^ 1));
        mbar_expect_tx(smem_u32(&k_full), 16384u);
        tma_load_3d(Ks, &p.kmap, smem_u32(&k_full), 0, j * AT_KB, b);
        if (it >= first2) {
          const int jj = it - first2;
          mbar_wait(smem_u32(&v_empty), (uint32_t)((jj & 1) ^ 1));
          mbar_expect_tx(smem_u32(&v_full), vbytes);
          for (int c = 0; c < p.nvblk; ++c)
            tma_load_3d(Vs + (uint32_t)c * 16384u, &p.vmap, smem_u32(&v_full), c * 64, j * AT_KB, b);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc_s = umma_idesc_bf16(128, AT_KB, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, p.dv, 0, 1);
      const int ksteps = p.dq / 16;
      mbar_wait(smem_u32(&q_full), 0);
      for (int it = 0; it <= T; ++it) {
        if (it < T) {
          mbar_wait(smem_u32(&k_full), (uint32_t)(it & 1));
          mbar_wait(smem_u32(&s_empty), (uint32_t)((it & 1) ^ 1));
          tc_fence_after();
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tS, umma_desc_kmajor_sw128(Qs + k * 32), umma_desc_kmajor_sw128(Ks + k * 32), idesc_s, k ? 1u : 0u);
          umma_commit(smem_u32(&s_full));
          umma_commit(smem_u32(&k_empty));
        }
        const int pv = it - 1 - first2;           // the P V product of the previous (pass-2) iteration
        if (pv >= 0) {
          mbar_wait(smem_u32(&v_full), (uint32_t)(pv & 1));
          mbar_wait(smem_u32(&p_full), (uint32_t)(pv & 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AT_KB / 16; ++k) {
            const uint64_t ad = umma_desc_kmajor_sw128(Ps + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u);
            const uint64_t bd = umma_desc_mnmajor_sw128(Vs + (uint32_t)k * 2048u, 16384u, 1024u);
            umma_bf16(tO, ad, bd, idesc_o, (pv | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&p_empty));
          umma_commit(smem_u32(&v_empty));
        }
      }
      umma_commit(smem_u32(&o_full));
    }
  } else {
    const int q = warp & 3;                       // TMEM lane quarter
    const int half = (warp - 2) >> 2;             // which 64 of the 128 key columns of a block
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    float m = -INFINITY, sum_true = 0.f;
    for (int it = 0; it < T; ++it) {
      const int j = it >= p.nkb ? it - p.nkb : it;
      const bool pass2 = it >= first2;
      if (it == first2) {
        if (single) m = mb;                       // the shift is the per-image bound: no pass 1
        else {                                    // pass 1 done: combine the two halves' row maxima
          xch[0][half][row] = m;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          m = fmaxf(xch[0][0][row], xch[0][1][row]);
        }
      }
      mbar_wait(smem_u32(&s_full), (uint32_t)(it & 1));
      tc_fence_after();
      if (pass2) mbar_wait(smem_u32(&p_empty), (uint32_t)(((it - first2) & 1) ^ 1));
      const bool kmask = (j + 1) * AT_KB > p.Nk;       // only the last key block can be partial (block-uniform)
      const float mL = m * LOG2E;
      for (int cc = 0; cc < 64; cc += 16) {
        const int c0 = half * 64 + cc;
        uint32_t r[16];
        tmem_ld_x16(tS + lane_base + (uint32_t)c0, r);
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
        if (kmask) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (j * AT_KB + c0 + e >= p.Nk) v[e] = -INFINITY;
        }
        if (!pass2) {
#pragma unroll
          for (int e = 0; e < 16; ++e) m = fmaxf(m, v[e]);
        } else {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float p0 = fast_exp2(fmaf(v[2 * e], LOG2E, -mL)), p1 = fast_exp2(fmaf(v[2 * e + 1], LOG2E, -mL));
            __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
            sum_true += p0 + p1;      // fp32 normaliser (the bf16 rounding of P is unbiased noise, as in flash attention)
            pk[e] = *reinterpret_cast<uint32_t*>(&h);
          }
          // P[row][c0..c0+15] -> K-major SWIZZLE_128B tile(s): 16-byte chunk index XOR (row & 7)
          const uint32_t kb = (uint32_t)(c0 >> 6), ch = (uint32_t)((c0 & 63) >> 3);
          const uint32_t rbase = Ps + kb * 16384u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
          const uint32_t a0 = rbase + ((ch ^ (uint32_t)(row & 7)) << 4), a1 = rbase + (((ch + 1) ^ (uint32_t)(row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
        }
      }
      tc_fence_before();
      if (pass2) fence_proxy_async_smem();
      __syncwarp();
      if (elect_one_sync()) {
        mbar_arrive(smem_u32(&s_empty));
        if (pass2) mbar_arrive(smem_u32(&p_full));
      }
    }
    // combine the halves' row sums
    xch[1][half][row] = sum_true;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum_true = xch[1][0][row] + xch[1][1][row];
    mbar_wait(smem_u32(&o_full), 0);
    tc_fence_after();
    const int qi = q0 + row;
    const float inv = 1.f / sum_true;
    const int csplit = ((p.dv / 16 + 1) / 2) * 16;
    const int cbeg = half ? csplit : 0, cend = half ? p.dv : csplit;
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(tO + lane_base + (uint32_t)c0, r);
      if (qi < p.Nq) {
        uint4 o0, o1;
        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h0[e] = __floats2bfloat162_rn(__uint_as_float(r[2 * e]) * inv, __uint_as_float(r[2 * e + 1]) * inv);
          h1[e] = __floats2bfloat162_rn(__uint_as_float(r[8 + 2 * e]) * inv, __uint_as_float(r[8 + 2 * e + 1]) * inv);
        }
        uint4* op = reinterpret_cast<uint4*>(p.o + ((long)b * p.Nq + qi) * p.dv + c0);
        op[0] = o0; op[1] = o1;
      }
    }
    if (half == 0 && qi < p.Nq) p.lse[(long)b * p.Nq + qi] = m + logf(sum_true);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled2 attn_encode() {
  static PFN_tmapEncodeTiled2 fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled2)f;
  }
  return fn;
}

// [B, rows, C] bf16 -> 3D map (C, rows, B), box (64, box_rows, 1), 128B swizzle, zero OOB fill
int bg_attn_make_map(CUtensorMap* m, const void* base, int B, int rows, int C, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)rows * C * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  PFN_tmapEncodeTiled2 enc = attn_encode();
  if (!enc) { bg_set_error("cuTensorMapEncodeTiled unavailable"); return BG_ERR_CUDA; }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bg_set_error("cuTensorMapEncodeTiled(attn) failed: %d (C=%d rows=%d B=%d)", (int)r, C, rows, B); return BG_ERR_CUDA; }
  return BG_OK;
}

// bound[b] = max_i |q_i| * max_j |k_j| (+ margin): one CTA per image, bf16x2 loads, fixed-order block maximum
__global__ void __launch_bounds__(256) attn_bound_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, float* __restrict__ bound,
                                                         int Nq, int Nk, int dq) {
  __shared__ float sh[2][8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx[2] = {0.f, 0.f};
  for (int t = 0; t < 2; ++t) {
    const bf16* base = (t ? k : q) + (long)b * (t ? Nk : Nq) * dq;
    const int rows = t ? Nk : Nq;
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
      const uint4* rp = reinterpret_cast<const uint4*>(base + (long)r * dq);
      float ss = 0.f;
      for (int c = 0; c < dq / 8; ++c) {
        const uint4 u = __ldg(rp + c);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h[e]); ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss)); }
      }
      mx[t] = fmaxf(mx[t], ss);
    }
    for (int o = 16; o > 0; o >>= 1) mx[t] = fmaxf(mx[t], __shfl_xor_sync(0xffffffffu, mx[t], o));
    if (lane == 0) sh[t][warp] = mx[t];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < 8; ++w) { a = fmaxf(a, sh[0][w]); c = fmaxf(c, sh[1][w]); }
    bound[b] = sqrtf(a * c) * 1.001f + 1e-3f;      // margin for the fp32 rounding of the norms and of the tensor-core dot products
  }
}

static int attn2_fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Nq, int Nk, int dq, int dv,
                          float* bound, void* stream);
extern "C" size_t bg_attn2_fwd_workspace_bytes(int B) { return (size_t)(B > 0 ? B : 0) * sizeof(float) + 16; }
extern "C" int bg_attn2_fwd_ws(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Nq, int Nk, int dq, int dv,
                               void* ws, size_t ws_bytes, void* stream) {
  float* bound = nullptr;
  static int sp_env = -1;                           // A/B knob: 0 = always two passes
  if (sp_env < 0) { const char* e = getenv("BG_ATTN_SINGLE_PASS"); sp_env = e ? atoi(e) : 1; }
  if (sp_env && ws && ws_bytes >= bg_attn2_fwd_workspace_bytes(B) && (((uintptr_t)ws) & 3) == 0 && dq % 8 == 0 &&
      ((((uintptr_t)q) | ((uintptr_t)k)) & 15) == 0)
    bound = (float*)ws;
  const int rc = attn2_fwd_impl(q, k, v, o, lse, B, Nq, Nk, dq, dv, bound, stream);
  static int dbg_env = -1;                          // diagnostics (eager mode only): print the largest bound of each call
  if (dbg_env < 0) { const char* e = getenv("BG_ATTN_DEBUG"); dbg_env = e ? atoi(e) : 0; }
  if (dbg_env && bound && rc == BG_OK) {
    float* h = (float*)malloc(sizeof(float) * B);
    if (h && cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess && cudaMemcpy(h, bound, sizeof(float) * B, cudaMemcpyDeviceToHost) == cudaSuccess) {
      float mx = 0.f;
      for (int i = 0; i < B; ++i) mx = h[i] > mx ? h[i] : mx;
      fprintf(stderr, "[bg_attn2_fwd_ws] B=%d dq=%d dv=%d max score bound %.3f (single pass up to 30)\n", B, dq, dv, mx);
    }
    free(h);
  }
  return rc;
}
extern "C" int bg_attn2_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Nq, int Nk, int dq, int dv,
                            void* stream) {
  return attn2_fwd_impl(q, k, v, o, lse, B, Nq, Nk, dq, dv, nullptr, stream);
}
static int attn2_fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, int B, int Nq, int Nk, int dq, int dv,
                          float* bound, void* stream) {
  BG_REQUIRE(bg_has_tensor_core_path(), BG_ERR_UNSUPPORTED, "attn2: tensor-core path unavailable on this device");
  BG_REQUIRE(dq % 16 == 0 && dq >= 16 && dq <= 64, BG_ERR_UNSUPPORTED, "attn2: dq=%d must be 16..64, multiple of 16", dq);
  BG_REQUIRE(dv % 16 == 0 && dv >= 16 && dv <= 128, BG_ERR_UNSUPPORTED, "attn2: dv=%d must be 16..128, multiple of 16", dv);
  BG_REQUIRE(B > 0 && Nq > 0 && Nk > 0 && B <= 65535, BG_ERR_SHAPE, "attn2: bad shape");
  static AttnFwdParams P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = bg_attn_make_map(&P.qmap, q, B, Nq, dq, 128))) return rc;
  if ((rc = bg_attn_make_map(&P.kmap, k, B, Nk, dq, AT_KB))) return rc;
  if ((rc = bg_attn_make_map(&P.vmap, v, B, Nk, dv, AT_KB))) return rc;
  P.B = B; P.Nq = Nq; P.Nk = Nk; P.dq = dq; P.dv = dv;
  P.nkb = (int)bg_cdiv(Nk, AT_KB);
  P.nvblk = (dv + 63) / 64;
  P.o = (bf16*)o; P.lse = lse;
  P.bound = bound; P.bound_max = 30.f;
  if (bound) {
    attn_bound_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const bf16*)q, (const bf16*)k, bound, Nq, Nk, dq);
    BG_LAUNCH_CHECK();
  }
  const size_t smem = 16384 + 16384 + (size_t)P.nvblk * 16384 + 32768 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  dim3 grid((unsigned)bg_cdiv(Nq, 128), (unsigned)B);
  attn_fwd_kernel<<<grid, 320, smem, (cudaStream_t)stream>>>(P);
  BG_LAUNCH_CHECK();
  return BG_OK;
}
