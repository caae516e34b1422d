// Backward of the fused attention core (ops.py:481-485) on tcgen05: recomputes S = Q K^T and
// P = exp(S - lse) tile by tile (nothing of size [N, N/4] is ever stored), with
//   dP = dO V^T,  dS = P o (dP - delta),  delta_i = dO_i . O_i,
//   dV = P^T dO,  dK = dS^T Q   (kernel attn_bwd_dkdv: one CTA per (image, 128 keys), loop over query tiles)
//   dQ = dS K                   round 2: the SAME kernel also forms this key block's contribution dS K for every query tile
//                               (one more MMA group on the dS tile it already holds; fp32 partial [key block][B, Nq, dq] in the
//                               workspace) and attn_dq_reduce_kernel sums the <= 8 partials in fixed order -- the second kernel
//                               that recomputed S, P and dP for dQ (attn_bwd_dq, kept for callers without workspace) is not run.
// Every operand transposition is done by the UMMA descriptors: the same 128B-swizzled shared-memory tile
// [rows][64 ch] is read K-major (contract over channels) or MN-major (contract over rows).
// Deterministic: no atomics; each output element is produced by exactly one CTA.
#include "tc_ptx.cuh"

#define AB_T 128
int bg_attn_make_map(CUtensorMap* m, const void* base, int B, int rows, int C, int box_rows);

struct AttnBwdParams {
  CUtensorMap qmap, kmap, vmap, domap;
  int B, Nq, Nk, dq, dv, nqt, nkb, nvblk;
  const float* lse;
  const float* delta;
  bf16* dq_out;
  bf16* dk_out;
  bf16* dv_out;
  float* dq_part;            // [nkb][B][Nq][dq] fp32 partial dQ (NULL: the dkdv kernel does not form dQ)
};

// p / ds for 16 consecutive keys of one query row -> bf16 K-major SWIZZLE_128B tiles
__device__ __forceinline__ void store_row16(uint32_t tile, int row, int c0, const uint32_t* pk) {
  const uint32_t kb = (uint32_t)(c0 >> 6), ch = (uint32_t)((c0 & 63) >> 3);
  const uint32_t rbase = tile + kb * 16384u + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u;
  const uint32_t a0 = rbase + ((ch ^ (uint32_t)(row & 7)) << 4), a1 = rbase + (((ch + 1) ^ (uint32_t)(row & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
}

__device__ __forceinline__ void store_acc_bf16(bf16* dst, const uint32_t* r) {
  uint4 o0, o1;
  __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&o0);
  __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    h0[e] = __floats2bfloat162_rn(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
    h1[e] = __floats2bfloat162_rn(__uint_as_float(r[8 + 2 * e]), __uint_as_float(r[8 + 2 * e + 1]));
  }
  uint4* op = reinterpret_cast<uint4*>(dst);
  op[0] = o0; op[1] = o1;
}

// S = Q K^T and dP = dO V^T for one (query tile, key block) pair
__device__ __forceinline__ void issue_s_dp(uint32_t tS, uint32_t tdP, uint32_t Qs, uint32_t Ks, uint32_t dOs, uint32_t Vs, int dq, int dv) {
  const uint32_t idesc = umma_idesc_bf16(128, AB_T, 0, 0);
  for (int k = 0; k < dq / 16; ++k)
    umma_bf16(tS, umma_desc_kmajor_sw128(Qs + k * 32), umma_desc_kmajor_sw128(Ks + k * 32), idesc, k ? 1u : 0u);
  for (int k = 0; k < dv / 16; ++k) {
    const uint32_t off = (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u;
    umma_bf16(tdP, umma_desc_kmajor_sw128(dOs + off), umma_desc_kmajor_sw128(Vs + off), idesc, k ? 1u : 0u);
  }
}

// ------------------------------------------------------------------------------------------------
// 18 warps: TMA producer, MMA issuer, and SIXTEEN softmax warps -- four per TMEM lane quarter, each taking 32 of the 128 key
// columns of a block.  The P / dS production (exp, two bf16 packs, swizzled stores per score) is what bounds this kernel
// (12-18 % tensor-pipe activity); with 8 warps it took ~2x the MMA time of a (query tile, key block) pair.
#define AB_SW 16                                   // softmax warps
__global__ void __launch_bounds__(64 + 32 * AB_SW, 1) attn_bwd_dkdv_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t kv_full, qd_full[2], qd_empty[2], s_full, s_empty, pds_full, pds_empty, acc_full;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t vbytes = (uint32_t)p.nvblk * 16384u;
  const uint32_t Ks = base, Vs = Ks + 16384u, Qs = Vs + vbytes, dOs = Qs + 2u * 16384u, Ps = dOs + 2u * vbytes, dSs = Ps + 32768u;
  const int b = blockIdx.y, j = blockIdx.x;

  if (warp == 0 && elect_one_sync()) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); tma_prefetch_desc(&p.domap);
    mbar_init(smem_u32(&kv_full), 1);
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&qd_full[i]), 1); mbar_init(smem_u32(&qd_empty[i]), 1); }
    mbar_init(smem_u32(&s_full), 1); mbar_init(smem_u32(&s_empty), AB_SW);
    mbar_init(smem_u32(&pds_full), AB_SW); mbar_init(smem_u32(&pds_empty), 1); mbar_init(smem_u32(&acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128u, tdV = tmem + 256u, tdK = tmem + 384u, tdQ = tmem + 448u;   // dq <= 64 columns each

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(smem_u32(&kv_full), 16384u + vbytes);
      tma_load_3d(Ks, &p.kmap, smem_u32(&kv_full), 0, j * AB_T, b);
      for (int c = 0; c < p.nvblk; ++c) tma_load_3d(Vs + (uint32_t)c * 16384u, &p.vmap, smem_u32(&kv_full), c * 64, j * AB_T, b);
      for (int i = 0; i < p.nqt; ++i) {
        const int s = i & 1;
        mbar_wait(smem_u32(&qd_empty[s]), (uint32_t)(((i >> 1) & 1) ^ 1));
        mbar_expect_tx(smem_u32(&qd_full[s]), 16384u + vbytes);
        tma_load_3d(Qs + (uint32_t)s * 16384u, &p.qmap, smem_u32(&qd_full[s]), 0, i * AB_T, b);
        for (int c = 0; c < p.nvblk; ++c)
          tma_load_3d(dOs + (uint32_t)s * vbytes + (uint32_t)c * 16384u, &p.domap, smem_u32(&qd_full[s]), c * 64, i * AB_T, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc_dv = umma_idesc_bf16(128, p.dv, 1, 1);
      const uint32_t idesc_dk = umma_idesc_bf16(128, p.dq, 1, 1);
      const uint32_t idesc_dq = umma_idesc_bf16(128, p.dq, 0, 1);
      const bool want_dq = p.dq_part != nullptr;
      mbar_wait(smem_u32(&kv_full), 0);
      for (int i = 0; i <= p.nqt; ++i) {
        if (i < p.nqt) {
          const int s = i & 1;
          mbar_wait(smem_u32(&qd_full[s]), (uint32_t)((i >> 1) & 1));
          mbar_wait(smem_u32(&s_empty), (uint32_t)((i & 1) ^ 1));
          tc_fence_after();
          issue_s_dp(tS, tdP, Qs + (uint32_t)s * 16384u, Ks, dOs + (uint32_t)s * vbytes, Vs, p.dq, p.dv);
          umma_commit(smem_u32(&s_full));
        }
        const int pi = i - 1;
        if (pi >= 0) {
          const int s = pi & 1;
          mbar_wait(smem_u32(&pds_full), (uint32_t)(pi & 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k) {      // contract over the 128 queries of the tile
            const uint64_t aP = umma_desc_mnmajor_sw128(Ps + (uint32_t)k * 2048u, 16384u, 1024u);
            const uint64_t bdO = umma_desc_mnmajor_sw128(dOs + (uint32_t)s * vbytes + (uint32_t)k * 2048u, 16384u, 1024u);
            umma_bf16(tdV, aP, bdO, idesc_dv, (pi | k) != 0 ? 1u : 0u);
            const uint64_t adS = umma_desc_mnmajor_sw128(dSs + (uint32_t)k * 2048u, 16384u, 1024u);
            const uint64_t bQ = umma_desc_mnmajor_sw128(Qs + (uint32_t)s * 16384u + (uint32_t)k * 2048u, 16384u, 1024u);
            umma_bf16(tdK, adS, bQ, idesc_dk, (pi | k) != 0 ? 1u : 0u);
          }
          if (want_dq) {
#pragma unroll
            for (int k = 0; k < AB_T / 16; ++k) {    // dQ(tile pi) contribution of this key block: contract over its 128 keys
              const uint64_t adS = umma_desc_kmajor_sw128(dSs + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u);
              const uint64_t bK = umma_desc_mnmajor_sw128(Ks + (uint32_t)k * 2048u, 16384u, 1024u);
              umma_bf16(tdQ, adS, bK, idesc_dq, k != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&pds_empty));
          umma_commit(smem_u32(&qd_empty[s]));
        }
      }
      umma_commit(smem_u32(&acc_full));
    }
  } else {
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;          // AB_SW / 4 warps per TMEM lane quarter split the 128 key columns
    constexpr int NPART = AB_SW / 4, PCOLS = AB_T / NPART;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    // accumulator drains (dQ partials, dV, dK): 16-column chunks dealt round-robin to the warps of a lane quarter
    auto drain_dq = [&](int qp) {
      for (int c0 = part * 16; c0 < p.dq; c0 += NPART * 16) {
        uint32_t r[16];
        tmem_ld_x16(tdQ + lane_base + (uint32_t)c0, r);
        if (qp < p.Nq) {
          float4* o = reinterpret_cast<float4*>(p.dq_part + (((long)j * p.B + b) * p.Nq + qp) * p.dq + c0);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]), __uint_as_float(r[4 * e + 3]));
        }
      }
    };
    for (int i = 0; i < p.nqt; ++i) {
      const int qi = i * AB_T + row;
      const bool qv = qi < p.Nq;
      const float nlL = qv ? -p.lse[(long)b * p.Nq + qi] * LOG2E : -INFINITY;      // exp2(s*log2e - lse*log2e); 0 for rows past Nq
      const float dl = qv ? p.delta[(long)b * p.Nq + qi] : 0.f;
      const bool kmask = (j + 1) * AB_T > p.Nk;                                     // partial last key block (block-uniform)
      mbar_wait(smem_u32(&s_full), (uint32_t)(i & 1));
      tc_fence_after();
      mbar_wait(smem_u32(&pds_empty), (uint32_t)((i & 1) ^ 1));
      if (p.dq_part != nullptr && i > 0) {           // the MMAs of tile i - 1 are complete: its dQ partial leaves TMEM before
        tc_fence_after();                            // tile i's dS (and with it the next dQ MMA group) is released
        drain_dq((i - 1) * AB_T + row);
        tc_fence_before();
      }
      for (int cc = 0; cc < PCOLS; cc += 16) {
        const int c0 = part * PCOLS + cc;
        uint32_t rs[16], rp[16], pk[8], dk[8];
        tmem_ld_x16(tS + lane_base + (uint32_t)c0, rs);
        tmem_ld_x16(tdP + lane_base + (uint32_t)c0, rp);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float pv[2], dv2[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float pr = fast_exp2(fmaf(__uint_as_float(rs[2 * e + h]), LOG2E, nlL));     // nlL = -inf for rows past Nq
            if (kmask && j * AB_T + c0 + 2 * e + h >= p.Nk) pr = 0.f;
            pv[h] = pr;
            dv2[h] = pr * (__uint_as_float(rp[2 * e + h]) - dl);
          }
          __nv_bfloat162 hp = __floats2bfloat162_rn(pv[0], pv[1]), hd = __floats2bfloat162_rn(dv2[0], dv2[1]);
          pk[e] = *reinterpret_cast<uint32_t*>(&hp);
          dk[e] = *reinterpret_cast<uint32_t*>(&hd);
        }
        store_row16(Ps, row, c0, pk);
        store_row16(dSs, row, c0, dk);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(&s_empty)); mbar_arrive(smem_u32(&pds_full)); }
    }
    mbar_wait(smem_u32(&acc_full), 0);
    tc_fence_after();
    if (p.dq_part != nullptr) drain_dq((p.nqt - 1) * AB_T + row);      // dQ partial of the last query tile
    const int key = j * AB_T + row;          // accumulator rows are keys
    for (int c0 = part * 16; c0 < p.dv; c0 += NPART * 16) {
      uint32_t r[16];
      tmem_ld_x16(tdV + lane_base + (uint32_t)c0, r);
      if (key < p.Nk) store_acc_bf16(p.dv_out + ((long)b * p.Nk + key) * p.dv + c0, r);
    }
    for (int c0 = part * 16; c0 < p.dq; c0 += NPART * 16) {
      uint32_t r[16];
      tmem_ld_x16(tdK + lane_base + (uint32_t)c0, r);
      if (key < p.Nk) store_acc_bf16(p.dk_out + ((long)b * p.Nk + key) * p.dq + c0, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(320, 1) attn_bwd_dq_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qd_full, kv_full[2], kv_empty[2], s_full, s_empty, ds_full, ds_empty, acc_full;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t vbytes = (uint32_t)p.nvblk * 16384u;
  const uint32_t Qs = base, dOs = Qs + 16384u, Ks = dOs + vbytes, Vs = Ks + 2u * 16384u, dSs = Vs + 2u * vbytes;
  const int b = blockIdx.y, i = blockIdx.x;

  if (warp == 0 && elect_one_sync()) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); tma_prefetch_desc(&p.domap);
    mbar_init(smem_u32(&qd_full), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&kv_full[s]), 1); mbar_init(smem_u32(&kv_empty[s]), 1); }
    mbar_init(smem_u32(&s_full), 1); mbar_init(smem_u32(&s_empty), 8);
    mbar_init(smem_u32(&ds_full), 8); mbar_init(smem_u32(&ds_empty), 1); mbar_init(smem_u32(&acc_full), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t tS = tmem, tdP = tmem + 128u, tdQ = tmem + 256u;

  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_expect_tx(smem_u32(&qd_full), 16384u + vbytes);
      tma_load_3d(Qs, &p.qmap, smem_u32(&qd_full), 0, i * AB_T, b);
      for (int c = 0; c < p.nvblk; ++c) tma_load_3d(dOs + (uint32_t)c * 16384u, &p.domap, smem_u32(&qd_full), c * 64, i * AB_T, b);
      for (int j = 0; j < p.nkb; ++j) {
        const int s = j & 1;
        mbar_wait(smem_u32(&kv_empty[s]), (uint32_t)(((j >> 1) & 1) ^ 1));
        mbar_expect_tx(smem_u32(&kv_full[s]), 16384u + vbytes);
        tma_load_3d(Ks + (uint32_t)s * 16384u, &p.kmap, smem_u32(&kv_full[s]), 0, j * AB_T, b);
        for (int c = 0; c < p.nvblk; ++c)
          tma_load_3d(Vs + (uint32_t)s * vbytes + (uint32_t)c * 16384u, &p.vmap, smem_u32(&kv_full[s]), c * 64, j * AB_T, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc_dq = umma_idesc_bf16(128, p.dq, 0, 1);
      mbar_wait(smem_u32(&qd_full), 0);
      for (int j = 0; j <= p.nkb; ++j) {
        if (j < p.nkb) {
          const int s = j & 1;
          mbar_wait(smem_u32(&kv_full[s]), (uint32_t)((j >> 1) & 1));
          mbar_wait(smem_u32(&s_empty), (uint32_t)((j & 1) ^ 1));
          tc_fence_after();
          issue_s_dp(tS, tdP, Qs, Ks + (uint32_t)s * 16384u, dOs, Vs + (uint32_t)s * vbytes, p.dq, p.dv);
          umma_commit(smem_u32(&s_full));
        }
        const int pj = j - 1;
        if (pj >= 0) {
          const int s = pj & 1;
          mbar_wait(smem_u32(&ds_full), (uint32_t)(pj & 1));
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < AB_T / 16; ++k) {      // contract over the 128 keys of the block
            const uint64_t adS = umma_desc_kmajor_sw128(dSs + (uint32_t)(k >> 2) * 16384u + (uint32_t)(k & 3) * 32u);
            const uint64_t bK = umma_desc_mnmajor_sw128(Ks + (uint32_t)s * 16384u + (uint32_t)k * 2048u, 16384u, 1024u);
            umma_bf16(tdQ, adS, bK, idesc_dq, (pj | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&ds_empty));
          umma_commit(smem_u32(&kv_empty[s]));
        }
      }
      umma_commit(smem_u32(&acc_full));
    }
  } else {
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const int qi = i * AB_T + row;
    const bool qv = qi < p.Nq;
    const float nlL = qv ? -p.lse[(long)b * p.Nq + qi] * LOG2E : -INFINITY;
    const float dl = qv ? p.delta[(long)b * p.Nq + qi] : 0.f;
    for (int j = 0; j < p.nkb; ++j) {
      const bool kmask = (j + 1) * AB_T > p.Nk;
      mbar_wait(smem_u32(&s_full), (uint32_t)(j & 1));
      tc_fence_after();
      mbar_wait(smem_u32(&ds_empty), (uint32_t)((j & 1) ^ 1));
      for (int cc = 0; cc < 64; cc += 16) {
        const int c0 = half * 64 + cc;
        uint32_t rs[16], rp[16], dk[8];
        tmem_ld_x16(tS + lane_base + (uint32_t)c0, rs);
        tmem_ld_x16(tdP + lane_base + (uint32_t)c0, rp);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float d2[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float pr = fast_exp2(fmaf(__uint_as_float(rs[2 * e + h]), LOG2E, nlL));     // nlL = -inf for rows past Nq
            if (kmask && j * AB_T + c0 + 2 * e + h >= p.Nk) pr = 0.f;
            d2[h] = pr * (__uint_as_float(rp[2 * e + h]) - dl);
          }
          __nv_bfloat162 hd = __floats2bfloat162_rn(d2[0], d2[1]);
          dk[e] = *reinterpret_cast<uint32_t*>(&hd);
        }
        store_row16(dSs, row, c0, dk);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(&s_empty)); mbar_arrive(smem_u32(&ds_full)); }
    }
    mbar_wait(smem_u32(&acc_full), 0);
    tc_fence_after();
    const int qsplit = ((p.dq / 16 + 1) / 2) * 16;
    for (int c0 = half ? qsplit : 0; c0 < (half ? p.dq : qsplit); c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(tdQ + lane_base + (uint32_t)c0, r);
      if (qv) store_acc_bf16(p.dq_out + ((long)b * p.Nq + qi) * p.dq + c0, r);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

// dQ = sum over key blocks of the fp32 partials (fixed order) -> bf16; 8 values per thread
__global__ void __launch_bounds__(256) attn_dq_reduce_kernel(const float* __restrict__ part, long stride, int nkb, bf16* __restrict__ dq, long n8) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    for (int z = 0; z < nkb; ++z) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(part + (long)z * stride + i * 8));
      const float4 c = __ldg(reinterpret_cast<const float4*>(part + (long)z * stride + i * 8 + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += c.x; v[5] += c.y; v[6] += c.z; v[7] += c.w;
    }
    Vec<bf16>::st(dq, i * 8, v);
  }
}

// delta[b,i] = sum_c dO[b,i,c] * O[b,i,c]
__global__ void attn_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, float* __restrict__ delta, long rows, int dv) {
  const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = 0; c < dv; c += 4) {
    f4 a = bg_ld4(o, r * dv + c), g = bg_ld4(d_o, r * dv + c);
    s += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
  }
  delta[r] = s;
}

extern "C" size_t bg_attn2_bwd_workspace_bytes(int B, int Nq, int Nk, int dq) {
  return (size_t)bg_cdiv(Nk, AB_T) * B * Nq * dq * sizeof(float) + 256;
}

extern "C" int bg_attn2_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse, float* delta,
                            void* dq_out, void* dk_out, void* dv_out, int B, int Nq, int Nk, int dq, int dv, void* ws, size_t ws_bytes,
                            void* stream) {
  BG_REQUIRE(bg_has_tensor_core_path(), BG_ERR_UNSUPPORTED, "attn2: tensor-core path unavailable on this device");
  BG_REQUIRE(dq % 16 == 0 && dq >= 16 && dq <= 64, BG_ERR_UNSUPPORTED, "attn2: dq=%d must be 16..64, multiple of 16", dq);
  BG_REQUIRE(dv % 16 == 0 && dv >= 16 && dv <= 128, BG_ERR_UNSUPPORTED, "attn2 backward: dv=%d must be 16..128, multiple of 16", dv);
  BG_REQUIRE(B > 0 && Nq > 0 && Nk > 0 && B <= 65535, BG_ERR_SHAPE, "attn2: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const long rows = (long)B * Nq;
  attn_delta_kernel<<<(unsigned)bg_cdiv(rows, 256), 256, 0, st>>>((const bf16*)o, (const bf16*)d_o, delta, rows, dv);
  BG_LAUNCH_CHECK();
  static AttnBwdParams P;
  memset(&P, 0, sizeof(P));
  int rc;
  if ((rc = bg_attn_make_map(&P.qmap, q, B, Nq, dq, AB_T))) return rc;
  if ((rc = bg_attn_make_map(&P.kmap, k, B, Nk, dq, AB_T))) return rc;
  if ((rc = bg_attn_make_map(&P.vmap, v, B, Nk, dv, AB_T))) return rc;
  if ((rc = bg_attn_make_map(&P.domap, d_o, B, Nq, dv, AB_T))) return rc;
  P.B = B; P.Nq = Nq; P.Nk = Nk; P.dq = dq; P.dv = dv;
  P.nqt = (int)bg_cdiv(Nq, AB_T); P.nkb = (int)bg_cdiv(Nk, AB_T); P.nvblk = (dv + 63) / 64;
  P.lse = lse; P.delta = delta; P.dq_out = (bf16*)dq_out; P.dk_out = (bf16*)dk_out; P.dv_out = (bf16*)dv_out;
  // with workspace: the dkdv kernel also writes this key block's dQ contribution and a reduce pass replaces the dq kernel
  const long part_stride = (long)B * Nq * dq;
  const bool fuse_dq = ws != nullptr && ws_bytes >= (size_t)P.nkb * part_stride * sizeof(float) && (((uintptr_t)ws) & 15) == 0 &&
                       (part_stride % 8) == 0;
  P.dq_part = fuse_dq ? (float*)ws : nullptr;
  static bool attr_set = false;
  if (!attr_set) {
    BG_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    BG_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  const size_t vb = (size_t)P.nvblk * 16384;
  const size_t smem_kv = 16384 + vb + 2 * 16384 + 2 * vb + 32768 + 32768 + 1024;
  const size_t smem_q = 16384 + vb + 2 * 16384 + 2 * vb + 32768 + 1024;
  attn_bwd_dkdv_kernel<<<dim3((unsigned)P.nkb, (unsigned)B), 64 + 32 * AB_SW, smem_kv, st>>>(P);
  BG_LAUNCH_CHECK();
  if (fuse_dq) {
    const long n8 = part_stride / 8;
    const int blocks = (int)min(bg_cdiv(n8, 256), (long)bg_num_sms() * 16);
    attn_dq_reduce_kernel<<<blocks, 256, 0, st>>>(P.dq_part, part_stride, P.nkb, (bf16*)dq_out, n8);
  } else {
    attn_bwd_dq_kernel<<<dim3((unsigned)P.nqt, (unsigned)B), 320, smem_q, st>>>(P);
  }
  BG_LAUNCH_CHECK();
  return BG_OK;
}
