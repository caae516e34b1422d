// tcgen05 / TMEM / TMA implicit-GEMM kernels (bf16 operands, fp32 accumulate) for the dense
// contractions of the BigGAN step: conv, conv2d_transpose and their dgrad (ops.py:49-139) share ONE
// forward-type kernel driven by the phase/tap geometry of bg_common.cuh; wgrad has its own kernel.
//
// forward-type tile: D[128 output pixels, BN channels] (fp32, TMEM) = sum over taps, 64-channel
// k-blocks of  A[128 pixels, 64 ch] (TMA box of the NHWC activation, OOB zero-fill = zero padding,
// SWIZZLE_128B)  x  B[BN, 64 ch] (TMA box of the K-major weights [tap][n][k]).
// Warp roles (320 threads, persistent CTA): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issue (one ELECTED lane each:
// `elect_one_sync()`, never `lane == 0`, see tc_ptx.cuh), warps 2-9 = epilogue (tcgen05.ld -> bias / residual / act -> bf16).
// Variants: halo kernel (taps of one dy share an activation box), split-K work items, fold mode (reflect-conv dgrad).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include "bg_common.cuh"

#include "tc_ptx.cuh"

// ------------------------------------------------------------------------------------------------
// forward-type kernel
// ------------------------------------------------------------------------------------------------
#define TC_MAX_TAPS 16
struct TcTap { int map, dy, dx, widx; };
#define TC_MAX_GROUPS 8
#define TC_HALO_ABYTES 17408   // activation box region of a halo stage: up to 136 pixel rows of 128 B, 1024-aligned
struct TcGroup { int dy, dx0, ntap; int widx[3]; int dxo[3]; };
struct TcFwdParams {
  CUtensorMap amap[4];
  CUtensorMap bmap;
  int B, Ho, Wo, Co;
  int nphase, ostep;
  int bw, bh, bb;            // pixel box (bw*bh*bb == 128)
  int tw, th, tb;            // tiles per phase
  int nkb, BN, act, stages;
  // fold mode (dgrad of a reflect-padded conv): the kernel iterates the PADDED domain; interior pixels go straight to `out`
  // (unpadded tensor, + residual), the ring of pad pixels to `ring` (padded-domain buffer), folded back by a band kernel
  int fold, f_pt, f_pl, f_Hi, f_Wi;
  bf16* ring;
  const bf16* mask;          // fold mode: ReLU mask (the conv's input, same layout as `out`): interior gradients are zeroed where mask <= 0
  int ksplit;                // split-K: each tile's (tap, k-block) stage sequence is cut into ksplit work items whose fp32
  float* part;               //   partial accumulators go to part[split * part_stride + out offset] (summed by tc_splitk_reduce_kernel)
  long part_stride;
  int dry;                   // tuning: skip the MMAs (measures the bare TMA ingest rate of the tile stream)
  int ktail;                 // 16-channel MMA steps holding data in the last 64-channel k-block (the rest is OOB zero fill)
  int oy0[4], ox0[4], Hp[4], Wp[4], ntaps[4];
  TcTap taps[4][TC_MAX_TAPS];
  const float* bias;
  const bf16* residual;
  bf16* out;
  // halo mode (tc_fwd_halo_kernel): taps that differ only in dx share ONE activation box widened by their dx span
  int ngroups, halo_w;       // tap groups; box width in pixels (128 + span)
  TcGroup groups[TC_MAX_GROUPS];
  // TMA-store epilogue (layers whose main loop is short): the bf16 tile is staged in shared memory as 16-channel slabs
  // [128 pixels][32 B] and written with cp.async.bulk.tensor -- full 32-byte sectors in a few bulk requests instead of one
  // 16-byte store per thread and slab, which saturated the SM's L1 -> crossbar request port (ncu: l1tex__m_l1tex2xbar_req
  // 65 % busy at 33 % tensor-pipe activity on the 96-channel 3x3 layers)
  // resident-weights mode (plain kernel, Co == BN, all weight tiles fit): the (tap, k-block) weight tiles are loaded ONCE per
  // persistent CTA into their own shared-memory region; ring stages then carry activations only.  For the layers with few
  // output channels / small K (1x1 projections, the RGB layers) the weights were 40-75 % of every tile's ingest.
  // paired-tile mode (plain kernel, BN <= 128, streamed weights): a work item is TWO consecutive 128-pixel tiles sharing every weight
  // stage (two TMEM accumulators): the weight tiles enter the SM once per 256 pixels instead of once per 128
  int mt;                    // m-tiles per work item: 1 or 2
  int resident_b;
  uint32_t bres_off;         // resident weight region, bytes from the 1024-aligned dynamic shared memory base
  int nwidx;                 // weight tiles per k-block in the region (max widx + 1)
  int tma_store;             // 0 = direct stores
  uint32_t ostage_off;       // staging area, bytes from the 1024-aligned dynamic shared memory base
  uint32_t ostage_half;      // bytes per epilogue half and set
  int ostage_sets;           // 1, or 2 when the ring can spare the memory (tile i's bulk stores overlap the staging of tile i + 1)
  CUtensorMap omap[4];       // per phase: the output sampled with step `ostep` from (oy0, ox0); box (16, bw, bh, bb), no swizzle
  int o_di[4], o_dj[4];      // fold mode: omap covers the UNPADDED tensor from the phase's first interior row / column; tile row i0 is
                             //   its row i0 - o_di (tiles that are interior pixels only take the bulk-store path, the others store directly)
  int phase_minor;           // 4-phase (strided) launches: the phase is the fastest index after the n-tile, so the four interleaved
                             //   sub-pixel phases of one output region run at the same time on neighbouring SMs and their partial-line
                             //   writes (and residual / mask reads) meet in L2 instead of visiting DRAM four times, far apart
  int nwork;                 // work items x ksplit (loop bound of the persistent roles, kept in the parameter block)
  int aux_batch;             // BN <= 96: all residual / mask vectors of a half-tile requested up front (A/B knob BG_TC_AUX_BATCH)
  FastDiv fd_S, fd_ntn, fd_tpp, fd_tw, fd_th;   // per-tile index decode without runtime divisions (ksplit, n-tiles, items per phase, tw, th)
};

// tile index within a phase -> first pixel of its box
__device__ __forceinline__ void tc_tile_origin(const TcFwdParams& p, int t, int& j0, int& i0, int& b0) {
  int q = fast_div(t, p.fd_tw); j0 = (t - q * p.tw) * p.bw; t = q;
  q = fast_div(t, p.fd_th); i0 = (t - q * p.th) * p.bh; b0 = q * p.bb;
}
// work item -> (split, n-tile, phase, item within the phase)
__device__ __forceinline__ void tc_item_decode(const TcFwdParams& p, int wi, int ntn, int tpp_items, int& split, int& nt, int& phase, int& ti) {
  const int item = fast_div(wi, p.fd_S); split = wi - item * p.ksplit;
  const int q = fast_div(item, p.fd_ntn); nt = item - q * ntn;
  if (p.phase_minor) { ti = q >> 2; phase = (q + ti) & 3; }           // nphase == 4; rotated so that a persistent CTA (stride = grid size) meets every phase
  else { phase = fast_div(q, p.fd_tpp); ti = q - phase * tpp_items; }
}

// Epilogue warps (8): drain the double-buffered TMEM accumulator of every tile this CTA owns:
// tcgen05.ld -> bias / residual / activation -> bf16 -> global.  Shared by the plain and the halo kernel.
__device__ __forceinline__ void tc_fwd_epilogue(const TcFwdParams& p, uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                int warp, int lane, int ntiles, int ntn, int tiles_pp, uint32_t smem_base) {
    const int ew = warp - 2;                      // 0..7
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = ew >> 2;                     // which half of the BN columns
    const int csplit = ((p.BN / 16 + 1) / 2) * 16;
    const int cbeg = half ? csplit : 0, cend = half ? p.BN : csplit;
    const int row = q * 32 + lane;
    const int w = row % p.bw, h = (row / p.bw) % p.bh, bl = row / (p.bw * p.bh);
    int lt = 0, nstored = 0;
    const int S = p.ksplit;
    const int MT = p.mt;
    // (items per phase and n-tiles are read from the parameter block where needed: p.fd_tpp.d, p.fd_ntn.d -- no live registers)
    for (int wi = blockIdx.x; wi < p.nwork; wi += gridDim.x, ++lt) {
      int split, nt, phase, ti;
      tc_item_decode(p, wi, p.fd_ntn.d, p.fd_tpp.d, split, nt, phase, ti);
     for (int m = 0; m < MT; ++m) {
      int j0, i0, b0;                                        // a tile index past the phase decodes to b0 >= B: nothing valid, boxes clipped
      tc_tile_origin(p, ti * MT + m, j0, i0, b0);
      const int n0 = nt * p.BN;
      const int pi = i0 + h, pj = j0 + w, b = b0 + bl;
      const bool valid = (b < p.B) && (pi < p.Hp[phase]) && (pj < p.Wp[phase]);
      const int oy = p.oy0[phase] + pi * p.ostep, ox = p.ox0[phase] + pj * p.ostep;
      long obase = (((long)b * p.Ho + oy) * p.Wo + ox) * p.Co + n0;
      bf16* outb = p.out;
      bool res_ok = true;
      if (p.fold) {
        const int iy = oy - p.f_pt, ix = ox - p.f_pl;
        if (iy >= 0 && iy < p.f_Hi && ix >= 0 && ix < p.f_Wi) obase = (((long)b * p.f_Hi + iy) * p.f_Wi + ix) * p.Co + n0;
        else { outb = p.ring; res_ok = false; }
      }
      const int buf = lt & 1;
      const uint32_t tcol = (uint32_t)(buf * MT * p.BN + m * p.BN);          // this tile's accumulator columns
      // residual (fused accumulate / skip connection) is read in groups of 2 x 16 channels, one group ahead of its use; the
      // first group is in flight while this warp still waits for the accumulator, so the global-load latency is off the
      // per-tile critical path
      if (S > 1) {                                   // split-K work item: raw fp32 partials, finished by the reduce kernel
        mbar_wait(smem_u32(&tfull_bar[buf]), (uint32_t)((lt >> 1) & 1));
        tc_fence_after();
        float* pp = p.part + (long)split * p.part_stride + obase;
        for (int c0 = cbeg; c0 < cend; c0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)c0, r);
          if (valid) {
            float4* o = reinterpret_cast<float4*>(pp + c0);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]),
                                 __uint_as_float(r[4 * e + 3]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
        continue;
      }
      const bool use_res = (p.residual != nullptr) && valid && res_ok;
      const bool use_mask = (p.mask != nullptr) && valid && res_ok;      // fold mode, interior pixel: d relu
      bool tstore = p.tma_store != 0;
      if (tstore && p.fold) {                    // fold mode: bulk stores only for tiles made of interior pixels (tile-uniform)
        const int y_lo = p.oy0[phase] + i0 * p.ostep - p.f_pt, y_hi = y_lo + (p.bh - 1) * p.ostep;
        const int x_lo = p.ox0[phase] + j0 * p.ostep - p.f_pl, x_hi = x_lo + (p.bw - 1) * p.ostep;
        tstore = y_lo >= 0 && y_hi < p.f_Hi && x_lo >= 0 && x_hi < p.f_Wi;
      }
      // staging slabs are double-buffered per half: tile lt writes set lt & 1 while the bulk stores of tile lt - 1 still read
      // the other set; the stores of tile lt - 2 (same set) must have finished reading
      const uint32_t hset = smem_base + p.ostage_off + (uint32_t)(p.ostage_sets * half + (p.ostage_sets == 2 ? (nstored & 1) : 0)) * p.ostage_half;
      if (tstore) ++nstored;                     // sets alternate per BULK-STORED tile (direct tiles of the fold mode in between commit nothing)
      const uint32_t hb = hset + (uint32_t)row * 32u;
      // one 16-channel slab: accumulator -> bias -> d relu mask -> residual -> activation -> bf16 -> staging slab / global
      auto do_slab = [&](int c0, const uint4& r0, const uint4& r1, const uint4& m0, const uint4& m1) {
        uint32_t r[16];
        tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + tcol + (uint32_t)c0, r);
        if (valid || tstore) {
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c0);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float4 bv = __ldg(bp + e);
              v[4 * e] += bv.x; v[4 * e + 1] += bv.y; v[4 * e + 2] += bv.z; v[4 * e + 3] += bv.w;
            }
          }
          if (use_mask) {                                    // before the residual joins
            const __nv_bfloat162* g0 = reinterpret_cast<const __nv_bfloat162*>(&m0);
            const __nv_bfloat162* g1 = reinterpret_cast<const __nv_bfloat162*>(&m1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (!(__low2float(g0[e]) > 0.f)) v[2 * e] = 0.f;
              if (!(__high2float(g0[e]) > 0.f)) v[2 * e + 1] = 0.f;
              if (!(__low2float(g1[e]) > 0.f)) v[8 + 2 * e] = 0.f;
              if (!(__high2float(g1[e]) > 0.f)) v[8 + 2 * e + 1] = 0.f;
            }
          }
          if (use_res) {
            const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
            const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[2 * e] += __low2float(h0[e]); v[2 * e + 1] += __high2float(h0[e]);
              v[8 + 2 * e] += __low2float(h1[e]); v[8 + 2 * e + 1] += __high2float(h1[e]);
            }
          }
          if (p.act == BG_ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          } else if (p.act == BG_ACT_TANH) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fast_tanh(v[e]);
          }
          uint4 o0, o1;
          __nv_bfloat162* q0 = reinterpret_cast<__nv_bfloat162*>(&o0);
          __nv_bfloat162* q1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            q0[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
            q1[e] = __floats2bfloat162_rn(v[8 + 2 * e], v[8 + 2 * e + 1]);
          }
          if (tstore) {
            const uint32_t a = hb + (uint32_t)((c0 - cbeg) >> 4) * 4096u;
            st_shared_v4(a, o0);
            st_shared_v4(a + 16u, o1);
          } else {
            uint4* op = reinterpret_cast<uint4*>(outb + obase + c0);
            op[0] = o0;
            op[1] = o1;
          }
        }
      };
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      if (p.aux_batch && cend - cbeg <= 48) {
        // BN <= 96: EVERY residual / mask vector of this half-tile is requested up front (3 slabs x 2 x 16 B each), before the
        // staging-set and accumulator waits -- one global-memory latency per tile instead of one per slab (ncu on the 96-channel
        // fold-mode dgrads: 62 % of the epilogue warps' samples were long-scoreboard stalls on the mask loads, issued one slab at a time)
        uint4 ra[3][2], ma[3][2];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          ra[i][0] = ra[i][1] = ma[i][0] = ma[i][1] = z4;
          const int c0 = cbeg + 16 * i;
          if (c0 < cend) {
            if (use_res) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + c0);
              ra[i][0] = __ldg(rp); ra[i][1] = __ldg(rp + 1);
            }
            if (use_mask) {
              const uint4* mp = reinterpret_cast<const uint4*>(p.mask + obase + c0);
              ma[i][0] = __ldg(mp); ma[i][1] = __ldg(mp + 1);
            }
          }
        }
        if (tstore) {
          if ((ew & 3) == 0 && elect_one_sync()) {                          // same elected lane as the one issuing the stores
            if (p.ostage_sets == 2) tma_store_wait_read1(); else tma_store_wait_read0();
          }
          named_bar_sync(1u + (uint32_t)half, 128u);
        }
        if (m == 0) {
          mbar_wait(smem_u32(&tfull_bar[buf]), (uint32_t)((lt >> 1) & 1));
          tc_fence_after();
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int c0 = cbeg + 16 * i;
          if (c0 < cend) do_slab(c0, ra[i][0], ra[i][1], ma[i][0], ma[i][1]);
        }
      } else {
        // BN > 96 (long main loops): residual in groups of 2 x 16 channels, one group ahead of its use; the first group is in
        // flight while this warp still waits for the accumulator
        uint4 rb[2][2][2];
        auto load_group = [&](int g, int cg) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (use_res && cg + j * 16 < cend) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + cg + j * 16);
              rb[g][j][0] = __ldg(rp);
              rb[g][j][1] = __ldg(rp + 1);
            }
        };
        load_group(0, cbeg);
        if (tstore) {
          if ((ew & 3) == 0 && elect_one_sync()) {
            if (p.ostage_sets == 2) tma_store_wait_read1(); else tma_store_wait_read0();
          }
          named_bar_sync(1u + (uint32_t)half, 128u);
        }
        if (m == 0) {
          mbar_wait(smem_u32(&tfull_bar[buf]), (uint32_t)((lt >> 1) & 1));
          tc_fence_after();
        }
        int gsel = 0;
        for (int cg = cbeg; cg < cend; cg += 32, gsel ^= 1) {
          if (gsel == 0) load_group(1, cg + 32); else load_group(0, cg + 32);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c0 = cg + j * 16;
            if (c0 >= cend) break;
            uint4 m0 = z4, m1 = z4;
            if (use_mask) {
              const uint4* mp = reinterpret_cast<const uint4*>(p.mask + obase + c0);
              m0 = __ldg(mp); m1 = __ldg(mp + 1);
            }
            do_slab(c0, gsel == 0 ? rb[0][j][0] : rb[1][j][0], gsel == 0 ? rb[0][j][1] : rb[1][j][1], m0, m1);
          }
        }
      }
      if (m == MT - 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[buf]));
      }
      if (tstore) {
        fence_proxy_async_smem();
        named_bar_sync(1u + (uint32_t)half, 128u);
        if ((ew & 3) == 0 && elect_one_sync()) {
          uint32_t src = hset;
          for (int c0 = cbeg; c0 < cend; c0 += 16, src += 4096u)
            tma_store_4d(&p.omap[phase], src, n0 + c0, j0 - p.o_dj[phase], i0 - p.o_di[phase], b0);
          tma_store_commit();
        }
      }
     }
    }
    if (p.tma_store && (ew & 3) == 0 && elect_one_sync()) tma_store_wait_all();     // shared memory must outlive the last bulk store
}

// Persistent: one CTA per SM walks the tile list (n-tile fastest, so concurrently running CTAs share the
// activation tile in L2).  The TMA ring never drains between tiles, and the accumulator is double-buffered
// in TMEM so the 8 epilogue warps drain tile i while the MMA warp already accumulates tile i+1.
template <int TMEM_COLS>
__global__ void __launch_bounds__(320, 1) tc_fwd_kernel(const __grid_constant__ TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wtile = (uint32_t)p.BN * 128u;
  const int MT = p.mt;                                   // m-tiles per work item (1 or 2)
  const uint32_t abytes = (uint32_t)MT * 16384u;         // activation tiles of a stage
  const uint32_t stage_bytes = p.resident_b ? abytes : abytes + wtile;
  const int tiles_pp = p.tw * p.th * p.tb;
  const int tpp_items = (tiles_pp + MT - 1) / MT;
  const int ntn = p.Co / p.BN;
  const int ntiles = tpp_items * p.nphase * ntn;          // work items

  if (warp == 0 && elect_one_sync()) {
    for (int m = 0; m < 4; ++m) tma_prefetch_desc(&p.amap[m]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), 8); }
    mbar_init(smem_u32(&bres_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // Both issue loops are single dependent instruction streams: stage index / parity / shared-memory addresses advance
  // incrementally (no runtime divisions), descriptor words by plain adds.
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, par = 1, sa = smem_base;
      const int S = p.ksplit;
      if (p.resident_b) {                            // every weight tile of the layer, once per CTA
        const uint32_t bb = smem_u32(&bres_bar);
        mbar_expect_tx(bb, (uint32_t)(p.nwidx * p.nkb) * wtile);
        uint32_t dst = smem_base + p.bres_off;
        for (int w = 0; w < p.nwidx; ++w)
          for (int kb = 0; kb < p.nkb; ++kb, dst += wtile) tma_load_3d(dst, &p.bmap, bb, kb * 64, 0, w);
      }
      for (int wi = blockIdx.x; wi < ntiles * S; wi += gridDim.x) {
        int split, nt, phase, ti;
        tc_item_decode(p, wi, ntn, tpp_items, split, nt, phase, ti);
        int j0, i0, b0, j1 = 0, i1 = 0, b1 = 0;
        tc_tile_origin(p, ti * MT, j0, i0, b0);
        if (MT == 2) tc_tile_origin(p, ti * MT + 1, j1, i1, b1);   // second tile of a pair (past the phase: b >= B, zero fill)
        const int n0 = nt * p.BN;
        const int total = p.ntaps[phase] * p.nkb;
        int lo = 0, hi = total, tp = 0, kb = 0;                              // this work item's (tap, k-block) stages
        if (S > 1) { lo = total * split / S; hi = total * (split + 1) / S; tp = lo / p.nkb; kb = lo - tp * p.nkb; }
        TcTap tap = p.taps[phase][tp];
        for (int li = lo; li < hi; ++li) {
          mbar_wait(empty0 + s * 8u, par);
          const uint32_t fb = full0 + s * 8u;
          mbar_expect_tx(fb, stage_bytes);
          tma_load_4d(sa, &p.amap[tap.map], fb, kb * 64, j0 + tap.dx, i0 + tap.dy, b0);
          if (MT == 2) tma_load_4d(sa + 16384u, &p.amap[tap.map], fb, kb * 64, j1 + tap.dx, i1 + tap.dy, b1);
          if (!p.resident_b) tma_load_3d(sa + abytes, &p.bmap, fb, kb * 64, n0, tap.widx);
          if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; sa = smem_base; } else sa += stage_bytes;
          if (++kb == p.nkb) { kb = 0; ++tp; if (li + 1 < hi) tap = p.taps[phase][tp]; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN, 0, 0);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t a_lo0 = umma_desc_lo(smem_base, 16), stage_q = stage_bytes >> 4;
      const uint32_t dry = (uint32_t)p.dry;
      uint32_t s = 0, par = 0, a_lo = a_lo0;
      int lt = 0;
      const int S = p.ksplit;
      const uint32_t res = (uint32_t)p.resident_b;
      const uint32_t bres_lo = umma_desc_lo(smem_base + p.bres_off, 16), wtile_q = wtile >> 4;
      if (res) { mbar_wait(smem_u32(&bres_bar), 0); tc_fence_after(); }
      for (int wi = blockIdx.x; wi < ntiles * S; wi += gridDim.x, ++lt) {
        int split, nt_, phase, ti_;
        tc_item_decode(p, wi, ntn, tpp_items, split, nt_, phase, ti_);
        const int total = p.ntaps[phase] * p.nkb;
        int lo = 0, hi = total, kb = 0, tp = 0;
        if (S > 1) { lo = total * split / S; hi = total * (split + 1) / S; tp = lo / p.nkb; kb = lo - tp * p.nkb; }
        const int buf = lt & 1;
        mbar_wait(smem_u32(&tempty_bar[buf]), (uint32_t)(((lt >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * MT * p.BN);
        uint32_t acc = 0u;
        uint32_t bw_lo = res ? bres_lo + (uint32_t)(p.taps[phase][tp].widx * p.nkb + kb) * wtile_q : 0u;
        for (int li = lo; li < hi; ++li) {
          mbar_wait(full0 + s * 8u, par);
          tc_fence_after();
          const int nk = (kb == p.nkb - 1) ? p.ktail : 4;
          if (!dry) {
            const uint32_t b_lo = res ? bw_lo : a_lo + (abytes >> 4);     // weight tile: resident region, or behind the activation tile(s)
            umma_bf16_w(tacc, a_lo, UMMA_DESC_HI_KMAJOR_SW128, b_lo, UMMA_DESC_HI_KMAJOR_SW128, idesc, acc);
            if (nk > 1) umma_bf16_w(tacc, a_lo + 2u, UMMA_DESC_HI_KMAJOR_SW128, b_lo + 2u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
            if (nk > 2) umma_bf16_w(tacc, a_lo + 4u, UMMA_DESC_HI_KMAJOR_SW128, b_lo + 4u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
            if (nk > 3) umma_bf16_w(tacc, a_lo + 6u, UMMA_DESC_HI_KMAJOR_SW128, b_lo + 6u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
            if (MT == 2) {                                     // second tile of the pair: its own activation tile and accumulator, same weights
              const uint32_t a1 = a_lo + 1024u, t1 = tacc + (uint32_t)p.BN;
              umma_bf16_w(t1, a1, UMMA_DESC_HI_KMAJOR_SW128, b_lo, UMMA_DESC_HI_KMAJOR_SW128, idesc, acc);
              if (nk > 1) umma_bf16_w(t1, a1 + 2u, UMMA_DESC_HI_KMAJOR_SW128, b_lo + 2u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
              if (nk > 2) umma_bf16_w(t1, a1 + 4u, UMMA_DESC_HI_KMAJOR_SW128, b_lo + 4u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
              if (nk > 3) umma_bf16_w(t1, a1 + 6u, UMMA_DESC_HI_KMAJOR_SW128, b_lo + 6u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
            }
            acc = 1u;
          }
          umma_commit(empty0 + s * 8u);
          if (++kb == p.nkb) {
            kb = 0; ++tp;
            if (res && li + 1 < hi) bw_lo = bres_lo + (uint32_t)(p.taps[phase][tp].widx * p.nkb) * wtile_q;
          } else bw_lo += wtile_q;
          if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; a_lo = a_lo0; } else a_lo += stage_q;
        }
        umma_commit(smem_u32(&tfull_bar[buf]));
      }
    }
  } else {
    tc_fwd_epilogue(p, tmem_base, tfull_bar, tempty_bar, warp, lane, ntiles, ntn, tiles_pp, smem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// Halo variant for stride-1 layers whose tile is one 128-pixel image-row segment: the taps (dy, dx0..dx0+2) of a
// group read the SAME activation rows shifted by one pixel, i.e. by one 128-byte row of the K-major SWIZZLE_128B tile.
// One TMA box of 128+span pixels feeds all of them: tap j's A descriptor starts dxo[j] rows into the box (matrix base
// offset = dxo[j], the row phase of the swizzle pattern), so a 3x3 conv loads each activation row 3x instead of 9x.
// These layers are bound by the L2->SM ingest of the activation stream, not by the tensor pipe.
template <int TMEM_COLS>
__global__ void __launch_bounds__(320, 1) tc_fwd_halo_kernel(const __grid_constant__ TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t wbytes = (uint32_t)p.BN * 128u;                       // one weight tile
  const uint32_t stage_bytes = TC_HALO_ABYTES + 3u * wbytes;
  const uint32_t abox_bytes = (uint32_t)p.halo_w * 128u;
  const int tiles_pp = p.tw * p.th * p.tb;
  const int ntn = p.Co / p.BN;
  const int ntiles = tiles_pp * ntn;

  if (warp == 0 && elect_one_sync()) {
    tma_prefetch_desc(&p.amap[0]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tfull_bar[i]), 1); mbar_init(smem_u32(&tempty_bar[i]), 8); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, par = 1, sa = smem_base;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t = fast_div(tile, p.fd_ntn), nt = tile - t * ntn;
        int j0, i0, b0;
        tc_tile_origin(p, t, j0, i0, b0);
        const int n0 = nt * p.BN;
        for (int g = 0; g < p.ngroups; ++g) {
          const TcGroup& grp = p.groups[g];
          const int ntap = grp.ntap, cx = j0 + grp.dx0, cy = i0 + grp.dy;
          const int w0 = grp.widx[0], w1 = grp.widx[1], w2 = grp.widx[2];
          const uint32_t tx = abox_bytes + (uint32_t)ntap * wbytes;
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(empty0 + s * 8u, par);
            const uint32_t fb = full0 + s * 8u;
            mbar_expect_tx(fb, tx);
            tma_load_4d(sa, &p.amap[0], fb, kb * 64, cx, cy, b0);
            const uint32_t sb = sa + TC_HALO_ABYTES;
            tma_load_3d(sb, &p.bmap, fb, kb * 64, n0, w0);
            if (ntap > 1) tma_load_3d(sb + wbytes, &p.bmap, fb, kb * 64, n0, w1);
            if (ntap > 2) tma_load_3d(sb + 2u * wbytes, &p.bmap, fb, kb * 64, n0, w2);
            if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; sa = smem_base; } else sa += stage_bytes;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN, 0, 0);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t a_lo0 = umma_desc_lo(smem_base, 16), stage_q = stage_bytes >> 4, w_q = wbytes >> 4;
      const uint32_t dry = (uint32_t)p.dry;
      uint32_t s = 0, par = 0, a_lo = a_lo0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
        const int buf = lt & 1;
        mbar_wait(smem_u32(&tempty_bar[buf]), (uint32_t)(((lt >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * p.BN);
        uint32_t acc = 0u;
        for (int g = 0; g < p.ngroups; ++g) {
          const int ntap = p.groups[g].ntap;
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(full0 + s * 8u, par);
            tc_fence_after();
            const int nk = (kb == p.nkb - 1) ? p.ktail : 4;
            if (!dry) {
              // tap j reads the activation box j pixel rows (j * 128 B = 8 descriptor units) further in; the UMMA swizzle
              // is a function of the absolute shared-memory address, so no matrix-base-offset is needed (checked on B200)
              uint32_t aj = a_lo, bj = a_lo + (TC_HALO_ABYTES >> 4);
              for (int j = 0; j < ntap; ++j, aj += 8u, bj += w_q) {
                umma_bf16_w(tacc, aj, UMMA_DESC_HI_KMAJOR_SW128, bj, UMMA_DESC_HI_KMAJOR_SW128, idesc, acc);
                acc = 1u;
                if (nk > 1) umma_bf16_w(tacc, aj + 2u, UMMA_DESC_HI_KMAJOR_SW128, bj + 2u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
                if (nk > 2) umma_bf16_w(tacc, aj + 4u, UMMA_DESC_HI_KMAJOR_SW128, bj + 4u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
                if (nk > 3) umma_bf16_w(tacc, aj + 6u, UMMA_DESC_HI_KMAJOR_SW128, bj + 6u, UMMA_DESC_HI_KMAJOR_SW128, idesc, 1u);
              }
            }
            umma_commit(empty0 + s * 8u);
            if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; a_lo = a_lo0; } else a_lo += stage_q;
          }
        }
        umma_commit(smem_u32(&tfull_bar[buf]));
      }
    }
  } else {
    tc_fwd_epilogue(p, tmem_base, tfull_bar, tempty_bar, warp, lane, ntiles, ntn, tiles_pp, smem_base);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)f;
  }
  return fn;
}

static bool tc_device_ok() {
  static int ok = -1;
  if (ok < 0) {
    ok = 0;
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
        major == 10 && get_encode() != nullptr)
      ok = 1;
    const char* e = getenv("BG_DISABLE_TC");
    if (e && e[0] == '1') ok = 0;
  }
  return ok == 1;
}

extern "C" int bg_has_tensor_core_path(void) { return tc_device_ok() ? 1 : 0; }

// activation map [B,H,W,C] bf16 sampled with step s starting at parity (pa,pb); box (64, bw, bh, bb)
static int make_act_map(CUtensorMap* m, const void* base, int B, int H, int W, int C, int s, int pa, int pb, int bw, int bh, int bb,
                        int box_c = 64, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  const int Hs = (H - pa + s - 1) / s, Ws = (W - pb + s - 1) / s;
  if (Hs <= 0 || Ws <= 0) { memset(m, 0, sizeof(*m)); return BG_OK; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ws, (cuuint64_t)Hs, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)s * C * 2, (cuuint64_t)s * W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t es[4] = {1, 1, 1, 1};
  const char* ptr = (const char*)base + ((long)pa * W + pb) * C * 2;
  static int promo = -1;                          // tuning knob: L2 promotion of the activation maps (default 256 B)
  if (promo < 0) { const char* e = getenv("BG_TC_PROMO"); promo = e ? atoi(e) : 3; if (promo < 0 || promo > 3) promo = 3; }
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swz, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bg_set_error("cuTensorMapEncodeTiled(act) failed: %d (C=%d W=%d H=%d B=%d s=%d)", (int)r, C, Ws, Hs, B, s); return BG_ERR_CUDA; }
  return BG_OK;
}

// weights [taps][N][K] bf16, K contiguous; box (64, BN, 1)
static int make_w_map(CUtensorMap* m, const void* base, int taps, int N, int K, int BN) {
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)BN, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = get_encode()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bg_set_error("cuTensorMapEncodeTiled(w) failed: %d (K=%d N=%d taps=%d)", (int)r, K, N, taps); return BG_ERR_CUDA; }
  return BG_OK;
}

// pixel box (bw,bh,bb), product `npix`, maximising the fraction of useful rows
static void choose_box(int Wp, int Hp, int B, int npix, int* bw, int* bh, int* bb) {
  double best = -1.0;
  for (int w = 1; w <= npix; w *= 2)
    for (int h = 1; w * h <= npix; h *= 2) {
      const int b = npix / (w * h);
      if (w > 256 || h > 256 || b > 256) continue;
      const double cover = (double)bg_cdiv(Wp, w) * w * (double)bg_cdiv(Hp, h) * h * (double)bg_cdiv(B, b) * b;
      const double util = (double)Wp * Hp * B / cover + 1e-9 * w + 1e-12 * h;   // ties: wider boxes
      if (util > best) { best = util; *bw = w; *bh = h; *bb = b; }
    }
}

static int pick_bn(int Co) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (Co % bn == 0) return bn;
  return 0;
}

static inline int floordiv2(int a) { return (a >= 0) ? a / 2 : -((-a + 1) / 2); }

// out = act(sum_z part[z] + bias + residual) -> bf16, 8 channels per thread (Co % 16 == 0)
__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float* __restrict__ part, long stride, int nz, const float* __restrict__ bias, const bf16* residual,
                        bf16* out, long n8, int Co, int act) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    const long e0 = i * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    for (int z = 0; z < nz; ++z) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(part + (long)z * stride + e0));
      const float4 b = __ldg(reinterpret_cast<const float4*>(part + (long)z * stride + e0 + 4));
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (bias) {
      const int c = (int)(e0 % Co);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += bias[c + e];
    }
    if (residual) {
      float r[8];
      Vec<bf16>::ld(residual, e0, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += r[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = bg_apply_act(v[e], act);
    Vec<bf16>::st(out, e0, v);
  }
}

// Tiling + split-K plan of the forward-type kernel (shared by the launcher and the workspace query).
// Split-K: when the tile count leaves most SMs idle (small maps / small per-GPU batch) or its last wave is mostly empty,
// each tile's (tap, k-block) stage sequence is cut into S work items with fp32 partials in workspace.
struct TcFwdPlan { int ok, BN, bw, bh, bb, tw, th, tb, ntiles, ksplit; };
struct BgTcFold { int pt, pl, Hi, Wi; void* ring; const void* mask; };   // ring == nullptr: no fold, only the ReLU mask (same layout as `out`)
static TcFwdPlan tc_fwd_plan(const BgGeom& g) {
  TcFwdPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!tc_device_ok()) return pl;
  if (g.Ci % 8 != 0 || g.Ci < 16 || g.Co % 16 != 0 || g.nphase < 1) return pl;
  int min_taps = 1 << 30;
  for (int ph = 0; ph < g.nphase; ++ph) {
    if (g.ph[ph].Hp != g.ph[0].Hp || g.ph[ph].Wp != g.ph[0].Wp) return pl;
    if (g.ph[ph].ntaps < 1 || g.ph[ph].ntaps > TC_MAX_TAPS) return pl;
    min_taps = min(min_taps, g.ph[ph].ntaps);
  }
  pl.BN = pick_bn(g.Co);
  if (!pl.BN) return pl;
  pl.bw = 128; pl.bh = 1; pl.bb = 1;
  choose_box(g.ph[0].Wp, g.ph[0].Hp, g.B, 128, &pl.bw, &pl.bh, &pl.bb);
  pl.tw = (int)bg_cdiv(g.ph[0].Wp, pl.bw); pl.th = (int)bg_cdiv(g.ph[0].Hp, pl.bh); pl.tb = (int)bg_cdiv(g.B, pl.bb);
  pl.ntiles = pl.tw * pl.th * pl.tb * g.nphase * (g.Co / pl.BN);
  pl.ksplit = 1;
  static int sk_env = -1;
  if (sk_env < 0) { const char* e = getenv("BG_TC_SPLITK"); sk_env = e ? atoi(e) : 8; }
  const int sms = bg_num_sms();
  const int total = min_taps * (int)bg_cdiv(g.Ci, 64);
  if (sk_env > 1 && pl.ntiles < 4 * sms && total >= 12) {
    // microsecond model: waves x (stages per item x ~0.3 us + 1 us per item) + the reduce pass (launch + fp32 partial
    // write and read at ~4 TB/s); splitting pays for the 4x4 / 8x8 maps at small per-GPU batch, not for large outputs
    const double out_mb = (double)g.B * g.Ho * g.Wo * g.Co * 4.0 / 1e6;
    double best = (double)bg_cdiv(pl.ntiles, sms) * (total * 0.3 + 1.0);
    for (int S = 2; S <= sk_env && S <= 8 && total / S >= 6; ++S) {
      const double t = (double)bg_cdiv((long)pl.ntiles * S, sms) * ((double)total / S * 0.3 + 1.0) + 3.0 + (2.0 * S + 1.0) * out_mb / 4.0;
      if (t < 0.85 * best) { best = t; pl.ksplit = S; }
    }
  }
  pl.ok = 1;
  return pl;
}
size_t bg_tc_fwd_splitk_bytes(const BgGeom& g) {
  const TcFwdPlan pl = tc_fwd_plan(g);
  if (!pl.ok || pl.ksplit <= 1) return 0;
  return (size_t)pl.ksplit * g.B * g.Ho * g.Wo * g.Co * sizeof(float) + 256;
}

static int g_cnt_tc_fwd = 0, g_cnt_tc_wgrad = 0;   // launches taken by the tensor-core kernels (tests / bench)
// `in` must already be zero-pad semantics (reflect inputs are pre-padded by the caller, with
// tap offsets shifted accordingly through pad_shift_y/x).
int bg_tc_fwd_launch(const BgGeom& g, const void* in, int in_H, int in_W, int shift_y, int shift_x, const void* w_kmajor,
                     const float* bias, const void* residual, void* out, int act, void* sk_ws, size_t sk_bytes, const BgTcFold* fold_in,
                     cudaStream_t st, int* taken) {
  *taken = 0;
  const void* mask_only = (fold_in && !fold_in->ring) ? fold_in->mask : nullptr;
  const BgTcFold* fold = (fold_in && fold_in->ring) ? fold_in : nullptr;
  const TcFwdPlan pl = tc_fwd_plan(g);
  if (!pl.ok) return BG_OK;
  if (((uintptr_t)in | (uintptr_t)w_kmajor | (uintptr_t)out | (uintptr_t)residual) & 15) return BG_OK;
  static TcFwdParams P;      // large struct: keep it off the stack (host calls are single-threaded per process)
  memset(&P, 0, sizeof(P));
  const int BN = pl.BN;
  int bw = pl.bw, bh = pl.bh, bb = pl.bb;
  bool box_override = false;
  {
    const char* e = getenv("BG_TC_BOX");          // tuning knob: "bw,bh,bb" (product 128)
    int a = 0, b = 0, c = 0;
    if (e && sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && a * b * c == 128) { bw = a; bh = b; bb = c; box_override = true; }
    const char* dry = getenv("BG_TC_DRY");
    P.dry = (dry && dry[0] == '1') ? 1 : 0;
  }
  const int s = g.istep;
  int rc;
  for (int pa = 0; pa < s; ++pa)
    for (int pb = 0; pb < s; ++pb)
      if ((rc = make_act_map(&P.amap[pa * s + pb], in, g.B, in_H, in_W, g.Ci, s, pa, pb, bw, bh, bb))) return rc;
  for (int m = s * s; m < 4; ++m) P.amap[m] = P.amap[0];
  int maxw = 0;
  for (int ph = 0; ph < g.nphase; ++ph)
    for (int t = 0; t < g.ph[ph].ntaps; ++t) maxw = max(maxw, g.ph[ph].taps[t].widx);
  if ((rc = make_w_map(&P.bmap, w_kmajor, maxw + 1, g.Co, g.Ci, BN))) return rc;
  P.B = g.B; P.Ho = g.Ho; P.Wo = g.Wo; P.Co = g.Co;
  P.nphase = g.nphase; P.ostep = g.ostep;
  P.bw = bw; P.bh = bh; P.bb = bb;
  P.tw = (int)bg_cdiv(g.ph[0].Wp, bw); P.th = (int)bg_cdiv(g.ph[0].Hp, bh); P.tb = (int)bg_cdiv(g.B, bb);
  P.nkb = (int)bg_cdiv(g.Ci, 64);
  P.ktail = (int)bg_cdiv(g.Ci - (P.nkb - 1) * 64, 16);
  P.BN = BN; P.act = act;
  static int smem_kb = -1;                        // A/B knob (<= 224: + 1 KB alignment slack stays under the 227 KB per-block limit);
  // measured: 224 KB gives the halo layers a 4th stage but changes the step by < 0.3 % either way, so the default stays 200
  if (smem_kb < 0) { const char* e = getenv("BG_TC_SMEM_KB"); smem_kb = e ? atoi(e) : 200; if (smem_kb < 96 || smem_kb > 224) smem_kb = 200; }
  int ring_budget = smem_kb * 1024;
  // resident-weights mode (see TcFwdParams): one n-tile, no split-K, not a halo layer, all weight tiles fit beside >= 4 stages
  int bres_bytes = 0;
  {
    static int rb_env = -1;                       // A/B knob
    if (rb_env < 0) { const char* e = getenv("BG_TC_RESIDENT_B"); rb_env = e ? atoi(e) : 1; }
    const char* he0 = getenv("BG_TC_HALO");
    const bool halo_possible = (he0 ? atoi(he0) : 1) && g.nphase == 1 && s == 1 && bw == 128 && bh == 1 && bb == 1 && g.ph[0].ntaps > 1;
    const int region = (maxw + 1) * P.nkb * BN * 128;
    if (rb_env && g.Co == BN && pl.ksplit <= 1 && !halo_possible && !box_override && region <= 112 * 1024 &&
        (ring_budget - region) / 16384 >= 4) {
      P.resident_b = 1;
      P.nwidx = maxw + 1;
      bres_bytes = region;
      ring_budget -= region;
    }
  }
  // paired-tile mode (see TcFwdParams): streamed weights that are a sizeable part of a stage, two accumulators fit TMEM
  P.mt = 1;
  {
    static int pair_env = -1;                     // A/B knob
    if (pair_env < 0) { const char* e = getenv("BG_TC_PAIR"); pair_env = e ? atoi(e) : 1; }
    const char* he1 = getenv("BG_TC_HALO");
    const bool halo_possible = (he1 ? atoi(he1) : 1) && g.nphase == 1 && s == 1 && bw == 128 && bh == 1 && bb == 1 && g.ph[0].ntaps > 1;
    const long tiles_pp0 = (long)P.tw * P.th * P.tb;
    if (pair_env && !P.resident_b && pl.ksplit <= 1 && !halo_possible && !box_override && BN >= 64 && BN <= 128 && tiles_pp0 >= 2 &&
        tiles_pp0 * g.nphase * (g.Co / BN) >= 4L * bg_num_sms() && ring_budget / (32768 + BN * 128) >= 3)
      P.mt = 2;
  }
  const int stage_bytes = P.resident_b ? 16384 : P.mt * 16384 + BN * 128;
  // TMA-store epilogue for layers with a short main loop (see TcFwdParams): the staging slabs come out of the ring's budget
  {
    static int ts_env = -1;                       // A/B knob: 0 = never, 1 = automatic (default), 2 = whenever possible
    if (ts_env < 0) { const char* e = getenv("BG_TC_TMA_STORE"); ts_env = e ? atoi(e) : 1; }
    int min_taps2 = 1 << 30;
    for (int ph = 0; ph < g.nphase; ++ph) min_taps2 = min(min_taps2, g.ph[ph].ntaps);
    const long mma_cycles = (long)min_taps2 * bg_cdiv(g.Ci, 16) * (BN <= 128 ? 60 : BN / 2);   // issue-rate model of one tile
    const int half_bytes = ((BN / 16 + 1) / 2) * 4096;
    // (not in fold mode: boxes starting at negative coordinates are fine for TMA loads but fault on stores -- measured)
    static int ts_k = -1;                         // threshold of the issue-rate model, in units of 16 x BN cycles (tuning knob)
    if (ts_k < 0) { const char* e = getenv("BG_TC_TMA_STORE_K"); ts_k = e ? atoi(e) : 3; if (ts_k < 1) ts_k = 3; }
    const char* tsf = getenv("BG_TC_TMA_STORE_FOLD");       // A/B knob (read per call): bulk stores for the interior tiles of the fold mode
    const int tsf_env = tsf ? atoi(tsf) : 1;
    // (fold mode: stride-1 layers only -- measured on the strided dgrads the bulk stores of four interleaved phases lose 20 %)
    if (ts_env > 0 && (!fold || (tsf_env && g.ostep == 1) || tsf_env == 2) && pl.ksplit <= 1 && g.ostep >= 1 && (ts_env == 2 || mma_cycles < (long)ts_k * 16 * BN) &&
        (ring_budget - 2 * half_bytes) / stage_bytes >= 3) {
      P.tma_store = 1;
      P.ostage_half = (uint32_t)half_bytes;
      P.ostage_sets = (ring_budget - 4 * half_bytes) / stage_bytes >= 4 ? 2 : 1;
      ring_budget -= 2 * P.ostage_sets * half_bytes;
      for (int ph = 0; ph < g.nphase; ++ph) {
        if (fold) {
          // the kernel iterates the padded domain; the map covers the unpadded tensor from this phase's first interior row / column
          const int di = g.ph[ph].oy0 >= fold->pt ? 0 : (fold->pt - g.ph[ph].oy0 + g.ostep - 1) / g.ostep;
          const int dj = g.ph[ph].ox0 >= fold->pl ? 0 : (fold->pl - g.ph[ph].ox0 + g.ostep - 1) / g.ostep;
          P.o_di[ph] = di; P.o_dj[ph] = dj;
          const int y0 = g.ph[ph].oy0 + di * g.ostep - fold->pt, x0 = g.ph[ph].ox0 + dj * g.ostep - fold->pl;
          if (y0 >= fold->Hi || x0 >= fold->Wi) { P.omap[ph] = P.omap[0]; continue; }      // no interior pixel in this phase: never used
          if ((rc = make_act_map(&P.omap[ph], out, g.B, fold->Hi, fold->Wi, g.Co, g.ostep, y0, x0, bw, bh, bb, 16, CU_TENSOR_MAP_SWIZZLE_NONE)))
            return rc;
        } else if ((rc = make_act_map(&P.omap[ph], out, g.B, g.Ho, g.Wo, g.Co, g.ostep, g.ph[ph].oy0, g.ph[ph].ox0, bw, bh, bb, 16,
                                      CU_TENSOR_MAP_SWIZZLE_NONE)))
          return rc;
      }
      for (int ph = g.nphase; ph < 4; ++ph) P.omap[ph] = P.omap[0];
    }
  }
  int stages = ring_budget / stage_bytes;        // one persistent CTA per SM: a deep ring that never drains
  if (stages > 8) stages = 8;
  {
    static int st_env = -1;                       // tuning knob (tools/layer_bench.py)
    if (st_env < 0) { const char* e = getenv("BG_TC_FWD_STAGES"); st_env = e ? atoi(e) : 0; }
    if (st_env >= 2 && st_env < stages) stages = st_env;
  }
  if (stages < 2) return BG_OK;
  P.stages = stages;
  for (int ph = 0; ph < g.nphase; ++ph) {
    P.oy0[ph] = g.ph[ph].oy0; P.ox0[ph] = g.ph[ph].ox0; P.Hp[ph] = g.ph[ph].Hp; P.Wp[ph] = g.ph[ph].Wp;
    P.ntaps[ph] = g.ph[ph].ntaps;
    for (int t = 0; t < g.ph[ph].ntaps; ++t) {
      const int dy = g.ph[ph].taps[t].dy + shift_y, dx = g.ph[ph].taps[t].dx + shift_x;
      TcTap& tt = P.taps[ph][t];
      if (s == 1) { tt.map = 0; tt.dy = dy; tt.dx = dx; }
      else { const int pa = dy & 1, pb = dx & 1; tt.map = pa * 2 + pb; tt.dy = floordiv2(dy); tt.dx = floordiv2(dx); }
      tt.widx = g.ph[ph].taps[t].widx;
    }
  }
  P.bias = bias; P.residual = (const bf16*)residual; P.out = (bf16*)out;
  const long ntiles = bg_cdiv((long)P.tw * P.th * P.tb, (long)P.mt) * g.nphase * (g.Co / BN);      // work items
  const long out_numel = (long)g.B * g.Ho * g.Wo * g.Co;
  if (mask_only) {
    if (((uintptr_t)mask_only) & 15) return BG_OK;                         // (split-K is never combined with the mask: its reduce pass has none)
    P.mask = (const bf16*)mask_only;
  }
  if (fold) { P.fold = 1; P.f_pt = fold->pt; P.f_pl = fold->pl; P.f_Hi = fold->Hi; P.f_Wi = fold->Wi; P.ring = (bf16*)fold->ring; P.mask = (const bf16*)fold->mask; }
  P.ksplit = 1;
  if (pl.ksplit > 1 && !fold && !mask_only && !box_override && sk_ws && sk_bytes >= (size_t)pl.ksplit * out_numel * sizeof(float) &&
      (((uintptr_t)sk_ws) & 15) == 0) {
    P.ksplit = pl.ksplit; P.part = (float*)sk_ws; P.part_stride = out_numel;
  }
  {
    // A/B knob, read per call (tools/layer_sweep.py toggles it in one process).  (An L2 prefetch of the NEXT tile's residual / mask
    // rows from the epilogue warps was measured too: it cost 0.7 ms per step -- more requests on an already busy L1 -> crossbar port.)
    const char* e = getenv("BG_TC_AUX_BATCH");
    P.aux_batch = e ? atoi(e) : 1;
  }
  P.fd_S = make_fastdiv(P.ksplit); P.fd_ntn = make_fastdiv(g.Co / BN); P.fd_tw = make_fastdiv(P.tw); P.fd_th = make_fastdiv(P.th);
  P.fd_tpp = make_fastdiv((int)bg_cdiv((long)P.tw * P.th * P.tb, (long)P.mt));
  {
    const char* e = getenv("BG_TC_PHASE_MINOR");       // A/B knob, read per call
    P.phase_minor = (g.nphase == 4 && (e ? atoi(e) : 1)) ? 1 : 0;
  }
  P.nwork = (int)(ntiles * P.ksplit);
  dim3 grid((unsigned)min(ntiles * P.ksplit, (long)bg_num_sms()));
  // halo mode: stride-1 layer tiled in full 128-pixel row segments -> taps differing only in dx share one activation box
  {
    const char* he = getenv("BG_TC_HALO");        // tuning / A-B knob, read per call
    const int halo_env = he ? atoi(he) : 1;
    const int hstage = TC_HALO_ABYTES + 3 * BN * 128;
    int hstages = ring_budget / hstage;
    if (P.tma_store && P.ostage_sets == 2 && hstages < 4 && halo_env) {   // halo stages are larger: give the second set back
      ring_budget += 2 * (int)P.ostage_half;
      P.ostage_sets = 1;
      hstages = ring_budget / hstage;
      stages = ring_budget / stage_bytes;
      if (stages > 8) stages = 8;
      P.stages = stages;
    }
    if (hstages > 8) hstages = 8;
    if (halo_env && P.ksplit == 1 && g.nphase == 1 && s == 1 && bw == 128 && bh == 1 && bb == 1 && hstages >= 3) {
      TcTap srt[TC_MAX_TAPS];
      const int nt = g.ph[0].ntaps;
      for (int t = 0; t < nt; ++t) srt[t] = P.taps[0][t];
      for (int a = 1; a < nt; ++a)                        // insertion sort by (dy, dx)
        for (int b = a; b > 0 && (srt[b].dy < srt[b - 1].dy || (srt[b].dy == srt[b - 1].dy && srt[b].dx < srt[b - 1].dx)); --b) {
          TcTap tmp = srt[b]; srt[b] = srt[b - 1]; srt[b - 1] = tmp;
        }
      int ng = 0, span = 0, shared = 0;
      bool ok = true;
      for (int t = 0; t < nt && ok;) {
        if (ng == TC_MAX_GROUPS) { ok = false; break; }
        TcGroup& gr = P.groups[ng];
        gr.dy = srt[t].dy; gr.dx0 = srt[t].dx; gr.ntap = 0;
        while (t < nt && gr.ntap < 3 && srt[t].dy == gr.dy && srt[t].dx == gr.dx0 + gr.ntap) {
          gr.widx[gr.ntap] = srt[t].widx; gr.dxo[gr.ntap] = gr.ntap; gr.ntap++; t++;
        }
        span = max(span, gr.ntap - 1);
        shared += gr.ntap - 1;
        ng++;
      }
      if (ok && shared > 0) {
        P.ngroups = ng; P.halo_w = 128 + span; P.stages = hstages;
        if ((rc = make_act_map(&P.amap[0], in, g.B, in_H, in_W, g.Ci, 1, 0, 0, P.halo_w, 1, 1))) return rc;
        P.ostage_off = (uint32_t)(hstages * hstage);
        const size_t hsmem = (size_t)hstages * hstage + 1024 + (P.tma_store ? 2 * P.ostage_sets * P.ostage_half : 0);
#define TC_LAUNCH_H(COLS)                                                                                        \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      BG_CHECK_CUDA(cudaFuncSetAttribute(tc_fwd_halo_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    tc_fwd_halo_kernel<COLS><<<grid, 320, hsmem, st>>>(P);                                                       \
  } while (0)
        const int hcols = 2 * BN;
        if (hcols <= 32) TC_LAUNCH_H(32);
        else if (hcols <= 64) TC_LAUNCH_H(64);
        else if (hcols <= 128) TC_LAUNCH_H(128);
        else if (hcols <= 256) TC_LAUNCH_H(256);
        else TC_LAUNCH_H(512);
#undef TC_LAUNCH_H
        BG_LAUNCH_CHECK();
        g_cnt_tc_fwd++;
        *taken = 1;
        return BG_OK;
      }
    }
  }
  P.bres_off = (uint32_t)(stages * stage_bytes);
  P.ostage_off = P.bres_off + (uint32_t)bres_bytes;
  const size_t smem = (size_t)stages * stage_bytes + bres_bytes + 1024 + (P.tma_store ? 2 * P.ostage_sets * P.ostage_half : 0);
#define TC_LAUNCH(COLS)                                                                                          \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      BG_CHECK_CUDA(cudaFuncSetAttribute(tc_fwd_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    tc_fwd_kernel<COLS><<<grid, 320, smem, st>>>(P);                                                             \
  } while (0)
  const int cols = 2 * P.mt * BN;                // double-buffered accumulator(s)
  if (cols <= 32) TC_LAUNCH(32);
  else if (cols <= 64) TC_LAUNCH(64);
  else if (cols <= 128) TC_LAUNCH(128);
  else if (cols <= 256) TC_LAUNCH(256);
  else TC_LAUNCH(512);
#undef TC_LAUNCH
  BG_LAUNCH_CHECK();
  if (P.ksplit > 1) {
    const long n8 = out_numel / 8;
    const int blocks = (int)min(bg_cdiv(n8, 256), (long)bg_num_sms() * 8);
    tc_splitk_reduce_kernel<<<blocks, 256, 0, st>>>(P.part, P.part_stride, P.ksplit, bias, (const bf16*)residual, (bf16*)out, n8, g.Co, act);
    BG_LAUNCH_CHECK();
  }
  g_cnt_tc_fwd++;
  *taken = 1;
  return BG_OK;
}

// ------------------------------------------------------------------------------------------------
// wgrad kernel: D[128 in-channels, BN out-channels] (fp32, TMEM) = sum over pixel boxes of
//   A^T[64 pixels, 128 ch] (two 64-channel TMA boxes of the gathered tensor, tap-shifted, OOB = 0)
//   x B[64 pixels, BN ch] (ceil(BN/64) boxes of the paired tensor).  Both operands are MN-major
// (the contraction runs over pixels, channels are contiguous): SWIZZLE_128B, LBO = 8192 (next
// 64-channel block), SBO = 1024 (next 8 pixels).  Deterministic split-K over pixel boxes (grid.z).
// ------------------------------------------------------------------------------------------------
struct TcWgParams {
  CUtensorMap amap[4];
  CUtensorMap bmap;
  int bw, bh, bb, tw, th, tb;   // pixel box (product 64) and boxes per grid
  int Ci, Co, BN, nblkB;        // M total, N total, N tile, 64-channel blocks per B tile
  int ntile_n, nsplit, nboxes, stages;
  int ntaps;
  TcTap taps[TC_MAX_TAPS];
  float* part;
  long zstride, wtap, wk;
};

// TPC taps per CTA: they share every Og (B operand) stage -- the shared tile is loaded once instead of once per tap, which
// is what bounds the C <= 192 layers (L2 -> SM ingest) -- and each tap owns BN accumulator columns of TMEM.
template <int TMEM_COLS, int TPC>
__global__ void __launch_bounds__(192, 1) tc_wgrad_kernel(const __grid_constant__ TcWgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[8];
  __shared__ __align__(8) uint64_t empty_bar[8];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t B_OFF = (uint32_t)TPC * 16384u;                  // stage layout: [A tap 0][A tap 1]...[B]
  const uint32_t stage_bytes = B_OFF + (uint32_t)p.nblkB * 8192u;
  const int m0 = (blockIdx.x / p.ntile_n) * 128, n0 = (blockIdx.x % p.ntile_n) * p.BN;
  const int tap0 = blockIdx.y * TPC;
  const int nt = min(TPC, p.ntaps - tap0);                             // taps handled by this CTA
  const int box_beg = (int)(((long)blockIdx.z * p.nboxes) / p.nsplit);
  const int box_end = (int)(((long)(blockIdx.z + 1) * p.nboxes) / p.nsplit);
  const int total = box_end - box_beg;

  if (warp == 0 && elect_one_sync()) {
    for (int m = 0; m < 4; ++m) tma_prefetch_desc(&p.amap[m]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&accum_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(smem_u32(&tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const CUtensorMap* am[TPC];
      int tdx[TPC], tdy[TPC];
#pragma unroll
      for (int t2 = 0; t2 < TPC; ++t2) {
        const TcTap tp = p.taps[min(tap0 + t2, p.ntaps - 1)];
        am[t2] = &p.amap[tp.map]; tdx[t2] = tp.dx; tdy[t2] = tp.dy;
      }
      const uint32_t tx = (uint32_t)nt * 16384u + (uint32_t)p.nblkB * 8192u;
      // box coordinates advance incrementally (the issue thread is a single dependent instruction stream)
      int t = box_beg;
      int j0 = (t % p.tw) * p.bw; t /= p.tw;
      int i0 = (t % p.th) * p.bh; t /= p.th;
      int b0 = t * p.bb;
      const int jend = p.tw * p.bw, iend = p.th * p.bh;
      uint32_t s = 0, par = 1, sa = smem_base;
      for (int it = 0; it < total; ++it) {
        mbar_wait(empty0 + s * 8u, par);
        const uint32_t fb = full0 + s * 8u;
        mbar_expect_tx(fb, tx);
#pragma unroll
        for (int t2 = 0; t2 < TPC; ++t2)
          if (t2 < nt) {
            tma_load_4d(sa + (uint32_t)t2 * 16384u, am[t2], fb, m0, j0 + tdx[t2], i0 + tdy[t2], b0);
            tma_load_4d(sa + (uint32_t)t2 * 16384u + 8192u, am[t2], fb, m0 + 64, j0 + tdx[t2], i0 + tdy[t2], b0);
          }
        for (int nb = 0; nb < p.nblkB; ++nb)
          tma_load_4d(sa + B_OFF + (uint32_t)nb * 8192u, &p.bmap, fb, n0 + nb * 64, j0, i0, b0);
        j0 += p.bw;
        if (j0 >= jend) { j0 = 0; i0 += p.bh; if (i0 >= iend) { i0 = 0; b0 += p.bb; } }
        if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; sa = smem_base; } else sa += stage_bytes;
      }
    }
  } else if (warp == 1) {
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_bf16(128, p.BN, 1, 1);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      // MN-major SWIZZLE_128B operands: LBO = 8192 (next 64-channel block), SBO = 1024 (next 8 pixels); one MMA covers
      // 16 pixels = 2048 B = 128 descriptor units
      const uint32_t a_lo0 = umma_desc_lo(smem_base, 8192), stage_q = stage_bytes >> 4;
      const uint32_t hi = umma_desc_hi_mnmajor_sw128(1024);
      const uint32_t bn = (uint32_t)p.BN;
      uint32_t s = 0, par = 0, a_lo = a_lo0, acc = 0u;
      for (int it = 0; it < total; ++it) {
        mbar_wait(full0 + s * 8u, par);
        tc_fence_after();
        const uint32_t b_lo = a_lo + (B_OFF >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int t2 = 0; t2 < TPC; ++t2)
            if (t2 < nt)
              umma_bf16_w(tmem_base + (uint32_t)t2 * bn, a_lo + (uint32_t)t2 * 1024u + (uint32_t)k * 128u, hi, b_lo + (uint32_t)k * 128u, hi,
                          idesc, k == 0 ? acc : 1u);
        }
        acc = 1u;
        umma_commit(empty0 + s * 8u);
        if (++s == (uint32_t)p.stages) { s = 0; par ^= 1u; a_lo = a_lo0; } else a_lo += stage_q;
      }
      umma_commit(smem_u32(&accum_bar));
    }
  } else {
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    if (total > 0) {
      mbar_wait(smem_u32(&accum_bar), 0);
      tc_fence_after();
    }
#pragma unroll
    for (int t2 = 0; t2 < TPC; ++t2) {
      if (t2 >= nt) break;
      float* dst = p.part + (long)blockIdx.z * p.zstride + (long)p.taps[tap0 + t2].widx * p.wtap + (long)m * p.wk + n0;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16];
        if (total > 0) tmem_ld_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t2 * p.BN + c0), r);
        else {
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = 0u;
        }
        if (m < p.Ci) {
          float4* o = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = make_float4(__uint_as_float(r[4 * e]), __uint_as_float(r[4 * e + 1]), __uint_as_float(r[4 * e + 2]),
                               __uint_as_float(r[4 * e + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

extern "C" int bg_debug_counters(int* tc_fwd, int* tc_wgrad) {
  if (tc_fwd) *tc_fwd = g_cnt_tc_fwd;
  if (tc_wgrad) *tc_wgrad = g_cnt_tc_wgrad;
  return BG_OK;
}

// split-K factor: the grid is tiles*ns CTAs on `slots` = SMs x resident CTAs.  Minimise a microsecond model of
//   waves x (boxes per CTA x t_box + t_fixed) + t_reduce
// t_box: one 64-pixel stage at this CTA's share of the L2->SM bandwidth (~80 B/ns per SM); t_fixed: prologue, pipeline
// fill and the fp32 accumulator drain at the SM's share of HBM write bandwidth; t_reduce: the split-K reduction pass
// (absent for ns == 1).  297 CTAs on 296 slots cost two waves; splitting a 1536x1536 layer costs more than it buys.
static int wg_nsplit(int tiles, int nboxes, int ctas_per_sm, int BN, int TPC, long numel) {
  const long slots = (long)bg_num_sms() * ctas_per_sm;
  long mx = nboxes / 4;
  if (mx > 64) mx = 64;
  if (mx < 1) mx = 1;
  const double t_box = (TPC * 16384.0 + ((BN + 63) / 64) * 8192.0) / (80e3 / ctas_per_sm);       // us
  const double t_fixed = 2.0 + TPC * 128.0 * BN * 4.0 / (43e3 / ctas_per_sm);                      // us
  long best = 1;
  double best_cost = 1e30;
  for (long ns = 1; ns <= mx; ++ns) {
    double cost = (double)bg_cdiv((long)tiles * ns, slots) * ((double)bg_cdiv(nboxes, ns) * t_box + t_fixed);
    if (ns > 1) cost += 3.0 + (double)(ns + 1) * numel * 4.0 / 5e6;
    if (cost < best_cost * 0.999) { best_cost = cost; best = ns; }
  }
  return (int)best;
}

struct WgPlan { int ok, BN, bw, bh, bb, tw, th, tb, nboxes, tiles, nsplit, TPC, stages; };

static WgPlan wg_plan(const BgGeom& g) {
  WgPlan pl;
  memset(&pl, 0, sizeof(pl));
  if (!tc_device_ok()) return pl;
  if (g.nphase != 1 || g.ostep != 1 || g.wn != 1) return pl;
  if (g.Ci % 8 != 0 || g.Ci < 16 || g.Co % 16 != 0 || g.wk % 4 != 0) return pl;
  if (g.ph[0].ntaps < 1 || g.ph[0].ntaps > TC_MAX_TAPS) return pl;
  pl.BN = pick_bn(g.Co);
  if (!pl.BN) return pl;
  pl.bw = 64; pl.bh = 1; pl.bb = 1;
  choose_box(g.ph[0].Wp, g.ph[0].Hp, g.B, 64, &pl.bw, &pl.bh, &pl.bb);
  pl.tw = (int)bg_cdiv(g.ph[0].Wp, pl.bw); pl.th = (int)bg_cdiv(g.ph[0].Hp, pl.bh); pl.tb = (int)bg_cdiv(g.B, pl.bb);
  pl.nboxes = pl.tw * pl.th * pl.tb;
  // taps per CTA (TPC): sharing the Og stage pays when the layer is ingest-bound, i.e. when one stage's MMAs
  // (4 x max(64, BN/2) cycles per tap) take less time than loading it (~48 B/clk/SM); needs TPC*BN <= 512 TMEM columns
  // and at least two stages of shared memory
  {
    const int nt = g.ph[0].ntaps, nblk = (pl.BN + 63) / 64;
    static int tpc_env = -1;
    if (tpc_env < 0) { const char* e = getenv("BG_WGRAD_TPC"); tpc_env = e ? atoi(e) : 4; }
    int best = 1;
    for (int t = 2; t <= tpc_env && t <= 4; ++t) {
      if (t * pl.BN > 512 || t > nt) continue;
      if (2 * (t * 16384 + nblk * 8192) > 200 * 1024) continue;
      const double mma_cyc = 4.0 * (pl.BN / 2 > 64 ? pl.BN / 2 : 64), load_cyc = (16384.0 + nblk * 8192.0) / 48.0;
      if (mma_cyc >= load_cyc) continue;                       // tensor-pipe bound already
      const int idle_t = (int)bg_cdiv(nt, t) * t - nt, idle_b = (int)bg_cdiv(nt, best) * best - nt;
      if (best == 1 || idle_t <= idle_b) best = t;             // least idle accumulator slots, prefer larger
    }
    pl.TPC = best;
  }
  {
    const int stage_bytes = pl.TPC * 16384 + ((pl.BN + 63) / 64) * 8192;
    int stages = (stage_bytes <= 32 * 1024) ? 3 : 4;            // 2 CTAs per SM for the small tiles
    if ((size_t)stages * stage_bytes > 200 * 1024) stages = (200 * 1024) / stage_bytes;
    if (stages < 2) stages = 2;
    pl.stages = stages;
  }
  pl.tiles = (int)(bg_cdiv(g.Ci, 128) * (g.Co / pl.BN) * bg_cdiv(g.ph[0].ntaps, pl.TPC));
  pl.nsplit = wg_nsplit(pl.tiles, pl.nboxes, (pl.stages * (pl.TPC * 16384 + ((pl.BN + 63) / 64) * 8192) <= 100 * 1024) ? 2 : 1, pl.BN, pl.TPC,
                        (long)g.ph[0].ntaps * g.Ci * g.Co);
  pl.ok = 1;
  return pl;
}

size_t bg_tc_conv_wgrad_ws(const BgGeom& g, const bg_conv_desc* d) {
  WgPlan pl = wg_plan(g);
  if (!pl.ok) return 0;
  size_t pad = 0;
  if (!d->transposed && d->pad_mode == BG_PAD_REFLECT && (d->pad_t | d->pad_b | d->pad_l | d->pad_r))
    pad = ((size_t)d->B * (d->Hi + d->pad_t + d->pad_b) * (d->Wi + d->pad_l + d->pad_r) * d->Ci * 2 + 1023) / 1024 * 1024;
  size_t part = pl.nsplit > 1 ? (size_t)pl.nsplit * d->kh * d->kw * d->Ci * d->Co * sizeof(float) : 0;
  return pad + part + 1024;
}

__global__ void splitk_reduce_kernel2(const float* __restrict__ part, float* __restrict__ dw, long n, long zstride, int nsplit) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float s = 0.f;
    for (int z = 0; z < nsplit; ++z) s += part[(long)z * zstride + i];
    dw[i] = s;
  }
}

int bg_tc_conv_wgrad(const BgGeom& g, const bg_conv_desc* d, const void* in, const void* og, float* dw, void* ws, size_t ws_bytes,
                     cudaStream_t st, int* taken) {
  *taken = 0;
  WgPlan pl = wg_plan(g);
  if (!pl.ok) return BG_OK;
  if (((uintptr_t)in | (uintptr_t)og | (uintptr_t)dw) & 15) return BG_OK;
  if (ws_bytes < bg_tc_conv_wgrad_ws(g, d) || (!ws && bg_tc_conv_wgrad_ws(g, d) > 1024)) return BG_OK;
  static TcWgParams P;
  memset(&P, 0, sizeof(P));
  int rc;
  // reflect-padded input: pre-pad (the gathered tensor is x for a plain conv)
  const void* in_ptr = in;
  int inH = g.Hi, inW = g.Wi, sy = 0, sx = 0;
  char* wsp = (char*)ws;
  if (!d->transposed && d->pad_mode == BG_PAD_REFLECT && (d->pad_t | d->pad_b | d->pad_l | d->pad_r)) {
    rc = bg_reflect_pad(in, wsp, d->B, d->Hi, d->Wi, d->Ci, d->pad_t, d->pad_l, d->pad_b, d->pad_r, BG_BF16, (void*)st);
    if (rc) return rc;
    in_ptr = wsp;
    inH += d->pad_t + d->pad_b; inW += d->pad_l + d->pad_r; sy = d->pad_t; sx = d->pad_l;
    wsp += ((size_t)d->B * inH * inW * d->Ci * 2 + 1023) / 1024 * 1024;
  }
  const int s = g.istep;
  for (int pa = 0; pa < s; ++pa)
    for (int pb = 0; pb < s; ++pb)
      if ((rc = make_act_map(&P.amap[pa * s + pb], in_ptr, g.B, inH, inW, g.Ci, s, pa, pb, pl.bw, pl.bh, pl.bb))) return rc;
  for (int m = s * s; m < 4; ++m) P.amap[m] = P.amap[0];
  if ((rc = make_act_map(&P.bmap, og, g.B, g.Ho, g.Wo, g.Co, 1, 0, 0, pl.bw, pl.bh, pl.bb))) return rc;
  P.bw = pl.bw; P.bh = pl.bh; P.bb = pl.bb; P.tw = pl.tw; P.th = pl.th; P.tb = pl.tb;
  P.Ci = g.Ci; P.Co = g.Co; P.BN = pl.BN; P.nblkB = (pl.BN + 63) / 64;
  P.ntile_n = g.Co / pl.BN; P.nsplit = pl.nsplit; P.nboxes = pl.nboxes;
  const int stage_bytes = pl.TPC * 16384 + P.nblkB * 8192;
  const int stages = pl.stages;
  P.stages = stages;
  P.ntaps = g.ph[0].ntaps;
  for (int t = 0; t < P.ntaps; ++t) {
    const int dy = g.ph[0].taps[t].dy + sy, dx = g.ph[0].taps[t].dx + sx;
    TcTap& tt = P.taps[t];
    if (s == 1) { tt.map = 0; tt.dy = dy; tt.dx = dx; }
    else { const int pa = dy & 1, pb = dx & 1; tt.map = pa * 2 + pb; tt.dy = floordiv2(dy); tt.dx = floordiv2(dx); }
    tt.widx = g.ph[0].taps[t].widx;
  }
  const long numel = (long)d->kh * d->kw * d->Ci * d->Co;
  P.part = pl.nsplit > 1 ? (float*)wsp : dw;
  P.zstride = numel; P.wtap = g.wtap; P.wk = g.wk;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  dim3 grid((unsigned)(bg_cdiv(g.Ci, 128) * P.ntile_n), (unsigned)bg_cdiv(P.ntaps, pl.TPC), (unsigned)pl.nsplit);
#define TC_LAUNCH_WG(COLS, T)                                                                                    \
  do {                                                                                                           \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      BG_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel<COLS, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024)); \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    tc_wgrad_kernel<COLS, T><<<grid, 192, smem, st>>>(P);                                                        \
  } while (0)
#define TC_LAUNCH_WG_T(T)                                                                                        \
  do {                                                                                                           \
    const int wcols = T * pl.BN;                                                                                 \
    if (wcols <= 32) TC_LAUNCH_WG(32, T);                                                                        \
    else if (wcols <= 64) TC_LAUNCH_WG(64, T);                                                                   \
    else if (wcols <= 128) TC_LAUNCH_WG(128, T);                                                                 \
    else if (wcols <= 256) TC_LAUNCH_WG(256, T);                                                                 \
    else TC_LAUNCH_WG(512, T);                                                                                   \
  } while (0)
  if (pl.TPC == 1) TC_LAUNCH_WG_T(1);
  else if (pl.TPC == 2) TC_LAUNCH_WG_T(2);
  else if (pl.TPC == 3) TC_LAUNCH_WG_T(3);
  else TC_LAUNCH_WG_T(4);
#undef TC_LAUNCH_WG_T
#undef TC_LAUNCH_WG
  BG_LAUNCH_CHECK();
  if (pl.nsplit > 1) {
    const int blocks = (int)min(bg_cdiv(numel, 256), (long)bg_num_sms() * 8);
    splitk_reduce_kernel2<<<blocks, 256, 0, st>>>(P.part, dw, numel, numel, pl.nsplit);
    BG_LAUNCH_CHECK();
  }
  g_cnt_tc_wgrad++;
  *taken = 1;
  return BG_OK;
}
