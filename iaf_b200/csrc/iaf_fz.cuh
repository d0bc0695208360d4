// Fused IAF step for stacks with ONE hidden layer (depth_ar = 1: BASELINE configs C1 / C2a), second generation.
// Included by iaf_tc.cu (shares its PTX wrappers, parameter structs and the weight-prep kernel).
//
// Same formulation as iaf_tc_kernel (slot stream, tap = slot shift, fp16 hi/lo split operands, fp32 accumulation in
// TMEM), but a different schedule, designed around what bounded the first kernel (profiles/r1_c2a_ncu.md: every warp
// stalled on the same global loads and the E0 -> M1 -> E0 chain through a two-tile ring):
//
//   * INDEPENDENT OVERLAPPED TILES.  A tile emits TO = 128 - MIR output slots but computes the hidden layer for all
//     128 rows of its own z window (WIN = 128 + MIR slots), so the MIR halo rows the heads need are recomputed instead
//     of borrowed from the neighbour tile: no cross-tile dependency, no mirrored ring, no halo tile, and two tiles are
//     in flight with two plain 128-slot h buffers (MIR / 128 = 14 % more stage-1 rows at 16x16).
//   * WARP-SPECIALISED PRODUCERS.  All global LOADS of the hidden layer live in loader warps that run up to two tiles
//     ahead of the pipeline and may stall as long as they like.  Per tile they
//       - turn the fp32 z window into the fp16 hi/lo operand window in shared memory (layer mode: the posterior sample),
//       - write context + bias (+ Theano pad-channel terms) INTO the stage-0 accumulator with tcgen05.st before the MMAs
//         run, which then accumulate on top of it (accumulate = 1 from the first instruction; a loader warp is tied to
//         its TMEM lane quadrant).  The hidden epilogue never touches global memory.
//   * EPILOGUE WARPS only move TMEM -> registers -> shared memory (E0: elu, hi/lo split, next operand) or -> global
//     stores (E1: affine update, outputs, deterministic partial sums); E0(k) runs while the tensor pipe works on
//     M0(k+1), E1(k-1) while it works on M1(k).
//   * one MMA warp (convergent, one elected lane) issues M0(0), then M0(k+1), M1(k) alternately; a reducer warp folds
//     the per-tile partial sums exactly as in iaf_tc_kernel.
#pragma once

#ifndef FZ_EPI
#define FZ_EPI 8            // epilogue warps (multiple of 4: TMEM lane quadrants)
#endif
#ifndef FZ_LD
#define FZ_LD 8             // loader warps (multiple of 4: the context part writes TMEM, so a warp is tied to its lane quadrant)
#endif
#define FZ_W_LD0 FZ_EPI
#define FZ_W_MMA (FZ_EPI + FZ_LD)
#define FZ_W_RED (FZ_W_MMA + 1)
#define FZ_W_TMA (FZ_W_MMA + 2)   // bulk-copy producer of the staged z window (idle in the gathered variant)
#define FZ_THREADS ((FZ_W_TMA + 1) * 32)
#define FZ_LTHREADS (FZ_LD * 32)
#define FZ_ZB 3             // (slot, chunk) items of the z window per loader thread (FZ_ZB * loader threads >= items)
#define FZ_CXG 2            // context channel groups (of 16) per loader warp (FZ_CXG * FZ_LGS * 16 >= hidden width)
#define FZ_LGS (FZ_LD / 4)  // loader warps sharing one lane quadrant (they split the context channel groups)
#define FZ_CGS (FZ_EPI / 4) // epilogue warps sharing one lane quadrant

enum {
  FB_W = 0,         // bias tables + stage-0 weights have landed (bulk copies); the stage-1 weights: FB_W1
  FB_ZFULL = 1,     // + b: z window b written
  FB_ZEMPTY = 3,    // + b: M0 committed (z window b read)
  FB_A0_INIT = 5,   // + b: context + bias written into accumulator b of stage 0
  FB_A0_FULL = 7,   // + b: M0 committed
  FB_A0_EMPTY = 9,  // + b: E0 has read it
  FB_H_FULL = 11,   // + b: E0 has written h buffer b
  FB_H_EMPTY = 13,  // + b: M1 committed (h buffer b read)
  FB_A1_FULL = 15,  // + b
  FB_A1_EMPTY = 17, // + b
  FB_PART = 19,     // + tile parity
  FB_PART_EMPTY = 21,
  FB_ZST_FULL = 23, // + b, staged variant: the TMA copies of a tile's fp32 z rows have landed in staging buffer b
  FB_ZST_EMPTY = 25,// + b: the loader warps have read them
  FB_W1 = 27,       // stage-1 (heads) weights have landed: only M1(0) waits for them, M0(0) starts 80 KB earlier
  FB_COUNT = 28
};

// Optional wait-time probe (compile with -DIAF_FZ_PROBE; development aid): the lead lane of each role in CTA 1 accumulates
// the cycles it spends in each mbarrier wait and in between, and dumps them at the end (iaf_fz_probe_dump()).
#ifdef IAF_FZ_PROBE
__device__ long long g_fz_probe[4][8];
#define PROBE_DECL long long pr_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_ = clock64();
#define PROBE(slot) { const long long n_ = clock64(); pr_[slot] += n_ - pt_; pt_ = n_; }
#define PROBE_DUMP(role) if (blockIdx.x == 1 && lane == 0) { for (int i_ = 0; i_ < 8; ++i_) g_fz_probe[role][i_] = pr_[i_]; }
#else
#define PROBE_DECL
#define PROBE(slot)
#define PROBE_DUMP(role)
#endif

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMA tiled copy of one 4-D box (global -> shared), completing on an mbarrier
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

// Incremental slot decoding for the loaders: a thread's slots advance by TO per tile, so (sample n, offset r inside the
// sample) is carried from tile to tile and only the row / column split is recomputed (one multiply-high).
__device__ __forceinline__ SlotInfo decode_nr(const IafTcParams& p, int n, int r, int HW) {
  SlotInfo si;
  si.n = n;
  si.y = fast_div(r, p.Wp, p.mg_wp);
  si.x = r - si.y * p.Wp;
  si.valid = (n < p.B) && (si.y < p.H) && (si.x < p.W);
  const int pix = si.y * p.W + si.x;
  si.gp = p.flip ? HW - 1 - pix : pix;
  return si;
}
__device__ __forceinline__ void advance_nr(const IafTcParams& p, int& n, int& r, int step) {
  r += step;
  while (r >= p.SPS) { r -= p.SPS; ++n; }
}

// Rows of the (at most two) samples a tile's z window touches, for the staged variant.  Stream slot s -> sample s / SPS,
// offset r = s % SPS, stream row y = r / Wp (row H is the zero row and is never fetched).
struct ZstGeo {
  int n0, n1;        // samples (n1 = n0 + 1)
  int y0, rows0;     // sample n0: first stream row in the window, number of image rows staged
  int rows1;         // sample n1: image rows 0 .. rows1-1 staged
};
__device__ __forceinline__ ZstGeo zst_geometry(const IafTcParams& p, int s0) {
  ZstGeo g;
  g.n0 = fast_div(s0, p.SPS, p.mg_sps);
  const int r0 = s0 - g.n0 * p.SPS;
  g.y0 = fast_div(r0, p.Wp, p.mg_wp);
  const int s1 = s0 + p.WIN - 1;
  const int nb = fast_div(s1, p.SPS, p.mg_sps);
  const int yb = min(p.H - 1, fast_div(s1 - nb * p.SPS, p.Wp, p.mg_wp));
  g.n1 = g.n0 + 1;
  const int ylast0 = (nb == g.n0) ? yb : p.H - 1;
  g.rows0 = (g.n0 < p.B && g.y0 <= ylast0) ? ylast0 - g.y0 + 1 : 0;
  g.rows1 = (nb > g.n0 && g.n1 < p.B) ? yb + 1 : 0;
  return g;
}

template <bool PADW, int MODE, int NLT, int THW>
__global__ void __launch_bounds__(FZ_THREADS, 1) iaf_fz_kernel(const __grid_constant__ IafTcParams p) {
  const int HW = THW ? THW : p.HW;
  // staged variant: z reaches shared memory by bulk copies (the 16x16 instantiation, step / multiconv modes)
  constexpr bool ZST = (THW == 256) && (MODE != IAF_MODE_LAYER);
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[FB_COUNT];
  __shared__ uint32_t s_tmem;
  TL_DECL

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int t0 = (int)((long long)blockIdx.x * p.NT / G);
  const int t1 = (int)((long long)(blockIdx.x + 1) * p.NT / G);
  const int nt = t1 - t0;
  const IafTcStage& S0 = p.st[0];
  const IafTcStage& S1 = p.st[1];
  const int TO = p.TO;

  if (p.dbg & 64) return;  // timing only: the cost of launching this grid with its shared-memory footprint
  // ---- one-time setup (programmatic dependent launch: see iaf_tc_kernel).  Nothing here waits for global data: the
  // weights AND the bias tables arrive by bulk copy behind one mbarrier that their consumers wait on at first use.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == FZ_W_MMA) {
    tmem_alloc(&s_tmem, (uint32_t)p.tmem_cols);
    if (lane == 0) {
      mbar_init(&bars[FB_W], 1);
      mbar_init(&bars[FB_W1], 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars[FB_ZST_FULL + b], 1);
        mbar_init(&bars[FB_ZST_EMPTY + b], FZ_LD);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&bars[FB_ZFULL + b], FZ_LD);
        mbar_init(&bars[FB_ZEMPTY + b], 1);
        mbar_init(&bars[FB_A0_INIT + b], FZ_LD);
        mbar_init(&bars[FB_A0_FULL + b], 1);
        mbar_init(&bars[FB_A0_EMPTY + b], FZ_EPI);
        mbar_init(&bars[FB_H_FULL + b], FZ_EPI);
        mbar_init(&bars[FB_H_EMPTY + b], 1);
        mbar_init(&bars[FB_A1_FULL + b], 1);
        mbar_init(&bars[FB_A1_EMPTY + b], FZ_EPI);
        mbar_init(&bars[FB_PART + b], FZ_EPI);
        mbar_init(&bars[FB_PART_EMPTY + b], 1);
      }
      fence_barrier_init();
      asm volatile("griddepcontrol.wait;" ::: "memory");
      const uint32_t tab0 = 5u * (uint32_t)S0.N * 4u, tab1 = 5u * (uint32_t)S1.N * 4u;
      const bool now = (p.dbg & 32) != 0;  // timing only: no weight load
      mbar_expect_tx(&bars[FB_W], tab0 + tab1 + (now ? 0u : 2u * (uint32_t)S0.w_bytes));
      mbar_expect_tx(&bars[FB_W1], now ? 0u : 2u * (uint32_t)S1.w_bytes);
      // bias tables [5][N] (row 0 bias, rows 1..4 the Theano pad-channel weights; zero for the TF variant)
      bulk_g2s(smem + S0.sm_bias, S0.bias, tab0, &bars[FB_W]);
      bulk_g2s(smem + S1.sm_bias, S1.bias, tab1, &bars[FB_W]);
      for (int j = 0; j < (now ? 0 : 2); ++j) {
        for (int off = 0; off < 2 * p.st[j].w_bytes; off += 32768) {
          const uint32_t n = (uint32_t)min(32768, 2 * p.st[j].w_bytes - off);
          bulk_g2s(smem + p.st[j].sm_whi + off, reinterpret_cast<const uint8_t*>(p.st[j].whi) + off, n, &bars[j ? FB_W1 : FB_W]);
        }
      }
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (p.dbg & 128) {  // timing only: prologue + teardown, no roles
    if (warp == FZ_W_MMA) { mbar_wait(&bars[FB_W], 0); mbar_wait(&bars[FB_W1], 0); }
  } else
  if (warp == FZ_W_MMA) {
    // =====================================================================================
    // MMA warp: M0(0), then M0(k+1) and M1(k) alternately; one elected lane issues
    // =====================================================================================
    PROBE_DECL
    mbar_wait(&bars[FB_W], 0);
    PROBE(6)
    const uint32_t idesc0 = umma_idesc(S0.N), idesc0m = umma_idesc(2 * S0.N);
    const uint32_t idesc1 = umma_idesc(S1.N), idesc1m = umma_idesc(2 * S1.N);
    // stage 0 operands: z window (plane pitch WIN), weights image [K/8][2N][8]
    const uint32_t a0_plane = (uint32_t)S0.in_slots * 16u;
    const uint32_t a0h_0 = umma_desc_lo(smem_u32(smem + S0.sm_in), a0_plane);
    const uint32_t a0l_0 = umma_desc_lo(smem_u32(smem + S0.sm_in) + (uint32_t)(S0.cin >> 3) * a0_plane, a0_plane);
    const uint32_t zbuf_step = (uint32_t)p.z_bytes >> 4;  // second z window, in descriptor units
    const uint32_t b0_plane = (uint32_t)(2 * S0.N) * 16u;
    const uint32_t b0h = umma_desc_lo(smem_u32(smem + S0.sm_whi), b0_plane);
    const uint32_t b0l = umma_desc_lo(smem_u32(smem + S0.sm_whi) + (uint32_t)S0.N * 16u, b0_plane);
    const uint32_t a0_kstep = (2u * a0_plane) >> 4, b0_kstep = (2u * b0_plane) >> 4;
    // stage 1 operands: h buffer k & 1 (plane pitch 128 slots)
    const uint32_t a1_plane = (uint32_t)S1.in_slots * 16u;
    const uint32_t b1_plane = (uint32_t)(2 * S1.N) * 16u;
    const uint32_t b1h = umma_desc_lo(smem_u32(smem + S1.sm_whi), b1_plane);
    const uint32_t b1l = umma_desc_lo(smem_u32(smem + S1.sm_whi) + (uint32_t)S1.N * 16u, b1_plane);
    const uint32_t a1_kstep = (2u * a1_plane) >> 4, b1_kstep = (2u * b1_plane) >> 4;
    const int nks0 = S0.cin >> 4, nks1 = S1.cin >> 4;
    const int mg0 = S0.merged, mg1 = S1.merged;

    for (int k = 0; k <= nt; ++k) {
      if (k < nt) {
        // ---------------- M0(k): accumulates ON TOP of the context + bias the loader wrote ----------------
        const int b = k & 1;
        const int zb = (p.nzw == 2) ? b : 0;                 // z operand window of tile k
        const int zuse = (p.nzw == 2) ? (k >> 1) : k;        // how many times it has been filled before
        PROBE(7)
        mbar_wait(&bars[FB_ZFULL + zb], (uint32_t)(zuse & 1));
        PROBE(0)
        mbar_wait(&bars[FB_A0_INIT + b], (uint32_t)((k >> 1) & 1));
        PROBE(1)
        const uint32_t a0h = a0h_0 + (uint32_t)zb * zbuf_step, a0l = a0l_0 + (uint32_t)zb * zbuf_step;
        tc_fence_after();
        if (lane == 0) TL(0, 100, k);
        const uint32_t d = tmem_base + (uint32_t)(S0.tmem_col + b * S0.acc_cols);
        if (elect_one_sync()) {
          // (compact on purpose: fully unrolled, this issue code was 35 KB of SASS and missed the instruction cache
          //  on every burst; the loops below are a few hundred bytes)
          if (!(p.dbg & 16)) {
            uint32_t bh = b0h, bl = b0l;
            bool first = true;
#pragma unroll 1
            for (int tp = 0; tp < IAF_NTAPS; ++tp) {
              const uint32_t shv = tp < 2 ? (uint32_t)tp : (uint32_t)(p.Wp + tp - 3);  // slot shifts 0, 1, Wp-1, Wp, Wp+1
              uint32_t ah = a0h + shv, al = a0l + shv;
#pragma unroll 1
              for (int ks = 0; ks < nks0; ++ks) {
                if (!mg0 || first) {
                  umma_f16(d, mk_desc(al), mk_desc(bh), idesc0, 1u);  // lo * hi
                  umma_f16(d, mk_desc(ah), mk_desc(bh), idesc0, 1u);  // hi * hi
                  // hi * lo: own columns [N, 2N) when merged (nothing was written there: start from zero)
                  umma_f16(mg0 ? d + (uint32_t)S0.N : d, mk_desc(ah), mk_desc(bl), idesc0, mg0 ? 0u : 1u);
                } else {
                  umma_f16(d, mk_desc(ah), mk_desc(bh), idesc0m, 1u);  // hi * [hi | lo] as one N' = 2N instruction
                  umma_f16(d, mk_desc(al), mk_desc(bh), idesc0, 1u);
                }
                first = false;
                ah += a0_kstep; al += a0_kstep; bh += b0_kstep; bl += b0_kstep;
              }
            }
          }
          umma_commit(&bars[FB_A0_FULL + b]);
          umma_commit(&bars[FB_ZEMPTY + zb]);
          TL(0, 200, k);
        }
        __syncwarp();
        PROBE(2)
      }
      if (k >= 1) {
        // ---------------- M1(k-1): heads ----------------
        const int kk = k - 1, b = kk & 1;
        PROBE(7)
        if (kk == 0) mbar_wait(&bars[FB_W1], 0);
        const int hb = (p.nhb == 2) ? b : 0;               // hidden-activation buffer of tile kk
        const int huse = (p.nhb == 2) ? (kk >> 1) : kk;     // how many times it has been filled before
        mbar_wait(&bars[FB_H_FULL + hb], (uint32_t)(huse & 1));
        PROBE(3)
        if (kk >= 2) mbar_wait(&bars[FB_A1_EMPTY + b], (uint32_t)(((kk >> 1) - 1) & 1));
        PROBE(4)
        tc_fence_after();
        if (lane == 0) TL(0, 101, kk);
        const uint32_t d = tmem_base + (uint32_t)(S1.tmem_col + b * S1.acc_cols);
        const uint32_t a_base = smem_u32(smem + S1.sm_in + hb * p.h_bytes);
        const uint32_t a1h = umma_desc_lo(a_base, a1_plane);
        const uint32_t a1l = umma_desc_lo(a_base + (uint32_t)(S1.cin >> 3) * a1_plane, a1_plane);
        if (elect_one_sync()) {
          if (!(p.dbg & 16)) {
            uint32_t acc = 0;
            uint32_t bh = b1h, bl = b1l;
#pragma unroll 1
            for (int tp = 0; tp < IAF_NTAPS; ++tp) {
              const uint32_t shv = tp < 2 ? (uint32_t)tp : (uint32_t)(p.Wp + tp - 3);
              uint32_t ah = a1h + shv, al = a1l + shv;
#pragma unroll 1
              for (int ks = 0; ks < nks1; ++ks) {
                if (mg1) {
                  umma_f16(d, mk_desc(ah), mk_desc(bh), idesc1m, acc);
                  umma_f16(d, mk_desc(al), mk_desc(bh), idesc1, 1u);
                } else {
                  umma_f16(d, mk_desc(al), mk_desc(bh), idesc1, acc);
                  umma_f16(d, mk_desc(ah), mk_desc(bl), idesc1, 1u);
                  umma_f16(d, mk_desc(ah), mk_desc(bh), idesc1, 1u);
                }
                acc = 1;
                ah += a1_kstep; al += a1_kstep; bh += b1_kstep; bl += b1_kstep;
              }
            }
          }
          umma_commit(&bars[FB_A1_FULL + b]);
          umma_commit(&bars[FB_H_EMPTY + hb]);
          TL(0, 201, kk);
        }
        __syncwarp();
        PROBE(5)
      }
    }
    PROBE_DUMP(0)
  } else if (warp >= FZ_W_LD0 && warp < FZ_W_MMA) {
    // =====================================================================================
    // loader warps.  Per tile k, with every load of the tile in flight before the first wait:
    //   (1) fp32 z window -> fp16 hi / lo operand window (layer mode: the posterior sample);
    //   (2) accumulator b of stage 0 := context + bias (+ pad-channel terms)   (ar.py:402 / layers.py:163)
    // =====================================================================================
    const int lw = warp - FZ_W_LD0;
    const int lt = lw * 32 + lane;
    const int q = warp & 3, hf = lw >> 2;
    const int sl = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* tb = reinterpret_cast<const float*>(smem + S0.sm_bias);
    const int ngroups = S0.N >> 4;
    const int nch0 = S0.cin >> 3;
    const int n_items = p.WIN * nch0;
    const int plane = S0.in_slots * 16;
    const int lo_off = nch0 * plane;
    bool tables = false;
    // tile-invariant geometry of this thread's z-window items and of its context slot
    int z_dst[FZ_ZB], z_cho[FZ_ZB], z_n[FZ_ZB], z_r[FZ_ZB];
#pragma unroll
    for (int it = 0; it < FZ_ZB; ++it) {
      const int idx = it * FZ_LTHREADS + lt;
      z_dst[it] = -1; z_cho[it] = 0; z_n[it] = 0; z_r[it] = 0;
      if (idx < n_items) {
        const int ch = fast_div(idx, p.WIN, p.mg_win);
        const int zs = idx - ch * p.WIN;
        z_dst[it] = ch * plane + zs * 16;
        z_cho[it] = ch * 8 * HW;
        const int s = t0 * TO + zs;
        z_n[it] = fast_div(s, p.SPS, p.mg_sps);
        z_r[it] = s - z_n[it] * p.SPS;
      }
    }
    int c_n = fast_div(t0 * TO + sl, p.SPS, p.mg_sps);
    int c_r = t0 * TO + sl - c_n * p.SPS;
    const int CHW = p.C * HW;
    PROBE_DECL
    for (int k = 0; k < nt; ++k) {
      const int b = k & 1;
      // ---- issue every global load of the tile first: z window items, then this warp's context channel groups ----
      float v[FZ_ZB][8];
      if (!ZST) {
#pragma unroll
      for (int it = 0; it < FZ_ZB; ++it) {
        if (z_dst[it] >= 0) {
          const SlotInfo zi = decode_nr(p, z_n[it], z_r[it], HW);
          advance_nr(p, z_n[it], z_r[it], TO);
          if (zi.valid && !(p.dbg & 1)) {
            const size_t g = (size_t)zi.n * CHW + (size_t)(z_cho[it] + zi.gp);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[it][e] = __ldg(p.z + g + (size_t)e * HW);
            if (MODE == IAF_MODE_LAYER) {  // z0 = mean + exp(logsd) * eps   (tf_train.py:57, distributions.py:20)
#pragma unroll
              for (int e = 0; e < 8; ++e)
                v[it][e] = fmaf(fast_exp(__ldg(p.post_logsd + g + (size_t)e * HW)), v[it][e],
                                __ldg(p.post_mean + g + (size_t)e * HW));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[it][e] = 0.f;
          }
        }
      }
      }
      const SlotInfo si = decode_nr(p, c_n, c_r, HW);
      advance_nr(p, c_n, c_r, TO);
      const bool bx0 = (si.x == 0), bxW = (si.x == p.W - 1), byH = (si.y == p.H - 1);
      const uint32_t t_acc = t_lane + (uint32_t)(S0.tmem_col + b * S0.acc_cols);
      float cx[FZ_CXG][16];
#pragma unroll
      for (int gg = 0; gg < FZ_CXG; ++gg) {
        const int g = hf + gg * FZ_LGS;
        if (g < ngroups && si.valid && !(p.dbg & 1)) {
          const float* cp = p.ctx + ((size_t)si.n * S0.N + g * 16) * HW + si.gp;
#pragma unroll
          for (int e = 0; e < 16; ++e) cx[gg][e] = __ldg(cp + (size_t)e * HW);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) cx[gg][e] = 0.f;
        }
      }
      if (lw == 0 && lane == 0) TL(3, 41, k);
      PROBE(0)
      if (ZST) {
        // the tile's fp32 rows were staged by the TMA warp as [sample][channel][row][x]: read this thread's items
        // into registers and hand the staging buffer back BEFORE waiting for the operand window, so that the copies of
        // tile k+1 are in flight while the MMAs of tile k-1 still own the window
        const ZstGeo g = zst_geometry(p, (t0 + k) * TO);
        const int sb = (p.nzs == 2) ? (k & 1) : 0;
        const int suse = (p.nzs == 2) ? (k >> 1) : k;
        mbar_wait(&bars[FB_ZST_FULL + sb], (uint32_t)(suse & 1));
        const float* stg = reinterpret_cast<const float*>(smem + p.sm_zst + sb * p.zst_bytes);
#pragma unroll
        for (int it = 0; it < FZ_ZB; ++it) {
          if (z_dst[it] >= 0) {
            const SlotInfo zi = decode_nr(p, z_n[it], z_r[it], HW);
            advance_nr(p, z_n[it], z_r[it], TO);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[it][e] = 0.f;
            if (zi.valid && !(p.dbg & 1)) {
              const bool second = zi.n != g.n0;
              const int rows = second ? g.rows1 : g.rows0;
              const int ylo = second ? 0 : g.y0;
              const int rel = p.flip ? (ylo + rows - 1 - zi.y) : (zi.y - ylo);
              const int xx = p.flip ? (p.W - 1 - zi.x) : zi.x;
              const int cstride = rows * p.W;
              const float* sp = stg + (second ? p.C * g.rows0 * p.W : 0) + (z_cho[it] / HW) * cstride + rel * p.W + xx;
#pragma unroll
              for (int e = 0; e < 8; ++e) v[it][e] = sp[e * cstride];
            }
          }
        }
        // the loads above must have returned before the buffer is released: make the arrive depend on their values
        float dep = 0.f;
#pragma unroll
        for (int it = 0; it < FZ_ZB; ++it)
#pragma unroll
          for (int e = 0; e < 8; ++e) dep += v[it][e];
        __syncwarp();
        if (lane == 0 || dep == 1.0e38f) mbar_arrive(&bars[FB_ZST_EMPTY + sb]);
      }
      PROBE(6)
      if (!tables) {  // the bias tables travel with the weights
        mbar_wait(&bars[FB_W], 0);
        tables = true;
      }
      PROBE(1)
      // ---- (1) context + bias -> accumulator b of stage 0 (free once E0(k-2) has read it) ----
      if (k >= 2) {
        mbar_wait(&bars[FB_A0_EMPTY + b], (uint32_t)(((k >> 1) - 1) & 1));
        tc_fence_after();
      }
      PROBE(2)
#pragma unroll
      for (int gg = 0; gg < FZ_CXG; ++gg) {
        const int g = hf + gg * FZ_LGS;
        if (g < ngroups && !(p.dbg & 2)) {
          const int c0 = g * 16;
          const float4* tb4 = reinterpret_cast<const float4*>(tb + c0);
          uint32_t r[16];
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 t4 = tb4[e4];
            r[4 * e4] = __float_as_uint(cx[gg][4 * e4] + t4.x);
            r[4 * e4 + 1] = __float_as_uint(cx[gg][4 * e4 + 1] + t4.y);
            r[4 * e4 + 2] = __float_as_uint(cx[gg][4 * e4 + 2] + t4.z);
            r[4 * e4 + 3] = __float_as_uint(cx[gg][4 * e4 + 3] + t4.w);
          }
          if (PADW) {  // conv.py:77-83: the pad channel is 1 where a tap falls outside the image
            const float f1 = bxW ? 1.f : 0.f, f2 = (byH || bx0) ? 1.f : 0.f, f3 = byH ? 1.f : 0.f,
                        f4 = (byH || bxW) ? 1.f : 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e)
              r[e] = __float_as_uint(__uint_as_float(r[e]) + f1 * tb[S0.N + c0 + e] + f2 * tb[2 * S0.N + c0 + e] +
                                     f3 * tb[3 * S0.N + c0 + e] + f4 * tb[4 * S0.N + c0 + e]);
          }
          tmem_st16(t_acc + (uint32_t)c0, r);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[FB_A0_INIT + b]);
      if (lw == 0 && lane == 0) TL(3, 40, k);
      PROBE(3)
      // ---- (2) z operand window ----
      const int zb = (p.nzw == 2) ? b : 0;
      const int zuse = (p.nzw == 2) ? (k >> 1) : k;
      if (zuse >= 1) mbar_wait(&bars[FB_ZEMPTY + zb], (uint32_t)((zuse - 1) & 1));  // the M0 that read the window last
      PROBE(4)
#pragma unroll
      for (int it = 0; it < FZ_ZB; ++it) {
        if (z_dst[it] >= 0 && !(p.dbg & 2)) {
          uint8_t* dst = smem + S0.sm_in + zb * p.z_bytes + z_dst[it];
          split_store8(v[it], dst, dst + lo_off);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[FB_ZFULL + zb]);
      if (lw == 0 && lane == 0) TL(2, 30, k);
      PROBE(5)
    }
    if (lw == 0) { PROBE_DUMP(1) }
  } else if (warp < FZ_EPI) {
    // =====================================================================================
    // epilogue warps: E0(k) (hidden layer -> next operand), then E1(k-1) (heads -> outputs)
    // =====================================================================================
    const int q = warp & 3, cg = warp >> 2;
    const int sl = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* s_part = reinterpret_cast<float*>(smem + p.sm_part);
    const float* tb1 = reinterpret_cast<const float*>(smem + S1.sm_bias);
    const int ngroups0 = S0.N >> 4, ngroups1 = S1.N >> 4;
    const int plane1 = S1.in_slots * 16;
    const int lo_off1 = (S0.N >> 3) * plane1;
    PROBE_DECL
    mbar_wait(&bars[FB_W], 0);  // the heads' bias table travels with the weights
    PROBE(6)

    // this thread's slot advances by TO per tile: (sample, offset) carried incrementally; E1(k-1) reuses E0(k-1)'s decode
    int e_n = fast_div(t0 * TO + sl, p.SPS, p.mg_sps);
    int e_r = t0 * TO + sl - e_n * p.SPS;
    int t_n = fast_div(t0 * TO, p.SPS, p.mg_sps);   // same for the tile's first slot
    int t_r = t0 * TO - t_n * p.SPS;
    SlotInfo si_prev;
    si_prev.n = 0; si_prev.y = 0; si_prev.x = 0; si_prev.gp = 0; si_prev.valid = false;
    int tn_prev = 0, tr_prev = 0;
    for (int k = 0; k <= nt; ++k) {
      SlotInfo si_cur = si_prev;
      int tn_cur = tn_prev, tr_cur = tr_prev;
      if (k < nt) {
        // ---------------- E0(k) ----------------
        const int b = k & 1;
        const SlotInfo si = decode_nr(p, e_n, e_r, HW);
        advance_nr(p, e_n, e_r, TO);
        si_cur = si;
        tn_cur = t_n; tr_cur = t_r;
        advance_nr(p, t_n, t_r, TO);
        const uint32_t t_acc = t_lane + (uint32_t)(S0.tmem_col + b * S0.acc_cols);
        const int hb = (p.nhb == 2) ? b : 0;
        const int huse = (p.nhb == 2) ? (k >> 1) : k;
        uint8_t* obase = smem + S1.sm_in + hb * p.h_bytes + sl * 16;
        if (warp == 0 && lane == 0) TL(1, 9, k);
        PROBE(7)
        mbar_wait(&bars[FB_A0_FULL + b], (uint32_t)((k >> 1) & 1));
        tc_fence_after();
        PROBE(0)
        if (huse >= 1) mbar_wait(&bars[FB_H_EMPTY + hb], (uint32_t)((huse - 1) & 1));  // the M1 that read the buffer last
        PROBE(1)
        if (warp == 0 && lane == 0) TL(1, 10, k);
        for (int g = (p.dbg & 4) ? ngroups0 : cg; g < ngroups0; g += FZ_CGS) {
          const int c0 = g * 16;
          uint32_t r[16];
          tmem_ld16(t_acc + (uint32_t)c0, r);
          if (S0.merged) {  // hi*lo partial products sit in columns [N, 2N)
            uint32_t r2[16];
            tmem_ld16(t_acc + (uint32_t)(S0.N + c0), r2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float a0 = __uint_as_float(r[e]), a1 = __uint_as_float(r[e + 1]);
              add2(a0, a1, __uint_as_float(r2[e]), __uint_as_float(r2[e + 1]));
              r[e] = __float_as_uint(a0); r[e + 1] = __float_as_uint(a1);
            }
          } else {
            tmem_ld_wait();
          }
          float v[16];
          if (NLT == IAF_NL_ELU) {
            // elu(a) = max(a, exp(min(a, 0)) - 1): exp(a) - 1 >= a for a < 0, and = 0 <= a otherwise; packed fp32 pairs
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              const float a0 = __uint_as_float(r[e]), a1 = __uint_as_float(r[e + 1]);
              float x0 = fminf(a0, 0.f), x1 = fminf(a1, 0.f);
              mul2(x0, x1, 1.4426950408889634f, 1.4426950408889634f);
              x0 = ex2_approx(x0); x1 = ex2_approx(x1);
              add2(x0, x1, -1.0f, -1.0f);
              v[e] = fmaxf(a0, x0); v[e + 1] = fmaxf(a1, x1);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = tc_apply_nl<NLT>(__uint_as_float(r[e]), p.nl);
          }
          if (!si.valid) {  // pad column / zero row / past the end: this zero IS the next conv's padding
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.f;
          }
          if (S0.hid_out && si.valid && sl < TO) {  // training forward: keep the activations for iaf_step_bwd_saved
            float* hp = S0.hid_out + ((size_t)si.n * S0.N + c0) * HW + si.gp;
#pragma unroll
            for (int e = 0; e < 16; ++e) hp[(size_t)e * HW] = v[e];
          }
#pragma unroll
          for (int hch = 0; hch < 2; ++hch) {
            uint8_t* dst = obase + ((c0 >> 3) + hch) * plane1;
            split_store8(v + 8 * hch, dst, dst + lo_off1);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars[FB_A0_EMPTY + b]);
          mbar_arrive(&bars[FB_H_FULL + hb]);
        }
        if (warp == 0 && lane == 0) TL(1, 20, k);
        PROBE(2)
      }
      if (k >= 1) {
        // ---------------- E1(k-1): columns in groups of 16 = (m x 8, s x 8) of 8 channels ----------------
        const int kk = k - 1, b = kk & 1;
        const SlotInfo si = si_prev;
        const bool act = si.valid && sl < TO;  // rows >= TO belong to the next tile (they only fed this tile's halo)
        const bool bx0 = (si.x == 0), bxW = (si.x == p.W - 1), byH = (si.y == p.H - 1);
        const uint32_t t_acc = t_lane + (uint32_t)(S1.tmem_col + b * S1.acc_cols);
        constexpr int NRED = (MODE == IAF_MODE_LAYER) ? 8 : 1;
        float red[NRED];
#pragma unroll
        for (int i = 0; i < NRED; ++i) red[i] = 0.f;
        const int n_first = tn_prev;
        const int n_last = min(p.B - 1, tn_prev + fast_div(tr_prev + TO - 1, p.SPS, p.mg_sps));
        const int ns = (n_first < p.B) ? (n_last - n_first + 1) : 0;
        const int pb = kk & 1;  // partial-sum buffer
        const bool want_red = (MODE != IAF_MODE_MULTICONV) && (p.persample_out || p.bc_out);
        if (MODE == IAF_MODE_LAYER && want_red && kk >= 2)
          mbar_wait(&bars[FB_PART_EMPTY + pb], (uint32_t)(((kk >> 1) - 1) & 1));
        bool waited = false;
        if (warp == 0 && lane == 0) TL(1, 11, kk);
        for (int g = (p.dbg & 8) ? ngroups1 : cg; g < ngroups1; g += FZ_CGS) {
          const int c0 = g * 16;
          const int ch0 = g * 8;
          float zv[8];
          size_t gi = 0;
          if (act) {
            gi = ((size_t)si.n * p.C + ch0) * HW + si.gp;
            if (MODE != IAF_MODE_MULTICONV) {
#pragma unroll
              for (int e = 0; e < 8; ++e) zv[e] = __ldg(p.z + gi + (size_t)e * HW);
            }
          }
          if (!waited) {
            PROBE(7)
            mbar_wait(&bars[FB_A1_FULL + b], (uint32_t)((kk >> 1) & 1));
            tc_fence_after();
            PROBE(3)
            waited = true;
            if (warp == 0 && lane == 0) TL(1, 51, kk);
          }
          uint32_t r[16];
          tmem_ld16(t_acc + (uint32_t)c0, r);
          if (S1.merged) {
            uint32_t r2[16];
            tmem_ld16(t_acc + (uint32_t)(S1.N + c0), r2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float a0 = __uint_as_float(r[e]), a1 = __uint_as_float(r[e + 1]);
              add2(a0, a1, __uint_as_float(r2[e]), __uint_as_float(r2[e + 1]));
              r[e] = __float_as_uint(a0); r[e + 1] = __float_as_uint(a1);
            }
          } else {
            tmem_ld_wait();
          }
          if (MODE == IAF_MODE_LAYER) {
#pragma unroll
            for (int i = 0; i < NRED; ++i) red[i] = 0.f;
          }
          if (act) {
            const float4* tb4 = reinterpret_cast<const float4*>(tb1 + c0);
            const float4 bm0 = tb4[0], bm1 = tb4[1], bs0 = tb4[2], bs1 = tb4[3];
            const float bm[8] = {bm0.x, bm0.y, bm0.z, bm0.w, bm1.x, bm1.y, bm1.z, bm1.w};
            const float bs[8] = {bs0.x, bs0.y, bs0.z, bs0.w, bs1.x, bs1.y, bs1.z, bs1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float m = __uint_as_float(r[e]) + bm[e];
              float sv = __uint_as_float(r[8 + e]) + bs[e];
              if (PADW) {
                if (bxW) { m += tb1[S1.N + c0 + e]; sv += tb1[S1.N + c0 + 8 + e]; }
                if (byH || bx0) { m += tb1[2 * S1.N + c0 + e]; sv += tb1[2 * S1.N + c0 + 8 + e]; }
                if (byH) { m += tb1[3 * S1.N + c0 + e]; sv += tb1[3 * S1.N + c0 + 8 + e]; }
                if (byH || bxW) { m += tb1[4 * S1.N + c0 + e]; sv += tb1[4 * S1.N + c0 + 8 + e]; }
              }
              const size_t ge = gi + (size_t)e * HW;
              if (MODE == IAF_MODE_MULTICONV) {  // the un-fused operator: raw heads (ar.py:405-411 / layers.py:166)
                p.z_out[ge] = m;
                p.elem[ge] = sv;
                continue;
              }
              const float arw_mean = p.scale * m, arw_logsd = p.scale * sv;  // models.py:282-285
              float z0 = zv[e];
              float eps = 0.f, pls = 0.f;
              if (MODE == IAF_MODE_LAYER) {
                eps = z0;
                pls = __ldg(p.post_logsd + ge);
                z0 = fmaf(fast_exp(pls), eps, __ldg(p.post_mean + ge));
              }
              const float zn = (z0 - arw_mean) * fast_exp(-arw_logsd);
              p.z_out[ge] = zn;
              if (MODE == IAF_MODE_STEP) {
                if (p.elem) p.elem[ge] = arw_logsd;
                red[0] += arw_logsd;
              } else {
                // logqs of the pre-flow sample + arw_logsd, prior logps at z'  (tf_train.py:68-75)
                const float logqs = -0.9189385332046727f - pls - 0.5f * eps * eps + arw_logsd;
                const float pl = __ldg(p.prior_logsd + ge);
                const float dd = zn - __ldg(p.prior_mean + ge);
                const float logps = -0.9189385332046727f - pl - 0.5f * dd * dd * fast_exp(-2.0f * pl);
                const float kl = logqs - logps;
                if (p.elem) p.elem[ge] = kl;
                red[e] = kl;
              }
            }
          }
          if (MODE == IAF_MODE_LAYER && want_red) {
            // per-(sample, channel) sums over this warp's 32 slots, fixed butterfly order
            for (int nl_ = 0; nl_ < ns; ++nl_) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float x = (act && si.n == n_first + nl_) ? red[e] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                if (lane == 0) s_part[((pb * 4 + q) * p.MAXS + nl_) * p.C + ch0 + e] = x;
              }
            }
          }
        }
        if (!waited) mbar_wait(&bars[FB_A1_FULL + b], (uint32_t)((kk >> 1) & 1));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[FB_A1_EMPTY + b]);
        if (warp == 0 && lane == 0) TL(1, 21, kk);

        // deterministic per-sample reductions: fixed-order partial sums deposited in smem (double-buffered by tile
        // parity); the reducer warp folds them (see iaf_tc_kernel)
        PROBE(4)
        if (want_red) {
          if (MODE != IAF_MODE_LAYER) {
            if (kk >= 2) mbar_wait(&bars[FB_PART_EMPTY + pb], (uint32_t)(((kk >> 1) - 1) & 1));
            PROBE(5)
            for (int nl_ = 0; nl_ < ns; ++nl_) {
              float x = (act && si.n == n_first + nl_) ? red[0] : 0.f;
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
              if (lane == 0) s_part[(pb * FZ_EPI + warp) * p.MAXS + nl_] = x;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[FB_PART + pb]);
        }
      }
      si_prev = si_cur; tn_prev = tn_cur; tr_prev = tr_cur;
    }
    if (warp == 0) { PROBE_DUMP(2) }
  } else if (ZST && warp == FZ_W_TMA) {
    // =====================================================================================
    // TMA producer (staged variant): the fp32 rows of z a tile's window touches -> staging buffer, as tiled tensor
    // copies (cp.async.bulk.tensor.4d over z viewed as (x, y, channel, sample)): ONE box per sample = its rows of every
    // channel (the box height is a property of the descriptor, so the parameters carry one descriptor per row count),
    // <= 2 boxes per tile, completing on one mbarrier (expect_tx).
    // (Measured before this: per-(sample, channel) 1-D bulk copies, 64 copies of <= 640 B per tile: ~3 K cycles of TMA
    //  issue per tile, slower than gathering; one box per image ROW, 10 boxes of 32 x 64-byte segments: faster than
    //  gathering, the loader warps still waited ~2.5 K cycles per tile for them.)
    // =====================================================================================
    const uint32_t row_bytes = (uint32_t)(p.C * p.W * 4);
    // (L2 prefetch of the rows of the tiles further ahead -- cp.async.bulk.prefetch.tensor for z and for the context --
    //  was measured: 26.8 us against 25.0 us without; the extra descriptor-based requests delay the copies themselves.)
    for (int k = 0; k < nt; ++k) {
      const ZstGeo g = zst_geometry(p, (t0 + k) * TO);
      const int sb = (p.nzs == 2) ? (k & 1) : 0;
      const int suse = (p.nzs == 2) ? (k >> 1) : k;
      uint8_t* stg = smem + p.sm_zst + sb * p.zst_bytes;
      if (suse >= 1) mbar_wait(&bars[FB_ZST_EMPTY + sb], (uint32_t)((suse - 1) & 1));
      const int nrows = g.rows0 + g.rows1;
      if (lane == 0) mbar_expect_tx(&bars[FB_ZST_FULL + sb], (p.dbg & 1) ? 0u : (uint32_t)nrows * row_bytes);
      __syncwarp();
      if (!(p.dbg & 1) && lane < 2) {
        // lane 0 fetches sample n0's rows, lane 1 sample n1's: ONE box each (x 0..W-1, `rows` image rows, every channel),
        // taken from the descriptor whose box has exactly that many rows.  First staged memory row of each sample (Theano
        // orientation: the stream is the point-reflected image):
        const int m0 = p.flip ? p.H - g.y0 - g.rows0 : g.y0;
        const int m1 = p.flip ? p.H - g.rows1 : 0;
        const int rws = lane ? g.rows1 : g.rows0;
        if (rws > 0)
          tma_load_4d(stg + (lane ? (size_t)g.rows0 * row_bytes : 0), p.tmap_z[rws - 1], 0, lane ? m1 : m0, 0, lane ? g.n1 : g.n0,
                      &bars[FB_ZST_FULL + sb]);
      }
      __syncwarp();
    }
  } else if (warp == FZ_W_RED && (p.persample_out || p.bc_out) && MODE != IAF_MODE_MULTICONV) {
    // =====================================================================================
    // reducer warp: per-tile partials -> per-sample outputs (a sample's last tile sums all of its tiles in fixed order)
    // =====================================================================================
    float* s_part = reinterpret_cast<float*>(smem + p.sm_part);
    constexpr bool LAY = (MODE == IAF_MODE_LAYER);
    for (int k = 0; k < nt; ++k) {
      const int u = t0 + k;
      const int pb = k & 1;
      const int tile_s0 = u * TO;
      const int n_first = fast_div(tile_s0, p.SPS, p.mg_sps);
      const int n_last = min(p.B - 1, fast_div(tile_s0 + TO - 1, p.SPS, p.mg_sps));
      const int ns = (tile_s0 < p.S) ? (n_last - n_first + 1) : 0;
      mbar_wait(&bars[FB_PART + pb], (uint32_t)((k >> 1) & 1));
      const int cred = LAY ? p.C : 1;
      // fold the workers' partials (fixed order) into registers, hand the scratch buffer back at once -- the global
      // round trips below (fence, counter atomic, possibly the final sum) then overlap the next tiles' epilogues
      constexpr int RMAX = LAY ? 8 : 1;  // values per lane: ns * cred <= 32 * RMAX is checked by the host layout
      float tot[RMAX];
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        const int i = lane + 32 * j;
        tot[j] = 0.f;
        if (i < ns * cred) {
          if (LAY) {
            const int nl_ = i / p.C, c = i - nl_ * p.C;
            for (int qq = 0; qq < 4; ++qq) tot[j] += s_part[((pb * 4 + qq) * p.MAXS + nl_) * p.C + c];
          } else {
            for (int w = 0; w < FZ_EPI; ++w) tot[j] += s_part[(pb * FZ_EPI + w) * p.MAXS + i];
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[FB_PART_EMPTY + pb]);
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        const int i = lane + 32 * j;
        if (i < ns * cred) p.tilepart[((size_t)u * p.MAXS) * cred + i] = tot[j];
      }
      // Publish, count, and (last tile of a sample) sum.  Step / multiconv-free modes: lane i stores sample i's partial AND
      // bumps sample i's counter, so one acq_rel atomic orders both directions and the two device-wide fences -- each a
      // round trip of the kernel's tail -- are not needed.  Layer mode: a sample's partials come from 32 lanes: fences.
#ifdef IAF_FZ_FENCE_TAIL
      constexpr bool FENCE = true;   // A/B switch: the round-2 mid-state tail (fence, atomic, fence)
#else
      constexpr bool FENCE = LAY;
#endif
      if (FENCE) {
        __threadfence();
        __syncwarp();
      }
      for (int i = lane; i < ns; i += 32) {
        const int n = n_first + i;
        const int a = n * p.SPS, bb = a + p.SPS - 1;
        const int ta = a / TO, tbk = bb / TO;
        const unsigned expected = (unsigned)(tbk - ta + 1);
        unsigned prev;
        if (FENCE) {
          prev = atomicAdd(p.counter + n, 1u);
        } else {
          asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(p.counter + n) : "memory");
        }
        if (prev == expected - 1u) {
          if (FENCE) __threadfence();
          p.counter[n] = 0u;  // ready for the next launch
          float cost = 0.f;
          for (int c = 0; c < cred; ++c) {
            float t = 0.f;
            for (int tt = ta; tt <= tbk; ++tt) {
              const int nf = fast_div(tt * TO, p.SPS, p.mg_sps);
              t += __ldcg(p.tilepart + ((size_t)tt * p.MAXS + (n - nf)) * cred + c);
            }
            if (LAY && p.bc_out) p.bc_out[(size_t)n * p.C + c] = t;
            cost += t;
          }
          if (p.persample_out) p.persample_out[n] = LAY ? cost : -cost;  // logdet = -sum(arw_logsd)
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  TL_FLUSH
  if (warp == FZ_W_MMA) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}
