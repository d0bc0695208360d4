// Weight gradient of one masked conv layer on the tensor cores (included by iaf_tc.cu).
//
//   dW[tap][ci][co] = sum over slots r of  X[r][ci] * G[r + shift_tap][co]          (point-reflected stream)
//
// X (the layer's input activations) and G (the gradient at its pre-activation output) are the SAME operand images the
// forward and the data gradient use -- [chunk of 8 channels][slot][8] fp16 hi / lo -- but read with the SLOT stream as
// K: 8 consecutive slots of a chunk plane are exactly one core matrix of the canonical no-swizzle MN-MAJOR layout
// (8 K-rows x 16 bytes of 8 contiguous channels), LBO = 128 B between K groups, SBO = the plane pitch between channel
// groups, and a tap is still `shift x 16` bytes on G's start address (tools/mma_mnmajor.cu, profiles/r2_mma_mnmajor.log:
// exact for shifts 0 / 1 / 8 / 17 / 18).  So per tap and per 16 slots: D_tap[ci][co] += X^T G with M = 128 input channels
// (a block; M = 64 for a block of at most 64; rows past the layer's channels read whatever follows in shared memory and
// are never stored), N = a part of
// the columns (5 accumulators of N <= 96 columns fill the 512 TMEM columns), the same three split-operand products as
// everywhere (X_lo G_hi, X_hi G_lo, X_hi G_hi; A-operand collector on the pair).
//
// Work split: one CTA per (channel block, column part, split-K group); the group walks K tiles of WG_KT slots through a
// bulk-copy ring; partial sums go to part[group][...] and the fixed-order reduction kernel of iaf_bwd.cu adds them.
// The images carry per-sample scales (s_n on G, c / s_n on X, c = the smallest s_n; see iaf_dg_image_kernel), so the
// product carries the single factor c, removed here.
#pragma once

#define WG_KT 64          // slots per K tile (4 MMAs of K = 16 per tap and product)
#define WG_HALO 24        // G slots past the tile a tap can reach (>= Wp + 1, multiple of 8)
#define WG_MAX_STAGES 6
#define WG_EPI 4          // epilogue warps (one per TMEM lane quadrant)
#define WG_W_MMA WG_EPI
#define WG_W_TMA (WG_EPI + 1)
#define WG_THREADS ((WG_EPI + 2) * 32)

struct IafWgTcParams {
  const __nv_bfloat16* x_hi;  // [x_planes/8][S_pad][8]
  const __nv_bfloat16* x_lo;
  const __nv_bfloat16* g_hi;  // [g_planes/8][S_pad][8]
  const __nv_bfloat16* g_lo;
  float* part;                // [NG][5 * cin * ncol (+ 5 * ncol unused here)]
  const float* amax;          // [B]: c = scale of the largest
  int B, cin, ncol, S_pad, Wp;
  int n_mb, n_np, Np;         // channel blocks of 128, column parts of Np
  int NTK, NG;                // K tiles in all, split-K groups
  int part_stride;            // floats per group in `part`
  int n_stages, stage_bytes, xa_bytes;  // ring: per stage [X hi planes][X lo planes][G hi planes][G lo planes]
  int xplanes, gplanes;       // chunk planes staged per tile: of this CTA's channel block / column part
};

__global__ void __launch_bounds__(WG_THREADS, 1) iaf_wg_kernel(const __grid_constant__ IafWgTcParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * WG_MAX_STAGES + 1];
  __shared__ uint32_t s_tmem;
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_MAX_STAGES;
  uint64_t* acc_full = bars + 2 * WG_MAX_STAGES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int bid = blockIdx.x;
  const int np = bid % p.n_np; bid /= p.n_np;
  const int mb = bid % p.n_mb;
  const int g = bid / p.n_mb;
  const int x_pitch = WG_KT * 16, g_pitch = (WG_KT + WG_HALO) * 16;  // bytes per chunk plane in a stage
  const int n_my = (p.NTK - g + p.NG - 1) / p.NG;                      // K tiles u = g, g + NG, ...
  const int xpl = min(p.xplanes, (p.cin >> 3) - mb * 16);               // chunk planes this channel block really has
  // a block with at most 64 channels issues M = 64 instructions (half the A fetch; the accumulator then sits in lanes
  // 0-15 of every 32-lane quadrant: row m -> lane (m / 16) * 32 + m % 16, tools/mma_mnmajor.cu)
  const bool m64 = xpl <= 8;

  if (warp == WG_W_MMA) {
    tmem_alloc(&s_tmem, 512u);
    if (lane == 0) {
      for (int i = 0; i < WG_MAX_STAGES; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == WG_W_TMA) {
    // one bulk copy per chunk plane and image; the lanes of the warp issue them side by side
    const uint32_t tx = (uint32_t)(2 * xpl * x_pitch + 2 * p.gplanes * g_pitch);
    const int ncp = 2 * xpl + 2 * p.gplanes;
    for (int i = 0; i < n_my; ++i) {
      const int u = g + i * p.NG;
      const int stg = i % p.n_stages, use = i / p.n_stages;
      if (use >= 1) mbar_wait(&empty[stg], (uint32_t)((use - 1) & 1));
      uint8_t* dst = smem + (size_t)stg * p.stage_bytes;
      if (lane == 0) mbar_expect_tx(&full[stg], tx);
      __syncwarp();
      const size_t s0 = (size_t)u * WG_KT;
      for (int k = lane; k < ncp; k += 32) {
        if (k < 2 * xpl) {
          const int lo = k >= xpl, c = lo ? k - xpl : k;
          const size_t go = ((size_t)(mb * 16 + c) * p.S_pad + s0) * 8;
          bulk_g2s(dst + ((lo ? p.xplanes : 0) + c) * x_pitch, (lo ? p.x_lo : p.x_hi) + go, (uint32_t)x_pitch, &full[stg]);
        } else {
          const int kk = k - 2 * xpl;
          const int lo = kk >= p.gplanes, c = lo ? kk - p.gplanes : kk;
          const size_t go = ((size_t)(np * (p.Np >> 3) + c) * p.S_pad + s0) * 8;
          bulk_g2s(dst + p.xa_bytes + ((lo ? p.gplanes : 0) + c) * g_pitch, (lo ? p.g_lo : p.g_hi) + go, (uint32_t)g_pitch,
                   &full[stg]);
        }
      }
      __syncwarp();
    }
  } else if (warp == WG_W_MMA) {
    // instruction descriptor: fp16 x fp16 -> f32, BOTH operands MN-major (bits 15, 16), M = 128, N = Np
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)((m64 ? 64 : 128) >> 4) << 24);
    const uint32_t xh_hi = ((uint32_t)x_pitch >> 4) | (1u << 14);  // high words: SBO = plane pitch, descriptor version 1
    const uint32_t gh_hi = ((uint32_t)g_pitch >> 4) | (1u << 14);
    const uint32_t sh[IAF_NTAPS] = {0u, 1u, (uint32_t)(p.Wp - 1), (uint32_t)p.Wp, (uint32_t)(p.Wp + 1)};
    for (int i = 0; i < n_my; ++i) {
      const int stg = i % p.n_stages, use = i / p.n_stages;
      mbar_wait(&full[stg], (uint32_t)(use & 1));
      tc_fence_after();
      const uint32_t sbase = smem_u32(smem + (size_t)stg * p.stage_bytes);
      // low words: start address (16-byte units) | LBO = 128 B (next 8 slots) << 16
      const uint32_t xh0 = ((sbase >> 4) & 0x3FFFu) | ((128u >> 4) << 16);
      const uint32_t xl0 = (((sbase + (uint32_t)(p.xplanes * x_pitch)) >> 4) & 0x3FFFu) | ((128u >> 4) << 16);
      const uint32_t gh0 = (((sbase + (uint32_t)p.xa_bytes) >> 4) & 0x3FFFu) | ((128u >> 4) << 16);
      const uint32_t gl0 = (((sbase + (uint32_t)p.xa_bytes + (uint32_t)(p.gplanes * g_pitch)) >> 4) & 0x3FFFu) | ((128u >> 4) << 16);
      if (elect_one_sync()) {
#pragma unroll 1
        for (int t = 0; t < IAF_NTAPS; ++t) {
          const uint32_t d = tmem_base + (uint32_t)(t * p.Np);
#pragma unroll 1
          for (int ks = 0; ks < WG_KT / 16; ++ks) {
            const uint32_t xo = (uint32_t)(ks * 16), go = (uint32_t)(ks * 16) + sh[t];  // 16-byte units = slots
            const uint64_t xh = ((uint64_t)xh_hi << 32) | (xh0 + xo), xl = ((uint64_t)xh_hi << 32) | (xl0 + xo);
            const uint64_t gh = ((uint64_t)gh_hi << 32) | (gh0 + go), gl = ((uint64_t)gh_hi << 32) | (gl0 + go);
            const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
            umma_f16(d, xl, gh, idesc, acc);
            umma_f16_afill(d, xh, gl, idesc, 1u);
            umma_f16_alast(d, xh, gh, idesc, 1u);
          }
        }
        umma_commit(&empty[stg]);
        if (i == n_my - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps: D_tap[ci][co] (lanes = input channels of this block) -> part[g][(tap * cin + ci) * ncol + co] / c
    const int ci = m64 ? (lane < 16 ? mb * 128 + warp * 16 + lane : p.cin) : mb * 128 + warp * 32 + lane;
    float* out = p.part + (size_t)g * p.part_stride;
    if (n_my > 0) {
      float am = 0.f;
      for (int n = 0; n < p.B; ++n) am = fmaxf(am, __ldg(p.amax + n));
      const float inv_c = 1.0f / dg_scale_from_amax(am);
      mbar_wait(acc_full, 0u);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int t = 0; t < IAF_NTAPS; ++t)
        for (int c0 = 0; c0 < p.Np; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(t_lane + (uint32_t)(t * p.Np + c0), r);
          tmem_ld_wait();
          if (ci < p.cin) {
            float* o = out + ((size_t)t * p.cin + ci) * p.ncol + np * p.Np + c0;
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4)
              *reinterpret_cast<float4*>(o + 4 * e4) =
                  make_float4(__uint_as_float(r[4 * e4]) * inv_c, __uint_as_float(r[4 * e4 + 1]) * inv_c,
                              __uint_as_float(r[4 * e4 + 2]) * inv_c, __uint_as_float(r[4 * e4 + 3]) * inv_c);
          }
        }
    } else if (ci < p.cin) {
      for (int t = 0; t < IAF_NTAPS; ++t)
        for (int c0 = 0; c0 < p.Np; ++c0) out[((size_t)t * p.cin + ci) * p.ncol + np * p.Np + c0] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_W_MMA) tmem_dealloc(tmem_base, 512u);
}
