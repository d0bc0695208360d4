// Development microbenchmark (not part of the library): how fast does tcgen05.mma (kind::f16, M = 128, cta_group::1)
// run from shared-memory operands in the layouts the IAF kernels use or could use?
//   layout 0: SWIZZLE_NONE K-major, rows at 16-byte pitch inside a [K chunk][slot] plane -- a tap shift is +16 B on the
//             start address, so every shift that is not a multiple of 8 slots leaves the 8-row core matrices straddling
//             two 128-byte lines (what iaf_tc_kernel / iaf_ly_kernel do today);
//   layout 1: SWIZZLE_128B K-major, one 128-byte row (64 bf16) per slot -- a tap shift is +128 B.
// It also checks NUMERICALLY that a row-shifted SWIZZLE_128B operand (start address not 1024-byte aligned) gives the
// right product, with the descriptor's base-offset field set to 0 and to (addr >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_bench tools/mma_bench.cu && ./mma_bench
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t idesc_for(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// descriptor: start address, LBO, SBO (bytes), layout type, base offset
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14) | ((base_off & 7u) << 17) | (layout << 29);
  return ((uint64_t)hi << 32) | lo;
}

struct Cfg {
  int layout;   // 0 none, 1 sw128
  int N;        // MMA N
  int shift;    // A row shift (slots)
  int nmma;     // MMAs per timed burst
  int pair;     // 1: alternate N and N/2 (the merged hi*[hi|lo] + lo*hi pattern)
  int bo_mode;  // sw128: 0 base offset 0, 1 base offset (addr >> 7) & 7
  int plane;    // pair == 3: A plane pitch in bytes (0 = 4480)
  int nks;      // pair == 3: K-steps per tap (0 = 4)
  int unm;      // pair == 3: 1 = the first K-step of a burst is issued un-merged (3 N/2-wide MMAs), as iaf_fz_kernel's M0
  int dcol;     // pair == 3: accumulator column offset
};

// timing kernel: one CTA per SM; warp 0 walks the issue loop convergently and ONE elected lane issues (descriptors then
// live in uniform registers: the same issue pattern as the library kernels); warps 1..3 optionally hammer shared memory
// with 16-byte stores (bg = 1) or loads (bg = 2) in a disjoint region to expose interference with the operand fetch.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
__device__ __forceinline__ uint64_t desc_from(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// bg: 0 none, 1 st.shared.v4 stream, 2 ld.shared.v4 stream, 3 ld.global.nc (L1-allocating) stream over an L2-resident
// buffer, 4 the same with L1::no_allocate, 5 tcgen05.ld stream, 6 tcgen05.st stream; issued by warps 4..15 (384 threads).
// c.pair == 2: the library's real stage-1 pattern: 5 taps (A shifts 0,1,Wp-1,Wp,Wp+1) x 4 K-steps x (N'=128, N=64).
__global__ void __launch_bounds__(512, 1) k_time(Cfg c, long long* out, int bg, const float* gbuf) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  __shared__ volatile int s_stop;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 216 * 1024 / 4; i += 512) {
    uint32_t x = 0x3c003c00u + (i & 0xff);
    if (c.bo_mode) {  // random bf16 pairs in roughly [-2, 2]: sign random, exponent 0x3c..0x3f, mantissa random
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      x = (h & 0x83ff83ffu) | 0x3c003c00u | ((h >> 3) & 0x03000300u);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = x;
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    s_stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t d = s_tmem;
  const uint32_t a_base = smem_u32(smem);                 // A region: 128 KB
  const uint32_t b_base = smem_u32(smem) + 128 * 1024;    // B region: 64 KB; background traffic region: 192..216 KB
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    uint32_t a_lo0, a_hi, a_step, b_lo0, b_step;
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t plane = 280 * 16;
    if (c.layout == 0) {
      a_lo0 = (((a_base + c.shift * 16) >> 4) & 0x3FFFu) | (((plane >> 4) & 0x3FFFu) << 16);
      a_hi = (128u >> 4) | (1u << 14);
      a_step = (2 * plane) >> 4;
    } else {
      a_lo0 = (((a_base + c.shift * 128) >> 4) & 0x3FFFu) | (1u << 16);
      a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      a_step = 32 >> 4;
    }
    const uint32_t bplane = (uint32_t)c.N * 16;
    b_lo0 = ((b_base >> 4) & 0x3FFFu) | (((bplane >> 4) & 0x3FFFu) << 16);
    b_step = (2 * bplane) >> 4;
    const uint32_t idN = idesc_for(c.N), idH = idesc_for(c.N / 2);
    const uint32_t sh[5] = {0u, 1u, 16u, 17u, 18u};
    uint32_t par = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (c.pair == 3) {
        // bursts (5 taps x nks K-steps x (N'=N, N/2)), each followed by commit + wait, like the library's M phases
        const uint32_t pl = c.plane ? (uint32_t)c.plane : plane;
        const int nks = c.nks ? c.nks : 4;
        const uint32_t pa_lo0 = (((a_base) >> 4) & 0x3FFFu) | (((pl >> 4) & 0x3FFFu) << 16);
        const uint32_t pa_step = (2 * pl) >> 4;
        const uint32_t dd = d + (uint32_t)c.dcol;
        const uint32_t b_lo_half = b_lo0 + (((uint32_t)c.N / 2 * 16) >> 4);  // the lo rows of the [hi | lo] image
        for (int done = 0; done < c.nmma; done += 10 * nks) {
          if (elect_one()) {
            uint32_t acc = 0;
            const uint32_t al0 = pa_lo0 + ((2 * nks * pl) >> 4);
            uint32_t bh = b_lo0, bl = b_lo_half;
            bool first = c.unm != 0;
#pragma unroll
            for (int tp = 0; tp < 5; ++tp) {
              uint32_t ah = pa_lo0 + sh[tp], al = al0 + sh[tp];
              for (int ks = 0; ks < nks; ++ks) {
                if (first) {
                  umma(dd, desc_from(al, a_hi), desc_from(bh, b_hi), idH, 1u);
                  umma(dd, desc_from(ah, a_hi), desc_from(bh, b_hi), idH, 1u);
                  umma(dd + (uint32_t)c.N / 2, desc_from(ah, a_hi), desc_from(bl, b_hi), idH, 0u);
                } else {
                  umma(dd, desc_from(ah, a_hi), desc_from(bh, b_hi), idN, acc);
                  umma(dd, desc_from(al, a_hi), desc_from(bh, b_hi), idH, 1u);
                }
                first = false;
                acc = 1;
                ah += pa_step; al += pa_step; bh += b_step; bl += b_step;
              }
            }
            commit(&bar);
          }
          __syncwarp();
          mbar_wait(&bar, par);
          par ^= 1;
        }
        t1 = clock64();
        continue;
      }
      if (elect_one()) {
        uint32_t acc = 0;
        if (c.pair == 2) {
          // real stage-1 pattern (layout 0): hi planes at a_base, lo planes 8 planes further; weights image [K/8][2N][8]
          const uint32_t al0 = a_lo0 + ((8 * plane) >> 4);
          for (int burst = 0; burst < c.nmma / 40; ++burst) {
            uint32_t bh = b_lo0;
#pragma unroll
            for (int tp = 0; tp < 5; ++tp) {
              uint32_t ah = a_lo0 + sh[tp], al = al0 + sh[tp];
              for (int ks = 0; ks < 4; ++ks) {
                umma(d, desc_from(ah, a_hi), desc_from(bh, b_hi), idN, acc);
                umma(d, desc_from(al, a_hi), desc_from(bh, b_hi), idH, 1u);
                acc = 1;
                ah += a_step; al += a_step; bh += b_step;
              }
            }
          }
        } else {
          for (int i = 0; i < c.nmma; i += 4) {
            uint32_t al = a_lo0, bl = b_lo0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              umma(d, desc_from(al, a_hi), desc_from(bl, b_hi), (c.pair && (j & 1)) ? idH : idN, acc);
              acc = 1;
              al += a_step; bl += b_step;
            }
          }
        }
        commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, par);
      par ^= 1;
      t1 = clock64();
    }
    s_stop = 1;
  } else if (bg == 7 && warp >= 1) {
    __shared__ __align__(8) uint64_t idle_bar;
    if (tid == 32) mbar_init(&idle_bar, 1);
    __syncwarp();
    while (!s_stop) {
      uint32_t ok;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&idle_bar)), "r"(0u) : "memory");
      if (ok) break;
    }
  } else if (bg && warp >= 4) {
    const int bt = tid - 128;  // 0..383
    uint8_t* reg = smem + 192 * 1024 + (bt & 127) * 16;
    uint4 v = make_uint4(tid, 1, 2, 3);
    const float* gp = gbuf + (size_t)blockIdx.x * 65536 + bt;
    uint32_t r[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) r[e] = tid + e;
    const uint32_t taddr = d + ((uint32_t)((warp & 3) * 32) << 16) + 256;  // columns 256.. (the MMAs use 0..255)
    int it = 0;
    float accf = 0.f;
    while (!s_stop) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint8_t* q = reg + ((it + u) & 7) * 2048;
        if (bg == 1) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(q)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        else if (bg == 2) { uint4 w; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(smem_u32(q)) : "memory"); v.x ^= w.x; }
        else if (bg == 3) { float w; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(w) : "l"(gp + (size_t)((it + u) & 127) * 512)); accf += w; }
        else if (bg == 4) { float w; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(w) : "l"(gp + (size_t)((it + u) & 127) * 512)); accf += w; }
        else if (bg == 5) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                         "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                       : "r"(taddr + (uint32_t)(u * 16)) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else if (bg == 6) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + (uint32_t)(u * 16)),
                       "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                       "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
      it += 8;
    }
    if (v.x == 0x12345 || accf == 1.2345f || r[3] == 0x7654321) out[1] = v.x;
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(512) : "memory");
}

// correctness kernel: A[256 slots][64] bf16 (SW128 rows), B[64 n][64 k] (no-swizzle image [K/8][N][8]); D = A[s..s+128) * B^T
__global__ void __launch_bounds__(128, 1) k_check(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int shift, int bo_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x;
  // A: row r at r*128, 16-byte chunk c stored at position c ^ (r & 7)
  for (int i = tid; i < 256 * 8; i += 128) {
    const int r = i >> 3, ch = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * 64 + ch * 8);
    *reinterpret_cast<uint4*>(smem + r * 128 + ((ch ^ (r & 7)) << 4)) = v;
  }
  // B image [K/8][N][8]
  for (int i = tid; i < 8 * 64; i += 128) {
    const int kc = i / 64, n = i % 64;
    const uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)n * 64 + kc * 8);
    *reinterpret_cast<uint4*>(smem + 64 * 1024 + (kc * 64 + n) * 16) = v;
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t d = s_tmem;
  if (tid == 0) {
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t addr = smem_u32(smem) + shift * 128 + ks * 32;
      const uint64_t ad = mk_desc(addr, 16, 1024, 2, bo_mode ? ((addr >> 7) & 7) : 0);
      const uint64_t bd = mk_desc(smem_u32(smem) + 64 * 1024 + ks * 2 * 64 * 16, 64 * 16, 128, 0, 0);
      umma(d, ad, bd, idesc_for(64), ks ? 1u : 0u);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = tid >> 5;
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(d + ((uint32_t)(warp * 32) << 16) + c0)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 16; ++e) D[(size_t)tid * 64 + c0 + e] = __uint_as_float(r[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(64) : "memory");
}

int main() {
  cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
  cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  // ---- correctness of row-shifted SWIZZLE_128B operands ----
  std::vector<__nv_bfloat16> hA(256 * 64), hB(64 * 64);
  std::vector<float> fA(256 * 64), fB(64 * 64);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 17) - 8); hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 9) - 4); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift : {0, 1, 17}) {
      cudaMemset(dD, 0, 128 * 64 * 4);
      k_check<<<1, 128, 96 * 1024>>>(dA, dB, dD, shift, bo);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> hD(128 * 64);
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 64; ++n) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += fA[(size_t)(i + shift) * 64 + k] * fB[(size_t)n * 64 + k];
          if (ref != hD[(size_t)i * 64 + n]) ++bad;
        }
      printf("check sw128 shift=%2d base_offset_mode=%d : %s (%d mismatches) %s\n", shift, bo, bad ? "WRONG" : "ok", bad,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  // ---- timing ----
  long long* dout;
  cudaMalloc(&dout, 16);
  float* gbuf;
  cudaMalloc(&gbuf, (size_t)148 * 65536 * 4 + 4096);
  cudaMemset(gbuf, 0, (size_t)148 * 65536 * 4 + 4096);
  std::vector<Cfg> rows;
  // {layout, N, shift, nmma, pair, data, plane, nks, unm, dcol}
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 4480, 4, 0, 0});    // stage 1, gen1 ring pitch
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2048, 4, 0, 0});    // stage 1, fz h buffer pitch
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2048, 4, 0, 256});  // ... accumulator at column 256
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2176, 4, 0, 0});    // ... pitch 2048 + 128
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2080, 4, 0, 0});    // ... pitch 2048 + 32
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2336, 2, 0, 0});    // stage 0 (2 K-steps per tap), fz z window pitch, all merged
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2336, 2, 1, 0});    // ... first K-step un-merged (what M0 issues)
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2432, 2, 0, 0});    // stage 0, gen1 z window pitch
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 4480, 2, 0, 0});    // stage 0 with the ring pitch
  rows.push_back(Cfg{0, 128, 0, 160, 3, 1, 2048, 2, 0, 0});
  for (int bg : {0})
    for (auto& c : rows) {
      k_time<<<148, 512, 216 * 1024>>>(c, dout, bg, gbuf);
      cudaError_t e = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
      const int per = 10 * c.nks, nb = (c.nmma + per - 1) / per;
      printf("time plane=%4d nks=%d unmerged_first=%d dcol=%3d : %7.1f cycles per burst of %d K-steps (%d bursts) %s\n", c.plane, c.nks,
             c.unm, c.dcol, (double)cyc / nb, 5 * c.nks, nb, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
