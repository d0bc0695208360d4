// Development microbenchmark (not part of the library): does the A-operand COLLECTOR of tcgen05.mma (PTX qualifiers
// .collector::a::fill / ::use / ::lastuse) save the shared-memory fetch of A when consecutive MMAs read the same A tile?
// The split-operand scheme issues (A_lo x B_hi), (A_hi x B_hi), (A_hi x B_lo) per K-step: the third could take A_hi from
// the collector.  Measured model so far (tools/mma_bench.cu): cycles per M=128 K=16 MMA = max(tensor, (A + B bytes) / 128).
//   part 1 (numerics): one triple with small-integer fp16 data, each usage pattern, compared exactly with the host; plus
//           a "liar" pattern whose ::use instruction names a DIFFERENT A descriptor -- the result tells whether the
//           hardware really took A from the collector;
//   part 2 (timing): bursts of triples (commit + wait per burst) for N = 32 / 64 / 128 / 160, patterns default / fill+lastuse.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_collector tools/mma_collector.cu
#include <cuda_fp16.h>
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
#define UMMA_VARIANT(name, qual)                                                                                              \
  __device__ __forceinline__ void name(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {                    \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16" qual                      \
                 " [%0], %1, %2, %3, p;\n\t}" ::"r"(d),                                                                        \
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)                                                                          \
                 : "memory");                                                                                                  \
  }
UMMA_VARIANT(umma_plain, "")
UMMA_VARIANT(umma_fill, ".collector::a::fill")
UMMA_VARIANT(umma_use, ".collector::a::use")
UMMA_VARIANT(umma_lastuse, ".collector::a::lastuse")
UMMA_VARIANT(umma_discard, ".collector::a::discard")

__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t idesc_for(int N) {  // fp16 A and B, fp32 accumulate, K-major both, M = 128
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {  // SWIZZLE_NONE
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

// operand images (no swizzle, K-major): [K chunk of 8][row][8 halves]; rows = 160 for A (128 + room for shifts), N for B
constexpr int A_ROWS = 160;
constexpr int KS_MAX = 8;   // K-steps (of 16) resident (N = 160: 4 * 8 * (160 + 160) * 16 B = 164 KB)
__host__ __device__ inline int a_val_hi(int r, int k) { return (r * 3 + k) % 7 - 3; }
__host__ __device__ inline int a_val_lo(int r, int k) { return (r + 2 * k) % 5 - 2; }
__host__ __device__ inline int b_val_hi(int n, int k) { return (n + k) % 9 - 4; }
__host__ __device__ inline int b_val_lo(int n, int k) { return (2 * n + 3 * k) % 5 - 2; }

struct Cfg {
  int N;
  int pattern;  // 0 plain x3; 1 plain, fill, lastuse; 2 discard, fill, lastuse; 3 liar: plain, fill(A_hi), lastuse(descriptor of A_lo)
                // 4: fill(A_hi x B_hi), use(A_hi x B_lo), then plain(A_lo x B_hi)  [reordered triple]
  int ksteps;   // K-steps per burst (each 3 MMAs)
  int bursts;
  int shift;    // A row shift of every second K-step (tap shifts: the A descriptor changes every K-step anyway)
};

__global__ void __launch_bounds__(128, 1) k_col(Cfg c, long long* cyc_out, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = c.N;
  const uint32_t a_plane = A_ROWS * 16, b_plane = (uint32_t)N * 16;
  uint8_t* Ah = smem;
  uint8_t* Al = Ah + 2 * KS_MAX * a_plane;
  uint8_t* Bh = Al + 2 * KS_MAX * a_plane;
  uint8_t* Bl = Bh + 2 * KS_MAX * b_plane;
  for (int i = tid; i < 2 * KS_MAX * A_ROWS * 8; i += 128) {
    const int e = i & 7, r = (i >> 3) % A_ROWS, ch = (i >> 3) / A_ROWS;
    const int k = ch * 8 + e;
    reinterpret_cast<__half*>(Ah)[i] = __int2half_rn(a_val_hi(r, k));
    reinterpret_cast<__half*>(Al)[i] = __int2half_rn(a_val_lo(r, k));
  }
  for (int i = tid; i < 2 * KS_MAX * N * 8; i += 128) {
    const int e = i & 7, n = (i >> 3) % N, ch = (i >> 3) / N;
    const int k = ch * 8 + e;
    reinterpret_cast<__half*>(Bh)[i] = __int2half_rn(b_val_hi(n, k));
    reinterpret_cast<__half*>(Bl)[i] = __int2half_rn(b_val_lo(n, k));
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
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
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    const uint32_t id = idesc_for(N);
    uint32_t par = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      for (int b = 0; b < c.bursts; ++b) {
        if (elect_one()) {
          uint32_t acc = 0;
#pragma unroll 1
          for (int ks = 0; ks < c.ksteps; ++ks) {
            const int kk = ks % KS_MAX;
            const uint32_t sh = (ks & 1) ? (uint32_t)c.shift * 16 : 0u;
            const uint64_t ah = mk_desc(smem_u32(Ah) + 2 * kk * a_plane + sh, a_plane, 128);
            const uint64_t al = mk_desc(smem_u32(Al) + 2 * kk * a_plane + sh, a_plane, 128);
            const uint64_t bh = mk_desc(smem_u32(Bh) + 2 * kk * b_plane, b_plane, 128);
            const uint64_t bl = mk_desc(smem_u32(Bl) + 2 * kk * b_plane, b_plane, 128);
            switch (c.pattern) {
              case 0:
                umma_plain(d, al, bh, id, acc); umma_plain(d, ah, bh, id, 1u); umma_plain(d, ah, bl, id, 1u); break;
              case 1:
                umma_plain(d, al, bh, id, acc); umma_fill(d, ah, bh, id, 1u); umma_lastuse(d, ah, bl, id, 1u); break;
              case 2:
                umma_discard(d, al, bh, id, acc); umma_fill(d, ah, bh, id, 1u); umma_lastuse(d, ah, bl, id, 1u); break;
              case 3:
                umma_plain(d, al, bh, id, acc); umma_fill(d, ah, bh, id, 1u); umma_lastuse(d, al, bl, id, 1u); break;
              default:
                umma_fill(d, ah, bh, id, acc); umma_lastuse(d, ah, bl, id, 1u); umma_plain(d, al, bh, id, 1u); break;
            }
            acc = 1;
          }
          commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, par);
        par ^= 1;
      }
      t1 = clock64();
    }
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x == 0) cyc_out[0] = t1 - t0;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(d + ((uint32_t)(warp * 32) << 16) + c0)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int e = 0; e < 16; ++e) D[(size_t)tid * N + c0 + e] = __uint_as_float(r[e]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(256) : "memory");
}

static size_t smem_for(int N) { return (size_t)4 * KS_MAX * A_ROWS * 16 + (size_t)4 * KS_MAX * N * 16 + 1024; }

int main() {
  cudaFuncSetAttribute(k_col, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  // every case below fits (checked by the launch error)
  long long* dcyc;
  cudaMalloc(&dcyc, 16);
  float* dD;
  cudaMalloc(&dD, 128 * 256 * 4);
  const char* pname[] = {"plain x3", "plain, fill, lastuse", "discard, fill, lastuse", "LIAR: lastuse names A_lo", "fill, lastuse, plain (reordered)"};
  // ---- numerics: ONE K-step (3 MMAs), the last burst of the last rep is what stays in TMEM ----
  for (int N : {64, 160})
    for (int pat = 0; pat < 5; ++pat) {
      Cfg c{N, pat, 1, 1, 0};
      cudaMemset(dD, 0, 128 * 256 * 4);
      k_col<<<1, 128, smem_for(N)>>>(c, dcyc, dD);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      std::vector<float> hD((size_t)128 * N);
      cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
      int bad_true = 0, bad_refetch = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
          // "true" = what the split scheme wants: lo*hi + hi*hi + hi*lo;  "refetch" (liar only) = third product with A_lo
          float t = 0.f, rf = 0.f;
          for (int k = 0; k < 16; ++k) {
            const float ahv = a_val_hi(i, k), alv = a_val_lo(i, k), bhv = b_val_hi(n, k), blv = b_val_lo(n, k);
            t += alv * bhv + ahv * bhv + ahv * blv;
            rf += alv * bhv + ahv * bhv + alv * blv;
          }
          if (t != hD[(size_t)i * N + n]) ++bad_true;
          if (rf != hD[(size_t)i * N + n]) ++bad_refetch;
        }
      printf("numerics N=%3d pattern %d (%s): %d mismatches vs lo*hi+hi*hi+hi*lo, %d vs the liar's re-fetched product %s\n", N, pat, pname[pat],
             bad_true, bad_refetch, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  // ---- timing: 148 CTAs, bursts of 10 K-steps (30 MMAs), 8 bursts ----
  for (int shift : {0, 1})
    for (int N : {32, 64, 128, 160})
      for (int pat : {0, 1, 2, 4}) {
        Cfg c{N, pat, 10, 8, shift};
        k_col<<<148, 128, smem_for(N)>>>(c, dcyc, nullptr);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("time N=%3d shift=%d pattern %d (%-32s): %7.1f cycles per MMA (bursts of 30 incl. commit+wait) %s\n", N, shift, pat, pname[pat],
               (double)cyc / (8 * 30), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
