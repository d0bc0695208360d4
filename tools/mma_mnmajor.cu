// Development microbenchmark (not part of the library): are the library's operand images [chunk of 8 channels][slot][8]
// usable as MN-MAJOR tcgen05.mma operands with the SLOT stream as K?  That is what a weight gradient needs:
//     dW[ci][co] = sum over slots s of  X[s][ci] * G[s + shift][co]
// i.e. M = ci, N = co, K = slots -- both operands "transposed" with respect to the forward use of the same images.
// In the canonical no-swizzle MN-major layout a core matrix is 8 K-rows of 16 bytes (8 contiguous MN elements each),
// which is exactly 8 consecutive slots of one chunk plane.  Unknown without documentation: which of LBO / SBO is the
// stride between K groups (8 slots = 128 B) and which the stride between MN groups (the plane pitch), and whether a
// start address that is not 128-byte aligned (a tap shift of `shift` slots = shift * 16 B) is honoured.
// The test tries both assignments and several shifts with small-integer data and compares exactly with the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_mnmajor tools/mma_mnmajor.cu
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
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// fp16 A and B, fp32 accumulate, M = 128, BOTH operands MN-major (bits 15 and 16)
__device__ __forceinline__ uint32_t idesc_mn(int N, int M = 128) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {  // SWIZZLE_NONE
  const uint32_t lo = ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16);
  const uint32_t hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
  return ((uint64_t)hi << 32) | lo;
}

constexpr int SLOTS = 192;  // slots per chunk plane (K extent incl. room for shifts)
constexpr int MA = 128;     // A channels (M)
constexpr int KT = 64;      // slots contracted (4 MMAs of K = 16)
__host__ __device__ inline int a_val(int m, int s) { return (m * 3 + s * 5) % 7 - 3 + (s == (m & 63) ? 2 : 0); }  // rows pairwise different
__host__ __device__ inline int b_val(int n, int s) { return (n + 2 * s) % 9 - 4; }

struct Cfg { int N; int shift; int swap; int M; };  // M = 128 or 64  // swap 0: LBO = 128 (K groups), SBO = plane pitch (MN groups); 1: the other way round

__global__ void __launch_bounds__(128, 1) k_mn(Cfg c, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = c.N;
  const uint32_t plane = SLOTS * 16;
  __half* A = reinterpret_cast<__half*>(smem);                                 // [MA/8][SLOTS][8]
  __half* B = reinterpret_cast<__half*>(smem + (MA / 8) * plane);              // [N/8][SLOTS][8]
  for (int i = tid; i < (MA / 8) * SLOTS * 8; i += 128) {
    const int e = i & 7, s = (i >> 3) % SLOTS, ch = (i >> 3) / SLOTS;
    A[i] = __int2half_rn(a_val(ch * 8 + e, s));
  }
  for (int i = tid; i < (N / 8) * SLOTS * 8; i += 128) {
    const int e = i & 7, s = (i >> 3) % SLOTS, ch = (i >> 3) / SLOTS;
    B[i] = __int2half_rn(b_val(ch * 8 + e, s));
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
  if (tid == 0) {
    const uint32_t lbo = c.swap ? plane : 128u, sbo = c.swap ? 128u : plane;
    for (int ks = 0; ks < KT / 16; ++ks) {
      const uint64_t ad = mk_desc(smem_u32(A) + (uint32_t)(ks * 16) * 16u, lbo, sbo);
      const uint64_t bd = mk_desc(smem_u32(B) + (uint32_t)(ks * 16 + c.shift) * 16u, lbo, sbo);
      umma(d, ad, bd, idesc_mn(N, c.M), ks ? 1u : 0u);
    }
    commit(&bar);
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(256) : "memory");
}

int main() {
  cudaFuncSetAttribute(k_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  float* dD;
  cudaMalloc(&dD, 128 * 256 * 4);
  for (int N : {64, 160})
    for (int swap = 0; swap < 2; ++swap)
      for (int shift : {0, 1, 8, 17, 18}) {
        Cfg c{N, shift, swap, 128};
        const size_t smem = (size_t)(MA / 8 + N / 8) * SLOTS * 16 + 1024;
        cudaMemset(dD, 0, 128 * 256 * 4);
        k_mn<<<1, 128, smem>>>(c, dD);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        std::vector<float> hD((size_t)128 * N);
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int s = 0; s < KT; ++s) ref += (float)a_val(m, s) * (float)b_val(n, s + shift);
            if (ref != hD[(size_t)m * N + n]) ++bad;
          }
        printf("MN-major N=%3d %s shift=%2d : %s (%d of %d mismatches) %s\n", N,
               swap ? "LBO=plane SBO=128" : "LBO=128 SBO=plane", shift, bad ? "WRONG" : "ok", bad, 128 * N,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
      }
  // ---- M = 64: where do the 64 rows of D land in the 128 TMEM lanes? ----
  {
    const int N = 64;
    Cfg c{N, 1, 0, 64};
    const size_t smem = (size_t)(MA / 8 + N / 8) * SLOTS * 16 + 1024;
    cudaMemset(dD, 0xff, 128 * 256 * 4);  // NaN pattern: untouched lanes stay recognisable
    k_mn<<<1, 128, smem>>>(c, dD);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    std::vector<float> hD((size_t)128 * N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    printf("M=64 N=64 shift=1: %s\n", e == cudaSuccess ? "ran" : cudaGetErrorString(e));
    for (int m = 0; m < 64; ++m) {
      int found = -1, nfound = 0;
      for (int lane = 0; lane < 128; ++lane) {
        bool all = true;
        for (int n = 0; n < N && all; ++n) {
          float ref = 0.f;
          for (int s = 0; s < KT; ++s) ref += (float)a_val(m, s) * (float)b_val(n, s + 1);
          all = ref == hD[(size_t)lane * N + n];
        }
        if (all) { if (found < 0) found = lane; ++nfound; }
      }
      if (m % 8 == 0 || found != m) printf("  row %2d -> lane %d (%d lanes match)\n", m, found, nfound);
    }
  }
  return 0;
}
