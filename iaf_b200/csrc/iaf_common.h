// Internal declarations shared by the CUDA translation units of libiaf_b200.so.
#pragma once
#ifdef IAF_EMU
// Host emulation of the CUDA subset the SIMT kernels use (tests/emu/cuda_emu.h): TEST INFRASTRUCTURE ONLY, it lets the
// CPU test-suite execute the kernels' index logic without a GPU.  Never compiled into libiaf_b200.so.
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
// kernel<<<grid, block, smem, stream>>>(args...) and the dynamic shared-memory window, spelled as macros so that the
// same sources also compile under the host emulation above
#define IAF_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define IAF_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
// 4-byte asynchronous global -> shared copy (LDGSTS), zero-filled when !valid; src must be a mapped address either way
__device__ __forceinline__ void iaf_cp_async4(float* dst, const float* src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void iaf_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void iaf_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif
// Raise a kernel's dynamic shared-memory limit to everything the device allows (opt-in maximum minus the kernel's static
// shared memory).  The attribute is per kernel, not per plan: setting it to one plan's size would LOWER it for plans
// created earlier with a larger footprint, so it is only ever set to the device maximum (idempotent).
template <class K>
static inline cudaError_t iaf_smem_optin(K kernel) {
#ifdef IAF_EMU
  (void)kernel;
  return cudaSuccess;
#else
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes fa;
  e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
#endif
}
#include <stdint.h>
#include "../../include/iaf_b200.h"

#define IAF_NTAPS 5          // live taps of the 3x3 AR mask: (ky,kx) = (1,1)c (1,2) (2,0) (2,1) (2,2)
#define IAF_MAX_STAGES (IAF_MAX_HIDDEN + 1)

// One conv stage as the SIMT kernel sees it (weights already masked, normalised, scaled).
struct IafStageDev {
  const float* w;     // [5][cin][cout_pad], cout contiguous (heads: see iaf_pack.cu for the column order)
  const float* bias;  // [cout_pad]
  const float* padw;  // [4][cout_pad]: Theano pad-channel weights of taps 1..4, nullptr for TF
  int cin, cout, cout_pad;
};

struct IafSimtParams {
  // inputs
  const float* z;          // [B,C,H,W]   (layer mode: eps)
  const float* ctx;        // [B,hidden0,H,W]
  const float* post_mean;  // layer mode only
  const float* post_logsd;
  const float* prior_mean;
  const float* prior_logsd;
  // outputs (nullable)
  float* z_out;
  float* logsd_out;        // step mode: arw_logsd; layer mode: kl per element
  float* m_out;            // multiconv mode: head 0
  float* s_out;            // multiconv mode: head 1
  float* bc_out;           // layer mode: [B,C] sum over (h,w) of kl
  float* persample_out;    // step: logdet [B]; layer: kl_cost [B]
  float* hid_out[IAF_MAX_HIDDEN];  // training forward: hidden activations [B][hidden[j]][HW], nullable
  float* partial;          // [B][n_bands][C] per-band per-channel partial sums
  unsigned* counter;       // [B] band arrival counters (self-resetting)
  IafStageDev stage[IAF_MAX_STAGES];
  int n_stages;            // n_hidden + 1 (last = heads)
  int n_heads, head_c, head_pad;
  int B, C, H, W, P;       // P = smem row pitch = 8*ceil(W/8) + 2
  int band_rows, n_bands;
  int flip;                // 1: Theano orientation (data point-reflected on load/store)
  int nl;
  int mode;                // 0 multiconv, 1 step, 2 layer
  float scale;             // 0.1
  int bufz_elems, bufa_elems, bufb_elems;
};

enum { IAF_MODE_MULTICONV = 0, IAF_MODE_STEP = 1, IAF_MODE_LAYER = 2 };

// Raw-parameter description handed to the pack kernel.
struct IafPackLayer {
  const float* w;      // reference layout (TF [3,3,Cin,Cout] | Theano [Cout,Cin+1,3,3])
  const float* scale;  // g | s
  const float* bias;   // b
  float* w_out;        // simt packed
  float* bias_out;
  float* padw_out;     // Theano only
  int cin, cout, cout_pad;
  int zerodiag;        // heads: 1
  int head_pairs;      // 1: two equal heads interleaved in groups of 4 columns (m0..3,s0..3,...)
  int head_c, head_pad;
  int col0;            // first packed column this (head) layer owns when head_pairs (0 or 4)
};

struct IafPackParams {
  IafPackLayer layer[IAF_MAX_HIDDEN + IAF_MAX_HEADS];
  int n_layers;
  int variant;
};

cudaError_t iaf_launch_pack(const IafPackParams& p, int max_cout, cudaStream_t stream);
cudaError_t iaf_launch_simt(const IafSimtParams& p, size_t smem_bytes, cudaStream_t stream);
cudaError_t iaf_simt_set_smem();
