// Weight preparation for the IAF step: mask o V -> per-output-channel l2 normalise -> scale,
// written straight into the packed layouts the step kernels read.  One launch for the whole
// stack (the reference re-derives W inside the graph on every step through ~6 tiny kernels
// per conv: tf_utils/layers.py:53-60, graphy/nodes/ar.py:312-321 with 267-281).
//
// Only the 5 live taps of the 3x3 AR mask are kept (tf_utils/layers.py:134-141 ==
// graphy/nodes/ar.py:241-264): t -> (ky,kx) = (1,1) (1,2) (2,0) (2,1) (2,2); the centre tap
// carries the MADE channel mask (layers.py:115-131).
#include "iaf_common.h"

__device__ __forceinline__ bool iaf_centre_visible(int ci, int co, int cin, int cout, int zd) {
  if (cout >= cin) {
    const int k = cout / cin;
    const int i = co / k;
    return zd ? (ci < i) : (ci <= i);
  }
  const int k = cin / cout;
  return zd ? (ci < co * k) : (ci < (co + 1) * k);
}

__device__ __forceinline__ int iaf_tap_ky(int t) { return t < 2 ? 1 : 2; }
__device__ __forceinline__ int iaf_tap_kx(int t) { return t == 0 ? 1 : (t == 1 ? 2 : t - 2); }

__device__ __forceinline__ float iaf_raw_weight(const IafPackLayer& L, int variant, int t, int ci, int co) {
  const int ky = iaf_tap_ky(t), kx = iaf_tap_kx(t);
  if (variant == IAF_VARIANT_TF) return L.w[((size_t)(ky * 3 + kx) * L.cin + ci) * L.cout + co];
  return L.w[(((size_t)co * (L.cin + 1) + ci) * 3 + ky) * 3 + kx];
}

__global__ void __launch_bounds__(128) iaf_pack_kernel(const __grid_constant__ IafPackParams p) {
  const IafPackLayer& L = p.layer[blockIdx.y];
  const int co = blockIdx.x;
  if (co >= L.cout) return;
  const int tid = threadIdx.x;
  const int n_real = L.cin * IAF_NTAPS;
  const int n_pad = (p.variant == IAF_VARIANT_THEANO) ? 4 : 0;  // pad channel: taps 1..4 (centre masked, ar.py:249-262)

  // pass 1: sum of squares of the masked row
  float ss = 0.f;
  for (int e = tid; e < n_real + n_pad; e += blockDim.x) {
    float v;
    if (e < n_real) {
      const int t = e / L.cin, ci = e % L.cin;
      v = iaf_raw_weight(L, p.variant, t, ci, co);
      if (t == 0 && !iaf_centre_visible(ci, co, L.cin, L.cout, L.zerodiag)) v = 0.f;
    } else {
      v = iaf_raw_weight(L, p.variant, e - n_real + 1, L.cin, co);
    }
    ss = fmaf(v, v, ss);
  }
  __shared__ float red[128];
  red[tid] = ss;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (tid < s) red[tid] += red[tid + s];
    __syncthreads();
  }
  ss = red[0];
  float factor;
  if (p.variant == IAF_VARIANT_TF)
    factor = expf(L.scale[co]) / sqrtf(fmaxf(ss, 1e-12f));        // layers.py:60, l2_normalize eps
  else
    factor = expf(3.0f * L.scale[co]) / (sqrtf(ss) + 1e-8f);      // ar.py:277-281,316 (logscale_scale = 3)

  const int col = L.head_pairs ? ((co >> 2) * 8 + L.col0 + (co & 3)) : co;
  for (int e = tid; e < n_real + n_pad; e += blockDim.x) {
    if (e < n_real) {
      const int t = e / L.cin, ci = e % L.cin;
      float v = iaf_raw_weight(L, p.variant, t, ci, co);
      if (t == 0 && !iaf_centre_visible(ci, co, L.cin, L.cout, L.zerodiag)) v = 0.f;
      L.w_out[((size_t)t * L.cin + ci) * L.cout_pad + col] = v * factor;
    } else {
      const int t = e - n_real + 1;
      L.padw_out[(size_t)(t - 1) * L.cout_pad + col] = iaf_raw_weight(L, p.variant, t, L.cin, co) * factor;
    }
  }
  if (tid == 0) L.bias_out[col] = L.bias[co];
}

cudaError_t iaf_launch_pack(const IafPackParams& p, int max_cout, cudaStream_t stream) {
  dim3 grid(max_cout, p.n_layers);
  IAF_LAUNCH(iaf_pack_kernel, grid, 128, 0, stream, p);
  return cudaGetLastError();
}
