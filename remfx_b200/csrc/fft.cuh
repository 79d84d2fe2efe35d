// cuFFT-free shared-memory FFT building blocks (Stockham autosort, radix-2 family: radix-4 = two fused
// radix-2 stages, plus one plain radix-2 stage when log2(N) is odd).
//
// A real FFT of n_fft points is computed as a complex FFT of NC = n_fft/2 packed points
// z[n] = x[2n] + i x[2n+1] followed by the standard even/odd split.  The pass functions are
// __host__ __device__ and take the "thread index" as an argument so that the index arithmetic
// can be emulated and checked on the CPU (tests/host/fft_emul.cpp); on the GPU one pass is executed
// by NC/4 threads with a __syncthreads() between passes.
//
// tw[m] = exp(-2 pi i m / n_fft), m in [0, n_fft): precomputed in double precision on the host.
#pragma once
#include <cuda_runtime.h>

namespace rfx {

struct cpx {
  float x, y;
};

__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// One radix-4 Stockham pass over NC points with NC/4 work items; `j` is the work-item index,
// `Ns` the size of the sub-transforms already completed (1, 4, 16, ...), `twstride` = n_fft / (4 Ns).
__host__ __device__ __forceinline__ void fft_pass_r4(const float2* __restrict__ in, float2* __restrict__ out,
                                                    const float2* __restrict__ tw, int NC, int Ns, int twstride, int j) {
  const int T = NC >> 2;
  const int k = j & (Ns - 1);
  float2 v0 = in[j], v1 = in[j + T], v2 = in[j + 2 * T], v3 = in[j + 3 * T];
  if (Ns > 1) {
    const int m = k * twstride;
    v1 = cmul(v1, tw[m]);
    v2 = cmul(v2, tw[2 * m]);
    v3 = cmul(v3, tw[3 * m]);
  }
  const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
  const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
  const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
  const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);  // (v1 - v3) * (-i)
  const int j0 = ((j - k) << 2) + k;
  out[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
  out[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
  out[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
  out[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
}

// Final radix-2 pass (only when log2(NC) is odd): Ns = NC/2, work item j in [0, NC/2).
__host__ __device__ __forceinline__ void fft_pass_r2_last(const float2* __restrict__ in, float2* __restrict__ out,
                                                         const float2* __restrict__ tw, int NC, int j) {
  const int h = NC >> 1;
  const float2 v0 = in[j];
  const float2 v1 = cmul(in[j + h], tw[2 * j]);  // exp(-2 pi i j / NC) = tw[j * n_fft / NC] = tw[2 j]
  out[j] = make_float2(v0.x + v1.x, v0.y + v1.y);
  out[j + h] = make_float2(v0.x - v1.x, v0.y - v1.y);
}

// Even/odd split: spectrum bin k (0..NC) of the real signal from the packed transform Z (NC points).
__host__ __device__ __forceinline__ float2 rfft_post(const float2* __restrict__ Z, const float2* __restrict__ tw, int NC, int k) {
  const float2 zk = Z[k & (NC - 1)];
  const float2 zn = Z[(NC - k) & (NC - 1)];
  const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
  const float2 d = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y));
  const float2 o = make_float2(d.y, -d.x);  // d / i
  const float2 t = cmul(o, tw[k]);
  return make_float2(e.x + t.x, e.y + t.y);
}

// Inverse of rfft_post: packed-spectrum point k (0..NC-1) from Hermitian half spectrum X (bins 0..NC);
// returns conj(Zc[k]) so that a FORWARD complex FFT followed by a conjugate gives NC * z[n].
__host__ __device__ __forceinline__ float2 irfft_pre(float2 xk, float2 xn, float2 twk) {
  const float2 e = make_float2(0.5f * (xk.x + xn.x), 0.5f * (xk.y - xn.y));
  const float2 d = make_float2(0.5f * (xk.x - xn.x), 0.5f * (xk.y + xn.y));
  const float2 o = cmul(d, make_float2(twk.x, -twk.y));  // d * exp(+2 pi i k / n_fft)
  // Zc = e + i o ; return conj(Zc)
  return make_float2(e.x - o.y, -(e.y + o.x));
}

#ifdef __CUDACC__
// Cooperative complex FFT of NC = 2^LOG2NC points by NC/4 threads.  Input in `a`; returns the buffer
// (a or b) that holds the result.  Ends with a __syncthreads() so the result is visible to all threads.
template <int LOG2NC>
__device__ __forceinline__ float2* fft_block(float2* a, float2* b, const float2* __restrict__ tw, int tid) {
  constexpr int NC = 1 << LOG2NC;
  constexpr int NFFT = 2 * NC;
  float2* in = a;
  float2* out = b;
  __syncthreads();
#pragma unroll
  for (int Ns = 1; Ns * 4 <= NC; Ns *= 4) {
    fft_pass_r4(in, out, tw, NC, Ns, NFFT / (4 * Ns), tid);
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
  }
  if (LOG2NC & 1) {
    fft_pass_r2_last(in, out, tw, NC, tid);
    fft_pass_r2_last(in, out, tw, NC, tid + NC / 4);
    __syncthreads();
    float2* t = in;
    in = out;
    out = t;
  }
  return in;
}
#endif

}  // namespace rfx
