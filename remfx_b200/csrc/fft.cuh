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

// The bin pair (k, NC - k), 0 <= k < NC/2, from zk = Z[k], zn = Z[(NC - k) mod NC] and twk = tw[k] (tw[NC - k] = -conj(tw[k])):
//   X[k] = E + T,  X[NC - k] = conj(E - T),  E = (zk + conj(zn)) / 2,  T = tw[k] (zk - conj(zn)) / (2 i).
// Returns 2 X (the factor 1/2 is exact in binary floating point and is folded into the caller's scale), so the values equal
// rfft_post()'s bit for bit up to the table rounding of tw[NC - k].  k = 0 gives the DC bin and the Nyquist bin (k = NC).
__host__ __device__ __forceinline__ void rfft_post_pair2(float2 zk, float2 zn, float2 twk, float2& xk2, float2& xn2) {
  const float2 e = make_float2(zk.x + zn.x, zk.y - zn.y);
  const float2 d = make_float2(zk.x - zn.x, zk.y + zn.y);
  const float2 t = cmul(make_float2(d.y, -d.x), twk);
  xk2 = make_float2(e.x + t.x, e.y + t.y);
  xn2 = make_float2(e.x - t.x, t.y - e.y);
}

// Inverse of the above for the packed-spectrum points k and NC - k (0 < k < NC/2) from the Hermitian half-spectrum bins
// xk = X[k], xn = X[NC - k]: returns 2 conj(Zc[k]) and 2 conj(Zc[NC - k]) (see irfft_pre; the 1/2 is folded into the caller's scale).
__host__ __device__ __forceinline__ void irfft_pre_pair2(float2 xk, float2 xn, float2 twk, float2& zk2, float2& zn2) {
  const float2 e = make_float2(xk.x + xn.x, xk.y - xn.y);
  const float2 d = make_float2(xk.x - xn.x, xk.y + xn.y);
  const float2 o = cmul(d, make_float2(twk.x, -twk.y));  // d * exp(+2 pi i k / n_fft)
  zk2 = make_float2(e.x - o.y, -(e.y + o.x));
  zn2 = make_float2(e.x + o.y, e.y - o.x);
}

// ------------------------------------------------------------------------------------------------------
// 1024-point complex FFT (n_fft = 2048) as THREE register passes -- radix 16, 16, 4 -- by 64 threads:
// two shared-memory exchanges per frame instead of five, 16-point butterflies entirely in registers.
// Same Stockham indexing as fft_pass_r4, generalised: work item j loads in[j + r N/R], multiplies by
// exp(-2 pi i r k / (Ns R)) with k = j mod Ns, transforms, and stores to out[(j - k) R + k + r Ns].
// Intermediate buffers are padded by one element every 16 (fft1024_pad) so that the stride-16 accesses of the
// radix-16 passes spread over the banks; the last pass writes the natural, unpadded order.
// ------------------------------------------------------------------------------------------------------
constexpr int FFT1024_BUF = 1024 + 64;  // float2 elements of one (padded) frame buffer

__host__ __device__ __forceinline__ int fft1024_pad(int i) { return i + (i >> 4); }

__host__ __device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = make_float2(a0.x + a2.x, a0.y + a2.y), d02 = make_float2(a0.x - a2.x, a0.y - a2.y);
  const float2 s13 = make_float2(a1.x + a3.x, a1.y + a3.y);
  const float2 d13 = make_float2(a1.y - a3.y, a3.x - a1.x);  // (a1 - a3) * (-i)
  a0 = make_float2(s02.x + s13.x, s02.y + s13.y);
  a1 = make_float2(d02.x + d13.x, d02.y + d13.y);
  a2 = make_float2(s02.x - s13.x, s02.y - s13.y);
  a3 = make_float2(d02.x - d13.x, d02.y - d13.y);
}

// In-register 16-point DFT: v[n], n = g + 4 m  ->  v[k], k = q + 4 p (natural order on return).
__host__ __device__ __forceinline__ void dft16(float2 (&v)[16]) {
  // stage 1: DFT4 over m for every g (elements g, g+4, g+8, g+12) -> y[g][q] stored at v[g + 4 q]
#pragma unroll
  for (int g = 0; g < 4; ++g) dft4(v[g], v[g + 4], v[g + 8], v[g + 12]);
  // twiddle y[g][q] *= W16^(g q)
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  // W16^1 = (C1, -S1), W16^2 = (R2, -R2), W16^3 = (S1, -C1), W16^4 = (0, -1), W16^6 = (-R2, -R2), W16^9 = (-C1, S1)
  v[1 + 4] = cmul(v[1 + 4], make_float2(C1, -S1));
  v[1 + 8] = cmul(v[1 + 8], make_float2(R2, -R2));
  v[1 + 12] = cmul(v[1 + 12], make_float2(S1, -C1));
  v[2 + 4] = cmul(v[2 + 4], make_float2(R2, -R2));
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);  // * (-i)
  v[2 + 12] = cmul(v[2 + 12], make_float2(-R2, -R2));
  v[3 + 4] = cmul(v[3 + 4], make_float2(S1, -C1));
  v[3 + 8] = cmul(v[3 + 8], make_float2(-R2, -R2));
  v[3 + 12] = cmul(v[3 + 12], make_float2(-C1, S1));
  // stage 2: DFT4 over g for every q (elements 4 q .. 4 q + 3) -> X[q + 4 p] lands at v[4 q + p]
#pragma unroll
  for (int q = 0; q < 4; ++q) dft4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  // v[4 q + p] holds X[q + 4 p]: transpose the 4 x 4 index to natural order
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int p = q + 1; p < 4; ++p) {
      const float2 t = v[4 * q + p];
      v[4 * q + p] = v[4 * p + q];
      v[4 * p + q] = t;
    }
}

// Radix-16 pass for work item j in [0, 64), in three steps so that the GPU version can run IN PLACE (all loads of the CTA,
// barrier, all stores).  IN_PAD / OUT_PAD: the buffer uses the padded index map.
template <bool IN_PAD>
__host__ __device__ __forceinline__ void fft1024_r16_load(const float2* __restrict__ in, int j, float2 (&v)[16]) {
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = in[IN_PAD ? fft1024_pad(j + 64 * r) : j + 64 * r];
}
__host__ __device__ __forceinline__ void fft1024_r16_compute(const float2* __restrict__ tw, int Ns, int j, float2 (&v)[16]) {
  if (Ns > 1) {
    const int m = (j & (Ns - 1)) * (2048 / (Ns * 16));  // exp(-2 pi i r k / (Ns 16)) = tw[r k 2048 / (16 Ns)]
#pragma unroll
    for (int r = 1; r < 16; ++r) v[r] = cmul(v[r], tw[r * m]);
  }
  dft16(v);
}
template <bool OUT_PAD>
__host__ __device__ __forceinline__ void fft1024_r16_store(float2* __restrict__ out, int Ns, int j, const float2 (&v)[16]) {
  const int k = j & (Ns - 1);
  const int j0 = (j - k) * 16 + k;
#pragma unroll
  for (int r = 0; r < 16; ++r) out[OUT_PAD ? fft1024_pad(j0 + r * Ns) : j0 + r * Ns] = v[r];
}
template <bool IN_PAD, bool OUT_PAD>
__host__ __device__ __forceinline__ void fft1024_pass_r16(const float2* __restrict__ in, float2* __restrict__ out,
                                                          const float2* __restrict__ tw, int Ns, int j) {
  float2 v[16];
  fft1024_r16_load<IN_PAD>(in, j, v);
  fft1024_r16_compute(tw, Ns, j, v);
  fft1024_r16_store<OUT_PAD>(out, Ns, j, v);
}

// Last pass: radix 4 with Ns = 256 for work item j in [0, 256); padded input, natural output.
__host__ __device__ __forceinline__ void fft1024_r4_last_load(const float2* __restrict__ in, const float2* __restrict__ tw, int j,
                                                              float2 (&v)[4]) {
  v[0] = in[fft1024_pad(j)];
  const int m = 2 * j;  // exp(-2 pi i r j / 1024) = tw[2 r j]
  v[1] = cmul(in[fft1024_pad(j + 256)], tw[m]);
  v[2] = cmul(in[fft1024_pad(j + 512)], tw[2 * m]);
  v[3] = cmul(in[fft1024_pad(j + 768)], tw[3 * m]);
  dft4(v[0], v[1], v[2], v[3]);
}
__host__ __device__ __forceinline__ void fft1024_pass_r4_last(const float2* __restrict__ in, float2* __restrict__ out,
                                                              const float2* __restrict__ tw, int j) {
  float2 v[4];
  fft1024_r4_last_load(in, tw, j, v);
#pragma unroll
  for (int r = 0; r < 4; ++r) out[j + 256 * r] = v[r];
}

#ifdef __CUDACC__
// Barrier of one frame's 64 threads (named barrier 1 + frame slot) or of the whole CTA (GROUP_SYNC = false).  With the
// per-frame form the four frames of a CTA drift apart, so one frame's shared-memory phases overlap another's butterflies.
template <bool GROUP_SYNC>
__device__ __forceinline__ void fft_sync(int tid) {
  if (GROUP_SYNC) asm volatile("bar.sync %0, 64;" ::"r"(1 + (tid >> 6)) : "memory");
  else __syncthreads();
}

// Four 1024-point FFTs at once by a 256-thread CTA, IN PLACE: frame g = tid / 64 lives in buf + g FFT1024_BUF, natural order
// on entry and on return.  Every pass is load-all / barrier / store-all.  Ends with a barrier (of the kind chosen).
template <bool LEAD_SYNC = true, bool GROUP_SYNC = false>
__device__ __forceinline__ void fft1024_x4(float2* buf, const float2* __restrict__ tw, int tid) {
  const int t = tid & 63;
  float2* f = buf + (tid >> 6) * FFT1024_BUF;
  float2 v[16];
  if (LEAD_SYNC) fft_sync<GROUP_SYNC>(tid);  // false: the caller has already put a barrier behind the writes of the input
  fft1024_r16_load<false>(f, t, v);
  fft1024_r16_compute(tw, 1, t, v);
  fft_sync<GROUP_SYNC>(tid);
  fft1024_r16_store<true>(f, 1, t, v);
  fft_sync<GROUP_SYNC>(tid);
  fft1024_r16_load<true>(f, t, v);
  fft1024_r16_compute(tw, 16, t, v);
  fft_sync<GROUP_SYNC>(tid);
  fft1024_r16_store<true>(f, 16, t, v);
  fft_sync<GROUP_SYNC>(tid);
  float2 w[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m) fft1024_r4_last_load(f, tw, t + 64 * m, w[m]);
  fft_sync<GROUP_SYNC>(tid);
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int r = 0; r < 4; ++r) f[t + 64 * m + 256 * r] = w[m][r];
  fft_sync<GROUP_SYNC>(tid);
}

// Same, with the first pass's inputs already in registers (v[r] = element t + 64 r of the thread's frame): the windowed samples
// (STFT) go from their staging buffer straight into the radix-16 butterflies.  The frame buffer must not be in use on entry.
template <bool GROUP_SYNC = false>
__device__ __forceinline__ void fft1024_x4_regs(float2* buf, const float2* __restrict__ tw, int tid, float2 (&v)[16]) {
  const int t = tid & 63;
  float2* f = buf + (tid >> 6) * FFT1024_BUF;
  fft1024_r16_compute(tw, 1, t, v);
  fft1024_r16_store<true>(f, 1, t, v);
  fft_sync<GROUP_SYNC>(tid);
  fft1024_r16_load<true>(f, t, v);
  fft1024_r16_compute(tw, 16, t, v);
  fft_sync<GROUP_SYNC>(tid);
  fft1024_r16_store<true>(f, 16, t, v);
  fft_sync<GROUP_SYNC>(tid);
  float2 w[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m) fft1024_r4_last_load(f, tw, t + 64 * m, w[m]);
  fft_sync<GROUP_SYNC>(tid);
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int r = 0; r < 4; ++r) f[t + 64 * m + 256 * r] = w[m][r];
  fft_sync<GROUP_SYNC>(tid);
}

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
