// Element-wise / reduction / small SIMT kernels of the Hybrid-Demucs path (torchaudio/models/_hdemucs.py).
// Activations are channel-last (B, Y, X, C): freq-branch tensors use Y = time frame, X = frequency bin group;
// time-branch tensors use Y = 1, X = time.  "split" = two bf16 planes (hi, lo), see gemm2.cu.
#pragma once
#include "hd_internal.h"

namespace rfx {
namespace hd {

__device__ __forceinline__ float gelu_exact(float v) { return gelu_fast(v); }  // erf-form GELU, |err| <= 3.4e-7

__device__ __forceinline__ void store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float (&o)[8]) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(o[2 * i] - hf.x, o[2 * i + 1] - hf.y);
    ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}
__device__ __forceinline__ void load_split8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&o)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(hi + off);
  const uint4 l = *reinterpret_cast<const uint4*>(lo + off);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[i]));
    const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[i]));
    o[2 * i] = hf.x + lf.x;
    o[2 * i + 1] = hf.y + lf.y;
  }
}

// ---- per-item mean / unbiased std of a contiguous fp32 block (HDemucs input normalisation, _hdemucs.py:553-563) ----
// stats[2b] = mean, stats[2b+1] = std.  One block per item.
__global__ void __launch_bounds__(1024) item_stats_kernel(const float* __restrict__ x, long long n, float* __restrict__ stats) {
  __shared__ double r0[32], r1[32];
  const float* p = x + (size_t)blockIdx.x * n;
  double s = 0.0, ss = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = p[i];
    s += v;
    ss += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = s; r1[threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += r0[w]; b += r1[w]; }
    const double mean = a / (double)n;
    const double var = (b - a * mean) / (double)(n - 1);
    stats[2 * blockIdx.x] = (float)mean;
    stats[2 * blockIdx.x + 1] = (float)sqrt(var > 0.0 ? var : 0.0);
  }
}

// (v - mean_b) / (1e-5 + std_b): fp32 [B][n] -> split planes [B][n] (n % 8 == 0), or fp32 when ohi == nullptr
__global__ void __launch_bounds__(256) item_normalize_kernel(const float* __restrict__ x, long long n, const float* __restrict__ stats,
                                                             __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo,
                                                             float* __restrict__ of) {
  const int b = blockIdx.y;
  const long long i8 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8;
  if (i8 >= n) return;
  const float mean = stats[2 * b], inv = 1.0f / (1e-5f + stats[2 * b + 1]);
  const float4 a = *reinterpret_cast<const float4*>(x + (size_t)b * n + i8);
  const float4 c = *reinterpret_cast<const float4*>(x + (size_t)b * n + i8 + 4);
  float o[8] = {(a.x - mean) * inv, (a.y - mean) * inv, (a.z - mean) * inv, (a.w - mean) * inv,
                (c.x - mean) * inv, (c.y - mean) * inv, (c.z - mean) * inv, (c.w - mean) * inv};
  if (ohi) {
    store_split8(ohi, olo, (size_t)b * n + i8, o);
  } else {
    *reinterpret_cast<float4*>(of + (size_t)b * n + i8) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(of + (size_t)b * n + i8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// same, one element per thread: lengths that are not a multiple of 8 (whole files; the benchmark chunks take the vector form)
__global__ void __launch_bounds__(256) item_normalize_scalar_kernel(const float* __restrict__ x, long long n, const float* __restrict__ stats,
                                                                    float* __restrict__ of) {
  const int b = blockIdx.y;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  of[(size_t)b * n + i] = (x[(size_t)b * n + i] - stats[2 * b]) * (1.0f / (1e-5f + stats[2 * b + 1]));
}

// ---- (B, Y, X, C) split planes -> (B, Y, Xp, C) with zero rows in [X, Xp): the zero padding a strided time conv applies to an input
//      whose length is not a multiple of its stride (TA:147-150); only taken for lengths off the 1024-sample grid ----
__global__ void __launch_bounds__(256) pad_rows_kernel(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo, int Y, int X, int Xp,
                                                       int C, __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)Y * Xp * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int x = (int)((idx / groups) % Xp);
  const int y = (int)(idx / ((long long)groups * Xp));
  float u[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (x < X) load_split8(ihi, ilo, (((size_t)b * Y + y) * X + x) * C + c0, u);
  store_split8(ohi, olo, (((size_t)b * Y + y) * Xp + x) * C + c0, u);
}

// ---- first time-branch conv: Conv1d(1 -> C, k, stride s, pad p) + bias + GELU on the normalised waveform ----
// xt fp32 [B][T] -> split [B][T/s][C]; one thread per (output step, 8 channels); weights transposed into smem [K][C]
template <int K>
__global__ void __launch_bounds__(256) time_first_kernel(const float* __restrict__ xt, int T, int Lo, int C, int S, int P,
                                                         const float* __restrict__ w /*[C][K]*/, const float* __restrict__ bias,
                                                         __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  extern __shared__ float hd_smem[];  // [K][C] weights, then [C] bias
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) hd_smem[(i % K) * C + i / K] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) hd_smem[K * C + i] = bias[i];
  __syncthreads();
  const unsigned groups = C / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (unsigned)Lo * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int t = (int)(idx / groups);
  float xs[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int i = t * S + j - P;
    xs[j] = (i >= 0 && i < T) ? __ldg(xt + (size_t)b * T + i) : 0.0f;
  }
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = hd_smem[K * C + c0 + i];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const float4 w0 = *reinterpret_cast<const float4*>(hd_smem + j * C + c0);
    const float4 w1 = *reinterpret_cast<const float4*>(hd_smem + j * C + c0 + 4);
    o[0] = fmaf(w0.x, xs[j], o[0]); o[1] = fmaf(w0.y, xs[j], o[1]); o[2] = fmaf(w0.z, xs[j], o[2]); o[3] = fmaf(w0.w, xs[j], o[3]);
    o[4] = fmaf(w1.x, xs[j], o[4]); o[5] = fmaf(w1.y, xs[j], o[5]); o[6] = fmaf(w1.z, xs[j], o[6]); o[7] = fmaf(w1.w, xs[j], o[7]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = gelu_exact(o[i]);
  store_split8(ohi, olo, ((size_t)b * Lo + t) * C + c0, o);
}

// ---- first freq-branch conv: Conv2d(2 -> C, (k,1), (s,1), pad (p,0)) + bias + GELU, input normalised on the fly ----
// Z fp32 [B][T][Fr][2] (re, im), stats (mean, std) per item -> split [B][T][Fr/s][C]; one thread per (pixel, 8 channels)
template <int K>
__global__ void __launch_bounds__(256) freq_first_kernel(const float* __restrict__ Z, const float* __restrict__ stats, int Tf, int Fr, int Fo,
                                                         int C, int S, int P, const float* __restrict__ w /*[C][2][K]*/,
                                                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ ohi,
                                                         __nv_bfloat16* __restrict__ olo) {
  extern __shared__ float hd_smem[];  // [2K][C] weights (row = part * K + tap), then [C] bias
  for (int i = threadIdx.x; i < C * 2 * K; i += blockDim.x) hd_smem[(i % (2 * K)) * C + i / (2 * K)] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) hd_smem[2 * K * C + i] = bias[i];
  __syncthreads();
  const unsigned groups = C / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (unsigned)Tf * Fo * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const unsigned pix = idx / groups;
  const int fo = (int)(pix % Fo);
  const int t = (int)(pix / Fo);
  const float mean = stats[2 * b], inv = 1.0f / (1e-5f + stats[2 * b + 1]);
  const float* zr = Z + (((size_t)b * Tf + t) * Fr) * 2;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = hd_smem[2 * K * C + c0 + i];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int f = fo * S + j - P;
    float2 v = make_float2(0.0f, 0.0f);
    if (f >= 0 && f < Fr) {
      v = __ldg(reinterpret_cast<const float2*>(zr + 2 * f));
      v.x = (v.x - mean) * inv;
      v.y = (v.y - mean) * inv;
    }
    const float4 r0 = *reinterpret_cast<const float4*>(hd_smem + j * C + c0);
    const float4 r1 = *reinterpret_cast<const float4*>(hd_smem + j * C + c0 + 4);
    const float4 i0 = *reinterpret_cast<const float4*>(hd_smem + (K + j) * C + c0);
    const float4 i1 = *reinterpret_cast<const float4*>(hd_smem + (K + j) * C + c0 + 4);
    // same summation order as before: acc = w_re * x_re + (w_im * x_im + acc)
    o[0] = fmaf(r0.x, v.x, fmaf(i0.x, v.y, o[0])); o[1] = fmaf(r0.y, v.x, fmaf(i0.y, v.y, o[1]));
    o[2] = fmaf(r0.z, v.x, fmaf(i0.z, v.y, o[2])); o[3] = fmaf(r0.w, v.x, fmaf(i0.w, v.y, o[3]));
    o[4] = fmaf(r1.x, v.x, fmaf(i1.x, v.y, o[4])); o[5] = fmaf(r1.y, v.x, fmaf(i1.y, v.y, o[5]));
    o[6] = fmaf(r1.z, v.x, fmaf(i1.z, v.y, o[6])); o[7] = fmaf(r1.w, v.x, fmaf(i1.w, v.y, o[7]));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = gelu_exact(o[i]);
  store_split8(ohi, olo, (((size_t)b * Tf + t) * Fo + fo) * C + c0, o);
}

// ---- last freq decoder, fused: ConvTranspose2d(C -> 2, (k,1), (s,1)) + bias, crop `pad`, de-normalise -> complex Z ----
// y split [B][T][Fi][C]; Z[b][t][bin] = (convtr(y)[bin + pad][re, im]) * std + mean.  One thread per (t, bin).
template <int K, int S>
__global__ void __launch_bounds__(256) final_freq_convtr_kernel(const __nv_bfloat16* __restrict__ yhi, const __nv_bfloat16* __restrict__ ylo,
                                                                int Tf, int Fi, int C, int pad, int bins,
                                                                const float* __restrict__ w /*[C][2][K]*/, const float* __restrict__ bias,
                                                                const float* __restrict__ stats, float2* __restrict__ Z) {
  extern __shared__ float hd_smem[];  // [K][C + 1] float2 (re, im) weights; the odd row pitch spreads the S tap phases over banks
  const int pitch = 2 * (C + 1);
  for (int i = threadIdx.x; i < C * 2 * K; i += blockDim.x) {
    const int c = i / (2 * K), part = (i / K) & 1, j = i % K;
    hd_smem[j * pitch + 2 * c + part] = w[i];
  }
  __syncthreads();
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (unsigned)Tf * bins) return;
  const int k = (int)(idx % bins), t = (int)(idx / bins);
  const int pos = k + pad;
  float re = bias[0], im = bias[1];
#pragma unroll
  for (int jj = 0; jj < K / S; ++jj) {
    const int j = pos % S + jj * S;
    const int i = (pos - j) / S;
    if (i < 0 || i >= Fi) continue;
    const size_t off = (((size_t)b * Tf + t) * Fi + i) * C;
    const float2* wj = reinterpret_cast<const float2*>(hd_smem + j * pitch);
    for (int c = 0; c < C; c += 8) {
      float u[8];
      load_split8(yhi, ylo, off + c, u);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float2 wv = wj[c + e];
        re = fmaf(wv.x, u[e], re);
        im = fmaf(wv.y, u[e], im);
      }
    }
  }
  const float mean = stats[2 * b], sd = stats[2 * b + 1];
  Z[((size_t)b * Tf + t) * bins + k] = make_float2(re * sd + mean, im * sd + mean);
}

// ---- GroupNorm statistics over an fp32 tensor (B, Y, X, C): one segment per (b [, x]) and group ----
// accum[(seg * G + g) * 2 + {0,1}] += (sum, sum of squares) in fp64; grid = (nsplit, nseg).
__global__ void __launch_bounds__(256) gn_accum_kernel(const float* __restrict__ raw, int Y, int X, int C, int G, int per_x,
                                                       double* __restrict__ accum) {
  __shared__ double red[8][8];
  const int seg = blockIdx.y;
  const int b = per_x ? seg / X : seg;
  const int xk = per_x ? seg % X : 0;
  const int cpg = C / G;
  const long long npix = per_x ? Y : (long long)Y * X;
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wid = blockIdx.x * 8ll + warp, nw = gridDim.x * 8ll;
  for (long long p = wid; p < npix; p += nw) {
    const size_t pix = per_x ? ((size_t)b * Y + p) * X + xk : (size_t)b * Y * X + p;
    const float* r = raw + pix * C;
    for (int c = lane; c < C; c += 32) {
      const double v = r[c];
      const int g = c / cpg;
      s[g] += v;
      ss[g] += v * v;
    }
  }
  for (int g = 0; g < G; ++g) {
    for (int o = 16; o > 0; o >>= 1) {
      s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
      ss[g] += __shfl_xor_sync(0xffffffffu, ss[g], o);
    }
    if (lane == 0) { red[warp][2 * g] = s[g]; red[warp][2 * g + 1] = ss[g]; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * G) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    atomicAdd(&accum[(size_t)seg * G * 2 + threadIdx.x], t);
  }
}
// stats[(seg * G + g) * 2] = mean, [+1] = 1 / sqrt(biased var + eps)
__global__ void gn_final_kernel(const double* __restrict__ accum, long long count, int n, float eps, float* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = accum[2 * i] / (double)count;
  double var = accum[2 * i + 1] / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// ---- GroupNorm apply + activation; raw fp32 (B, Y, Xr, Cr) -> split (B, Y, Xo, Co), reading x + x_off (crop) ----
// mode 0: y = gn(raw)                     (Co >= Cr; padded channels are written as zero)
// mode 1: y = gelu(gn(raw))
// mode 2: y = gn(raw)[c] * sigmoid(gn(raw)[c + Co'])  with Co' = Cr / 2 (GLU over channels)
// mode 3: y = raw[2c] * sigmoid(raw[2c + 1])            (GLU over interleaved (value, gate) column pairs: the layout the
//         un-normalised layers' weights are packed in for the fused epilogue; training forward only, stats == nullptr)
// then  y = y * scale[c] (if scale) ; y += res (if res: split (B, Y, Xo, Co)).  stats == nullptr -> identity norm.
__device__ __forceinline__ void ld8f(const float* p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
// VEC: channel counts are multiples of 8, every 8-channel run sits in one group and all vectors are 16-byte aligned
template <bool VEC>
__global__ void __launch_bounds__(256) gn_apply_kernel(const GnApply a) {
  const unsigned groups = a.Co / 8;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (unsigned)a.Y * a.Xo * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const unsigned pix = idx / groups;
  const int x = (int)(pix % a.Xo);
  const int y = (int)(pix / a.Xo);
  const int xr = x + a.x_off;
  const float* r = a.raw + (((size_t)b * a.Y + y) * a.Xr + xr) * a.Cr;
  const int seg = a.per_x ? b * a.Xr + xr : b;
  const int cvalid = a.mode >= 2 ? a.Cr / 2 : a.Cr;
  const int cpg = a.Cr / a.G;
  float o[8];
  if (a.mode == 3) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float v = 0.0f;
      if (c < cvalid) {
        const float2 pr = *reinterpret_cast<const float2*>(r + 2 * c);
        v = pr.x * sigmoidf_fast(pr.y);
        if (a.scale) v *= a.scale[c];
      }
      o[i] = v;
    }
  } else if (VEC) {
    if (c0 < cvalid) {
      ld8f(r + c0, o);
      if (a.stats) {
        const float2 st = *reinterpret_cast<const float2*>(a.stats + ((size_t)seg * a.G + c0 / cpg) * 2);
        float ga[8], be[8];
        ld8f(a.gamma + c0, ga);
        ld8f(a.beta + c0, be);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (o[i] - st.x) * st.y * ga[i] + be[i];
      }
      if (a.mode == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = gelu_exact(o[i]);
      } else if (a.mode == 2) {
        const int c2 = c0 + cvalid;
        float gt[8];
        ld8f(r + c2, gt);
        if (a.stats) {
          const float2 st = *reinterpret_cast<const float2*>(a.stats + ((size_t)seg * a.G + c2 / cpg) * 2);
          float ga[8], be[8];
          ld8f(a.gamma + c2, ga);
          ld8f(a.beta + c2, be);
#pragma unroll
          for (int i = 0; i < 8; ++i) gt[i] = (gt[i] - st.x) * st.y * ga[i] + be[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= sigmoidf_fast(gt[i]);
      }
      if (a.scale) {
        float sc[8];
        ld8f(a.scale + c0, sc);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= sc[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.0f;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      float v = 0.0f;
      if (c < cvalid) {
        float u = r[c];
        if (a.stats) {
          const float* st = a.stats + ((size_t)seg * a.G + c / cpg) * 2;
          u = (u - st[0]) * st[1] * a.gamma[c] + a.beta[c];
        }
        if (a.mode == 1) {
          v = gelu_exact(u);
        } else if (a.mode == 2) {
          const int c2 = c + cvalid;
          float gte = r[c2];
          if (a.stats) {
            const float* st = a.stats + ((size_t)seg * a.G + c2 / cpg) * 2;
            gte = (gte - st[0]) * st[1] * a.gamma[c2] + a.beta[c2];
          }
          v = u * sigmoidf_fast(gte);
        } else {
          v = u;
        }
        if (a.scale) v *= a.scale[c];
      }
      o[i] = v;
    }
  }
  const size_t off = (((size_t)b * a.Y + y) * a.Xo + x) * a.Co + c0;
  if (a.rhi) {
    float rr[8];
    load_split8(a.rhi, a.rlo, off, rr);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += rr[i];
  }
  store_split8(a.ohi, a.olo, off, o);
}

// ---- out = a[(b, y, x + x_off)] + skip[(b, y, x)]  (decoder `x + skip` with the transposed-conv crop folded in) ----
__global__ void __launch_bounds__(256) add_crop_kernel(const __nv_bfloat16* __restrict__ ahi, const __nv_bfloat16* __restrict__ alo, int Ya, int Xa, int x_off,
                                                       const __nv_bfloat16* __restrict__ shi, const __nv_bfloat16* __restrict__ slo,
                                                       __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo, int Y, int X, int C) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)Y * X * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int x = (int)((idx / groups) % X);
  const int y = (int)(idx / ((long long)groups * X));
  float u[8], v[8];
  load_split8(ahi, alo, (((size_t)b * Ya + y) * Xa + x + x_off) * C + c0, u);   // (a may have more rows than the skip: Ya >= Y)
  const size_t off = (((size_t)b * Y + y) * X + x) * C + c0;
  load_split8(shi, slo, off, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] += v[i];
  store_split8(ohi, olo, off, u);
}

// ---- z[b][y][x][c] += w * emb[x][c]  (frequency embedding after freq layer 0, _hdemucs.py:586-591), in place ----
__global__ void __launch_bounds__(256) freq_emb_kernel(__nv_bfloat16* __restrict__ zhi, __nv_bfloat16* __restrict__ zlo, int Y, int X, int C,
                                                       const float* __restrict__ emb /*[X][C]*/, float w) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)Y * X * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int x = (int)((idx / groups) % X);
  const size_t off = ((size_t)b * Y * X) * C + (size_t)(idx / groups) * C + c0;
  float u[8];
  load_split8(zhi, zlo, off, u);
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] += w * emb[(size_t)x * C + c0 + i];
  store_split8(zhi, zlo, off, u);
}

// ---- a + b on fp32 tensors (time-branch injection into freq layer 4, _hdemucs.py:164-169) ----
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}

// ---- last freq decoder: transposed-conv output (fp32 [B][T][Xg][4*2], bias included) -> de-normalised complex Z ----
// bin k of frame t = position k + pad in the (Xg * 4)-long output; channel 0 = re, 1 = im; Z = v * std + mean.
__global__ void __launch_bounds__(256) final_freq_kernel(const float* __restrict__ raw, int Tf, int Xg, int pad, int bins,
                                                         const float* __restrict__ stats, float2* __restrict__ Z) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)Tf * bins) return;
  const int k = (int)(idx % bins), t = (int)(idx / bins);
  const int pos = k + pad;
  const float* r = raw + (((size_t)b * Tf + t) * Xg + pos / 4) * 8 + (pos % 4) * 2;
  const float mean = stats[2 * b], sd = stats[2 * b + 1];
  Z[((size_t)b * Tf + t) * bins + k] = make_float2(r[0] * sd + mean, r[1] * sd + mean);
}

// ---- last time decoder: ConvTranspose1d(C -> 1, k, stride s) + bias, crop [pad, pad + T), de-normalise, add to out ----
// y split [B][L][C]; out[b][t] += (convtr(y)[t + pad]) * std_t + mean_t.  One thread per output sample.
__global__ void __launch_bounds__(256) final_time_kernel(const __nv_bfloat16* __restrict__ yhi, const __nv_bfloat16* __restrict__ ylo, int L, int C,
                                                         int K, int S, int pad, const float* __restrict__ w /*[C][1][K]*/,
                                                         const float* __restrict__ bias, const float* __restrict__ stats, int T,
                                                         float* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const int pos = (int)t + pad;
  float acc = bias[0];
  for (int j = pos % S; j < K; j += S) {
    const int i = (pos - j) / S;
    if (i < 0 || i >= L) continue;
    const size_t off = ((size_t)b * L + i) * C;
    for (int c = 0; c < C; c += 8) {
      float u[8];
      load_split8(yhi, ylo, off + c, u);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(w[(c + e) * K + j], u[e], acc);
    }
  }
  out[(size_t)b * T + t] += acc * stats[2 * b + 1] + stats[2 * b];
}

// ---- _BLSTM framing (TA:758-768): frames of `width` with stride width/2, zero padded; in (B,1,T,C) -> (B*nf,1,width,C) ----
__global__ void __launch_bounds__(256) blstm_frame_kernel(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo, int T, int C,
                                                          int nf, int width, int stride, __nv_bfloat16* __restrict__ ohi,
                                                          __nv_bfloat16* __restrict__ olo) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int bf = blockIdx.y;  // b * nf + k
  if (idx >= (long long)width * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int pos = (int)(idx / groups);
  const int b = bf / nf, k = bf % nf;
  const int t = k * stride + pos;
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (t < T) load_split8(ihi, ilo, ((size_t)b * T + t) * C + c0, o);
  store_split8(ohi, olo, ((size_t)bf * width + pos) * C + c0, o);
}
// ---- _BLSTM stitch + skip (TA:772-788): lin fp32 (B*nf, width, C) -> out split (B,1,T,C) = lin[frame(t)][pos(t)] + skip ----
__global__ void __launch_bounds__(256) blstm_merge_kernel(const float* __restrict__ lin, int T, int C, int nf, int width, int stride,
                                                          const __nv_bfloat16* __restrict__ shi, const __nv_bfloat16* __restrict__ slo,
                                                          __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)T * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int t = (int)(idx / groups);
  int k = 0, pos = t;
  if (nf > 1) {
    k = (t - stride / 2) / stride;
    if (t < stride / 2) k = 0;
    if (k > nf - 1) k = nf - 1;
    pos = t - k * stride;
  }
  const float* r = lin + (((size_t)b * nf + k) * width + pos) * C + c0;
  float o[8];
  load_split8(shi, slo, ((size_t)b * T + t) * C + c0, o);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] += r[i];
  store_split8(ohi, olo, ((size_t)b * T + t) * C + c0, o);
}

// ---- _LocalState attention core (TA:832-857) ----
// qkv fp32 [(b*T + t)][ld]: columns [0,C) queries, [C,2C) keys, [2C,3C) content, [3C, 3C + heads*ndecay) decay logits.
// For head h (channels h*Ch .. +Ch):  dots[t][s] = k_t . q_s / sqrt(Ch) - sum_f (f+1) |t-s| / sqrt(nd) * sigmoid(dq[f][s]) / 2,
// diagonal = -100, softmax over t, result[s][c] = sum_t w[t][s] content[t][c].  Output split [(b*T + s)][C].
// grid = (ceil(T / 32), heads, B), 256 threads; dynamic smem: K[T][Ch+1] + V[T][Ch+1] + W[T][33] + Q[32][Ch+1].
constexpr int LA_QT = 32;
__global__ void __launch_bounds__(256) local_attn_kernel(const float* __restrict__ qkv, int ld, int T, int C, int heads, int nd,
                                                         __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  extern __shared__ float la_smem[];
  const int Ch = C / heads, Chp = Ch + 1;
  float* Ks = la_smem;                 // [T][Chp]
  float* Vs = Ks + (size_t)T * Chp;    // [T][Chp]
  float* Ws = Vs + (size_t)T * Chp;    // [T][LA_QT + 1]
  float* Qs = Ws + (size_t)T * (LA_QT + 1);  // [LA_QT][Chp]
  float* Ds = Qs + LA_QT * Chp;        // [LA_QT]: sum_f (f+1) * sigmoid(dq)/2 / sqrt(nd)
  const int tid = threadIdx.x;
  const int hd = blockIdx.y, b = blockIdx.z;
  const int s0 = blockIdx.x * LA_QT;
  const float* base = qkv + (size_t)b * T * ld;
  for (int i = tid; i < T * Ch; i += 256) {
    const int t = i / Ch, c = i % Ch;
    Ks[t * Chp + c] = base[(size_t)t * ld + C + hd * Ch + c];
    Vs[t * Chp + c] = base[(size_t)t * ld + 2 * C + hd * Ch + c];
  }
  for (int i = tid; i < LA_QT * Ch; i += 256) {
    const int sl = i / Ch, c = i % Ch;
    Qs[sl * Chp + c] = (s0 + sl < T) ? base[(size_t)(s0 + sl) * ld + hd * Ch + c] : 0.0f;
  }
  if (tid < LA_QT) {
    float d = 0.0f;
    if (s0 + tid < T)
      for (int f = 0; f < nd; ++f) d += (float)(f + 1) * (sigmoidf_acc(base[(size_t)(s0 + tid) * ld + 3 * C + hd * nd + f]) * 0.5f);
    Ds[tid] = d / sqrtf((float)nd);
  }
  __syncthreads();
  const float inv = 1.0f / sqrtf((float)Ch);
  for (int i = tid; i < T * LA_QT; i += 256) {
    const int t = i / LA_QT, sl = i % LA_QT;
    const int s = s0 + sl;
    float acc = 0.0f;
    for (int c = 0; c < Ch; ++c) acc = fmaf(Ks[t * Chp + c], Qs[sl * Chp + c], acc);
    float v = acc * inv - fabsf((float)(t - s)) * Ds[sl];
    if (t == s) v = -100.0f;
    Ws[t * (LA_QT + 1) + sl] = v;
  }
  __syncthreads();
  {  // softmax over t for each query: one warp handles 4 queries
    const int warp = tid >> 5, lane = tid & 31;
    for (int sl = warp * 4; sl < warp * 4 + 4; ++sl) {
      float mx = -INFINITY;
      for (int t = lane; t < T; t += 32) mx = fmaxf(mx, Ws[t * (LA_QT + 1) + sl]);
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.0f;
      for (int t = lane; t < T; t += 32) {
        const float e = expf(Ws[t * (LA_QT + 1) + sl] - mx);
        Ws[t * (LA_QT + 1) + sl] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float r = 1.0f / sum;
      for (int t = lane; t < T; t += 32) Ws[t * (LA_QT + 1) + sl] *= r;
    }
  }
  __syncthreads();
  for (int i = tid; i < LA_QT * Ch; i += 256) {
    const int sl = i / Ch, c = i % Ch;
    const int s = s0 + sl;
    if (s >= T) continue;
    float acc = 0.0f;
    for (int t = 0; t < T; ++t) acc = fmaf(Ws[t * (LA_QT + 1) + sl], Vs[t * Chp + c], acc);
    __nv_bfloat16 hh, ll;
    split_bf16(acc, hh, ll);
    const size_t off = ((size_t)b * T + s) * C + hd * Ch + c;
    ohi[off] = hh;
    olo[off] = ll;
  }
}

// ---- weight gather for the implicit-GEMM convolutions (see hdemucs.cu: ConvSpec) ----
__global__ void gather_w_kernel(const float* __restrict__ w, const float* __restrict__ bias, GatherSpec g, float* __restrict__ wcat,
                                float* __restrict__ bcat) {
  const long long total = (long long)g.Nout * g.taps * g.Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % g.Kp);
    const int tap = (int)((i / g.Kp) % g.taps);
    const int n = (int)(i / ((long long)g.Kp * g.taps));
    float v = 0.0f;
    if (g.kind == 0) {
      int co = n;
      if (g.glu) co = (n & 1) ? (g.Co / 2 + n / 2) : n / 2;
      if (col < g.Ci) v = w[((size_t)co * g.Ci + col) * g.k + tap];
    } else if (g.kind == 1) {
      int co = n;
      const int r = col / g.Ci, ci = col % g.Ci;
      const int j = g.s * (g.tau_min + tap) + r + g.p;
      if (col < g.s * g.Ci && j >= 0 && j < g.k) v = w[((size_t)co * g.Ci + ci) * g.k + j];
    } else {
      const int r = n / g.Co, co = n % g.Co;  // weight [Ci][Co][k]
      const int j = r + g.s * tap;
      if (col < g.Ci && j < g.k) v = w[((size_t)col * g.Co + co) * g.k + j];
    }
    wcat[i] = v;
    if (tap == 0 && col == 0 && bcat) {
      float bv = 0.0f;
      if (bias) {
        if (g.kind == 2) bv = bias[n % g.Co];
        else if (g.glu) bv = bias[(n & 1) ? (g.Co / 2 + n / 2) : n / 2];
        else bv = bias[n];
      }
      bcat[n] = bv;
    }
  }
}

}  // namespace hd
}  // namespace rfx
