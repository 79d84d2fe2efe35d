// Backward kernels of the Hybrid-Demucs training step (torchaudio/models/_hdemucs.py under torch autograd, as the reference's
// Lightning step differentiates it: remfx/models.py:217-220 -> loss.backward(); cfg/exp/5-5_full.yaml:3).  Included by
// hdemucs_bwd.cu only.  Conventions (see hd_internal.h): activations are channel-last (B, Y, X, C); the gradient of a split-bf16
// ACTIVATION is an fp32 tensor of the same shape; the gradient of an fp32 PRE-ACTIVATION (a conv output) is written as split
// planes (B, Y, X, Cg = ceil8(C)) because it is the A operand of the input-gradient GEMM and of the weight-gradient contraction.
// tools/hd_bwd_emul.py holds the same formulas in fp64 against autograd.
#pragma once
#include "hd_internal.h"

namespace rfx {
namespace hd {

__device__ __forceinline__ void bw_store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t off, const float (&o)[8]) {
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
__device__ __forceinline__ void bw_load_split8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&o)[8]) {
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
// d/du gelu(u), erf form:  Phi(u) + u phi(u)
__device__ __forceinline__ float gelu_grad(float u) {
  const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * u * u);
  return fmaf(u, pdf, cdf);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (optional) + activation (+ LayerScale, + residual, + crop) backward.  Forward: gn_apply_kernel (hd_kernels.cuh).
//   u = stats ? xhat * gamma + beta : raw,  xhat = (raw - mean) * rstd;   v = act(u);   out = v * scale + res
//   PASS 1 (only when stats != nullptr): per-channel sums  dbeta += du, dgamma += du xhat, dscale += dout v   and the two
//          per-(segment, group) sums  S1 = sum(gamma du), S2 = sum(gamma du xhat)  (fp64 atomics)
//   PASS 2: d raw = stats ? rstd (gamma du - S1 / n - xhat S2 / n) : du  -> split planes;  d res (+)= dout
// Work split: items = (segment, pixel chunk), a 1-D grid of at most four CTAs per SM strides over them; a thread owns one octet of
// OUTPUT channels and strides over the item's pixels, so per-channel sums stay in registers for the CTA's whole life (the
// tcn_act_bwd_kernel pattern) and reach global memory once per CTA.  The pixel domain is the whole UNcropped raw extent:
// positions outside the crop window have dout = 0 but, under GroupNorm, a non-zero d raw.
// ------------------------------------------------------------------------------------------------
struct GnBwd {
  GnApply a;              // the forward arguments
  const float* dy;        // gradient of out, fp32 (B, Y, Xo, Co)
  float* dres;            // gradient of res (same shape as out) or nullptr
  int res_accum;          // 1: dres += dout, 0: dres = dout
  __nv_bfloat16* ghi;     // d raw planes (B, Y, Xr, Cg)
  __nv_bfloat16* glo;
  int Cg;
  double* gsum;           // [segments * G][2]
  float* dgamma; float* dbeta; float* dscale;
  long long count;        // elements per (segment, group)
  int rows_per_cta;       // pixels per work item
  int nseg, nchunks;      // work items = nseg * nchunks; the grid strides over them
};

template <int PASS, int MODE>   // MODE = GnApply::mode, a compile-time constant: the plain / GELU forms carry none of the gate octet's state
__global__ void __launch_bounds__(256, ((PASS == 2 && MODE != 2) || (PASS == 1 && MODE <= 1)) ? 3 : 2) gn_bwd_kernel(const GnBwd p) {
  extern __shared__ __align__(16) float gb_smem[];  // PASS 1: [2 * Cr] (dbeta, dgamma) + [Co] (dscale), then doubles [2 * G] (16-byte aligned)
  const GnApply& a = p.a;
  const int groups = a.Co / 8;
  const int rows = 256 / groups;
  const int g8 = threadIdx.x % groups, r = threadIdx.x / groups;
  const int c0 = g8 * 8;
  constexpr int mode = MODE;
  const int Ch = mode >= 2 ? a.Cr / 2 : a.Cr;  // valid output channels
  const int cpg = a.Cr / a.G;
  const bool has_stats = a.stats != nullptr;
  const long long npix = a.per_x ? a.Y : (long long)a.Y * a.Xr;
  double* sm_d = nullptr;
  if (PASS == 1) {
    const int nf = 2 * a.Cr + a.Co;
    for (int i = threadIdx.x; i < nf; i += 256) gb_smem[i] = 0.0f;
    sm_d = reinterpret_cast<double*>(gb_smem + ((nf + 3) & ~3));
    if (threadIdx.x < 2 * a.G) sm_d[threadIdx.x] = 0.0;
    __syncthreads();
  }
  // per-thread constants: affine parameters of the value octet and (GLU) the gate octet
  float ga[8], be[8], ga2[8], be2[8], sc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    ga[i] = 1.0f; be[i] = 0.0f; ga2[i] = 1.0f; be2[i] = 0.0f; sc[i] = 1.0f;
    if (c < Ch) {
      if (has_stats) {
        ga[i] = a.gamma[c]; be[i] = a.beta[c];
        if (mode == 2) { ga2[i] = a.gamma[c + Ch]; be2[i] = a.beta[c + Ch]; }
      }
      if (a.scale) sc[i] = a.scale[c];
    }
  }
  const int grp1 = has_stats ? min(c0, a.Cr - 1) / cpg : 0;
  const int grp2 = (has_stats && mode == 2) ? min(c0 + Ch, a.Cr - 1) / cpg : 0;
  float acc_b[8] = {}, acc_g[8] = {}, acc_b2[8] = {}, acc_g2[8] = {}, acc_s[8] = {};  // PASS 1: per-channel sums over ALL of this CTA's work
  const bool active = r < rows && c0 < a.Co;
  // whole, 16-byte aligned octets: Cr % 4 == 0 keeps every pixel row aligned, Ch % 8 == 0 keeps the octets whole
  const bool vec_ok = (a.Cr % 4 == 0) && (Ch % 8 == 0) && ((reinterpret_cast<uintptr_t>(a.raw) & 15) == 0);
  // work items = (segment, pixel chunk); a CTA strides over them so that the per-channel sums are flushed once per CTA, not per item
  const long long n_items = (long long)p.nseg * p.nchunks;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seg = (int)(item / p.nchunks), chunk = (int)(item % p.nchunks);
    const int b = a.per_x ? seg / a.Xr : seg;
    const int xfix = a.per_x ? seg % a.Xr : 0;
    float mean1 = 0.0f, rstd1 = 1.0f, mean2 = 0.0f, rstd2 = 1.0f;
    double s1a = 0.0, s2a = 0.0, s1b = 0.0, s2b = 0.0;   // PASS 1: this item's group sums; PASS 2: the segment's means
    if (has_stats) {
      const float* st = a.stats + ((size_t)seg * a.G + grp1) * 2;
      mean1 = st[0]; rstd1 = st[1];
      if (mode == 2) { const float* st2 = a.stats + ((size_t)seg * a.G + grp2) * 2; mean2 = st2[0]; rstd2 = st2[1]; }
      if (PASS == 2) {
        s1a = p.gsum[((size_t)seg * a.G + grp1) * 2] / (double)p.count;
        s2a = p.gsum[((size_t)seg * a.G + grp1) * 2 + 1] / (double)p.count;
        if (mode == 2) {
          s1b = p.gsum[((size_t)seg * a.G + grp2) * 2] / (double)p.count;
          s2b = p.gsum[((size_t)seg * a.G + grp2) * 2 + 1] / (double)p.count;
        }
      }
    }
    const float m1a = (float)s1a, m2a = (float)s2a, m1b = (float)s1b, m2b = (float)s2b;
    if (PASS == 1) { s1a = s2a = s1b = s2b = 0.0; }
    const long long p_begin = (long long)chunk * p.rows_per_cta;
    const long long p_end = min(npix, p_begin + p.rows_per_cta);
    if (active) {
      for (long long pp = p_begin + r; pp < p_end; pp += rows) {
        int y, xr;
        if (a.per_x) { y = (int)pp; xr = xfix; }
        else { y = (int)(pp / a.Xr); xr = (int)(pp % a.Xr); }
        const int xo = xr - a.x_off;
        const bool inwin = xo >= 0 && xo < a.Xo;
        const float* rawp = a.raw + (((size_t)b * a.Y + y) * a.Xr + xr) * a.Cr;
        float dout[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dout[i] = 0.0f;
        if (inwin) {
          const size_t ooff = (((size_t)b * a.Y + y) * a.Xo + xo) * a.Co + c0;
          const float4 d0 = *reinterpret_cast<const float4*>(p.dy + ooff), d1 = *reinterpret_cast<const float4*>(p.dy + ooff + 4);
          dout[0] = d0.x; dout[1] = d0.y; dout[2] = d0.z; dout[3] = d0.w; dout[4] = d1.x; dout[5] = d1.y; dout[6] = d1.z; dout[7] = d1.w;
          if (PASS == 2 && p.dres) {
            float4 r0 = d0, r1 = d1;
            if (p.res_accum) {
              const float4 o0 = *reinterpret_cast<const float4*>(p.dres + ooff), o1 = *reinterpret_cast<const float4*>(p.dres + ooff + 4);
              r0.x += o0.x; r0.y += o0.y; r0.z += o0.z; r0.w += o0.w; r1.x += o1.x; r1.y += o1.y; r1.z += o1.z; r1.w += o1.w;
            }
            *reinterpret_cast<float4*>(p.dres + ooff) = r0;
            *reinterpret_cast<float4*>(p.dres + ooff + 4) = r1;
          }
        }
        float dr1[8], dr2[8];   // d raw of the value octet / the gate octet
        float rv1[8], rv2[8];   // raw values of the two octets (16-byte loads when the octet is whole and aligned)
        if (vec_ok) {
          if (mode == 3) {
            const float4 q0 = *reinterpret_cast<const float4*>(rawp + 2 * c0), q1 = *reinterpret_cast<const float4*>(rawp + 2 * c0 + 4);
            const float4 q2 = *reinterpret_cast<const float4*>(rawp + 2 * c0 + 8), q3 = *reinterpret_cast<const float4*>(rawp + 2 * c0 + 12);
            rv1[0] = q0.x; rv2[0] = q0.y; rv1[1] = q0.z; rv2[1] = q0.w; rv1[2] = q1.x; rv2[2] = q1.y; rv1[3] = q1.z; rv2[3] = q1.w;
            rv1[4] = q2.x; rv2[4] = q2.y; rv1[5] = q2.z; rv2[5] = q2.w; rv1[6] = q3.x; rv2[6] = q3.y; rv1[7] = q3.z; rv2[7] = q3.w;
          } else {
            const float4 q0 = *reinterpret_cast<const float4*>(rawp + c0), q1 = *reinterpret_cast<const float4*>(rawp + c0 + 4);
            rv1[0] = q0.x; rv1[1] = q0.y; rv1[2] = q0.z; rv1[3] = q0.w; rv1[4] = q1.x; rv1[5] = q1.y; rv1[6] = q1.z; rv1[7] = q1.w;
            if (mode == 2) {
              const float4 g0 = *reinterpret_cast<const float4*>(rawp + Ch + c0), g1 = *reinterpret_cast<const float4*>(rawp + Ch + c0 + 4);
              rv2[0] = g0.x; rv2[1] = g0.y; rv2[2] = g0.z; rv2[3] = g0.w; rv2[4] = g1.x; rv2[5] = g1.y; rv2[6] = g1.z; rv2[7] = g1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) rv2[i] = 0.0f;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            rv1[i] = 0.0f; rv2[i] = 0.0f;
            if (c >= Ch) continue;
            if (mode == 3) { rv1[i] = rawp[2 * c]; rv2[i] = rawp[2 * c + 1]; }
            else { rv1[i] = rawp[c]; if (mode == 2) rv2[i] = rawp[c + Ch]; }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = c0 + i;
          dr1[i] = 0.0f; dr2[i] = 0.0f;
          if (c >= Ch) continue;
          const float raw1 = rv1[i], raw2 = rv2[i];
          const float xh1 = (raw1 - mean1) * rstd1, xh2 = (raw2 - mean2) * rstd2;
          const float u1 = has_stats ? fmaf(xh1, ga[i], be[i]) : raw1;
          const float u2 = has_stats ? fmaf(xh2, ga2[i], be2[i]) : raw2;
          const float dv = dout[i] * sc[i];
          float du1, du2 = 0.0f, v;
          if (mode == 0) { du1 = dv; v = u1; }
          else if (mode == 1) { du1 = dv * gelu_grad(u1); v = gelu_fast(u1); }
          else {
            const float s = sigmoidf_acc(u2);
            du1 = dv * s;
            du2 = dv * u1 * s * (1.0f - s);
            v = u1 * s;
          }
          if (PASS == 1) {
            acc_b[i] += du1; acc_g[i] += du1 * xh1; acc_s[i] += dout[i] * v;
            s1a += (double)(ga[i] * du1); s2a += (double)(ga[i] * du1 * xh1);
            if (mode == 2) {
              acc_b2[i] += du2; acc_g2[i] += du2 * xh2;
              s1b += (double)(ga2[i] * du2); s2b += (double)(ga2[i] * du2 * xh2);
            }
          } else {
            dr1[i] = has_stats ? rstd1 * (ga[i] * du1 - m1a - xh1 * m2a) : du1;
            if (mode >= 2) dr2[i] = has_stats ? rstd2 * (ga2[i] * du2 - m1b - xh2 * m2b) : du2;
          }
        }
        if (PASS == 2) {
          const size_t gpix = (((size_t)b * a.Y + y) * a.Xr + xr) * p.Cg;
          if (mode == 3) {  // interleaved (value, gate) pairs: raw channels 2 c0 .. 2 c0 + 15
            if (c0 < Ch) {
              float lo8[8], hi8[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) { lo8[2 * i] = dr1[i]; lo8[2 * i + 1] = dr2[i]; hi8[2 * i] = dr1[4 + i]; hi8[2 * i + 1] = dr2[4 + i]; }
              bw_store_split8(p.ghi, p.glo, gpix + 2 * c0, lo8);
              bw_store_split8(p.ghi, p.glo, gpix + 2 * c0 + 8, hi8);
            }
          } else {
            if (c0 < p.Cg) bw_store_split8(p.ghi, p.glo, gpix + c0, dr1);   // channels in [Cr, Cg) are written as zero
            if (mode == 2 && c0 < Ch) bw_store_split8(p.ghi, p.glo, gpix + Ch + c0, dr2);
          }
        }
      }
    }
    if (PASS == 1 && p.gsum) {   // this item's group sums -> the segment's accumulators
      // warp-reduce per group first: shared-memory fp64 atomics are CAS loops, and 250 threads on two addresses serialise badly
      for (int g = 0; g < a.G; ++g) {
        double v1 = 0.0, v2 = 0.0;
        if (active) {
          if (grp1 == g) { v1 += s1a; v2 += s2a; }
          if (mode == 2 && grp2 == g) { v1 += s1b; v2 += s2b; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if ((threadIdx.x & 31) == 0 && (v1 != 0.0 || v2 != 0.0)) { atomicAdd(&sm_d[2 * g], v1); atomicAdd(&sm_d[2 * g + 1], v2); }
      }
      __syncthreads();
      if (threadIdx.x < 2 * a.G) {
        atomicAdd(p.gsum + (size_t)seg * a.G * 2 + threadIdx.x, sm_d[threadIdx.x]);
        sm_d[threadIdx.x] = 0.0;
      }
      __syncthreads();
    }
  }
  if (PASS == 1) {
    float* sm_b = gb_smem;             // [Cr] dbeta
    float* sm_g = gb_smem + a.Cr;      // [Cr] dgamma
    float* sm_s = gb_smem + 2 * a.Cr;  // [Co] dscale
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c0 + i;
        if (c >= Ch) continue;
        atomicAdd(&sm_b[c], acc_b[i]); atomicAdd(&sm_g[c], acc_g[i]); atomicAdd(&sm_s[c], acc_s[i]);
        if (mode == 2) { atomicAdd(&sm_b[c + Ch], acc_b2[i]); atomicAdd(&sm_g[c + Ch], acc_g2[i]); }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < a.Cr; i += 256) {
      if (p.dbeta) atomicAdd(p.dbeta + i, sm_b[i]);
      if (p.dgamma) atomicAdd(p.dgamma + i, sm_g[i]);
    }
    if (p.dscale)
      for (int i = threadIdx.x; i < Ch; i += 256) atomicAdd(p.dscale + i, sm_s[i]);
  }
}

// ---- out = a[x + x_off] + skip  backward:  d a = (dout inside the crop window, 0 outside);  d skip (+)= dout ----
__global__ void __launch_bounds__(256) addcrop_bwd_kernel(const float* __restrict__ dy, int Y, int X, int C, float* __restrict__ da, int Xa, int x_off,
                                                          float* __restrict__ dskip, int skip_accum) {
  const int groups = C / 4;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)Y * Xa * groups) return;
  const int c0 = (int)(idx % groups) * 4;
  const int xa = (int)((idx / groups) % Xa);
  const int y = (int)(idx / ((long long)groups * Xa));
  const int x = xa - x_off;
  float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x >= 0 && x < X) {
    const size_t off = (((size_t)b * Y + y) * X + x) * C + c0;
    g = *reinterpret_cast<const float4*>(dy + off);
    if (dskip) {
      float4 o = g;
      if (skip_accum) { const float4 q = *reinterpret_cast<const float4*>(dskip + off); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
      *reinterpret_cast<float4*>(dskip + off) = o;
    }
  }
  *reinterpret_cast<float4*>(da + (((size_t)b * Y + y) * Xa + xa) * C + c0) = g;
}

// ---- dst (+)= src, fp32 ----
__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}

// ---- frequency embedding backward:  demb[x][c] += w * sum_{b, y} dz[b][y][x][c]  (TA:586-591; z itself passes through) ----
__global__ void __launch_bounds__(256) freqemb_bwd_kernel(const float* __restrict__ dz, int B, int Y, int X, int C, float w, float* __restrict__ demb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (x, c)
  if (i >= X * C) return;
  const int yb = blockIdx.y, ny = gridDim.y;             // slices of the (b, y) range
  const long long rows = (long long)B * Y;
  float acc = 0.0f;
  for (long long ry = yb; ry < rows; ry += ny) acc += dz[(size_t)ry * X * C + i];
  atomicAdd(demb + i, w * acc);
}

// ---- _BLSTM framing backward (adjoint of _unfold, TA:888-905): dx[b][t][c] (+)= sum over the frames covering t ----
__global__ void __launch_bounds__(256) frame_bwd_kernel(const float* __restrict__ dfr /*(B*nf, width, C)*/, int T, int C, int nf, int width, int stride,
                                                        float* __restrict__ dx, int accum) {
  const int groups = C / 4;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)T * groups) return;
  const int c0 = (int)(idx % groups) * 4;
  const int t = (int)(idx / groups);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int k_hi = min(nf - 1, t / stride);
  for (int k = k_hi; k >= 0 && t - k * stride < width; --k) {
    const float4 v = *reinterpret_cast<const float4*>(dfr + (((size_t)b * nf + k) * width + (t - k * stride)) * C + c0);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float* o = dx + ((size_t)b * T + t) * C + c0;
  if (accum) { const float4 q = *reinterpret_cast<const float4*>(o); acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w; }
  *reinterpret_cast<float4*>(o) = acc;
}

// ---- _BLSTM stitch + skip backward (TA:772-788): d lin[(b, k), pos][c] = dout[b][t][c] where frame(t) == k, else 0 -> split planes;
//      d skip (+)= dout.  One thread per (frame row, 8 channels) of lin. ----
__global__ void __launch_bounds__(256) merge_bwd_kernel(const float* __restrict__ dy /*(B, T, C)*/, int T, int C, int nf, int width, int stride,
                                                        __nv_bfloat16* __restrict__ ghi, __nv_bfloat16* __restrict__ glo) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int bf = blockIdx.y;  // b * nf + k
  if (idx >= (long long)width * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int pos = (int)(idx / groups);
  const int b = bf / nf, k = bf % nf;
  const int t = k * stride + pos;
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (t < T) {
    int ksel = 0;
    if (nf > 1) {
      ksel = (t - stride / 2) / stride;
      if (t < stride / 2) ksel = 0;
      if (ksel > nf - 1) ksel = nf - 1;
    }
    if (ksel == k) {
      const float* s = dy + ((size_t)b * T + t) * C + c0;
      const float4 d0 = *reinterpret_cast<const float4*>(s), d1 = *reinterpret_cast<const float4*>(s + 4);
      o[0] = d0.x; o[1] = d0.y; o[2] = d0.z; o[3] = d0.w; o[4] = d1.x; o[5] = d1.y; o[6] = d1.z; o[7] = d1.w;
    }
  }
  bw_store_split8(ghi, glo, ((size_t)bf * width + pos) * C + c0, o);
}

// ---- fp32 (rows, cols) -> split planes (rows, cols_pad) with zero padding columns ----
__global__ void __launch_bounds__(256) split_pad_kernel(const float* __restrict__ src, long long rows, int cols, int cols_pad,
                                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int groups = cols_pad / 8;
  const long long total = rows * groups;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long rr = i / groups;
    const int c0 = (int)(i % groups) * 8;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (c0 + k < cols) ? src[(size_t)rr * cols + c0 + k] : 0.0f;
    bw_store_split8(hi, lo, (size_t)rr * cols_pad + c0, o);
  }
}

// ---- column sums of split planes: out[n] += sum_rows G[row][col0 + n]   (bias gradients); grid = (row chunks, column blocks of 2048) ----
__global__ void __launch_bounds__(256) colsum_split_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, long long rows, int ld,
                                                           int col0, int N, float* __restrict__ out, int rows_per_cta) {
  __shared__ float cs_smem[2048];
  const int cb = blockIdx.y * 2048;
  const int nloc = min(2048, N - cb);
  const int groups = (nloc + 7) / 8;
  const int nrow = 256 / groups;
  const int g8 = threadIdx.x % groups, r = threadIdx.x / groups;
  for (int i = threadIdx.x; i < groups * 8; i += 256) cs_smem[i] = 0.0f;
  __syncthreads();
  float acc[8] = {};
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  if (r < nrow) {
    for (long long rr = r0 + r; rr < r1; rr += nrow) {
      float v[8];
      bw_load_split8(hi, lo, (size_t)rr * ld + col0 + cb + g8 * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&cs_smem[g8 * 8 + i], acc[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nloc; i += 256) atomicAdd(out + cb + i, cs_smem[i]);
}
// staged column sums [Nout] -> the bias parameter's gradient (inverse of gather_w_kernel's bias re-ordering)
__global__ void scatter_bias_kernel(const float* __restrict__ stage, GatherSpec g, float* __restrict__ db, float* __restrict__ db2) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= g.Nout) return;
  int idx;
  if (g.kind == 2) idx = n % g.Co;
  else if (g.glu) idx = (n & 1) ? (g.Co / 2 + n / 2) : n / 2;
  else idx = n;
  atomicAdd(db + idx, stage[n]);
  if (db2) atomicAdd(db2 + idx, stage[n]);
}

// ---- weight-gradient staging [Nout][taps][Kp] -> the parameter's layout (inverse of gather_w_kernel; every parameter element
//      has exactly one staging slot, so this WRITES) ----
__global__ void scatter_w_kernel(const float* __restrict__ stage, GatherSpec g, float* __restrict__ dw) {
  const long long total = (long long)g.Nout * g.taps * g.Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % g.Kp);
    const int tap = (int)((i / g.Kp) % g.taps);
    const int n = (int)(i / ((long long)g.Kp * g.taps));
    if (g.kind == 0) {
      int co = n;
      if (g.glu) co = (n & 1) ? (g.Co / 2 + n / 2) : n / 2;
      if (col < g.Ci) dw[((size_t)co * g.Ci + col) * g.k + tap] = stage[i];
    } else if (g.kind == 1) {
      const int rr = col / g.Ci, ci = col % g.Ci;
      const int j = g.s * (g.tau_min + tap) + rr + g.p;
      if (col < g.s * g.Ci && j >= 0 && j < g.k) dw[((size_t)n * g.Ci + ci) * g.k + j] = stage[i];
    } else {
      const int rr = n / g.Co, co = n % g.Co;  // weight [Ci][Co][k]
      const int j = rr + g.s * tap;
      if (col < g.Ci && j < g.k) dw[((size_t)col * g.Co + co) * g.k + j] = stage[i];
    }
  }
}

// ---- transposed weights for the input gradient:  Wt[k][tap * Np + n] = Wcat[n][tap * Kp + k]   (k < Kt rows, Np = ceil64(Nout)) ----
__global__ void transpose_w_kernel(const float* __restrict__ wcat, int Nout, int taps, int Kp, int Kt, int Np, float* __restrict__ wt) {
  const long long total = (long long)Kt * taps * Np;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % Np);
    const int tap = (int)((i / Np) % taps);
    const int k = (int)(i / ((long long)Np * taps));
    wt[i] = (n < Nout && k < Kp) ? wcat[((size_t)n * taps + tap) * Kp + k] : 0.0f;
  }
}


// ------------------------------------------------------------------------------------------------
// Weight gradient of an implicit-GEMM convolution (the generalisation of tcn_wgrad_kernel, tcn_bwd.cu):
//   dW[n][tap][k] += sum_{b, y, x} G[b, y, x, gcol0 + n] * A[b, y + dy[tap], x + dx[tap], k]      (A reads outside its extent are 0)
// The contraction runs over PIXELS and both operands are pixel-major in HBM ([pixel][channel]), i.e. MN-major for the MMA:
// fragments come out of ldmatrix.trans; mma.sync.m16n8k16 bf16x3 (lo*hi + hi*lo + hi*hi) with fp32 accumulation; a 3-stage
// cp.async ring of 32 pixels x 4 plane tiles; one CTA = (tap, n x k tile, item, pixel chunk); fp32 atomics into the
// [N][taps][Kp] staging buffer.
// ------------------------------------------------------------------------------------------------
constexpr int HW_BK = 32, HW_STAGES = 3;
// CTA tile = (WM * MT * 16) output columns n  x  (WN * NT * 8) input channels k, 8 warps as WM x WN; the small-channel layers of the
// network (N or K of 12 .. 96) would waste most of a 128 x 128 tile, so the launcher picks the variant with the least padding.
template <int WM, int WN, int MT, int NT>
struct HwCfg {
  static constexpr int BM = WM * MT * 16, BN = WN * NT * 8;
  static constexpr int LDG = BM + 8, LDA = BN + 8;          // padded rows: conflict-free ldmatrix
  static constexpr int G_TILE = HW_BK * LDG, A_TILE = HW_BK * LDA;
  static constexpr int STAGE_ELEMS = 2 * G_TILE + 2 * A_TILE;  // G hi, G lo, A hi, A lo
  static constexpr int SMEM = HW_STAGES * STAGE_ELEMS * 2;
};

struct WgP {
  const __nv_bfloat16* g; long long g_bs, g_ldy, g_ld, g_plane; int gcol0;
  const __nv_bfloat16* a; long long a_bs, a_ldy, a_ld, a_plane;
  int Y, X, Ay, Ax;
  int N, K, taps, Kp;
  int dx[16], dy[16];
  int nchunks, rchunk;
  float* dW;
};

__device__ __forceinline__ void hw_cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void hw_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void hw_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void hw_ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hw_mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int WM, int WN, int MT, int NT>
__global__ void __launch_bounds__(256, 2) hd_wgrad_kernel(const WgP p) {
  using Cfg = HwCfg<WM, WN, MT, NT>;
  static_assert(WM * WN == 8 && NT % 2 == 0, "8 warps; B fragments come in pairs of n-tiles");
  extern __shared__ __align__(16) unsigned char hw_smem[];
  __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(hw_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tilesN = (p.N + Cfg::BM - 1) / Cfg::BM, tilesK = (p.K + Cfg::BN - 1) / Cfg::BN;
  int bx = blockIdx.x;
  const int tk = bx % tilesK; bx /= tilesK;
  const int tn = bx % tilesN;
  const int tap = bx / tilesN;
  const int b = blockIdx.y / p.nchunks, chunk = blockIdx.y % p.nchunks;
  const long long R = (long long)p.Y * p.X;
  const long long r_begin = (long long)chunk * p.rchunk, r_end = min(R, r_begin + p.rchunk);
  const int n0 = tn * Cfg::BM, k0 = tk * Cfg::BN;
  const int ddx = p.dx[tap], ddy = p.dy[tap];
  const __nv_bfloat16* gsrc = p.g + (size_t)b * p.g_bs + p.gcol0;
  const __nv_bfloat16* asrc = p.a + (size_t)b * p.a_bs;
  const int iters = (int)((r_end - r_begin + HW_BK - 1) / HW_BK);

  auto load_stage = [&](int it, int stage) {
    const long long rr0 = r_begin + (long long)it * HW_BK;
    __nv_bfloat16* st = sm + (size_t)stage * Cfg::STAGE_ELEMS;
    constexpr int GP = Cfg::BM / 8, AP = Cfg::BN / 8;  // 16-byte pieces per row
    for (int idx = tid; idx < HW_BK * GP; idx += 256) {
      const int rr = idx / GP, cc = (idx % GP) * 8;
      const long long pix = rr0 + rr;
      const bool gok = pix < r_end && (n0 + cc < p.N);
      const int y = gok ? (int)(pix / p.X) : 0, x = gok ? (int)(pix % p.X) : 0;
      const __nv_bfloat16* gp = gok ? gsrc + (size_t)y * p.g_ldy + (size_t)x * p.g_ld + n0 + cc : p.g;
      const uint32_t d = smem_u32(st + rr * Cfg::LDG + cc);
      hw_cp_async16(d, gp, gok);
      hw_cp_async16(d + Cfg::G_TILE * 2, gok ? gp + p.g_plane : p.g, gok);
    }
    for (int idx = tid; idx < HW_BK * AP; idx += 256) {
      const int rr = idx / AP, cc = (idx % AP) * 8;
      const long long pix = rr0 + rr;
      const bool row_ok = pix < r_end && (k0 + cc < p.K);
      const int y = row_ok ? (int)(pix / p.X) : 0, x = row_ok ? (int)(pix % p.X) : 0;
      const int ya = y + ddy, xa = x + ddx;
      const bool aok = row_ok && ya >= 0 && ya < p.Ay && xa >= 0 && xa < p.Ax;
      const __nv_bfloat16* ap = aok ? asrc + (size_t)ya * p.a_ldy + (size_t)xa * p.a_ld + k0 + cc : p.a;
      const uint32_t d = smem_u32(st + 2 * Cfg::G_TILE + rr * Cfg::LDA + cc);
      hw_cp_async16(d, ap, aok);
      hw_cp_async16(d + Cfg::A_TILE * 2, aok ? ap + p.a_plane : p.a, aok);
    }
  };

  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

  for (int s = 0; s < HW_STAGES - 1; ++s) {
    if (s < iters) load_stage(s, s);
    hw_cp_async_commit();
  }
  const int wm = warp / WN, wn = warp % WN;
  const int lj = lane >> 3, lr = lane & 7;
  // ldmatrix.trans source rows are pixels.  A-operand (n x pixel) fragment order: (m lo, k lo), (m hi, k lo), (m lo, k hi), (m hi, k hi);
  // B-operand (pixel x k) fragments for two n-tiles: (k lo, n0), (k hi, n0), (k lo, n1), (k hi, n1).
  const int a_row = (lj >> 1) * 8 + lr, a_col = wm * (MT * 16) + (lj & 1) * 8;
  const int b_row = (lj & 1) * 8 + lr, b_col = wn * (NT * 8) + (lj >> 1) * 8;

  for (int it = 0; it < iters; ++it) {
    hw_cp_async_wait<HW_STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + HW_STAGES - 1;
      if (nx < iters) load_stage(nx, nx % HW_STAGES);
      hw_cp_async_commit();
    }
    const __nv_bfloat16* st = sm + (size_t)(it % HW_STAGES) * Cfg::STAGE_ELEMS;
    const uint32_t g_hi = smem_u32(st), g_lo = g_hi + Cfg::G_TILE * 2, x_hi = g_hi + 2 * Cfg::G_TILE * 2, x_lo = x_hi + Cfg::A_TILE * 2;
#pragma unroll
    for (int kk = 0; kk < HW_BK; kk += 16) {
      uint32_t bh[NT / 2][4], bl[NT / 2][4];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        const uint32_t o = (uint32_t)(((kk + b_row) * Cfg::LDA + b_col + np * 16) * 2);
        hw_ldsm_x4_t(x_hi + o, bh[np]);
        hw_ldsm_x4_t(x_lo + o, bl[np]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t ah[4], al[4];
        const uint32_t o = (uint32_t)(((kk + a_row) * Cfg::LDG + a_col + mt * 16) * 2);
        hw_ldsm_x4_t(g_hi + o, ah);
        hw_ldsm_x4_t(g_lo + o, al);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int np = nt >> 1, q = (nt & 1) * 2;
          hw_mma16816(acc[mt][nt], al, bh[np][q], bh[np][q + 1]);
          hw_mma16816(acc[mt][nt], ah, bl[np][q], bl[np][q + 1]);
          hw_mma16816(acc[mt][nt], ah, bh[np][q], bh[np][q + 1]);
        }
      }
    }
  }
  hw_cp_async_wait<0>();

  const size_t ldn = (size_t)p.taps * p.Kp;
  float* dst = p.dW + (size_t)tap * p.Kp;
  const int gq = lane >> 2, qq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int n = n0 + wm * (MT * 16) + mt * 16 + gq;
      const int k = k0 + wn * (NT * 8) + nt * 8 + qq * 2;
      if (k < p.K) {  // K is a multiple of 8 here (channel counts are padded): k + 1 < K as well
        if (n < p.N) {
          atomicAdd(dst + (size_t)n * ldn + k, acc[mt][nt][0]);
          atomicAdd(dst + (size_t)n * ldn + k + 1, acc[mt][nt][1]);
        }
        if (n + 8 < p.N) {
          atomicAdd(dst + (size_t)(n + 8) * ldn + k, acc[mt][nt][2]);
          atomicAdd(dst + (size_t)(n + 8) * ldn + k + 1, acc[mt][nt][3]);
        }
      }
    }
}

// ------------------------------------------------------------------------------------------------
// The same contraction on tcgen05 (the generalisation of tcn_wgrad_tc_kernel, tcn_bwd.cu): both operands are pixel-major in HBM, i.e.
// MN-major for the MMA.  A TMA box {64 channels, bx, by, 1 item, 2 planes} (bx * by = 32 pixels, SWIZZLE_128B) lands in shared memory as
// the canonical MN-major SW128 layout (128-byte rows = 64 channels of one pixel, 8 pixels per 1024-byte atom); a conv tap is a shifted
// box origin and the convolution's zero padding, ragged channel counts and ragged pixel extents are all TMA out-of-bounds fill.
// Roles swapped for coalesced atomics: M = k (input channels, halves of 128), N = n (output-gradient channels, up to 256):
// D[k][n] in TMEM (2 x 256 columns); a thread (= TMEM lane = k) adds 32 consecutive n to dW[n][tap][k].
// One CTA = (tap, 256 n x 256 k tile, pixel-tile chunk, item); warp 0 TMA producer, warp 1 MMA issuer (bf16x3), warps 2-5 epilogue.
// ------------------------------------------------------------------------------------------------
constexpr int HT_BK = 32, HT_STAGES = 3;
constexpr int HT_BOX = 64 * HT_BK * 2 * 2;   // one TMA box: 64 channels x 32 pixels x (hi, lo) = 8 KB
constexpr int HT_PLANE = 64 * HT_BK * 2;     // lo plane offset inside a box
constexpr int HT_STAGE_MAX = 8 * HT_BOX;     // up to 4 A boxes + 4 G boxes
constexpr int HT_SMEM = HT_STAGES * HT_STAGE_MAX + 1024 + 256;

struct HtParams {
  int taps, N, K, Kp;
  int dx[16], dy[16];
  int bx, by;              // pixel box
  int tiles_x, ptiles;     // pixel tiles per row of tiles / per item
  int tchunk, nchunks;     // pixel tiles per CTA, chunks per item
  int ntn, ntk;            // 256-wide n tiles / k tiles
  float* dW;               // [N][taps][Kp]
};
struct alignas(64) HtMap { unsigned char b[128]; };

__device__ __forceinline__ void ht_tma_load_5d(void* dst, const void* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint64_t ht_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1) hd_wgrad_tc_kernel(const __grid_constant__ HtMap mapG, const __grid_constant__ HtMap mapA, const HtParams p) {
  extern __shared__ uint8_t ht_smem_raw[];
  const uint32_t raw = smem_u32(ht_smem_raw);
  uint8_t* smem = ht_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + HT_STAGES * HT_STAGE_MAX);
  uint64_t* empty_bar = full_bar + HT_STAGES;
  uint64_t* tfull_bar = empty_bar + HT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int bxi = blockIdx.x;
  const int tk = bxi % p.ntk; bxi /= p.ntk;
  const int tn = bxi % p.ntn;
  const int tap = bxi / p.ntn;
  const int chunk = blockIdx.y, b = blockIdx.z;
  const int n0 = tn * 256, k0 = tk * 256;
  const int n_left = min(256, p.N - n0), k_left = min(256, p.K - k0);
  const int nb_g = (n_left + 63) >> 6;            // G boxes (64 channels each)
  const int halves = (k_left + 127) >> 7;         // M = 128 halves of the k tile; each half is two A boxes (a missing one is all OOB zero)
  const int n_eff = (n_left + 15) & ~15;          // MMA N
  const uint32_t stage_bytes = (uint32_t)(2 * halves + nb_g) * HT_BOX;
  const int t_begin = chunk * p.tchunk, t_end = min(p.ptiles, t_begin + p.tchunk);
  const int iters = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        uint8_t* st = smem + s * HT_STAGE_MAX;
        const int pt = t_begin + it;
        const int x0 = (pt % p.tiles_x) * p.bx, y0 = (pt / p.tiles_x) * p.by;
        for (int cb = 0; cb < 2 * halves; ++cb) ht_tma_load_5d(st + cb * HT_BOX, &mapA, k0 + cb * 64, x0 + p.dx[tap], y0 + p.dy[tap], b, 0, &full_bar[s]);
        for (int cb = 0; cb < nb_g; ++cb) ht_tma_load_5d(st + (2 * halves + cb) * HT_BOX, &mapG, n0 + cb * 64, x0, y0, b, 0, &full_bar[s]);
        if (++s == HT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && iters > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, n_eff) | (1u << 15) | (1u << 16);   // MN-major A and B
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t xb = smem_u32(smem + s * HT_STAGE_MAX), gb = xb + (uint32_t)(2 * halves) * HT_BOX;
#pragma unroll
        for (int kk = 0; kk < HT_BK / 16; ++kk) {
          const uint32_t ko = kk * 2048;  // 16 pixels = two 1024-byte atoms
          const uint64_t g_hi = ht_desc_mn_sw128(gb + ko, HT_BOX), g_lo = ht_desc_mn_sw128(gb + HT_PLANE + ko, HT_BOX);
          for (int h = 0; h < halves; ++h) {
            const uint64_t x_hi = ht_desc_mn_sw128(xb + h * 2 * HT_BOX + ko, HT_BOX), x_lo = ht_desc_mn_sw128(xb + h * 2 * HT_BOX + HT_PLANE + ko, HT_BOX);
            const uint32_t d = tmem + h * 256;
            const uint32_t first = (it == 0 && kk == 0) ? 0u : 1u;
            umma_f16(d, x_lo, g_hi, idesc, first);
            umma_f16(d, x_hi, g_lo, idesc, 1u);
            umma_f16(d, x_hi, g_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (++s == HT_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else if (iters > 0) {
    // epilogue: warp w may only read TMEM lanes [32 (w % 4), +32)
    const int q = warp & 3;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const size_t ldn = (size_t)p.taps * p.Kp;
    float* dst = p.dW + (size_t)tap * p.Kp;
    for (int h = 0; h < halves; ++h) {
      const int k = k0 + h * 128 + q * 32 + lane;
      for (int cc = 0; cc * 32 < n_eff; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 256 + cc * 32, v);
        tmem_ld_wait();
        if (k < p.K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = n0 + cc * 32 + j;
            if (n < p.N) atomicAdd(dst + (size_t)n * ldn + k, __uint_as_float(v[j]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
// Narrow layers (N <= 128 output-gradient channels, K <= 64 input channels: the 48 -> 96 3x3 / k = 3 convs and DConv's 12 ... 48 channel
// convs at full resolution).  The kernel above is bound by shared-memory fill there: one CTA per tap re-loads the output-gradient
// boxes for every tap and the M = 128 MMA needs a second, all-zero activation box (a 9-tap 48 -> 96 conv: 288 KB per 32-pixel tile).
// Here the roles are swapped -- M = n (the two output-gradient boxes, loaded ONCE per pixel tile), N = ceil16(K) input channels (one
// activation box per tap) -- and the accumulators of ALL taps sit side by side in tensor memory (taps * kcols <= 512 columns), so one
// CTA = (pixel chunk, item) and a pixel tile costs (2 + taps) boxes: 88 KB for the 9-tap conv.  Same maps, same staging layout
// dW[n][tap][k] as above (thread = TMEM lane = n).
// ------------------------------------------------------------------------------------------------
constexpr int HF_MAX_STAGE_BOXES = 2 + 16;
__global__ void __launch_bounds__(192, 1) hd_wgrad_tc_fused_kernel(const __grid_constant__ HtMap mapG, const __grid_constant__ HtMap mapA, const HtParams p,
                                                                    int stages, int kcols, int tmem_cols, float* __restrict__ dB) {
  // dB != nullptr: the four epilogue warps, idle during the main loop, also sum the output-gradient boxes of every stage over their
  // 32 pixels (thread = channel n) -- the bias gradient, for which a separate pass would read the whole gradient tensor again
  extern __shared__ uint8_t hf_smem_raw[];
  const uint32_t raw = smem_u32(hf_smem_raw);
  uint8_t* smem = hf_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t stage_bytes = (uint32_t)(2 + p.taps) * HT_BOX;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* tfull_bar = empty_bar + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int k_eff = (p.K + 15) & ~15;   // MMA N
  const int t_begin = chunk * p.tchunk, t_end = min(p.ptiles, t_begin + p.tchunk);
  const int iters = t_end - t_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], dB ? 5 : 1); }   // MMA commit (+ 4 summing warps)
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, (uint32_t)tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        uint8_t* st = smem + (size_t)s * stage_bytes;
        const int pt = t_begin + it;
        const int x0 = (pt % p.tiles_x) * p.bx, y0 = (pt / p.tiles_x) * p.by;
        for (int cb = 0; cb < 2; ++cb) ht_tma_load_5d(st + cb * HT_BOX, &mapG, cb * 64, x0, y0, b, 0, &full_bar[s]);   // (n >= N: zero fill)
        for (int t = 0; t < p.taps; ++t) ht_tma_load_5d(st + (2 + t) * HT_BOX, &mapA, 0, x0 + p.dx[t], y0 + p.dy[t], b, 0, &full_bar[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && iters > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, k_eff) | (1u << 15) | (1u << 16);   // MN-major A and B
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t gb = smem_u32(smem + (size_t)s * stage_bytes), xb = gb + 2 * HT_BOX;
#pragma unroll
        for (int kk = 0; kk < HT_BK / 16; ++kk) {
          const uint32_t ko = kk * 2048;  // 16 pixels = two 1024-byte atoms
          const uint64_t g_hi = ht_desc_mn_sw128(gb + ko, HT_BOX), g_lo = ht_desc_mn_sw128(gb + HT_PLANE + ko, HT_BOX);
          const uint32_t first = (it == 0 && kk == 0) ? 0u : 1u;
          for (int t = 0; t < p.taps; ++t) {
            const uint64_t x_hi = ht_desc_mn_sw128(xb + t * HT_BOX + ko, HT_BOX), x_lo = ht_desc_mn_sw128(xb + t * HT_BOX + HT_PLANE + ko, HT_BOX);
            const uint32_t d = tmem + (uint32_t)(t * kcols);
            umma_f16(d, g_lo, x_hi, idesc, first);
            umma_f16(d, g_hi, x_lo, idesc, 1u);
            umma_f16(d, g_hi, x_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else if (iters > 0) {
    // epilogue: warp w may only read TMEM lanes [32 (w % 4), +32); lane = n
    const int q = warp & 3;
    if (dB) {
      const int nn = q * 32 + lane;
      // SW128 MN-major box: pixel p = 128-byte row, channel c of the box in 16-byte chunk ((c >> 3) ^ (p & 7)), hi plane then lo plane
      const uint32_t coff = (uint32_t)(nn >> 6) * HT_BOX + (uint32_t)(nn & 7) * 2u;
      const uint32_t cch = (uint32_t)((nn & 63) >> 3);
      float acc = 0.0f;
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait_backoff(&full_bar[s], ph);
        const uint8_t* st = smem + (size_t)s * stage_bytes + coff;
#pragma unroll 8
        for (int px = 0; px < HT_BK; ++px) {
          const uint32_t o = (uint32_t)px * 128u + ((cch ^ (uint32_t)(px & 7)) << 4);
          acc += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(st + o)) +
                 __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(st + HT_PLANE + o));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      if (nn < p.N) atomicAdd(dB + nn, acc);
    }
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int n = q * 32 + lane;
    const size_t ldn = (size_t)p.taps * p.Kp;
    float* dst = p.dW + (size_t)n * ldn;
    for (int t = 0; t < p.taps; ++t) {
      for (int c8 = 0; c8 < k_eff; c8 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * kcols + c8), v);
        tmem_ld_wait();
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c8 + j < p.K) atomicAdd(dst + (size_t)t * p.Kp + c8 + j, __uint_as_float(v[j]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, (uint32_t)tmem_cols); }
}

// ------------------------------------------------------------------------------------------------
// LSTM backward (one bidirectional layer; tools/hd_bwd_emul.py:lstm_dir_bwd).  Gx = W_ih x + b (saved by the forward), R = W_hh h_prev
// for every step at once (one GEMM over the saved h), both fp32 [Bs][T][8H], column = dir * 4H + gate * H + unit (gates i, f, g, o).
// ------------------------------------------------------------------------------------------------
// cell state by an element-wise scan: cs[b][t][dir * H + u]
__global__ void __launch_bounds__(256) lstm_cscan_kernel(const float* __restrict__ Gx, const float* __restrict__ R, int Bs, int T, int H,
                                                         float* __restrict__ cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Bs * 2 * H) return;
  const int u = i % H, dir = (i / H) % 2, b = i / (2 * H);
  float c = 0.0f;
  // only c is carried from step to step: the pre-activations of CS_U steps are loaded (and their transcendentals evaluated) as one
  // independent batch, so 6 CS_U loads are in flight per thread instead of 6 (the scan was bound by one DRAM round trip per step)
  constexpr int CS_U = 8;
  for (int k0 = 0; k0 < T; k0 += CS_U) {
    float pi[CS_U], pf[CS_U], pg[CS_U];
#pragma unroll
    for (int j = 0; j < CS_U; ++j) {
      const int k = k0 + j;
      pi[j] = pf[j] = pg[j] = 0.0f;
      if (k < T) {
        const int t = dir ? T - 1 - k : k;
        const size_t go = ((size_t)b * T + t) * 8 * H + (size_t)dir * 4 * H + u;
        pi[j] = Gx[go] + R[go];
        pf[j] = Gx[go + H] + R[go + H];
        pg[j] = Gx[go + 2 * H] + R[go + 2 * H];
      }
    }
#pragma unroll
    for (int j = 0; j < CS_U; ++j) {
      pi[j] = sigmoidf_acc(pi[j]);
      pf[j] = sigmoidf_acc(pf[j]);
      pg[j] = tanhf(pg[j]);
    }
#pragma unroll
    for (int j = 0; j < CS_U; ++j) {
      const int k = k0 + j;
      if (k < T) {
        const int t = dir ? T - 1 - k : k;
        c = pf[j] * c + pi[j] * pg[j];
        cs[((size_t)b * T + t) * 2 * H + dir * H + u] = c;
      }
    }
  }
}

// One step of the reverse-time chain for both directions.  Step k handles time t = T - 1 - k (dir 0) / t = k (dir 1):
//   dh = dH[t] + W_hh^T dG[t'] (t' = the step handled by launch k - 1),  gate derivatives -> dG[t],  dc carry.
// grid = (ceil(H / 128), 2 dirs, ceil(Bs / 8)); 128 threads: thread = unit, 8 sequences per CTA share every W_hh load.
constexpr int LB_NB = 8;
__global__ void __launch_bounds__(128) lstm_bwd_step_kernel(const float* __restrict__ Gx, const float* __restrict__ R, const float* __restrict__ cs,
                                                            const float* __restrict__ dH, const float* __restrict__ Whh /*[2][4H][H]*/,
                                                            float* __restrict__ dG, float* __restrict__ dc_carry /*[Bs][2H]*/, int Bs, int T, int H,
                                                            int k) {
  extern __shared__ float lb_smem[];  // [LB_NB][4H]: dG of the previous launch's time step
  const int dir = blockIdx.y;
  const int b0 = blockIdx.z * LB_NB;
  const int u = blockIdx.x * 128 + threadIdx.x;
  const int t = dir ? k : T - 1 - k;
  const int tn = dir ? t - 1 : t + 1;   // time step of the previous launch
  const int tp = dir ? t + 1 : t - 1;   // forward-previous time (c_prev)
  const int nb = min(LB_NB, Bs - b0);
  float acc[LB_NB];
#pragma unroll
  for (int j = 0; j < LB_NB; ++j) acc[j] = 0.0f;
  if (k > 0) {
    for (int i = threadIdx.x; i < nb * 4 * H; i += 128) {
      const int j = i / (4 * H), g = i % (4 * H);
      lb_smem[j * 4 * H + g] = dG[((size_t)(b0 + j) * T + tn) * 8 * H + (size_t)dir * 4 * H + g];
    }
    __syncthreads();
    if (u < H) {
      const float* w = Whh + (size_t)dir * 4 * H * H + u;
      for (int g = 0; g < 4 * H; ++g) {
        const float wv = __ldg(w + (size_t)g * H);
#pragma unroll
        for (int j = 0; j < LB_NB; ++j)
          if (j < nb) acc[j] = fmaf(lb_smem[j * 4 * H + g], wv, acc[j]);
      }
    }
  }
  if (u >= H) return;
  for (int j = 0; j < nb; ++j) {
    const int b = b0 + j;
    const size_t go = ((size_t)b * T + t) * 8 * H + (size_t)dir * 4 * H + u;
    const size_t ho = ((size_t)b * T + t) * 2 * H + dir * H + u;
    const float gi = sigmoidf_acc(Gx[go] + R[go]);
    const float gf = sigmoidf_acc(Gx[go + H] + R[go + H]);
    const float gg = tanhf(Gx[go + 2 * H] + R[go + 2 * H]);
    const float go_ = sigmoidf_acc(Gx[go + 3 * H] + R[go + 3 * H]);
    const float c = cs[ho];
    const float cprev = (tp >= 0 && tp < T) ? cs[((size_t)b * T + tp) * 2 * H + dir * H + u] : 0.0f;
    const float tc = tanhf(c);
    const float dh = dH[ho] + acc[j];
    float dc = (k > 0 ? dc_carry[(size_t)b * 2 * H + dir * H + u] : 0.0f) + dh * go_ * (1.0f - tc * tc);
    dG[go] = dc * gg * gi * (1.0f - gi);
    dG[go + H] = dc * cprev * gf * (1.0f - gf);
    dG[go + 2 * H] = dc * gi * (1.0f - gg * gg);
    dG[go + 3 * H] = dh * tc * go_ * (1.0f - go_);
    dc_carry[(size_t)b * 2 * H + dir * H + u] = dc * gf;
  }
}

// Persistent form of the same chain (one cooperative launch per layer instead of one launch per step): grid = (H / 16, 2 dirs,
// ceil(Bs / 16)); a CTA owns 16 hidden units of one direction for 16 sequences and keeps its W_hh slice [4H][16] in shared memory for
// the whole launch; thread = (sequence row, unit), so the dc carry lives in a register.  Per step the CTAs of one (direction,
// sequence chunk) group exchange dG[t] through L2 behind a monotonic arrive / wait counter (all CTAs are co-resident: the launch is
// cooperative).  The step's own operands (gates, cell states, dH) are loaded BEFORE the wait, so their latency hides behind it.
constexpr int LBP_U = 16, LBP_B = 16;
__device__ __forceinline__ unsigned lbp_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) lstm_bwd_persist_kernel(const float* __restrict__ Gx, const float* __restrict__ R, const float* __restrict__ cs,
                                                               const float* __restrict__ dH, const float* __restrict__ Whh /*[2][4H][H]*/,
                                                               float* dG, int Bs, int T, int H, unsigned* bar /*[2 * gridDim.z]*/) {
  extern __shared__ __align__(16) float lp_smem[];
  const int G4 = 4 * H, lda = G4 + 4;
  float* Ws = lp_smem;                       // [4H][LBP_U]
  float* As = lp_smem + (size_t)G4 * LBP_U;  // [LBP_B][4H + 4]
  float* red = As + (size_t)LBP_B * lda;     // [4 gates][LBP_B][LBP_U] partial sums of the W_hh^T dG product
  const int dir = blockIdx.y, u0 = blockIdx.x * LBP_U, b0 = blockIdx.z * LBP_B;
  const int tu = threadIdx.x % LBP_U, tb = threadIdx.x / LBP_U;
  const int u = u0 + tu, b = b0 + tb;
  const bool live = b < Bs;
  unsigned* ctr = bar + dir * gridDim.z + blockIdx.z;
  const unsigned gsz = gridDim.x;
  for (int i = threadIdx.x; i < G4 * LBP_U; i += 256) {
    const int g = i / LBP_U, uu = i % LBP_U;
    Ws[i] = Whh[(size_t)dir * G4 * H + (size_t)g * H + u0 + uu];
  }
  float dc = 0.0f;
  __syncthreads();
  for (int k = 0; k < T; ++k) {
    const int t = dir ? k : T - 1 - k;
    const int tn = dir ? t - 1 : t + 1;
    const int tp = dir ? t + 1 : t - 1;
    // this step's own operands: issued before the wait
    float pi = 0.f, pf = 0.f, pg = 0.f, po = 0.f, c = 0.f, cprev = 0.f, dh = 0.f;
    size_t go = 0;
    if (live) {
      go = ((size_t)b * T + t) * 8 * H + (size_t)dir * G4 + u;
      const size_t ho = ((size_t)b * T + t) * 2 * H + dir * H + u;
      pi = Gx[go] + R[go]; pf = Gx[go + H] + R[go + H]; pg = Gx[go + 2 * H] + R[go + 2 * H]; po = Gx[go + 3 * H] + R[go + 3 * H];
      c = cs[ho];
      cprev = (tp >= 0 && tp < T) ? cs[((size_t)b * T + tp) * 2 * H + dir * H + u] : 0.0f;
      dh = dH[ho];
    }
    if (k > 0) {
      if (threadIdx.x == 0) {
        const unsigned target = (unsigned)k * gsz;
        while (lbp_ld_acquire(ctr) < target) { }
      }
      __syncthreads();
      // dG of the previous step for this chunk's sequences (every unit of this direction): L2 -> shared memory
      for (int i = threadIdx.x; i < LBP_B * (G4 / 4); i += 256) {
        const int rr = i / (G4 / 4), g4 = (i % (G4 / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + rr < Bs) v = __ldcg(reinterpret_cast<const float4*>(dG + ((size_t)(b0 + rr) * T + tn) * 8 * H + (size_t)dir * G4 + g4));
        *reinterpret_cast<float4*>(As + (size_t)rr * lda + g4) = v;
      }
      __syncthreads();
      // W_hh^T dG for (16 sequences x 16 units), register-tiled: thread = (gate kq, 4 sequences sg, unit tu) keeps 4 x 2 partial
      // sums, so every W value read from shared memory feeds 4 FMAs (8 shared-memory loads per 16 FMAs; with thread = (sequence,
      // unit) over all 4H rows it was 5 loads per 4 FMAs and the launch was bound by the shared-memory pipe at 12-14.5 us per step).
      // The four gates' partial sums meet in `red`.
      {
        const int kq = threadIdx.x >> 6, sg = (threadIdx.x >> 4) & 3;
        float a[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i][0] = a[i][1] = 0.f;
        const float* wk = Ws + (size_t)kq * H * LBP_U + tu;
        const float* ak = As + (size_t)(4 * sg) * lda + kq * H;
#pragma unroll 2
        for (int g = 0; g < H; g += 4) {
          const float w0 = wk[(g + 0) * LBP_U], w1 = wk[(g + 1) * LBP_U], w2 = wk[(g + 2) * LBP_U], w3 = wk[(g + 3) * LBP_U];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 av = *reinterpret_cast<const float4*>(ak + (size_t)i * lda + g);
            a[i][0] = fmaf(av.x, w0, a[i][0]);
            a[i][1] = fmaf(av.y, w1, a[i][1]);
            a[i][0] = fmaf(av.z, w2, a[i][0]);
            a[i][1] = fmaf(av.w, w3, a[i][1]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) red[(kq * LBP_B + 4 * sg + i) * LBP_U + tu] = a[i][0] + a[i][1];
      }
      __syncthreads();
      dh += (red[(0 * LBP_B + tb) * LBP_U + tu] + red[(1 * LBP_B + tb) * LBP_U + tu]) +
            (red[(2 * LBP_B + tb) * LBP_U + tu] + red[(3 * LBP_B + tb) * LBP_U + tu]);
    }
    if (live) {
      const float gi = sigmoidf_acc(pi), gf = sigmoidf_acc(pf), gg = tanhf(pg), go_ = sigmoidf_acc(po);
      const float tc = tanhf(c);
      dc += dh * go_ * (1.0f - tc * tc);
      __stcg(dG + go, dc * gg * gi * (1.0f - gi));
      __stcg(dG + go + H, dc * cprev * gf * (1.0f - gf));
      __stcg(dG + go + 2 * H, dc * gi * (1.0f - gg * gg));
      __stcg(dG + go + 3 * H, dh * tc * go_ * (1.0f - go_));
      dc *= gf;
    }
    __threadfence();
    __syncthreads();   // every thread's dG stores (and its reads of As) are done
    if (threadIdx.x == 0) atomicAdd(ctr, 1u);
  }
}

// ------------------------------------------------------------------------------------------------
// _LocalState attention backward (TA:832-857; tools/hd_bwd_emul.py:check_local_state).  qkv fp32 [(b, t)][ld]: columns [0, C) query,
// [C, 2C) key, [2C, 3C) content, [3C, 3C + heads * nd) decay logits; dres fp32 [(b, s)][C] = gradient of the attention output.
// One CTA = (32 queries, head, item): recomputes the softmax column W[t][s] of its queries, then
//   dcontent[t] += sum_s W[t][s] dres[s];   dW = content dres^T;   dd = W (dW - sum_t W dW), diagonal 0;
//   dkey[t] += sum_s dd[t][s] q[s] / sqrt(Ch);   dquery[s] = sum_t dd[t][s] k[t] / sqrt(Ch);
//   ddecay[s][f] = -(f + 1) / sqrt(nd) * (sum_t dd[t][s] |t - s|) * sg (1 - sg) / 2.
// dqkv (fp32, same layout as qkv) must be zero on entry (keys / content accumulate with atomics across the query tiles).
// ------------------------------------------------------------------------------------------------
constexpr int LAB_QT = 32;
__global__ void __launch_bounds__(256) local_attn_bwd_kernel(const float* __restrict__ qkv, int ld, int T, int C, int heads, int nd,
                                                             const float* __restrict__ dres, float* __restrict__ dqkv) {
  extern __shared__ float lab_smem[];
  const int Ch = C / heads, Chp = Ch + 1;
  float* Ks = lab_smem;                          // [T][Chp]
  float* Vs = Ks + (size_t)T * Chp;              // [T][Chp]
  float* Ws = Vs + (size_t)T * Chp;              // [T][LAB_QT + 1]  softmax weights
  float* Ds = Ws + (size_t)T * (LAB_QT + 1);     // [T][LAB_QT + 1]  dW, then dd
  float* Qs = Ds + (size_t)T * (LAB_QT + 1);     // [LAB_QT][Chp]
  float* Gs = Qs + LAB_QT * Chp;                 // [LAB_QT][Chp]    dres tile
  float* Dc = Gs + LAB_QT * Chp;                 // [LAB_QT]         decay slope
  float* Rs = Dc + LAB_QT;                       // [LAB_QT]         sum_t W dW
  float* As = Rs + LAB_QT;                       // [LAB_QT]         sum_t dd |t - s|
  const int tid = threadIdx.x;
  const int hd = blockIdx.y, b = blockIdx.z;
  const int s0 = blockIdx.x * LAB_QT;
  const float* base = qkv + (size_t)b * T * ld;
  const float* gbase = dres + (size_t)b * T * C;
  float* obase = dqkv + (size_t)b * T * ld;
  for (int i = tid; i < T * Ch; i += 256) {
    const int t = i / Ch, c = i % Ch;
    Ks[t * Chp + c] = base[(size_t)t * ld + C + hd * Ch + c];
    Vs[t * Chp + c] = base[(size_t)t * ld + 2 * C + hd * Ch + c];
  }
  for (int i = tid; i < LAB_QT * Ch; i += 256) {
    const int sl = i / Ch, c = i % Ch;
    const bool ok = s0 + sl < T;
    Qs[sl * Chp + c] = ok ? base[(size_t)(s0 + sl) * ld + hd * Ch + c] : 0.0f;
    Gs[sl * Chp + c] = ok ? gbase[(size_t)(s0 + sl) * C + hd * Ch + c] : 0.0f;
  }
  if (tid < LAB_QT) {
    float d = 0.0f;
    if (s0 + tid < T)
      for (int f = 0; f < nd; ++f) d += (float)(f + 1) * (sigmoidf_acc(base[(size_t)(s0 + tid) * ld + 3 * C + hd * nd + f]) * 0.5f);
    Dc[tid] = d / sqrtf((float)nd);
  }
  __syncthreads();
  const float inv = 1.0f / sqrtf((float)Ch);
  for (int i = tid; i < T * LAB_QT; i += 256) {
    const int t = i / LAB_QT, sl = i % LAB_QT;
    const int s = s0 + sl;
    float acc = 0.0f, accd = 0.0f;
    for (int c = 0; c < Ch; ++c) {
      acc = fmaf(Ks[t * Chp + c], Qs[sl * Chp + c], acc);
      accd = fmaf(Vs[t * Chp + c], Gs[sl * Chp + c], accd);
    }
    float v = acc * inv - fabsf((float)(t - s)) * Dc[sl];
    if (t == s) v = -100.0f;
    Ws[t * (LAB_QT + 1) + sl] = v;
    Ds[t * (LAB_QT + 1) + sl] = accd;   // dW[t][s] = content[t] . dres[s]
  }
  __syncthreads();
  {  // softmax over t per query, then rowsum = sum_t W dW: one warp handles 4 queries
    const int warp = tid >> 5, lane = tid & 31;
    for (int sl = warp * 4; sl < warp * 4 + 4; ++sl) {
      float mx = -INFINITY;
      for (int t = lane; t < T; t += 32) mx = fmaxf(mx, Ws[t * (LAB_QT + 1) + sl]);
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.0f;
      for (int t = lane; t < T; t += 32) {
        const float e = expf(Ws[t * (LAB_QT + 1) + sl] - mx);
        Ws[t * (LAB_QT + 1) + sl] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float rinv = 1.0f / sum;
      float rs = 0.0f;
      for (int t = lane; t < T; t += 32) {
        const float w = Ws[t * (LAB_QT + 1) + sl] * rinv;
        Ws[t * (LAB_QT + 1) + sl] = w;
        rs = fmaf(w, Ds[t * (LAB_QT + 1) + sl], rs);
      }
      rs = warp_sum(rs);
      const int s = s0 + sl;
      float as = 0.0f;
      for (int t = lane; t < T; t += 32) {
        float dd = Ws[t * (LAB_QT + 1) + sl] * (Ds[t * (LAB_QT + 1) + sl] - rs);
        if (t == s || s >= T) dd = 0.0f;   // masked diagonal: no gradient; padded queries contribute nothing
        Ds[t * (LAB_QT + 1) + sl] = dd;
        as = fmaf(dd, fabsf((float)(t - s)), as);
      }
      as = warp_sum(as);
      if (lane == 0) { Rs[sl] = rs; As[sl] = as; }
    }
  }
  __syncthreads();
  // dcontent[t][c] += sum_s W[t][s] dres[s][c];  dkey[t][c] += sum_s dd[t][s] q[s][c] * inv
  for (int i = tid; i < T * Ch; i += 256) {
    const int t = i / Ch, c = i % Ch;
    float av = 0.0f, ak = 0.0f;
    for (int sl = 0; sl < LAB_QT; ++sl) {
      if (s0 + sl >= T) break;
      av = fmaf(Ws[t * (LAB_QT + 1) + sl], Gs[sl * Chp + c], av);
      ak = fmaf(Ds[t * (LAB_QT + 1) + sl], Qs[sl * Chp + c], ak);
    }
    atomicAdd(obase + (size_t)t * ld + 2 * C + hd * Ch + c, av);
    atomicAdd(obase + (size_t)t * ld + C + hd * Ch + c, ak * inv);
  }
  // dquery[s][c] = sum_t dd[t][s] k[t][c] * inv
  for (int i = tid; i < LAB_QT * Ch; i += 256) {
    const int sl = i / Ch, c = i % Ch;
    const int s = s0 + sl;
    if (s >= T) continue;
    float aq = 0.0f;
    for (int t = 0; t < T; ++t) aq = fmaf(Ds[t * (LAB_QT + 1) + sl], Ks[t * Chp + c], aq);
    obase[(size_t)s * ld + hd * Ch + c] = aq * inv;
  }
  if (tid < LAB_QT * nd) {
    const int sl = tid / nd, f = tid % nd;
    const int s = s0 + sl;
    if (s < T) {
      const float sg = sigmoidf_acc(base[(size_t)s * ld + 3 * C + hd * nd + f]);
      obase[(size_t)s * ld + 3 * C + hd * nd + f] = -(float)(f + 1) / sqrtf((float)nd) * As[sl] * sg * (1.0f - sg) * 0.5f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The four "narrow" layers: first encoders (1 or 2 input channels -> C, k 8, stride 4, TA:124) and last decoders (C -> 1 or 2,
// transposed k 8 stride 4, TA:243).  With `wide` the C-channel side ([line][x][C]) and `narrow` the P-channel side ([line][n][P]):
//   narrow position of (wide row x, tap j):  n = x * S + j - pad        weight w[c][p][j]  (Conv [Co][Ci][k] / ConvTranspose [Ci][Co][k])
//   ENC:  pre[x][c] = b_c + sum_{j,p} w[c][p][j] nv[n][p];  wide value = dwide[x][c] * gelu'(pre)   (the forward fuses GELU: pre is recomputed)
//         dW[c][p][j] += wide value * nv[n][p],  db[c] += wide value
//   DEC:  wide value = y[x][c];  dW[c][p][j] += y * nv[n][p];  dy[x][c] = sum_{j,p} w[c][p][j] nv[n][p]  (nv = output gradient);  db[p] = sum nv
// nv = (raw - sub_b) * mul_b * (ck ? (n == 0 ? 1 : 2) : 1): input normalisation (TA:553-563) on the way in, de-normalisation and the
// irfft bin weights on the way out.
// ------------------------------------------------------------------------------------------------
struct NarrowP {
  const float* narrow; long long n_line;   // line stride (elements); lines = B * Y
  int Xn;                                   // narrow positions per line
  const float* stats;                       // per item (mean, std) or nullptr
  int norm_mode;                            // 0: nv = raw; 1: (raw - mean) / (1e-5 + std); 2: raw * std
  int ck;                                   // 1: multiply by (n == 0 ? 1 : 2)
  int Y;                                    // lines per item
  int Xw, C, S, pad;                        // wide extent per line, channels, stride, padding
  const float* w; const float* bias;        // [C][P][K], [C] (ENC) or [P] (DEC)
  const float* dwide;                       // ENC: gradient of the wide activation, fp32
  const __nv_bfloat16* yhi; const __nv_bfloat16* ylo;  // DEC: wide activation planes
  float* dy;                                // DEC: gradient of the wide activation (written)
  float* dW; float* db;
  int rows_per_cta;
  int nlines, nchunks;                      // work items = nlines * nchunks; the grid strides over them
};

// Persistent: a 1-D grid of (at most) two CTAs per SM strides over the work items (line, chunk of wide rows), a thread owns CH
// wide channels (CH * P * K = 64 weight-gradient sums in registers for the CTA's whole life) and the sums reach shared / global
// memory once per CTA.  The weights sit in shared memory TRANSPOSED ([P * K][C]): the threads of a warp read consecutive channels of
// one tap (conflict-free; the [C][P * K] order put every channel group on the same bank).
template <int CH>
__device__ __forceinline__ void nb_load_wide(const __nv_bfloat16* hi, const __nv_bfloat16* lo, size_t off, float (&o)[CH]) {
  if constexpr (CH == 8) {
    bw_load_split8(hi, lo, off, o);
  } else {
    static_assert(CH == 4, "narrow layers: 4 or 8 channels per thread");
    const uint2 h = *reinterpret_cast<const uint2*>(hi + off), l = *reinterpret_cast<const uint2*>(lo + off);
    const uint32_t hw[2] = {h.x, h.y}, lw[2] = {l.x, l.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[i]));
      const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[i]));
      o[2 * i] = hf.x + lf.x;
      o[2 * i + 1] = hf.y + lf.y;
    }
  }
}
template <int K, int P, bool ENC, int CH>
__global__ void __launch_bounds__(256, 2) narrow_bwd_kernel(const NarrowP p) {
  extern __shared__ __align__(16) float nb_smem[];  // [P * K][C] weights (transposed), [C * P * K + C] accumulators
  constexpr int PK = P * K;
  const int C = p.C, WN = C * PK;
  float* sw = nb_smem;
  float* sa = nb_smem + WN;
  for (int i = threadIdx.x; i < WN; i += 256) sw[(i % PK) * C + i / PK] = p.w[i];
  for (int i = threadIdx.x; i < WN + C; i += 256) sa[i] = 0.0f;
  __syncthreads();
  const int groups = C / CH, rows = 256 / groups;
  const int g8 = threadIdx.x % groups, r = threadIdx.x / groups, c0 = g8 * CH;
  float accw[CH][PK];
  float accb[CH], bias[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    accb[i] = 0.0f;
    bias[i] = (ENC && r < rows) ? p.bias[c0 + i] : 0.0f;
#pragma unroll
    for (int q = 0; q < PK; ++q) accw[i][q] = 0.0f;
  }
  const long long n_items = (long long)p.nlines * p.nchunks;
  if (r < rows) {
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int line = (int)(item / p.nchunks), chunk = (int)(item % p.nchunks), b = line / p.Y;
      float sub = 0.0f, mul = 1.0f;
      if (p.stats) {
        if (p.norm_mode == 1) { sub = p.stats[2 * b]; mul = 1.0f / (1e-5f + p.stats[2 * b + 1]); }
        else if (p.norm_mode == 2) { mul = p.stats[2 * b + 1]; }
      }
      const float* nl = p.narrow + (size_t)line * p.n_line;
      const int x_begin = chunk * p.rows_per_cta, x_end = min(p.Xw, x_begin + p.rows_per_cta);
      for (int x = x_begin + r; x < x_end; x += rows) {
        float nv[PK];
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int n = x * p.S + j - p.pad;
          const bool ok = n >= 0 && n < p.Xn;
#pragma unroll
          for (int q = 0; q < P; ++q) {
            float v = 0.0f;
            if (ok) {
              v = (__ldg(nl + (size_t)n * P + q) - sub) * mul;
              if (p.ck && n != 0) v *= 2.0f;
            }
            nv[q * K + j] = v;
          }
        }
        const size_t woff = ((size_t)line * p.Xw + x) * C + c0;
        float wide[CH];
        if (ENC) {
          float dv[CH];
#pragma unroll
          for (int i = 0; i < CH; i += 4) {
            const float4 d = *reinterpret_cast<const float4*>(p.dwide + woff + i);
            dv[i] = d.x; dv[i + 1] = d.y; dv[i + 2] = d.z; dv[i + 3] = d.w;
          }
          float pre[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) pre[i] = bias[i];
#pragma unroll
          for (int q = 0; q < PK; ++q) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(sw + q * C + c0 + i);
              pre[i] = fmaf(w4.x, nv[q], pre[i]); pre[i + 1] = fmaf(w4.y, nv[q], pre[i + 1]);
              pre[i + 2] = fmaf(w4.z, nv[q], pre[i + 2]); pre[i + 3] = fmaf(w4.w, nv[q], pre[i + 3]);
            }
          }
#pragma unroll
          for (int i = 0; i < CH; ++i) {
            wide[i] = dv[i] * gelu_grad(pre[i]);
            accb[i] += wide[i];
          }
        } else {
          nb_load_wide<CH>(p.yhi, p.ylo, woff, wide);
          float dyv[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) dyv[i] = 0.0f;
#pragma unroll
          for (int q = 0; q < PK; ++q) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
              const float4 w4 = *reinterpret_cast<const float4*>(sw + q * C + c0 + i);
              dyv[i] = fmaf(w4.x, nv[q], dyv[i]); dyv[i + 1] = fmaf(w4.y, nv[q], dyv[i + 1]);
              dyv[i + 2] = fmaf(w4.z, nv[q], dyv[i + 2]); dyv[i + 3] = fmaf(w4.w, nv[q], dyv[i + 3]);
            }
          }
#pragma unroll
          for (int i = 0; i < CH; i += 4)
            *reinterpret_cast<float4*>(p.dy + woff + i) = make_float4(dyv[i], dyv[i + 1], dyv[i + 2], dyv[i + 3]);
        }
#pragma unroll
        for (int i = 0; i < CH; ++i)
#pragma unroll
          for (int q = 0; q < PK; ++q) accw[i][q] = fmaf(wide[i], nv[q], accw[i][q]);
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
#pragma unroll
      for (int q = 0; q < PK; ++q) atomicAdd(&sa[(c0 + i) * PK + q], accw[i][q]);
      if (ENC) atomicAdd(&sa[WN + c0 + i], accb[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < WN; i += 256) atomicAdd(p.dW + i, sa[i]);
  if (ENC)
    for (int i = threadIdx.x; i < C; i += 256) atomicAdd(p.db + i, sa[WN + i]);
}

// DEC bias gradient: db[q] += sum over every narrow position of nv (each output position receives the bias once)
template <int P>
__global__ void __launch_bounds__(256) narrow_bias_kernel(const NarrowP p) {
  __shared__ float red[P][8];
  const int line = blockIdx.y, b = line / p.Y;
  float mul = 1.0f;
  if (p.stats && p.norm_mode == 2) mul = p.stats[2 * b + 1];
  const float* nl = p.narrow + (size_t)line * p.n_line;
  float acc[P];
#pragma unroll
  for (int q = 0; q < P; ++q) acc[q] = 0.0f;
  for (int n = blockIdx.x * 256 + threadIdx.x; n < p.Xn; n += gridDim.x * 256) {
    const float f = (p.ck && n != 0) ? 2.0f * mul : mul;
#pragma unroll
    for (int q = 0; q < P; ++q) acc[q] = fmaf(nl[(size_t)n * P + q], f, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < P; ++q) {
    acc[q] = warp_sum(acc[q]);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = acc[q];
  }
  __syncthreads();
  if (threadIdx.x < P) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(p.db + threadIdx.x, t);
  }
}

// ---- iSTFT adjoint, step 1: ghat[b][P0 + i] = dout[b][i] / envelope(i), zero in the P0 / P1 pads (TA:941-961 backward).
//      envelope(i) = sum over the frames t in [-env_pad, F + env_pad) covering i of window[i + frame_off - t * hop]^2 ----
__global__ void __launch_bounds__(256) istft_adj_prep_kernel(const float* __restrict__ dout, int T, const float* __restrict__ window, int n_fft, int hop,
                                                             int frame_off, int F, int env_pad, int P0, int Ltot, float* __restrict__ ghat) {
  const long long j = blockIdx.x * 256ll + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= Ltot) return;
  const long long i = j - P0;
  float v = 0.0f;
  if (i >= 0 && i < T) {
    const long long pos = i + frame_off;
    long long t_hi = pos >= 0 ? pos / hop : -((-pos + hop - 1) / hop);
    if (t_hi > F + env_pad - 1) t_hi = F + env_pad - 1;
    float env = 0.0f;
    for (long long t = t_hi; t >= -env_pad; --t) {
      const long long n = pos - t * hop;
      if (n >= n_fft) break;
      const float w = window[n];
      env = fmaf(w, w, env);
    }
    v = dout[(size_t)b * T + i] / env;
  }
  ghat[(size_t)b * Ltot + j] = v;
}

}  // namespace hd
}  // namespace rfx
