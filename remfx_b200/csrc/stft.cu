// STFT / iSTFT kernels (cuFFT-free).  Reference semantics: torch.stft / torch.istft with center=True,
// pad_mode="reflect", onesided, as called by
//   umx/openunmix/transforms.py:106-116 (TorchSTFT), :168-177 (TorchISTFT), :198-216 (ComplexNorm),
//   remfx/utils.py:138-159 (spectrogram), torchaudio MelSpectrogram (remfx/classifier.py:156-161),
//   auraloss STFTLoss (remfx/models.py:289-291).
// Data layout is frame-major: row m = b * F + t, bins contiguous (so the FC layers that follow read
// K-contiguous rows and no transpose is ever materialised).
#include "kernels.h"
#include "fft.cuh"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>
#include <cmath>

namespace rfx {

// -------------------------------------------------------------------------------------------------
// Twiddle tables exp(-2 pi i m / n_fft), computed in double on the host, cached per (device, n_fft).
// -------------------------------------------------------------------------------------------------
static std::mutex g_tw_mu;
static std::map<std::pair<int, int>, float2*> g_tw;

const float2* twiddles(int n_fft) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_tw_mu);
  auto key = std::make_pair(dev, n_fft);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) return it->second;
  std::vector<float2> h(n_fft);
  for (int m = 0; m < n_fft; ++m) {
    const double a = -2.0 * M_PI * (double)m / (double)n_fft;
    h[m] = make_float2((float)cos(a), (float)sin(a));
  }
  float2* d = nullptr;
  if (cudaMalloc(&d, sizeof(float2) * n_fft) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, h.data(), sizeof(float2) * n_fft, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
  g_tw[key] = d;
  return d;
}

__device__ __forceinline__ int reflect_index(int i, int T) {
  if (i < 0) i = -i;
  if (i >= T) i = 2 * (T - 1) - i;
  return i;
}

// -------------------------------------------------------------------------------------------------
// Forward STFT.  grid = (ceil(F / FPC), B), block = NC/4 threads; each CTA transforms FPC frames.
// -------------------------------------------------------------------------------------------------
constexpr int STFT_FPC = 4;

template <int LOG2NC>
__global__ void __launch_bounds__((1 << LOG2NC) / 4) stft_kernel(StftParams p, int groups, int n_work) {
  constexpr int NC = 1 << LOG2NC;
  constexpr int T4 = NC / 4;
  constexpr int NFFT = 2 * NC;
  __shared__ float2 sa[NC];
  __shared__ float2 sb[NC];
  const int j = threadIdx.x;
  // work item = (batch item, group of STFT_FPC frames); the grid strides over them (it is capped when other kernels
  // must keep part of the chip, see StftParams::max_ctas)
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
  const int b = work / groups;
  const int fg = work - b * groups;
  const float* __restrict__ x = p.x + (size_t)b * p.x_bstride;
  const int f_end = min(p.F, (fg + 1) * STFT_FPC);
  for (int f = fg * STFT_FPC; f < f_end; ++f) {
    const int base = f * p.hop - p.frame_off;  // first sample of the frame (frame_off = n_fft/2 for centre padding)
    const bool interior = (base >= 0) && (base + NFFT <= p.T);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = j + r * T4;
      const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
      float2 v;
      if (interior && p.x_aligned8) {
        v = *reinterpret_cast<const float2*>(x + base + 2 * n);
      } else {
        v.x = x[reflect_index(base + 2 * n, p.T)];
        v.y = x[reflect_index(base + 2 * n + 1, p.T)];
      }
      sa[n] = make_float2(v.x * w.x, v.y * w.y);
    }
    const float2* Zp = fft_block<LOG2NC>(sa, sb, p.tw, j);
    const size_t m = (size_t)b * p.F + f;
    for (int k = j; k < p.nbins; k += T4) {
      float2 X = rfft_post(Zp, p.tw, NC, k);
      X.x *= p.scale;
      X.y *= p.scale;
      if (p.Z) p.Z[m * p.ldz + k] = X;
      if (p.mode != STFT_COMPLEX) {
        const float pw = X.x * X.x + X.y * X.y;
        float a;
        switch (p.mode) {
          case STFT_UMX_MAG: a = (sqrtf(pw) + p.in_mean[k]) * p.in_scale[k]; break;  // ComplexNorm (transforms.py:211) + input affine (model.py:127-128)
          case STFT_MAG: a = sqrtf(pw); break;
          case STFT_POWER: a = pw; break;
          case STFT_MAG_CLAMP: a = sqrtf(fmaxf(pw, 1e-8f)); break;
          case STFT_UMX_POW: a = (powf(sqrtf(pw) + 1e-8f, p.alpha) + p.in_mean[k]) * p.in_scale[k]; break;
          default: a = powf(sqrtf(pw) + 1e-8f, p.alpha); break;  // STFT_MAG_POW
        }
        if (p.A) p.A[m * p.lda + k] = a;
        if (p.Ahi) {  // split-bf16 copy for the tensor-core layer that consumes it
          __nv_bfloat16 h, l;
          split_bf16(a, h, l);
          p.Ahi[m * p.ldas + k] = h;
          p.Alo[m * p.ldas + k] = l;
        }
      }
    }
    if (p.A && p.lda > NC + 1) {
      for (int k = NC + 1 + j; k < p.lda; k += T4) p.A[m * p.lda + k] = 0.0f;
    }
    __syncthreads();  // sa/sb are reused by the next frame
  }
  }
}

// n_fft = 2048 (Open-Unmix, the Cnn14 mel front end, the widest loss resolution): FOUR frames per 256-thread CTA at once, one per
// 64-thread group, on the register-pass FFT (fft1024_x4: radix 16, 16, 4; two shared-memory exchanges instead of five).
// Every thread keeps 16 independent loads in flight in the load phase, which is what hides the HBM latency.
template <int MODE>
__device__ __forceinline__ float stft_mag_of(float pw, float in_mean, float in_scale, float alpha) {
  if (MODE == STFT_UMX_MAG) return (sqrtf(pw) + in_mean) * in_scale;  // ComplexNorm (transforms.py:211) + input affine (model.py:127-128)
  if (MODE == STFT_MAG) return sqrtf(pw);
  if (MODE == STFT_POWER) return pw;
  if (MODE == STFT_MAG_CLAMP) return sqrtf(fmaxf(pw, 1e-8f));
  if (MODE == STFT_UMX_POW) return (powf(sqrtf(pw) + 1e-8f, alpha) + in_mean) * in_scale;
  return powf(sqrtf(pw) + 1e-8f, alpha);  // STFT_MAG_POW
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) stft2048_kernel(StftParams p, int groups, int n_work) {
  constexpr int NC = 1024, NFFT = 2048;
  __shared__ float2 buf[4 * FFT1024_BUF];
  const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
  float2* fb = buf + g * FFT1024_BUF;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {  // work item = (batch item, group of 4 frames)
    const int b = work / groups;
    const int f = (work - b * groups) * 4 + g;
    const bool active = f < p.F;
    const float* __restrict__ x = p.x + (size_t)b * p.x_bstride;
    const int base = f * p.hop - p.frame_off;  // first sample of the frame (frame_off = n_fft/2 for centre padding)
    const bool interior = active && (base >= 0) && (base + NFFT <= p.T);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const int n = t + 64 * r;
      float2 v = make_float2(0.f, 0.f);
      if (active) {
        const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
        if (interior && p.x_aligned8) {
          v = *reinterpret_cast<const float2*>(x + base + 2 * n);
        } else {
          v.x = x[reflect_index(base + 2 * n, p.T)];
          v.y = x[reflect_index(base + 2 * n + 1, p.T)];
        }
        v.x *= w.x;
        v.y *= w.y;
      }
      fb[n] = v;
    }
    fft1024_x4(buf, p.tw, tid);
    if (active) {
      // epilogue: the mode is a compile-time constant and the row pointers are hoisted, so an iteration is ~30 instructions
      const size_t m = (size_t)b * p.F + f;
      float2* __restrict__ Zrow = p.Z ? p.Z + m * p.ldz : nullptr;
      float* __restrict__ Arow = p.A ? p.A + m * p.lda : nullptr;
      __nv_bfloat16* __restrict__ Hrow = p.Ahi ? p.Ahi + m * p.ldas : nullptr;
      __nv_bfloat16* __restrict__ Lrow = p.Ahi ? p.Alo + m * p.ldas : nullptr;
      const float scale = p.scale;
      for (int k = t; k < p.nbins; k += 64) {
        float2 X = rfft_post(fb, p.tw, NC, k);
        X.x *= scale;
        X.y *= scale;
        if (Zrow) Zrow[k] = X;
        if (MODE != STFT_COMPLEX) {
          const float pw = X.x * X.x + X.y * X.y;
          constexpr bool AFF = MODE == STFT_UMX_MAG || MODE == STFT_UMX_POW;
          const float a = stft_mag_of<MODE>(pw, AFF ? p.in_mean[k] : 0.f, AFF ? p.in_scale[k] : 1.f, p.alpha);
          if (Arow) Arow[k] = a;
          if (Hrow) {  // split-bf16 copy for the tensor-core layer that consumes it
            __nv_bfloat16 h, l;
            split_bf16(a, h, l);
            Hrow[k] = h;
            Lrow[k] = l;
          }
        }
      }
      if (Arow && p.lda > NC + 1) {
        for (int k = NC + 1 + t; k < p.lda; k += 64) Arow[k] = 0.0f;
      }
    }
    __syncthreads();  // buf is refilled by the next work item
  }
}

// ---- n_fft = 2048, bulk-copy staged ("TMA-staged window" of the north star) ----------------------------------------------
// Persistent CTAs; work item = (batch item, 4 consecutive frames).  The 3 hop + 2048 samples the four frames share are brought in
// by ONE cp.async.bulk (SASS UBLKCP): the copy of item i + 1 is issued as soon as every frame of item i has its samples in
// registers and lands during item i's FFT and epilogue, and every sample is read from L2 / HBM once per item instead of once per
// overlapping frame.  The windowed samples go from the staging buffer straight into the first radix-16 butterflies (registers); the
// window AND the twiddle table live in shared memory (with three CTAs of shared memory per SM the L1 is down to ~28 KB and table
// reads through it showed up as the kernel's main stall); the barriers inside the FFT are per frame (64 threads), so the four
// frames of a CTA drift apart; the epilogue handles the bins in (k, NC - k) pairs -- both come from the same two packed-FFT points
// and the same twiddle.  Items that touch the reflect padding (the first / last groups of an item) fill the staging buffer with
// ordinary loads.
template <int MODE>
__device__ __forceinline__ void stft_emit_bin(const StftParams& p, int k, float2 X, float2* __restrict__ Zrow, float* __restrict__ Arow,
                                              __nv_bfloat16* __restrict__ Hrow, __nv_bfloat16* __restrict__ Lrow) {
  if (Zrow) Zrow[k] = X;
  if (MODE != STFT_COMPLEX) {
    const float pw = X.x * X.x + X.y * X.y;
    float mean = 0.f, scl = 1.f;
    if (MODE == STFT_UMX_MAG || MODE == STFT_UMX_POW) {
      if (p.in_ms) {
        const float2 ms = __ldg(p.in_ms + k);
        mean = ms.x; scl = ms.y;
      } else {
        mean = p.in_mean[k]; scl = p.in_scale[k];
      }
    }
    const float a = stft_mag_of<MODE>(pw, mean, scl, p.alpha);
    if (Arow) Arow[k] = a;
    if (Hrow) {  // split-bf16 copy for the tensor-core layer that consumes it
      __nv_bfloat16 h, l;
      split_bf16(a, h, l);
      Hrow[k] = h;
      Lrow[k] = l;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256, 3) stft2048_tma_kernel(StftParams p, int groups, int n_work, int seg) {
  constexpr int NC = 1024, NFFT = 2048;
  extern __shared__ __align__(128) unsigned char stft_smem[];
  float2* buf = reinterpret_cast<float2*>(stft_smem);            // [4][FFT1024_BUF]
  float2* tws = buf + 4 * FFT1024_BUF;                           // [NFFT] twiddles
  float* win = reinterpret_cast<float*>(tws + NFFT);             // [NFFT]
  float* xs = win + NFFT;                                        // [seg] staged samples
  uint64_t* bar = reinterpret_cast<uint64_t*>(xs + seg);
  const int tid = threadIdx.x, g = tid >> 6, t = tid & 63;
  float2* fb = buf + g * FFT1024_BUF;
  for (int i = tid; i < NC; i += 256) reinterpret_cast<float2*>(win)[i] = reinterpret_cast<const float2*>(p.window)[i];
  for (int i = tid; i < NFFT; i += 256) tws[i] = p.tw[i];
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int span = 4 * p.hop;
  auto issue = [&](int work) {  // one thread: start the copy of a work item's samples (interior items only)
    const int b = work / groups;
    const int base = (work - b * groups) * span - p.frame_off;
    if (base >= 0 && base + seg <= p.T) {
      mbar_arrive_expect_tx(bar, (uint32_t)seg * 4u);
      bulk_g2s(xs, p.x + (size_t)b * p.x_bstride + base, (uint32_t)seg * 4u, bar);
    }
  };
  if (tid == 0 && (int)blockIdx.x < n_work) issue(blockIdx.x);
  uint32_t phase = 0;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
    const int b = work / groups;
    const int fg = work - b * groups;
    const int base = fg * span - p.frame_off;
    if (base >= 0 && base + seg <= p.T) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    } else {
      const float* __restrict__ x = p.x + (size_t)b * p.x_bstride;
      for (int i = tid; i < seg; i += 256) {
        int r = base + i;
        if (r < 0) r = -r;
        if (r >= p.T) r = 2 * (p.T - 1) - r;
        xs[i] = (r >= 0 && r < p.T) ? x[r] : 0.0f;  // beyond the reflected range only frames >= F read (and they are not emitted)
      }
      fence_proxy_async_smem();  // generic writes, later overwritten through the async proxy
      __syncthreads();
    }
    const int f = fg * 4 + g;
    float2 v[16];
    {
      const float* xf = xs + g * p.hop;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int n = t + 64 * r;
        const float2 xv = *reinterpret_cast<const float2*>(xf + 2 * n);
        const float2 w = *reinterpret_cast<const float2*>(win + 2 * n);
        v[r] = make_float2(xv.x * w.x, xv.y * w.y);
      }
    }
    __syncthreads();  // every frame has its samples: the staging buffer is free for the next item
    if (tid == 0 && work + (int)gridDim.x < n_work) issue(work + gridDim.x);
    fft1024_x4_regs<true>(buf, tws, tid, v);
    if (f < p.F) {
      const size_t m = (size_t)b * p.F + f;
      float2* __restrict__ Zrow = p.Z ? p.Z + m * p.ldz : nullptr;
      float* __restrict__ Arow = p.A ? p.A + m * p.lda : nullptr;
      __nv_bfloat16* __restrict__ Hrow = p.Ahi ? p.Ahi + m * p.ldas : nullptr;
      __nv_bfloat16* __restrict__ Lrow = p.Ahi ? p.Alo + m * p.ldas : nullptr;
      const float hs = 0.5f * p.scale;
#pragma unroll 2
      for (int i = 0; i < 8; ++i) {
        const int k = t + 64 * i;  // [0, NC/2)
        float2 xk, xn;
        rfft_post_pair2(fb[k], fb[(NC - k) & (NC - 1)], tws[k], xk, xn);
        stft_emit_bin<MODE>(p, k, make_float2(xk.x * hs, xk.y * hs), Zrow, Arow, Hrow, Lrow);
        if (NC - k < p.nbins) stft_emit_bin<MODE>(p, NC - k, make_float2(xn.x * hs, xn.y * hs), Zrow, Arow, Hrow, Lrow);
      }
      if (t == 0) {
        float2 X = rfft_post(fb, tws, NC, NC / 2);
        stft_emit_bin<MODE>(p, NC / 2, make_float2(X.x * p.scale, X.y * p.scale), Zrow, Arow, Hrow, Lrow);
      }
      if (Arow && p.lda > NC + 1) {
        for (int k = NC + 1 + t; k < p.lda; k += 64) Arow[k] = 0.0f;
      }
    }
    fft_sync<true>(tid);  // this frame's buffer is rewritten by the next item's first pass
  }
}

typedef void (*Stft2048TmaFn)(StftParams, int, int, int);
static Stft2048TmaFn stft2048_tma_fn(int mode) {
  switch (mode) {
    case STFT_COMPLEX: return stft2048_tma_kernel<STFT_COMPLEX>;
    case STFT_UMX_MAG: return stft2048_tma_kernel<STFT_UMX_MAG>;
    case STFT_MAG: return stft2048_tma_kernel<STFT_MAG>;
    case STFT_POWER: return stft2048_tma_kernel<STFT_POWER>;
    case STFT_MAG_CLAMP: return stft2048_tma_kernel<STFT_MAG_CLAMP>;
    case STFT_UMX_POW: return stft2048_tma_kernel<STFT_UMX_POW>;
    default: return stft2048_tma_kernel<STFT_MAG_POW>;
  }
}

static int device_sms() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 1;
}

// 0 = the legacy per-frame-load kernels (RFX_STFT_LEGACY=1; kept as cross-check and for shapes the staged kernels do not take)
static bool stft_staged_enabled() {
  static const bool on = [] { const char* e = getenv("RFX_STFT_LEGACY"); return !(e && atoi(e) != 0); }();
  return on;
}

static bool stft2048_tma_ok(const StftParams& p) {
  return stft_staged_enabled() && p.n_fft == 2048 && p.hop % 4 == 0 && p.hop >= 64 && p.hop <= 1024 && p.T % 4 == 0 && p.x_bstride % 4 == 0 &&
         p.frame_off % 4 == 0 && ((uintptr_t)p.x & 15) == 0 && p.T >= 3 * p.hop + 2048 + p.frame_off;
}

static int launch_stft2048_tma(const StftParams& p, int B, cudaStream_t stream) {
  const int seg = 3 * p.hop + 2048;
  const size_t smem = sizeof(float2) * (4 * FFT1024_BUF + 2048) + sizeof(float) * (2048 + seg) + 16;
  const int groups = ceil_div(p.F, 4);
  const long long n_work_ll = (long long)groups * B;
  RFX_REQUIRE(n_work_ll < (1ll << 31), "stft: too many frames");
  Stft2048TmaFn fn = stft2048_tma_fn(p.mode);
  RFX_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  RFX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, smem));
  const int sms = p.max_sms > 0 ? p.max_sms : (p.sms_avail > 0 ? p.sms_avail : device_sms());
  const long long cap = (long long)sms * (per_sm > 0 ? per_sm : 1);
  const int grid = (int)std::min<long long>(n_work_ll, cap);
  fn<<<grid, 256, smem, stream>>>(p, groups, (int)n_work_ll, seg);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

typedef void (*Stft2048Fn)(StftParams, int, int);
static Stft2048Fn stft2048_fn(int mode) {
  switch (mode) {
    case STFT_COMPLEX: return stft2048_kernel<STFT_COMPLEX>;
    case STFT_UMX_MAG: return stft2048_kernel<STFT_UMX_MAG>;
    case STFT_MAG: return stft2048_kernel<STFT_MAG>;
    case STFT_POWER: return stft2048_kernel<STFT_POWER>;
    case STFT_MAG_CLAMP: return stft2048_kernel<STFT_MAG_CLAMP>;
    case STFT_UMX_POW: return stft2048_kernel<STFT_UMX_POW>;
    default: return stft2048_kernel<STFT_MAG_POW>;
  }
}

int launch_stft(const StftParams& p, int B, cudaStream_t stream) {
  RFX_REQUIRE(p.tw != nullptr, "twiddle table");
  RFX_REQUIRE(p.T > p.n_fft / 2, "reflect padding needs T > n_fft/2");
  RFX_REQUIRE(p.frame_off >= 0 && p.frame_off < p.T && p.nbins >= 1 && p.nbins <= p.n_fft / 2 + 1, "stft frame_off / nbins");
  static const bool generic2048 = [] { const char* e = getenv("RFX_STFT_GENERIC"); return e && atoi(e) != 0; }();
  const bool fast = p.n_fft == 2048 && !generic2048;
  if (fast && stft2048_tma_ok(p)) return launch_stft2048_tma(p, B, stream);
  const int groups = ceil_div(p.F, fast ? 4 : STFT_FPC);
  const long long n_work_ll = (long long)groups * B;
  RFX_REQUIRE(n_work_ll < (1ll << 31), "stft: too many frames");
  const int n_work = (int)n_work_ll;
  int grid = n_work;
  if (p.max_sms > 0) {
    int per_sm = 0;
    cudaError_t e = cudaSuccess;
    switch (p.n_fft) {
      case 512: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_kernel<8>, 64, 0); break;
      case 1024: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_kernel<9>, 128, 0); break;
      case 2048: e = fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft2048_fn(p.mode), 256, 0)
                          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_kernel<10>, 256, 0); break;
      case 4096: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_kernel<11>, 512, 0); break;
      default: break;
    }
    RFX_CHECK_CUDA(e);
    const long long cap = (long long)p.max_sms * (per_sm > 0 ? per_sm : 1);
    if (cap < grid) grid = (int)cap;
  }
  switch (p.n_fft) {
    case 512: stft_kernel<8><<<grid, 64, 0, stream>>>(p, groups, n_work); break;
    case 1024: stft_kernel<9><<<grid, 128, 0, stream>>>(p, groups, n_work); break;
    case 2048:
      if (fast) stft2048_fn(p.mode)<<<grid, 256, 0, stream>>>(p, groups, n_work);
      else stft_kernel<10><<<grid, 256, 0, stream>>>(p, groups, n_work);
      break;
    case 4096: stft_kernel<11><<<grid, 512, 0, stream>>>(p, groups, n_work); break;
    default: set_error("stft: n_fft must be 512, 1024, 2048 or 4096"); return 2;
  }
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Inverse STFT with fused mask multiply, window, overlap-add in shared memory, envelope division and
// crop.  grid = (ceil(length / S), B) with S = hops_per_cta * hop output samples per CTA; each CTA
// inverse-transforms every frame overlapping its segment (sequentially, so the OLA needs no atomics).
// -------------------------------------------------------------------------------------------------
template <int LOG2NC>
__global__ void __launch_bounds__((1 << LOG2NC) / 4) istft_kernel(IstftParams p, int segs, int n_work) {
  constexpr int NC = 1 << LOG2NC;
  constexpr int T4 = NC / 4;
  constexpr int NFFT = 2 * NC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* sa = reinterpret_cast<float2*>(smem_raw);
  float2* sb = sa + NC;
  float* ola = reinterpret_cast<float*>(sb + NC);
  const int j = threadIdx.x;
  const int S = p.hops_per_cta * p.hop;
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {  // work item = (batch item, output segment)
  const int b = work / segs;
  const int s0 = (work - b * segs) * S;  // first output sample of this segment
  for (int i = j; i < S; i += T4) ola[i] = 0.0f;
  // frames whose support [t*hop - NC, t*hop + NC) (output coordinates) intersects [s0, s0 + S)
  // frame t covers output samples [t*hop - frame_off, t*hop - frame_off + NFFT)
  const int lo_num = s0 + p.frame_off - NFFT;  // t*hop > lo_num
  int t_lo = lo_num < 0 ? 0 : lo_num / p.hop + 1;
  int t_hi = (s0 + S + p.frame_off + p.hop - 1) / p.hop - 1;
  if (t_hi > p.F - 1) t_hi = p.F - 1;
  const float inv = p.scale / (float)NC;
  for (int t = t_lo; t <= t_hi; ++t) {
    const size_t m = (size_t)b * p.F + t;
    const float2* __restrict__ Zr = p.Z + m * p.ldz;
    const float* __restrict__ Mr = p.mask ? p.mask + m * p.ldm : nullptr;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = j + r * T4;  // k in [0, NC/2)
      float2 xk = Zr[k], xn = (NC - k < p.nbins) ? Zr[NC - k] : make_float2(0.f, 0.f);
      if (Mr) {
        const float mk = Mr[k], mn = (NC - k < p.nbins) ? Mr[NC - k] : 0.f;
        xk.x *= mk; xk.y *= mk; xn.x *= mn; xn.y *= mn;
      }
      if (k == 0) {  // irfft ignores the imaginary part of the DC and Nyquist bins
        xk.y = 0.0f;
        xn.y = 0.0f;
        sa[0] = irfft_pre(xk, xn, p.tw[0]);
      } else {
        sa[k] = irfft_pre(xk, xn, p.tw[k]);
        sa[NC - k] = irfft_pre(xn, xk, p.tw[NC - k]);
      }
    }
    if (j == 0) {
      float2 xh = Zr[NC / 2];
      if (Mr) { const float mh = Mr[NC / 2]; xh.x *= mh; xh.y *= mh; }
      sa[NC / 2] = irfft_pre(xh, xh, p.tw[NC / 2]);
    }
    const float2* res = fft_block<LOG2NC>(sa, sb, p.tw, j);
    const int off = t * p.hop - p.frame_off - s0;  // segment-relative position of frame sample 0
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = j + r * T4;
      const float2 v = res[n];
      const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
      const int q = off + 2 * n;
      if (q >= 0 && q < S) ola[q] += v.x * inv * w.x;
      if (q + 1 >= 0 && q + 1 < S) ola[q + 1] += -v.y * inv * w.y;
    }
    __syncthreads();
  }
  // envelope sum_t w^2 (torch.istft window_envelop), then crop to `length`
  float* __restrict__ out = p.out + (size_t)b * p.out_bstride;
  for (int i = j; i < S; i += T4) {
    const int s = s0 + i;
    if (s >= p.length) break;
    // window envelope: frames t in [-env_pad, F + env_pad) with 0 <= q - t*hop < NFFT, q = s + frame_off
    const int q = s + p.frame_off + p.env_pad * p.hop;  // shift so that frame indices start at 0
    const int Fe = p.F + 2 * p.env_pad;
    int ta = (q - NFFT + p.hop) / p.hop;  // ceil((q - NFFT + 1) / hop) for q - NFFT + 1 >= 0
    if (q - NFFT + 1 <= 0) ta = 0;
    int tb = q / p.hop;
    if (tb > Fe - 1) tb = Fe - 1;
    float env = 0.0f;
    for (int t = ta; t <= tb; ++t) {
      const float w = p.window[q - t * p.hop];
      env += w * w;
    }
    out[s] = (env > 1e-11f) ? ola[i] / env : 0.0f;
  }
  __syncthreads();  // ola is zeroed again by the next work item
  }
}

// n_fft = 2048 version: the frames overlapping the CTA's output segment are inverse-transformed FOUR at a time (one per 64-thread
// group, fft1024_x4), so four frames' spectrum / mask loads are in flight together; the overlap-add of the four results is done by
// all 256 threads, frame after frame in a fixed order (deterministic, no atomics).
__global__ void __launch_bounds__(256) istft2048_kernel(IstftParams p, int segs, int n_work) {
  constexpr int NC = 1024, NFFT = 2048;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);          // [4][FFT1024_BUF]
  float* ola = reinterpret_cast<float*>(buf + 4 * FFT1024_BUF);  // [S]
  const int tid = threadIdx.x, g = tid >> 6, tl = tid & 63;
  float2* fb = buf + g * FFT1024_BUF;
  const int S = p.hops_per_cta * p.hop;
  float* env_int = ola + S;  // [hop] window envelope of an interior sample, by q mod hop (all NFFT / hop frames present)
  const float inv = p.scale / (float)NC;
  const int fph = NFFT / p.hop;  // frames covering an interior sample
  for (int r = tid; r < p.hop; r += 256) {
    float e = 0.0f;
    for (int j = 0; j < fph; ++j) {
      const float w = p.window[r + j * p.hop];
      e += w * w;
    }
    env_int[r] = e;
  }
  const bool pair_ok = ((p.hop | p.frame_off | S) & 1) == 0;  // frame samples (2n, 2n+1) land on an even output index: float2 overlap-add
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {  // work item = (batch item, output segment)
    const int b = work / segs;
    const int s0 = (work - b * segs) * S;  // first output sample of this segment
    for (int i = tid; i < S; i += 256) ola[i] = 0.0f;
    // frame t covers output samples [t*hop - frame_off, t*hop - frame_off + NFFT): those intersecting [s0, s0 + S)
    const int lo_num = s0 + p.frame_off - NFFT;  // t*hop > lo_num
    const int t_lo = lo_num < 0 ? 0 : lo_num / p.hop + 1;
    int t_hi = (s0 + S + p.frame_off + p.hop - 1) / p.hop - 1;
    if (t_hi > p.F - 1) t_hi = p.F - 1;
    for (int tb = t_lo; tb <= t_hi; tb += 4) {
      const int t = tb + g;
      if (t <= t_hi) {
        const size_t m = (size_t)b * p.F + t;
        const float2* __restrict__ Zr = p.Z + m * p.ldz;
        const float* __restrict__ Mr = p.mask ? p.mask + m * p.ldm : nullptr;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int k = tl + 64 * r;  // k in [0, NC/2)
          float2 xk = Zr[k], xn = (NC - k < p.nbins) ? Zr[NC - k] : make_float2(0.f, 0.f);
          if (Mr) {
            const float mk = Mr[k], mn = (NC - k < p.nbins) ? Mr[NC - k] : 0.f;
            xk.x *= mk; xk.y *= mk; xn.x *= mn; xn.y *= mn;
          }
          if (k == 0) {  // irfft ignores the imaginary part of the DC and Nyquist bins
            xk.y = 0.0f;
            xn.y = 0.0f;
            fb[0] = irfft_pre(xk, xn, p.tw[0]);
          } else {
            fb[k] = irfft_pre(xk, xn, p.tw[k]);
            fb[NC - k] = irfft_pre(xn, xk, p.tw[NC - k]);
          }
        }
        if (tl == 0) {
          float2 xh = Zr[NC / 2];
          if (Mr) { const float mh = Mr[NC / 2]; xh.x *= mh; xh.y *= mh; }
          fb[NC / 2] = irfft_pre(xh, xh, p.tw[NC / 2]);
        }
      }
      fft1024_x4(buf, p.tw, tid);
      for (int gg = 0; gg < 4 && tb + gg <= t_hi; ++gg) {  // overlap-add, one frame at a time (CTA-uniform loop)
        const float2* res = buf + gg * FFT1024_BUF;
        const int off = (tb + gg) * p.hop - p.frame_off - s0;  // segment-relative position of frame sample 0
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int n = tid + r * 256;
          const float2 v = res[n];
          const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
          const int q = off + 2 * n;
          if (pair_ok) {
            if (q >= 0 && q < S) {
              float2* o2 = reinterpret_cast<float2*>(ola + q);
              float2 acc = *o2;
              acc.x += v.x * inv * w.x;
              acc.y += -v.y * inv * w.y;
              *o2 = acc;
            }
          } else {
            if (q >= 0 && q < S) ola[q] += v.x * inv * w.x;
            if (q + 1 >= 0 && q + 1 < S) ola[q + 1] += -v.y * inv * w.y;
          }
        }
        __syncthreads();
      }
    }
    // envelope sum_t w^2 (torch.istft window_envelop), then crop to `length`
    float* __restrict__ out = p.out + (size_t)b * p.out_bstride;
    for (int i = tid; i < S; i += 256) {
      const int s = s0 + i;
      if (s >= p.length) break;
      const int q = s + p.frame_off + p.env_pad * p.hop;  // shift so that frame indices start at 0
      const int Fe = p.F + 2 * p.env_pad;
      const int tq = q / p.hop;
      float env;
      if (tq >= fph - 1 && tq <= Fe - 1) {  // interior: every frame tq - fph + 1 .. tq exists -> tabulated envelope
        env = env_int[q - tq * p.hop];
      } else {
        int ta = (q - NFFT + p.hop) / p.hop;  // ceil((q - NFFT + 1) / hop) for q - NFFT + 1 >= 0
        if (q - NFFT + 1 <= 0) ta = 0;
        const int tbb = tq > Fe - 1 ? Fe - 1 : tq;
        env = 0.0f;
        for (int tt = ta; tt <= tbb; ++tt) {
          const float w = p.window[q - tt * p.hop];
          env += w * w;
        }
      }
      out[s] = (env > 1e-11f) ? ola[i] / env : 0.0f;
    }
    __syncthreads();  // ola is zeroed again by the next work item
  }
}

// ---- n_fft = 2048, bulk-copy staged --------------------------------------------------------------------------------------
// Persistent CTAs; work item = (batch item, output segment of hops_per_cta hops), transformed in ROUNDS of four frames.  The four
// spectrum rows of a round are contiguous in Z and arrive by ONE cp.async.bulk (UBLKCP); the mask values of the next round are
// prefetched into registers.  Both are issued as soon as the current round's spectra have been consumed, so they are in flight
// during its FFT and overlap-add.  The spectrum is unpacked in (k, NC - k) pairs.  OWN (hop == 512): a thread always meets the same
// output positions (index == tid mod 256), so the four frames of a round are summed in registers with the synthesis window (and
// the 1/N scale) held in registers, then added to the segment once -- no barrier between the frames, a fixed order: deterministic.
template <bool OWN>
__global__ void __launch_bounds__(256, 2) istft2048_tma_kernel(IstftParams p, int segs, int n_work) {
  constexpr int NC = 1024, NFFT = 2048;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);                    // [4][FFT1024_BUF]
  float2* zst = buf + 4 * FFT1024_BUF;                                  // [4][ldz] staged spectrum rows
  float* ola = reinterpret_cast<float*>(zst + 4 * (size_t)p.ldz);       // [S]
  const int S = p.hops_per_cta * p.hop;
  float* env_int = ola + S;  // [hop] window envelope of an interior sample, by q mod hop (all NFFT / hop frames present)
  float2* tws = reinterpret_cast<float2*>(env_int + p.hop);  // [NFFT] twiddles (the L1 left beside this much shared memory is too small)
  uint64_t* bar = reinterpret_cast<uint64_t*>(tws + NFFT);
  const int tid = threadIdx.x, g = tid >> 6, tl = tid & 63;
  float2* fb = buf + g * FFT1024_BUF;
  for (int i = tid; i < NFFT; i += 256) tws[i] = p.tw[i];
  const float inv = 0.5f * p.scale / (float)NC;  // the unpacking below returns twice the packed spectrum
  const int fph = NFFT / p.hop;                  // frames covering an interior sample
  for (int r = tid; r < p.hop; r += 256) {
    float e = 0.0f;
    for (int j = 0; j < fph; ++j) {
      const float w = p.window[r + j * p.hop];
      e += w * w;
    }
    env_int[r] = e;
  }
  for (int i = tid; i < S; i += 256) ola[i] = 0.0f;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  float2 wv[4];  // synthesis window (times the scale) of the frame elements this thread overlap-adds: n = tid + 256 jj
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * (tid + 256 * jj));
    wv[jj] = make_float2(w.x * inv, -w.y * inv);
  }
  __syncthreads();

  // round state: (work, tb) = first frame of the round; the frames of a work item are t_lo .. t_hi
  auto item = [&](int work, int& b, int& s0, int& t_lo, int& t_hi) {
    b = work / segs;
    s0 = (work - b * segs) * S;
    const int lo_num = s0 + p.frame_off - NFFT;  // frame t covers output samples [t*hop - frame_off, t*hop - frame_off + NFFT)
    t_lo = lo_num < 0 ? 0 : lo_num / p.hop + 1;
    t_hi = (s0 + S + p.frame_off + p.hop - 1) / p.hop - 1;
    if (t_hi > p.F - 1) t_hi = p.F - 1;
  };
  float mk[8], mn[8], mh = 1.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mk[i] = mn[i] = 1.0f;
  auto fetch = [&](int b, int tb, int t_hi) {  // start the loads of a round: spectrum rows (one thread), mask values (all)
    if (tid == 0) {
      const int nf = min(4, t_hi - tb + 1);
      const uint32_t bytes = (uint32_t)nf * (uint32_t)p.ldz * 8u;
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(zst, p.Z + ((size_t)b * p.F + tb) * p.ldz, bytes, bar);
    }
    if (p.mask && tb + g <= t_hi) {
      const float* __restrict__ Mr = p.mask + ((size_t)b * p.F + tb + g) * p.ldm;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = tl + 64 * i;
        mk[i] = __ldg(Mr + k);
        mn[i] = (NC - k < p.nbins) ? __ldg(Mr + NC - k) : 0.0f;
      }
      if (tl == 0) mh = __ldg(Mr + NC / 2);
    }
  };

  int work = blockIdx.x;
  if (work >= n_work) return;
  int b, s0, t_lo, t_hi;
  item(work, b, s0, t_lo, t_hi);
  int tb = t_lo;
  fetch(b, tb, t_hi);
  uint32_t phase = 0;
  while (true) {
    // ---- unpack the round's spectra (times the mask) into the packed inverse-FFT inputs ----
    mbar_wait(bar, phase);
    phase ^= 1u;
    if (tb + g <= t_hi) {
      const float2* __restrict__ zr = zst + (size_t)g * p.ldz;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = tl + 64 * i;  // [0, NC/2)
        float2 xk = zr[k], xn = (NC - k < p.nbins) ? zr[NC - k] : make_float2(0.f, 0.f);
        xk.x *= mk[i]; xk.y *= mk[i]; xn.x *= mn[i]; xn.y *= mn[i];
        if (k == 0) {  // irfft ignores the imaginary part of the DC and Nyquist bins
          xk.y = 0.0f;
          xn.y = 0.0f;
        }
        float2 zk2, zn2;
        irfft_pre_pair2(xk, xn, tws[k], zk2, zn2);
        fb[k] = zk2;
        if (k != 0) fb[NC - k] = zn2;
      }
      if (tl == 0) {
        float2 xh = zr[NC / 2];
        xh.x *= mh; xh.y *= mh;
        float2 zk2, zn2;
        irfft_pre_pair2(xh, xh, tws[NC / 2], zk2, zn2);
        fb[NC / 2] = zk2;
      }
    }
    __syncthreads();  // the staged rows are consumed, the frame buffers complete
    // ---- next round (possibly of the next work item): its loads fly during this round's FFT and overlap-add ----
    int n_work_i = work, nb = b, ns0 = s0, nt_lo = t_lo, nt_hi = t_hi, ntb = tb + 4;
    bool more = true;
    if (ntb > t_hi) {
      n_work_i = work + gridDim.x;
      more = n_work_i < n_work;
      if (more) {
        item(n_work_i, nb, ns0, nt_lo, nt_hi);
        ntb = nt_lo;
      }
    }
    if (more) fetch(nb, ntb, nt_hi);
    fft1024_x4<false, true>(buf, tws, tid);
    __syncthreads();  // all four frames transformed
    // ---- window + overlap-add into the segment ----
    if (OWN) {
      float2 acc[7];
#pragma unroll
      for (int r = 0; r < 7; ++r) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        if (tb + gg <= t_hi) {  // CTA-uniform
          const float2* res = buf + gg * FFT1024_BUF;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float2 v = res[tid + 256 * jj];
            acc[gg + jj].x = fmaf(v.x, wv[jj].x, acc[gg + jj].x);
            acc[gg + jj].y = fmaf(v.y, wv[jj].y, acc[gg + jj].y);
          }
        }
      }
      const int P0 = (tb * p.hop - p.frame_off - s0) / 2 + tid;  // float2 index of frame tb's element tid (exact: all terms are even)
      float2* ola2 = reinterpret_cast<float2*>(ola);
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const int P = P0 + 256 * r;
        if (P >= 0 && 2 * P < S) {
          float2 o = ola2[P];
          o.x += acc[r].x;
          o.y += acc[r].y;
          ola2[P] = o;
        }
      }
    } else {
      for (int gg = 0; gg < 4 && tb + gg <= t_hi; ++gg) {  // one frame at a time (CTA-uniform loop)
        const float2* res = buf + gg * FFT1024_BUF;
        const int off = (tb + gg) * p.hop - p.frame_off - s0;  // segment-relative position of frame sample 0
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float2 v = res[tid + 256 * jj];
          const int q = off + 2 * (tid + 256 * jj);
          if (q >= 0 && q < S) ola[q] += v.x * wv[jj].x;
          if (q + 1 >= 0 && q + 1 < S) ola[q + 1] += v.y * wv[jj].y;
        }
        __syncthreads();
      }
    }
    if (tb + 4 > t_hi) {
      // ---- last round of the segment: envelope sum_t w^2 (torch.istft window_envelop), crop to `length`, clear for the next ----
      __syncthreads();
      float* __restrict__ out = p.out + (size_t)b * p.out_bstride;
      for (int i = tid; i < S; i += 256) {
        const int s = s0 + i;
        if (s >= p.length) break;
        const int q = s + p.frame_off + p.env_pad * p.hop;  // shift so that frame indices start at 0
        const int Fe = p.F + 2 * p.env_pad;
        const int tq = q / p.hop;
        float env;
        if (tq >= fph - 1 && tq <= Fe - 1) {  // interior: every frame tq - fph + 1 .. tq exists -> tabulated envelope
          env = env_int[q - tq * p.hop];
        } else {
          int ta = (q - NFFT + p.hop) / p.hop;  // ceil((q - NFFT + 1) / hop) for q - NFFT + 1 >= 0
          if (q - NFFT + 1 <= 0) ta = 0;
          const int tbb = tq > Fe - 1 ? Fe - 1 : tq;
          env = 0.0f;
          for (int tt = ta; tt <= tbb; ++tt) {
            const float w = p.window[q - tt * p.hop];
            env += w * w;
          }
        }
        out[s] = (env > 1e-11f) ? ola[i] / env : 0.0f;
      }
      __syncthreads();
      for (int i = tid; i < S; i += 256) ola[i] = 0.0f;  // ordered before the next accumulation by the barrier after the unpack
    } else if (OWN) {
      __syncthreads();  // the frame buffers are rewritten by the next unpack
    }
    if (!more) break;
    work = n_work_i; b = nb; s0 = ns0; t_lo = nt_lo; t_hi = nt_hi; tb = ntb;
  }
}

static bool istft2048_tma_ok(const IstftParams& p) {
  return stft_staged_enabled() && p.n_fft == 2048 && p.ldz % 2 == 0 && p.ldz >= p.nbins && ((uintptr_t)p.Z & 15) == 0 && 2048 % p.hop == 0 &&
         p.hop % 2 == 0 && p.frame_off % 2 == 0 && p.hop >= 128;
}

// hops per work item: the one that needs the fewest (waves of CTAs) x (rounds of four frames) on the SMs at hand
static int istft2048_choose_hops(const IstftParams& p, int B, int slots, size_t smem_fixed, size_t smem_cap) {
  const int fph = 2048 / p.hop;
  int best = 0;
  double best_cost = 0.0;
  const int total_hops = ceil_div(p.length, p.hop);
  for (int hpc = 4; hpc <= 64; ++hpc) {
    if (smem_fixed + (size_t)hpc * p.hop * 4 > smem_cap) break;
    const long long items = (long long)B * ceil_div(total_hops, hpc);
    const long long waves = (items + slots - 1) / slots;
    const int rounds = ceil_div(hpc + fph - 1, 4);
    const double cost = (double)waves * (rounds + 0.02 * hpc + 0.3);
    if (best == 0 || cost < best_cost - 1e-9) { best = hpc; best_cost = cost; }
  }
  return best;
}

static int launch_istft2048_tma(const IstftParams& p_in, int B, cudaStream_t stream) {
  IstftParams p = p_in;
  const bool own = p.hop == 512 && p.frame_off % 512 == 0;
  auto kern = own ? istft2048_tma_kernel<true> : istft2048_tma_kernel<false>;
  const size_t smem_fixed = sizeof(float2) * (4 * FFT1024_BUF + 2048) + sizeof(float2) * 4 * (size_t)p.ldz + sizeof(float) * p.hop + 16;
  const int sms = p.max_sms > 0 ? p.max_sms : (p.sms_avail > 0 ? p.sms_avail : device_sms());
  const size_t smem_cap = (227 * 1024) / 2 - 1024;  // two CTAs per SM
  static const int force_hpc = [] { const char* e = getenv("RFX_ISTFT_HOPS"); return e ? atoi(e) : 0; }();
  int hpc = force_hpc > 0 ? force_hpc : istft2048_choose_hops(p, B, 2 * sms, smem_fixed, smem_cap);
  RFX_REQUIRE(hpc > 0, "istft: no segment size fits shared memory");
  if (own) hpc = std::max(1, hpc);
  p.hops_per_cta = hpc;
  const int S = hpc * p.hop;
  const size_t smem = smem_fixed + sizeof(float) * S;
  RFX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int segs = ceil_div(p.length, S);
  const long long n_work_ll = (long long)segs * B;
  RFX_REQUIRE(n_work_ll < (1ll << 31), "istft: too many segments");
  int per_sm = 0;
  RFX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
  const long long cap = (long long)sms * (per_sm > 0 ? per_sm : 1);
  const int grid = (int)std::min<long long>(n_work_ll, cap);
  kern<<<grid, 256, smem, stream>>>(p, segs, (int)n_work_ll);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_istft2048(const IstftParams& p, int B, cudaStream_t stream) {
  const int S = p.hops_per_cta * p.hop;
  const size_t smem = sizeof(float2) * 4 * FFT1024_BUF + sizeof(float) * (S + p.hop);
  RFX_REQUIRE(2048 % p.hop == 0, "istft: hop must divide n_fft");
  RFX_CHECK_CUDA(cudaFuncSetAttribute(istft2048_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int segs = ceil_div(p.length, S);
  const long long n_work_ll = (long long)segs * B;
  RFX_REQUIRE(n_work_ll < (1ll << 31), "istft: too many segments");
  const int n_work = (int)n_work_ll;
  int grid = n_work;
  if (p.max_sms > 0) {
    int per_sm = 0;
    RFX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, istft2048_kernel, 256, smem));
    const long long cap = (long long)p.max_sms * (per_sm > 0 ? per_sm : 1);
    if (cap < grid) grid = (int)cap;
  }
  istft2048_kernel<<<grid, 256, smem, stream>>>(p, segs, n_work);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int LOG2NC>
static int launch_istft_t(const IstftParams& p, int B, cudaStream_t stream) {
  constexpr int NC = 1 << LOG2NC;
  const int S = p.hops_per_cta * p.hop;
  const size_t smem = sizeof(float2) * 2 * NC + sizeof(float) * S;
  RFX_CHECK_CUDA(cudaFuncSetAttribute(istft_kernel<LOG2NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int segs = ceil_div(p.length, S);
  const long long n_work_ll = (long long)segs * B;
  RFX_REQUIRE(n_work_ll < (1ll << 31), "istft: too many segments");
  const int n_work = (int)n_work_ll;
  int grid = n_work;
  if (p.max_sms > 0) {
    int per_sm = 0;
    RFX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, istft_kernel<LOG2NC>, NC / 4, smem));
    const long long cap = (long long)p.max_sms * (per_sm > 0 ? per_sm : 1);
    if (cap < grid) grid = (int)cap;
  }
  istft_kernel<LOG2NC><<<grid, NC / 4, smem, stream>>>(p, segs, n_work);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_istft(const IstftParams& p, int B, cudaStream_t stream) {
  RFX_REQUIRE(p.tw != nullptr, "twiddle table");
  RFX_REQUIRE(p.hops_per_cta > 0, "hops_per_cta");
  switch (p.n_fft) {
    case 512: return launch_istft_t<8>(p, B, stream);
    case 1024: return launch_istft_t<9>(p, B, stream);
    case 2048: {
      static const bool generic2048 = [] { const char* e = getenv("RFX_STFT_GENERIC"); return e && atoi(e) != 0; }();
      if (!generic2048 && istft2048_tma_ok(p)) return launch_istft2048_tma(p, B, stream);
      return generic2048 ? launch_istft_t<10>(p, B, stream) : launch_istft2048(p, B, stream);
    }
    case 4096: return launch_istft_t<11>(p, B, stream);
    default: set_error("istft: n_fft must be 512, 1024, 2048 or 4096"); return 2;
  }
}

}  // namespace rfx
