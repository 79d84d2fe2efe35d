// RemFx training/eval loss  L = MRSTFT(out, target) + 100 * L1(out, target)   (remfx/models.py:299,320,385)
// with auraloss.freq.MultiResolutionSTFTLoss defaults (fft 1024/2048/512, hop 120/240/50, win 600/1200/240,
// hann; per resolution: spectral convergence ||Y-X||_F/||Y||_F per item -> batch mean, plus
// mean |log X - log Y|; magnitudes sqrt(clamp(re^2+im^2, 1e-8))).  See oracle/loss.py for the restatement
// (auraloss itself is an un-vendored dependency: parity unpinned upstream).
//
// One kernel per resolution transforms frame t of BOTH signals in the same CTA and reduces the three
// sums on chip, so no spectrogram ever reaches HBM (algorithmic traffic: the two signals, read ~n_fft/hop
// times from L2).  Partial sums are written per CTA and reduced in a fixed order in fp64 -> deterministic.
#include "kernels.h"
#include "fft.cuh"
#include "../../include/remfx_b200.h"

namespace rfx {

constexpr int LOSS_FPC = 8;  // frames per CTA

__device__ __forceinline__ int reflect_idx(int i, int T) {
  if (i < 0) i = -i;
  if (i >= T) i = 2 * (T - 1) - i;
  return i;
}

struct LossStftParams {
  const float* x;  // prediction
  const float* y;  // target
  long long x_bstride, y_bstride;
  int T, x_al8, y_al8;
  const float* window;
  const float2* tw;
  int hop, F;
  float* partials;  // [B][gridDim.x][3] = {sum (ym-xm)^2, sum ym^2, sum |log xm - log ym|}
};

template <int LOG2NC>
__global__ void __launch_bounds__((1 << LOG2NC) / 4) stft_loss_kernel(LossStftParams p) {
  constexpr int NC = 1 << LOG2NC;
  constexpr int T4 = NC / 4;
  constexpr int NFFT = 2 * NC;
  __shared__ float2 sa[NC];
  __shared__ float2 sb[NC];
  __shared__ float magx[NC + 1];
  __shared__ float red[3][32];
  const int j = threadIdx.x;
  const int b = blockIdx.y;
  float a_sc = 0.f, a_y2 = 0.f, a_lm = 0.f;
  const int f_end = min(p.F, (int)(blockIdx.x + 1) * LOSS_FPC);
  for (int f = blockIdx.x * LOSS_FPC; f < f_end; ++f) {
    const int base = f * p.hop - NC;
    const bool interior = (base >= 0) && (base + NFFT <= p.T);
#pragma unroll
    for (int sig = 0; sig < 2; ++sig) {
      const float* __restrict__ s = sig ? p.y + (size_t)b * p.y_bstride : p.x + (size_t)b * p.x_bstride;
      const bool al8 = sig ? p.y_al8 : p.x_al8;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + r * T4;
        const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
        float2 v;
        if (interior && al8) {
          v = *reinterpret_cast<const float2*>(s + base + 2 * n);
        } else {
          v.x = s[reflect_idx(base + 2 * n, p.T)];
          v.y = s[reflect_idx(base + 2 * n + 1, p.T)];
        }
        sa[n] = make_float2(v.x * w.x, v.y * w.y);
      }
      const float2* Zp = fft_block<LOG2NC>(sa, sb, p.tw, j);
      for (int k = j; k <= NC; k += T4) {
        const float2 X = rfft_post(Zp, p.tw, NC, k);
        const float mag = sqrtf(fmaxf(X.x * X.x + X.y * X.y, 1e-8f));
        if (sig == 0) {
          magx[k] = mag;  // same thread reads it back for the target pass: no sync needed for magx itself
        } else {
          const float xm = magx[k];
          const float d = mag - xm;
          a_sc += d * d;
          a_y2 += mag * mag;
          a_lm += fabsf(logf(xm) - logf(mag));
        }
      }
      __syncthreads();  // sa/sb reused
    }
  }
  // block reduction
  a_sc = warp_sum(a_sc);
  a_y2 = warp_sum(a_y2);
  a_lm = warp_sum(a_lm);
  const int warp = j >> 5, lane = j & 31;
  if (lane == 0) { red[0][warp] = a_sc; red[1][warp] = a_y2; red[2][warp] = a_lm; }
  __syncthreads();
  if (j < 3) {
    float s = 0.f;
    for (int w = 0; w < (T4 + 31) / 32; ++w) s += red[j][w];
    p.partials[((size_t)b * gridDim.x + blockIdx.x) * 3 + j] = s;
  }
}

__global__ void __launch_bounds__(256) l1_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, long long xbs, long long ybs,
                                                         int T, float* __restrict__ partials) {
  // grid = (nblk, B): partials[b * nblk + blk] = sum |x - y| over a strided slice
  __shared__ float red[8];
  const int b = blockIdx.y;
  const float* xr = x + (size_t)b * xbs;
  const float* yr = y + (size_t)b * ybs;
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) acc += fabsf(xr[i] - yr[i]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partials[(size_t)b * gridDim.x + blockIdx.x] = s;
  }
}

struct LossFinalParams {
  const float* part[3];  // per-resolution partials
  int nblk[3];
  int bins[3], F[3];
  const float* l1part;
  int l1blk;
  int B, T;
  float l1_weight;
  double* sums;   // [4][B][3] per-item totals (resolutions 0..2, then L1 in slot 0)
  float* result;  // [0] loss, [1] mrstft, [2] mean |x-y|, [3+2r] sc_r, [4+2r] lm_r
};

// grid = (B, 4): blockIdx.y = resolution (0..2) or 3 = L1.  Fixed strided assignment + fixed tree -> deterministic.
__global__ void __launch_bounds__(256) loss_reduce_kernel(LossFinalParams p) {
  __shared__ double red[3][256];
  const int b = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (r < 3) {
    const float* q = p.part[r] + (size_t)b * p.nblk[r] * 3;
    for (int k = tid; k < p.nblk[r]; k += 256) {
      a0 += q[3 * k];
      a1 += q[3 * k + 1];
      a2 += q[3 * k + 2];
    }
  } else {
    for (int k = tid; k < p.l1blk; k += 256) a0 += p.l1part[(size_t)b * p.l1blk + k];
  }
  red[0][tid] = a0; red[1][tid] = a1; red[2][tid] = a2;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) {
      red[0][tid] += red[0][tid + s];
      red[1][tid] += red[1][tid + s];
      red[2][tid] += red[2][tid + s];
    }
    __syncthreads();
  }
  if (tid < 3) p.sums[((size_t)r * p.B + b) * 3 + tid] = red[tid][0];
}

__global__ void loss_final_kernel(LossFinalParams p) {
  if (threadIdx.x != 0) return;
  double mr = 0.0;
  for (int r = 0; r < 3; ++r) {
    double sc_mean = 0.0, lm = 0.0;
    for (int b = 0; b < p.B; ++b) {
      const double* q = p.sums + ((size_t)r * p.B + b) * 3;
      sc_mean += sqrt(q[0]) / sqrt(q[1]);
      lm += q[2];
    }
    sc_mean /= p.B;
    lm /= (double)p.B * p.bins[r] * p.F[r];
    p.result[3 + 2 * r] = (float)sc_mean;
    p.result[4 + 2 * r] = (float)lm;
    mr += sc_mean + lm;
  }
  mr /= 3.0;
  double l1 = 0.0;
  for (int b = 0; b < p.B; ++b) l1 += p.sums[((size_t)3 * p.B + b) * 3];
  l1 /= (double)p.B * p.T;
  p.result[1] = (float)mr;
  p.result[2] = (float)l1;
  p.result[0] = (float)(mr + p.l1_weight * l1);
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the loss with respect to the prediction (the first link of the training step):
//   d/dX [ (1/3) sum_r ( mean_b ||Y-X||_F / ||Y||_F + mean |log X - log Y| ) ]  with X = sqrt(clamp(|STFT(out)|^2, 1e-8)),
// chained through the magnitude (zero where the clamp is active), the real FFT (adjoint = un-normalised inverse real FFT with the
// DC / Nyquist bins counted once), the window and the reflect-padded framing (overlapping frames and mirrored edge samples
// accumulate with atomic adds).  Per frame the CTA recomputes both spectra exactly as the forward kernel does -- nothing is
// stored between the passes except the per-item norms the forward already leaves in its workspace.
// ---------------------------------------------------------------------------------------------------------
struct LossBwdParams {
  const float* x;  // prediction
  const float* y;  // target
  long long x_bstride, y_bstride, g_bstride;
  int T, x_al8, y_al8;
  const float* window;
  const float2* tw;
  int hop, F, B;
  const double* sums;      // [B][3] of this resolution: sum (Y-X)^2, sum Y^2, (unused)
  const float* grad_loss;  // upstream scalar (device)
  float* grad;             // (B, T): accumulated into
};

template <int LOG2NC>
__global__ void __launch_bounds__((1 << LOG2NC) / 4) stft_loss_bwd_kernel(LossBwdParams p) {
  constexpr int NC = 1 << LOG2NC;
  constexpr int T4 = NC / 4;
  constexpr int NFFT = 2 * NC;
  __shared__ float2 sa[NC];
  __shared__ float2 sb[NC];
  __shared__ float2 hbuf[NC + 1];
  __shared__ float magy[NC + 1];
  const int j = threadIdx.x;
  const int b = blockIdx.y;
  const float gl = p.grad_loss[0];
  const double s0 = p.sums[(size_t)b * 3], s1 = p.sums[(size_t)b * 3 + 1];
  const float c_sc = (s0 > 0.0 && s1 > 0.0) ? (float)((double)gl / (3.0 * p.B) / (sqrt(s0) * sqrt(s1))) : 0.0f;
  const float c_lm = (float)((double)gl / (3.0 * (double)p.B * (double)(NC + 1) * (double)p.F));
  float* __restrict__ grow = p.grad + (size_t)b * p.g_bstride;
  const int f_end = min(p.F, (int)(blockIdx.x + 1) * LOSS_FPC);
  for (int f = blockIdx.x * LOSS_FPC; f < f_end; ++f) {
    const int base = f * p.hop - NC;
    const bool interior = (base >= 0) && (base + NFFT <= p.T);
#pragma unroll
    for (int sig = 1; sig >= 0; --sig) {  // target first (its magnitudes are needed while the prediction's spectrum is walked)
      const float* __restrict__ s = sig ? p.y + (size_t)b * p.y_bstride : p.x + (size_t)b * p.x_bstride;
      const bool al8 = sig ? p.y_al8 : p.x_al8;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + r * T4;
        const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
        float2 v;
        if (interior && al8) {
          v = *reinterpret_cast<const float2*>(s + base + 2 * n);
        } else {
          v.x = s[reflect_idx(base + 2 * n, p.T)];
          v.y = s[reflect_idx(base + 2 * n + 1, p.T)];
        }
        sa[n] = make_float2(v.x * w.x, v.y * w.y);
      }
      const float2* Zp = fft_block<LOG2NC>(sa, sb, p.tw, j);
      for (int k = j; k <= NC; k += T4) {
        const float2 X = rfft_post(Zp, p.tw, NC, k);
        const float pw = X.x * X.x + X.y * X.y;
        const float mag = sqrtf(fmaxf(pw, 1e-8f));
        if (sig == 1) {
          magy[k] = mag;  // read back by the same thread in the prediction pass
        } else {
          const float ym = magy[k];
          const float d = mag - ym;
          const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
          const float gm = c_sc * d + c_lm * sgn / mag;           // dL / d mag
          const float q = pw > 1e-8f ? gm / mag : 0.f;            // clamp active -> constant magnitude -> no gradient
          float2 H = make_float2(q * X.x, q * X.y);
          if (k == 0 || k == NC) H = make_float2(2.f * H.x, 0.f);  // bins without a mirror image count once
          hbuf[k] = H;
        }
      }
      __syncthreads();  // sa/sb reused; hbuf complete after the prediction pass
    }
    // adjoint of the real FFT: packed inverse transform of H (same pre-twiddle as the iSTFT)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int k = j + r * T4;  // k in [0, NC/2)
      const float2 xk = hbuf[k], xn = hbuf[NC - k];
      if (k == 0) {
        sa[0] = irfft_pre(xk, xn, p.tw[0]);
      } else {
        sa[k] = irfft_pre(xk, xn, p.tw[k]);
        sa[NC - k] = irfft_pre(xn, xk, p.tw[NC - k]);
      }
    }
    if (j == 0) sa[NC / 2] = irfft_pre(hbuf[NC / 2], hbuf[NC / 2], p.tw[NC / 2]);
    const float2* res = fft_block<LOG2NC>(sa, sb, p.tw, j);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = j + r * T4;
      const float2 v = res[n];
      const float2 w = *reinterpret_cast<const float2*>(p.window + 2 * n);
      const float g0 = v.x * w.x, g1 = -v.y * w.y;
      if (g0 != 0.f) atomicAdd(grow + reflect_idx(base + 2 * n, p.T), g0);
      if (g1 != 0.f) atomicAdd(grow + reflect_idx(base + 2 * n + 1, p.T), g1);
    }
    __syncthreads();  // sa/sb/hbuf reused by the next frame
  }
}

// grad = grad_loss * l1_weight / (B T) * sign(out - target): WRITES the gradient buffer (the spectral kernels then accumulate)
__global__ void __launch_bounds__(256) l1_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, long long xbs, long long ybs,
                                                     long long gbs, int B, int T, float l1_weight, const float* __restrict__ grad_loss,
                                                     float* __restrict__ grad) {
  const int b = blockIdx.y;
  const float c = grad_loss[0] * l1_weight / ((float)B * (float)T);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const float d = x[(size_t)b * xbs + i] - y[(size_t)b * ybs + i];
    grad[(size_t)b * gbs + i] = d > 0.f ? c : (d < 0.f ? -c : 0.f);
  }
}

static const int kRes[3][2] = {{1024, 120}, {2048, 240}, {512, 50}};  // (n_fft, hop); win lengths 600/1200/240 come via the windows
constexpr int L1_BLOCKS = 32;

static size_t loss_part_floats(int B, int T, int r) {
  const int F = T / kRes[r][1] + 1;
  return (size_t)B * ceil_div(F, LOSS_FPC) * 3;
}

}  // namespace rfx

using namespace rfx;

extern "C" {

size_t rfx_loss_workspace_bytes(int B, int T) {
  if (B <= 0 || T <= 0) return 0;
  size_t n = 0;
  for (int r = 0; r < 3; ++r) n += align_up(loss_part_floats(B, T, r) * 4, 256);
  n += align_up((size_t)B * L1_BLOCKS * 4, 256);
  n += align_up((size_t)4 * B * 3 * 8, 256);
  return n;
}

int rfx_remfx_loss(const float* out, long long out_bstride, const float* target, long long target_bstride, int B, int T,
                   const float* win1024, const float* win2048, const float* win512, float l1_weight, float* result, void* workspace,
                   size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(out && target && win1024 && win2048 && win512 && result && workspace, "null argument");
  RFX_REQUIRE(B > 0 && T > 1024, "need B > 0 and T > 1024 (reflect padding of the 2048-point STFT)");
  RFX_REQUIRE(workspace_bytes >= rfx_loss_workspace_bytes(B, T), "workspace too small (rfx_loss_workspace_bytes)");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const float* wins[3] = {win1024, win2048, win512};
  LossFinalParams fp{};
  size_t off = 0;
  for (int r = 0; r < 3; ++r) {
    const int n_fft = kRes[r][0], hop = kRes[r][1];
    LossStftParams p{};
    p.x = out; p.y = target; p.x_bstride = out_bstride; p.y_bstride = target_bstride;
    p.T = T;
    p.x_al8 = (((uintptr_t)out & 7) == 0 && out_bstride % 2 == 0) ? 1 : 0;
    p.y_al8 = (((uintptr_t)target & 7) == 0 && target_bstride % 2 == 0) ? 1 : 0;
    RFX_REQUIRE(((uintptr_t)wins[r] & 7) == 0, "windows must be 8-byte aligned");
    p.window = wins[r];
    p.tw = twiddles(n_fft);
    RFX_REQUIRE(p.tw != nullptr, "twiddle table allocation failed");
    p.hop = hop;
    p.F = T / hop + 1;
    p.partials = reinterpret_cast<float*>(ws + off);
    const int nblk = ceil_div(p.F, LOSS_FPC);
    fp.part[r] = p.partials; fp.nblk[r] = nblk; fp.bins[r] = n_fft / 2 + 1; fp.F[r] = p.F;
    off += align_up(loss_part_floats(B, T, r) * 4, 256);
    dim3 grid(nblk, B);
    if (n_fft == 1024) stft_loss_kernel<9><<<grid, 128, 0, s>>>(p);
    else if (n_fft == 2048) stft_loss_kernel<10><<<grid, 256, 0, s>>>(p);
    else stft_loss_kernel<8><<<grid, 64, 0, s>>>(p);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  float* l1part = reinterpret_cast<float*>(ws + off);
  l1_partial_kernel<<<dim3(L1_BLOCKS, B), 256, 0, s>>>(out, target, out_bstride, target_bstride, T, l1part);
  RFX_CHECK_CUDA(cudaGetLastError());
  off += align_up((size_t)B * L1_BLOCKS * 4, 256);
  fp.l1part = l1part; fp.l1blk = L1_BLOCKS; fp.B = B; fp.T = T; fp.l1_weight = l1_weight; fp.result = result;
  fp.sums = reinterpret_cast<double*>(ws + off);
  loss_reduce_kernel<<<dim3(B, 4), 256, 0, s>>>(fp);
  RFX_CHECK_CUDA(cudaGetLastError());
  loss_final_kernel<<<1, 32, 0, s>>>(fp);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int rfx_remfx_loss_backward(const float* out, long long out_bstride, const float* target, long long target_bstride, int B, int T,
                            const float* win1024, const float* win2048, const float* win512, float l1_weight, const float* grad_loss,
                            float* grad_out, long long grad_bstride, const void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(out && target && win1024 && win2048 && win512 && grad_loss && grad_out && workspace, "null argument");
  RFX_REQUIRE(B > 0 && T > 1024, "need B > 0 and T > 1024 (reflect padding of the 2048-point STFT)");
  RFX_REQUIRE(workspace_bytes >= rfx_loss_workspace_bytes(B, T), "workspace too small: pass the workspace rfx_remfx_loss ran in");
  cudaStream_t s = (cudaStream_t)stream;
  const uint8_t* ws = reinterpret_cast<const uint8_t*>(workspace);
  size_t off = 0;
  for (int r = 0; r < 3; ++r) off += align_up(loss_part_floats(B, T, r) * 4, 256);
  off += align_up((size_t)B * L1_BLOCKS * 4, 256);
  const double* sums = reinterpret_cast<const double*>(ws + off);  // [4][B][3], left there by the forward's reduction
  l1_bwd_kernel<<<dim3(64, B), 256, 0, s>>>(out, target, out_bstride, target_bstride, grad_bstride, B, T, l1_weight, grad_loss, grad_out);
  RFX_CHECK_CUDA(cudaGetLastError());
  const float* wins[3] = {win1024, win2048, win512};
  for (int r = 0; r < 3; ++r) {
    const int n_fft = kRes[r][0], hop = kRes[r][1];
    LossBwdParams p{};
    p.x = out; p.y = target; p.x_bstride = out_bstride; p.y_bstride = target_bstride; p.g_bstride = grad_bstride;
    p.T = T; p.B = B;
    p.x_al8 = (((uintptr_t)out & 7) == 0 && out_bstride % 2 == 0) ? 1 : 0;
    p.y_al8 = (((uintptr_t)target & 7) == 0 && target_bstride % 2 == 0) ? 1 : 0;
    RFX_REQUIRE(((uintptr_t)wins[r] & 7) == 0, "windows must be 8-byte aligned");
    p.window = wins[r];
    p.tw = twiddles(n_fft);
    RFX_REQUIRE(p.tw != nullptr, "twiddle table allocation failed");
    p.hop = hop;
    p.F = T / hop + 1;
    p.sums = sums + (size_t)r * B * 3;
    p.grad_loss = grad_loss;
    p.grad = grad_out;
    dim3 grid(ceil_div(p.F, LOSS_FPC), B);
    if (n_fft == 1024) stft_loss_bwd_kernel<9><<<grid, 128, 0, s>>>(p);
    else if (n_fft == 2048) stft_loss_bwd_kernel<10><<<grid, 256, 0, s>>>(p);
    else stft_loss_bwd_kernel<8><<<grid, 64, 0, s>>>(p);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// L3  SI-SDR metric (auraloss.time.SISDRLoss(zero_mean=True, eps=1e-8, reduction="mean"); oracle/loss.py):
//   x, y zero-meaned per item; alpha = <x,y>/(|y|^2 + eps); loss = -mean_b 10 log10(|alpha y|^2 / (|x - alpha y|^2 + eps) + eps)
// Five raw sums per item are accumulated in fp64 (fixed order: deterministic), the rest is closed form.
// ---------------------------------------------------------------------------------------------------------
namespace rfx {
constexpr int SISDR_BLOCKS = 64;

__global__ void __launch_bounds__(256) sisdr_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, long long xbs, long long ybs,
                                                            int T, double* __restrict__ partials /*[B][SISDR_BLOCKS][5]*/) {
  __shared__ double red[5][8];
  const int b = blockIdx.y;
  const float* xr = x + (size_t)b * xbs;
  const float* yr = y + (size_t)b * ybs;
  double s[5] = {0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    const double a = xr[i], c = yr[i];
    s[0] += a; s[1] += c; s[2] += a * c; s[3] += a * a; s[4] += c * c;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    partials[((size_t)b * gridDim.x + blockIdx.x) * 5 + threadIdx.x] = t;
  }
}

__global__ void sisdr_final_kernel(const double* __restrict__ partials, int B, int T, float* __restrict__ result) {
  if (threadIdx.x != 0) return;
  const double eps = 1e-8;
  double acc = 0.0;
  for (int b = 0; b < B; ++b) {
    double s[5] = {0, 0, 0, 0, 0};
    for (int k = 0; k < SISDR_BLOCKS; ++k)
      for (int j = 0; j < 5; ++j) s[j] += partials[((size_t)b * SISDR_BLOCKS + k) * 5 + j];
    const double n = (double)T, mx = s[0] / n, my = s[1] / n;
    const double sxy = s[2] - n * mx * my, sxx = s[3] - n * mx * mx, syy = s[4] - n * my * my;
    const double alpha = sxy / (syy + eps);
    const double et = alpha * alpha * syy;
    const double er = sxx - 2.0 * alpha * sxy + alpha * alpha * syy;
    acc += 10.0 * log10(et / (er + eps) + eps);
  }
  result[0] = (float)(-acc / B);
}
}  // namespace rfx

extern "C" {

size_t rfx_sisdr_workspace_bytes(int B) { return B > 0 ? (size_t)B * rfx::SISDR_BLOCKS * 5 * sizeof(double) : 0; }

int rfx_sisdr_loss(const float* x, long long x_bstride, const float* y, long long y_bstride, int B, int T, float* result, void* workspace,
                   size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(x && y && result && workspace, "null argument");
  RFX_REQUIRE(B > 0 && T > 0, "positive sizes");
  RFX_REQUIRE(workspace_bytes >= rfx_sisdr_workspace_bytes(B) && ((uintptr_t)workspace & 7) == 0, "workspace too small or misaligned");
  cudaStream_t s = (cudaStream_t)stream;
  double* part = reinterpret_cast<double*>(workspace);
  rfx::sisdr_partial_kernel<<<dim3(rfx::SISDR_BLOCKS, B), 256, 0, s>>>(x, y, x_bstride, y_bstride, T, part);
  RFX_CHECK_CUDA(cudaGetLastError());
  rfx::sisdr_final_kernel<<<1, 32, 0, s>>>(part, B, T, result);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
