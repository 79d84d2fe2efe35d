// Open-Unmix training step (SURVEY row L5 for this network): what `loss.backward()` does to `OpenUnmixModel.forward`'s output in
// the reference's Lightning step (remfx/models.py:217-221, 294-301) with the network in TRAINING mode:
//   * BatchNorm1d with batch statistics (umx/openunmix/model.py:135, 151, 157) -- the batch mean / biased variance of the three
//     layers are returned so that the host mirror can move the running statistics exactly as torch does;
//   * LSTM dropout 0.4 between the layers (model.py:62-69): the masks come from the caller (torch's generator on the host side of
//     the ABI: no RNG stream of the reference can be matched bit for bit anyway, and the parity tests need to inject theirs);
//   * the reference's extra pass `Y = self.model(spectrogram(x))` (remfx/models.py:296-297) whose only effect is on the running
//     statistics: `pow_pass = 1` runs the network on ((|Z| + 1e-8)^alpha + mean) * scale and stops after the third BatchNorm.
// Forward = the inference launch sequence with the epilogue fusions undone where a pre-activation must be kept (dense layer ->
// fp32 pre-BatchNorm rows -> column statistics -> normalise + activation), backward = reverse order on the shared building blocks of
// bwd_common.h (tcgen05 weight-gradient contraction, transposed-pack input-gradient GEMMs, persistent BPTT chain, iSTFT adjoint).
#include "umx_internal.h"
#include "bwd_common.h"

#include <algorithm>
#include <cstring>

namespace rfx {
namespace {

constexpr float kBnEps = 1e-5f;

// ------------------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------------------
// Column sums and sums of squares of a row-major fp32 matrix, in double (BatchNorm batch statistics; torch accumulates in fp32 with
// a Welford merge -- double sums are at least as accurate).  acc[2 n] += sum x, acc[2 n + 1] += sum x^2.
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ X, long long rows, int ld, int N, double* __restrict__ acc,
                                                        int rows_per_cta) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  double s = 0.0, ss = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const double v = (double)X[(size_t)r * ld + c];
    s += v;
    ss = fma(v, v, ss);
  }
  atomicAdd(acc + 2 * c, s);
  atomicAdd(acc + 2 * c + 1, ss);
}

// stats[0..N) = mean, [N..2N) = biased variance, [2N..3N) = 1 / sqrt(var + eps)
__global__ void bn_finalize_kernel(const double* __restrict__ acc, long long rows, int N, float eps, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const double m = acc[2 * c] / (double)rows;
  double var = acc[2 * c + 1] / (double)rows - m * m;
  if (var < 0.0) var = 0.0;
  stats[c] = (float)m;
  stats[N + c] = (float)var;
  stats[2 * N + c] = (float)(1.0 / sqrt(var + (double)eps));
}

enum BnAct { BN_TANH = 1, BN_RELU = 2, BN_MASK = 3 };

// y = act((x - mean) rstd gamma + beta).  BN_TANH / BN_RELU -> split planes (row stride ld_out); BN_MASK: the third BatchNorm
// followed by `* output_scale + output_mean` and the ReLU (model.py:157-164) -> fp32 mask rows.
template <int ACT>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ X, long long rows, int ld, int N, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ oscale, const float* __restrict__ omean,
                                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ outf,
                                                       int ld_out) {
  const long long total = rows * (long long)N;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / N;
    const int c = (int)(i - r * N);
    const float xh = (X[(size_t)r * ld + c] - stats[c]) * stats[2 * N + c];
    float y = fmaf(xh, gamma[c], beta[c]);
    if (ACT == BN_TANH) y = tanhf(y);
    if (ACT == BN_RELU) y = fmaxf(y, 0.0f);
    if (ACT == BN_MASK) {
      y = fmaxf(fmaf(y, oscale[c], omean[c]), 0.0f);
      outf[(size_t)r * ld_out + c] = y;
    } else {
      __nv_bfloat16 h, l;
      split_bf16(y, h, l);
      hi[(size_t)r * ld_out + c] = h;
      lo[(size_t)r * ld_out + c] = l;
    }
  }
}

// dropout between LSTM layers: planes (hi + lo) * mask -> planes
__global__ void __launch_bounds__(256) dropout_planes_kernel(const __nv_bfloat16* __restrict__ shi, const __nv_bfloat16* __restrict__ slo, int ld_in,
                                                             const float* __restrict__ mask, long long rows, int N, __nv_bfloat16* __restrict__ dhi,
                                                             __nv_bfloat16* __restrict__ dlo, int ld_out) {
  const long long total = rows * (long long)N;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / N;
    const int c = (int)(i - r * N);
    const size_t si = (size_t)r * ld_in + c;
    const float v = (__bfloat162float(shi[si]) + __bfloat162float(slo[si])) * mask[i];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    dhi[(size_t)r * ld_out + c] = h;
    dlo[(size_t)r * ld_out + c] = l;
  }
}

__global__ void __launch_bounds__(256) mul_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) a[i] *= b[i];
}

// Backward through  y = act(BN(x))  (and, for BN_MASK, through mask = relu(BN(x) os + om), est = mask * Z):
// the gradient arriving at the BatchNorm output, per element.
struct BnBwd {
  const float* X; int ld;            // pre-BatchNorm rows (kept by the forward)
  long long rows; int N;
  const float* stats; const float* gamma; const float* beta;
  const float* oscale; const float* omean;   // BN_MASK
  const float* dA; int ldA;          // gradient w.r.t. the activation output (BN_TANH / BN_RELU)
  const float* dB; int ldB;          // optional second addend of that gradient (skip connection)
  const float2* Z; const float2* gZ; int ldz;  // BN_MASK: mixture spectrum and FFT(window * dout / envelope)
  int NC; float inv_n;               // BN_MASK: n_fft / 2 and 1 / n_fft
  double* sums;                      // [2 N]: sum g, sum g xhat  (g = gradient at the BatchNorm output; BN_MASK: at u = BN os + om)
  __nv_bfloat16* ghi; __nv_bfloat16* glo; int ldg;  // pass 2: d x as split planes
};

template <int ACT>
__device__ __forceinline__ float bn_bwd_g(const BnBwd& p, long long r, int c, float xh) {
  const float y = fmaf(xh, p.gamma[c], p.beta[c]);
  if (ACT == BN_MASK) {
    const float u = fmaf(y, p.oscale[c], p.omean[c]);
    if (!(u > 0.0f)) return 0.0f;
    const float2 z = p.Z[(size_t)r * p.ldz + c], g = p.gZ[(size_t)r * p.ldz + c];
    // d mask = Re(conj(Z) dY) with dY = c_k / N FFT(.)_k; irfft ignores the imaginary part of the DC and Nyquist bins
    const bool edge = c == 0 || c == p.NC;
    return (edge ? z.x * g.x : 2.0f * (z.x * g.x + z.y * g.y)) * p.inv_n;
  }
  float d = p.dA[(size_t)r * p.ldA + c];
  if (p.dB) d += p.dB[(size_t)r * p.ldB + c];
  if (ACT == BN_TANH) {
    const float a = tanhf(y);
    return d * (1.0f - a * a);
  }
  return y > 0.0f ? d : 0.0f;
}

template <int ACT>
__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(BnBwd p, int rows_per_cta) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= p.N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(p.rows, r0 + rows_per_cta);
  const float mean = p.stats[c], rstd = p.stats[2 * p.N + c];
  double s1 = 0.0, s2 = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const float xh = (p.X[(size_t)r * p.ld + c] - mean) * rstd;
    const float g = bn_bwd_g<ACT>(p, r, c, xh);
    s1 += (double)g;
    s2 = fma((double)g, (double)xh, s2);
  }
  atomicAdd(p.sums + 2 * c, s1);
  atomicAdd(p.sums + 2 * c + 1, s2);
}

// d x = rstd gamma [os] (g - S1 / M - xhat S2 / M)  as split planes; columns [N, ldg) are zeroed
template <int ACT>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(BnBwd p) {
  const long long total = p.rows * (long long)p.ldg;
  const double invM = 1.0 / (double)p.rows;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long r = i / p.ldg;
    const int c = (int)(i - r * p.ldg);
    float dx = 0.0f;
    if (c < p.N) {
      const float rstd = p.stats[2 * p.N + c];
      const float xh = (p.X[(size_t)r * p.ld + c] - p.stats[c]) * rstd;
      const float g = bn_bwd_g<ACT>(p, r, c, xh);
      const float m1 = (float)(p.sums[2 * c] * invM), m2 = (float)(p.sums[2 * c + 1] * invM);
      float k = rstd * p.gamma[c];
      if (ACT == BN_MASK) k *= p.oscale[c];
      dx = k * (g - m1 - xh * m2);
    }
    __nv_bfloat16 h, l;
    split_bf16(dx, h, l);
    p.ghi[i] = h;
    p.glo[i] = l;
  }
}

// parameter gradients from the column sums: d beta = S1, d gamma = S2; BN_MASK: the sums are those of g_u, so
// d omean = S1, d oscale = gamma S2 + beta S1, d beta = os S1, d gamma = os S2
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int N, int mask_mode, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ oscale, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta, float* __restrict__ doscale, float* __restrict__ domean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const double s1 = sums[2 * c], s2 = sums[2 * c + 1];
  if (mask_mode) {
    if (domean) domean[c] = (float)s1;
    if (doscale) doscale[c] = (float)((double)gamma[c] * s2 + (double)beta[c] * s1);
    if (dbeta) dbeta[c] = (float)((double)oscale[c] * s1);
    if (dgamma) dgamma[c] = (float)((double)oscale[c] * s2);
  } else {
    if (dbeta) dbeta[c] = (float)s1;
    if (dgamma) dgamma[c] = (float)s2;
  }
}

// x_in = (A + mean) scale (model.py:127-128):  d mean[k] = scale[k] sum_m dxin,  d scale[k] = sum_m dxin (A + mean) = sum_m dxin x_in / scale[k]
__global__ void __launch_bounds__(256) input_affine_sums_kernel(const float* __restrict__ dxin, int ldd, const __nv_bfloat16* __restrict__ xhi,
                                                                const __nv_bfloat16* __restrict__ xlo, int ldx, long long rows, int N,
                                                                double* __restrict__ sums, int rows_per_cta) {
  const int c = blockIdx.y * 256 + threadIdx.x;
  if (c >= N) return;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  double s1 = 0.0, s2 = 0.0;
  for (long long r = r0; r < r1; ++r) {
    const float d = dxin[(size_t)r * ldd + c];
    const size_t xi = (size_t)r * ldx + c;
    const float x = __bfloat162float(xhi[xi]) + __bfloat162float(xlo[xi]);
    s1 += (double)d;
    s2 = fma((double)d, (double)x, s2);
  }
  atomicAdd(sums + 2 * c, s1);
  atomicAdd(sums + 2 * c + 1, s2);
}
__global__ void input_affine_grads_kernel(const double* __restrict__ sums, int N, const float* __restrict__ scale, float* __restrict__ dmean,
                                          float* __restrict__ dscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const float s = scale[c];
  if (dmean) dmean[c] = (float)((double)s * sums[2 * c]);
  if (dscale) dscale[c] = s != 0.0f ? (float)(sums[2 * c + 1] / (double)s) : 0.0f;
}

// ------------------------------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------------------------------
struct TrainLayout {
  int F = 0, M = 0, lda1 = 0, ldm = 0, ldz = 0, ldg3 = 0, L = 0;
  size_t pA1 = 0, pXC = 0, pH = 0, pG3 = 0, pG8 = 0;  // plane sizes in elements
  // kept by the forward
  size_t Z = 0, A1 = 0, Y1 = 0, XC = 0, G[8] = {0}, Hp[8] = {0}, Dp[8] = {0}, Y2 = 0, Y2p = 0, Y3 = 0, mask = 0, stats = 0, acc = 0;
  // backward scratch
  size_t ghat = 0, gZ = 0, g3p = 0, dY2 = 0, g2p = 0, dXC = 0, dH = 0, R = 0, cs = 0, dG = 0, dGp = 0, dX = 0, g1p = 0, dxin = 0, carry = 0, bar = 0;
  size_t total = 0;
};

TrainLayout train_layout(const rfx_umx* h, int B, int T) {
  TrainLayout L;
  const int hid = h->cfg.hidden, H = h->H, bins = h->bins;
  L.L = h->cfg.nb_layers;
  L.F = T / h->cfg.hop + 1;
  L.M = B * L.F;
  L.lda1 = ceil_div(bins, 8) * 8;
  L.ldm = ceil_div(bins, 4) * 4;
  L.ldz = ceil_div(bins, 2) * 2;
  L.ldg3 = ceil_div(bins, 8) * 8;
  const size_t M = L.M;
  L.pA1 = M * L.lda1; L.pXC = M * 2 * hid; L.pH = M * hid; L.pG3 = M * L.ldg3; L.pG8 = M * 8 * H;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  L.Z = take(M * L.ldz * 8);
  L.A1 = take(L.pA1 * 4);
  L.Y1 = take(M * hid * 4);
  L.XC = take(L.pXC * 4);
  for (int l = 0; l < L.L; ++l) L.G[l] = take(M * 8 * H * 4);
  for (int l = 0; l + 1 < L.L; ++l) { L.Hp[l] = take(L.pH * 4); L.Dp[l] = take(L.pH * 4); }
  L.Y2 = take(M * hid * 4);
  L.Y2p = take(L.pH * 4);
  L.Y3 = take(M * L.ldm * 4);
  L.mask = take(M * L.ldm * 4);
  L.stats = take((size_t)3 * 3 * L.lda1 * 4);      // three layers x (mean, var, rstd), each block 3 * lda1 floats apart
  L.acc = take((size_t)2 * L.lda1 * 8);
  L.ghat = take((size_t)B * (T + h->cfg.n_fft) * 4);
  L.gZ = take(M * L.ldz * 8);
  L.g3p = take(L.pG3 * 4);
  L.dY2 = take(M * hid * 4);
  L.g2p = take(L.pH * 4);
  L.dXC = take(M * 2 * hid * 4);
  L.dH = take(M * hid * 4);
  L.R = take(M * 8 * H * 4);
  L.cs = take(M * 2 * H * 4);
  L.dG = take(M * 8 * H * 4);
  L.dGp = take(L.pG8 * 4);
  L.dX = take(M * hid * 4);
  L.g1p = take(L.pH * 4);
  L.dxin = take(M * L.ldm * 4);
  L.carry = take((size_t)B * 2 * H * 4);
  L.bar = take(4096);
  L.total = o;
  return L;
}

int rows_per_cta_for(long long rows) {
  long long r = (rows + 147) / 148;
  return (int)std::max<long long>(r, 32);
}

int bn_stats(const float* X, long long rows, int ld, int N, double* acc, float* stats, cudaStream_t s) {
  RFX_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)2 * N * 8, s));
  const int rpc = rows_per_cta_for(rows);
  col_stats_kernel<<<dim3((unsigned)((rows + rpc - 1) / rpc), ceil_div(N, 256)), 256, 0, s>>>(X, rows, ld, N, acc, rpc);
  bn_finalize_kernel<<<ceil_div(N, 256), 256, 0, s>>>(acc, rows, N, kBnEps, stats);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// packs of the backward GEMMs (lazily, once per finalize)
int train_prepare(rfx_umx* h, cudaStream_t s) {
  if (h->train_ready) return 0;
  const int hid = h->cfg.hidden, H = h->H, bins = h->bins, L = h->cfg.nb_layers;
  // the backing buffers keep their sizes from one finalize to the next: allocate once, re-pack in place
  const size_t n_store = (size_t)2 * L + L + 3;
  if (h->train_store.size() != n_store) {
    for (auto& b : h->train_store) b.release();
    h->train_store.assign(n_store, DevBuf());
  }
  size_t next = 0;
  h->whhp.assign(2 * L, SplitW());
  h->wih_t.assign(L, SplitW());
  int rc;
  for (int l = 0; l < L; ++l)
    for (int d = 0; d < 2; ++d) {
      const int BN = g2_choose_bn(4 * H);
      DevBuf& st = h->train_store[next++];
      if (st.alloc(split_weight_elems(4 * H, H, BN))) return 1;
      if ((rc = pack_split_weights(h->whh_cat[l].p + (size_t)d * 4 * H * H, H, 4 * H, H, BN, reinterpret_cast<__nv_bfloat16*>(st.p), &h->whhp[2 * l + d], s)))
        return rc;
    }
  DevBuf& tmp = h->train_tmp;
  const size_t tmp_n = std::max({(size_t)hid * (ceil_div(8 * H, 64) * 64), (size_t)hid * (ceil_div(bins, 64) * 64), (size_t)2 * hid * (ceil_div(hid, 64) * 64),
                                 (size_t)bins * (ceil_div(hid, 64) * 64)});
  if (tmp.alloc(tmp_n)) return 1;
  auto tpack = [&](const float* W, int N, int K, SplitW* out) -> int {  // stream-ordered: the shared scratch is reused pack after pack
    DevBuf& st = h->train_store[next++];
    if (st.alloc(bw::transposed_pack_floats(N, K))) return 1;
    return bw::pack_transposed(W, N, K, tmp.p, st.p, out, s);
  };
  for (int l = 0; l < L; ++l)
    if ((rc = tpack(h->wih_cat[l].p, 8 * H, hid, &h->wih_t[l]))) return rc;
  if ((rc = tpack(umx_param(h, "fc1.weight"), hid, bins, &h->fc1_t)) || (rc = tpack(umx_param(h, "fc2.weight"), hid, 2 * hid, &h->fc2_t)) ||
      (rc = tpack(umx_param(h, "fc3.weight"), bins, hid, &h->fc3_t)))
    return rc;
  h->train_ready = true;
  return 0;
}

int check_train_call(rfx_umx* h, const void* x, int B, int T, const void* ws, size_t ws_bytes) {
  RFX_REQUIRE(h && x && ws, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_umx_finalize has not been called since the last parameter load");
  RFX_REQUIRE(B > 0 && T > h->cfg.n_fft / 2, "need B > 0 and T > n_fft/2 (reflect padding)");
  RFX_REQUIRE(h->H == 256 && h->cfg.nb_layers <= 8, "training path: hidden 512");
  RFX_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  RFX_REQUIRE(ws_bytes >= train_layout(h, B, T).total, "workspace too small (rfx_umx_train_workspace_bytes)");
  return 0;
}

}  // namespace
}  // namespace rfx

using namespace rfx;

extern "C" {

size_t rfx_umx_train_workspace_bytes(const rfx_umx_t* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return train_layout(h, B, T).total;
}

int rfx_umx_train_prepare(rfx_umx_t* h, void* stream) {
  RFX_REQUIRE(h, "null handle");
  RFX_REQUIRE(h->finalized, "rfx_umx_finalize has not been called since the last parameter load");
  return train_prepare(h, (cudaStream_t)stream);
}

int rfx_umx_forward_train(rfx_umx_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes, const float* drop_masks,
                          int pow_pass, float alpha, float* bn_stats_out, void* stream) {
  int rc;
  if ((rc = check_train_call(h, x, B, T, workspace, workspace_bytes))) return rc;
  RFX_REQUIRE(pow_pass || out, "out is required unless pow_pass is set");
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = train_prepare(h, s))) return rc;
  const TrainLayout L = train_layout(h, B, T);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int hid = h->cfg.hidden, H = h->H, bins = h->bins, nl = L.L, M = L.M;
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto fl = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  float2* Z = reinterpret_cast<float2*>(ws + L.Z);
  double* acc = reinterpret_cast<double*>(ws + L.acc);
  float* stats[3] = {fl(L.stats), fl(L.stats) + 3 * L.lda1, fl(L.stats) + 6 * L.lda1};
  const float2* tw = twiddles(h->cfg.n_fft);
  RFX_REQUIRE(tw != nullptr, "twiddle table allocation failed");
  const unsigned ew_grid = 148 * 8;

  // (1) STFT + magnitude (+ power) + input affine -> split planes; the complex spectrum is kept for the mask gradient
  StftParams sp{};
  sp.x = x; sp.x_bstride = T; sp.T = T;
  sp.x_aligned8 = (((uintptr_t)x & 7) == 0 && (T % 2 == 0)) ? 1 : 0;
  sp.window = umx_param(h, "window"); sp.tw = tw;
  sp.n_fft = h->cfg.n_fft; sp.hop = h->cfg.hop; sp.F = L.F; sp.frame_off = h->cfg.n_fft / 2; sp.nbins = bins;
  sp.scale = 1.0f; sp.alpha = alpha; sp.mode = pow_pass ? STFT_UMX_POW : STFT_UMX_MAG;
  sp.Z = pow_pass ? nullptr : Z; sp.ldz = L.ldz;
  sp.Ahi = bf(L.A1); sp.Alo = bf(L.A1) + L.pA1; sp.ldas = L.lda1;
  sp.in_mean = umx_param(h, "input_mean"); sp.in_scale = umx_param(h, "input_scale");
  if ((rc = launch_stft(sp, B, s))) return rc;

  // (2) fc1 -> batch statistics -> normalise + tanh -> first half of the skip-concat planes
  Epilogue none;
  if ((rc = umx_dense(bf(L.A1), L.pA1, L.lda1, M, bins, h->fc1p, fl(L.Y1), hid, nullptr, 0, 0, none, s))) return rc;
  if ((rc = bn_stats(fl(L.Y1), M, hid, hid, acc, stats[0], s))) return rc;
  bn_apply_kernel<BN_TANH><<<ew_grid, 256, 0, s>>>(fl(L.Y1), M, hid, hid, stats[0], umx_param(h, "bn1.weight"), umx_param(h, "bn1.bias"), nullptr,
                                                   nullptr, bf(L.XC), bf(L.XC) + L.pXC, nullptr, 2 * hid);
  RFX_CHECK_CUDA(cudaGetLastError());

  // (3) the BiLSTM stack with dropout between the layers
  for (int l = 0; l < nl; ++l) {
    const __nv_bfloat16* lin; size_t lin_plane; int ldin;
    if (l == 0) { lin = bf(L.XC); lin_plane = L.pXC; ldin = 2 * hid; }
    else { lin = bf(drop_masks ? L.Dp[l - 1] : L.Hp[l - 1]); lin_plane = L.pH; ldin = hid; }
    Epilogue eb; eb.t1 = h->lstm_bias[l].p;
    if ((rc = umx_dense(lin, lin_plane, ldin, M, hid, h->wihp[l], fl(L.G[l]), 8 * H, nullptr, 0, 0, eb, s))) return rc;
    __nv_bfloat16* hout; size_t hplane; int ldh;
    if (l == nl - 1) { hout = bf(L.XC) + hid; hplane = L.pXC; ldh = 2 * hid; }
    else { hout = bf(L.Hp[l]); hplane = L.pH; ldh = hid; }
    if ((rc = launch_lstm_layer_impl(fl(L.G[l]), 8 * H, h->whh_cat[l].p, nullptr, 0, hout, hout + hplane, ldh, B, L.F, H, -1, 0, s))) return rc;
    if (l + 1 < nl && drop_masks) {
      dropout_planes_kernel<<<ew_grid, 256, 0, s>>>(hout, hout + hplane, ldh, drop_masks + (size_t)l * M * hid, M, hid, bf(L.Dp[l]),
                                                    bf(L.Dp[l]) + L.pH, hid);
      RFX_CHECK_CUDA(cudaGetLastError());
    }
  }

  // (4) fc2 -> BatchNorm -> ReLU ; (5) fc3 -> BatchNorm -> output scale / mean -> ReLU = the mask
  if ((rc = umx_dense(bf(L.XC), L.pXC, 2 * hid, M, 2 * hid, h->fc2p, fl(L.Y2), hid, nullptr, 0, 0, none, s))) return rc;
  if ((rc = bn_stats(fl(L.Y2), M, hid, hid, acc, stats[1], s))) return rc;
  bn_apply_kernel<BN_RELU><<<ew_grid, 256, 0, s>>>(fl(L.Y2), M, hid, hid, stats[1], umx_param(h, "bn2.weight"), umx_param(h, "bn2.bias"), nullptr,
                                                   nullptr, bf(L.Y2p), bf(L.Y2p) + L.pH, nullptr, hid);
  RFX_CHECK_CUDA(cudaGetLastError());
  if ((rc = umx_dense(bf(L.Y2p), L.pH, hid, M, hid, h->fc3p, fl(L.Y3), L.ldm, nullptr, 0, 0, none, s))) return rc;
  if ((rc = bn_stats(fl(L.Y3), M, L.ldm, bins, acc, stats[2], s))) return rc;
  if (bn_stats_out) {  // [mean1, var1, mean2, var2, mean3, var3]
    float* o = bn_stats_out;
    const int n[3] = {hid, hid, bins};
    for (int i = 0; i < 3; ++i) {
      RFX_CHECK_CUDA(cudaMemcpyAsync(o, stats[i], (size_t)2 * n[i] * 4, cudaMemcpyDeviceToDevice, s));
      o += 2 * n[i];
    }
  }
  if (pow_pass) return 0;  // the reference discards this pass's output (remfx/models.py:297): only the statistics matter
  bn_apply_kernel<BN_MASK><<<ew_grid, 256, 0, s>>>(fl(L.Y3), M, L.ldm, bins, stats[2], umx_param(h, "bn3.weight"), umx_param(h, "bn3.bias"),
                                                   umx_param(h, "output_scale"), umx_param(h, "output_mean"), nullptr, nullptr, fl(L.mask), L.ldm);
  RFX_CHECK_CUDA(cudaGetLastError());

  // (6) mask * mixture spectrum -> iSTFT
  IstftParams ip{};
  ip.Z = Z; ip.ldz = L.ldz; ip.mask = fl(L.mask); ip.ldm = L.ldm;
  ip.window = umx_param(h, "window"); ip.tw = tw;
  ip.n_fft = h->cfg.n_fft; ip.hop = h->cfg.hop; ip.F = L.F; ip.length = T;
  ip.frame_off = h->cfg.n_fft / 2; ip.env_pad = 0; ip.nbins = bins;
  ip.scale = 1.0f; ip.out = out; ip.out_bstride = T; ip.hops_per_cta = 16;
  if ((rc = launch_istft(ip, B, s))) return rc;
  h->tape_B = B; h->tape_T = T; h->tape_ws = workspace; h->tape_masks = drop_masks;
  return 0;
}

int rfx_umx_backward(rfx_umx_t* h, const float* x, const float* dout, int B, int T, const char* const* keys, float* const* grads, int nkeys,
                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_train_call(h, x, B, T, workspace, workspace_bytes))) return rc;
  RFX_REQUIRE(dout && keys && grads && nkeys > 0, "null argument");
  RFX_REQUIRE(h->tape_ws == workspace && h->tape_B == B && h->tape_T == T && h->train_ready,
              "rfx_umx_backward must follow rfx_umx_forward_train (pow_pass = 0) with the same shape and workspace, with no finalize in between");
  cudaStream_t s = (cudaStream_t)stream;
  const TrainLayout L = train_layout(h, B, T);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int hid = h->cfg.hidden, H = h->H, bins = h->bins, nl = L.L, M = L.M, n_fft = h->cfg.n_fft;
  auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  auto fl = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
  std::map<std::string, float*> G;
  for (int i = 0; i < nkeys; ++i) {
    RFX_REQUIRE(keys[i] && grads[i], "null key / gradient buffer");
    auto it = h->params.find(keys[i]);
    if (it == h->params.end()) { set_error(std::string("umx backward: unknown parameter '") + keys[i] + "'"); return 2; }
    G[keys[i]] = grads[i];
    RFX_CHECK_CUDA(cudaMemsetAsync(grads[i], 0, it->second.n * sizeof(float), s));
  }
  auto pg = [&](const std::string& k) -> float* { auto it = G.find(k); return it == G.end() ? nullptr : it->second; };
  float2* Z = reinterpret_cast<float2*>(ws + L.Z);
  float2* gZ = reinterpret_cast<float2*>(ws + L.gZ);
  double* acc = reinterpret_cast<double*>(ws + L.acc);
  float* stats[3] = {fl(L.stats), fl(L.stats) + 3 * L.lda1, fl(L.stats) + 6 * L.lda1};
  const float2* tw = twiddles(n_fft);
  const unsigned ew_grid = 148 * 8;
  const int rpc = rows_per_cta_for(M);
  const dim3 sum_grid((unsigned)((M + rpc - 1) / rpc), 1);
  Epilogue none;
  const int zero_off[1] = {0};

  // (6') iSTFT adjoint: dY_k = c_k / N FFT(window * dout / envelope)_k -- an ordinary STFT of the zero-padded, envelope-divided gradient
  const int P0 = n_fft / 2, Ltot = T + n_fft;
  if ((rc = bw::istft_adjoint_prep(dout, B, T, umx_param(h, "window"), n_fft, h->cfg.hop, n_fft / 2, L.F, 0, P0, Ltot, fl(L.ghat), s))) return rc;
  StftParams sp{};
  sp.x = fl(L.ghat); sp.x_bstride = Ltot; sp.T = Ltot; sp.x_aligned8 = 0;
  sp.window = umx_param(h, "window"); sp.tw = tw;
  sp.n_fft = n_fft; sp.hop = h->cfg.hop; sp.F = L.F; sp.frame_off = 0; sp.nbins = bins;
  sp.scale = 1.0f; sp.alpha = 1.0f; sp.mode = STFT_COMPLEX;
  sp.Z = gZ; sp.ldz = L.ldz;
  if ((rc = launch_stft(sp, B, s))) return rc;

  // (5') mask -> ReLU -> output affine -> BatchNorm 3 -> fc3
  auto run_bn_bwd = [&](int act, BnBwd p, float* dgamma, float* dbeta, float* doscale, float* domean) -> int {
    RFX_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)2 * p.N * 8, s));
    p.sums = acc;
    const dim3 g1(sum_grid.x, ceil_div(p.N, 256));
    if (act == BN_TANH) bn_bwd_sums_kernel<BN_TANH><<<g1, 256, 0, s>>>(p, rpc);
    else if (act == BN_RELU) bn_bwd_sums_kernel<BN_RELU><<<g1, 256, 0, s>>>(p, rpc);
    else bn_bwd_sums_kernel<BN_MASK><<<g1, 256, 0, s>>>(p, rpc);
    bn_param_grads_kernel<<<ceil_div(p.N, 256), 256, 0, s>>>(acc, p.N, act == BN_MASK ? 1 : 0, p.gamma, p.beta, p.oscale, dgamma, dbeta, doscale, domean);
    if (act == BN_TANH) bn_bwd_apply_kernel<BN_TANH><<<ew_grid, 256, 0, s>>>(p);
    else if (act == BN_RELU) bn_bwd_apply_kernel<BN_RELU><<<ew_grid, 256, 0, s>>>(p);
    else bn_bwd_apply_kernel<BN_MASK><<<ew_grid, 256, 0, s>>>(p);
    RFX_CHECK_CUDA(cudaGetLastError());
    return 0;
  };
  {
    BnBwd p{};
    p.X = fl(L.Y3); p.ld = L.ldm; p.rows = M; p.N = bins; p.stats = stats[2];
    p.gamma = umx_param(h, "bn3.weight"); p.beta = umx_param(h, "bn3.bias"); p.oscale = umx_param(h, "output_scale"); p.omean = umx_param(h, "output_mean");
    p.Z = Z; p.gZ = gZ; p.ldz = L.ldz; p.NC = n_fft / 2; p.inv_n = 1.0f / (float)n_fft;
    p.ghi = bf(L.g3p); p.glo = bf(L.g3p) + L.pG3; p.ldg = L.ldg3;
    if ((rc = run_bn_bwd(BN_MASK, p, pg("bn3.weight"), pg("bn3.bias"), pg("output_scale"), pg("output_mean")))) return rc;
  }
  SplitAct aY2; aY2.hi = bf(L.Y2p); aY2.rows = M; aY2.ld = hid; aY2.plane_stride = (long long)L.pH;
  if (float* dw = pg("fc3.weight"))
    if ((rc = bw::wgrad(bf(L.g3p), L.pG3, L.ldg3, 0, 1, 1, M, aY2, zero_off, zero_off, 1, bins, hid, hid, dw, s))) return rc;
  if ((rc = umx_dense(bf(L.g3p), L.pG3, L.ldg3, M, bins, h->fc3_t, fl(L.dY2), hid, nullptr, 0, 0, none, s))) return rc;

  // (4') ReLU -> BatchNorm 2 -> fc2
  {
    BnBwd p{};
    p.X = fl(L.Y2); p.ld = hid; p.rows = M; p.N = hid; p.stats = stats[1];
    p.gamma = umx_param(h, "bn2.weight"); p.beta = umx_param(h, "bn2.bias");
    p.dA = fl(L.dY2); p.ldA = hid;
    p.ghi = bf(L.g2p); p.glo = bf(L.g2p) + L.pH; p.ldg = hid;
    if ((rc = run_bn_bwd(BN_RELU, p, pg("bn2.weight"), pg("bn2.bias"), nullptr, nullptr))) return rc;
  }
  SplitAct aXC; aXC.hi = bf(L.XC); aXC.rows = M; aXC.ld = 2 * hid; aXC.plane_stride = (long long)L.pXC;
  if (float* dw = pg("fc2.weight"))
    if ((rc = bw::wgrad(bf(L.g2p), L.pH, hid, 0, 1, 1, M, aXC, zero_off, zero_off, 1, hid, 2 * hid, 2 * hid, dw, s))) return rc;
  if ((rc = umx_dense(bf(L.g2p), L.pH, hid, M, hid, h->fc2_t, fl(L.dXC), 2 * hid, nullptr, 0, 0, none, s))) return rc;

  // (3') the BiLSTM stack, last layer first.  dH of the last layer = the second half of d(concat)
  RFX_CHECK_CUDA(cudaMemcpy2DAsync(fl(L.dH), (size_t)hid * 4, fl(L.dXC) + hid, (size_t)2 * hid * 4, (size_t)hid * 4, M, cudaMemcpyDeviceToDevice, s));
  for (int l = nl - 1; l >= 0; --l) {
    const __nv_bfloat16* hp; size_t hplane; int ldh;
    if (l == nl - 1) { hp = bf(L.XC) + hid; hplane = L.pXC; ldh = 2 * hid; }
    else { hp = bf(L.Hp[l]); hplane = L.pH; ldh = hid; }
    if ((rc = bw::lstm_layer_backward(fl(L.G[l]), B, L.F, H, hp, hplane, ldh, h->whhp[2 * l], h->whhp[2 * l + 1], h->whh_cat[l].p, fl(L.dH), fl(L.R),
                                      fl(L.cs), fl(L.dG), fl(L.carry), reinterpret_cast<unsigned*>(ws + L.bar), s)))
      return rc;
    if ((rc = bw::split_pad(fl(L.dG), M, 8 * H, 8 * H, bf(L.dGp), bf(L.dGp) + L.pG8, s))) return rc;
    // input of the layer: tanh output (l = 0) or the (dropped) output of the layer below
    SplitAct ain;
    if (l == 0) { ain.hi = bf(L.XC); ain.ld = 2 * hid; ain.plane_stride = (long long)L.pXC; }
    else { ain.hi = bf(h->tape_masks ? L.Dp[l - 1] : L.Hp[l - 1]); ain.ld = hid; ain.plane_stride = (long long)L.pH; }
    ain.rows = M;
    SplitAct ah;  // this layer's h, per item (the recurrent weight gradient pairs dG_t with h_{t -+ 1}: rows outside an item read as zero)
    ah.rows = L.F; ah.rows_y = 1; ah.ld = ldh; ah.ld_y = (long long)L.F * ldh; ah.batch_stride = (long long)L.F * ldh; ah.plane_stride = (long long)hplane;
    for (int d = 0; d < 2; ++d) {
      const std::string sfx = "_l" + std::to_string(l) + (d ? "_reverse" : "");
      if (float* dw = pg("lstm.weight_ih" + sfx))
        if ((rc = bw::wgrad(bf(L.dGp), L.pG8, 8 * H, d * 4 * H, 1, 1, M, ain, zero_off, zero_off, 1, 4 * H, hid, hid, dw, s))) return rc;
      if (float* dw = pg("lstm.weight_hh" + sfx)) {
        ah.hi = hp + d * H;
        const int dxo[1] = {d ? 1 : -1};
        if ((rc = bw::wgrad(bf(L.dGp), L.pG8, 8 * H, d * 4 * H, B, 1, L.F, ah, dxo, zero_off, 1, 4 * H, H, H, dw, s))) return rc;
      }
      float* db_ih = pg("lstm.bias_ih" + sfx);
      float* db_hh = pg("lstm.bias_hh" + sfx);
      float* db = db_ih ? db_ih : db_hh;
      if (db) {
        if ((rc = bw::colsum(bf(L.dGp), L.pG8, M, 8 * H, d * 4 * H, 4 * H, db, s))) return rc;
        if (db_ih && db_hh) RFX_CHECK_CUDA(cudaMemcpyAsync(db_hh, db_ih, (size_t)4 * H * 4, cudaMemcpyDeviceToDevice, s));
      }
    }
    // input gradient of the layer
    if ((rc = umx_dense(bf(L.dGp), L.pG8, 8 * H, M, 8 * H, h->wih_t[l], fl(L.dX), hid, nullptr, 0, 0, none, s))) return rc;
    if (l > 0) {
      if (h->tape_masks) {
        mul_inplace_kernel<<<ew_grid, 256, 0, s>>>(fl(L.dX), h->tape_masks + (size_t)(l - 1) * M * hid, (long long)M * hid);
        RFX_CHECK_CUDA(cudaGetLastError());
      }
      RFX_CHECK_CUDA(cudaMemcpyAsync(fl(L.dH), fl(L.dX), (size_t)M * hid * 4, cudaMemcpyDeviceToDevice, s));
    }
  }

  // (2') tanh -> BatchNorm 1 -> fc1 ; d(tanh output) = first half of d(concat) + the first layer's input gradient
  {
    BnBwd p{};
    p.X = fl(L.Y1); p.ld = hid; p.rows = M; p.N = hid; p.stats = stats[0];
    p.gamma = umx_param(h, "bn1.weight"); p.beta = umx_param(h, "bn1.bias");
    p.dA = fl(L.dXC); p.ldA = 2 * hid; p.dB = fl(L.dX); p.ldB = hid;
    p.ghi = bf(L.g1p); p.glo = bf(L.g1p) + L.pH; p.ldg = hid;
    if ((rc = run_bn_bwd(BN_TANH, p, pg("bn1.weight"), pg("bn1.bias"), nullptr, nullptr))) return rc;
  }
  SplitAct aA1; aA1.hi = bf(L.A1); aA1.rows = M; aA1.ld = L.lda1; aA1.plane_stride = (long long)L.pA1;
  if (float* dw = pg("fc1.weight"))
    if ((rc = bw::wgrad(bf(L.g1p), L.pH, hid, 0, 1, 1, M, aA1, zero_off, zero_off, 1, hid, bins, bins, dw, s))) return rc;
  // (1') input affine: x_in = (|Z| + mean) scale
  if (pg("input_mean") || pg("input_scale")) {
    if ((rc = umx_dense(bf(L.g1p), L.pH, hid, M, hid, h->fc1_t, fl(L.dxin), L.ldm, nullptr, 0, 0, none, s))) return rc;
    RFX_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)2 * bins * 8, s));
    input_affine_sums_kernel<<<dim3(sum_grid.x, ceil_div(bins, 256)), 256, 0, s>>>(fl(L.dxin), L.ldm, bf(L.A1), bf(L.A1) + L.pA1, L.lda1, M, bins, acc, rpc);
    input_affine_grads_kernel<<<ceil_div(bins, 256), 256, 0, s>>>(acc, bins, umx_param(h, "input_scale"), pg("input_mean"), pg("input_scale"));
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
