// Cnn14 effect classifier (remfx/classifier.py:134-284, eval mode) on the GPU.
//
//   power STFT (stft.cu) -> mel filterbank GEMM (gemm2) -> per-item standardise (stats kernel, applied in the
//   first conv's load) -> ConvBlock x 6: 3x3 conv + BN + ReLU twice, then 2x2 average pool
//     * block-1 conv1 (1 -> 64 channels): SIMT kernel
//     * the other eleven 3x3 convolutions: gemm2 implicit GEMM, 9 taps on NHWC split-bf16 activations, zero
//       padding from TMA out-of-bounds fill, BatchNorm (eval) + ReLU fused in the epilogue
//   -> mean over time, max + mean over mel -> fc1 + ReLU -> 5 heads + sigmoid (fp32 FFMA: tiny M = batch)
// The mel image is (H = mel bin, W = frame); activations are [B][H][W][C] split planes.
#include "kernels.h"
#include "../../include/remfx_b200.h"

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace rfx {

// per-item mean and 1/std (unbiased, no eps: classifier.py:207) of the mel image; one block per item
__global__ void __launch_bounds__(1024) cnn_stats_kernel(const float* __restrict__ mel, long long n, float* __restrict__ stats) {
  __shared__ double r0[32], r1[32];
  const float* x = mel + (size_t)blockIdx.x * n;
  double s = 0.0, ss = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
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
    stats[2 * blockIdx.x + 1] = (float)(1.0 / sqrt(var));
  }
}

// block-1 conv1: in[b][h][w] = (mel[b][w][h] - mean) * rstd (mel is frame-major [B][W][H]), 3x3, pad 1, 1 -> C
// channels, BN + ReLU; output NHWC split planes.  One thread per (pixel, 8 channels).
__global__ void __launch_bounds__(256) cnn_first_kernel(const float* __restrict__ mel, const float* __restrict__ stats, int H, int W, int C,
                                                        const float* __restrict__ w /*[C][9]*/, const float* __restrict__ bs,
                                                        const float* __restrict__ bt, __nv_bfloat16* __restrict__ ohi,
                                                        __nv_bfloat16* __restrict__ olo) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)H * W * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int x = (int)((idx / groups) % W);
  const int y = (int)(idx / ((long long)groups * W));
  const float mean = stats[2 * b], rstd = stats[2 * b + 1];
  const float* m = mel + (size_t)b * H * W;
  float in[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int yy = y + dy - 1, xx = x + dx - 1;
      in[dy * 3 + dx] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? (m[(size_t)xx * H + yy] - mean) * rstd : 0.0f;
    }
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc = fmaf(w[c * 9 + t], in[t], acc);
    o[i] = fmaxf(fmaf(acc, bs[c], bt[c]), 0.0f);
  }
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(o[2 * i] - hf.x, o[2 * i + 1] - hf.y);
    ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const size_t off = (((size_t)b * H + y) * W + x) * C + c0;
  *reinterpret_cast<uint4*>(ohi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(olo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// 2x2 average pool on NHWC split planes: in [B][Hi][Wi][C] (only the first 2*Ho x 2*Wo pixels are read) -> [B][Ho][Wo][C]
__global__ void __launch_bounds__(256) cnn_pool_kernel(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo, int Hi, int Wi,
                                                       int C, int Ho, int Wo, __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)Ho * Wo * groups) return;
  const int c0 = (int)(idx % groups) * 8;
  const int x = (int)((idx / groups) % Wo);
  const int y = (int)(idx / ((long long)groups * Wo));
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const size_t off = (((size_t)b * Hi + 2 * y + dy) * Wi + 2 * x + dx) * C + c0;
      const uint4 h = *reinterpret_cast<const uint4*>(ihi + off);
      const uint4 l = *reinterpret_cast<const uint4*>(ilo + off);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[i]));
        const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[i]));
        acc[2 * i] += hf.x + lf.x;
        acc[2 * i + 1] += hf.y + lf.y;
      }
    }
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = acc[2 * i] * 0.25f, c = acc[2 * i + 1] * 0.25f;
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, c);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - hf.x, c - hf.y);
    ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const size_t off = (((size_t)b * Ho + y) * Wo + x) * C + c0;
  *reinterpret_cast<uint4*>(ohi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(olo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// mean over W, then max + mean over H (classifier.py:221-225): [B][H][W][C] split -> [B][C] fp32
__global__ void __launch_bounds__(256) cnn_head_pool_kernel(const __nv_bfloat16* __restrict__ ihi, const __nv_bfloat16* __restrict__ ilo, int H, int W,
                                                            int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= C) return;
  float mx = -INFINITY, sm = 0.0f;
  for (int y = 0; y < H; ++y) {
    float row = 0.0f;
    for (int x = 0; x < W; ++x) {
      const size_t off = (((size_t)b * H + y) * W + x) * C + c;
      row += __bfloat162float(ihi[off]) + __bfloat162float(ilo[off]);
    }
    row /= (float)W;
    mx = fmaxf(mx, row);
    sm += row;
  }
  out[(size_t)b * C + c] = mx + sm / (float)H;
}

// gather conv weight [Co][Ci][3][3] -> Wcat [Co][9 * Ci] (tap-major, tap = dy * 3 + dx)
__global__ void cnn_gather_w_kernel(const float* __restrict__ w, int Co, int Ci, float* __restrict__ wcat) {
  const long long total = (long long)Co * 9 * Ci;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int tap = (int)((i / Ci) % 9);
    const int co = (int)(i / ((long long)Ci * 9));
    wcat[i] = w[((size_t)co * Ci + ci) * 9 + tap];
  }
}

// stack the per-class heads: W [K][2048] <- heads.k.weight, b [K] <- heads.k.bias
struct CBuf {
  float* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    RFX_CHECK_CUDA(cudaMalloc(&p, count * sizeof(float)));
    n = count;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace rfx

using namespace rfx;

struct rfx_cnn14 {
  rfx_cnn14_config cfg;
  std::map<std::string, CBuf> params;
  CBuf bn_s[12], bn_t[12];   // folded BatchNorm of conv index 0..11
  CBuf wsplit[12];           // split planes (conv index 1..11) -- index 0 is the SIMT conv
  SplitW wpack[12];
  CBuf fbT_split;            // mel filterbank transposed [n_mels][bins] as split planes
  SplitW fbpack;
  CBuf heads_w, heads_b;
  bool finalized = false;
  ~rfx_cnn14() {
    for (auto& kv : params) kv.second.release();
    for (int i = 0; i < 12; ++i) { bn_s[i].release(); bn_t[i].release(); wsplit[i].release(); }
    fbT_split.release(); heads_w.release(); heads_b.release();
  }
};

namespace {
const int kChan[7] = {1, 64, 128, 256, 512, 1024, 2048};
const float* CP(const rfx_cnn14* h, const std::string& k) {
  auto it = h->params.find(k);
  return it == h->params.end() ? nullptr : it->second.p;
}
__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {  // dst[c][r] = src[r][c]
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < (long long)rows * cols) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(size_t)c * rows + r] = src[i];
  }
}

struct CnnLayout {
  int F, bins, H, W;  // frames, bins, image height (mel) and width (frames)
  size_t off_P, off_mel, off_stats, off_a, off_b, off_emb0, off_emb1, total;
  size_t plane_P, plane_act;
};
CnnLayout cnn_layout(const rfx_cnn14* h, int B, int T) {
  CnnLayout L;
  L.F = T / h->cfg.hop + 1;
  L.bins = h->cfg.n_fft / 2 + 1;
  L.H = h->cfg.n_mels;
  L.W = L.F;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  const int ldp = ceil_div(L.bins, 8) * 8;
  L.plane_P = (size_t)B * L.F * ldp;
  L.plane_act = (size_t)B * L.H * L.W * 64;  // largest activation: block-1 output (64 channels at full resolution)
  L.off_P = take(L.plane_P * 2 * 2);
  L.off_mel = take((size_t)B * L.F * L.H * 4);
  L.off_stats = take((size_t)B * 2 * 4);
  L.off_a = take(L.plane_act * 2 * 2);
  L.off_b = take(L.plane_act * 2 * 2);
  L.off_emb0 = take((size_t)B * 2048 * 4);
  L.off_emb1 = take((size_t)B * 2048 * 4);
  L.total = o;
  return L;
}
}  // namespace

extern "C" {

int rfx_cnn14_create(const rfx_cnn14_config* cfg, rfx_cnn14_t** out) {
  RFX_REQUIRE(cfg && out, "null argument");
  RFX_REQUIRE(cfg->n_fft == 512 || cfg->n_fft == 1024 || cfg->n_fft == 2048 || cfg->n_fft == 4096, "n_fft must be 512/1024/2048/4096");
  RFX_REQUIRE(cfg->hop > 0 && cfg->hop % 2 == 0, "hop must be even");
  RFX_REQUIRE(cfg->n_mels >= 32 && cfg->n_mels % 32 == 0 && cfg->n_mels <= 256, "n_mels must be a multiple of 32 in [32, 256]");
  RFX_REQUIRE(cfg->num_classes >= 1 && cfg->num_classes <= 64, "num_classes in [1, 64]");
  rfx_cnn14* h = new rfx_cnn14();
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void rfx_cnn14_destroy(rfx_cnn14_t* h) { delete h; }

int rfx_cnn14_load_param(rfx_cnn14_t* h, const char* key, const float* src, int64_t numel, void* stream) {
  RFX_REQUIRE(h && key && src && numel > 0, "bad argument");
  CBuf& b = h->params[key];
  if (b.n != (size_t)numel) {
    if (b.alloc((size_t)numel)) return 1;
  }
  RFX_CHECK_CUDA(cudaMemcpyAsync(b.p, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->finalized = false;
  return 0;
}

int rfx_cnn14_finalize(rfx_cnn14_t* h, void* stream) {
  RFX_REQUIRE(h, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  const int bins = h->cfg.n_fft / 2 + 1, nm = h->cfg.n_mels, nc = h->cfg.num_classes;
  auto need = [&](const std::string& k, size_t n) -> int {
    auto it = h->params.find(k);
    if (it == h->params.end()) { set_error("cnn14: missing parameter '" + k + "'"); return 2; }
    if (it->second.n != n) { set_error("cnn14: parameter '" + k + "' has " + std::to_string(it->second.n) + " elements, expected " + std::to_string(n)); return 2; }
    return 0;
  };
  int rc;
  if ((rc = need("melspec.spectrogram.window", h->cfg.n_fft)) || (rc = need("melspec.mel_scale.fb", (size_t)bins * nm))) return rc;
  CBuf tmp;
  size_t tmp_n = (size_t)nm * bins;
  for (int i = 1; i < 12; ++i) tmp_n = std::max(tmp_n, (size_t)kChan[(i + 2) / 2] * 9 * kChan[(i + 1) / 2]);
  if (tmp.alloc(tmp_n)) return 1;
  // mel filterbank as the W operand of a GEMM: W[n_mels][bins]
  transpose_kernel<<<ceil_div(bins * nm, 256), 256, 0, s>>>(CP(h, "melspec.mel_scale.fb"), bins, nm, tmp.p);
  RFX_CHECK_CUDA(cudaGetLastError());
  if (h->fbT_split.alloc(split_weight_elems(nm, bins, 128))) { tmp.release(); return 1; }
  if ((rc = pack_split_weights(tmp.p, bins, nm, bins, 128, reinterpret_cast<__nv_bfloat16*>(h->fbT_split.p), &h->fbpack, s))) { tmp.release(); return rc; }
  for (int i = 0; i < 12; ++i) {  // conv index i: block i/2 + 1, conv i%2 + 1
    const int blk = i / 2 + 1, cv = i % 2 + 1;
    const int cin = (cv == 1) ? kChan[blk - 1] : kChan[blk], cout = kChan[blk];
    const std::string p = "conv_block" + std::to_string(blk);
    const std::string wk = p + ".conv" + std::to_string(cv) + ".weight", bk = p + ".bn" + std::to_string(cv);
    if ((rc = need(wk, (size_t)cout * cin * 9))) { tmp.release(); return rc; }
    for (const char* f : {".weight", ".bias", ".running_mean", ".running_var"})
      if ((rc = need(bk + f, cout))) { tmp.release(); return rc; }
    if (h->bn_s[i].alloc(cout) || h->bn_t[i].alloc(cout)) { tmp.release(); return 1; }
    if ((rc = launch_bn_fold(CP(h, bk + ".weight"), CP(h, bk + ".bias"), CP(h, bk + ".running_mean"), CP(h, bk + ".running_var"), 1e-5f,
                             h->bn_s[i].p, h->bn_t[i].p, cout, s))) { tmp.release(); return rc; }
    if (i == 0) continue;
    cnn_gather_w_kernel<<<148 * 8, 256, 0, s>>>(CP(h, wk), cout, cin, tmp.p);
    RFX_CHECK_CUDA(cudaGetLastError());
    const int BN = g2_choose_bn(cout);
    if (h->wsplit[i].alloc(split_weight_elems(cout, 9 * cin, BN))) { tmp.release(); return 1; }
    if ((rc = pack_split_weights(tmp.p, 9ll * cin, cout, 9 * cin, BN, reinterpret_cast<__nv_bfloat16*>(h->wsplit[i].p), &h->wpack[i], s))) {
      tmp.release();
      return rc;
    }
  }
  if ((rc = need("fc1.weight", (size_t)2048 * 2048)) || (rc = need("fc1.bias", 2048))) { tmp.release(); return rc; }
  if (h->heads_w.alloc((size_t)nc * 2048) || h->heads_b.alloc(nc)) { tmp.release(); return 1; }
  for (int k = 0; k < nc; ++k) {
    const std::string p = "heads." + std::to_string(k);
    if ((rc = need(p + ".weight", 2048)) || (rc = need(p + ".bias", 1))) { tmp.release(); return rc; }
    RFX_CHECK_CUDA(cudaMemcpyAsync(h->heads_w.p + (size_t)k * 2048, CP(h, p + ".weight"), 2048 * 4, cudaMemcpyDeviceToDevice, s));
    RFX_CHECK_CUDA(cudaMemcpyAsync(h->heads_b.p + k, CP(h, p + ".bias"), 4, cudaMemcpyDeviceToDevice, s));
  }
  RFX_CHECK_CUDA(cudaStreamSynchronize(s));
  tmp.release();
  h->finalized = true;
  return 0;
}

size_t rfx_cnn14_workspace_bytes(const rfx_cnn14_t* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return cnn_layout(h, B, T).total;
}

int rfx_cnn14_forward(rfx_cnn14_t* h, const float* x, int B, int T, float* probs, float* logits, void* workspace, size_t workspace_bytes,
                      void* stream) {
  RFX_REQUIRE(h && x && probs && workspace, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_cnn14_finalize has not been called since the last parameter load");
  RFX_REQUIRE(B > 0 && T > h->cfg.n_fft / 2, "need B > 0 and T > n_fft/2");
  const CnnLayout L = cnn_layout(h, B, T);
  RFX_REQUIRE(L.W >= 64 && L.H >= 32, "input too short for five 2x2 poolings");
  RFX_REQUIRE(workspace_bytes >= L.total, "workspace too small (rfx_cnn14_workspace_bytes)");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(ws + L.off_P);
  float* mel = reinterpret_cast<float*>(ws + L.off_mel);
  float* stats = reinterpret_cast<float*>(ws + L.off_stats);
  __nv_bfloat16* act[2] = {reinterpret_cast<__nv_bfloat16*>(ws + L.off_a), reinterpret_cast<__nv_bfloat16*>(ws + L.off_b)};
  float* emb0 = reinterpret_cast<float*>(ws + L.off_emb0);
  float* emb1 = reinterpret_cast<float*>(ws + L.off_emb1);
  const int nm = h->cfg.n_mels, bins = L.bins, ldp = ceil_div(bins, 8) * 8;
  int rc;

  // C1: power STFT (classifier.py:156-161,200) as split planes, then mel = P . fb  (frame-major [B*F][n_mels])
  StftParams sp{};
  sp.x = x; sp.x_bstride = T; sp.T = T;
  sp.x_aligned8 = (((uintptr_t)x & 7) == 0 && (T % 2 == 0)) ? 1 : 0;
  sp.window = CP(h, "melspec.spectrogram.window"); sp.tw = twiddles(h->cfg.n_fft);
  sp.n_fft = h->cfg.n_fft; sp.hop = h->cfg.hop; sp.F = L.F;
  sp.frame_off = h->cfg.n_fft / 2; sp.nbins = bins;
  sp.scale = 1.0f; sp.alpha = 1.0f; sp.mode = STFT_POWER;
  sp.Z = nullptr; sp.A = nullptr; sp.Ahi = P; sp.Alo = P + L.plane_P; sp.ldas = ldp;
  if ((rc = launch_stft(sp, B, s))) return rc;
  {
    G2Problem pr;
    pr.A.hi = P; pr.A.rows = (long long)B * L.F; pr.A.ld = ldp; pr.A.plane_stride = (long long)L.plane_P;
    pr.W = h->fbpack;
    pr.M = B * L.F; pr.N = nm; pr.batch = 1; pr.Ktap = bins; pr.taps = 1;
    pr.Cf = mel; pr.ldcf = nm;
    if ((rc = launch_gemm2(pr, s))) return rc;
  }
  // C2: per-item standardisation statistics (classifier.py:207)
  cnn_stats_kernel<<<B, 1024, 0, s>>>(mel, (long long)L.F * nm, stats);
  RFX_CHECK_CUDA(cudaGetLastError());

  // C3: conv blocks
  int H = L.H, W = L.W;  // current activation extent; Wa = allocated width of the current buffer
  int cur = 0;
  {
    const long long items = (long long)H * W * (64 / 8);
    dim3 grid((unsigned)((items + 255) / 256), B);
    cnn_first_kernel<<<grid, 256, 0, s>>>(mel, stats, H, W, 64, CP(h, "conv_block1.conv1.weight"), h->bn_s[0].p, h->bn_t[0].p, act[0],
                                          act[0] + L.plane_act);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  for (int i = 1; i < 12; ++i) {
    const int blk = i / 2 + 1, cv = i % 2 + 1;
    const int cin = (cv == 1) ? kChan[blk - 1] : kChan[blk], cout = kChan[blk];
    G2Problem pr;
    pr.A.hi = act[cur]; pr.A.rows = W; pr.A.rows_y = H; pr.A.ld = cin; pr.A.ld_y = (long long)W * cin;
    pr.A.batch_stride = (long long)H * W * cin; pr.A.plane_stride = (long long)L.plane_act;
    pr.W = h->wpack[i];
    pr.M = W; pr.My = H; pr.N = cout; pr.batch = B; pr.Ktap = cin; pr.taps = 9;
    pr.xt = W >= 128 ? 128 : (W >= 64 ? 64 : (W >= 32 ? 32 : 16));
    for (int t = 0; t < 9; ++t) { pr.row_off[t] = t % 3 - 1; pr.row_off_y[t] = t / 3 - 1; }
    pr.Chi = act[cur ^ 1]; pr.Clo = act[cur ^ 1] + L.plane_act;
    pr.ldcs = cout; pr.ldcs_y = (long long)W * cout; pr.bscs = (long long)H * W * cout;
    pr.epi.s1 = h->bn_s[i].p; pr.epi.t1 = h->bn_t[i].p; pr.epi.act = ACT_RELU;
    if ((rc = launch_gemm2(pr, s))) return rc;
    cur ^= 1;
    if (cv == 2 && blk <= 5) {  // avg_pool2d(2, 2) (floor): classifier.py:209-217; block 6 pools 1x1
      const int Ho = H / 2, Wo = W / 2;
      const long long items = (long long)Ho * Wo * (cout / 8);
      dim3 grid((unsigned)((items + 255) / 256), B);
      cnn_pool_kernel<<<grid, 256, 0, s>>>(act[cur], act[cur] + L.plane_act, H, W, cout, Ho, Wo, act[cur ^ 1], act[cur ^ 1] + L.plane_act);
      RFX_CHECK_CUDA(cudaGetLastError());
      cur ^= 1;
      H = Ho; W = Wo;
    }
  }
  // C4: pooled head (classifier.py:221-231)
  cnn_head_pool_kernel<<<dim3(ceil_div(2048, 256), B), 256, 0, s>>>(act[cur], act[cur] + L.plane_act, H, W, 2048, emb0);
  RFX_CHECK_CUDA(cudaGetLastError());
  Epilogue e1; e1.t1 = CP(h, "fc1.bias"); e1.act = ACT_RELU;
  if ((rc = launch_gemm_simt(emb0, 2048, B, CP(h, "fc1.weight"), 2048, 2048, 2048, emb1, 2048, e1, s))) return rc;
  const int nc = h->cfg.num_classes;
  if (logits) {
    Epilogue el; el.t1 = h->heads_b.p;
    if ((rc = launch_gemm_simt(emb1, 2048, B, h->heads_w.p, 2048, nc, 2048, logits, nc, el, s))) return rc;
  }
  Epilogue e2; e2.t1 = h->heads_b.p; e2.act = ACT_SIGMOID;
  return launch_gemm_simt(emb1, 2048, B, h->heads_w.p, 2048, nc, 2048, probs, nc, e2, s);
}

}  // extern "C"
