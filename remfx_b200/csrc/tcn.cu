// TCN effect-removal model (remfx/tcn.py:62-138 behind remfx/models.py:370-390) on the GPU.
//
//   block 0  (1 -> C channels): SIMT kernel, 7 taps + bias -> PReLU -> + 1x1 residual on the centre sample
//   blocks 1..N-1 (C -> C):     gemm2 implicit GEMM in dual-accumulator mode (tcgen05, bf16x3):
//                               D1 = sum_j W_j x[t + j d]  (7 taps), D2 = W_res x[t + 3 d];
//                               out = PReLU_c(D1 + b) + D2        (remfx/tcn.py:48-59, centre crop = offset 3d)
//   tail:    tanh(1x1 conv C -> 1)  (remfx/tcn.py:129)
// Activations are channel-last ([B][L][C]) split-bf16 planes, so every tap of the dilated convolution is a
// plain 2-D TMA box at a shifted row -- no im2col, no padding copies.  Two ping-pong buffers.
#include "tcn_internal.h"

namespace rfx {

// ---- block 0: x (B, L0) fp32 -> split planes [B][L1][C], one thread per (time step, 8 channels) ----
__global__ void __launch_bounds__(256) tcn_first_kernel(const float* __restrict__ x, long long x_bs, int L1, int C, int K, int dil, int res_off,
                                                        const float* __restrict__ w /*[C][K]*/, const float* __restrict__ bias,
                                                        const float* __restrict__ wres /*[C]*/, const float* __restrict__ slope,
                                                        __nv_bfloat16* __restrict__ ohi, __nv_bfloat16* __restrict__ olo, long long o_bs) {
  const int groups = C / 8;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= (long long)L1 * groups) return;
  const int t = (int)(idx / groups);
  const int c0 = (int)(idx % groups) * 8;
  const float* xr = x + (size_t)b * x_bs + t;
  float xs[16];
  for (int j = 0; j < K; ++j) xs[j] = xr[j * dil];
  const float xc = xr[res_off];
  uint32_t ph[4], pl[4];
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    float acc = bias[c];
    for (int j = 0; j < K; ++j) acc = fmaf(w[c * K + j], xs[j], acc);
    acc = acc >= 0.0f ? acc : acc * slope[c];
    o[i] = acc + wres[c] * xc;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(o[2 * i] - hf.x, o[2 * i + 1] - hf.y);
    ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  const size_t off = (size_t)b * o_bs + (size_t)t * C + c0;
  *reinterpret_cast<uint4*>(ohi + off) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  *reinterpret_cast<uint4*>(olo + off) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

// ---- tail: out[b][t] = tanh(sum_c w[c] * x[b][t][c] + bias); one warp per time step ----
__global__ void __launch_bounds__(256) tcn_tail_kernel(const __nv_bfloat16* __restrict__ xhi, const __nv_bfloat16* __restrict__ xlo, long long x_bs,
                                                       int L, int C, const float* __restrict__ w, const float* __restrict__ bias,
                                                       float* __restrict__ out, long long o_bs) {
  const int lane = threadIdx.x & 31;
  const long long t = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (t >= L) return;
  const __nv_bfloat16* rh = xhi + (size_t)b * x_bs + (size_t)t * C;
  const __nv_bfloat16* rl = xlo + (size_t)b * x_bs + (size_t)t * C;
  float acc = 0.0f;
  for (int c = lane * 8; c < C; c += 256) {
    const uint4 h = *reinterpret_cast<const uint4*>(rh + c);
    const uint4 l = *reinterpret_cast<const uint4*>(rl + c);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[i]));
      const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[i]));
      acc = fmaf(w[c + 2 * i], hf.x + lf.x, acc);
      acc = fmaf(w[c + 2 * i + 1], hf.y + lf.y, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)b * o_bs + t] = tanhf(acc + bias[0]);
}

// ---- gather conv1.weight [C][C][K] and res.weight [C][C][1] into Wcat [C][(K+1)*C] (tap-major) ----
__global__ void tcn_gather_w_kernel(const float* __restrict__ wconv, const float* __restrict__ wres, int C, int K, float* __restrict__ wcat) {
  const long long total = (long long)C * (K + 1) * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C);
    const int tap = (int)((i / C) % (K + 1));
    const int co = (int)(i / ((long long)C * (K + 1)));
    wcat[i] = tap < K ? wconv[((size_t)co * C + ci) * K + tap] : wres[(size_t)co * C + ci];
  }
}

}  // namespace rfx

using namespace rfx;

namespace rfx {

int tcn_run_forward(rfx_tcn* h, const float* x, int B, long long T, float* out, __nv_bfloat16* const* block_out, long long plane_elems,
                    cudaStream_t s) {
  const int C = h->cfg.channel_width, K = h->cfg.kernel_size, NBk = h->cfg.nblocks;
  const long long Lout = tcn_len_after(h, T, NBk);
  const long long L1 = tcn_len_after(h, T, 1);
  const long long bs = L1 * C;  // batch stride (elements) of every activation buffer
  // block 0
  {
    const int d = tcn_dilation_of(h, 0);
    const long long items = L1 * (C / 8);
    dim3 grid((unsigned)((items + 255) / 256), B);
    tcn_first_kernel<<<grid, 256, 0, s>>>(x, T, (int)L1, C, K, d, tcn_res_off(h, d), tcn_param(h, "process_blocks.0.conv1.weight"),
                                          tcn_param(h, "process_blocks.0.conv1.bias"), tcn_param(h, "process_blocks.0.res.weight"),
                                          tcn_param(h, "process_blocks.0.relu.weight"), block_out[0], block_out[0] + plane_elems, bs);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  long long Lin = L1;
  for (int n = 1; n < NBk; ++n) {
    const int d = tcn_dilation_of(h, n);
    const long long Lo = Lin - (long long)(K - 1) * d;
    const std::string p = "process_blocks." + std::to_string(n);
    G2Problem pr;
    pr.A.hi = block_out[n - 1]; pr.A.rows = Lin; pr.A.ld = C; pr.A.batch_stride = bs; pr.A.plane_stride = plane_elems;
    pr.W = h->wpack[n];
    pr.M = (int)Lo; pr.N = C; pr.batch = B; pr.Ktap = C; pr.taps = K + 1;
    for (int j = 0; j < K; ++j) pr.row_off[j] = j * d;
    pr.row_off[K] = tcn_res_off(h, d);  // causal_crop drops the last sample: offset (K-1) d - 1; center_crop: (K-1) d / 2
    pr.dual = true;
    pr.Chi = block_out[n]; pr.Clo = block_out[n] + plane_elems; pr.ldcs = C; pr.bscs = bs;
    pr.epi.t1 = tcn_param(h, p + ".conv1.bias");
    pr.epi.slope = tcn_param(h, p + ".relu.weight");
    pr.epi.act = ACT_PRELU;
    int rc = launch_gemm2(pr, s);
    if (rc) return rc;
    Lin = Lo;
  }
  {
    dim3 grid((unsigned)((Lin + 7) / 8), B);
    tcn_tail_kernel<<<grid, 256, 0, s>>>(block_out[NBk - 1], block_out[NBk - 1] + plane_elems, bs, (int)Lin, C, tcn_param(h, "output.weight"),
                                         tcn_param(h, "output.bias"), out, Lout);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // namespace rfx

extern "C" {

int rfx_tcn_create(const rfx_tcn_config* cfg, rfx_tcn_t** out) {
  RFX_REQUIRE(cfg && out, "null argument");
  RFX_REQUIRE(cfg->ninputs == 1 && cfg->noutputs == 1, "TCN: ninputs and noutputs must be 1 (mono audio, cfg/model/tcn.yaml)");
  RFX_REQUIRE(cfg->nblocks >= 1 && cfg->nblocks <= 64, "TCN: nblocks in [1, 64]");
  RFX_REQUIRE(cfg->channel_width % 64 == 0 && cfg->channel_width >= 64 && cfg->channel_width <= 256,
              "TCN: channel_width must be 64, 128, 192 or 256");
  RFX_REQUIRE(cfg->kernel_size >= 2 && cfg->kernel_size <= 15 && cfg->kernel_size % 2 == 1, "TCN: odd kernel_size in [3, 15]");
  RFX_REQUIRE(cfg->stack_size >= 1 && cfg->dilation_growth >= 1, "TCN: stack_size, dilation_growth >= 1");
  rfx_tcn* h = new rfx_tcn();
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void rfx_tcn_destroy(rfx_tcn_t* h) { delete h; }

int rfx_tcn_load_param(rfx_tcn_t* h, const char* key, const float* src, int64_t numel, void* stream) {
  RFX_REQUIRE(h && key && src && numel > 0, "bad argument");
  TcnBuf& b = h->params[key];
  if (b.n != (size_t)numel) {
    if (b.alloc((size_t)numel)) return 1;
  }
  RFX_CHECK_CUDA(cudaMemcpyAsync(b.p, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->finalized = false;
  h->transposed_ready = false;
  return 0;
}

int rfx_tcn_finalize(rfx_tcn_t* h, void* stream) {
  RFX_REQUIRE(h, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  const int C = h->cfg.channel_width, K = h->cfg.kernel_size, NBk = h->cfg.nblocks;
  auto need = [&](const std::string& k, size_t n) -> int {
    auto it = h->params.find(k);
    if (it == h->params.end()) { set_error("tcn: missing parameter '" + k + "'"); return 2; }
    if (it->second.n != n) { set_error("tcn: parameter '" + k + "' has " + std::to_string(it->second.n) + " elements, expected " + std::to_string(n)); return 2; }
    return 0;
  };
  int rc;
  if ((int)h->wsplit.size() != NBk) {
    for (auto& b : h->wsplit) b.release();
    h->wsplit.assign(NBk, TcnBuf());
  }
  h->wpack.assign(NBk, SplitW());
  h->transposed_ready = false;
  // all work below is ordered on `s`; the staging buffer and the packed planes are kept across calls (same sizes), so a
  // training loop that re-finalizes after every optimiser step neither allocates nor synchronises here
  if (NBk > 1 && h->wcat.alloc((size_t)C * (K + 1) * C)) return 1;
  for (int n = 0; n < NBk; ++n) {
    const std::string p = "process_blocks." + std::to_string(n);
    const int cin = n == 0 ? 1 : C;
    if ((rc = need(p + ".conv1.weight", (size_t)C * cin * K)) || (rc = need(p + ".conv1.bias", C)) || (rc = need(p + ".res.weight", (size_t)C * cin)) ||
        (rc = need(p + ".relu.weight", C)))
      return rc;
    if (n == 0) continue;
    tcn_gather_w_kernel<<<148 * 4, 256, 0, s>>>(tcn_param(h, p + ".conv1.weight"), tcn_param(h, p + ".res.weight"), C, K, h->wcat.p);
    RFX_CHECK_CUDA(cudaGetLastError());
    if (h->wsplit[n].alloc(split_weight_elems(C, (K + 1) * C, 256))) return 1;
    if ((rc = pack_split_weights(h->wcat.p, (long long)(K + 1) * C, C, (K + 1) * C, 256, reinterpret_cast<__nv_bfloat16*>(h->wsplit[n].p), &h->wpack[n], s)))
      return rc;
  }
  if ((rc = need("output.weight", C)) || (rc = need("output.bias", 1))) return rc;
  h->finalized = true;
  return 0;
}

long long rfx_tcn_out_length(const rfx_tcn_t* h, long long T) { return h ? tcn_len_after(h, T, h->cfg.nblocks) : 0; }

size_t rfx_tcn_workspace_bytes(const rfx_tcn_t* h, int B, long long T) {
  if (!h || B <= 0) return 0;
  if (tcn_len_after(h, T, 1) <= 0) return 0;
  return 4 * tcn_plane_bytes(h, B, T);  // two ping-pong buffers x (hi, lo)
}

int rfx_tcn_forward(rfx_tcn_t* h, const float* x, int B, long long T, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(h && x && out && workspace, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_tcn_finalize has not been called since the last parameter load");
  RFX_REQUIRE(B > 0 && tcn_len_after(h, T, h->cfg.nblocks) > 0, "input shorter than the receptive field");
  RFX_REQUIRE(T < (1ll << 31), "T too large");
  RFX_REQUIRE(workspace_bytes >= rfx_tcn_workspace_bytes(h, B, T), "workspace too small (rfx_tcn_workspace_bytes)");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const size_t plane_bytes = tcn_plane_bytes(h, B, T);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  std::vector<__nv_bfloat16*> outs(h->cfg.nblocks);
  for (int n = 0; n < h->cfg.nblocks; ++n) outs[n] = reinterpret_cast<__nv_bfloat16*>(ws + (size_t)(n & 1) * 2 * plane_bytes);  // ping-pong
  return tcn_run_forward(h, x, B, T, out, outs.data(), (long long)(plane_bytes / 2), (cudaStream_t)stream);
}

int rfx_tcn_launches_per_call(const rfx_tcn_t* h) { return h ? h->cfg.nblocks + 1 : 0; }

}  // extern "C"
