// Hybrid-Demucs handle, tensor descriptors and the training tape shared by hdemucs.cu (forward) and hdemucs_bwd.cu (backward).
#pragma once
#include "kernels.h"
#include "../../include/remfx_b200.h"

#include <map>
#include <string>
#include <vector>

namespace rfx {
namespace hd {

struct Buf {
  float* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (p && n == count) return 0;  // same size: keep the buffer (a training loop re-finalizes after every optimiser step)
    release();
    RFX_CHECK_CUDA(cudaMalloc(&p, (count > 0 ? count : 1) * sizeof(float)));
    n = count;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

// channel-last activation (B, Y, X, C): split planes (hi, lo = hi + plane) or fp32
struct Ten {
  int B = 0, Y = 1, X = 0, C = 0;
  __nv_bfloat16* hi = nullptr;
  size_t plane = 0;
  float* f = nullptr;
  __nv_bfloat16* lo() const { return hi + plane; }
  size_t elems() const { return (size_t)B * Y * X * C; }
  const void* key() const { return hi ? (const void*)hi : (const void*)f; }
};

// ---- weight gather for the implicit-GEMM convolutions ----
struct GatherSpec {
  int kind;      // 0 plain (taps = k, or kh*kw), 1 strided (regrouped by s), 2 transposed (regrouped by s)
  int Co, Ci, k; // logical conv dims; k = kernel extent along the conv axis (kh * kw for 2-D plain)
  int s, p;      // stride / padding (kinds 1, 2)
  int tau_min;   // first group offset (kind 1)
  int taps;      // number of GEMM taps
  int Kp;        // padded K per tap (multiple of 64)
  int glu;       // interleave output rows (value c, gate c) -> (2c, 2c+1)
  int Nout;      // GEMM N
};

// ---- GroupNorm apply + activation (see hd_kernels.cuh: gn_apply_kernel) ----
// mode 0: y = gn(raw); 1: gelu(gn(raw)); 2: GLU over channel halves; 3: GLU over interleaved (value, gate) column pairs
struct GnApply {
  const float* raw; int Y, Xr, Cr;
  const float* stats; int G, per_x;
  const float* gamma; const float* beta;
  int mode;
  const float* scale;
  const __nv_bfloat16* rhi; const __nv_bfloat16* rlo;
  __nv_bfloat16* ohi; __nv_bfloat16* olo; int Xo, Co, x_off;
};

// one convolution prepared for gemm2
struct Conv {
  GatherSpec g{};
  SplitW w;
  Buf wbuf;   // split planes
  Buf bias;   // [Nout] (re-ordered like the GEMM columns)
  int Ci = 0, Co = 0;
  int kh = 1, kw = 1;  // 2-D plain convs (kh along X = freq, kw along Y = time)
  int crop = 0;        // transposed convs: samples cropped on each side of the output (TA:288-294)
  std::string wkey, bkey, bkey2;  // state_dict keys the weight / bias came from (bkey2: LSTM bias_hh beside bias_ih)
  // backward: transposed pack  Wt[k][tap * ceil64(Nout) + n] = Wcat[n][tap][k]  (built lazily, re-built after a parameter load)
  SplitW wt;
  Buf wtbuf;
  bool wt_ready = false;
  int pack_ev = -1;   // backward: index of the event (rfx_hdemucs::ev_pack) that marks this pack complete on the preparation stream; -1 = none pending
};

// ---- training tape: one record per forward op, replayed in reverse by hdemucs_bwd.cu ----
enum OpKind {
  OP_CONV = 0,     // in (split) -> raw fp32 (possibly a column range of a wider tensor); pr = the forward problem
  OP_GN,           // raw fp32 [+ stats] -> out split (activation, LayerScale, residual, crop)
  OP_ADDCROP,      // out = a[x + x_off] + skip
  OP_FREQEMB,      // out += w * emb[x][c] (in place)
  OP_ADDF32,       // sum = raw + inject (fp32)
  OP_FRAME,        // BLSTM framing
  OP_LSTM,         // one bidirectional LSTM layer: in (split), gates G (fp32), hout (split)
  OP_MERGE,        // BLSTM stitch + skip
  OP_ATTN,         // local attention core: qkv fp32 -> res split
  OP_TIMEFIRST, OP_FREQFIRST, OP_FINALFREQ, OP_FINALTIME
};

struct Op {
  int kind = 0;
  std::string name;          // conv name / parameter prefix
  Ten in, in2, out, aux;     // roles depend on kind (see hdemucs.cu where each is recorded)
  G2Problem pr;              // OP_CONV
  int dst_col = 0;           // OP_CONV: first column of `out` this conv wrote
  GnApply gn{};              // OP_GN (forward arguments)
  std::string p_gamma, p_beta, p_scale;  // OP_GN: parameter keys (empty = none)
  int i0 = 0, i1 = 0, i2 = 0, i3 = 0, i4 = 0, i5 = 0;
  float f0 = 0.0f;
  const float* fp0 = nullptr;
  const float* fp1 = nullptr;
};

}  // namespace hd
}  // namespace rfx

struct rfx_hdemucs {
  rfx_hdemucs_config cfg;
  std::map<std::string, rfx::hd::Buf> params;
  std::map<std::string, rfx::hd::Conv> convs;
  std::map<std::string, rfx::hd::Buf> whh;  // "<blstm>.l<layer>": W_hh of both directions [2][4H][H]
  bool finalized = false;
  rfx::hd::Buf gather_tmp;   // fp32 staging of one conv's gathered weights (stream-ordered reuse)
  rfx::hd::Buf gather_tmp2;  // fp32 staging of one conv's transposed weights (backward)
  // debug taps of the last call: name -> tensor descriptor
  std::map<std::string, rfx::hd::Ten> taps;
  bool want_taps = false;
  // training: tape of the last rfx_hdemucs_forward_train call
  std::vector<rfx::hd::Op> tape;
  int tape_B = 0, tape_T = 0;
  const void* tape_ws = nullptr;
  size_t fwd_bytes = 0;                 // workspace bytes the forward used (the backward allocates after them)
  const float* st_f = nullptr;          // per-item (mean, std) of the spectrogram / waveform (TA:553-563)
  const float* st_t = nullptr;
  std::map<const void*, float*> act_grads;  // after a backward: tensor key -> fp32 gradient (for rfx_hdemucs_grad_tap)
  std::map<std::string, const float*> inject;  // debug: tap name -> device gradient to substitute during the backward
  // the time branch of the forward runs on its own stream beside the frequency branch (hdemucs.cu: run_forward)
  cudaStream_t s_time = nullptr;
  cudaEvent_t ev_branch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // the backward builds the transposed weight packs of its input-gradient GEMMs on a third stream, ahead of the reverse replay
  cudaStream_t s_prep = nullptr;
  bool grads_prezeroed = false;  // rfx_hdemucs_set_grads_prezeroed: the caller hands over gradient buffers that are already zero
  std::vector<cudaEvent_t> ev_pack;
  ~rfx_hdemucs() {
    if (s_time) cudaStreamDestroy(s_time);
    if (s_prep) cudaStreamDestroy(s_prep);
    for (auto e : ev_pack) if (e) cudaEventDestroy(e);
    for (auto e : ev_branch) if (e) cudaEventDestroy(e);
    for (auto& kv : params) kv.second.release();
    for (auto& kv : convs) { kv.second.wbuf.release(); kv.second.bias.release(); kv.second.wtbuf.release(); }
    for (auto& kv : whh) kv.second.release();
    gather_tmp.release();
    gather_tmp2.release();
  }
};

namespace rfx {
namespace hd {

inline const float* HP(const rfx_hdemucs* h, const std::string& k) {
  auto it = h->params.find(k);
  return it == h->params.end() ? nullptr : it->second.p;
}

// hdemucs.cu
int hd_run_forward(rfx_hdemucs* h, const float* x, int B, int T, float* out, uint8_t* ws, bool dry, bool train, cudaStream_t s, size_t* bytes,
                   int* launches);
void hd_unsplit(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* o, long long n, cudaStream_t s);
// hdemucs_bwd.cu: replays h->tape in reverse.  dry = size the extra workspace only.
int hd_run_backward(rfx_hdemucs* h, const float* x, const float* dout, int B, int T, const std::map<std::string, float*>& grads, uint8_t* ws,
                    size_t ws_off, bool dry, cudaStream_t s, size_t* bytes);

}  // namespace hd
}  // namespace rfx
