// Hybrid Demucs (torchaudio.models.HDemucs as RemFx instantiates it: remfx/models.py:308-324,
// cfg/model/demucs.yaml:11-16) forward pass on the GPU.  torchaudio/models/_hdemucs.py line references: "TA:".
//
// Every convolution runs on the gemm2 tcgen05 engine as an implicit GEMM over channel-last split-bf16 activations:
//   * strided encoder convs (k=8, s=4, TA:124)  -> the input is VIEWED as (X/4, 4C) (free regrouping), which turns the
//     strided conv into a stride-1 3-tap conv with K = 4C;
//   * transposed decoder convs (k=8, s=4, TA:243) -> a 2-tap stride-1 conv producing N = 4*Cout, whose output VIEWED as
//     (4X, Cout) is the upsampled signal (the crop TA:288-294 is folded into the consumer);
//   * DConv dilated k=3 convs (TA:694), 1x1 convs, the decoder's 3x3 Conv2d "rewrite" (TA:249): plain multi-tap.
// GroupNorm needs whole-tensor statistics, so norm'd convs write fp32, a reduction kernel accumulates (sum, sum^2) in
// fp64 and a fused apply kernel does norm + GELU / GLU + LayerScale + residual and re-splits.  Un-normalised layers fuse
// bias + GELU or GLU into the GEMM epilogue.  Layout: freq branch (B, T, Fr, C), time branch (B, 1, L, C).
#include "hd_kernels.cuh"
#include "../../include/remfx_b200.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

using namespace rfx;
using namespace rfx::hd;

__global__ void hd_unsplit_kernel(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* o, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) o[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}
namespace rfx {
namespace hd {
void hd_unsplit(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* o, long long n, cudaStream_t s) {
  hd_unsplit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(hi, lo, o, n);
}
}  // namespace hd
}  // namespace rfx

namespace {

int prep_conv(rfx_hdemucs* h, const std::string& name, int kind, int Co, int Ci, int k, int s, int p, int glu, int kh, int kw, Buf& tmp,
              cudaStream_t st, const std::string& wkey_in = "", const std::string& bkey_in = "") {
  const std::string wkey = wkey_in.empty() ? name + ".weight" : wkey_in;
  const std::string bkey = bkey_in.empty() ? name + ".bias" : bkey_in;
  auto wit = h->params.find(wkey);
  if (wit == h->params.end()) { set_error("hdemucs: missing parameter '" + wkey + "'"); return 2; }
  const size_t expect = (size_t)Co * Ci * k;
  if (wit->second.n != expect) {
    set_error("hdemucs: parameter '" + wkey + "' has " + std::to_string(wit->second.n) + " elements, expected " + std::to_string(expect));
    return 2;
  }
  Conv& c = h->convs[name];
  c.Ci = Ci; c.Co = Co; c.kh = kh; c.kw = kw;
  c.wkey = wkey; c.bkey = h->params.count(bkey) ? bkey : std::string();
  c.wt_ready = false;  // the transposed pack of the backward is rebuilt lazily from the new weights
  GatherSpec& g = c.g;
  g.kind = kind; g.Co = Co; g.Ci = Ci; g.k = k; g.s = s; g.p = p; g.glu = glu;
  if (kind == 0) {
    g.taps = k; g.Kp = ceil_div(Ci, 64) * 64; g.Nout = Co; g.tau_min = 0;
  } else if (kind == 1) {
    const int qlo = -p, qhi = k - 1 - p;
    auto fl = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
    g.tau_min = fl(qlo, s);
    g.taps = fl(qhi, s) - g.tau_min + 1;
    g.Kp = ceil_div(s * Ci, 64) * 64; g.Nout = Co;
  } else {
    g.taps = k / s; g.Kp = ceil_div(Ci, 64) * 64; g.Nout = s * Co; g.tau_min = 0;
  }
  const size_t wn = (size_t)g.Nout * g.taps * g.Kp;
  if (tmp.n < wn) { if (tmp.alloc(wn)) return 1; }
  if (c.bias.alloc(g.Nout)) return 1;
  gather_w_kernel<<<148 * 4, 256, 0, st>>>(wit->second.p, HP(h, bkey), g, tmp.p, c.bias.p);
  RFX_CHECK_CUDA(cudaGetLastError());
  const int BN = g2_choose_bn(g.Nout);
  if (c.wbuf.alloc(split_weight_elems(g.Nout, g.taps * g.Kp, BN))) return 1;
  int rc = pack_split_weights(tmp.p, (long long)g.taps * g.Kp, g.Nout, g.taps * g.Kp, BN, reinterpret_cast<__nv_bfloat16*>(c.wbuf.p), &c.w, st);
  if (rc) return rc;
  return 0;  // tmp is reused by the next conv: same stream, so the reuse is ordered
}

// ------------------------------------------------------------------------------------------------
// One forward pass (or a dry run that only sizes the workspace)
// ------------------------------------------------------------------------------------------------
struct Runner {
  rfx_hdemucs* h;
  int B, T;
  uint8_t* ws;
  size_t off = 0;
  bool dry;
  cudaStream_t s;
  bool train = false;  // training forward: every conv keeps its fp32 pre-activation, ops are recorded on h->tape
  int rc = 0;
  int launches = 0;

  void* take(size_t bytes) {
    const size_t r = off;
    off += align_up(bytes, 256);
    // a dry run hands out distinct fake addresses (never dereferenced) so that a dry TRAINING run can still key its tape by tensor
    return dry ? reinterpret_cast<void*>((uintptr_t)4096 + r) : ws + r;
  }
  void record(const Op& op) {
    if (train) h->tape.push_back(op);
  }
  Ten split(int Bn, int Y, int X, int C) {
    Ten t; t.B = Bn; t.Y = Y; t.X = X; t.C = C;
    t.plane = align_up(t.elems() * 2, 256) / 2;
    t.hi = reinterpret_cast<__nv_bfloat16*>(take(t.plane * 2 * 2));
    return t;
  }
  Ten f32(int Bn, int Y, int X, int C) {
    Ten t; t.B = Bn; t.Y = Y; t.X = X; t.C = C;
    t.f = reinterpret_cast<float*>(take(t.elems() * 4));
    return t;
  }
  void tap(const std::string& name, const Ten& t) {
    if ((h->want_taps || train) && !dry) h->taps[name] = t;
  }
  bool ok() const { return rc == 0; }
  void chk() {
    if (!dry && rc == 0) {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { set_error(std::string("hdemucs kernel launch: ") + cudaGetErrorString(e)); rc = 1; }
    }
    ++launches;
  }

  // ---- generic implicit-GEMM convolution -------------------------------------------------------
  // axis: 0 = taps along X, 1 = taps along Y (plain 1-D convs); dil = dilation; pad = zero padding (plain)
  // out_f32: write fp32 (pre-norm) instead of split; act: epilogue activation (ACT_NONE / GELU / GLU_PAIR)
  // gn_G > 0: also accumulate GroupNorm statistics of the output in the GEMM epilogue; *gn_out receives (mean, rstd) stats
  // also_act (training only, with out_f32): the same launch also writes the ACTIVATED output as split planes into *also_act (created
  // and recorded here as the element-wise op the backward expects) while the fp32 tensor receives the pre-activation.
  Ten conv(const std::string& name, const Ten& in, int axis, int dil, int pad, bool out_f32, int act, const Ten* dst = nullptr,
           int dst_col = 0, int gn_G = 0, int gn_per_x = 0, float** gn_out = nullptr, Ten* also_act = nullptr) {
    auto it = h->convs.find(name);
    if (it == h->convs.end()) { set_error("hdemucs: conv '" + name + "' was not prepared"); rc = 2; return Ten(); }
    Conv& c = it->second;
    const GatherSpec& g = c.g;
    if (g.kind == 1 && in.X % g.s != 0) {
      // a strided conv zero-pads its input to a multiple of the stride (TA:147-150): lengths off the 1024-sample grid only
      if (train) { set_error("hdemucs: training needs chunk lengths that are multiples of 1024 samples"); rc = 2; return Ten(); }
      const int Xp = ceil_div(in.X, g.s) * g.s;
      Ten padded = split(in.B, in.Y, Xp, in.C);
      if (!dry && rc == 0) {
        const long long items = (long long)in.Y * Xp * (in.C / 8);
        pad_rows_kernel<<<dim3((unsigned)((items + 255) / 256), in.B), 256, 0, s>>>(in.hi, in.lo(), in.Y, in.X, Xp, in.C, padded.hi, padded.lo());
        chk();
      } else ++launches;
      return conv(name, padded, axis, dil, pad, out_f32, act, dst, dst_col, gn_G, gn_per_x, gn_out, also_act);
    }
    G2Problem pr;
    int Xv = in.X, Cv = in.C;  // input view
    int Xo = in.X, Yo = in.Y;  // output pixel grid
    if (g.kind == 1) {
      Xv = in.X / g.s; Cv = in.C * g.s;
      Xo = (in.X + 2 * g.p - g.k) / g.s + 1;
      for (int t = 0; t < g.taps; ++t) { pr.row_off[t] = g.tau_min + t; pr.row_off_y[t] = 0; }
    } else if (g.kind == 2) {
      Xo = in.X + g.taps - 1;
      for (int t = 0; t < g.taps; ++t) { pr.row_off[t] = -t; pr.row_off_y[t] = 0; }
    } else if (c.kh * c.kw > 1 && c.kh > 1 && c.kw > 1) {  // 2-D plain (kh along X, kw along Y), pad 1
      for (int a = 0; a < c.kh; ++a)
        for (int b2 = 0; b2 < c.kw; ++b2) { pr.row_off[a * c.kw + b2] = a - c.kh / 2; pr.row_off_y[a * c.kw + b2] = b2 - c.kw / 2; }
    } else {
      for (int t = 0; t < g.taps; ++t) {
        const int o = t * dil - pad;
        pr.row_off[t] = axis == 0 ? o : 0;
        pr.row_off_y[t] = axis == 0 ? 0 : o;
      }
    }
    const int Nout = g.Nout;
    if (train && !out_f32 && !dst) {
      // training: keep the pre-activation.  conv (bias only) -> fp32, then the activation as its own (recorded) element-wise op
      static const bool fuse = [] { const char* e = getenv("RFX_HD_TRAIN_FUSE_ACT"); return !(e && atoi(e) == 0); }();
      const int Cact = act == ACT_GLU_PAIR ? Nout / 2 : Nout;
      if (fuse && (act == ACT_GELU || act == ACT_GLU_PAIR) && Cact % 8 == 0 && Nout % 4 == 0) {
        // one launch: pre-activation -> fp32 (kept for the backward), activation -> split planes (G2Problem::cf_pre_act)
        Ten act_out;
        conv(name, in, axis, dil, pad, true, act, nullptr, 0, 0, 0, nullptr, &act_out);
        return act_out;
      }
      Ten raw = conv(name, in, axis, dil, pad, true, ACT_NONE);
      const int mode = act == ACT_GELU ? 1 : (act == ACT_GLU_PAIR ? 3 : 0);
      return gn_apply(raw, nullptr, 1, 0, nullptr, nullptr, mode, nullptr, nullptr, raw.X, 0, 0);
    }
    const int Cout_store = (act == ACT_GLU_PAIR && !also_act) ? Nout / 2 : Nout;
    Ten out;
    if (dst) out = *dst;  // write columns [dst_col, dst_col + N) of an existing fp32 tensor
    else out = out_f32 ? f32(in.B, Yo, Xo, Cout_store) : split(in.B, Yo, Xo, Cout_store);
    double* gacc = nullptr;
    float* gst = nullptr;
    int gn_cmod = 0, nseg = 0;
    long long gcount = 0;
    if (gn_G > 0) {
      gn_cmod = g.kind == 2 ? c.Co : Nout;               // transposed convs: channel = column % Cout
      const int Xs = g.kind == 2 ? Xo * g.s : Xo;        // spatial extent in the (upsampled) view
      nseg = gn_per_x ? in.B * Xo : in.B;
      gacc = reinterpret_cast<double*>(take((size_t)nseg * gn_G * 2 * 8));
      gst = reinterpret_cast<float*>(take((size_t)nseg * gn_G * 2 * 4));
      gcount = (gn_per_x ? (long long)Yo : (long long)Yo * Xs) * (gn_cmod / gn_G);
      if (gn_out) *gn_out = gst;
    }
    pr.A.hi = in.hi; pr.A.rows = Xv; pr.A.rows_y = in.Y; pr.A.ld = Cv; pr.A.ld_y = (long long)Xv * Cv;
    pr.A.batch_stride = (long long)in.Y * Xv * Cv; pr.A.plane_stride = (long long)in.plane;
    pr.M = Xo; pr.My = Yo; pr.N = Nout; pr.batch = in.B; pr.Ktap = Cv; pr.taps = g.taps;
    if (train) {  // (also in a dry run: the backward is sized from the tape)
      if (!out_f32 && !dst) { set_error("hdemucs: internal: training conv must produce fp32"); rc = 2; return out; }
      Op op; op.kind = OP_CONV; op.name = name; op.in = in; op.out = out; op.pr = pr; op.dst_col = dst ? dst_col : 0;
      record(op);
    }
    if (also_act) {  // the element-wise op's output tensor + tape record, without its launch
      const int mode = act == ACT_GELU ? 1 : (act == ACT_GLU_PAIR ? 3 : 0);
      *also_act = gn_apply(out, nullptr, 1, 0, nullptr, nullptr, mode, nullptr, nullptr, out.X, 0, 0, /*launch=*/false);
    }
    if (dry || rc) { launches += gn_G > 0 ? 3 : 1; return out; }
    if (gn_G > 0 && cudaMemsetAsync(gacc, 0, (size_t)nseg * gn_G * 2 * 8, s) != cudaSuccess) { set_error("memset failed"); rc = 1; return out; }
    pr.A.hi = in.hi; pr.A.rows = Xv; pr.A.rows_y = in.Y; pr.A.ld = Cv; pr.A.ld_y = (long long)Xv * Cv;
    pr.A.batch_stride = (long long)in.Y * Xv * Cv; pr.A.plane_stride = (long long)in.plane;
    pr.W = c.w;
    pr.M = Xo; pr.My = Yo; pr.N = Nout; pr.batch = in.B; pr.Ktap = Cv; pr.taps = g.taps;
    // pixel tile = xt x (128 / xt).  Default: the largest power of two <= Xo; when that pads the X axis by more than 10 % (Xo = 129,
    // 33, 9: transposed-conv outputs before their crop -> half-empty tiles, the largest single launch of the forward ran at 145 TF/s
    // instead of ~270) take the widest tile whose padded pixel count is within 3 % of the best
    int xt = 128;
    while (xt > Xo && xt > 1) xt >>= 1;
    {
      auto padded = [&](int w) { return (long long)ceil_div(Xo, w) * w * ((long long)ceil_div(Yo, 128 / w) * (128 / w)); };
      if (padded(xt) * 10 > (long long)Xo * Yo * 11) {
        long long best = padded(1);
        for (int w = 2; w <= xt; w <<= 1) best = std::min(best, padded(w));
        for (int w = xt; w >= 1; w >>= 1)
          if (padded(w) * 100 <= best * 103) { xt = w; break; }
      }
    }
    pr.xt = xt;
    if (dst) { pr.Cf = out.f + dst_col; pr.ldcf = out.C; pr.ldcf_y = (long long)Xo * out.C; pr.bscf = (long long)Yo * Xo * out.C; }
    else if (out_f32) { pr.Cf = out.f; pr.ldcf = Cout_store; pr.ldcf_y = (long long)Xo * Cout_store; pr.bscf = (long long)Yo * Xo * Cout_store; }
    else { pr.Chi = out.hi; pr.Clo = out.lo(); pr.ldcs = Cout_store; pr.ldcs_y = (long long)Xo * Cout_store; pr.bscs = (long long)Yo * Xo * Cout_store; }
    if (also_act) {
      const int Ca = also_act->C;
      pr.Chi = also_act->hi; pr.Clo = also_act->lo(); pr.ldcs = Ca; pr.ldcs_y = (long long)Xo * Ca; pr.bscs = (long long)Yo * Xo * Ca;
      pr.cf_pre_act = true;
    }
    pr.epi.t1 = c.bias.p;
    pr.epi.act = act;
    pr.gn_acc = gacc; pr.gn_G = gn_G > 0 ? gn_G : 1; pr.gn_per_x = gn_per_x; pr.gn_cmod = gn_cmod;
    rc = launch_gemm2(pr, s);
    ++launches;
    if (gn_G > 0 && rc == 0) {
      gn_final_kernel<<<ceil_div(nseg * gn_G, 256), 256, 0, s>>>(gacc, gcount, nseg * gn_G, 1e-5f, gst);
      chk();
      ++launches;
    }
    return out;
  }

  // ---- GroupNorm statistics of an fp32 tensor -> stats (mean, rstd) per (segment, group) ----
  float* gn_stats(const Ten& raw, int G, int per_x) {
    const int nseg = per_x ? raw.B * raw.X : raw.B;
    double* acc = reinterpret_cast<double*>(take((size_t)nseg * G * 2 * 8));
    float* st = reinterpret_cast<float*>(take((size_t)nseg * G * 2 * 4));
    if (dry || rc) { launches += 3; return st; }
    if (cudaMemsetAsync(acc, 0, (size_t)nseg * G * 2 * 8, s) != cudaSuccess) { set_error("memset failed"); rc = 1; return st; }
    const long long npix = per_x ? raw.Y : (long long)raw.Y * raw.X;
    int nsplit = (int)std::max<long long>(1, std::min<long long>((npix + 63) / 64, std::max(1, 1184 / nseg)));
    gn_accum_kernel<<<dim3(nsplit, nseg), 256, 0, s>>>(raw.f, raw.Y, raw.X, raw.C, G, per_x, acc);
    chk();
    const long long count = npix * (raw.C / G);
    gn_final_kernel<<<ceil_div(nseg * G, 256), 256, 0, s>>>(acc, count, nseg * G, 1e-5f, st);
    chk();
    ++launches;
    return st;
  }

  // parameter keys of the NEXT gn_apply's gamma / beta / scale (training: where their gradients go); set via P()
  std::string pending_gamma, pending_beta, pending_scale;
  const float* P(const std::string& key, int which) {
    (which == 0 ? pending_gamma : (which == 1 ? pending_beta : pending_scale)) = key;
    return HP(h, key);
  }

  // ---- norm (optional) + activation (+ LayerScale, + residual) -> split ----
  Ten gn_apply(const Ten& raw, const float* stats, int G, int per_x, const float* gamma, const float* beta, int mode, const float* scale,
               const Ten* res, int Xo, int x_off, int Cpad, bool launch = true) {
    const int Cvalid = mode >= 2 ? raw.C / 2 : raw.C;
    const int Co = Cpad > 0 ? Cpad : ceil_div(Cvalid, 8) * 8;
    Ten out = split(raw.B, raw.Y, Xo, Co);
    if (train) {
      Op op; op.kind = OP_GN; op.in = raw; op.out = out; if (res) op.in2 = *res;
      GnApply& ga = op.gn;
      ga.raw = raw.f; ga.Y = raw.Y; ga.Xr = raw.X; ga.Cr = raw.C; ga.stats = stats; ga.G = G; ga.per_x = per_x; ga.gamma = gamma; ga.beta = beta;
      ga.mode = mode; ga.scale = scale; ga.rhi = res ? res->hi : nullptr; ga.rlo = res ? res->lo() : nullptr;
      ga.ohi = out.hi; ga.olo = out.lo(); ga.Xo = Xo; ga.Co = Co; ga.x_off = x_off;
      op.p_gamma = pending_gamma; op.p_beta = pending_beta; op.p_scale = pending_scale;
      record(op);
    }
    pending_gamma.clear(); pending_beta.clear(); pending_scale.clear();
    if (!launch) return out;   // the producing GEMM writes `out` itself (conv(..., also_act))
    if (dry || rc) { ++launches; return out; }
    GnApply a{};
    a.raw = raw.f; a.Y = raw.Y; a.Xr = raw.X; a.Cr = raw.C;
    a.stats = stats; a.G = G; a.per_x = per_x; a.gamma = gamma; a.beta = beta; a.mode = mode; a.scale = scale;
    a.rhi = res ? res->hi : nullptr; a.rlo = res ? res->lo() : nullptr;
    a.ohi = out.hi; a.olo = out.lo(); a.Xo = Xo; a.Co = Co; a.x_off = x_off;
    const long long items = (long long)raw.Y * Xo * (Co / 8);
    if (items >= (1LL << 31)) { set_error("gn_apply: item count overflows 32 bits"); rc = 2; return out; }
    const bool al16 = ((((uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)scale | (uintptr_t)raw.f) & 15) == 0);
    const bool vec = al16 && Cvalid % 8 == 0 && raw.C % 4 == 0 && (!stats || (raw.C / G) % 8 == 0);
    if (vec) gn_apply_kernel<true><<<dim3((unsigned)((items + 255) / 256), raw.B), 256, 0, s>>>(a);
    else gn_apply_kernel<false><<<dim3((unsigned)((items + 255) / 256), raw.B), 256, 0, s>>>(a);
    chk();
    return out;
  }

  Ten add_crop(const Ten& a, int x_off, const Ten& skip) {
    Ten out = split(skip.B, skip.Y, skip.X, skip.C);
    if (train) { Op op; op.kind = OP_ADDCROP; op.in = a; op.in2 = skip; op.out = out; op.i0 = x_off; record(op); }
    if (dry || rc) { ++launches; return out; }
    const long long items = (long long)skip.Y * skip.X * (skip.C / 8);
    if (a.Y < skip.Y || a.X < skip.X + x_off) { set_error("hdemucs: internal: skip sum reads beyond its input"); rc = 2; return out; }
    add_crop_kernel<<<dim3((unsigned)((items + 255) / 256), skip.B), 256, 0, s>>>(a.hi, a.lo(), a.Y, a.X, x_off, skip.hi, skip.lo(), out.hi, out.lo(),
                                                                                  skip.Y, skip.X, skip.C);
    chk();
    return out;
  }

  // ---- _BLSTM (TA:742-788): 2-layer BiLSTM (hidden = C) over overlapping 200-step frames, Linear(2C -> C), stitch, + skip ----
  Ten blstm(const std::string& base, const Ten& x) {  // x: split (B, 1, T, C)
    const int C = x.C, Tn = x.X, Bn = x.B;
    const int width = 200, stride = 100;
    const bool framed = Tn > width;
    const int nf = framed ? ceil_div(Tn, stride) : 1;
    const int Tf = framed ? width : Tn;
    Ten cur = x;
    if (framed) {
      cur = split(Bn * nf, 1, Tf, C);
      if (train) { Op op; op.kind = OP_FRAME; op.in = x; op.out = cur; op.i0 = nf; op.i1 = width; op.i2 = stride; record(op); }
      if (!dry && ok()) {
        const long long items = (long long)Tf * (C / 8);
        blstm_frame_kernel<<<dim3((unsigned)((items + 255) / 256), Bn * nf), 256, 0, s>>>(x.hi, x.lo(), Tn, C, nf, width, stride, cur.hi, cur.lo());
        chk();
      } else ++launches;
    }
    const int Bs = Bn * nf;
    for (int l = 0; l < 2 && ok(); ++l) {
      Ten G = f32(Bs, 1, Tf, 8 * C);
      conv(base + ".lstm.ih" + std::to_string(l) + "f", cur, 0, 1, 0, true, ACT_NONE, &G, 0);
      conv(base + ".lstm.ih" + std::to_string(l) + "r", cur, 0, 1, 0, true, ACT_NONE, &G, 4 * C);
      Ten hout = split(Bs, 1, Tf, 2 * C);
      if (train) { Op op; op.kind = OP_LSTM; op.name = base; op.in = cur; op.aux = G; op.out = hout; op.i0 = l; record(op); }
      if (!dry && ok()) {
        rc = launch_lstm_layer(G.f, 8 * C, h->whh[base + ".l" + std::to_string(l)].p, nullptr, 0, hout.hi, hout.lo(), 2 * C, Bs, Tf, C, s);
      }
      ++launches;
      cur = hout;
    }
    Ten lin = conv(base + ".linear", cur, 0, 1, 0, true, ACT_NONE);  // fp32 (Bs, 1, Tf, C)
    Ten out = split(Bn, 1, Tn, C);
    if (train) { Op op; op.kind = OP_MERGE; op.in = lin; op.in2 = x; op.out = out; op.i0 = nf; op.i1 = Tf; op.i2 = stride; record(op); }
    if (!dry && ok()) {
      const long long items = (long long)Tn * (C / 8);
      blstm_merge_kernel<<<dim3((unsigned)((items + 255) / 256), Bn), 256, 0, s>>>(lin.f, Tn, C, nf, Tf, stride, x.hi, x.lo(), out.hi, out.lo());
      chk();
    } else ++launches;
    return out;
  }

  // ---- _LocalState (TA:822-857): x + proj(attention(x)) ----
  Ten local_state(const std::string& base, const Ten& x) {  // x: split (B, 1, T, C)
    const int C = x.C, Tn = x.X, Bn = x.B, heads = 4, nd = 4;
    const int ld = 3 * C + heads * nd;
    Ten qkv = f32(Bn, 1, Tn, ld);
    conv(base + ".query", x, 0, 1, 0, true, ACT_NONE, &qkv, 0);
    conv(base + ".key", x, 0, 1, 0, true, ACT_NONE, &qkv, C);
    conv(base + ".content", x, 0, 1, 0, true, ACT_NONE, &qkv, 2 * C);
    conv(base + ".query_decay", x, 0, 1, 0, true, ACT_NONE, &qkv, 3 * C);
    Ten res = split(Bn, 1, Tn, C);
    if (train) { Op op; op.kind = OP_ATTN; op.in = qkv; op.out = res; op.i0 = heads; op.i1 = nd; record(op); }
    if (!dry && ok()) {
      const int Ch = C / heads;
      const size_t smem = ((size_t)2 * Tn * (Ch + 1) + (size_t)Tn * (LA_QT + 1) + (size_t)LA_QT * (Ch + 1) + LA_QT) * 4;
      if (smem > 220 * 1024) { set_error("hdemucs: LocalState sequence too long for the shared-memory attention kernel"); rc = 2; return res; }
      cudaFuncSetAttribute(local_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      local_attn_kernel<<<dim3(ceil_div(Tn, LA_QT), heads, Bn), 256, smem, s>>>(qkv.f, ld, Tn, C, heads, nd, res.hi, res.lo());
      chk();
    } else ++launches;
    Ten pr = conv(base + ".proj", res, 0, 1, 0, true, ACT_NONE);
    return gn_apply(pr, nullptr, 1, 0, nullptr, nullptr, 0, nullptr, &x, pr.X, 0, 0);
  }

  // ---- DConv residual branch (TA:709-721); y: split (B, Y, X, C); axis = conv axis (1 = Y for the freq branch) ----
  Ten dconv(const std::string& base, Ten y, int axis, int per_x, bool lstm_attn) {
    const int depth = h->cfg.dconv_depth;
    for (int d = 0; d < depth && ok(); ++d) {
      const std::string L = base + ".layers." + std::to_string(d);
      const int dil = 1 << d;
      // conv k=3 (dilated) -> GroupNorm(1, h) -> GELU
      float* st1 = nullptr;
      Ten r1 = conv(L + ".0", y, axis, dil, dil, true, ACT_NONE, nullptr, 0, 1, per_x, &st1);
      Ten a1 = gn_apply(r1, st1, 1, per_x, P(L + ".1.weight", 0), P(L + ".1.bias", 1), 1, nullptr, nullptr, r1.X, 0, 0);
      tap(L + ".2", a1);
      int ci = 3;
      if (lstm_attn) {
        // the tensors here are (B, T, 1, C) / (B, 1, T, C): the same memory; work on the time-like view
        Ten z = a1;
        if (z.Y > 1) { z.X = z.Y * z.X; z.Y = 1; }
        z = blstm(L + ".3", z);
        z = local_state(L + ".4", z);
        tap(L + ".4", z);
        a1 = z;
        if (y.Y > 1) { a1.Y = y.Y; a1.X = y.X; }
        ci = 5;
      }
      // 1x1 conv h -> 2C -> GroupNorm(1, 2C) -> GLU -> LayerScale -> residual
      float* st2 = nullptr;
      Ten r2 = conv(L + "." + std::to_string(ci), a1, 0, 1, 0, true, ACT_NONE, nullptr, 0, 1, per_x, &st2);
      const std::string gn2 = L + "." + std::to_string(ci + 1), ls = L + "." + std::to_string(ci + 3);
      y = gn_apply(r2, st2, 1, per_x, P(gn2 + ".weight", 0), P(gn2 + ".bias", 1), 2, P(ls + ".scale", 2), &y, r2.X, 0, 0);
    }
    return y;
  }
};

int run_forward(rfx_hdemucs* h, const float* x, int B, int T, float* out, uint8_t* ws, bool dry, bool train, cudaStream_t s, size_t* bytes,
                int* launches) {
  Runner R{h, B, T, ws, 0, dry, s};
  R.train = train;
  if (train) { h->tape.clear(); h->tape_B = B; h->tape_T = T; h->tape_ws = ws; h->act_grads.clear(); }
  const int nfft = h->cfg.nfft, hl = nfft / 4, bins = nfft / 2, depth = h->cfg.depth, ch0 = h->cfg.channels;
  const int le = ceil_div(T, hl);
  const bool lstm_attn_from = true;
  (void)lstm_attn_from;
  if (!dry) h->taps.clear();
  // The time branch (1-D convs on the waveform) and the frequency branch (convs on the spectrogram) only meet at the innermost
  // encoder layer, at the first time decoder and in the final sum: the time branch runs on its own stream, so that its launches
  // fill the SMs the frequency branch's small / latency-bound launches leave idle (two whole forwards side by side measured 1.11x
  // at B = 32, 1.18x at 16, 1.49x at 1: tools/hd_concurrency_probe.py).  RFX_HD_OVERLAP=0 keeps everything on the caller's stream.
  static const bool overlap_on = [] { const char* e = getenv("RFX_HD_OVERLAP"); return !(e && atoi(e) == 0); }();
  bool overlap = overlap_on && !dry;
  if (overlap && !h->s_time) {
    if (cudaStreamCreateWithFlags(&h->s_time, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); h->s_time = nullptr; overlap = false; }
    for (auto& e : h->ev_branch)
      if (overlap && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); overlap = false; }
  }
  cudaStream_t st = overlap ? h->s_time : s;  // the time branch's stream
  auto hand = [&](int e, cudaStream_t from, cudaStream_t to) {  // `to` continues only after everything queued on `from` so far
    if (overlap && (cudaEventRecord(h->ev_branch[e], from) != cudaSuccess || cudaStreamWaitEvent(to, h->ev_branch[e], 0) != cudaSuccess)) {
      set_error("hdemucs: stream hand-over failed");
      R.rc = 1;
    }
  };

  // ---------------- D1 / D2: spectrogram, (re, im) as channels, per-item normalisation (TA:465-487, 509-514, 553-563) ----
  float2* Z = reinterpret_cast<float2*>(R.take((size_t)B * le * bins * 8));
  float* st_f = reinterpret_cast<float*>(R.take((size_t)B * 2 * 4));
  float* st_t = reinterpret_cast<float*>(R.take((size_t)B * 2 * 4));
  float* xt = reinterpret_cast<float*>(R.take((size_t)B * T * 4));
  if (train) { h->st_f = st_f; h->st_t = st_t; }
  if (!dry) {
    StftParams sp{};
    sp.x = x; sp.x_bstride = T; sp.T = T; sp.x_aligned8 = 0;
    sp.window = HP(h, "__window__"); sp.tw = twiddles(nfft);
    sp.n_fft = nfft; sp.hop = hl; sp.F = le; sp.frame_off = hl / 2 * 3; sp.nbins = bins;
    sp.scale = 1.0f / sqrtf((float)nfft); sp.alpha = 1.0f; sp.mode = STFT_COMPLEX;
    sp.Z = Z; sp.ldz = bins;
    if ((R.rc = launch_stft(sp, B, s))) return R.rc;
    item_stats_kernel<<<B, 1024, 0, s>>>(reinterpret_cast<const float*>(Z), (long long)le * bins * 2, st_f);
    item_stats_kernel<<<B, 1024, 0, s>>>(x, (long long)T, st_t);
    if (T % 8 == 0 && ((uintptr_t)x & 15) == 0)
      item_normalize_kernel<<<dim3((unsigned)((T / 8 + 255) / 256), B), 256, 0, s>>>(x, (long long)T, st_t, nullptr, nullptr, xt);
    else
      item_normalize_scalar_kernel<<<dim3((unsigned)((T + 255) / 256), B), 256, 0, s>>>(x, (long long)T, st_t, xt);
    R.chk();
  }
  R.launches += 4;
  hand(0, s, st);  // the normalised waveform is ready

  // ---------------- encoders (TA:565-593) ----------------
  std::vector<Ten> saved, saved_t;
  Ten xcur;        // freq branch
  Ten tcur;        // time branch (split) after layer 0
  Ten inject;      // fp32, time conv output of the merge layer
  int freqs = bins;
  for (int idx = 0; idx < depth && R.ok(); ++idx) {
    const bool lstm_attn = idx >= h->cfg.dconv_lstm;  // lstm and attn start at the same layer in RemFx's config
    const bool normed = idx >= h->cfg.norm_starts;
    const bool freq = freqs > 1;
    const std::string fe = "freq_encoder." + std::to_string(idx), te = "time_encoder." + std::to_string(idx);
    if (freq) {
      const bool last_freq = freqs <= h->cfg.kernel_size;
      // ---- time branch ----
      R.s = st;
      if (idx == 0) {
        const int Lo = ceil_div(T, h->cfg.stride);   // (the conv zero-pads T to a multiple of its stride)
        tcur = R.split(B, 1, Lo, ch0);
        if (train) {
          Op op; op.kind = OP_TIMEFIRST; op.name = te + ".conv"; op.out = tcur; op.fp0 = xt; op.i0 = T; op.i1 = h->cfg.stride; op.i2 = h->cfg.kernel_size / 4;
          R.record(op);
        }
        if (!dry) {
          const long long items = (long long)Lo * (ch0 / 8);
          time_first_kernel<8><<<dim3((unsigned)((items + 255) / 256), B), 256, (8 + 1) * ch0 * 4, R.s>>>(
              xt, T, Lo, ch0, h->cfg.stride, h->cfg.kernel_size / 4, HP(h, te + ".conv.weight"), HP(h, te + ".conv.bias"), tcur.hi, tcur.lo());
          R.chk();
        } else ++R.launches;
      } else if (!last_freq) {
        tcur = R.conv(te + ".conv", tcur, 0, 1, 0, false, ACT_GELU);
      } else {
        inject = R.conv(te + ".conv", tcur, 0, 1, 0, true, ACT_NONE);  // "empty" layer: just the conv (TA:159-160)
      }
      if (!last_freq) {
        tcur = R.dconv(te + ".dconv", tcur, 0, 0, lstm_attn);
        tcur = R.conv(te + ".rewrite", tcur, 0, 1, 0, false, ACT_GLU_PAIR);
        R.tap(te, tcur);
        saved_t.push_back(tcur);
      }
      R.s = s;
      if (last_freq) hand(1, st, s);  // the frequency branch adds `inject` below
      // ---- freq branch ----
      if (idx == 0) {  // 2 -> C channels: SIMT kernel that normalises (TA:553-557) on the fly
        const int Fo = bins / h->cfg.stride;
        xcur = R.split(B, le, Fo, ch0);
        if (train) {
          Op op; op.kind = OP_FREQFIRST; op.name = fe + ".conv"; op.out = xcur; op.fp0 = reinterpret_cast<const float*>(Z); op.fp1 = st_f;
          op.i0 = bins; op.i1 = h->cfg.stride; op.i2 = h->cfg.kernel_size / 4;
          R.record(op);
        }
        if (!dry && R.ok()) {
          const long long items = (long long)le * Fo * (ch0 / 8);
          freq_first_kernel<8><<<dim3((unsigned)((items + 255) / 256), B), 256, (2 * 8 + 1) * ch0 * 4, s>>>(
              reinterpret_cast<const float*>(Z), st_f, le, bins, Fo, ch0, h->cfg.stride, h->cfg.kernel_size / 4, HP(h, fe + ".conv.weight"),
              HP(h, fe + ".conv.bias"), xcur.hi, xcur.lo());
          R.chk();
        } else ++R.launches;
      } else if (!normed) {
        xcur = R.conv(fe + ".conv", xcur, 0, 1, 0, false, ACT_GELU);
      } else {
        Ten raw = R.conv(fe + ".conv", xcur, 0, 1, 0, true, ACT_NONE);  // (B, T, 1, C)
        if (last_freq) {  // y = y + inject (TA:164-169); both are [B][T][C] in memory
          Ten sum = R.f32(raw.B, raw.Y, raw.X, raw.C);
          if (train) { Op op; op.kind = OP_ADDF32; op.in = raw; op.in2 = inject; op.out = sum; R.record(op); }
          if (!dry && R.ok()) {
            add_f32_kernel<<<(unsigned)((raw.elems() + 255) / 256), 256, 0, s>>>(raw.f, inject.f, sum.f, (long long)raw.elems());
            R.chk();
          } else ++R.launches;
          raw = sum;
        }
        float* st = R.gn_stats(raw, h->cfg.norm_groups, 0);
        xcur = R.gn_apply(raw, st, h->cfg.norm_groups, 0, R.P(fe + ".norm1.weight", 0), R.P(fe + ".norm1.bias", 1), 1, nullptr, nullptr, raw.X, 0, 0);
      }
      R.tap(fe + ".act1", xcur);
      xcur = R.dconv(fe + ".dconv", xcur, 1, 1, lstm_attn);
      if (!normed) {
        xcur = R.conv(fe + ".rewrite", xcur, 0, 1, 0, false, ACT_GLU_PAIR);
      } else {
        float* st = nullptr;
        Ten raw = R.conv(fe + ".rewrite", xcur, 0, 1, 0, true, ACT_NONE, nullptr, 0, h->cfg.norm_groups, 0, &st);
        xcur = R.gn_apply(raw, st, h->cfg.norm_groups, 0, R.P(fe + ".norm2.weight", 0), R.P(fe + ".norm2.bias", 1), 2, nullptr, nullptr, raw.X, 0, 0);
      }
      if (idx == 0 && h->cfg.freq_emb_weight != 0.0f) {  // TA:586-591
        if (train) {
          Op op; op.kind = OP_FREQEMB; op.name = "freq_emb.embedding.weight"; op.out = xcur; op.f0 = h->cfg.freq_emb_weight * h->cfg.freq_emb_scale;
          R.record(op);
        }
        if (!dry && R.ok()) {
          const long long items = (long long)xcur.Y * xcur.X * (xcur.C / 8);
          freq_emb_kernel<<<dim3((unsigned)((items + 255) / 256), B), 256, 0, s>>>(xcur.hi, xcur.lo(), xcur.Y, xcur.X, xcur.C,
                                                                                   HP(h, "freq_emb.embedding.weight"),
                                                                                   h->cfg.freq_emb_weight * h->cfg.freq_emb_scale);
          R.chk();
        } else ++R.launches;
      }
      R.tap(fe, xcur);
      saved.push_back(xcur);
      freqs = last_freq ? 1 : freqs / h->cfg.stride;
    } else {
      // merged layer (freq == false): Conv1d(k = 2*time_stride, s = time_stride, pad) on (B, 1, T, C) (TA:389-399)
      Ten xin = xcur;  // (B, T, 1, C) has the same memory order as (B, 1, T, C)
      xin.X = xcur.Y; xin.Y = 1;
      float* st = nullptr;
      Ten raw = R.conv(fe + ".conv", xin, 0, 1, 0, true, ACT_NONE, nullptr, 0, h->cfg.norm_groups, 0, &st);
      Ten y = R.gn_apply(raw, st, h->cfg.norm_groups, 0, R.P(fe + ".norm1.weight", 0), R.P(fe + ".norm1.bias", 1), 1, nullptr, nullptr, raw.X, 0, 0);
      y = R.dconv(fe + ".dconv", y, 0, 0, lstm_attn);
      float* st2 = nullptr;
      Ten raw2 = R.conv(fe + ".rewrite", y, 0, 1, 0, true, ACT_NONE, nullptr, 0, h->cfg.norm_groups, 0, &st2);
      xcur = R.gn_apply(raw2, st2, h->cfg.norm_groups, 0, R.P(fe + ".norm2.weight", 0), R.P(fe + ".norm2.bias", 1), 2, nullptr, nullptr, raw2.X, 0, 0);
      R.tap(fe, xcur);
      saved.push_back(xcur);
    }
  }

  // ---------------- decoders (TA:595-615); the decoder input is all-zero, so `x + skip` = skip at the first layer ----
  Ten xd, xtd;          // current decoder activations (split; transposed-conv outputs are kept UNcropped, see crop_*)
  int crop_f = 0, crop_t = 0;  // pending crop offset of xd / xtd along X
  bool have_x = false;
  const int n_tdec = (int)saved_t.size() + 1;  // time decoders incl. the "empty" one
  const int offset = depth - n_tdec;
  for (int idx = 0; idx < depth && R.ok(); ++idx) {
    const std::string fd = "freq_decoder." + std::to_string(idx);
    const int enc_idx = depth - 1 - idx;
    const bool normed = enc_idx >= h->cfg.norm_starts;
    const bool last = enc_idx == 0;
    Ten skip = saved.back();
    saved.pop_back();
    const bool freq_layer = skip.Y > 1 || (enc_idx < depth - 1);  // only the innermost layer is the merged (time-like) one
    Ten xin;
    if (!have_x) { xin = skip; have_x = true; }
    else {
      Ten xv = xd;
      if (freq_layer && xv.Y == 1) { xv.Y = xv.X; xv.X = 1; }  // (B, 1, T, C) -> (B, T, 1, C): same memory (TA:282-284)
      xin = R.add_crop(xv, crop_f, skip);
    }
    // rewrite (k=3 / 3x3, C -> 2C) [+ norm1] + GLU
    Ten y;
    const bool conv2d = h->convs[fd + ".rewrite"].kh > 1 && h->convs[fd + ".rewrite"].kw > 1;
    (void)conv2d;
    if (!normed) {
      y = R.conv(fd + ".rewrite", xin, 0, 1, 1, false, ACT_GLU_PAIR);
    } else {
      float* st = nullptr;
      Ten raw = R.conv(fd + ".rewrite", xin, xin.Y > 1 ? 1 : 0, 1, 1, true, ACT_NONE, nullptr, 0, h->cfg.norm_groups, 0, &st);
      y = R.gn_apply(raw, st, h->cfg.norm_groups, 0, R.P(fd + ".norm1.weight", 0), R.P(fd + ".norm1.bias", 1), 2, nullptr, nullptr, raw.X, 0, 0);
    }
    R.tap(fd + ".pre", y);
    // transposed conv [+ norm2] [+ GELU]
    const Conv& ctr = h->convs[fd + ".conv_tr"];
    const int pad = ctr.crop;  // (k - s) / 2, or 0 for the last_freq layer (TA:419-423)
    if (last) {
      // transposed conv (C -> 2) + crop + de-normalise -> complex, then iSTFT (TA:287-294, 516-521, 489-497, 624-633)
      float2* Zo = reinterpret_cast<float2*>(R.take((size_t)B * le * bins * 8));
      if (train) {
        Op op; op.kind = OP_FINALFREQ; op.name = fd + ".conv_tr"; op.in = y; op.i0 = pad; op.i1 = bins; op.i2 = le; op.i3 = nfft; op.fp1 = st_f;
        R.record(op);
      }
      if (!dry && R.ok()) {
        const long long items = (long long)le * bins;
        final_freq_convtr_kernel<8, 4><<<dim3((unsigned)((items + 255) / 256), B), 256, 8 * 2 * (y.C + 1) * 4, s>>>(
            y.hi, y.lo(), le, y.X, y.C, pad, bins, HP(h, fd + ".conv_tr.weight"), HP(h, fd + ".conv_tr.bias"), st_f, Zo);
        R.chk();
        IstftParams ip{};
        ip.Z = Zo; ip.ldz = bins; ip.mask = nullptr; ip.ldm = 0;
        ip.window = HP(h, "__window__"); ip.tw = twiddles(nfft);
        ip.n_fft = nfft; ip.hop = hl; ip.F = le; ip.length = T;
        ip.frame_off = hl / 2 * 3; ip.env_pad = 2; ip.nbins = bins;
        ip.scale = sqrtf((float)nfft); ip.out = out; ip.out_bstride = T; ip.hops_per_cta = 8;
        if ((R.rc = launch_istft(ip, B, s))) return R.rc;
      }
      R.launches += 2;
    } else if (!normed) {
      xd = R.conv(fd + ".conv_tr", y, 0, 1, 0, false, ACT_GELU);  // bias + GELU fused; crop deferred to the next add
      xd.X *= ctr.g.s; xd.C /= ctr.g.s;                           // view (Xg, s*Co) as (Xg*s, Co)
      crop_f = pad;
    } else {
      float* st = nullptr;
      Ten raw = R.conv(fd + ".conv_tr", y, 0, 1, 0, true, ACT_NONE, nullptr, 0, h->cfg.norm_groups, 0, &st);
      raw.X *= ctr.g.s; raw.C /= ctr.g.s;
      const int len = raw.X - 2 * pad;
      xd = R.gn_apply(raw, st, h->cfg.norm_groups, 0, R.P(fd + ".norm2.weight", 0), R.P(fd + ".norm2.bias", 1), 1, nullptr, nullptr, len, pad, 0);
      crop_f = 0;
    }
    if (!last) R.tap(fd, xd);

    // ---- time decoder ----
    if (idx >= offset) {
      const int ti = idx - offset;
      const std::string td = "time_decoder." + std::to_string(ti);
      const Conv& ttr = h->convs[td + ".conv_tr"];
      const int tpad = ttr.crop;
      const bool tnormed = normed;
      Ten yt;
      if (ti == 0) hand(2, s, st);  // the first time decoder starts from the frequency branch's `pre`
      R.s = st;
      if (ti == 0) {
        yt = y;  // "empty" decoder: pre[:, :, 0] (TA:603-607); (B, T, 1, C) == (B, 1, T, C) in memory
        yt.X = y.Y * y.X; yt.Y = 1;
      } else {
        Ten tskip = saved_t.back();
        saved_t.pop_back();
        Ten tin = R.add_crop(xtd, crop_t, tskip);
        yt = R.conv(td + ".rewrite", tin, 0, 1, 1, false, ACT_GLU_PAIR);
      }
      if (last) {
        if (train) {
          Op op; op.kind = OP_FINALTIME; op.name = td + ".conv_tr"; op.in = yt; op.i0 = tpad; op.i1 = ttr.g.k; op.i2 = ttr.g.s; op.fp1 = st_t;
          R.record(op);
        }
        hand(3, s, st);  // `out` holds the inverse STFT of the frequency branch: the time branch is added to it
        if (!dry && R.ok()) {
          final_time_kernel<<<dim3((unsigned)((T + 255) / 256), B), 256, 0, R.s>>>(yt.hi, yt.lo(), yt.X, yt.C, ttr.g.k, ttr.g.s, tpad,
                                                                                 HP(h, td + ".conv_tr.weight"), HP(h, td + ".conv_tr.bias"),
                                                                                 st_t, T, out);
          R.chk();
        } else ++R.launches;
      } else if (!tnormed) {
        xtd = R.conv(td + ".conv_tr", yt, 0, 1, 0, false, ACT_GELU);
        xtd.X *= ttr.g.s; xtd.C /= ttr.g.s;
        crop_t = tpad;
      } else {
        float* st = nullptr;
        Ten raw = R.conv(td + ".conv_tr", yt, 0, 1, 0, true, ACT_NONE, nullptr, 0, h->cfg.norm_groups, 0, &st);
        raw.X *= ttr.g.s; raw.C /= ttr.g.s;
        const int len = raw.X - 2 * tpad;
        xtd = R.gn_apply(raw, st, h->cfg.norm_groups, 0, R.P(td + ".norm2.weight", 0), R.P(td + ".norm2.bias", 1), 1, nullptr, nullptr, len, tpad, 0);
        crop_t = 0;
      }
      if (!last) R.tap(td, xtd);
      R.s = s;
    }
  }
  hand(4, st, s);  // join: the caller's stream sees the finished output
  if (bytes) *bytes = R.off;
  if (launches) *launches = R.launches;
  if (train) h->fwd_bytes = R.off;
  return R.rc;
}

}  // namespace

namespace rfx {
namespace hd {
// re-gather one prepared conv's fp32 weights [Nout][taps][Kp] (the backward transposes them for the input-gradient GEMM)
int hd_gather_weights(rfx_hdemucs* h, const Conv& c, float* dst, cudaStream_t st) {
  const float* w = HP(h, c.wkey);
  RFX_REQUIRE(w != nullptr, "hdemucs: conv weight is not loaded");
  gather_w_kernel<<<148 * 4, 256, 0, st>>>(w, nullptr, c.g, dst, nullptr);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int hd_run_forward(rfx_hdemucs* h, const float* x, int B, int T, float* out, uint8_t* ws, bool dry, bool train, cudaStream_t s, size_t* bytes,
                   int* launches) {
  return run_forward(h, x, B, T, out, ws, dry, train, s, bytes, launches);
}
}  // namespace hd
}  // namespace rfx

extern "C" {

int rfx_hdemucs_create(const rfx_hdemucs_config* cfg, rfx_hdemucs_t** out) {
  RFX_REQUIRE(cfg && out, "null argument");
  RFX_REQUIRE(cfg->audio_channels == 1 && cfg->n_sources == 1, "HDemucs: only audio_channels = 1, one source (cfg/model/demucs.yaml)");
  RFX_REQUIRE(cfg->nfft == 4096, "HDemucs: nfft must be 4096");
  RFX_REQUIRE(cfg->depth == 6 && cfg->kernel_size == 8 && cfg->stride == 4 && cfg->time_stride == 2 && cfg->growth == 2,
              "HDemucs: depth 6, kernel 8, stride 4, time_stride 2, growth 2 only");
  RFX_REQUIRE(cfg->channels % 8 == 0 && cfg->channels >= 16, "HDemucs: channels must be a multiple of 8");
  RFX_REQUIRE(cfg->context == 1 && cfg->context_enc == 0, "HDemucs: context 1 / context_enc 0 only");
  RFX_REQUIRE(cfg->dconv_depth >= 1 && cfg->dconv_comp == 4, "HDemucs: dconv_comp 4 only");
  RFX_REQUIRE(cfg->dconv_lstm == cfg->dconv_attn, "HDemucs: dconv_lstm and dconv_attn must start at the same layer");
  rfx_hdemucs* h = new rfx_hdemucs();
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void rfx_hdemucs_destroy(rfx_hdemucs_t* h) { delete h; }

int rfx_hdemucs_load_param(rfx_hdemucs_t* h, const char* key, const float* src, int64_t numel, void* stream) {
  RFX_REQUIRE(h && key && src && numel > 0, "bad argument");
  Buf& b = h->params[key];
  if (b.n != (size_t)numel) {
    if (b.alloc((size_t)numel)) return 1;
  }
  RFX_CHECK_CUDA(cudaMemcpyAsync(b.p, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  h->finalized = false;
  return 0;
}

int rfx_hdemucs_finalize(rfx_hdemucs_t* h, void* stream) {
  RFX_REQUIRE(h, "null handle");
  cudaStream_t s = (cudaStream_t)stream;
  const rfx_hdemucs_config& c = h->cfg;
  RFX_REQUIRE(h->params.count("__window__") && h->params["__window__"].n == (size_t)c.nfft, "hdemucs: load the hann window as '__window__'");
  Buf& tmp = h->gather_tmp;  // kept across calls: a training loop re-finalizes after every optimiser step
  int rc = 0;
  int chin = c.audio_channels, chin_z = 2 * c.audio_channels, chout = c.channels, chout_z = c.channels;
  int freqs = c.nfft / 2;
  for (int idx = 0; idx < c.depth && !rc; ++idx) {
    const bool la = idx >= c.dconv_lstm;
    const bool freq = freqs > 1;
    int ker = c.kernel_size, stri = c.stride;
    bool pad = true, last_freq = false;
    if (!freq) { ker = c.time_stride * 2; stri = c.time_stride; }
    if (freq && freqs <= c.kernel_size) { ker = freqs; pad = false; last_freq = true; }
    if (last_freq) { chout_z = std::max(chout, chout_z); chout = chout_z; }
    const std::string fe = "freq_encoder." + std::to_string(idx), te = "time_encoder." + std::to_string(idx);
    const std::string fd = "freq_decoder." + std::to_string(c.depth - 1 - idx), td = "time_decoder." + std::to_string(c.depth - 2 - idx);
    // encoder convs
    if (idx > 0) rc = prep_conv(h, fe + ".conv", 1, chout_z, chin_z, ker, stri, pad ? ker / 4 : 0, 0, 1, 1, tmp, s);
    if (!rc) rc = prep_conv(h, fe + ".rewrite", 0, 2 * chout_z, chout_z, 1, 1, 0, idx < c.norm_starts ? 1 : 0, 1, 1, tmp, s);
    auto prep_dconv = [&](const std::string& base, int C) -> int {
      const int hid = C / c.dconv_comp;
      for (int d = 0; d < c.dconv_depth; ++d) {
        const std::string L = base + ".layers." + std::to_string(d);
        int r = prep_conv(h, L + ".0", 0, hid, C, 3, 1, 0, 0, 1, 1, tmp, s);
        if (r) return r;
        const int ci = la ? 5 : 3;
        r = prep_conv(h, L + "." + std::to_string(ci), 0, 2 * C, hid, 1, 1, 0, 0, 1, 1, tmp, s);
        if (r) return r;
        if (la) {
          const std::string lb = L + ".3";
          for (int l = 0; l < 2; ++l) {
            const int in = l == 0 ? hid : 2 * hid;
            Buf& whh = h->whh[lb + ".l" + std::to_string(l)];
            if (whh.alloc((size_t)2 * 4 * hid * hid)) return 1;
            for (int d2 = 0; d2 < 2; ++d2) {
              const std::string sfx = "_l" + std::to_string(l) + (d2 ? "_reverse" : "");
              const float* bi = HP(h, lb + ".lstm.bias_ih" + sfx);
              const float* bh = HP(h, lb + ".lstm.bias_hh" + sfx);
              const float* wh = HP(h, lb + ".lstm.weight_hh" + sfx);
              if (!bi || !bh || !wh) { set_error("hdemucs: missing LSTM parameters under '" + lb + ".lstm'"); return 2; }
              Buf& bsum = h->params[lb + ".lstm.bias_sum" + sfx];
              if (bsum.alloc(4 * hid)) return 1;
              if ((r = launch_add_vec(bi, bh, bsum.p, 4 * hid, s))) return r;
              RFX_CHECK_CUDA(cudaMemcpyAsync(whh.p + (size_t)d2 * 4 * hid * hid, wh, (size_t)4 * hid * hid * 4, cudaMemcpyDeviceToDevice, s));
              const std::string ihname = lb + ".lstm.ih" + std::to_string(l) + (d2 ? "r" : "f");
              r = prep_conv(h, ihname, 0, 4 * hid, in, 1, 1, 0, 0, 1, 1, tmp, s, lb + ".lstm.weight_ih" + sfx, lb + ".lstm.bias_sum" + sfx);
              if (r) return r;
              h->convs[ihname].bkey = lb + ".lstm.bias_ih" + sfx;   // both biases receive the column sums of the gate gradient
              h->convs[ihname].bkey2 = lb + ".lstm.bias_hh" + sfx;
              // W_hh as a 1-tap "conv" too: the backward recomputes every step's gates with one GEMM over the saved h
              r = prep_conv(h, lb + ".lstm.hh" + std::to_string(l) + (d2 ? "r" : "f"), 0, 4 * hid, hid, 1, 1, 0, 0, 1, 1, tmp, s,
                            lb + ".lstm.weight_hh" + sfx, "__none__");
              if (r) return r;
            }
          }
          if ((r = prep_conv(h, lb + ".linear", 0, hid, 2 * hid, 1, 1, 0, 0, 1, 1, tmp, s))) return r;
          const std::string ab = L + ".4";
          if ((r = prep_conv(h, ab + ".query", 0, hid, hid, 1, 1, 0, 0, 1, 1, tmp, s))) return r;
          if ((r = prep_conv(h, ab + ".key", 0, hid, hid, 1, 1, 0, 0, 1, 1, tmp, s))) return r;
          if ((r = prep_conv(h, ab + ".content", 0, hid, hid, 1, 1, 0, 0, 1, 1, tmp, s))) return r;
          if ((r = prep_conv(h, ab + ".query_decay", 0, 16, hid, 1, 1, 0, 0, 1, 1, tmp, s))) return r;
          if ((r = prep_conv(h, ab + ".proj", 0, hid, hid, 1, 1, 0, 0, 1, 1, tmp, s))) return r;
        }
      }
      return 0;
    };
    if (!rc) rc = prep_dconv(fe + ".dconv", chout_z);
    if (freq && !rc) {
      if (idx > 0) rc = prep_conv(h, te + ".conv", 1, chout, chin, c.kernel_size, c.stride, c.kernel_size / 4, 0, 1, 1, tmp, s);
      if (!last_freq && !rc) {
        rc = prep_conv(h, te + ".rewrite", 0, 2 * chout, chout, 1, 1, 0, 1, 1, 1, tmp, s);
        if (!rc) rc = prep_dconv(te + ".dconv", chout);
      }
    }
    // decoder convs (decoder i mirrors encoder depth-1-i)
    const int cin_dec = (idx == 0) ? c.audio_channels * c.n_sources : chin;
    const int cin_dec_z = (idx == 0) ? 2 * c.audio_channels * c.n_sources : chin_z;
    if (!rc) {
      const bool two_d = freq;  // Conv2d rewrite on the freq branch: 3x3 (TA:249); the merged layer uses Conv1d k=3
      rc = prep_conv(h, fd + ".rewrite", 0, 2 * chout_z, chout_z, two_d ? 9 : 3, 1, 1, idx < c.norm_starts ? 1 : 0, two_d ? 3 : 1, two_d ? 3 : 3, tmp, s);
      if (!rc && two_d) { h->convs[fd + ".rewrite"].kh = 3; h->convs[fd + ".rewrite"].kw = 3; }
    }
    if (!rc && idx > 0) rc = prep_conv(h, fd + ".conv_tr", 2, cin_dec_z, chout_z, ker, stri, 0, 0, 1, 1, tmp, s);
    if (!rc && idx == 0) {  // the last freq decoder (C -> 2) is a fused SIMT kernel; it still needs a descriptor for (k, s)
      Conv& cc = h->convs[fd + ".conv_tr"];
      cc.g.kind = 2; cc.g.k = ker; cc.g.s = stri; cc.Ci = chout_z; cc.Co = cin_dec_z;
    }
    if (!rc) h->convs[fd + ".conv_tr"].crop = pad ? (ker - stri) / 2 : 0;
    if (freq && !rc) {
      if (!last_freq) rc = prep_conv(h, td + ".rewrite", 0, 2 * chout, chout, 3, 1, 1, idx < c.norm_starts ? 1 : 0, 1, 1, tmp, s);
      if (!rc && idx > 0) rc = prep_conv(h, td + ".conv_tr", 2, cin_dec, chout, c.kernel_size, c.stride, 0, 0, 1, 1, tmp, s);
      if (!rc) h->convs[td + ".conv_tr"].crop = (c.kernel_size - c.stride) / 2;
      if (!rc && idx == 0) {  // the last time decoder (C -> 1) is a SIMT kernel; it still needs a descriptor for (k, s)
        Conv& cc = h->convs[td + ".conv_tr"];
        cc.g.kind = 2; cc.g.k = c.kernel_size; cc.g.s = c.stride; cc.Ci = chout; cc.Co = cin_dec;
        cc.crop = (c.kernel_size - c.stride) / 2;
      }
    }
    chin = chout; chin_z = chout_z;
    chout *= c.growth; chout_z *= c.growth;
    if (freq) freqs = (freqs <= c.kernel_size) ? 1 : freqs / c.stride;
  }
  if (rc) return rc;
  h->finalized = true;
  return 0;
}

size_t rfx_hdemucs_workspace_bytes(rfx_hdemucs_t* h, int B, int T) {
  if (!h || !h->finalized || B <= 0 || T <= 0) return 0;
  size_t bytes = 0;
  int launches = 0;
  if (run_forward(h, nullptr, B, T, nullptr, nullptr, true, false, nullptr, &bytes, &launches)) return 0;
  return bytes;
}

int rfx_hdemucs_forward(rfx_hdemucs_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(h && x && out && workspace, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_hdemucs_finalize has not been called since the last parameter load");
  // any length the reference's reflect padding accepts without its zero-extension path (TA:465-487: left pad 3/8 nfft, right pad up
  // to 3/8 nfft + hop - 1); lengths off the hop grid run the zero-padded strided convs of the time branch (TA:147-150)
  RFX_REQUIRE(B > 0 && T >= h->cfg.nfft, "T must be at least nfft samples");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  size_t need = 0;
  int launches = 0;
  int rc = run_forward(h, nullptr, B, T, nullptr, nullptr, true, false, nullptr, &need, &launches);
  if (rc) return rc;
  RFX_REQUIRE(workspace_bytes >= need, "workspace too small (rfx_hdemucs_workspace_bytes)");
  return run_forward(h, x, B, T, out, reinterpret_cast<uint8_t*>(workspace), false, false, (cudaStream_t)stream, nullptr, nullptr);
}

int rfx_hdemucs_launches_per_call(rfx_hdemucs_t* h, int B, int T) {
  if (!h || !h->finalized) return 0;
  size_t bytes = 0;
  int launches = 0;
  run_forward(h, nullptr, B, T, nullptr, nullptr, true, false, nullptr, &bytes, &launches);
  return launches;
}

int rfx_hdemucs_set_taps(rfx_hdemucs_t* h, int on) {
  RFX_REQUIRE(h, "null handle");
  h->want_taps = on != 0;
  return 0;
}

/* Copy a named intermediate activation of the last forward (taps enabled) as fp32 (B, Y, X, C); dims -> 4 ints. */
int rfx_hdemucs_tap(rfx_hdemucs_t* h, const char* name, float* dst, int64_t capacity, int* dims, void* stream) {
  RFX_REQUIRE(h && name && dims, "null argument");
  auto it = h->taps.find(name);
  RFX_REQUIRE(it != h->taps.end(), "no such tap (enable taps and run a forward first)");
  const Ten& t = it->second;
  dims[0] = t.B; dims[1] = t.Y; dims[2] = t.X; dims[3] = t.C;
  if (!dst) return 0;
  RFX_REQUIRE((size_t)capacity >= t.elems(), "destination too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (t.f) {
    RFX_CHECK_CUDA(cudaMemcpyAsync(dst, t.f, t.elems() * 4, cudaMemcpyDeviceToDevice, s));
  } else {
    hd_unsplit(t.hi, t.lo(), dst, (long long)t.elems(), s);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
