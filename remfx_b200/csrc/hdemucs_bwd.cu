// Hybrid-Demucs training step, backward half: replays the tape recorded by rfx_hdemucs_forward_train (hdemucs.cu) in reverse.
// Reference: torchaudio.models.HDemucs under torch autograd inside remfx.models.DemucsModel.forward (remfx/models.py:317-321) as
// the Lightning step differentiates it (remfx/models.py:217-220; cfg/exp/5-5_full.yaml:3 is this network).
//
// Per recorded op (hd_internal.h: OpKind):
//   conv        dW = pixel contraction  G^T A   (hd_wgrad_kernel, mma.sync bf16x3, staged then scattered into the parameter layout)
//               db = column sums of G;  dA = gemm2 (tcgen05) on the TRANSPOSED packed weights with negated tap offsets
//   GroupNorm / GELU / GLU / LayerScale / residual / crop   gn_bwd_kernel (two passes under GroupNorm)
//   BLSTM       gates of every step recomputed by one GEMM over the saved h, cell state by an element-wise scan, reverse-time
//               chain with W_hh^T one launch per step (both directions), dW_hh / dW_ih as pixel contractions
//   local attention   local_attn_bwd_kernel (softmax recomputed in shared memory)
//   first / last layers (1-2 channels)   narrow_bwd_kernel;  iSTFT adjoint = an ordinary forward STFT launch of dout / envelope
// Gradient formats: fp32 (B, Y, X, C) for split-bf16 activations, split planes for fp32 pre-activations (they feed the GEMMs).
// No gradient is produced for the audio input (nothing upstream of the network has parameters; the reference never asks).
#include "hd_bwd_kernels.cuh"
#include "bwd_common.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <set>

using namespace rfx;
using namespace rfx::hd;

namespace rfx {
namespace hd {
int hd_gather_weights(rfx_hdemucs* h, const Conv& c, float* dst, cudaStream_t st);  // hdemucs.cu
}
}  // namespace rfx
int rfx_encode_tiled_bf16(void* map, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                          const unsigned* box, int swizzle128);  // gemm2.cu

static int g_hd_wgrad_impl = 0;  // 0 = tcgen05 (MN-major operands), 1 = mma.sync tile variants

// Shapes the fused narrow-layer kernel takes (launch_wgrad_fused) AND whose bias column sums it may produce on the side
static bool wgrad_fused_sums_bias(int N, int K, int taps) {
  static const bool allow = [] { const char* e = getenv("RFX_HD_WGRAD_FUSED"); return !(e && atoi(e) == 0); }();
  static const bool sums = [] { const char* e = getenv("RFX_HD_WGRAD_COLSUM"); return e && atoi(e) != 0; }();   // opt-in until verified on hardware
  const int k_eff = (K + 15) & ~15;
  return allow && sums && N <= 128 && K <= 64 && taps <= 16 && taps * k_eff <= 512 && (2 + taps) * HT_BOX * 2 <= 227 * 1024 - 2048;
}

// The fused narrow-layer form of the tcgen05 contraction (hd_wgrad_tc_fused_kernel): N <= 128, K <= 64, every tap's accumulator in
// tensor memory at once.  Returns -1 when the shape does not qualify (the caller takes the general kernel), else a status.
static int launch_wgrad_fused(const HtMap& mg, const HtMap& ma, HtParams tp, int Bn, cudaStream_t s, float* dB = nullptr) {
  static const bool allow = [] { const char* e = getenv("RFX_HD_WGRAD_FUSED"); return !(e && atoi(e) == 0); }();
  const int k_eff = (tp.K + 15) & ~15;
  if (!allow || tp.N > 128 || tp.K > 64 || tp.taps * k_eff > 512 || tp.taps > 16) return -1;
  const int kcols = tp.taps * 64 <= 512 ? 64 : k_eff;
  int tmem_cols = 32;
  while (tmem_cols < tp.taps * kcols) tmem_cols <<= 1;
  const int stage_bytes = (2 + tp.taps) * HT_BOX;
  const int stages = std::min(4, (227 * 1024 - 2048) / stage_bytes);
  if (stages < 2) return -1;
  // one CTA per SM (tensor memory), a single balanced wave: chunks per item = floor(SMs / items)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int want = std::max(1, sms / std::max(1, Bn));
  int tchunk = std::max(4, ceil_div(tp.ptiles, want));
  tp.tchunk = tchunk;
  tp.nchunks = ceil_div(tp.ptiles, tchunk);
  if (tp.nchunks > 65535 || Bn > 65535) return -1;
  const int smem = stages * stage_bytes + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    RFX_CHECK_CUDA(cudaFuncSetAttribute(hd_wgrad_tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  hd_wgrad_tc_fused_kernel<<<dim3(tp.nchunks, Bn), 192, smem, s>>>(mg, ma, tp, stages, kcols, tmem_cols, dB);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {

struct GradRec {
  float* f = nullptr;             // fp32 gradient of a split activation
  __nv_bfloat16* s = nullptr;     // split planes of the gradient of an fp32 pre-activation (lo plane at s + plane)
  size_t plane = 0;
  bool init = false;
};

struct BRunner {
  rfx_hdemucs* h;
  int B, T;
  uint8_t* ws;
  size_t off;
  bool dry;
  cudaStream_t s;
  const std::map<std::string, float*>& grads;
  int rc = 0;
  std::map<const void*, GradRec> G;
  float* stage = nullptr;   // weight-gradient staging [Nout][taps][Kp] (largest conv)
  float* bstage = nullptr;  // bias staging [largest Nout]
  std::set<std::string> injected;
  // Parameter gradients (weight-gradient contraction, bias column sums, scatters) are off the critical path of the reverse replay --
  // only the INPUT gradient feeds the next op -- so they run on a side stream (the handle's second stream) behind an event that
  // marks "this op's output gradient is complete"; the shared staging buffers are only ever touched on that stream, which keeps
  // their reuse stream-ordered.  The caller's stream joins at the end of the replay.  RFX_HD_OVERLAP=0 keeps one stream.
  cudaStream_t s_side = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_join = nullptr;
  void side_setup() {
    static const bool on = [] { const char* e = getenv("RFX_HD_OVERLAP"); return !(e && atoi(e) == 0); }();
    if (!on || dry) return;
    if (!h->s_time && cudaStreamCreateWithFlags(&h->s_time, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); h->s_time = nullptr; return; }
    for (int i = 0; i < 2; ++i)
      if (!h->ev_branch[i] && cudaEventCreateWithFlags(&h->ev_branch[i], cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); return; }
    s_side = h->s_time; ev_ready = h->ev_branch[0]; ev_join = h->ev_branch[1];
  }
  struct SideScope {  // while alive, the runner's launches go to the side stream
    BRunner& r; cudaStream_t keep;
    explicit SideScope(BRunner& rr) : r(rr), keep(rr.s) {
      if (r.s_side && !r.dry && !r.rc) {
        if (cudaEventRecord(r.ev_ready, r.s) != cudaSuccess || cudaStreamWaitEvent(r.s_side, r.ev_ready, 0) != cudaSuccess) r.fail("side-stream hand-over");
        else r.s = r.s_side;
      }
    }
    ~SideScope() { r.s = keep; }
  };
  void side_join() {
    if (s_side && !dry) {
      if (cudaEventRecord(ev_join, s_side) != cudaSuccess || cudaStreamWaitEvent(s, ev_join, 0) != cudaSuccess) fail("side-stream join");
    }
  }

  void* take(size_t bytes) {
    const size_t r = off;
    off += align_up(bytes, 256);
    return dry ? reinterpret_cast<void*>((uintptr_t)4096 + r) : ws + r;
  }
  bool ok() const { return rc == 0; }
  void chk(const char* what) {
    if (!dry && rc == 0) {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { set_error(std::string("hdemucs backward (") + what + "): " + cudaGetErrorString(e)); rc = 1; }
    }
  }
  void fail(const std::string& m) { if (!rc) { set_error("hdemucs backward: " + m); rc = 2; } }
  float* pgrad(const std::string& key) {
    if (key.empty()) return nullptr;
    auto it = grads.find(key);
    if (it == grads.end()) { if (!dry) fail("no gradient buffer for parameter '" + key + "'"); return nullptr; }
    return it->second;
  }
  // fp32 gradient buffer of an activation tensor (allocated on first use; `init` says whether it holds a value yet)
  GradRec& actgrad(const Ten& t) {
    GradRec& r = G[t.key()];
    if (!r.f) r.f = reinterpret_cast<float*>(take(t.elems() * 4));
    return r;
  }
  // split planes for the gradient of an fp32 pre-activation tensor (B, Y, X, C) -> (B, Y, X, ceil8(C))
  GradRec& rawgrad(const Ten& t) {
    GradRec& r = G[t.key()];
    if (!r.s) {
      const size_t n = (size_t)t.B * t.Y * t.X * (ceil_div(t.C, 8) * 8);
      r.plane = align_up(n * 2, 256) / 2;
      r.s = reinterpret_cast<__nv_bfloat16*>(take(r.plane * 2 * 2));
    }
    return r;
  }
  // dst (+)= src honoring the first-writer rule
  void add_or_copy(GradRec& r, const float* src, size_t n) {
    if (dry || rc) { r.init = true; return; }
    if (!r.init) {
      if (cudaMemcpyAsync(r.f, src, n * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) fail("memcpy");
    } else {
      add_inplace_kernel<<<(unsigned)std::min<size_t>((n / 4 + 255) / 256, 148 * 8), 256, 0, s>>>(r.f, src, (long long)(n / 4));
      chk("add");
    }
    r.init = true;
  }
  int rows_for(long long npix, int nseg, int min_rows) {
    const long long want = std::max<long long>(1, 592 / std::max(1, nseg));
    long long rows = (npix + want - 1) / want;
    if (rows < min_rows) rows = min_rows;
    return (int)std::min<long long>(rows, 1 << 30);
  }

  // ---- weight + bias gradient of one prepared conv from the planes of its output gradient ----
  void conv_param_grads(const Conv& c, const __nv_bfloat16* g, size_t g_plane, long long g_ld, int gcol0, int Bn, int Y, int X,
                        const SplitAct& A, const int* dx, const int* dy, int Ktap) {
    const GatherSpec& gs = c.g;
    float* dw = pgrad(c.wkey);
    float* db = pgrad(c.bkey);
    float* db2 = pgrad(c.bkey2);
    if (dry || rc) return;
    SideScope side(*this);
    // narrow layers: the fused weight-gradient kernel sums the bias columns from the boxes it has in shared memory anyway
    const bool bias_in_wgrad = db && dw && g_hd_wgrad_impl == 0 && (g_ld % 8) == 0 && (A.ld % 8) == 0 && (gcol0 % 8) == 0 &&
                               wgrad_fused_sums_bias(gs.Nout, Ktap, gs.taps);
    if (db && bias_in_wgrad) {
      if (cudaMemsetAsync(bstage, 0, (size_t)gs.Nout * 4, s) != cudaSuccess) { fail("memset"); return; }
    } else if (db) {
      if (cudaMemsetAsync(bstage, 0, (size_t)gs.Nout * 4, s) != cudaSuccess) { fail("memset"); return; }
      const long long rows = (long long)Bn * Y * X;
      const int rpc = rows_for(rows, 1, 64);
      colsum_split_kernel<<<dim3((unsigned)((rows + rpc - 1) / rpc), ceil_div(gs.Nout, 2048)), 256, 0, s>>>(g, g + g_plane, rows, (int)g_ld, gcol0,
                                                                                                             gs.Nout, bstage, rpc);
      chk("colsum");
      scatter_bias_kernel<<<ceil_div(gs.Nout, 256), 256, 0, s>>>(bstage, gs, db, db2);
      chk("scatter_bias");
    }
    if (dw) {
      const size_t sn = (size_t)gs.Nout * gs.taps * gs.Kp;
      if (cudaMemsetAsync(stage, 0, sn * 4, s) != cudaSuccess) { fail("memset"); return; }
      WgP p{};
      p.g = g; p.g_bs = (long long)Y * X * g_ld; p.g_ldy = (long long)X * g_ld; p.g_ld = g_ld; p.g_plane = (long long)g_plane; p.gcol0 = gcol0;
      p.a = A.hi; p.a_bs = A.batch_stride; p.a_ldy = A.ld_y; p.a_ld = A.ld; p.a_plane = A.plane_stride;
      p.Y = Y; p.X = X; p.Ay = (int)(A.rows_y > 0 ? A.rows_y : 1); p.Ax = (int)A.rows;
      p.N = gs.Nout; p.K = Ktap; p.taps = gs.taps; p.Kp = gs.Kp;
      for (int t = 0; t < gs.taps; ++t) { p.dx[t] = dx[t]; p.dy[t] = dy[t]; }
      if (g_hd_wgrad_impl == 0 && (g_ld % 8) == 0 && (A.ld % 8) == 0 && (gcol0 % 8) == 0) {
        // ---- tcgen05 path: 5-D tensor maps {channel, x, y, item, plane}; taps are shifted box origins, padding is OOB fill ----
        const int Ay = (int)(A.rows_y > 0 ? A.rows_y : 1), Ax = (int)A.rows;
        HtMap mg, ma;
        int bx = 32;
        while (bx > 1 && bx > X) bx >>= 1;           // largest power of two <= X, at most 32
        const int by = 32 / bx;
        const unsigned box[5] = {64, (unsigned)bx, (unsigned)by, 1, 2};
        const unsigned long long gd[5] = {(unsigned long long)gs.Nout, (unsigned long long)X, (unsigned long long)Y, (unsigned long long)Bn, 2};
        const unsigned long long gst[4] = {(unsigned long long)g_ld * 2, (unsigned long long)X * g_ld * 2, (unsigned long long)Y * X * g_ld * 2,
                                           (unsigned long long)g_plane * 2};
        const unsigned long long ad[5] = {(unsigned long long)Ktap, (unsigned long long)Ax, (unsigned long long)Ay, (unsigned long long)Bn, 2};
        const unsigned long long ast[4] = {(unsigned long long)A.ld * 2, (unsigned long long)(A.ld_y > 0 ? A.ld_y : (long long)Ax * A.ld) * 2,
                                           (unsigned long long)A.batch_stride * 2, (unsigned long long)A.plane_stride * 2};
        if (rfx_encode_tiled_bf16(&mg, g + gcol0, 5, gd, gst, box, 1) || rfx_encode_tiled_bf16(&ma, A.hi, 5, ad, ast, box, 1)) { rc = 1; return; }
        HtParams tp{};
        tp.taps = gs.taps; tp.N = gs.Nout; tp.K = Ktap; tp.Kp = gs.Kp;
        for (int t = 0; t < gs.taps; ++t) { tp.dx[t] = dx[t]; tp.dy[t] = dy[t]; }
        tp.bx = bx; tp.by = by;
        tp.tiles_x = ceil_div(X, bx);
        tp.ptiles = tp.tiles_x * ceil_div(Y, by);
        tp.ntn = ceil_div(gs.Nout, 256); tp.ntk = ceil_div(Ktap, 256);
        tp.dW = stage;
        {
          const int fr = launch_wgrad_fused(mg, ma, tp, Bn, s, bias_in_wgrad ? bstage : nullptr);
          if (fr > 0) { rc = fr; return; }
          if (fr == 0) {
            chk("wgrad tcgen05 (fused taps)");
            scatter_w_kernel<<<148 * 4, 256, 0, s>>>(stage, gs, dw);
            chk("scatter_w");
            if (bias_in_wgrad) {
              scatter_bias_kernel<<<ceil_div(gs.Nout, 256), 256, 0, s>>>(bstage, gs, db, db2);
              chk("scatter_bias");
            }
            return;
          }
          if (bias_in_wgrad) { fail("internal: fused weight-gradient kernel refused a shape it was expected to take"); return; }
        }
        const long long per = (long long)gs.taps * tp.ntn * tp.ntk * Bn;
        long long want = (148ll * 2 + per - 1) / per;   // about two waves of CTAs
        if (want < 1) want = 1;
        long long tchunk = (tp.ptiles + want - 1) / want;
        if (tchunk < 4) tchunk = 4;
        tp.tchunk = (int)tchunk;
        tp.nchunks = ceil_div(tp.ptiles, tp.tchunk);
        tp.dW = stage;
        if (tp.nchunks > 65535 || Bn > 65535) { fail("weight-gradient grid too large"); return; }
        hd_wgrad_tc_kernel<<<dim3(gs.taps * tp.ntn * tp.ntk, tp.nchunks, Bn), 192, HT_SMEM, s>>>(mg, ma, tp);
        chk("wgrad tcgen05");
        scatter_w_kernel<<<148 * 4, 256, 0, s>>>(stage, gs, dw);
        chk("scatter_w");
        return;
      }
      // tile variant with the least padded area (ties go to the larger tile)
      static const int cfgs[6][2] = {{128, 128}, {64, 128}, {128, 64}, {64, 64}, {32, 128}, {128, 32}};
      int best = 0;
      long long best_cost = -1;
      for (int i = 0; i < 6; ++i) {
        const long long cost = (long long)ceil_div(p.N, cfgs[i][0]) * cfgs[i][0] * ceil_div(p.K, cfgs[i][1]) * cfgs[i][1];
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = i; }
      }
      const int BMt = cfgs[best][0], BNt = cfgs[best][1];
      const int tiles = ceil_div(p.N, BMt) * ceil_div(p.K, BNt) * gs.taps;
      const long long R = (long long)Y * X;
      long long want = (148ll * 6 + (long long)tiles * Bn - 1) / ((long long)tiles * Bn);
      if (want < 1) want = 1;
      long long rchunk = (R + want - 1) / want;
      rchunk = (rchunk + HW_BK - 1) / HW_BK * HW_BK;
      if (rchunk < 8 * HW_BK) rchunk = 8 * HW_BK;
      p.rchunk = (int)rchunk;
      p.nchunks = (int)((R + rchunk - 1) / rchunk);
      p.dW = stage;
      if ((long long)Bn * p.nchunks > 65535) { fail("weight-gradient grid too large"); return; }
      const dim3 grid(tiles, (unsigned)(Bn * p.nchunks));
      switch (best) {
        case 0: hd_wgrad_kernel<2, 4, 4, 4><<<grid, 256, HwCfg<2, 4, 4, 4>::SMEM, s>>>(p); break;
        case 1: hd_wgrad_kernel<2, 4, 2, 4><<<grid, 256, HwCfg<2, 4, 2, 4>::SMEM, s>>>(p); break;
        case 2: hd_wgrad_kernel<2, 4, 4, 2><<<grid, 256, HwCfg<2, 4, 4, 2>::SMEM, s>>>(p); break;
        case 3: hd_wgrad_kernel<2, 4, 2, 2><<<grid, 256, HwCfg<2, 4, 2, 2>::SMEM, s>>>(p); break;
        case 4: hd_wgrad_kernel<1, 8, 2, 2><<<grid, 256, HwCfg<1, 8, 2, 2>::SMEM, s>>>(p); break;
        default: hd_wgrad_kernel<8, 1, 1, 4><<<grid, 256, HwCfg<8, 1, 1, 4>::SMEM, s>>>(p); break;
      }
      chk("wgrad");
      scatter_w_kernel<<<148 * 4, 256, 0, s>>>(stage, gs, dw);
      chk("scatter_w");
    }
  }

  int ensure_transposed(Conv& c, int Kt, cudaStream_t s) {   // (s shadows the runner's stream: the pack is built on the stream given)
    if (c.wt_ready) return 0;
    const GatherSpec& gs = c.g;
    const int Np = ceil_div(gs.Nout, 64) * 64;
    const size_t wn = (size_t)gs.Nout * gs.taps * gs.Kp, tn = (size_t)Kt * gs.taps * Np;
    if (h->gather_tmp.n < wn && h->gather_tmp.alloc(wn)) return 1;
    if (h->gather_tmp2.n < tn && h->gather_tmp2.alloc(tn)) return 1;
    int r = hd_gather_weights(h, c, h->gather_tmp.p, s);
    if (r) return r;
    transpose_w_kernel<<<148 * 4, 256, 0, s>>>(h->gather_tmp.p, gs.Nout, gs.taps, gs.Kp, Kt, Np, h->gather_tmp2.p);
    RFX_CHECK_CUDA(cudaGetLastError());
    const int BN = g2_choose_bn(Kt);
    if (c.wtbuf.alloc(split_weight_elems(Kt, gs.taps * Np, BN))) return 1;
    r = pack_split_weights(h->gather_tmp2.p, (long long)gs.taps * Np, Kt, gs.taps * Np, BN, reinterpret_cast<__nv_bfloat16*>(c.wtbuf.p), &c.wt, s);
    if (r) return r;
    c.wt_ready = true;
    return 0;
  }

  // ---- OP_CONV ----
  void conv_bwd(const Op& op) {
    auto it = h->convs.find(op.name);
    if (it == h->convs.end()) { fail("conv '" + op.name + "' not prepared"); return; }
    Conv& c = it->second;
    const G2Problem& pr = op.pr;
    auto git = G.find(op.out.key());
    if (git == G.end() || !git->second.s) { fail("no output gradient reached conv '" + op.name + "'"); return; }
    const GradRec gr = git->second;
    const int Yo = pr.My > 0 ? pr.My : 1, Xo = pr.M;
    const long long ldg = ceil_div(op.out.C, 8) * 8;
    conv_param_grads(c, gr.s, gr.plane, ldg, op.dst_col, pr.batch, Yo, Xo, pr.A, pr.row_off, pr.row_off_y, pr.Ktap);
    // input gradient
    GradRec& gi = actgrad(op.in);
    if (dry || rc) { gi.init = true; return; }
    if (c.pack_ev >= 0) {   // built ahead on the preparation stream: this stream continues behind its event
      if (cudaStreamWaitEvent(s, h->ev_pack[c.pack_ev], 0) != cudaSuccess) { fail("pack event"); return; }
      c.pack_ev = -1;
    }
    if ((rc = ensure_transposed(c, pr.Ktap, s))) return;
    G2Problem q;
    q.A.hi = gr.s + op.dst_col; q.A.rows = Xo; q.A.rows_y = Yo; q.A.ld = ldg; q.A.ld_y = (long long)Xo * ldg;
    q.A.batch_stride = (long long)Yo * Xo * ldg; q.A.plane_stride = (long long)gr.plane;
    q.W = c.wt;
    const int Xv = (int)pr.A.rows, Yv = (int)(pr.A.rows_y > 0 ? pr.A.rows_y : 1), Cv = pr.Ktap;
    q.M = Xv; q.My = Yv; q.N = Cv; q.batch = pr.batch; q.Ktap = c.g.Nout; q.taps = c.g.taps;
    for (int t = 0; t < c.g.taps; ++t) { q.row_off[t] = -pr.row_off[t]; q.row_off_y[t] = -pr.row_off_y[t]; }
    int xt = 128;
    while (xt > Xv && xt > 1) xt >>= 1;
    q.xt = xt;
    // residual branches: the second and later contributions accumulate in the GEMM's epilogue (Cf += ...)
    q.Cf = gi.f; q.cf_accum = gi.init; q.ldcf = Cv; q.ldcf_y = (long long)Xv * Cv; q.bscf = (long long)Yv * Xv * Cv;
    if ((rc = launch_gemm2(q, s))) return;
    gi.init = true;
  }

  // ---- OP_GN ----
  void gn_bwd(const Op& op) {
    auto git = G.find(op.out.key());
    if (git == G.end() || !git->second.f || !git->second.init) { fail("no gradient reached an activation (gn_apply output)"); return; }
    const float* dy = git->second.f;
    GradRec& graw = rawgrad(op.in);
    GradRec* gres = op.in2.key() ? &actgrad(op.in2) : nullptr;
    const GnApply& a = op.gn;
    const bool has_stats = a.stats != nullptr;
    const int nseg = a.per_x ? op.in.B * a.Xr : op.in.B;
    double* gsum = has_stats ? reinterpret_cast<double*>(take((size_t)nseg * a.G * 2 * 8)) : nullptr;
    float* dgamma = has_stats ? pgrad(op.p_gamma) : nullptr;
    float* dbeta = has_stats ? pgrad(op.p_beta) : nullptr;
    float* dscale = a.scale ? pgrad(op.p_scale) : nullptr;
    if (dry || rc) { graw.init = true; if (gres) gres->init = true; return; }
    if (a.Co % 8 || a.Co / 8 > 256) { fail("gn backward: unsupported channel count"); return; }
    if (has_stats && a.G > 1 && (a.Cr / a.G) % 8) { fail("gn backward: groups must hold whole channel octets"); return; }
    if (a.scale && !has_stats) { fail("gn backward: LayerScale without GroupNorm is not on the path"); return; }
    GnBwd p{};
    p.a = a; p.dy = dy;
    p.dres = gres ? gres->f : nullptr; p.res_accum = gres && gres->init ? 1 : 0;
    p.ghi = graw.s; p.glo = graw.s + graw.plane; p.Cg = ceil_div(a.Cr, 8) * 8;
    p.gsum = gsum; p.dgamma = dgamma; p.dbeta = dbeta; p.dscale = dscale;
    const long long npix = a.per_x ? a.Y : (long long)a.Y * a.Xr;
    p.count = npix * (a.Cr / a.G);
    const int groups = a.Co / 8, rows = 256 / groups;
    if (a.mode < 0 || a.mode > 3) { fail("gn backward: unknown activation mode"); return; }
    // The instantiations hold two or three CTAs per SM (registers).  The work is cut into ~888 = 148 x 6 items so that a grid of
    // 148 x (CTAs per SM) is ONE wave whose CTAs each stride over the same number of items, whichever the occupancy (a fixed grid of
    // 592 ran as 1.33 waves at three CTAs per SM).
    using GnFn = void (*)(GnBwd);
    static const GnFn fn1[4] = {gn_bwd_kernel<1, 0>, gn_bwd_kernel<1, 1>, gn_bwd_kernel<1, 2>, gn_bwd_kernel<1, 3>};
    static const GnFn fn2[4] = {gn_bwd_kernel<2, 0>, gn_bwd_kernel<2, 1>, gn_bwd_kernel<2, 2>, gn_bwd_kernel<2, 3>};
    const size_t smem1 = (size_t)((2 * a.Cr + a.Co + 3) & ~3) * 4 + 2 * a.G * 8;
    auto ctas_per_sm = [&](GnFn f, size_t smem) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, f, 256, smem) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = 1; }
      return std::min(n, 6);
    };
    {
      const long long want = std::max<long long>(1, 888 / std::max(1, nseg));
      p.rows_per_cta = (int)std::min<long long>(std::max<long long>((npix + want - 1) / want, rows * 8), 1 << 30);
    }
    p.nseg = nseg;
    p.nchunks = (int)((npix + p.rows_per_cta - 1) / p.rows_per_cta);
    const long long items = (long long)nseg * p.nchunks;
    if (has_stats) {
      if (cudaMemsetAsync(gsum, 0, (size_t)nseg * a.G * 2 * 8, s) != cudaSuccess) { fail("memset"); return; }
      const unsigned grid1 = (unsigned)std::min<long long>(items, 148 * ctas_per_sm(fn1[a.mode], smem1));
      fn1[a.mode]<<<grid1, 256, smem1, s>>>(p);
      chk("gn pass 1");
    }
    const unsigned grid2 = (unsigned)std::min<long long>(items, 148 * ctas_per_sm(fn2[a.mode], 0));
    fn2[a.mode]<<<grid2, 256, 0, s>>>(p);
    chk("gn pass 2");
    graw.init = true;
    if (gres) gres->init = true;
  }

  // ---- OP_LSTM ----
  void lstm_bwd(const Op& op) {
    const Ten& cur = op.in; const Ten& Gx = op.aux; const Ten& hout = op.out;
    const int Bs = hout.B, Tf = hout.X, H = hout.C / 2, l = op.i0;
    auto git = G.find(hout.key());
    if (git == G.end() || !git->second.init) { fail("no gradient reached an LSTM output"); return; }
    const float* dH = git->second.f;
    const size_t gate_elems = (size_t)Bs * Tf * 8 * H;
    float* R = reinterpret_cast<float*>(take(gate_elems * 4));
    float* cs = reinterpret_cast<float*>(take((size_t)Bs * Tf * 2 * H * 4));
    float* dG = reinterpret_cast<float*>(take(gate_elems * 4));
    float* carry = reinterpret_cast<float*>(take((size_t)Bs * 2 * H * 4));
    unsigned* bar = reinterpret_cast<unsigned*>(take((size_t)2 * ceil_div(Bs, LBP_B) * sizeof(unsigned)));
    GradRec& gg = rawgrad(Gx);
    const std::string nm[2] = {op.name + ".lstm.hh" + std::to_string(l) + "f", op.name + ".lstm.hh" + std::to_string(l) + "r"};
    for (int d = 0; d < 2 && ok(); ++d) {
      if (!h->convs.count(nm[d])) { fail("LSTM recurrent weights '" + nm[d] + "' not prepared"); return; }
      pgrad(h->convs[nm[d]].wkey);
    }
    if (dry || rc) { gg.init = true; return; }
    // 1. R = W_hh h_prev for every step (one GEMM per direction over the saved h, shifted by one step)
    for (int d = 0; d < 2; ++d) {
      const Conv& c = h->convs[nm[d]];
      G2Problem q;
      q.A.hi = hout.hi + d * H; q.A.rows = Tf; q.A.rows_y = 1; q.A.ld = 2 * H; q.A.ld_y = (long long)Tf * 2 * H;
      q.A.batch_stride = (long long)Tf * 2 * H; q.A.plane_stride = (long long)hout.plane;
      q.W = c.w;
      q.M = Tf; q.My = 1; q.N = 4 * H; q.batch = Bs; q.Ktap = H; q.taps = 1;
      q.row_off[0] = d ? 1 : -1;
      int xt = 128;
      while (xt > Tf && xt > 1) xt >>= 1;
      q.xt = xt;
      q.Cf = R + (size_t)d * 4 * H; q.ldcf = 8 * H; q.ldcf_y = (long long)Tf * 8 * H; q.bscf = (long long)Tf * 8 * H;
      if ((rc = launch_gemm2(q, s))) return;
    }
    // 2. cell states
    lstm_cscan_kernel<<<ceil_div(Bs * 2 * H, 256), 256, 0, s>>>(Gx.f, R, Bs, Tf, H, cs);
    chk("lstm cscan");
    // 3. reverse-time chain: one persistent cooperative launch (W_hh slices resident in shared memory, per-step exchange through L2);
    //    falls back to one launch per step when the grid cannot be co-resident (very large batches) or RFX_HD_LSTM_BWD_STEPWISE is set
    const float* whh = h->whh[op.name + ".l" + std::to_string(l)].p;
    bool persistent = false;
    static const bool mma_on = [] { const char* e = getenv("RFX_LSTM_BWD_MMA"); return !(e && atoi(e) == 0); }();
    if (mma_on && lstm_bwd_chain_mma_supported(H)) {  // cluster / tensor-core chain (lstm.cu)
      if ((rc = launch_lstm_bwd_chain_mma(Gx.f, R, cs, dH, whh, dG, Bs, Tf, H, s))) return;
      persistent = true;
    } else {
      static const bool stepwise = [] { const char* e = getenv("RFX_HD_LSTM_BWD_STEPWISE"); return e && atoi(e) != 0; }();
      const dim3 pg(H / LBP_U, 2, ceil_div(Bs, LBP_B));
      const size_t psmem = ((size_t)4 * H * LBP_U + (size_t)LBP_B * (4 * H + 4) + (size_t)4 * LBP_B * LBP_U) * 4;
      int sms = 148, dev = 0, per_sm = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (!stepwise && H % LBP_U == 0 && psmem <= 227 * 1024 &&
          cudaFuncSetAttribute(lstm_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_bwd_persist_kernel, 256, psmem) == cudaSuccess &&
          (long long)pg.x * pg.y * pg.z <= (long long)per_sm * sms) {
        if (cudaMemsetAsync(bar, 0, (size_t)2 * pg.z * sizeof(unsigned), s) != cudaSuccess) { fail("memset"); return; }
        const float* a0 = Gx.f; const float* a1 = R; const float* a2 = cs; const float* a3 = dH; const float* a4 = whh;
        float* a5 = dG; int a6 = Bs, a7 = Tf, a8 = H; unsigned* a9 = bar;
        void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9};
        const cudaError_t e = cudaLaunchCooperativeKernel((const void*)lstm_bwd_persist_kernel, pg, dim3(256), args, psmem, s);
        if (e == cudaSuccess) persistent = true;
        else (void)cudaGetLastError();   // not launchable as a cooperative grid here: take the step-wise path
      }
    }
    if (!persistent) {
      const size_t smem = (size_t)LB_NB * 4 * H * 4;
      if (cudaFuncSetAttribute(lstm_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { fail("smem attribute"); return; }
      dim3 grid(ceil_div(H, 128), 2, ceil_div(Bs, LB_NB));
      for (int k = 0; k < Tf && ok(); ++k) {
        lstm_bwd_step_kernel<<<grid, 128, smem, s>>>(Gx.f, R, cs, dH, whh, dG, carry, Bs, Tf, H, k);
        if (k == 0 || k == Tf - 1) chk("lstm step");
      }
    }
    // 4. gate gradient as split planes: the A operand of the W_ih input-gradient GEMM and of every weight contraction
    split_pad_kernel<<<148 * 4, 256, 0, s>>>(dG, (long long)Bs * Tf, 8 * H, 8 * H, gg.s, gg.s + gg.plane);
    chk("lstm split");
    gg.init = true;
    // 5. dW_hh = sum_t dG[t] (x) h_prev[t]
    for (int d = 0; d < 2 && ok(); ++d) {
      const Conv& c = h->convs[nm[d]];
      SplitAct A;
      A.hi = hout.hi + d * H; A.rows = Tf; A.rows_y = 1; A.ld = 2 * H; A.ld_y = (long long)Tf * 2 * H;
      A.batch_stride = (long long)Tf * 2 * H; A.plane_stride = (long long)hout.plane;
      const int dx[1] = {d ? 1 : -1}, dy0[1] = {0};
      conv_param_grads(c, gg.s, gg.plane, 8 * H, d * 4 * H, Bs, 1, Tf, A, dx, dy0, H);
    }
    (void)cur;
  }

  // ---- OP_ATTN ----
  void attn_bwd(const Op& op) {
    const Ten& qkv = op.in; const Ten& res = op.out;
    const int Bn = res.B, Tn = res.X, C = res.C, heads = op.i0, nd = op.i1, ld = qkv.C;
    auto git = G.find(res.key());
    if (git == G.end() || !git->second.init) { fail("no gradient reached the attention output"); return; }
    float* dq = reinterpret_cast<float*>(take((size_t)Bn * Tn * ld * 4));
    GradRec& gq = rawgrad(qkv);
    if (dry || rc) { gq.init = true; return; }
    if (ld % 8) { fail("attention: projection buffer width must be a multiple of 8"); return; }
    if (cudaMemsetAsync(dq, 0, (size_t)Bn * Tn * ld * 4, s) != cudaSuccess) { fail("memset"); return; }
    const int Ch = C / heads;
    const size_t smem = ((size_t)2 * Tn * (Ch + 1) + (size_t)2 * Tn * (LAB_QT + 1) + (size_t)2 * LAB_QT * (Ch + 1) + 3 * LAB_QT) * 4;
    if (smem > 226 * 1024) { fail("LocalState sequence too long for the shared-memory attention backward"); return; }
    cudaFuncSetAttribute(local_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    local_attn_bwd_kernel<<<dim3(ceil_div(Tn, LAB_QT), heads, Bn), 256, smem, s>>>(qkv.f, ld, Tn, C, heads, nd, git->second.f, dq);
    chk("attention backward");
    split_pad_kernel<<<148 * 2, 256, 0, s>>>(dq, (long long)Bn * Tn, ld, ld, gq.s, gq.s + gq.plane);
    chk("attention split");
    gq.init = true;
  }

  // ---- first / last (narrow) layers ----
  template <int P, bool ENC>
  void narrow(NarrowP p, int lines) {
    constexpr int CH = (P == 2 || ENC) ? 4 : 8;   // at most 64 weight-gradient sums per thread (no spills at two CTAs per SM)
    const int WN = p.C * P * 8;
    const size_t smem = (size_t)(2 * WN + p.C) * 4;
    const int groups = p.C / CH, rows = 256 / groups;
    if (p.C % 8 || groups > 256) { fail("narrow layer: channels must be a multiple of 8 (at most 1024)"); return; }
    // chunks of wide rows: enough items to balance 296 CTAs, at least four rows per thread
    const long long want = std::max<long long>(1, (8 * 296 + lines - 1) / lines);
    p.rows_per_cta = (int)std::max<long long>((p.Xw + want - 1) / want, rows * 4);
    p.nlines = lines;
    p.nchunks = ceil_div(p.Xw, p.rows_per_cta);
    const int grid = (int)std::min<long long>((long long)lines * p.nchunks, 296);
    if (smem > 48 * 1024) cudaFuncSetAttribute(narrow_bwd_kernel<8, P, ENC, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    narrow_bwd_kernel<8, P, ENC, CH><<<grid, 256, smem, s>>>(p);
    chk("narrow layer");
    if (!ENC) {
      narrow_bias_kernel<P><<<dim3(std::min(32, ceil_div(p.Xn, 256)), lines), 256, 0, s>>>(p);
      chk("narrow bias");
    }
  }

  // The transposed packs depend on the weights only.  Built inside conv_bwd they sat on the critical path of the replay (three small
  // launches per conv, 118 convs); here they are all queued on a third stream in the order the replay needs them, one event per
  // pack, and conv_bwd waits on its conv's event.  The staging buffers (gather_tmp / gather_tmp2) are only touched on that stream
  // during a backward; the caller's stream joins it at the end so that the next finalize finds them free.
  bool prep_joined = true;
  void prepare_packs() {
    if (!s_side || dry || rc) return;   // RFX_HD_OVERLAP=0: packs are built lazily on the caller's stream
    if (!h->s_prep && cudaStreamCreateWithFlags(&h->s_prep, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); h->s_prep = nullptr; return; }
    if (cudaEventRecord(ev_join, s) != cudaSuccess || cudaStreamWaitEvent(h->s_prep, ev_join, 0) != cudaSuccess) { fail("preparation-stream fork"); return; }
    prep_joined = false;
    size_t n_ev = 0;
    for (auto op = h->tape.rbegin(); op != h->tape.rend() && !rc; ++op) {
      if (op->kind != OP_CONV) continue;
      auto it = h->convs.find(op->name);
      if (it == h->convs.end() || it->second.wt_ready) continue;
      Conv& c = it->second;
      if ((rc = ensure_transposed(c, op->pr.Ktap, h->s_prep))) return;
      if (n_ev == h->ev_pack.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { fail("event"); return; }
        h->ev_pack.push_back(e);
      }
      if (cudaEventRecord(h->ev_pack[n_ev], h->s_prep) != cudaSuccess) { fail("pack event record"); return; }
      c.pack_ev = (int)n_ev++;
    }
  }
  void prep_join() {
    if (prep_joined || dry) return;
    prep_joined = true;
    for (auto& kv : h->convs) kv.second.pack_ev = -1;
    if (cudaEventRecord(ev_join, h->s_prep) != cudaSuccess || cudaStreamWaitEvent(s, ev_join, 0) != cudaSuccess) fail("preparation-stream join");
  }

  void run(const float* x, const float* dout) {
    (void)x;
    // scratch sized from the tape
    size_t max_stage = 1, max_n = 1;
    for (const Op& op : h->tape) {
      if (op.kind != OP_CONV) continue;
      auto it = h->convs.find(op.name);
      if (it == h->convs.end()) { fail("conv '" + op.name + "' not prepared"); return; }
      const GatherSpec& g = it->second.g;
      max_stage = std::max(max_stage, (size_t)g.Nout * g.taps * g.Kp);
      max_n = std::max(max_n, (size_t)g.Nout);
    }
    for (auto& kv : h->convs) {  // LSTM recurrent weights are not on the tape as convs
      const GatherSpec& g = kv.second.g;
      if (kv.first.find(".lstm.hh") != std::string::npos) {
        max_stage = std::max(max_stage, (size_t)g.Nout * g.taps * g.Kp);
        max_n = std::max(max_n, (size_t)g.Nout);
      }
    }
    stage = reinterpret_cast<float*>(take(max_stage * 4));
    bstage = reinterpret_cast<float*>(take(max_n * 4));
    side_setup();
    if (!dry) {
      bool attr_ok = true;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_kernel<2, 4, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, HwCfg<2, 4, 4, 4>::SMEM) == cudaSuccess;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_kernel<2, 4, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, HwCfg<2, 4, 2, 4>::SMEM) == cudaSuccess;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_kernel<2, 4, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, HwCfg<2, 4, 4, 2>::SMEM) == cudaSuccess;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_kernel<2, 4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, HwCfg<2, 4, 2, 2>::SMEM) == cudaSuccess;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_kernel<1, 8, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, HwCfg<1, 8, 2, 2>::SMEM) == cudaSuccess;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_kernel<8, 1, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, HwCfg<8, 1, 1, 4>::SMEM) == cudaSuccess;
      attr_ok &= cudaFuncSetAttribute(hd_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM) == cudaSuccess;
      if (!attr_ok) { fail("smem attribute"); return; }
      for (auto& kv : grads) {  // every parameter gradient starts from zero (kernels accumulate or overwrite)
        auto pit = h->params.find(kv.first);
        if (pit == h->params.end()) { fail("gradient buffer for unknown parameter '" + kv.first + "'"); return; }
        if (!h->grads_prezeroed && cudaMemsetAsync(kv.second, 0, pit->second.n * 4, s) != cudaSuccess) { fail("memset"); return; }
      }
    }
    prepare_packs();
    if (rc) return;
    // debug: map injected taps to tensor keys
    std::map<const void*, std::string> inject_at;
    if (!dry)
      for (auto& kv : h->inject) {
        auto t = h->taps.find(kv.first);
        if (t != h->taps.end()) inject_at[t->second.key()] = kv.first;
      }

    const int nfft = h->cfg.nfft, hl = nfft / 4;
    for (int i = (int)h->tape.size() - 1; i >= 0 && ok(); --i) {
      const Op& op = h->tape[i];
      if (!dry && !inject_at.empty() && op.kind != OP_FREQEMB) {
        auto ia = inject_at.find(op.out.key());
        if (ia != inject_at.end() && !injected.count(ia->second)) {
          GradRec& r = actgrad(op.out);
          if (cudaMemcpyAsync(r.f, h->inject[ia->second], op.out.elems() * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) { fail("inject memcpy"); return; }
          r.init = true;
          injected.insert(ia->second);
        }
      }
      switch (op.kind) {
        case OP_FINALTIME: {
          GradRec& gy = actgrad(op.in);
          float* dw = pgrad(op.name + ".weight"); float* db = pgrad(op.name + ".bias");
          if (dry || rc) { gy.init = true; break; }
          if (gy.init) { fail("final time layer: input gradient already written"); break; }
          NarrowP p{};
          p.narrow = dout; p.n_line = T; p.Xn = T; p.stats = op.fp1; p.norm_mode = 2; p.ck = 0; p.Y = 1;
          p.Xw = op.in.X; p.C = op.in.C; p.S = op.i2; p.pad = op.i0;
          p.w = HP(h, op.name + ".weight"); p.bias = nullptr;
          p.yhi = op.in.hi; p.ylo = op.in.lo(); p.dy = gy.f; p.dW = dw; p.db = db;
          if (op.i1 != 8) { fail("final time layer: kernel size 8 only"); break; }
          narrow<1, false>(p, B);
          gy.init = true;
          break;
        }
        case OP_FINALFREQ: {
          const int pad = op.i0, bins = op.i1, le = op.i2;
          const int P0 = hl / 2 * 3, Ltot = T + 2 * P0;
          float* ghat = reinterpret_cast<float*>(take((size_t)B * Ltot * 4));
          float2* gZ = reinterpret_cast<float2*>(take((size_t)B * le * bins * 8));
          GradRec& gy = actgrad(op.in);
          float* dw = pgrad(op.name + ".weight"); float* db = pgrad(op.name + ".bias");
          if (dry || rc) { gy.init = true; break; }
          if (gy.init) { fail("final freq layer: input gradient already written"); break; }
          istft_adj_prep_kernel<<<dim3(ceil_div(Ltot, 256), B), 256, 0, s>>>(dout, T, HP(h, "__window__"), nfft, hl, P0, le, 2, P0, Ltot, ghat);
          chk("istft adjoint prep");
          StftParams sp{};
          sp.x = ghat; sp.x_bstride = Ltot; sp.T = Ltot; sp.x_aligned8 = 0;
          sp.window = HP(h, "__window__"); sp.tw = twiddles(nfft);
          sp.n_fft = nfft; sp.hop = hl; sp.F = le; sp.frame_off = 0; sp.nbins = bins;
          sp.scale = 1.0f / sqrtf((float)nfft); sp.alpha = 1.0f; sp.mode = STFT_COMPLEX;
          sp.Z = gZ; sp.ldz = bins;
          if ((rc = launch_stft(sp, B, s))) break;
          NarrowP p{};
          p.narrow = reinterpret_cast<const float*>(gZ); p.n_line = (long long)bins * 2; p.Xn = bins; p.stats = op.fp1; p.norm_mode = 2; p.ck = 1;
          p.Y = le; p.Xw = op.in.X; p.C = op.in.C; p.S = 4; p.pad = pad;
          p.w = HP(h, op.name + ".weight"); p.bias = nullptr;
          p.yhi = op.in.hi; p.ylo = op.in.lo(); p.dy = gy.f; p.dW = dw; p.db = db;
          narrow<2, false>(p, B * le);
          gy.init = true;
          break;
        }
        case OP_TIMEFIRST: {
          auto git = G.find(op.out.key());
          float* dw = pgrad(op.name + ".weight"); float* db = pgrad(op.name + ".bias");
          if (dry || rc) break;
          if (git == G.end() || !git->second.init) { fail("no gradient reached the first time layer"); break; }
          NarrowP p{};
          p.narrow = op.fp0; p.n_line = op.i0; p.Xn = op.i0; p.stats = nullptr; p.norm_mode = 0; p.ck = 0; p.Y = 1;
          p.Xw = op.out.X; p.C = op.out.C; p.S = op.i1; p.pad = op.i2;
          p.w = HP(h, op.name + ".weight"); p.bias = HP(h, op.name + ".bias");
          p.dwide = git->second.f; p.dW = dw; p.db = db;
          narrow<1, true>(p, B);
          break;
        }
        case OP_FREQFIRST: {
          auto git = G.find(op.out.key());
          float* dw = pgrad(op.name + ".weight"); float* db = pgrad(op.name + ".bias");
          if (dry || rc) break;
          if (git == G.end() || !git->second.init) { fail("no gradient reached the first freq layer"); break; }
          NarrowP p{};
          p.narrow = op.fp0; p.n_line = (long long)op.i0 * 2; p.Xn = op.i0; p.stats = op.fp1; p.norm_mode = 1; p.ck = 0; p.Y = op.out.Y;
          p.Xw = op.out.X; p.C = op.out.C; p.S = op.i1; p.pad = op.i2;
          p.w = HP(h, op.name + ".weight"); p.bias = HP(h, op.name + ".bias");
          p.dwide = git->second.f; p.dW = dw; p.db = db;
          narrow<2, true>(p, B * op.out.Y);
          break;
        }
        case OP_CONV: conv_bwd(op); break;
        case OP_GN: gn_bwd(op); break;
        case OP_ADDCROP: {
          auto git = G.find(op.out.key());
          GradRec& ga = actgrad(op.in);
          GradRec& gs = actgrad(op.in2);
          if (dry || rc) { ga.init = true; gs.init = true; break; }
          if (git == G.end() || !git->second.init) { fail("no gradient reached a decoder skip sum"); break; }
          if (ga.init) { fail("add_crop: the cropped operand already has a gradient"); break; }
          const Ten& sk = op.in2;
          const long long items = (long long)sk.Y * op.in.X * (sk.C / 4);
          addcrop_bwd_kernel<<<dim3((unsigned)((items + 255) / 256), sk.B), 256, 0, s>>>(git->second.f, sk.Y, sk.X, sk.C, ga.f, op.in.X, op.i0, gs.f,
                                                                                         gs.init ? 1 : 0);
          chk("add_crop backward");
          ga.init = true; gs.init = true;
          break;
        }
        case OP_FREQEMB: {
          auto git = G.find(op.out.key());
          float* de = pgrad(op.name);
          if (dry || rc) break;
          if (git == G.end() || !git->second.init) { fail("no gradient reached the frequency embedding"); break; }
          const Ten& z = op.out;
          freqemb_bwd_kernel<<<dim3(ceil_div(z.X * z.C, 256), 64), 256, 0, s>>>(git->second.f, z.B, z.Y, z.X, z.C, op.f0, de);
          chk("freq_emb backward");
          break;
        }
        case OP_ADDF32: {
          auto git = G.find(op.out.key());
          if (git == G.end() || !git->second.s) { if (!dry) fail("no gradient reached the branch merge"); else { rawgrad(op.out); git = G.find(op.out.key()); } }
          if (rc) break;
          const GradRec r = git->second;
          G[op.in.key()] = r;
          G[op.in2.key()] = r;
          break;
        }
        case OP_FRAME: {
          auto git = G.find(op.out.key());
          GradRec& gx = actgrad(op.in);
          if (dry || rc) { gx.init = true; break; }
          if (git == G.end() || !git->second.init) { fail("no gradient reached the BLSTM frames"); break; }
          const Ten& xin = op.in;
          const long long items = (long long)xin.X * (xin.C / 4);
          frame_bwd_kernel<<<dim3((unsigned)((items + 255) / 256), xin.B), 256, 0, s>>>(git->second.f, xin.X, xin.C, op.i0, op.i1, op.i2, gx.f,
                                                                                        gx.init ? 1 : 0);
          chk("frame backward");
          gx.init = true;
          break;
        }
        case OP_MERGE: {
          auto git = G.find(op.out.key());
          GradRec& gl = rawgrad(op.in);
          GradRec& gx = actgrad(op.in2);
          if (dry || rc) { gl.init = true; gx.init = true; break; }
          if (git == G.end() || !git->second.init) { fail("no gradient reached a BLSTM output"); break; }
          const Ten& o = op.out;
          const int nf = op.i0, Tf = op.i1, stride = op.i2;
          const long long items = (long long)Tf * (o.C / 8);
          merge_bwd_kernel<<<dim3((unsigned)((items + 255) / 256), o.B * nf), 256, 0, s>>>(git->second.f, o.X, o.C, nf, Tf, stride, gl.s, gl.s + gl.plane);
          chk("merge backward");
          gl.init = true;
          add_or_copy(gx, git->second.f, o.elems());
          break;
        }
        case OP_LSTM: lstm_bwd(op); break;
        case OP_ATTN: attn_bwd(op); break;
        default: fail("unknown tape op"); break;
      }
    }
    prep_join();
    side_join();
    if (!dry && ok()) {
      h->act_grads.clear();
      for (auto& kv : G)
        if (kv.second.f && kv.second.init) h->act_grads[kv.first] = kv.second.f;
    }
  }
};

}  // namespace

namespace rfx {
namespace hd {
int hd_run_backward(rfx_hdemucs* h, const float* x, const float* dout, int B, int T, const std::map<std::string, float*>& grads, uint8_t* ws,
                    size_t ws_off, bool dry, cudaStream_t s, size_t* bytes) {
  BRunner R{h, B, T, ws, ws_off, dry, s, grads};
  R.run(x, dout);
  if (bytes) *bytes = R.off;
  return R.rc;
}
}  // namespace hd
}  // namespace rfx

// ------------------------------------------------------------------------------------------------
// Free-function forms of the building blocks (bwd_common.h), used by the Open-Unmix training path as well
// ------------------------------------------------------------------------------------------------
namespace rfx {
namespace bw {

int wgrad(const __nv_bfloat16* g, size_t g_plane, long long g_ld, int gcol0, int Bn, int Y, int X, const SplitAct& A, const int* dx, const int* dy,
          int taps, int N, int K, int Kp, float* stage, cudaStream_t s) {
  RFX_REQUIRE(taps >= 1 && taps <= 16 && (g_ld % 8) == 0 && (A.ld % 8) == 0 && (gcol0 % 8) == 0, "wgrad: operand strides must be multiples of 8");
  static bool attr = false;
  if (!attr) {
    RFX_CHECK_CUDA(cudaFuncSetAttribute(hd_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HT_SMEM));
    attr = true;
  }
  const int Ay = (int)(A.rows_y > 0 ? A.rows_y : 1), Ax = (int)A.rows;
  HtMap mg, ma;
  int bx = 32;
  while (bx > 1 && bx > X) bx >>= 1;
  const int by = 32 / bx;
  const unsigned box[5] = {64, (unsigned)bx, (unsigned)by, 1, 2};
  const unsigned long long gd[5] = {(unsigned long long)N, (unsigned long long)X, (unsigned long long)Y, (unsigned long long)Bn, 2};
  const unsigned long long gst[4] = {(unsigned long long)g_ld * 2, (unsigned long long)X * g_ld * 2, (unsigned long long)Y * X * g_ld * 2,
                                     (unsigned long long)g_plane * 2};
  const unsigned long long ad[5] = {(unsigned long long)K, (unsigned long long)Ax, (unsigned long long)Ay, (unsigned long long)Bn, 2};
  const long long a_ldy = A.ld_y > 0 ? A.ld_y : (long long)Ax * A.ld;
  const long long a_bs = A.batch_stride > 0 ? A.batch_stride : a_ldy * Ay;
  const unsigned long long ast[4] = {(unsigned long long)A.ld * 2, (unsigned long long)a_ldy * 2, (unsigned long long)a_bs * 2,
                                     (unsigned long long)A.plane_stride * 2};
  int rc;
  if ((rc = rfx_encode_tiled_bf16(&mg, g + gcol0, 5, gd, gst, box, 1)) || (rc = rfx_encode_tiled_bf16(&ma, A.hi, 5, ad, ast, box, 1))) return rc;
  HtParams tp{};
  tp.taps = taps; tp.N = N; tp.K = K; tp.Kp = Kp;
  for (int t = 0; t < taps; ++t) { tp.dx[t] = dx[t]; tp.dy[t] = dy[t]; }
  tp.bx = bx; tp.by = by;
  tp.tiles_x = ceil_div(X, bx);
  tp.ptiles = tp.tiles_x * ceil_div(Y, by);
  tp.ntn = ceil_div(N, 256); tp.ntk = ceil_div(K, 256);
  tp.dW = stage;
  {
    const int fr = launch_wgrad_fused(mg, ma, tp, Bn, s);
    if (fr >= 0) return fr;
  }
  const long long per = (long long)taps * tp.ntn * tp.ntk * Bn;
  long long want = (148ll * 2 + per - 1) / per;
  if (want < 1) want = 1;
  long long tchunk = (tp.ptiles + want - 1) / want;
  if (tchunk < 4) tchunk = 4;
  tp.tchunk = (int)tchunk;
  tp.nchunks = ceil_div(tp.ptiles, tp.tchunk);
  tp.dW = stage;
  RFX_REQUIRE(tp.nchunks <= 65535 && Bn <= 65535, "wgrad: grid too large");
  hd_wgrad_tc_kernel<<<dim3(taps * tp.ntn * tp.ntk, tp.nchunks, Bn), 192, HT_SMEM, s>>>(mg, ma, tp);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int colsum(const __nv_bfloat16* g, size_t g_plane, long long rows, int ld, int col0, int N, float* out, cudaStream_t s) {
  long long rpc = (rows + 591) / 592;
  if (rpc < 64) rpc = 64;
  colsum_split_kernel<<<dim3((unsigned)((rows + rpc - 1) / rpc), ceil_div(N, 2048)), 256, 0, s>>>(g, g + g_plane, rows, ld, col0, N, out, (int)rpc);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int split_pad(const float* src, long long rows, int cols, int cols_pad, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s) {
  RFX_REQUIRE(cols_pad % 8 == 0 && cols_pad >= cols, "split_pad: padded width must be a multiple of 8");
  split_pad_kernel<<<148 * 4, 256, 0, s>>>(src, rows, cols, cols_pad, hi, lo);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t transposed_pack_floats(int N, int K) {
  const int Np = ceil_div(N, 64) * 64;
  return split_weight_elems(K, Np, g2_choose_bn(K));
}

int pack_transposed(const float* W, int N, int K, float* tmp, float* store, SplitW* out, cudaStream_t s) {
  const int Np = ceil_div(N, 64) * 64;
  transpose_w_kernel<<<148 * 4, 256, 0, s>>>(W, N, 1, K, K, Np, tmp);   // wcat [n][1][Kp = K] -> wt [k][Np]
  RFX_CHECK_CUDA(cudaGetLastError());
  return pack_split_weights(tmp, Np, K, Np, g2_choose_bn(K), reinterpret_cast<__nv_bfloat16*>(store), out, s);
}

int lstm_layer_backward(const float* Gx, int Bs, int T, int H, const __nv_bfloat16* h_hi, size_t h_plane, int ldh, const SplitW& whh_f,
                        const SplitW& whh_r, const float* whh_cat, const float* dH, float* R, float* cs, float* dG, float* carry, unsigned* bar,
                        cudaStream_t s) {
  int rc;
  // 1. R = W_hh h_prev for every step: one GEMM per direction over the saved h shifted by one step
  for (int d = 0; d < 2; ++d) {
    G2Problem q;
    q.A.hi = h_hi + d * H; q.A.rows = T; q.A.rows_y = 1; q.A.ld = ldh; q.A.ld_y = (long long)T * ldh;
    q.A.batch_stride = (long long)T * ldh; q.A.plane_stride = (long long)h_plane;
    q.W = d ? whh_r : whh_f;
    q.M = T; q.My = 1; q.N = 4 * H; q.batch = Bs; q.Ktap = H; q.taps = 1;
    q.row_off[0] = d ? 1 : -1;
    int xt = 128;
    while (xt > T && xt > 1) xt >>= 1;
    q.xt = xt;
    q.Cf = R + (size_t)d * 4 * H; q.ldcf = 8 * H; q.ldcf_y = (long long)T * 8 * H; q.bscf = (long long)T * 8 * H;
    if ((rc = launch_gemm2(q, s))) return rc;
  }
  // 2. cell states
  lstm_cscan_kernel<<<ceil_div(Bs * 2 * H, 256), 256, 0, s>>>(Gx, R, Bs, T, H, cs);
  RFX_CHECK_CUDA(cudaGetLastError());
  // 3. reverse-time chain: H = 256 (Open-Unmix) on the cluster / tensor-core kernel of lstm.cu; else the persistent cooperative
  //    launch, or one launch per step when that cannot be co-resident
  bool persistent = false;
  {
    static const bool mma_on = [] { const char* e = getenv("RFX_LSTM_BWD_MMA"); return !(e && atoi(e) == 0); }();
    if (mma_on && lstm_bwd_chain_mma_supported(H)) {
      if ((rc = launch_lstm_bwd_chain_mma(Gx, R, cs, dH, whh_cat, dG, Bs, T, H, s))) return rc;
      return 0;
    }
  }
  {
    static const bool stepwise = [] { const char* e = getenv("RFX_HD_LSTM_BWD_STEPWISE"); return e && atoi(e) != 0; }();
    const dim3 pg(H / LBP_U, 2, ceil_div(Bs, LBP_B));
    const size_t psmem = ((size_t)4 * H * LBP_U + (size_t)LBP_B * (4 * H + 4) + (size_t)4 * LBP_B * LBP_U) * 4;
    int sms = 148, dev = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!stepwise && H % LBP_U == 0 && psmem <= 227 * 1024 &&
        cudaFuncSetAttribute(lstm_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_bwd_persist_kernel, 256, psmem) == cudaSuccess &&
        (long long)pg.x * pg.y * pg.z <= (long long)per_sm * sms) {
      RFX_CHECK_CUDA(cudaMemsetAsync(bar, 0, (size_t)2 * pg.z * sizeof(unsigned), s));
      const float* a0 = Gx; const float* a1 = R; const float* a2 = cs; const float* a3 = dH; const float* a4 = whh_cat;
      float* a5 = dG; int a6 = Bs, a7 = T, a8 = H; unsigned* a9 = bar;
      void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6, &a7, &a8, &a9};
      const cudaError_t e = cudaLaunchCooperativeKernel((const void*)lstm_bwd_persist_kernel, pg, dim3(256), args, psmem, s);
      if (e == cudaSuccess) persistent = true;
      else (void)cudaGetLastError();
    }
  }
  if (!persistent) {
    const size_t smem = (size_t)LB_NB * 4 * H * 4;
    RFX_CHECK_CUDA(cudaFuncSetAttribute(lstm_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(H, 128), 2, ceil_div(Bs, LB_NB));
    for (int k = 0; k < T; ++k) lstm_bwd_step_kernel<<<grid, 128, smem, s>>>(Gx, R, cs, dH, whh_cat, dG, carry, Bs, T, H, k);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int istft_adjoint_prep(const float* dout, int B, int T, const float* window, int n_fft, int hop, int frame_off, int F, int env_pad, int P0, int Ltot,
                       float* ghat, cudaStream_t s) {
  istft_adj_prep_kernel<<<dim3(ceil_div(Ltot, 256), B), 256, 0, s>>>(dout, T, window, n_fft, hop, frame_off, F, env_pad, P0, Ltot, ghat);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bw
}  // namespace rfx

extern "C" {

size_t rfx_hdemucs_train_workspace_bytes(rfx_hdemucs_t* h, int B, int T) {
  if (!h || !h->finalized || B <= 0 || T <= 0) return 0;
  size_t fwd = 0, total = 0;
  int launches = 0;
  std::vector<Op> keep;
  keep.swap(h->tape);  // a sizing call must not disturb the tape of a forward that is waiting for its backward
  const int kB = h->tape_B, kT = h->tape_T;
  const void* kws = h->tape_ws;
  const size_t kf = h->fwd_bytes;
  int rc = hd_run_forward(h, nullptr, B, T, nullptr, nullptr, true, true, nullptr, &fwd, &launches);
  if (!rc) {
    std::map<std::string, float*> none;
    rc = hd_run_backward(h, nullptr, nullptr, B, T, none, nullptr, fwd, true, nullptr, &total);
  }
  h->tape.swap(keep);
  h->tape_B = kB; h->tape_T = kT; h->tape_ws = kws; h->fwd_bytes = kf;
  return rc ? 0 : total;
}

int rfx_hdemucs_forward_train(rfx_hdemucs_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(h && x && out && workspace, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_hdemucs_finalize has not been called since the last parameter load");
  RFX_REQUIRE(B > 0 && T > 0 && T % (h->cfg.nfft / 4) == 0 && T % 1024 == 0, "T must be a positive multiple of 1024");
  RFX_REQUIRE(T > h->cfg.nfft / 8 * 3, "input shorter than the reflect padding");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const size_t need = rfx_hdemucs_train_workspace_bytes(h, B, T);
  RFX_REQUIRE(need > 0 && workspace_bytes >= need, "workspace too small (rfx_hdemucs_train_workspace_bytes)");
  return hd_run_forward(h, x, B, T, out, reinterpret_cast<uint8_t*>(workspace), false, true, (cudaStream_t)stream, nullptr, nullptr);
}

int rfx_hdemucs_backward(rfx_hdemucs_t* h, const float* x, const float* dout, int B, int T, const char* const* keys, float* const* grads, int nkeys,
                         void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(h && x && dout && workspace && keys && grads, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_hdemucs_finalize has not been called since the last parameter load");
  RFX_REQUIRE(!h->tape.empty() && h->tape_B == B && h->tape_T == T && h->tape_ws == workspace,
              "rfx_hdemucs_backward needs the workspace of the matching rfx_hdemucs_forward_train call (same B, T, pointer)");
  std::map<std::string, float*> gmap;
  for (int i = 0; i < nkeys; ++i) {
    RFX_REQUIRE(keys[i] && grads[i], "null gradient key / pointer");
    gmap[keys[i]] = grads[i];
  }
  size_t used = 0;
  int rc = hd_run_backward(h, x, dout, B, T, gmap, reinterpret_cast<uint8_t*>(workspace), h->fwd_bytes, false, (cudaStream_t)stream, &used);
  if (!rc && used > workspace_bytes) { set_error("hdemucs backward overran its workspace"); return 1; }
  return rc;
}

/* The caller guarantees that every gradient buffer handed to the following rfx_hdemucs_backward calls is already zero (e.g. views
 * of one freshly zeroed flat buffer): the backward then skips its ~400 per-parameter memsets. */
int rfx_hdemucs_set_grads_prezeroed(rfx_hdemucs_t* h, int on) {
  RFX_REQUIRE(h != nullptr, "null handle");
  h->grads_prezeroed = on != 0;
  return 0;
}

/* Weight-gradient kernel selector, process-wide: 0 = tcgen05 (default), 1 = the mma.sync tile variants (cross-check). */
int rfx_hdemucs_set_wgrad_impl(int impl) {
  RFX_REQUIRE(impl == 0 || impl == 1, "impl 0 (tcgen05, MN-major operands) or 1 (mma.sync tile variants)");
  g_hd_wgrad_impl = impl;
  return 0;
}

/* Debug: gradient of a tapped activation (names as rfx_hdemucs_tap) after rfx_hdemucs_backward, fp32 (B, Y, X, C). */
int rfx_hdemucs_grad_tap(rfx_hdemucs_t* h, const char* name, float* dst, int64_t capacity, int* dims, void* stream) {
  RFX_REQUIRE(h && name && dims, "null argument");
  auto it = h->taps.find(name);
  RFX_REQUIRE(it != h->taps.end(), "no such tap (run a training forward first)");
  const Ten& t = it->second;
  dims[0] = t.B; dims[1] = t.Y; dims[2] = t.X; dims[3] = t.C;
  if (!dst) return 0;
  auto g = h->act_grads.find(t.key());
  RFX_REQUIRE(g != h->act_grads.end(), "no gradient recorded for that tap (run rfx_hdemucs_backward first)");
  RFX_REQUIRE((size_t)capacity >= t.elems(), "destination too small");
  RFX_CHECK_CUDA(cudaMemcpyAsync(dst, g->second, t.elems() * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

/* Debug: substitute `grad` (device fp32, the tap's (B, Y, X, C) layout) for the computed gradient of tap `name` during the next
 * backward calls, so that one run can judge every layer independently of the layers after it.  grad == nullptr removes it. */
int rfx_hdemucs_inject_grad(rfx_hdemucs_t* h, const char* name, const float* grad) {
  RFX_REQUIRE(h && name, "null argument");
  if (grad) h->inject[name] = grad;
  else h->inject.erase(name);
  return 0;
}

}  // extern "C"
