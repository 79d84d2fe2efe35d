// Internal (C++) launch interfaces of the remfx_b200 kernels.  The public C ABI is include/remfx_b200.h.
#pragma once
#include "common.cuh"

namespace rfx {

// ---------------------------------------------------------------- STFT / iSTFT (stft.cu)
enum StftMode {
  STFT_COMPLEX = 0,    // Z only
  STFT_UMX_MAG = 1,    // Z + A = (|Z| + in_mean[k]) * in_scale[k]          (Open-Unmix front end)
  STFT_MAG = 2,        // A = |Z|
  STFT_POWER = 3,      // A = |Z|^2                                         (MelSpectrogram power=2)
  STFT_MAG_CLAMP = 4,  // A = sqrt(max(|Z|^2, 1e-8))                        (auraloss STFT magnitude)
  STFT_MAG_POW = 5,    // A = (|Z| + 1e-8)^alpha                            (remfx.utils.spectrogram)
  STFT_UMX_POW = 6,    // A = ((|Z| + 1e-8)^alpha + in_mean[k]) * in_scale[k]  (the reference's training-mode pass of the network on spectrogram(x))
};

struct StftParams {
  const float* x;       // (B, T) signal, batch stride x_bstride
  long long x_bstride;
  int T;
  int x_aligned8;       // float2 loads of interior frames are legal
  const float* window;  // n_fft taps (already zero-padded / centred to n_fft)
  const float2* tw;     // exp(-2 pi i m / n_fft)
  int n_fft, hop, F;
  int frame_off;        // frame f starts at sample f * hop - frame_off (n_fft/2 for torch.stft(center=True))
  int nbins;            // bins written per frame (n_fft/2 + 1, or n_fft/2 to drop the Nyquist bin)
  float scale;          // 1, or n_fft^-1/2 for normalized=True
  float alpha;
  int mode;
  float2* Z;            // [B*F, ldz] complex, may be null
  int ldz;
  float* A;             // [B*F, lda] real, may be null; columns [bins, lda) are zero-filled
  int lda;
  __nv_bfloat16* Ahi;   // optional split-bf16 copy of A (planes hi / lo, row stride ldas) for the tensor-core layers
  __nv_bfloat16* Alo;
  int ldas;
  const float* in_mean;   // STFT_UMX_MAG only
  const float* in_scale;
  const float2* in_ms = nullptr;  // optional interleaved (mean, scale) copy of the two: one load per bin in the staged kernel
  int max_sms = 0;        // > 0: size the grid to fill at most this many SMs (grid-stride over the frame groups); 0 = one CTA per group
  int sms_avail = 0;      // SMs the launching stream can use (a green-context partition); 0 = the whole device.  Sizes persistent grids
};

struct IstftParams {
  const float2* Z;  // [B*F, ldz]
  int ldz;
  const float* mask;  // [B*F, ldm] real multiplier or null
  int ldm;
  const float* window;
  const float2* tw;
  int n_fft, hop, F, length;
  int frame_off;      // frame t starts at output sample t * hop - frame_off (n_fft/2 for torch.istft(center=True))
  int env_pad;        // extra (all-zero) frames on each side that still count in the window envelope (HDemucs: 2)
  int nbins;          // bins present per frame in Z (n_fft/2 + 1, or n_fft/2 when the Nyquist bin is implicitly zero)
  float scale;  // 1, or n_fft^1/2 for normalized=True
  float* out;   // (B, length)
  long long out_bstride;
  int hops_per_cta;
  int max_sms = 0;    // > 0: size the grid to fill at most this many SMs (grid-stride over the output segments)
  int sms_avail = 0;  // SMs the launching stream can use (a green-context partition); 0 = the whole device.  Sizes persistent grids
};

const float2* twiddles(int n_fft);
int launch_stft(const StftParams& p, int B, cudaStream_t stream);
int launch_istft(const IstftParams& p, int B, cudaStream_t stream);

// ---------------------------------------------------------------- GEMM (gemm.cu)
// C[m, n] = act( ((sum_k A[m,k] W[n,k]) * s1[n] + t1[n]) * s2[n] + t2[n] )   (null vectors = identity)
enum Act { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3, ACT_PRELU = 4, ACT_GELU = 5,
           ACT_GLU_PAIR = 6 /* columns (2c, 2c+1) = (value, gate) -> output column c = value * sigmoid(gate); gemm2 only */ };

struct Epilogue {
  const float* s1 = nullptr;
  const float* t1 = nullptr;
  const float* s2 = nullptr;
  const float* t2 = nullptr;
  const float* slope = nullptr;  // ACT_PRELU: per-column negative slope (gemm2 only)
  int act = ACT_NONE;
};

// plain fp32 FFMA GEMM on raw fp32 operands (cross-check only; the product path is gemm2 below).
int launch_gemm_simt(const float* A, int lda, int M, const float* W, int ldw, int N, int K, float* C, int ldc, const Epilogue& e,
                     cudaStream_t stream);

// ---------------------------------------------------------------- gemm2 (gemm2.cu): TMA-fed persistent engine
// Activations between tensor-core layers live in HBM as two bf16 planes (hi, lo): v ~= hi + lo.
struct SplitAct {
  const __nv_bfloat16* hi = nullptr;  // element (b, y, x, c) at hi[b * batch_stride + y * ld_y + x * ld + c]; lo plane at + plane_stride
  long long rows = 0;                 // X extent per batch item (positions outside [0, rows) read as zero)
  long long rows_y = 0;               // Y extent (0 / 1 = plain 1-D row space)
  long long ld = 0, ld_y = 0, batch_stride = 0, plane_stride = 0;  // in elements; multiples of 8
};
struct SplitW {
  const __nv_bfloat16* hi = nullptr;  // [Npad][Kpad] row-major, zero padded; lo plane follows at + Npad * Kpad
  const __nv_bfloat16* lo = nullptr;
  int N = 0, K = 0, Npad = 0, Kpad = 0, BN = 0;
};
struct G2Problem {
  SplitAct A;
  SplitW W;
  int M = 0, N = 0, batch = 1;  // output X extent per batch item, output columns
  int My = 0;                   // output Y extent (0 / 1 = 1-D); pixel tiles are xt x (128 / xt)
  int xt = 0;                   // pixel tile width (0 = 128)
  int Ktap = 0, taps = 1;       // K extent per tap (A columns); W column index = tap * ceil64(Ktap) + k
  int row_off[16] = {0};        // A x-offset per tap (implicit-GEMM convolution)
  int row_off_y[16] = {0};      // A y-offset per tap
  bool dual = false;            // last tap -> second accumulator, added after the activation (TCN residual)
  float* Cf = nullptr;          // fp32 output (optional): element (b, y, x, n) at b * bscf + y * ldcf_y + x * ldcf + n
  long long ldcf = 0, ldcf_y = 0, bscf = 0;
  bool cf_accum = false;        // Cf += result instead of Cf = result (fp32 output only, ACT_NONE)
  bool cf_pre_act = false;      // both outputs given: Cf receives the pre-activation (bias / scale applied), Chi / Clo the activated value
  __nv_bfloat16* Chi = nullptr;  // split-bf16 output (optional)
  __nv_bfloat16* Clo = nullptr;
  long long ldcs = 0, ldcs_y = 0, bscs = 0;
  Epilogue epi;
  // optional fused GroupNorm statistics of the output (see gemm2.cu): accum must be zeroed by the caller
  double* gn_acc = nullptr;
  int gn_G = 1, gn_per_x = 0, gn_cmod = 0;
  int max_ctas = 0;  // > 0: persistent grid of at most this many CTAs (leaves the other SMs to concurrent kernels)
};
// Process-wide matmul precision of the tensor-core kernels (gemm2, the tcgen05 LSTM recurrence):
//   0 = fp32-parity: bf16x3 (lo*hi + hi*lo + hi*hi), the default and the mode every parity gate is stated in;
//   1 = bf16-fast: the hi*hi pass only (BASELINE.json configs[1] says "bf16"): 3x fewer MMAs, ~3e-4 relative error on Open-Unmix.
void set_matmul_precision(int mode);
int get_matmul_precision();
int g2_choose_bn(int N);
size_t split_weight_elems(int N, int K, int BN);  // elements of ONE plane
int pack_split_weights(const float* W, long long ldw, int N, int K, int BN, __nv_bfloat16* dst, SplitW* out, cudaStream_t stream);
int launch_split_rows(const float* src, long long ld_src, int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, long long ld_dst,
                      int rows_pad, int cols_pad, cudaStream_t stream);
int launch_gemm2(const G2Problem& pr, cudaStream_t stream);

// ---------------------------------------------------------------- LSTM recurrence (lstm.cu)
// One direction-pair of one layer: G [B*F, ldg] holds W_ih x + b_ih + b_hh for both directions
// (column = dir * 4H + gate * H + unit, gate order i,f,g,o), Whh [2][4H][H]; writes h to
// Hout[b*F + t][dir*H + unit].
int lstm_max_active_clusters();  // co-resident 8-CTA clusters of the recurrence kernel on this device
int lstm_choose_nb(int B);       // batch items per cluster used for batch size B (FFMA variant)
void lstm_set_impl(int impl);    // 0 = tensor-core mma.sync bf16x3 (default), 1 = fp32 FFMA
int lstm_get_impl();
// Hout (fp32) and/or Hhi/Hlo (split bf16 planes, row stride ldhs) may be given.
int launch_lstm_layer(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs,
                      int B, int F, int H, cudaStream_t stream);
// Same, with the batch slots per cluster fixed by the caller (1..8; 0 = as few as keeps the launch one wave).  8 packs the
// launch into the fewest SMs (2 * ceil(B / 8) clusters of 8), which is what the multi-lane Open-Unmix pipeline wants.
int launch_lstm_layer_slots(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo,
                            int ldhs, int B, int F, int H, int slots, cudaStream_t stream);
// Same with the kernel chosen per call: impl 0 = mma.sync bf16x3, 1 = fp32 FFMA, 2 = tcgen05 (H = 256; 16 slots per cluster),
// -1 = the process-wide default (lstm_set_impl).
int launch_lstm_layer_impl(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo,
                           int ldhs, int B, int F, int H, int impl, int slots, cudaStream_t stream);
int lstm_clusters_for(int B, int slots);  // clusters (of 8 CTAs) one launch of the tensor-core recurrence uses
// Reverse-time chain of the LSTM backward for H = 192 / 256 / 384 on the cluster / mma.sync machinery (see lstm.cu): Gx, R = W_hh h_prev, cell
// states cs, dH ([Bs][T][...] layouts of bw::lstm_layer_backward) -> dG.
bool lstm_bwd_chain_mma_supported(int H);  // 192, 256, 384
int launch_lstm_bwd_chain_mma(const float* Gx, const float* R, const float* cs, const float* dH, const float* Whh, float* dG, int Bs, int T, int H,
                              cudaStream_t stream);

// ---------------------------------------------------------------- small utility kernels (util.cu)
int launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale, float* shift,
                   int n, cudaStream_t stream);
int launch_add_vec(const float* a, const float* b, float* out, int n, cudaStream_t stream);

}  // namespace rfx
