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
};

struct StftParams {
  const float* x;       // (B, T) signal, batch stride x_bstride
  long long x_bstride;
  int T;
  int x_aligned8;       // float2 loads of interior frames are legal
  const float* window;  // n_fft taps (already zero-padded / centred to n_fft)
  const float2* tw;     // exp(-2 pi i m / n_fft)
  int n_fft, hop, F;
  float scale;          // 1, or n_fft^-1/2 for normalized=True
  float alpha;
  int mode;
  float2* Z;            // [B*F, ldz] complex, may be null
  int ldz;
  float* A;             // [B*F, lda] real, may be null; columns [bins, lda) are zero-filled
  int lda;
  const float* in_mean;   // STFT_UMX_MAG only
  const float* in_scale;
};

struct IstftParams {
  const float2* Z;  // [B*F, ldz]
  int ldz;
  const float* mask;  // [B*F, ldm] real multiplier or null
  int ldm;
  const float* window;
  const float2* tw;
  int n_fft, hop, F, length;
  float scale;  // 1, or n_fft^1/2 for normalized=True
  float* out;   // (B, length)
  long long out_bstride;
  int hops_per_cta;
};

const float2* twiddles(int n_fft);
int launch_stft(const StftParams& p, int B, cudaStream_t stream);
int launch_istft(const IstftParams& p, int B, cudaStream_t stream);

// ---------------------------------------------------------------- GEMM (gemm.cu)
// C[m, n] = act( ((sum_k A[m,k] W[n,k]) * s1[n] + t1[n]) * s2[n] + t2[n] )   (null vectors = identity)
enum Act { ACT_NONE = 0, ACT_TANH = 1, ACT_RELU = 2, ACT_SIGMOID = 3 };

struct Epilogue {
  const float* s1 = nullptr;
  const float* t1 = nullptr;
  const float* s2 = nullptr;
  const float* t2 = nullptr;
  int act = ACT_NONE;
};

// Weight matrix W[N, K] (row-major fp32, as torch.nn.Linear stores it) pre-split into bf16 hi/lo and
// pre-tiled into the exact SWIZZLE_128B shared-memory images the tcgen05 kernel consumes.
struct PackedW {
  void* data = nullptr;   // device; owned by whoever called pack_weights
  size_t bytes = 0;
  int N = 0, K = 0;       // logical
  int Npad = 0, Kpad = 0; // padded to BN / 64
  int BN = 0;             // 128 or 256
};
size_t packed_weight_bytes(int N, int K, int BN);
int choose_bn(int N);
int pack_weights(const float* W, int ldw, int N, int K, int BN, void* dst, PackedW* out, cudaStream_t stream);

// bf16x3 tensor-core GEMM (tcgen05, fp32-grade accuracy). A: fp32 [M, lda] with lda % 4 == 0, 16B aligned,
// readable and zero (or finite * zero weight) up to Kpad columns.
int launch_gemm_tc(const float* A, int lda, int M, const PackedW& W, float* C, int ldc, const Epilogue& e, cudaStream_t stream);
// plain fp32 FFMA GEMM with the same contract on raw weights (cross-check / RFX_GEMM=simt mode).
int launch_gemm_simt(const float* A, int lda, int M, const float* W, int ldw, int N, int K, float* C, int ldc, const Epilogue& e,
                     cudaStream_t stream);

// ---------------------------------------------------------------- LSTM recurrence (lstm.cu)
// One direction-pair of one layer: G [B*F, ldg] holds W_ih x + b_ih + b_hh for both directions
// (column = dir * 4H + gate * H + unit, gate order i,f,g,o), Whh [2][4H][H]; writes h to
// Hout[b*F + t][dir*H + unit].
int lstm_max_active_clusters();  // co-resident 8-CTA clusters of the recurrence kernel on this device
int lstm_choose_nb(int B);       // batch items per cluster used for batch size B
int launch_lstm_layer(const float* G, int ldg, const float* Whh, float* Hout, int ldh, int B, int F, int H, cudaStream_t stream);

// ---------------------------------------------------------------- small utility kernels (util.cu)
int launch_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale, float* shift,
                   int n, cudaStream_t stream);
int launch_add_vec(const float* a, const float* b, float* out, int n, cudaStream_t stream);

}  // namespace rfx
