// Backward building blocks shared by the training paths (implemented in hdemucs_bwd.cu next to their kernels, used by umx.cu too).
#pragma once
#include "kernels.h"

namespace rfx {
namespace bw {

// dW[n][tap][k] (+)= sum_{b,y,x} G[b,y,x,gcol0+n] * A[b,y+dy[tap],x+dx[tap],k]   (A reads outside its extent are 0); `stage` is the
// [N][taps][Kp] fp32 destination (NOT zeroed here: the kernels accumulate with atomics).  G: split planes (B, Y, X, g_ld) with the lo
// plane at + g_plane elements; A as a SplitAct (rows = X extent, rows_y = Y extent).  tcgen05 form unless impl == 1 (mma.sync).
int wgrad(const __nv_bfloat16* g, size_t g_plane, long long g_ld, int gcol0, int Bn, int Y, int X, const SplitAct& A, const int* dx, const int* dy,
          int taps, int N, int K, int Kp, float* stage, cudaStream_t s);
// out[n] += sum_rows G[row][col0 + n]
int colsum(const __nv_bfloat16* g, size_t g_plane, long long rows, int ld, int col0, int N, float* out, cudaStream_t s);
// fp32 (rows, cols) -> split planes (rows, cols_pad), zero padding columns
int split_pad(const float* src, long long rows, int cols, int cols_pad, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s);
// transposed split pack for an input-gradient GEMM:  Wt[k][n] = W[n][k]  (W: [N][K] fp32 row-major); `store` must hold
// transposed_pack_floats(N, K) floats, `tmp` K * ceil64(N) floats
size_t transposed_pack_floats(int N, int K);
int pack_transposed(const float* W, int N, int K, float* tmp, float* store, SplitW* out, cudaStream_t s);
// One bidirectional LSTM layer's backward chain.  Gx [Bs][T][8H] = W_ih x + b (saved), h planes [Bs][T][ldh] (this layer's output,
// columns [0, 2H)), whh_f / whh_r = forward packs of W_hh ([4H][H], one per direction), whh_cat = fp32 [2][4H][H], dH [Bs][T][2H].
// Scratch: R, dG [Bs][T][8H], cs [Bs][T][2H], carry [Bs][2H] fp32, bar (2 * ceil(Bs / 16) unsigned).  Result: dG.
int lstm_layer_backward(const float* Gx, int Bs, int T, int H, const __nv_bfloat16* h_hi, size_t h_plane, int ldh, const SplitW& whh_f,
                        const SplitW& whh_r, const float* whh_cat, const float* dH, float* R, float* cs, float* dG, float* carry, unsigned* bar,
                        cudaStream_t s);
// iSTFT adjoint, step 1: ghat[b][P0 + i] = dout[b][i] / envelope(i), zeros in the pads (rows of length Ltot)
int istft_adjoint_prep(const float* dout, int B, int T, const float* window, int n_fft, int hop, int frame_off, int F, int env_pad, int P0, int Ltot,
                       float* ghat, cudaStream_t s);

}  // namespace bw
}  // namespace rfx
