// C ABI glue for the stand-alone operators (include/remfx_b200.h) + thread-local error state.
#include "kernels.h"
#include "../../include/remfx_b200.h"

#include <cmath>

namespace rfx {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* get_error() { return g_err.c_str(); }
}  // namespace rfx

using namespace rfx;

extern "C" {

int rfx_abi_version(void) { return RFX_ABI_VERSION; }
const char* rfx_last_error(void) { return get_error(); }

int rfx_device_supported(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

int rfx_stft(const float* x, int B, int T, int n_fft, int hop, const float* window, int normalized, int mode, float alpha, float* Z_ri,
             float* A, void* stream) {
  RFX_REQUIRE(x && window, "null argument");
  RFX_REQUIRE(B > 0 && T > 0 && hop > 0, "positive sizes");
  RFX_REQUIRE(Z_ri || A, "at least one output");
  RFX_REQUIRE(mode == STFT_COMPLEX || (mode >= STFT_MAG && mode <= STFT_MAG_POW), "mode must be 0, 2, 3, 4 or 5");
  RFX_REQUIRE(mode == STFT_COMPLEX || A, "real-valued modes need A");
  RFX_REQUIRE(((uintptr_t)window & 7) == 0, "window must be 8-byte aligned");
  StftParams p{};
  p.x = x; p.x_bstride = T; p.T = T;
  p.x_aligned8 = (((uintptr_t)x & 7) == 0 && T % 2 == 0 && hop % 2 == 0) ? 1 : 0;
  p.window = window;
  p.tw = twiddles(n_fft);
  p.n_fft = n_fft; p.hop = hop; p.F = T / hop + 1;
  p.frame_off = n_fft / 2; p.nbins = n_fft / 2 + 1;
  p.scale = normalized ? 1.0f / sqrtf((float)n_fft) : 1.0f;
  p.alpha = alpha; p.mode = mode;
  p.Z = reinterpret_cast<float2*>(Z_ri); p.ldz = n_fft / 2 + 1;
  p.A = (mode == STFT_COMPLEX) ? nullptr : A; p.lda = n_fft / 2 + 1;
  return launch_stft(p, B, (cudaStream_t)stream);
}

int rfx_istft(const float* Z_ri, const float* mask, int B, int F, int n_fft, int hop, const float* window, int normalized, int length,
              float* out, void* stream) {
  RFX_REQUIRE(Z_ri && window && out, "null argument");
  RFX_REQUIRE(B > 0 && F > 0 && hop > 0 && length > 0, "positive sizes");
  RFX_REQUIRE(((uintptr_t)window & 7) == 0 && ((uintptr_t)Z_ri & 7) == 0, "window / Z must be 8-byte aligned");
  IstftParams p{};
  p.Z = reinterpret_cast<const float2*>(Z_ri); p.ldz = n_fft / 2 + 1;
  p.mask = mask; p.ldm = n_fft / 2 + 1;
  p.window = window; p.tw = twiddles(n_fft);
  p.n_fft = n_fft; p.hop = hop; p.F = F; p.length = length;
  p.frame_off = n_fft / 2; p.env_pad = 0; p.nbins = n_fft / 2 + 1;
  p.scale = normalized ? sqrtf((float)n_fft) : 1.0f;
  p.out = out; p.out_bstride = length;
  p.hops_per_cta = 16;
  while (p.hops_per_cta > 1 && (size_t)p.hops_per_cta * hop * 4 > 96 * 1024) p.hops_per_cta >>= 1;
  return launch_istft(p, B, (cudaStream_t)stream);
}

size_t rfx_gemm_scratch_bytes(int M, int N, int K) {
  const size_t kpad = (size_t)ceil_div(K, 64) * 64;
  return align_up((size_t)M * kpad * 2 * 2, 256) + align_up(split_weight_elems(N, K, g2_choose_bn(N)) * 2 * 2, 256);
}

int rfx_gemm(int impl, const float* A, int lda, int M, const float* W, int N, int K, float* C, int ldc, const float* s1, const float* t1,
             const float* s2, const float* t2, int act, void* scratch, void* stream) {
  RFX_REQUIRE(A && W && C, "null argument");
  RFX_REQUIRE(impl == 0 || impl == 1, "impl 0 or 1");
  RFX_REQUIRE(M > 0 && N > 0 && K > 0, "positive sizes");
  Epilogue e;
  e.s1 = s1; e.t1 = t1; e.s2 = s2; e.t2 = t2; e.act = act;
  cudaStream_t s = (cudaStream_t)stream;
  if (impl == 1) return launch_gemm_simt(A, lda, M, W, K, N, K, C, ldc, e, s);
  RFX_REQUIRE(scratch != nullptr && ((uintptr_t)scratch & 255) == 0, "impl 0 needs 256-byte aligned scratch (rfx_gemm_scratch_bytes)");
  // split both operands into bf16 hi/lo planes, then run the TMA-fed tcgen05 engine
  const int kpad = ceil_div(K, 64) * 64;
  __nv_bfloat16* a_hi = reinterpret_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* a_lo = a_hi + (size_t)M * kpad;
  int rc = launch_split_rows(A, lda, M, K, a_hi, a_lo, kpad, M, kpad, s);
  if (rc) return rc;
  __nv_bfloat16* wdst = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(scratch) + align_up((size_t)M * kpad * 4, 256));
  G2Problem pr;
  if ((rc = pack_split_weights(W, K, N, K, g2_choose_bn(N), wdst, &pr.W, s))) return rc;
  pr.A.hi = a_hi; pr.A.rows = M; pr.A.ld = kpad; pr.A.batch_stride = 0; pr.A.plane_stride = (long long)M * kpad;
  pr.M = M; pr.N = N; pr.batch = 1; pr.Ktap = K; pr.taps = 1;
  pr.Cf = C; pr.ldcf = ldc; pr.bscf = 0;
  pr.epi = e;
  return launch_gemm2(pr, s);
}

int rfx_lstm_info(int B, int* max_active_clusters, int* batch_per_cluster) {
  RFX_REQUIRE(B > 0 && max_active_clusters && batch_per_cluster, "bad argument");
  *max_active_clusters = lstm_max_active_clusters();
  *batch_per_cluster = lstm_choose_nb(B);
  return 0;
}

int rfx_set_matmul_precision(int mode) {
  RFX_REQUIRE(mode == 0 || mode == 1, "mode 0 (fp32-parity: bf16x3) or 1 (bf16-fast: single pass)");
  set_matmul_precision(mode);
  return 0;
}
int rfx_get_matmul_precision(void) { return get_matmul_precision(); }

int rfx_lstm_set_impl(int impl) {
  RFX_REQUIRE(impl >= 0 && impl <= 2, "impl 0 (mma.sync tensor-core), 1 (fp32 FFMA) or 2 (tcgen05, H = 256)");
  lstm_set_impl(impl);
  return 0;
}

int rfx_lstm_layer(const float* G, const float* Whh, float* Hout, int ldh, int B, int F, int H, void* stream) {
  RFX_REQUIRE(G && Whh && Hout, "null argument");
  return launch_lstm_layer(G, 8 * H, Whh, Hout, ldh, nullptr, nullptr, 0, B, F, H, (cudaStream_t)stream);
}

int rfx_lstm_layer_slots(const float* G, const float* Whh, float* Hout, int ldh, int B, int F, int H, int slots, void* stream) {
  RFX_REQUIRE(G && Whh && Hout, "null argument");
  RFX_REQUIRE(slots >= 0 && slots <= 32, "slots per cluster must be 0 (automatic) or 1..32");
  return launch_lstm_layer_slots(G, 8 * H, Whh, Hout, ldh, nullptr, nullptr, 0, B, F, H, slots, (cudaStream_t)stream);
}

}  // extern "C"
