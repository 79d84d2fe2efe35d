// TCN handle and helpers shared by tcn.cu (inference / training forward) and tcn_bwd.cu (backward).
#pragma once
#include "kernels.h"
#include "../../include/remfx_b200.h"

#include <map>
#include <string>
#include <vector>

namespace rfx {

struct TcnBuf {
  float* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (p && n == count) return 0;  // same size: keep the buffer (a training loop re-finalizes after every optimiser step)
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

struct rfx_tcn {
  rfx_tcn_config cfg;
  std::map<std::string, rfx::TcnBuf> params;
  std::vector<rfx::TcnBuf> wsplit;   // per block >= 1: split-bf16 planes of Wcat  [co][tap * C + ci]   (forward)
  std::vector<rfx::SplitW> wpack;
  std::vector<rfx::TcnBuf> wsplit_t;  // per block >= 1: split-bf16 planes of WcatT [ci][tap * C + co]   (input gradient), built lazily
  std::vector<rfx::SplitW> wpack_t;
  rfx::TcnBuf wcat;  // fp32 staging of one block's gathered weights (finalize / transposed pack), reused stream-ordered
  bool finalized = false;
  bool transposed_ready = false;
  ~rfx_tcn() {
    for (auto& kv : params) kv.second.release();
    for (auto& b : wsplit) b.release();
    for (auto& b : wsplit_t) b.release();
    wcat.release();
  }
};

namespace rfx {

inline int tcn_dilation_of(const rfx_tcn* h, int n) {
  int d = 1;
  for (int i = 0; i < n % h->cfg.stack_size; ++i) d *= h->cfg.dilation_growth;
  return d;
}
inline const float* tcn_param(const rfx_tcn* h, const std::string& k) {
  auto it = h->params.find(k);
  return it == h->params.end() ? nullptr : it->second.p;
}
// signal length after the first `nblocks` blocks (= input length of block `nblocks`)
inline long long tcn_len_after(const rfx_tcn* h, long long T, int nblocks) {
  long long L = T;
  for (int n = 0; n < nblocks; ++n) L -= (long long)(h->cfg.kernel_size - 1) * tcn_dilation_of(h, n);
  return L;
}
// offset of the residual's input sample relative to the first tap (center_crop / causal_crop, remfx/utils.py:202-211)
inline int tcn_res_off(const rfx_tcn* h, int d) {
  const int K = h->cfg.kernel_size;
  return h->cfg.causal ? (K - 1) * d - 1 : ((K - 1) * d) / 2;
}
// bytes of one bf16 plane of an activation buffer ([B][L1][C], L1 = length after block 0); every buffer uses batch stride L1 * C
inline size_t tcn_plane_bytes(const rfx_tcn* h, int B, long long T) {
  return align_up((size_t)B * (size_t)tcn_len_after(h, T, 1) * h->cfg.channel_width * 2, 256);
}

// The whole forward: block n writes its output planes (hi at block_out[n], lo at block_out[n] + plane_elems).
int tcn_run_forward(rfx_tcn* h, const float* x, int B, long long T, float* out, __nv_bfloat16* const* block_out, long long plane_elems,
                    cudaStream_t s);

}  // namespace rfx
