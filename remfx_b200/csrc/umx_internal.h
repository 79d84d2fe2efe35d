// Open-Unmix handle, workspace layout and small helpers shared by umx.cu (inference, pipeline) and umx_train.cu (training step).
#pragma once
#include "kernels.h"
#include "../../include/remfx_b200.h"

#include <cuda.h>

#include <map>
#include <string>
#include <vector>

namespace rfx {

struct DevBuf {
  float* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    if (p && n == count) return 0;  // same size: keep the buffer (every user overwrites it completely; a training loop re-finalizes
                                    // after each optimiser step and cudaFree / cudaMalloc would synchronise the device every time)
    release();
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

struct rfx_umx {
  rfx_umx_config cfg;
  int bins = 0, H = 0;
  std::map<std::string, rfx::DevBuf> params;
  // derived at finalize()
  rfx::DevBuf bn_s[3], bn_t[3], in_ms;
  std::vector<rfx::DevBuf> lstm_bias, wih_cat, whh_cat;
  std::vector<rfx::DevBuf> packed_store;  // split-bf16 (hi, lo) weight planes
  rfx::SplitW fc1p, fc2p, fc3p;
  std::vector<rfx::SplitW> wihp;
  bool finalized = false;
  // training path (umx_train.cu): packs of the backward GEMMs, built lazily after a finalize
  bool train_ready = false;
  std::vector<rfx::DevBuf> train_store;
  rfx::DevBuf train_tmp;           // transposition scratch of the packs
  std::vector<rfx::SplitW> whhp;   // [2 L] W_hh per (layer, direction) as forward packs (gate recompute of the backward)
  std::vector<rfx::SplitW> wih_t;  // [L] transposed packs of [W_ih ; W_ih_reverse] (input gradient of a layer)
  rfx::SplitW fc1_t, fc2_t, fc3_t;
  int tape_B = 0, tape_T = 0;      // shape / workspace / dropout flag of the forward_train a backward may follow
  const void* tape_ws = nullptr;
  const float* tape_masks = nullptr;
  // optional per-stage timing (cudaEvents recorded on the caller's stream between the launches)
  bool profiling = false;
  std::vector<cudaEvent_t> events;
  int mark_idx = 0;
  // host-buffer pipeline (rfx_umx_sample_host / submit_host / wait_host): two slots, item-chunked copies on two internal streams.
  // The multi-lane pipeline (rfx_umx_pipe_*) reuses the same copy streams and per-slot events with slot = lane.
  static constexpr int kSlots = 8, kHostSlots = 2, kChunks = 4;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  cudaEvent_t ev_in[kSlots][kChunks] = {}, ev_ist[kSlots][kChunks] = {}, ev_out[kSlots] = {};
  bool pending[kSlots] = {};
  // multi-lane pipeline state (see rfx_umx_pipe_push)
  struct Lane {
    cudaStream_t s = nullptr;
    cudaEvent_t ev_x = nullptr, ev_pre = nullptr, ev_rec = nullptr, ev_stft = nullptr;
    const float* x = nullptr; float* out = nullptr;          // device buffers of the step in the lane
    const float* x_host = nullptr; float* out_host = nullptr;  // host buffers (null = device-resident)
    long long seq = -1;
    int next_stage = 0;
    bool live = false, stft_recorded = false;
    cudaEvent_t host_out_pending = nullptr;  // completion event of the last D2H out of this lane's output staging buffer
  };
  static constexpr int kRing = 16;  // completion events are kept for the last kRing steps
  struct Done { cudaEvent_t ev = nullptr; long long seq = -1; bool recorded = false; };
  struct Pipe {
    bool ready = false;
    int depth = 0, B = 0, T = 0;
    cudaStream_t rec[4] = {nullptr, nullptr, nullptr, nullptr};  // recurrence launches, round-robin in issue order, at the highest stream priority
    int rec_n = 1;                              // recurrence streams in use (2 = two launches side by side)
    long long rec_count = 0;
    Lane lane[kSlots];
    Done done[kRing];
    long long pushed = 0;
    void* ws = nullptr;
    bool free_run = false;
    int sms = 0, max_sms = 0, gemm_ctas = 0, lstm_slots = 0, lstm_impl = -1;
    // SM partition (CUDA green contexts): the recurrence streams own `rec_sms_granted` SMs, every other stream the rest.
    // When the driver cannot provide it the pipeline falls back to capping the grids of the non-recurrent kernels.
    CUgreenCtx gctx_rec = nullptr, gctx_rest = nullptr;
    int rec_sms_granted = 0, rest_sms_granted = 0;
    // optional timing of the recurrence launches (a pair of timing events around each, on the recurrence stream)
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    int prof_n = 0;
  } pipe;

  ~rfx_umx();
};


namespace rfx {

// Workspace layout.  Activations between tensor-core layers are split-bf16 planes (hi then lo).
struct UmxLayout {
  int F, M, lda1, ldm, ldz;
  size_t off_x[2], off_out[2], off_Z, off_A1, off_XC, off_G, off_H1, off_H2, off_Y2, off_mask, total;
  size_t plane_A1, plane_XC, plane_H, plane_Y2;  // elements per plane
};

UmxLayout umx_layout(const rfx_umx* h, int B, int T);
const float* umx_param(const rfx_umx* h, const std::string& k);
// One dense layer on the tensor-core engine: A (split planes, K columns) x W^T -> fp32 and/or split output.
int umx_dense(const __nv_bfloat16* a_hi, size_t a_plane, int lda, int M, int K, const SplitW& W, float* Cf, int ldcf, __nv_bfloat16* c_hi,
              size_t c_plane, int ldcs, const Epilogue& e, cudaStream_t s, int max_ctas = 0);

}  // namespace rfx
