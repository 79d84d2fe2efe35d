/* remfx_b200 -- C ABI of the Blackwell-native RemFx hot path (libremfx_b200.so).
 *
 * Plain C: opaque handles, raw device pointers, sizes; no torch / C++ types.  Every pointer named
 * `*_dev` (or undecorated) is a CUDA device pointer on the current device unless the function name ends
 * in `_host`.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All functions
 * return 0 on success; on failure they return non-zero and rfx_last_error() describes the problem
 * (thread-local).  Work is enqueued asynchronously on `stream` unless stated otherwise.
 *
 * Each entry point replaces one piece of the reference's Python plug-in surface (file:line under
 * mhrice/RemFx @ 85d5030); see INTEGRATION.md for the reference-side binding.
 */
#ifndef REMFX_B200_H
#define REMFX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RFX_ABI_VERSION 1

int rfx_abi_version(void);
/* Description of the most recent failure on the calling thread ("" if none). */
const char* rfx_last_error(void);
/* 1 if the current device is a compute-capability 10.x part the library was built for. */
int rfx_device_supported(void);

/* ---------------------------------------------------------------------------------------------
 * S1/S2/S4  torch.stft(center=True, reflect, onesided) + epilogue
 *   replaces umx/openunmix/transforms.py:89-120 (TorchSTFT.forward), :198-216 (ComplexNorm),
 *            remfx/utils.py:138-159 (spectrogram)
 * x: (B, T) fp32.  window: n_fft taps (caller zero-pads shorter windows, as torch.stft does).
 * Output layout is FRAME-major: Z[(b*F + t)*bins + k] (re,im interleaved), A likewise (real), with
 * F = T/hop + 1, bins = n_fft/2 + 1.  mode: 0 complex only (A unused), 2 |Z|, 3 |Z|^2,
 * 4 sqrt(max(|Z|^2,1e-8)), 5 (|Z|+1e-8)^alpha.  Z may be NULL when only A is wanted.
 * n_fft in {512, 1024, 2048, 4096}.
 * ------------------------------------------------------------------------------------------- */
int rfx_stft(const float* x, int B, int T, int n_fft, int hop, const float* window, int normalized,
             int mode, float alpha, float* Z_ri, float* A, void* stream);

/* S3  torch.istft(center=True, length=length)
 *   replaces umx/openunmix/transforms.py:164-181 (TorchISTFT.forward)
 * Z_ri: frame-major complex [(b*F + t)*bins + k]; mask (optional, same indexing, real) is multiplied in. */
int rfx_istft(const float* Z_ri, const float* mask, int B, int F, int n_fft, int hop, const float* window,
              int normalized, int length, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense layer  C[m,n] = act(((A[m,:] . W[n,:]) * s1[n] + t1[n]) * s2[n] + t2[n])
 *   the primitive behind nn.Linear + BatchNorm1d(eval) + activation in umx/openunmix/model.py:119-161
 * impl: 0 = TMA-fed tcgen05 bf16x3 tensor-core engine, 1 = fp32 FFMA cross-check kernel.  act: 0 none,
 * 1 tanh, 2 relu, 3 sigmoid.  s1/t1/s2/t2 may be NULL.  W is the raw nn.Linear weight [N, K] (row-major);
 * for impl 0 both operands are split into bf16 hi/lo planes in `scratch` (rfx_gemm_scratch_bytes) first.
 * Test / utility entry point -- the model handles below keep weights pre-split and activations split.
 * ------------------------------------------------------------------------------------------- */
size_t rfx_gemm_scratch_bytes(int M, int N, int K);
int rfx_gemm(int impl, const float* A, int lda, int M, const float* W, int N, int K, float* C, int ldc,
             const float* s1, const float* t1, const float* s2, const float* t2, int act,
             void* scratch, void* stream);

/* One bidirectional LSTM layer recurrence (torch.nn.LSTM, gate order i,f,g,o; zero initial state)
 *   replaces the recurrent half of umx/openunmix/model.py:141 (self.lstm)
 * G: [B*F, 8H] input projections incl. both biases, column = dir*4H + gate*H + unit, row = b*F + t.
 * Whh: [2, 4H, H].  Hout: [B*F, ldh], column = dir*H + unit.  H must be 256. */
/* Scheduling facts of the recurrence kernel on the current device: how many 8-CTA clusters are co-resident
 * and how many batch items each cluster takes for batch size B (all clusters of a launch run as one wave). */
int rfx_lstm_info(int B, int* max_active_clusters, int* batch_per_cluster);
/* Matmul precision of the tensor-core kernels (gemm2 and the tcgen05 LSTM recurrence), process-wide:
 *   0 = fp32-parity (default): every product as bf16x3 (lo*hi + hi*lo + hi*hi), the mode all parity gates are stated in
 *       (the reference computes in fp32, cfg/config.yaml:112);
 *   1 = bf16-fast: the hi*hi pass only -- what BASELINE.json configs[1] ("1xB200 bf16") names; ~3e-4 relative RMS on Open-Unmix
 *       (3.5e-4), reported by bench.py beside the parity-mode headline. */
int rfx_set_matmul_precision(int mode);
int rfx_get_matmul_precision(void);

/* Process-wide choice of the recurrence kernel: 0 = tensor-core (mma.sync bf16x3, default), 1 = fp32 FFMA, 2 = tcgen05 (H = 256). */
int rfx_lstm_set_impl(int impl);
int rfx_lstm_layer(const float* G, const float* Whh, float* Hout, int ldh, int B, int F, int H, void* stream);
/* Same with the batch slots per 8-CTA cluster fixed by the caller: 1..8 for the mma.sync / FFMA kernels, 1..16 for the tcgen05
 * kernel (rfx_lstm_set_impl(2): H = 256, W_hh as the TMEM A operand, 16 slots per cluster = half the SMs per launch, used by
 * the Open-Unmix pipeline); 0 = automatic as above. */
int rfx_lstm_layer_slots(const float* G, const float* Whh, float* Hout, int ldh, int B, int F, int H, int slots, void* stream);

/* ---------------------------------------------------------------------------------------------
 * U1-U5  Open-Unmix effect-removal model
 *   replaces remfx/models.py:259-304 (OpenUnmixModel.sample / eval-mode forward output) =
 *   umx/openunmix/model.py:242-319 (Separator.forward) around :107-166 (OpenUnmix.forward) and
 *   umx/openunmix/filtering.py:442-459 (wiener, niter=0)
 * ------------------------------------------------------------------------------------------- */
typedef struct rfx_umx rfx_umx_t;

typedef struct {
  int n_fft;       /* 2048 */
  int hop;         /* 512  */
  int hidden;      /* 512 (LSTM hidden = hidden/2 per direction; must be 512) */
  int nb_layers;   /* 3 */
  int gemm_impl;   /* must be 0: TMA-fed tcgen05 bf16x3 engine */
} rfx_umx_config;

int rfx_umx_create(const rfx_umx_config* cfg, rfx_umx_t** out);
void rfx_umx_destroy(rfx_umx_t* h);
/* Copy one tensor of the reference state_dict into the handle (device-to-device, on `stream`).
 * `key` is the OpenUnmix parameter name without the wrapper prefix, e.g. "fc1.weight",
 * "bn1.running_var", "lstm.weight_hh_l0_reverse", "input_mean", or "window" for the STFT window. */
int rfx_umx_load_param(rfx_umx_t* h, const char* key, const float* src, int64_t numel, void* stream);
/* Fold BatchNorm statistics, add LSTM biases and pack the GEMM weights; call after loading params. */
int rfx_umx_finalize(rfx_umx_t* h, void* stream);
size_t rfx_umx_workspace_bytes(const rfx_umx_t* h, int B, int T);
/* x: (B, 1, T) fp32 device -> out: (B, 1, T) fp32 device.  workspace: >= rfx_umx_workspace_bytes. */
int rfx_umx_sample(rfx_umx_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* Same, from / to (pinned) HOST buffers.  The batch travels in up to 4 item chunks on two internal copy streams: the STFT of
 * chunk c starts as soon as chunk c is in HBM, and each chunk's iSTFT is followed by its own D2H copy, so most of the PCIe
 * time hides behind the kernels; returns when out_host is complete. */
int rfx_umx_sample_host(rfx_umx_t* h, const float* x_host, int B, int T, float* out_host, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Pipelined form of the above: submit returns as soon as the work is enqueued; wait blocks until that slot's out_host is
 * complete.  Two slots (0, 1) with private device staging inside the workspace: submit(slot k+1) overlaps its H2D with the
 * kernels of slot k, whose D2H in turn overlaps the kernels of slot k+1.  Re-submitting a slot waits for its previous use.
 * x_host / out_host must stay valid (and should be pinned) until the matching wait; all submits of one handle must use the
 * same compute `stream` and the same workspace. */
int rfx_umx_submit_host(rfx_umx_t* h, int slot, const float* x_host, int B, int T, float* out_host, void* workspace,
                        size_t workspace_bytes, void* stream);
int rfx_umx_wait_host(rfx_umx_t* h, int slot);
/* Multi-lane pipeline: the throughput form of rfx_umx_sample for a stream of equally shaped batches (a serving loop over
 * remfx/models.py:303-304).  The bidirectional recurrence is a chain of strictly dependent steps that occupies a fraction of
 * the chip, and nothing inside one batch can overlap it; consecutive batches can.  push(n) enqueues the whole of batch n on lane
 * n mod depth (a stream + private workspace; depth = rfx_umx_pipe_depth batches in flight); the recurrence of LSTM layer l
 * always goes to recurrence stream l, so those streams run back to back (tcgen05 kernel, 32 slots per cluster), on their own SM
 * partition (CUDA green contexts; grid caps as the fallback), while every other kernel of the batches in flight runs beside
 * them.  Results equal rfx_umx_sample's to ~1e-6 (same math, the recurrence kernel sums in a different order).
 *   push   enqueues batch `*seq` (0, 1, 2, ... since the handle was created) and returns; x / out are device buffers, or
 *          (pinned) host buffers when x_on_host / out_on_host is non-zero (H2D / D2H then ride the internal copy streams).
 *          The lane starts after the work already enqueued on `stream`.  x and out must stay valid and untouched until
 *          the batch's output is complete.
 *   flush  makes `stream` wait for every output enqueued so far (and, in the staggered schedule -- as many lanes as LSTM
 *          layers, RFX_UMX_PIPE_LANES -- first runs the stages still owed to the batches in flight).
 *   wait / stream_wait  block the host / make `stream` wait until batch seq's output is complete (completion records are kept
 *          for the last 16 batches; in the staggered schedule a batch must first have left the pipeline).  A consumer that
 *          stays `depth` or more batches behind its pushes never starves the device.
 * workspace: rfx_umx_pipe_workspace_bytes (= depth private lanes), the same pointer for every push until a flush. */
size_t rfx_umx_pipe_workspace_bytes(const rfx_umx_t* h, int B, int T);
int rfx_umx_pipe_depth(const rfx_umx_t* h);
/* Schedule facts after the first push: SMs owned by the recurrence streams (a CUDA green-context partition; 0 when the driver
 * could not provide one and the other kernels' grids are capped instead), SMs left to every other kernel, recurrence launches
 * that run side by side, batch slots per recurrence cluster. */
int rfx_umx_pipe_info(const rfx_umx_t* h, int* rec_sms, int* rest_sms, int* rec_streams, int* slots_per_cluster);
int rfx_umx_pipe_push(rfx_umx_t* h, const float* x, int x_on_host, int B, int T, float* out, int out_on_host, void* workspace,
                      size_t workspace_bytes, void* stream, long long* seq);
int rfx_umx_pipe_flush(rfx_umx_t* h, void* stream);
int rfx_umx_pipe_wait(rfx_umx_t* h, long long seq);
/* Non-blocking: *done = 1 once batch `seq`'s output (and, in host mode, its D2H copy) is complete; 0 while it is in flight. */
int rfx_umx_pipe_query(rfx_umx_t* h, long long seq, int* done);
int rfx_umx_pipe_stream_wait(rfx_umx_t* h, long long seq, void* stream);
/* Timing of the pipeline's recurrence launches: set_profiling(n > 0) brackets the next n launches with cudaEvents on the
 * recurrence stream (0 switches it off); rec_times returns their durations in ms, in launch order, once they have run. */
int rfx_umx_pipe_set_profiling(rfx_umx_t* h, int max_launches);
int rfx_umx_pipe_rec_times(rfx_umx_t* h, float* ms, int capacity, int* n_out);
/* Number of kernels one rfx_umx_sample call launches (for bench.py's gpu_launches). */
int rfx_umx_launches_per_call(const rfx_umx_t* h);
/* Per-stage device timing: when enabled, rfx_umx_sample records a cudaEvent on `stream` between its
 * kernel launches; after the stream has been synchronised rfx_umx_stage_times returns the elapsed ms of
 * the last call's stages in launch order: stft, fc1, (w_ih GEMM, lstm recurrence) x nb_layers, fc2, fc3, istft. */
int rfx_umx_set_profiling(rfx_umx_t* h, int on);
int rfx_umx_stage_times(rfx_umx_t* h, float* ms, int capacity, int* n_out);
/* Debug tap: copy the ratio mask (what = 3; M x ldm fp32) of the last call into dst; row stride via *ld. */
int rfx_umx_debug_tap(rfx_umx_t* h, int what, const void* workspace, int B, int T, float* dst, int* ld, void* stream);

/* Training (row L5 for Open-Unmix): the reference's Lightning step runs `OpenUnmixModel.forward` (remfx/models.py:294-301) in
 * TRAINING mode and calls loss.backward() on its output (remfx/models.py:217-221):
 *   rfx_umx_forward_train  one pass of the network with BatchNorm1d batch statistics (umx/openunmix/model.py:135,151,157) and
 *                          inter-layer LSTM dropout (model.py:62-69), every pre-BatchNorm activation kept in `workspace`.
 *        drop_masks   NULL (no dropout) or (nb_layers - 1, B * frames, hidden) fp32 inverted-dropout masks (0 or 1 / (1 - p)),
 *                     drawn by the caller -- no reference RNG stream can be matched, and the parity tests inject theirs;
 *        pow_pass     1 = the reference's extra pass `Y = self.model(spectrogram(x))` (remfx/models.py:296-297): the network
 *                     input is ((|STFT| + 1e-8)^alpha + input_mean) * input_scale, the pass stops after the third BatchNorm
 *                     (`out` may be NULL) and only bn_stats_out matters; 0 = the separator pass (|STFT| input, mask, iSTFT -> out);
 *        bn_stats_out NULL or 2 * (hidden + hidden + bins) floats: [mean1, var1, mean2, var2, mean3, var3], biased batch
 *                     variances, for the caller's running_mean / running_var update (torch: momentum 0.1, unbiased variance).
 *   rfx_umx_backward       given dout = dLoss / dout (B, T): dLoss / dparameter for every state_dict key in keys[] / grads[] (device
 *                          fp32 buffers of the parameter's size, overwritten).  Must follow a pow_pass = 0 forward_train with the
 *                          same shape and workspace pointer, no rfx_umx_finalize in between.  No gradient is produced for x (the
 *                          reference feeds the network a detached spectrogram, model.py:267).
 * workspace: rfx_umx_train_workspace_bytes(h, B, T) bytes, 256-byte aligned.  Any T > n_fft / 2 with T % 4 == 0. */
size_t rfx_umx_train_workspace_bytes(const rfx_umx_t* h, int B, int T);
/* Builds the backward's weight packs (transposed / per-direction) for the current parameters on `stream`; forward_train does it
 * implicitly, a caller that runs the two passes of a step on two streams calls it first so that both see finished packs. */
int rfx_umx_train_prepare(rfx_umx_t* h, void* stream);
int rfx_umx_forward_train(rfx_umx_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                          const float* drop_masks, int pow_pass, float alpha, float* bn_stats_out, void* stream);
int rfx_umx_backward(rfx_umx_t* h, const float* x, const float* dout, int B, int T, const char* const* keys, float* const* grads,
                     int nkeys, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * T1-T3  TCN effect-removal model
 *   replaces remfx/models.py:370-390 (TCNModel.forward / sample) = remfx/tcn.py:126-130 (TCN.forward):
 *   nblocks x TCNBlock (remfx/tcn.py:48-59: dilated Conv1d, PReLU, 1x1 residual + crop) then tanh(1x1 conv).
 * Parameter keys are the reference's names below `model.`: "process_blocks.3.conv1.weight", "output.bias", ...
 * x: (B, 1, T) fp32 device -> out: (B, 1, rfx_tcn_out_length(T)) fp32 device.
 * ------------------------------------------------------------------------------------------- */
typedef struct rfx_tcn rfx_tcn_t;

typedef struct {
  int ninputs;         /* 1 */
  int noutputs;        /* 1 */
  int nblocks;         /* 20 */
  int channel_width;   /* 256 (multiple of 64, <= 256) */
  int kernel_size;     /* 7 (odd, <= 15) */
  int stack_size;      /* 10 */
  int dilation_growth; /* 2 */
  int causal;          /* 0: center_crop residual (cfg/model/tcn.yaml), 1: causal_crop */
} rfx_tcn_config;

int rfx_tcn_create(const rfx_tcn_config* cfg, rfx_tcn_t** out);
void rfx_tcn_destroy(rfx_tcn_t* h);
int rfx_tcn_load_param(rfx_tcn_t* h, const char* key, const float* src, int64_t numel, void* stream);
int rfx_tcn_finalize(rfx_tcn_t* h, void* stream);
long long rfx_tcn_out_length(const rfx_tcn_t* h, long long T);
size_t rfx_tcn_workspace_bytes(const rfx_tcn_t* h, int B, long long T);
int rfx_tcn_forward(rfx_tcn_t* h, const float* x, int B, long long T, float* out, void* workspace, size_t workspace_bytes, void* stream);
int rfx_tcn_launches_per_call(const rfx_tcn_t* h);

/* Training path of the TCN (config 5's step for this network: remfx/models.py:217-220 -> loss.backward() differentiates
 * remfx/tcn.py:48-59,126-130 through torch autograd; here the same gradients come from hand-written kernels).
 *   rfx_tcn_forward_train  = rfx_tcn_forward that keeps every block's output in `workspace` (split-bf16 planes);
 *   rfx_tcn_backward       = dL/d(parameters) from dL/d(out): `x`, `out` and `workspace` are those of the matching
 *                            rfx_tcn_forward_train call; `dout` is (B, 1, rfx_tcn_out_length(T)) fp32.
 * keys[i] names a parameter as in rfx_tcn_load_param and grads[i] is a device buffer of that parameter's size; every
 * parameter of the network must be listed; each buffer is OVERWRITTEN with the gradient (summed over the batch).
 * Sums over time use fp32 atomics: reproducible to rounding, not bit-for-bit. */
size_t rfx_tcn_train_workspace_bytes(const rfx_tcn_t* h, int B, long long T);
int rfx_tcn_forward_train(rfx_tcn_t* h, const float* x, int B, long long T, float* out, void* workspace, size_t workspace_bytes, void* stream);
int rfx_tcn_backward(rfx_tcn_t* h, const float* x, const float* out, const float* dout, int B, long long T, const char* const* keys,
                     float* const* grads, int nkeys, void* workspace, size_t workspace_bytes, void* stream);
int rfx_tcn_backward_launches_per_call(const rfx_tcn_t* h);
/* Weight-gradient kernel selector, process-wide: 0 = the product path (tcgen05 bf16x3 with MN-major operands when the channel
 * width is 256 as in cfg/model/tcn.yaml, else mma.sync bf16x3), 1 = plain fp32 FFMA kernel (slow; cross-check in tests only),
 * 2 = mma.sync bf16x3 for every width. */
int rfx_tcn_set_wgrad_impl(int impl);

/* ---------------------------------------------------------------------------------------------
 * C1-C4  Cnn14 effect classifier (eval mode)
 *   replaces remfx/classifier.py:193-233 (Cnn14.forward) as called by FXClassifier.forward
 *   (remfx/models.py:490-491) and the cascade (remfx/models.py:63).
 * Parameter keys are the reference state_dict names ("conv_block3.bn2.running_var", "melspec.mel_scale.fb",
 * "heads.4.bias", ...).  x: (B, 1, T) fp32 device.  probs: (B, num_classes) sigmoid outputs (row-major, i.e.
 * torch.hstack of the reference's list); logits (optional, may be NULL): pre-sigmoid values.
 * The model and input sample rates must be equal (no resampler on this path).
 * ------------------------------------------------------------------------------------------- */
typedef struct rfx_cnn14 rfx_cnn14_t;

typedef struct {
  int num_classes; /* 5 */
  int n_fft;       /* 2048 */
  int hop;         /* 512 */
  int n_mels;      /* 128 */
} rfx_cnn14_config;

int rfx_cnn14_create(const rfx_cnn14_config* cfg, rfx_cnn14_t** out);
void rfx_cnn14_destroy(rfx_cnn14_t* h);
int rfx_cnn14_load_param(rfx_cnn14_t* h, const char* key, const float* src, int64_t numel, void* stream);
int rfx_cnn14_finalize(rfx_cnn14_t* h, void* stream);
size_t rfx_cnn14_workspace_bytes(const rfx_cnn14_t* h, int B, int T);
int rfx_cnn14_forward(rfx_cnn14_t* h, const float* x, int B, int T, float* probs, float* logits, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * D1-D10  Hybrid Demucs effect-removal model
 *   replaces remfx/models.py:308-324 (DemucsModel.forward / sample) = torchaudio.models.HDemucs.forward
 *   (torchaudio/models/_hdemucs.py:523-634, the un-vendored third-party module the reference imports).
 * Parameter keys are the HDemucs state_dict names ("freq_encoder.2.dconv.layers.1.3.weight", ...) plus
 * "__window__" = hann_window(nfft).  x: (B, 1, T) fp32 device -> out: (B, 1, T) fp32 device; T % 1024 == 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct rfx_hdemucs rfx_hdemucs_t;

typedef struct {
  int audio_channels; /* 1 */
  int n_sources;      /* 1 (sources = ["mixture"]) */
  int channels;       /* 48 */
  int growth;         /* 2 */
  int nfft;           /* 4096 */
  int depth;          /* 6 */
  int kernel_size;    /* 8 */
  int stride;         /* 4 */
  int time_stride;    /* 2 */
  int context;        /* 1 */
  int context_enc;    /* 0 */
  int norm_starts;    /* 4 */
  int norm_groups;    /* 4 */
  int dconv_depth;    /* 2 */
  int dconv_comp;     /* 4 */
  int dconv_attn;     /* 4 */
  int dconv_lstm;     /* 4 */
  float freq_emb_weight; /* 0.2 (0 disables the frequency embedding) */
  float freq_emb_scale;  /* 10 */
} rfx_hdemucs_config;

int rfx_hdemucs_create(const rfx_hdemucs_config* cfg, rfx_hdemucs_t** out);
void rfx_hdemucs_destroy(rfx_hdemucs_t* h);
int rfx_hdemucs_load_param(rfx_hdemucs_t* h, const char* key, const float* src, int64_t numel, void* stream);
int rfx_hdemucs_finalize(rfx_hdemucs_t* h, void* stream);
size_t rfx_hdemucs_workspace_bytes(rfx_hdemucs_t* h, int B, int T);
int rfx_hdemucs_forward(rfx_hdemucs_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                        void* stream);
int rfx_hdemucs_launches_per_call(rfx_hdemucs_t* h, int B, int T);
/* Debug taps: when enabled, the next forward records named intermediate activations; rfx_hdemucs_tap copies one
 * as fp32 in (B, Y, X, C) channel-last order (dst may be NULL to query dims[4] only). */
int rfx_hdemucs_set_taps(rfx_hdemucs_t* h, int on);
int rfx_hdemucs_tap(rfx_hdemucs_t* h, const char* name, float* dst, int64_t capacity, int* dims, void* stream);

/* Training (row L5 with Hybrid Demucs, the network of cfg/exp/5-5_full.yaml:3): what `loss.backward()` does to
 * `DemucsModel.forward`'s output in the reference's Lightning step (remfx/models.py:217-221, 317-321).
 *   rfx_hdemucs_forward_train  the forward with every pre-activation kept in `workspace` and a tape of its ops in the handle;
 *                              same result as rfx_hdemucs_forward to rounding (activations are un-fused from the GEMM epilogues)
 *   rfx_hdemucs_backward       given dout = dLoss / dout (B, T), writes dLoss / dparameter for every state_dict key passed in
 *                              keys[] / grads[] (device fp32 buffers of the parameter's size; all are overwritten).  Must be
 *                              called with the SAME workspace pointer as the matching forward_train.  No gradient is produced for x.
 * workspace: rfx_hdemucs_train_workspace_bytes(h, B, T) bytes (forward + backward scratch), 256-byte aligned. */
size_t rfx_hdemucs_train_workspace_bytes(rfx_hdemucs_t* h, int B, int T);
int rfx_hdemucs_forward_train(rfx_hdemucs_t* h, const float* x, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                              void* stream);
int rfx_hdemucs_backward(rfx_hdemucs_t* h, const float* x, const float* dout, int B, int T, const char* const* keys,
                         float* const* grads, int nkeys, void* workspace, size_t workspace_bytes, void* stream);
/* Debug: after a backward, the gradient of a tapped activation (fp32, (B, Y, X, C)); and a way to substitute a reference
 * gradient at a tap during the following backward calls so that each layer can be judged on its own (grad NULL removes it). */
/* Weight-gradient contraction of the Hybrid-Demucs backward, process-wide: 0 = tcgen05 with MN-major TMA-staged operands (default),
 * 1 = the mma.sync tile variants (cross-check in tests). */
int rfx_hdemucs_set_wgrad_impl(int impl);
/* on != 0: the caller guarantees that the gradient buffers of the following rfx_hdemucs_backward calls are already zero (views of
 * one freshly zeroed flat bucket -- what remfx_b200.optim.FlatBucket hands out), so the backward skips its per-parameter memsets.
 * Replaces nothing in the reference: torch autograd allocates each parameter gradient itself (remfx/models.py:217-220). */
int rfx_hdemucs_set_grads_prezeroed(rfx_hdemucs_t* h, int on);
int rfx_hdemucs_grad_tap(rfx_hdemucs_t* h, const char* name, float* dst, int64_t capacity, int* dims, void* stream);
int rfx_hdemucs_inject_grad(rfx_hdemucs_t* h, const char* name, const float* grad);

/* ---------------------------------------------------------------------------------------------
 * L1/L2  RemFx loss = MultiResolutionSTFTLoss(out, target) + l1_weight * mean|out - target|
 *   replaces `self.mrstftloss(out, target) + self.l1loss(out, target) * 100`
 *   (remfx/models.py:299,320,385; auraloss.freq.MultiResolutionSTFTLoss defaults, see oracle/loss.py)
 * out/target: (B, T) with the given batch strides (in floats), so a cropped view of a longer target can
 * be passed directly.  win1024/win2048/win512: the hann windows of length 600/1200/240 zero-padded
 * (centred) to n_fft, as torch.stft does.  result (device, 9 floats): [0] loss, [1] MR-STFT term,
 * [2] mean|out-target|, [3+2r] spectral-convergence and [4+2r] log-magnitude term of resolution r.
 * ------------------------------------------------------------------------------------------- */
size_t rfx_loss_workspace_bytes(int B, int T);
int rfx_remfx_loss(const float* out, long long out_bstride, const float* target, long long target_bstride, int B, int T,
                   const float* win1024, const float* win2048, const float* win512, float l1_weight, float* result,
                   void* workspace, size_t workspace_bytes, void* stream);
/* Gradient of that loss with respect to `out` (the first link of the training step, remfx/models.py:299 under autograd):
 * grad_out (B rows, stride grad_bstride) = *grad_loss * dL/d out.  `workspace` must be the one rfx_remfx_loss just ran in on the
 * same (out, target): it holds the per-item spectral norms.  Overlapping frames accumulate with fp32 atomic adds (the last bits
 * of the gradient vary from run to run). */
int rfx_remfx_loss_backward(const float* out, long long out_bstride, const float* target, long long target_bstride, int B, int T,
                            const float* win1024, const float* win2048, const float* win512, float l1_weight,
                            const float* grad_loss, float* grad_out, long long grad_bstride, const void* workspace,
                            size_t workspace_bytes, void* stream);

/* L3  SI-SDR metric: auraloss.time.SISDRLoss() as constructed at remfx/models.py:41,173 (zero_mean, eps 1e-8, mean
 * over the batch; returns the NEGATIVE SI-SDR in dB -- the reference negates it again when logging, models.py:230-233).
 * x = estimate, y = target, (B, T) with batch strides in floats.  result: 1 float (device). */
size_t rfx_sisdr_workspace_bytes(int B);
int rfx_sisdr_loss(const float* x, long long x_bstride, const float* y, long long y_bstride, int B, int T, float* result,
                   void* workspace, size_t workspace_bytes, void* stream);

/* L5  optimiser half of the training step: torch.optim.AdamW as configured at remfx/models.py:185-191 (lr 1e-4, betas
 * (0.95, 0.999), eps 1e-6, weight_decay 1e-3) + Lightning's `gradient_clip_val: 10.0` (cfg/config.yaml:119 =
 * torch.nn.utils.clip_grad_norm_) over ONE flat fp32 bucket holding every parameter (16-byte aligned, same length for
 * param / grad / exp_avg / exp_avg_sq).
 *   rfx_grad_sumsq : workspace[0] (double) (+)= sum(grad^2); accumulate != 0 adds to the previous value (several buckets)
 *   rfx_adamw_step : g' = grad * grad_scale * min(1, max_norm / (sqrt(sumsq) * grad_scale + 1e-6)) (no clipping when
 *                    max_norm <= 0), then the decoupled-decay Adam update with bias correction for the 1-based `step`;
 *                    grad_scale = 1 / world_size after a SUM all-reduce of the bucket; total_norm (device, optional)
 *                    receives the pre-clip norm, as clip_grad_norm_ returns it.  The learning rate is an argument so the
 *                    host applies the MultiStepLR schedule (models.py:192-196). */
size_t rfx_optim_workspace_bytes(void);
int rfx_grad_sumsq(const float* grad, long long n, void* workspace, int accumulate, void* stream);
int rfx_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, float max_norm, const void* workspace, float* total_norm,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * N2  Batch ingest (the step in front of the hot path)
 *   replaces remfx/datasets.py:461-468 (EffectDataset.__getitem__: two torchaudio.load per item) + the DataLoader collate
 * Host-side, no GPU work: decodes n mono RIFF/WAVE files (IEEE float 32/64, integer PCM 8/16/24/32, plain or
 * WAVE_FORMAT_EXTENSIBLE headers, unknown chunks skipped; integer samples scaled like torchaudio.load(normalize=True)) with a
 * pool of n_threads threads straight into row i of dst_host[n][T] -- meant to be one pinned (B, 1, T) buffer that
 * rfx_umx_pipe_push / rfx_umx_sample_host consume.  Files shorter than T are zero-padded, longer ones cut; frames[i] /
 * sample_rates[i] (optional) report what each file holds.  Returns 2 and sets rfx_last_error() on the first unreadable file.
 * ------------------------------------------------------------------------------------------- */
int rfx_wav_info(const char* path, int* sample_rate, int* channels, long long* frames, int* format_tag, int* bits);
int rfx_ingest_wav_batch(const char* const* paths, int n, float* dst_host, long long T, int n_threads, long long* frames,
                         int* sample_rates);

#ifdef __cplusplus
}
#endif
#endif /* REMFX_B200_H */
