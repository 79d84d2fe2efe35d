// Bidirectional LSTM recurrence (torch.nn.LSTM semantics, gate order i,f,g,o, zero initial state) for the
// Open-Unmix 3-layer BiLSTM (umx/openunmix/model.py:62-69,141) -- the latency-critical part of the path:
// 3 layers x 513 strictly dependent steps.
//
// The input projections W_ih x + b_ih + b_hh of all time steps are one big tensor-core GEMM (gemm.cu);
// this kernel only runs the recurrent part  g_t = G_t + W_hh h_{t-1}.  Mapping:
//   * one thread-block CLUSTER of 8 CTAs per (direction, group of NB = 4 batch items);
//     grid = 8 x ceil(B/4) x 2  (= 128 CTAs at B = 32: one per SM, both directions concurrently);
//   * CTA rank r owns hidden units [32 r, 32 r + 32) = 128 gate rows of W_hh, held ENTIRELY IN REGISTERS
//     (128 fp32 per thread: row = tid % 128, K-half = tid / 128) for the whole sequence, so a step reads no
//     weights from memory at all;
//   * h_{t-1} (NB x 256 fp32) lives in every CTA's shared memory (double-buffered); after the gate math
//     the 32 new h values per batch item are pushed to all 8 CTAs with st.shared::cluster (DSMEM) and one
//     cluster barrier per step orders the exchange -- no global-memory round trip, no grid sync.
// fp32 FFMA throughout (the recurrence amplifies rounding over 513 steps; bf16 would fail the 1e-4 gate).
#include "kernels.h"

namespace rfx {

constexpr int LSTM_H = 256;
constexpr int LSTM_CL = 8;                  // CTAs per cluster
constexpr int LSTM_UPC = LSTM_H / LSTM_CL;  // hidden units per CTA (32)
constexpr int LSTM_ROWS = 4 * LSTM_UPC;     // gate rows per CTA (128)
constexpr int LSTM_NB = 4;                  // batch items per cluster
constexpr int LSTM_KH = LSTM_H / 2;         // K elements per thread (128)

__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(256, 1)
    lstm_rec_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Whh, float* __restrict__ Hout, int ldh, int B, int F) {
  __shared__ __align__(16) float h_buf[2][LSTM_NB][LSTM_H];
  __shared__ float part[2][LSTM_NB][LSTM_ROWS];

  const int tid = threadIdx.x;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.z;
  const int b0 = blockIdx.y * LSTM_NB;
  const int row_local = tid & (LSTM_ROWS - 1);  // gate * 32 + unit
  const int khalf = tid >> 7;
  const int gate = row_local >> 5, unit = row_local & 31;

  // W_hh slice -> registers (one-time, 512 B per thread)
  float w[LSTM_KH];
  {
    const float* wrow = Whh + ((size_t)dir * 4 * LSTM_H + (size_t)gate * LSTM_H + rank * LSTM_UPC + unit) * LSTM_H + khalf * LSTM_KH;
#pragma unroll
    for (int i = 0; i < LSTM_KH; i += 4) {
      const float4 t = *reinterpret_cast<const float4*>(wrow + i);
      w[i] = t.x; w[i + 1] = t.y; w[i + 2] = t.z; w[i + 3] = t.w;
    }
  }
  for (int i = tid; i < 2 * LSTM_NB * LSTM_H; i += 256) (&h_buf[0][0][0])[i] = 0.0f;

  // finalize role (threads 0..127): batch item fb, unit fu; cell state lives in a register
  const int fb = tid >> 5, fu = tid & 31;
  const bool fin = tid < LSTM_NB * LSTM_UPC;
  const bool fvalid = fin && (b0 + fb < B);
  float c_state = 0.0f;
  const size_t gcol = (size_t)dir * 4 * LSTM_H + rank * LSTM_UPC + fu;  // + gate * H
  float gin[4] = {0.f, 0.f, 0.f, 0.f};
  if (fvalid) {
    const int tt0 = dir ? F - 1 : 0;
    const float* g = G + ((size_t)(b0 + fb) * F + tt0) * ldg + gcol;
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = g[q * LSTM_H];
  }
  // remote addresses of h_buf[0][fb][rank*32 + fu] in every CTA of the cluster
  uint32_t remote[LSTM_CL];
  {
    const uint32_t local = smem_u32(&h_buf[0][fb & (LSTM_NB - 1)][rank * LSTM_UPC + fu]);
#pragma unroll
    for (int r = 0; r < LSTM_CL; ++r) remote[r] = mapa_u32(local, r);
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();

  for (int step = 0; step < F; ++step) {
    const int cur = step & 1;
    const int tt = dir ? F - 1 - step : step;
    // prefetch next step's input-projection values (independent of the recurrence)
    float gnext[4] = {0.f, 0.f, 0.f, 0.f};
    if (fvalid && step + 1 < F) {
      const int tn = dir ? tt - 1 : tt + 1;
      const float* g = G + ((size_t)(b0 + fb) * F + tn) * ldg + gcol;
#pragma unroll
      for (int q = 0; q < 4; ++q) gnext[q] = g[q * LSTM_H];
    }
    // partial dot products over this thread's K half for the NB batch items
    float acc[LSTM_NB];
#pragma unroll
    for (int b = 0; b < LSTM_NB; ++b) acc[b] = 0.0f;
    const float* hb = &h_buf[cur][0][khalf * LSTM_KH];
#pragma unroll
    for (int i = 0; i < LSTM_KH; i += 4) {
#pragma unroll
      for (int b = 0; b < LSTM_NB; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hb + b * LSTM_H + i);
        acc[b] = fmaf(w[i], hv.x, acc[b]);
        acc[b] = fmaf(w[i + 1], hv.y, acc[b]);
        acc[b] = fmaf(w[i + 2], hv.z, acc[b]);
        acc[b] = fmaf(w[i + 3], hv.w, acc[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < LSTM_NB; ++b) part[khalf][b][row_local] = acc[b];
    __syncthreads();
    if (fin) {
      const float pi = part[0][fb][0 * LSTM_UPC + fu] + part[1][fb][0 * LSTM_UPC + fu] + gin[0];
      const float pf = part[0][fb][1 * LSTM_UPC + fu] + part[1][fb][1 * LSTM_UPC + fu] + gin[1];
      const float pg = part[0][fb][2 * LSTM_UPC + fu] + part[1][fb][2 * LSTM_UPC + fu] + gin[2];
      const float po = part[0][fb][3 * LSTM_UPC + fu] + part[1][fb][3 * LSTM_UPC + fu] + gin[3];
      const float ig = sigmoidf_acc(pi), fg = sigmoidf_acc(pf), gg = tanhf(pg), og = sigmoidf_acc(po);
      c_state = fg * c_state + ig * gg;
      const float h = og * tanhf(c_state);
      if (fvalid) Hout[((size_t)(b0 + fb) * F + tt) * ldh + dir * LSTM_H + rank * LSTM_UPC + fu] = h;
      const uint32_t boff = (uint32_t)((cur ^ 1) * LSTM_NB * LSTM_H * sizeof(float));
#pragma unroll
      for (int r = 0; r < LSTM_CL; ++r) st_cluster_f32(remote[r] + boff, h);
#pragma unroll
      for (int q = 0; q < 4; ++q) gin[q] = gnext[q];
    }
    // orders: DSMEM writes of h_t (release) before anyone reads them (acquire); also protects part[]
    cluster_arrive();
    cluster_wait();
  }
}

int launch_lstm_layer(const float* G, int ldg, const float* Whh, float* Hout, int ldh, int B, int F, int H, cudaStream_t stream) {
  RFX_REQUIRE(H == LSTM_H, "lstm: hidden size per direction must be 256");
  RFX_REQUIRE(B > 0 && F > 0, "lstm: positive sizes");
  RFX_REQUIRE(((uintptr_t)Whh & 15) == 0, "lstm: W_hh must be 16-byte aligned");
  dim3 grid(LSTM_CL, ceil_div(B, LSTM_NB), 2);
  lstm_rec_kernel<<<grid, 256, 0, stream>>>(G, ldg, Whh, Hout, ldh, B, F);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rfx
