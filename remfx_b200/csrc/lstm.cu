// Bidirectional LSTM recurrence (torch.nn.LSTM semantics, gate order i,f,g,o, zero initial state) for the
// Open-Unmix 3-layer BiLSTM (umx/openunmix/model.py:62-69,141) -- the latency-critical part of the path:
// 3 layers x 513 strictly dependent steps.
//
// The input projections W_ih x + b_ih + b_hh of all time steps are one big tensor-core GEMM (gemm.cu);
// this kernel only runs the recurrent part  g_t = G_t + W_hh h_{t-1}.  Mapping:
//   * one thread-block CLUSTER of 8 CTAs per (direction, group of NB batch items), NB in 4..8 chosen so
//     that every cluster of the launch is co-resident (B = 32 -> NB = 5: 14 clusters = 112 SMs, both
//     directions concurrently);
//   * CTA rank r owns hidden units [32 r, 32 r + 32); warp w of the CTA owns units 4w..4w+3, i.e. 16 gate
//     rows of W_hh, held ENTIRELY IN REGISTERS for the whole sequence (lane = unit-in-warp x K-slice:
//     4 gates x 32 K-values = 128 fp32 per thread), so a step reads no weights from memory at all;
//   * h_{t-1} (NB x 256 fp32) lives in every CTA's shared memory, double-buffered, laid out
//     [source CTA][batch][32 units] so that each CTA's contribution is one contiguous 512-byte block.
//     A step is: 128 NB FFMA/thread against float4 LDS of h, a 28-shuffle reduce-scatter over the 8
//     K-slices (lane ks ends up owning batch slot ks), gate math, the new h values staged in local smem,
//     one __syncthreads, then
//     8 DSMEM BULK copies (cp.async.bulk shared::cta -> shared::cluster, 512 B each) push them to all 8
//     CTAs with mbarrier complete_tx signalling -- 8 remote transactions per CTA per step instead of
//     hundreds of small remote stores (measured: fine-grained st.async/st.shared::cluster exchange cost
//     ~7 k cycles per step).  No cluster barrier and no memory fence in the loop: the only cross-CTA wait
//     is the mbarrier counting the 4 KB of h_t arriving; double buffering + data dependence make the
//     buffers hazard-free.
// fp32 FFMA throughout (the recurrence amplifies rounding over 513 steps; bf16 would fail the 1e-4 gate).
#include "kernels.h"

#include <cstdlib>

namespace rfx {

constexpr int LSTM_H = 256;
constexpr int LSTM_CL = 8;                  // CTAs per cluster
constexpr int LSTM_UPC = LSTM_H / LSTM_CL;  // hidden units per CTA (32)
constexpr int LSTM_WARPS = 8;

// local shared memory -> (possibly remote) shared memory of a CTA in the cluster; completion on the
// destination CTA's mbarrier.
__device__ __forceinline__ void bulk_s2cluster(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes, uint32_t cluster_mbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster_addr),
               "r"(src_cta_addr), "r"(bytes), "r"(cluster_mbar)
               : "memory");
}

// NB = batch items per cluster (4..8).  All clusters of a launch must be co-resident (only ~15 clusters of
// 8 CTAs fit on a B200 at once: a 16th would run as a second wave and double the time), so the host picks
// the smallest NB for which 2 * ceil(B / NB) clusters fit (cudaOccupancyMaxActiveClusters).
template <int NB>
__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(LSTM_WARPS * 32, 1)
    lstm_rec_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Whh, float* __restrict__ Hout, int ldh,
                    __nv_bfloat16* __restrict__ Hhi, __nv_bfloat16* __restrict__ Hlo, int ldhs, int B, int F) {
  constexpr int TX_BYTES = LSTM_CL * NB * LSTM_UPC * 4;  // bytes of h_t every CTA receives per step
  __shared__ __align__(128) float h_buf[2][LSTM_CL][NB][LSTM_UPC];  // [buffer][source CTA][batch][unit]
  __shared__ __align__(128) float stage[2][NB][LSTM_UPC];           // this CTA's new h values
  __shared__ __align__(8) uint64_t h_bar[2];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.z;
  const int b0 = blockIdx.y * NB;
  const int rg = lane >> 3;  // unit within the warp (0..3)
  const int ks = lane & 7;   // K slice during the mat-vec; batch slot after the reduce-scatter
  const int unit = rank * LSTM_UPC + warp * 4 + rg;  // hidden unit this lane group works on

  // ---- W_hh rows of the 4 gates of `unit`, K-slice ks: k = 4*(8c + ks) + e, c = 0..7, e = 0..3 ----
  float w[4][32];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float* wrow = Whh + ((size_t)dir * 4 * LSTM_H + (size_t)g * LSTM_H + unit) * LSTM_H;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 t = *reinterpret_cast<const float4*>(wrow + 4 * (8 * c + ks));
      w[g][4 * c + 0] = t.x; w[g][4 * c + 1] = t.y; w[g][4 * c + 2] = t.z; w[g][4 * c + 3] = t.w;
    }
  }
  for (int i = tid; i < 2 * LSTM_CL * NB * LSTM_UPC; i += LSTM_WARPS * 32) (&h_buf[0][0][0][0])[i] = 0.0f;
  if (tid == 0) {
    mbar_init(&h_bar[0], 1);
    mbar_init(&h_bar[1], 1);
    mbar_fence_init();
  }

  // After the reduce-scatter lane (rg, ks) holds all four gate pre-activations of (unit, batch slot ks).
  const bool bslot = ks < NB;
  const bool bvalid = bslot && (b0 + ks) < B;
  const size_t gcol = (size_t)dir * 4 * LSTM_H + unit;  // + gate * H
  float c_state = 0.0f;
  float gin[4] = {0.f, 0.f, 0.f, 0.f};
  if (bvalid) {
    const float* g = G + ((size_t)(b0 + ks) * F + (dir ? F - 1 : 0)) * ldg + gcol;
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = g[q * LSTM_H];
  }
  // Sender role (warp 0, lanes 0..7): lane d bulk-copies this CTA's block to CTA d.
  const uint32_t dst_h = mapa_u32(smem_u32(&h_buf[0][rank][0][0]), lane & 7);
  const uint32_t dst_bar = mapa_u32(smem_u32(&h_bar[0]), lane & 7);
  // Global writer role (threads 0 .. NB*32-1): batch item tid >> 5, unit tid & 31.
  const int wb = tid >> 5, wu = tid & 31;
  const bool wvalid = wb < NB && (b0 + wb) < B;

  __syncthreads();
  cluster_arrive();  // barriers initialised and h_buf zeroed everywhere before anyone sends
  cluster_wait();

  for (int step = 0; step < F; ++step) {
    const int cur = step & 1;
    const int tt = dir ? F - 1 - step : step;
    if (tid == 0) mbar_arrive_expect_tx(&h_bar[cur ^ 1], TX_BYTES);  // phase that will receive h_t
    // prefetch next step's input projections (independent of the recurrence)
    float gn[4] = {0.f, 0.f, 0.f, 0.f};
    if (bvalid && step + 1 < F) {
      const float* g = G + ((size_t)(b0 + ks) * F + (dir ? tt - 1 : tt + 1)) * ldg + gcol;
#pragma unroll
      for (int q = 0; q < 4; ++q) gn[q] = g[q * LSTM_H];
    }
    if (step > 0) mbar_wait(&h_bar[cur], ((step - 1) >> 1) & 1);  // all of h_{t-1} has landed

    float acc[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[g][b] = 0.0f;
    const float* hb = &h_buf[cur][0][0][4 * ks];  // k = 32 c + 4 ks + e lives at [c][b][4 ks + e]
#pragma unroll
    for (int c = 0; c < 8; ++c) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(hb + c * (NB * LSTM_UPC) + b * LSTM_UPC);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          acc[g][b] = fmaf(w[g][4 * c + 0], hv.x, acc[g][b]);
          acc[g][b] = fmaf(w[g][4 * c + 1], hv.y, acc[g][b]);
          acc[g][b] = fmaf(w[g][4 * c + 2], hv.z, acc[g][b]);
          acc[g][b] = fmaf(w[g][4 * c + 3], hv.w, acc[g][b]);
        }
      }
    }
    // ---- reduce-scatter over the 8 K-slices (lanes 8 rg .. 8 rg + 7): lane ks ends with batch slot ks ----
    float r1[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const float send = (ks & 4) ? acc[g][bb] : acc[g][bb + 4];
        const float keep = (ks & 4) ? acc[g][bb + 4] : acc[g][bb];
        r1[g][bb] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
    float r2[4][2];
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const float send = (ks & 2) ? r1[g][bb] : r1[g][bb + 2];
        const float keep = (ks & 2) ? r1[g][bb + 2] : r1[g][bb];
        r2[g][bb] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
    float pre[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float send = (ks & 1) ? r2[g][0] : r2[g][1];
      const float keep = (ks & 1) ? r2[g][1] : r2[g][0];
      pre[g] = keep + __shfl_xor_sync(0xffffffffu, send, 1) + gin[g];
    }
    // ---- gate math (i, f, g, o) for (unit, batch slot ks) ----
    if (bslot) {
      const float ig = sigmoidf_acc(pre[0]), fg = sigmoidf_acc(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_acc(pre[3]);
      c_state = fg * c_state + ig * gg;
      stage[cur ^ 1][ks][warp * 4 + rg] = og * tanhf(c_state);
    }
    // ---- 8 bulk DSMEM copies per CTA push the new h block to every CTA of the cluster ----
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy (async proxy) reads
    __syncthreads();
    if (warp == 0 && lane < LSTM_CL) {
      const uint32_t boff = (uint32_t)((cur ^ 1) * LSTM_CL * NB * LSTM_UPC * sizeof(float));
      bulk_s2cluster(dst_h + boff, smem_u32(&stage[cur ^ 1][0][0]), NB * LSTM_UPC * sizeof(float),
                     dst_bar + (uint32_t)((cur ^ 1) * sizeof(uint64_t)));
    }
    if (wvalid) {
      const float hv = stage[cur ^ 1][wb][wu];
      const size_t row = (size_t)(b0 + wb) * F + tt;
      const int col = dir * LSTM_H + rank * LSTM_UPC + wu;
      if (Hout) Hout[row * ldh + col] = hv;
      if (Hhi) {  // the next tensor-core layer consumes split-bf16 planes
        __nv_bfloat16 h, l;
        split_bf16(hv, h, l);
        Hhi[row * ldhs + col] = h;
        Hlo[row * ldhs + col] = l;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = gn[q];
  }
  // Nobody may exit while peers can still write into its shared memory / barriers: wait for the last h to land.
  mbar_wait(&h_bar[F & 1], ((F - 1) >> 1) & 1);
  cluster_arrive();
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------------
// Tensor-core variant of the recurrence: the per-step mat-vec W_hh h_{t-1} runs on mma.sync.m16n8k16 (bf16,
// fp32 accumulate) with the bf16x3 split (W_lo*h_hi + W_hi*h_lo + W_hi*h_hi), so one warp needs 48 HMMA
// instead of ~640 FFMA + 28 shuffles per step.  Each warp owns 16 gate rows (4 units x {i,f,g,o}) as the M
// dimension, the 8 batch slots of the cluster are the N dimension, K = 256.  W_hh hi/lo fragments live in
// registers for the whole sequence; h is exchanged between the 8 CTAs as split-bf16 (hi, lo) -- exactly the
// B-fragment format -- laid out [source CTA][plane][slot][32 units] so that (i) every CTA's contribution is
// one contiguous 1 KB block (one DSMEM bulk copy per destination) and (ii) a lane's B fragments for two
// K-steps are one conflict-free LDS.128 (K is permuted consistently in the A fragments).
// Gate math uses ex2-based sigmoid/tanh (abs error ~1e-7), c and the emitted h stay fp32.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hmma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_hi2(float x, float y, float& rx, float& ry) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h2);
  rx = x - hf.x;
  ry = y - hf.y;
  return *reinterpret_cast<const uint32_t*>(&h2);
}
__device__ __forceinline__ uint32_t pack2(float x, float y) {
  const __nv_bfloat162 h2 = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h2);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

constexpr int LSTM_SLOTS = 8;  // batch slots per MMA n-tile

// H = hidden units per direction (multiple of 32 with H/8 a multiple of 8: 192, 256, 384).  Each of the 8 CTAs owns
// UPC = H/8 units (UPC/4 warps).  The W_hh hi fragments always live in registers (H/16 k-steps x 4 regs); when they
// would not both fit (H = 384, or two n-tiles) the lo fragments are kept in shared memory instead (LO_SMEM) and re-read
// every step.  NT = MMA n-tiles per cluster: the cluster serves 8 NT batch slots.  The legacy HMMA pipe is the busiest
// unit of a step (its time grows with NT), but a cluster of NT = 2 does twice the work on the same 8 SMs, which is what
// the multi-lane pipeline wants: two recurrence launches then fit side by side.
template <int H, bool LO_SMEM, int NT>
__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(H / 8 / 4 * 32, 1)
    lstm_rec_mma_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Whh, float* __restrict__ Hout, int ldh,
                        __nv_bfloat16* __restrict__ Hhi, __nv_bfloat16* __restrict__ Hlo, int ldhs, int B, int F, int NB) {
  constexpr int UPC = H / LSTM_CL;          // units per CTA
  constexpr int WARPS = UPC / 4;
  constexpr int THREADS = WARPS * 32;
  constexpr int KS = H / 16;                // MMA k-steps
  constexpr int NP = H / 32;                // k-step pairs (one LDS.128 of h per plane each)
  constexpr int SLOTS = LSTM_SLOTS * NT;
  constexpr int BLK_BYTES = 2 * SLOTS * UPC * 2;  // one CTA's h block: 2 planes x SLOTS x UPC units bf16
  constexpr int TX = LSTM_CL * BLK_BYTES;
  static_assert(UPC % 8 == 0 && H % 32 == 0, "unsupported hidden size");
  extern __shared__ __align__(128) uint8_t lstm_smem[];
  __nv_bfloat16* h_buf = reinterpret_cast<__nv_bfloat16*>(lstm_smem);                         // [2][CL][2][SLOTS][UPC]
  __nv_bfloat16* stage = h_buf + 2 * LSTM_CL * 2 * SLOTS * UPC;                               // [2][2][SLOTS][UPC]
  uint64_t* h_bar = reinterpret_cast<uint64_t*>(stage + 2 * 2 * SLOTS * UPC);                 // [2]
  uint4* alo_smem = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(h_bar) + 128);         // [KS][THREADS] (LO_SMEM only)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.z;
  const int b0 = blockIdx.y * NB;
  const int gid = lane >> 2, tig = lane & 3;
  const int u = gid >> 1, pp = gid & 1;  // unit within the warp; pp = 0: rows (i, g), pp = 1: rows (f, o)
  const int unit = rank * UPC + warp * 4 + u;
  const int gate0 = pp ? 1 : 0, gate1 = pp ? 3 : 2;

  // ---- A fragments: W_hh rows (gate0, unit) and (gate1, unit), true k = 32 P + 8 tig + [0, 8) for P = 0..NP-1 ----
  uint32_t a_hi[KS][4];
  uint32_t a_lo[LO_SMEM ? 1 : KS][4];
  {
    const float* w0 = Whh + ((size_t)dir * 4 * H + (size_t)gate0 * H + unit) * H + 8 * tig;
    const float* w1 = Whh + ((size_t)dir * 4 * H + (size_t)gate1 * H + unit) * H + 8 * tig;
#pragma unroll
    for (int P = 0; P < NP; ++P) {
      const float4 x0 = *reinterpret_cast<const float4*>(w0 + 32 * P), x1 = *reinterpret_cast<const float4*>(w0 + 32 * P + 4);
      const float4 y0 = *reinterpret_cast<const float4*>(w1 + 32 * P), y1 = *reinterpret_cast<const float4*>(w1 + 32 * P + 4);
      float rx, ry;
      uint32_t lo0[4], lo1[4];
      // K-step 2P: slots (2tig, 2tig+1) <- k+0,1 ; slots (2tig+8, +9) <- k+2,3.  K-step 2P+1: k+4,5 ; k+6,7.
      a_hi[2 * P][0] = pack_hi2(x0.x, x0.y, rx, ry); lo0[0] = pack2(rx, ry);
      a_hi[2 * P][1] = pack_hi2(y0.x, y0.y, rx, ry); lo0[1] = pack2(rx, ry);
      a_hi[2 * P][2] = pack_hi2(x0.z, x0.w, rx, ry); lo0[2] = pack2(rx, ry);
      a_hi[2 * P][3] = pack_hi2(y0.z, y0.w, rx, ry); lo0[3] = pack2(rx, ry);
      a_hi[2 * P + 1][0] = pack_hi2(x1.x, x1.y, rx, ry); lo1[0] = pack2(rx, ry);
      a_hi[2 * P + 1][1] = pack_hi2(y1.x, y1.y, rx, ry); lo1[1] = pack2(rx, ry);
      a_hi[2 * P + 1][2] = pack_hi2(x1.z, x1.w, rx, ry); lo1[2] = pack2(rx, ry);
      a_hi[2 * P + 1][3] = pack_hi2(y1.z, y1.w, rx, ry); lo1[3] = pack2(rx, ry);
      if (LO_SMEM) {
        alo_smem[(2 * P) * THREADS + tid] = make_uint4(lo0[0], lo0[1], lo0[2], lo0[3]);
        alo_smem[(2 * P + 1) * THREADS + tid] = make_uint4(lo1[0], lo1[1], lo1[2], lo1[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { a_lo[LO_SMEM ? 0 : 2 * P][q] = lo0[q]; a_lo[LO_SMEM ? 0 : 2 * P + 1][q] = lo1[q]; }
      }
    }
  }
  for (int i = tid; i < 2 * LSTM_CL * 2 * SLOTS * UPC / 2; i += THREADS) reinterpret_cast<uint32_t*>(h_buf)[i] = 0u;
  if (tid == 0) {
    mbar_init(&h_bar[0], 1);
    mbar_init(&h_bar[1], 1);
    mbar_fence_init();
  }

  // This lane's accumulator columns of n-tile j are batch slots 8 j + 2 tig and 8 j + 2 tig + 1.
  const int n0 = 2 * tig;
  bool v0[NT], v1[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    v0[j] = (8 * j + n0) < NB && (b0 + 8 * j + n0) < B;
    v1[j] = (8 * j + n0 + 1) < NB && (b0 + 8 * j + n0 + 1) < B;
  }
  const size_t gc0 = (size_t)dir * 4 * H + (size_t)gate0 * H + unit;
  const size_t gc1 = (size_t)dir * 4 * H + (size_t)gate1 * H + unit;
  float c_state[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) c_state[j][0] = c_state[j][1] = 0.f;
  // Input projections are prefetched PF steps ahead (scattered 4-byte loads from HBM stay off the critical path).
  constexpr int PF = NT == 1 ? 3 : 2;
  float gq[PF][NT][4];  // (gate0, n0), (gate0, n0+1), (gate1, n0), (gate1, n0+1) per n-tile
  auto load_g = [&](int st, float (&dst)[NT][4]) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      dst[j][0] = dst[j][1] = dst[j][2] = dst[j][3] = 0.f;
      if (st < F) {
        const int tq = dir ? F - 1 - st : st;
        if (v0[j]) { const float* g = G + ((size_t)(b0 + 8 * j + n0) * F + tq) * ldg; dst[j][0] = g[gc0]; dst[j][2] = g[gc1]; }
        if (v1[j]) { const float* g = G + ((size_t)(b0 + 8 * j + n0 + 1) * F + tq) * ldg; dst[j][1] = g[gc0]; dst[j][3] = g[gc1]; }
      }
    }
  };
#pragma unroll
  for (int j = 0; j < PF; ++j) load_g(j, gq[j]);
  const uint32_t dst_h = mapa_u32(smem_u32(h_buf) + rank * BLK_BYTES, lane & 7);
  const uint32_t dst_bar = mapa_u32(smem_u32(&h_bar[0]), lane & 7);
  // B-fragment byte offsets inside one h buffer: true k0 = 32 P + 8 tig lives in source CTA k0 / UPC at unit k0 % UPC
  uint32_t boffs[NP];
#pragma unroll
  for (int P = 0; P < NP; ++P) {
    const int k0 = 32 * P + 8 * tig;
    boffs[P] = (uint32_t)((k0 / UPC) * BLK_BYTES + gid * (UPC * 2) + (k0 % UPC) * 2);
  }

  __syncthreads();
  cluster_arrive();
  cluster_wait();

  for (int step = 0; step < F; ++step) {
    const int cur = step & 1;
    const int tt = dir ? F - 1 - step : step;
    if (tid == 0) mbar_arrive_expect_tx(&h_bar[cur ^ 1], TX);
    float gin[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) gin[j][q] = gq[0][j][q];
#pragma unroll
    for (int i = 0; i + 1 < PF; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) gq[i][j][q] = gq[i + 1][j][q];
    load_g(step + PF, gq[PF - 1]);
    if (step > 0) mbar_wait(&h_bar[cur], ((step - 1) >> 1) & 1);

    // ---- mat-vec on the tensor cores: three independent accumulation chains per n-tile ----
    float d0[NT][4], d1[NT][4], d2[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) d0[j][q] = d1[j][q] = d2[j][q] = 0.f;
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(h_buf) + cur * (LSTM_CL * BLK_BYTES);
#pragma unroll
    for (int P = 0; P < NP; ++P) {
      uint32_t al0[4], al1[4];
      if (LO_SMEM) {
        const uint4 l0 = alo_smem[(2 * P) * THREADS + tid], l1 = alo_smem[(2 * P + 1) * THREADS + tid];
        al0[0] = l0.x; al0[1] = l0.y; al0[2] = l0.z; al0[3] = l0.w;
        al1[0] = l1.x; al1[1] = l1.y; al1[2] = l1.z; al1[3] = l1.w;
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { al0[q] = a_lo[LO_SMEM ? 0 : 2 * P][q]; al1[q] = a_lo[LO_SMEM ? 0 : 2 * P + 1][q]; }
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const uint4 bh = *reinterpret_cast<const uint4*>(hb + boffs[P] + j * (LSTM_SLOTS * UPC * 2));
        const uint4 bl = *reinterpret_cast<const uint4*>(hb + boffs[P] + j * (LSTM_SLOTS * UPC * 2) + SLOTS * UPC * 2);
        hmma16816(d0[j], al0, bh.x, bh.y);
        hmma16816(d1[j], a_hi[2 * P], bl.x, bl.y);
        hmma16816(d2[j], a_hi[2 * P], bh.x, bh.y);
        hmma16816(d0[j], al1, bh.z, bh.w);
        hmma16816(d1[j], a_hi[2 * P + 1], bl.z, bl.w);
        hmma16816(d2[j], a_hi[2 * P + 1], bh.z, bh.w);
      }
    }
    __nv_bfloat16* stg = stage + (cur ^ 1) * (2 * SLOTS * UPC);
    float h0[NT], h1[NT];
    uint32_t hh[NT], hl[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      // d[0], d[1]: row gate0, slots n0, n0+1 ; d[2], d[3]: row gate1
      float pre[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pre[i] = (d0[j][i] + d1[j][i]) + d2[j][i] + gin[j][i];
      // ---- gate math.  pp = 0: (i, g) -> i * tanh(g);  pp = 1: (f, o).  tanh(x) = 2 sigmoid(2x) - 1 keeps it branch-free ----
      const float sa0 = fast_sigmoid(pre[0]), sa1 = fast_sigmoid(pre[1]);  // sigmoid(i) | sigmoid(f)
      const float sc = pp ? 1.0f : 2.0f;
      float sb0 = fast_sigmoid(sc * pre[2]), sb1 = fast_sigmoid(sc * pre[3]);  // sigmoid(o) | sigmoid(2g)
      if (!pp) { sb0 = 2.0f * sb0 - 1.0f; sb1 = 2.0f * sb1 - 1.0f; }          // tanh(g)
      const float ig0 = __shfl_xor_sync(0xffffffffu, sa0 * sb0, 4);           // partner lane (gid ^ 1): i * tanh(g)
      const float ig1 = __shfl_xor_sync(0xffffffffu, sa1 * sb1, 4);
      h0[j] = h1[j] = 0.f;
      hh[j] = hl[j] = 0u;
      if (pp) {
        c_state[j][0] = sa0 * c_state[j][0] + ig0;
        c_state[j][1] = sa1 * c_state[j][1] + ig1;
        h0[j] = sb0 * (2.0f * fast_sigmoid(2.0f * c_state[j][0]) - 1.0f);
        h1[j] = sb1 * (2.0f * fast_sigmoid(2.0f * c_state[j][1]) - 1.0f);
        float r0, r1;
        hh[j] = pack_hi2(h0[j], h1[j], r0, r1);
        hl[j] = pack2(r0, r1);
        __nv_bfloat16* st = stg + (8 * j + n0) * UPC + warp * 4 + u;
        st[0] = reinterpret_cast<const __nv_bfloat16*>(&hh[j])[0];
        st[UPC] = reinterpret_cast<const __nv_bfloat16*>(&hh[j])[1];
        st[SLOTS * UPC] = reinterpret_cast<const __nv_bfloat16*>(&hl[j])[0];
        st[SLOTS * UPC + UPC] = reinterpret_cast<const __nv_bfloat16*>(&hl[j])[1];
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0 && lane < LSTM_CL) {
      const uint32_t boff = (uint32_t)((cur ^ 1) * LSTM_CL * BLK_BYTES);
      bulk_s2cluster(dst_h + boff, smem_u32(stg), BLK_BYTES, dst_bar + (uint32_t)((cur ^ 1) * sizeof(uint64_t)));
    }
    // layer output to HBM: off the critical path (overlaps the DSMEM exchange)
    if (pp) {
      const int col = dir * H + unit;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if (v0[j]) {
          const size_t row = (size_t)(b0 + 8 * j + n0) * F + tt;
          if (Hout) Hout[row * ldh + col] = h0[j];
          if (Hhi) { Hhi[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hh[j])[0]; Hlo[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hl[j])[0]; }
        }
        if (v1[j]) {
          const size_t row = (size_t)(b0 + 8 * j + n0 + 1) * F + tt;
          if (Hout) Hout[row * ldh + col] = h1[j];
          if (Hhi) { Hhi[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hh[j])[1]; Hlo[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hl[j])[1]; }
        }
      }
    }
  }
  mbar_wait(&h_bar[F & 1], ((F - 1) >> 1) & 1);
  cluster_arrive();
  cluster_wait();
}

template <int H, bool LO_SMEM, int NT>
static size_t lstm_mma_smem() {
  constexpr int UPC = H / LSTM_CL;
  constexpr int SLOTS = LSTM_SLOTS * NT;
  size_t n = (size_t)(2 * LSTM_CL * 2 * SLOTS * UPC + 2 * 2 * SLOTS * UPC) * 2 + 128;
  if (LO_SMEM) n += (size_t)(H / 16) * (UPC / 4 * 32) * 16;
  return n;
}

template <int H, bool LO_SMEM, int NT>
static int launch_mma(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs, int B,
                      int F, int slots, cudaStream_t stream) {
  constexpr int SLOTS = LSTM_SLOTS * NT;
  const size_t smem = lstm_mma_smem<H, LO_SMEM, NT>();
  RFX_CHECK_CUDA(cudaFuncSetAttribute(lstm_rec_mma_kernel<H, LO_SMEM, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // batch slots per cluster: as few as possible while all clusters stay co-resident (the MMA cost does not depend on it)
  static int maxc = -1;
  if (maxc < 0) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(LSTM_CL, 64, 2);
    cfg.blockDim = dim3(H / 8 / 4 * 32);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = LSTM_CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_rec_mma_kernel<H, LO_SMEM, NT>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    maxc = n > 0 ? n : 14;
  }
  int nb = SLOTS;
  if (slots > 0) {
    nb = slots < SLOTS ? slots : SLOTS;
  } else {
    for (int cand = 1; cand <= SLOTS; ++cand)
      if (2 * ceil_div(B, cand) <= maxc) { nb = cand; break; }
  }
  dim3 grid(LSTM_CL, ceil_div(B, nb), 2);
  lstm_rec_mma_kernel<H, LO_SMEM, NT><<<grid, H / 8 / 4 * 32, smem, stream>>>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, nb);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int NB>
static int max_clusters() {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(LSTM_CL, 64, 2);
  cfg.blockDim = dim3(LSTM_WARPS * 32);
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = LSTM_CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_rec_kernel<NB>, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

template <int NB>
static int launch_nb(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs, int B,
                     int F, cudaStream_t stream) {
  dim3 grid(LSTM_CL, ceil_div(B, NB), 2);
  lstm_rec_kernel<NB><<<grid, LSTM_WARPS * 32, 0, stream>>>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int lstm_max_active_clusters() {
  static int cached = -1;
  if (cached < 0) cached = max_clusters<4>();
  return cached;
}

int lstm_choose_nb(int B) {
  const int maxc = lstm_max_active_clusters();
  for (int nb = 4; nb <= 8; ++nb)
    if (2 * ceil_div(B, nb) <= maxc) return nb;
  return 8;  // more clusters than fit: several waves of the widest variant
}

static int g_lstm_impl = 0;  // 0 = tensor-core (mma.sync bf16x3), 1 = fp32 FFMA
void lstm_set_impl(int impl) { g_lstm_impl = impl; }
int lstm_get_impl() { return g_lstm_impl; }

int lstm_clusters_for(int B, int slots) {
  const int s = slots <= 0 ? LSTM_SLOTS : (slots < 2 * LSTM_SLOTS ? slots : 2 * LSTM_SLOTS);
  return 2 * ceil_div(B, s);
}

int launch_lstm_layer(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs,
                      int B, int F, int H, cudaStream_t stream) {
  return launch_lstm_layer_slots(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, H, 0, stream);
}

int launch_lstm_layer_slots(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo,
                            int ldhs, int B, int F, int H, int slots, cudaStream_t stream) {
  RFX_REQUIRE(H == 192 || H == 256 || H == 384, "lstm: hidden size per direction must be 192, 256 or 384");
  RFX_REQUIRE(B > 0 && F > 0, "lstm: positive sizes");
  RFX_REQUIRE(((uintptr_t)Whh & 15) == 0, "lstm: W_hh must be 16-byte aligned");
  RFX_REQUIRE(Hout || (Hhi && Hlo), "lstm: no output given");
  if (g_lstm_impl == 0 || H != LSTM_H) {
    // more than 8 slots per cluster asked for: the two-n-tile variant (H = 256 only)
    // (measured, B = 32: 1.01 ms per launch against 0.59 ms with one n-tile -- 1.7x the time for 2x the slots per SM)
    if (H == 256 && slots > LSTM_SLOTS) return launch_mma<256, false, 2>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
    if (H == 256) return launch_mma<256, false, 1>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
    if (H == 192) return launch_mma<192, false, 1>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
    return launch_mma<384, true, 1>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
  }
  switch (lstm_choose_nb(B)) {
    case 4: return launch_nb<4>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, stream);
    case 5: return launch_nb<5>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, stream);
    case 6: return launch_nb<6>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, stream);
    case 7: return launch_nb<7>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, stream);
    default: return launch_nb<8>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, stream);
  }
}

}  // namespace rfx
