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

#include <algorithm>
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1 / (1 + 2^(-x log2 e)): two MUFU ops, no range fix-up (ex2 saturates to 0 / inf and rcp(inf) = 0, which is what the
// sigmoid needs there)
__device__ __forceinline__ float lean_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ float fast_sigmoid(float x) { return lean_sigmoid(x); }

constexpr int LSTM_SLOTS = 8;  // batch slots per cluster (MMA N)

// H = hidden units per direction (multiple of 32 with H/8 a multiple of 8: 192, 256, 384).  Each of the 8 CTAs owns
// UPC = H/8 units (UPC/4 warps).  The W_hh hi fragments always live in registers (H/16 k-steps x 4 regs); when they
// would not both fit (H = 384) the lo fragments are kept in shared memory instead (LO_SMEM) and re-read every step.
template <int H, bool LO_SMEM>
__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(H / 8 / 4 * 32, 1)
    lstm_rec_mma_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Whh, float* __restrict__ Hout, int ldh,
                        __nv_bfloat16* __restrict__ Hhi, __nv_bfloat16* __restrict__ Hlo, int ldhs, int B, int F, int NB) {
  constexpr int UPC = H / LSTM_CL;          // units per CTA
  constexpr int WARPS = UPC / 4;
  constexpr int THREADS = WARPS * 32;
  constexpr int KS = H / 16;                // MMA k-steps
  constexpr int NP = H / 32;                // k-step pairs (one LDS.128 of h per plane each)
  constexpr int BLK_BYTES = 2 * LSTM_SLOTS * UPC * 2;  // one CTA's h block: 2 planes x 8 slots x UPC units bf16
  constexpr int TX = LSTM_CL * BLK_BYTES;
  static_assert(UPC % 8 == 0 && H % 32 == 0, "unsupported hidden size");
  extern __shared__ __align__(128) uint8_t lstm_smem[];
  __nv_bfloat16* h_buf = reinterpret_cast<__nv_bfloat16*>(lstm_smem);                         // [2][CL][2][SLOTS][UPC]
  __nv_bfloat16* stage = h_buf + 2 * LSTM_CL * 2 * LSTM_SLOTS * UPC;                           // [2][2][SLOTS][UPC]
  uint64_t* h_bar = reinterpret_cast<uint64_t*>(stage + 2 * 2 * LSTM_SLOTS * UPC);             // [2]
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
  for (int i = tid; i < 2 * LSTM_CL * 2 * LSTM_SLOTS * UPC / 2; i += THREADS) reinterpret_cast<uint32_t*>(h_buf)[i] = 0u;
  if (tid == 0) {
    mbar_init(&h_bar[0], 1);
    mbar_init(&h_bar[1], 1);
    mbar_fence_init();
  }

  // This lane's accumulator columns are batch slots n0 = 2 tig and n0 + 1.
  const int n0 = 2 * tig;
  const bool v0 = n0 < NB && (b0 + n0) < B, v1 = (n0 + 1) < NB && (b0 + n0 + 1) < B;
  const size_t gc0 = (size_t)dir * 4 * H + (size_t)gate0 * H + unit;
  const size_t gc1 = (size_t)dir * 4 * H + (size_t)gate1 * H + unit;
  float c_state[2] = {0.f, 0.f};
  // Input projections are prefetched PF steps ahead (scattered 4-byte loads from HBM stay off the critical path).
  constexpr int PF = 3;
  float gq[PF][4];  // (gate0, n0), (gate0, n0+1), (gate1, n0), (gate1, n0+1)
  auto load_g = [&](int st, float (&dst)[4]) {
    dst[0] = dst[1] = dst[2] = dst[3] = 0.f;
    if (st < F) {
      const int tq = dir ? F - 1 - st : st;
      if (v0) { const float* g = G + ((size_t)(b0 + n0) * F + tq) * ldg; dst[0] = g[gc0]; dst[2] = g[gc1]; }
      if (v1) { const float* g = G + ((size_t)(b0 + n0 + 1) * F + tq) * ldg; dst[1] = g[gc0]; dst[3] = g[gc1]; }
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
    float gin[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) gin[q] = gq[0][q];
#pragma unroll
    for (int j = 0; j + 1 < PF; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) gq[j][q] = gq[j + 1][q];
    load_g(step + PF, gq[PF - 1]);
    if (step > 0) mbar_wait(&h_bar[cur], ((step - 1) >> 1) & 1);

    // ---- mat-vec on the tensor cores: three independent accumulation chains ----
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(h_buf) + cur * (LSTM_CL * BLK_BYTES);
#pragma unroll
    for (int P = 0; P < NP; ++P) {
      const uint4 bh = *reinterpret_cast<const uint4*>(hb + boffs[P]);
      const uint4 bl = *reinterpret_cast<const uint4*>(hb + boffs[P] + LSTM_SLOTS * UPC * 2);
      if (LO_SMEM) {
        const uint4 l0 = alo_smem[(2 * P) * THREADS + tid], l1 = alo_smem[(2 * P + 1) * THREADS + tid];
        const uint32_t al0[4] = {l0.x, l0.y, l0.z, l0.w}, al1[4] = {l1.x, l1.y, l1.z, l1.w};
        hmma16816(d0, al0, bh.x, bh.y);
        hmma16816(d1, a_hi[2 * P], bl.x, bl.y);
        hmma16816(d2, a_hi[2 * P], bh.x, bh.y);
        hmma16816(d0, al1, bh.z, bh.w);
        hmma16816(d1, a_hi[2 * P + 1], bl.z, bl.w);
        hmma16816(d2, a_hi[2 * P + 1], bh.z, bh.w);
      } else {
        hmma16816(d0, a_lo[LO_SMEM ? 0 : 2 * P], bh.x, bh.y);
        hmma16816(d1, a_hi[2 * P], bl.x, bl.y);
        hmma16816(d2, a_hi[2 * P], bh.x, bh.y);
        hmma16816(d0, a_lo[LO_SMEM ? 0 : 2 * P + 1], bh.z, bh.w);
        hmma16816(d1, a_hi[2 * P + 1], bl.z, bl.w);
        hmma16816(d2, a_hi[2 * P + 1], bh.z, bh.w);
      }
    }
    // d[0], d[1]: row gate0, slots n0, n0+1 ; d[2], d[3]: row gate1
    float pre[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pre[i] = (d0[i] + d1[i]) + d2[i] + gin[i];
    // ---- gate math.  pp = 0: (i, g) -> i * tanh(g);  pp = 1: (f, o).  tanh(x) = 2 sigmoid(2x) - 1 keeps it branch-free ----
    const float sa0 = fast_sigmoid(pre[0]), sa1 = fast_sigmoid(pre[1]);  // sigmoid(i) | sigmoid(f)
    const float sc = pp ? 1.0f : 2.0f;
    float sb0 = fast_sigmoid(sc * pre[2]), sb1 = fast_sigmoid(sc * pre[3]);  // sigmoid(o) | sigmoid(2g)
    if (!pp) { sb0 = 2.0f * sb0 - 1.0f; sb1 = 2.0f * sb1 - 1.0f; }          // tanh(g)
    const float ig0 = __shfl_xor_sync(0xffffffffu, sa0 * sb0, 4);           // partner lane (gid ^ 1): i * tanh(g)
    const float ig1 = __shfl_xor_sync(0xffffffffu, sa1 * sb1, 4);
    float h0 = 0.f, h1 = 0.f;
    uint32_t hh = 0u, hl = 0u;
    __nv_bfloat16* stg = stage + (cur ^ 1) * (2 * LSTM_SLOTS * UPC);
    if (pp) {
      c_state[0] = sa0 * c_state[0] + ig0;
      c_state[1] = sa1 * c_state[1] + ig1;
      h0 = sb0 * (2.0f * fast_sigmoid(2.0f * c_state[0]) - 1.0f);
      h1 = sb1 * (2.0f * fast_sigmoid(2.0f * c_state[1]) - 1.0f);
      float r0, r1;
      hh = pack_hi2(h0, h1, r0, r1);
      hl = pack2(r0, r1);
      __nv_bfloat16* st = stg + n0 * UPC + warp * 4 + u;
      st[0] = reinterpret_cast<const __nv_bfloat16*>(&hh)[0];
      st[UPC] = reinterpret_cast<const __nv_bfloat16*>(&hh)[1];
      st[LSTM_SLOTS * UPC] = reinterpret_cast<const __nv_bfloat16*>(&hl)[0];
      st[LSTM_SLOTS * UPC + UPC] = reinterpret_cast<const __nv_bfloat16*>(&hl)[1];
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
      if (v0) {
        const size_t row = (size_t)(b0 + n0) * F + tt;
        if (Hout) Hout[row * ldh + col] = h0;
        if (Hhi) { Hhi[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hh)[0]; Hlo[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hl)[0]; }
      }
      if (v1) {
        const size_t row = (size_t)(b0 + n0 + 1) * F + tt;
        if (Hout) Hout[row * ldh + col] = h1;
        if (Hhi) { Hhi[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hh)[1]; Hlo[row * ldhs + col] = reinterpret_cast<const __nv_bfloat16*>(&hl)[1]; }
      }
    }
  }
  mbar_wait(&h_bar[F & 1], ((F - 1) >> 1) & 1);
  cluster_arrive();
  cluster_wait();
}

// ------------------------------------------------------------------------------------------------------
// Reverse-time chain of the LSTM backward (BPTT) on the same machinery, H = 256: one 8-CTA cluster per (direction, 8 sequences).
//   dh_t = dH_t + W_hh^T dG_{t'}   (t' = the step handled before),   gate derivatives -> dG_t,   dc carry.
// CTA r owns the hidden units [32 r, 32 r + 32) -- it produces dG_t for their 4 x 32 gate rows and needs dh_t for them -- and keeps
// W_hh^T restricted to ITS OWN gate rows as mma.sync A fragments in registers (M = all 256 units, K = 128 own gate rows: warp w
// holds units [32 w, 32 w + 32) = 2 m-tiles x 8 k-steps, hi and lo planes = 128 registers, the forward kernel's budget).  The B
// operand is the CTA's OWN dG_{t'} (split bf16 in shared memory, written by its own threads at the end of the previous step), so
// nothing has to be gathered: each warp's result is the partial sum over this CTA's gate rows of dh for the units of CTA w, sent
// there as one 1152-byte DSMEM bulk copy (a reduce-scatter: 8 KB out and in per CTA and step instead of the 32 KB an all-gather of
// dG would move), summed by the owner (thread = (unit, sequence)), followed by the gate-derivative math in fp32 with the
// accurate transcendentals of the other backward kernels.  Inputs / outputs are those of hd::lstm_bwd_persist_kernel (Gx, R =
// W_hh h_prev, cell states, dH -> dG), which it replaces for H = 256: 11 us -> ~1 us per time step.
// ------------------------------------------------------------------------------------------------------
constexpr int LBM_BLK = 8 * 36 * 4;  // one partial block: [8 sequences][32 units + 4 pad] fp32
constexpr int LBM_KP = 136;          // row stride (bf16) of the dG operand: [8 sequences][128 own gate rows + 8 pad]
// CL = CTAs per cluster = H / 32 (6: H = 192, 8: H = 256, 12: H = 384 -- a non-portable cluster size); one MMA warp per peer CTA,
// eight element warps (32 units x 8 sequences); LO_SMEM keeps the lo-plane A fragments in shared memory (H = 384: 384 threads
// cannot hold 128 fragment registers each).
template <int CL, bool LO_SMEM>
__global__ void __launch_bounds__(32 * (CL > 8 ? CL : 8), 1)
    lstm_bwd_mma_kernel(const float* __restrict__ Gx, const float* __restrict__ R, const float* __restrict__ cs, const float* __restrict__ dH,
                        const float* __restrict__ Whh /*[2][4H][H]*/, float* __restrict__ dG, int Bs, int T) {
  constexpr int H = 32 * CL, UPC = 32;
  constexpr int NW = CL > 8 ? CL : 8;  // warps per CTA
  constexpr int THREADS = 32 * NW;
  extern __shared__ __align__(128) uint8_t lbm_smem[];
  uint8_t* recv = lbm_smem;                                   // [2][CL sources][LBM_BLK]
  uint8_t* send = recv + 2 * CL * LBM_BLK;                    // [2][CL warps][LBM_BLK]
  __nv_bfloat16* dgs = reinterpret_cast<__nv_bfloat16*>(send + 2 * CL * LBM_BLK);  // [2 planes][8][LBM_KP]
  uint64_t* rbar = reinterpret_cast<uint64_t*>(dgs + 2 * 8 * LBM_KP);              // [2]
  uint4* alo_smem = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(rbar) + 64);  // [2 mt][8 ks][THREADS] (LO_SMEM only)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.y, b0 = blockIdx.z * 8;
  const bool mma_warp = warp < CL;

  // ---- A fragments: A[m = unit 32 w + 16 mt + row][k] = W_hh[gate row (k / 32) * H + 32 rank + k % 32][unit] ----
  uint32_t a_hi[2][8][4];
  uint32_t a_lo[LO_SMEM ? 1 : 2][LO_SMEM ? 1 : 8][4];
  if (mma_warp) {
    const float* W = Whh + (size_t)dir * 4 * H * H;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        uint32_t lo4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = gid + (q & 1) * 8;                 // a0/a2: gid, a1/a3: gid + 8
          const int k = 16 * ks + 2 * tig + (q >> 1) * 8;    // a0/a1: 2 tig, a2/a3: 2 tig + 8
          const int unit = 32 * warp + 16 * mt + row;
          const size_t g0 = (size_t)((k >> 5) * H + 32 * (int)rank + (k & 31));
          const float w0 = W[g0 * H + unit], w1 = W[(g0 + 1) * H + unit];  // k and k + 1 are in the same gate (k is even)
          float r0, r1;
          a_hi[mt][ks][q] = pack_hi2(w0, w1, r0, r1);
          lo4[q] = pack2(r0, r1);
        }
        if (LO_SMEM) {
          alo_smem[(mt * 8 + ks) * THREADS + tid] = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) a_lo[LO_SMEM ? 0 : mt][LO_SMEM ? 0 : ks][q] = lo4[q];
        }
      }
  }
  for (int i = tid; i < 2 * 8 * LBM_KP / 2; i += THREADS) reinterpret_cast<uint32_t*>(dgs)[i] = 0u;
  if (tid == 0) {
    mbar_init(&rbar[0], 1);
    mbar_init(&rbar[1], 1);
    mbar_fence_init();
  }
  // element role (warps 0-7): thread = (unit uu of this CTA, sequence n)
  const bool elem = warp < 8;
  const int uu = lane, n = warp;
  const int u = 32 * (int)rank + uu, b = b0 + n;
  const bool live = elem && b < Bs;
  const int G4 = 4 * H;
  // destination of this warp's partial block: CTA `warp`, slot [source = rank]
  const uint32_t dst_blk = mma_warp ? mapa_u32(smem_u32(recv) + rank * LBM_BLK, (uint32_t)warp) : 0u;
  const uint32_t dst_bar = mma_warp ? mapa_u32(smem_u32(&rbar[0]), (uint32_t)warp) : 0u;
  float dc = 0.0f;
  __syncthreads();
  cluster_arrive();
  cluster_wait();

  // operands of the step, fetched one step ahead
  float pi = 0.f, pf = 0.f, pg = 0.f, po = 0.f, cc = 0.f, cprev = 0.f, dh0 = 0.f;
  auto fetch = [&](int k) {
    pi = pf = pg = po = cc = cprev = dh0 = 0.f;
    if (live && k < T) {
      const int t = dir ? k : T - 1 - k;
      const int tp = dir ? t + 1 : t - 1;
      const size_t go = ((size_t)b * T + t) * 8 * H + (size_t)dir * G4 + u;
      const size_t ho = ((size_t)b * T + t) * 2 * H + dir * H + u;
      pi = Gx[go] + R[go]; pf = Gx[go + H] + R[go + H]; pg = Gx[go + 2 * H] + R[go + 2 * H]; po = Gx[go + 3 * H] + R[go + 3 * H];
      cc = cs[ho];
      cprev = (tp >= 0 && tp < T) ? cs[((size_t)b * T + tp) * 2 * H + dir * H + u] : 0.0f;
      dh0 = dH[ho];
    }
  };
  fetch(0);
  for (int k = 0; k < T; ++k) {
    const int buf = k & 1;
    const int t = dir ? k : T - 1 - k;
    float dh = dh0;
    const float xi = pi, xf = pf, xg = pg, xo = po, c = cc, cp = cprev;
    if (k > 0) {
      if (tid == 0) mbar_arrive_expect_tx(&rbar[buf], CL * LBM_BLK);
      if (mma_warp) {
        // ---- partial dh of the units of CTA `warp` from this CTA's gate rows: 2 m-tiles x 8 k-steps x bf16x3 ----
        float d[2][3][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int q = 0; q < 4; ++q) d[mt][ch][q] = 0.f;
        const uint8_t* bh = reinterpret_cast<const uint8_t*>(dgs) + gid * (LBM_KP * 2) + tig * 4;
        const uint8_t* bl = bh + 8 * LBM_KP * 2;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t h0 = *reinterpret_cast<const uint32_t*>(bh + ks * 32), h1 = *reinterpret_cast<const uint32_t*>(bh + ks * 32 + 16);
          const uint32_t l0 = *reinterpret_cast<const uint32_t*>(bl + ks * 32), l1 = *reinterpret_cast<const uint32_t*>(bl + ks * 32 + 16);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            if (LO_SMEM) {
              const uint4 lv = alo_smem[(mt * 8 + ks) * THREADS + tid];
              const uint32_t al[4] = {lv.x, lv.y, lv.z, lv.w};
              hmma16816(d[mt][0], al, h0, h1);
            } else {
              hmma16816(d[mt][0], a_lo[LO_SMEM ? 0 : mt][LO_SMEM ? 0 : ks], h0, h1);
            }
            hmma16816(d[mt][1], a_hi[mt][ks], l0, l1);
            hmma16816(d[mt][2], a_hi[mt][ks], h0, h1);
          }
        }
        float* blk = reinterpret_cast<float*>(send + (buf * CL + warp) * LBM_BLK);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float c0 = (d[mt][0][0] + d[mt][1][0]) + d[mt][2][0], c1 = (d[mt][0][1] + d[mt][1][1]) + d[mt][2][1];
          const float c2 = (d[mt][0][2] + d[mt][1][2]) + d[mt][2][2], c3 = (d[mt][0][3] + d[mt][1][3]) + d[mt][2][3];
          blk[(2 * tig) * 36 + 16 * mt + gid] = c0;
          blk[(2 * tig + 1) * 36 + 16 * mt + gid] = c1;
          blk[(2 * tig) * 36 + 16 * mt + gid + 8] = c2;
          blk[(2 * tig + 1) * 36 + 16 * mt + gid + 8] = c3;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0)
          bulk_s2cluster(dst_blk + (uint32_t)(buf * CL * LBM_BLK), smem_u32(blk), LBM_BLK, dst_bar + (uint32_t)(buf * sizeof(uint64_t)));
      }
    }
    fetch(k + 1);     // next step's operands: in flight during the exchange
    __syncthreads();  // every warp is done reading dG_{t'} (the element phase below overwrites it)
    if (elem) {
      if (k > 0) {
        mbar_wait(&rbar[buf], ((k - 1) >> 1) & 1);
        const float* rb = reinterpret_cast<const float*>(recv + buf * CL * LBM_BLK) + n * 36 + uu;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int src = 0; src < CL; src += 2) {
          s0 += rb[src * (LBM_BLK / 4)];
          s1 += rb[(src + 1) * (LBM_BLK / 4)];
        }
        dh += s0 + s1;
      }
      // ---- gate derivatives (tools/hd_bwd_emul.py: lstm_dir_bwd), as in hd::lstm_bwd_persist_kernel ----
      float g4[4] = {0.f, 0.f, 0.f, 0.f};
      if (live) {
        const float gi = sigmoidf_acc(xi), gf = sigmoidf_acc(xf), gg = tanhf(xg), go_ = sigmoidf_acc(xo);
        const float tc = tanhf(c);
        dc += dh * go_ * (1.0f - tc * tc);
        g4[0] = dc * gg * gi * (1.0f - gi);
        g4[1] = dc * cp * gf * (1.0f - gf);
        g4[2] = dc * gi * (1.0f - gg * gg);
        g4[3] = dh * tc * go_ * (1.0f - go_);
        dc *= gf;
        const size_t go = ((size_t)b * T + t) * 8 * H + (size_t)dir * G4 + u;
#pragma unroll
        for (int q = 0; q < 4; ++q) dG[go + (size_t)q * H] = g4[q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // this CTA's dG_t as the next step's B operand: [plane][sequence][k = gate * 32 + unit]
        __nv_bfloat16 hq, lq;
        split_bf16(g4[q], hq, lq);
        dgs[n * LBM_KP + q * UPC + uu] = hq;
        dgs[8 * LBM_KP + n * LBM_KP + q * UPC + uu] = lq;
      }
    }
    __syncthreads();
  }
  // nobody leaves while a peer may still be writing into its receive buffers (all sends of the last step are consumed above)
  cluster_arrive();
  cluster_wait();
}

template <int CL, bool LO_SMEM>
static int launch_lbm(const float* Gx, const float* R, const float* cs, const float* dH, const float* Whh, float* dG, int Bs, int T, cudaStream_t stream) {
  constexpr int NW = CL > 8 ? CL : 8;
  const size_t smem = (size_t)4 * CL * LBM_BLK + (size_t)2 * 8 * LBM_KP * 2 + 16 + 64 + (LO_SMEM ? (size_t)16 * 32 * NW * 16 : 0);
  auto kern = lstm_bwd_mma_kernel<CL, LO_SMEM>;
  RFX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CL > 8) RFX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL, 2, ceil_div(Bs, 8));
  cfg.blockDim = dim3(32 * NW);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  RFX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, Gx, R, cs, dH, Whh, dG, Bs, T));
  return 0;
}

bool lstm_bwd_chain_mma_supported(int H) { return H == 192 || H == 256 || H == 384; }

int launch_lstm_bwd_chain_mma(const float* Gx, const float* R, const float* cs, const float* dH, const float* Whh, float* dG, int Bs, int T, int H,
                              cudaStream_t stream) {
  if (H == 256) return launch_lbm<8, false>(Gx, R, cs, dH, Whh, dG, Bs, T, stream);
  if (H == 192) return launch_lbm<6, false>(Gx, R, cs, dH, Whh, dG, Bs, T, stream);
  if (H == 384) return launch_lbm<12, true>(Gx, R, cs, dH, Whh, dG, Bs, T, stream);
  set_error("lstm_bwd_mma_kernel: H must be 192, 256 or 384");
  return 2;
}

template <int H, bool LO_SMEM>
static size_t lstm_mma_smem() {
  constexpr int UPC = H / LSTM_CL;
  size_t n = (size_t)(2 * LSTM_CL * 2 * LSTM_SLOTS * UPC + 2 * 2 * LSTM_SLOTS * UPC) * 2 + 128;
  if (LO_SMEM) n += (size_t)(H / 16) * (UPC / 4 * 32) * 16;
  return n;
}

template <int H, bool LO_SMEM>
static int launch_mma(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs, int B,
                      int F, int slots, cudaStream_t stream) {
  const size_t smem = lstm_mma_smem<H, LO_SMEM>();
  RFX_CHECK_CUDA(cudaFuncSetAttribute(lstm_rec_mma_kernel<H, LO_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
    if (cudaOccupancyMaxActiveClusters(&n, lstm_rec_mma_kernel<H, LO_SMEM>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    maxc = n > 0 ? n : 14;
  }
  int nb = LSTM_SLOTS;
  if (slots > 0) {
    nb = slots < LSTM_SLOTS ? slots : LSTM_SLOTS;
  } else {
    for (int cand = 1; cand <= LSTM_SLOTS; ++cand)
      if (2 * ceil_div(B, cand) <= maxc) { nb = cand; break; }
  }
  dim3 grid(LSTM_CL, ceil_div(B, nb), 2);
  lstm_rec_mma_kernel<H, LO_SMEM><<<grid, H / 8 / 4 * 32, smem, stream>>>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, nb);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------------
// tcgen05 recurrence (H = 256): the per-step product W_hh h_{t-1} on the 5th-generation tensor cores.
//
//   D[128 gate rows x 16 slots] (TMEM, fp32)  =  A[128 x 256] (TMEM)  x  B[256 x 16] (shared memory),  bf16x3
//
// * A = this CTA's 128 rows of W_hh (gate-major: row = gate * 32 + unit), split into bf16 hi / lo planes and written ONCE
//   into tensor memory with tcgen05.st (lane = row, two K-elements per 32-bit column: 128 columns per plane).  A from
//   TMEM (the ".ts" operand form) matters here: with N = 16 an MMA from shared memory would be bound by re-reading the
//   4 KB A tile every instruction; from TMEM the instruction paces at N/2 = 8 cycles.
// * B = h_{t-1} of the cluster's 16 batch slots as split-bf16, K-major, NO swizzle: 8 x 8 core matrices (128 B each) laid
//   out [source CTA][plane][slot group][unit group], so every CTA's contribution is ONE contiguous 2 KB block (one DSMEM
//   bulk copy per destination, as before) and an MMA k-step (16 units) is two K-adjacent core matrices
//   (descriptor: LBO = 128 B between K-adjacent core matrices, SBO = 512 B between 8-slot groups).
// * One thread issues the 48 MMAs of a step (lo*hi, hi*lo, hi*hi over 16 k-steps) and a tcgen05.commit; 8 epilogue warps
//   read D with tcgen05.ld (thread = gate row, 8 slots each), add the input projections, apply the gate non-linearity,
//   transpose through shared memory so that one thread owns (unit pair, slot), update c, emit h (fp32 / split planes to
//   HBM, split planes to the staging block) and the block is pushed to the 8 CTAs.
// A cluster serves 16 slots, so B = 32 needs 4 clusters = 32 SMs per launch.
// ------------------------------------------------------------------------------------------------------
constexpr int TC_N = 16;             // batch slots per cluster = MMA N (default); a 32-slot instantiation serves B = 32 per direction
constexpr int TC_N_MAX = 32;         //   with ONE cluster (8 SMs): slower per launch, half the SMs again
constexpr int TC_ISSUERS = 4;        // MMA-issue warps: each issues the k-steps of a quarter of K into its own accumulator
constexpr int TC_THREADS = 256 + 32 * TC_ISSUERS;  // 8 epilogue warps + the issue warps

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, the shared-memory descriptor given as two 32-bit halves (the low half is the only part that changes between the
// MMAs of a step, by a compile-time constant, which keeps the single issuing thread's instruction count down)
__device__ __forceinline__ void umma_ts_f16_split(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// K-major operand without swizzle: 8 x 16-byte core matrices; lbo = bytes between K-adjacent core matrices, sbo = bytes
// between 8-row groups (cute::UMMA::SmemDescriptor, version 1, layout type 0)
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// GROUPS = 2: TWO independent N-slot machines share a CTA (threads [0, TC_THREADS) and [TC_THREADS, 2 TC_THREADS)): each has its own
// h buffers, barriers, accumulators and named barriers, only W_hh in tensor memory is shared.  Nothing synchronises the two after the
// set-up, so the warp schedulers overlap one group's exchange / hand-off latencies with the other group's MMAs and gate math -- the
// "two slot groups out of phase" of VERDICT r1 item 6 without a software pipeline.  A cluster then serves 2 N slots.
template <int N, int GROUPS = 1>  // batch slots per group = MMA N (16 or 32)
__global__ void __cluster_dims__(LSTM_CL, 1, 1) __launch_bounds__(TC_THREADS * GROUPS, 1)
    lstm_rec_tc_kernel(const float* __restrict__ G, int ldg, const float* __restrict__ Whh, float* __restrict__ Hout, int ldh,
                       __nv_bfloat16* __restrict__ Hhi, __nv_bfloat16* __restrict__ Hlo, int ldhs, int B, int F, int NB, int fast) {
  constexpr int H = 256;
  const int dbg = fast >> 8;  // timing experiments only (RFX_LSTM_TC_DEBUG, wrong results on purpose): 1 = exchange the hi plane only, 4 = no MMAs,
                              // 8 = no G loads, 16 = no phase lock between the two groups
  fast &= 1;
  constexpr int UPC = H / LSTM_CL;  // 32 units per CTA -> 128 gate rows = MMA M
  constexpr int TC_BLKP = N * 64;   // bytes of one plane of one CTA's h block: N slots x 32 units bf16
  constexpr int TC_BLK = 2 * TC_BLKP;  // hi + lo
  constexpr int HS = N / 2;         // slots per epilogue thread in the activation phase
  constexpr int SPT = N / 16;       // slots per epilogue thread in the cell phase
  extern __shared__ __align__(128) uint8_t lstm_smem[];
  constexpr int GROUP_SMEM = 2 * LSTM_CL * TC_BLK + 2 * TC_BLK + 4 * N * UPC * 4 + 256;  // per-group shared-memory region (multiple of 128)
  const int grp = GROUPS > 1 ? (int)threadIdx.x / TC_THREADS : 0;
  uint8_t* gbase = lstm_smem + (size_t)grp * GROUP_SMEM;
  uint8_t* h_buf = gbase;                                                    // [2][CL][TC_BLK]
  uint8_t* stage = h_buf + 2 * LSTM_CL * TC_BLK;                             // [2][TC_BLK]
  float* act = reinterpret_cast<float*>(stage + 2 * TC_BLK);                 // [4 gates][N slots][32 units]
  uint64_t* h_bar = reinterpret_cast<uint64_t*>(act + 4 * N * UPC);          // [2][CL]: one per (buffer, source CTA)
  uint64_t* mma_bar = h_bar + 2 * LSTM_CL;                                   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lstm_smem + 2 * LSTM_CL * TC_BLK + 2 * TC_BLK + 4 * N * UPC * 4 + (2 * LSTM_CL + 1) * 8);  // group 0's

  const int tid = GROUPS > 1 ? (int)threadIdx.x % TC_THREADS : (int)threadIdx.x;   // thread index inside the group
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int dir = blockIdx.z;
  const int b0 = blockIdx.y * NB + grp * N;           // first batch item of this group
  const int NBg = GROUPS > 1 ? max(0, min(N, NB - grp * N)) : NB;   // slots of this group that hold items
  const int bar1 = 1 + 2 * grp, bar2 = 2 + 2 * grp;   // named barriers of this group's epilogue warps

  for (int i = tid; i < 2 * LSTM_CL * TC_BLK / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(h_buf)[i] = 0u;
  if (tid == 0) {
    for (int i = 0; i < 2 * LSTM_CL; ++i) mbar_init(&h_bar[i], 1);
    mbar_init(mma_bar, TC_ISSUERS);
    mbar_fence_init();
  }
  if (warp == 8 && grp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // the zeroed h buffer is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d = tmem_base + 256 + (uint32_t)(grp * TC_ISSUERS * N);  // columns [0,128) A hi, [128,256) A lo, then 4 N accumulator columns per group

  // ---- W_hh -> tensor memory (once, shared by the groups).  Warps 0-3 of group 0: thread = row (gate = warp, unit = lane). ----
  if (warp < 4 && grp == 0) {
    const float* wrow = Whh + ((size_t)dir * 4 * H + (size_t)warp * H + rank * UPC + lane) * H;
    const uint32_t t_row = tmem_base + ((uint32_t)(32 * warp) << 16);
    for (int c0 = 0; c0 < H / 2; c0 += 8) {  // 8 columns = 16 K-elements
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(wrow + 2 * c0 + 4 * q);
        float r0, r1, r2, r3;  // even K-element in the low half of the column (verified on the device against the oracle)
        hi[2 * q] = pack_hi2(v.x, v.y, r0, r1); lo[2 * q] = pack2(r0, r1);
        hi[2 * q + 1] = pack_hi2(v.z, v.w, r2, r3); lo[2 * q + 1] = pack2(r2, r3);
      }
      tmem_st8(t_row + c0, hi);
      tmem_st8(t_row + 128 + c0, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();  // barriers initialised and h zeroed everywhere before anyone sends
  cluster_wait();

  if (warp >= 8) {
    // =============================== MMA issue warps ===============================
    // Issuer w owns k-steps 4 w .. 4 w + 3 (units 64 w .. 64 w + 63) of all three passes = 12 MMAs into accumulator w: four
    // threads issue in parallel and the accumulation order inside every accumulator is fixed (deterministic results).
    // every operand of the MMAs is made provably warp-uniform (shuffles from lane 0) and the issuing lane is picked with
    // elect.sync, so the compiler feeds UTCHMMA from uniform registers directly instead of a per-MMA broadcast loop
    const int w = __shfl_sync(0xffffffffu, warp - 8, 0);
    const uint32_t tb_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t hb_u = __shfl_sync(0xffffffffu, smem_u32(h_buf), 0);
    const uint32_t mb_u = __shfl_sync(0xffffffffu, smem_u32(mma_bar), 0);
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    constexpr uint32_t lbo = 128u, sbo = 512u;
    constexpr uint32_t desc_hi = (sbo >> 4) | (1u << 14);  // SBO, descriptor version 1, no swizzle
    const uint32_t d_acc = tb_u + 256u + (uint32_t)(grp * TC_ISSUERS * N) + (uint32_t)(w * N);
    // The input projections G (134 MB per layer at B = 32: HBM-resident) are needed by the epilogue warps once per step, one 128-byte
    // line per (slot, gate).  Issued one step ahead their DRAM latency is not always hidden (single-group launch: 0.739 -> 0.658 ms
    // with this prefetch); the issue warps, idle most of the step, pull the lines of step + TC_PF into L2: issuer w = gate w, lane = slot.
    constexpr int TC_PF = 4;
    const bool pf_on = lane < N && lane < NBg && (b0 + lane) < B && !(dbg & 8);
    const float* pf_base = G + (size_t)(b0 + lane) * (size_t)F * (size_t)ldg + (size_t)(dir * 4 * H + w * H + (int)rank * UPC);
    auto prefetch_g = [&](int st) {
      if (pf_on && st < F) {
        const int tq = dir ? F - 1 - st : st;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf_base + (size_t)tq * (size_t)ldg));
      }
    };
    for (int st = 1; st < TC_PF; ++st) prefetch_g(st);
    // Phase lock of the two groups (GROUPS = 2).  Left alone the groups settle at an arbitrary relative phase: out of phase a
    // 32-slot launch takes ~0.75 ms, in phase (both on the tensor pipe, then both in the gate math, then both exchanging) 1.0+ ms,
    // and which one a launch got changed with unrelated code edits.  The lock: group 1 issues the MMAs of step s only when group
    // 0's MMAs of step s have completed, group 0 those of step s + 1 only when group 1's of step s have (each waits on the OTHER
    // group's mma_bar, which the epilogue warps of that group wait on anyway).  That keeps the groups at least one MMA phase apart
    // and at most one step minus an MMA phase, without adding a dependency longer than a step; measured 1.00 -> 0.76 ms.  Locks
    // on later events (gate math done: 0.91 ms; h pushed: 1.10 ms) serialise too much of the two steps.
    const bool lock = GROUPS > 1 && !(dbg & 16);
    uint64_t* mma_bar_other = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(mma_bar) + (GROUPS > 1 ? (grp ? -GROUP_SMEM : GROUP_SMEM) : 0));
    for (int step = 0; step < F; ++step) {
      const int cur = step & 1;
      prefetch_g(step + TC_PF);
      if (lock) {
        if (grp == 1) mbar_wait(mma_bar_other, step & 1);
        else if (step > 0) mbar_wait(mma_bar_other, (step - 1) & 1);
      }
      // units [16 kk, 16 kk + 16) live in source CTA kk / 2, unit groups 2 (kk & 1), + 1 of its block.  Every source block has
      // its own mbarrier and the senders rotate their destinations, so the blocks of a step land spread over the exchange:
      // this issuer starts on source 2 w as soon as THAT block is here, then source 2 w + 1.
      const uint32_t lo0 = (((hb_u + (uint32_t)cur * (LSTM_CL * TC_BLK) + (uint32_t)(2 * w) * TC_BLK) & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
      const uint32_t a0 = tb_u + 32u * w;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (step > 0) mbar_wait(&h_bar[cur * LSTM_CL + 2 * w + half], ((step - 1) >> 1) & 1);  // h_{t-1} of source CTA 2 w + half
        tc_fence_after();
        if (elect_one()) {
          if (dbg & 4) {
          } else if (fast) {  // bf16-fast (set_matmul_precision(1)): W_hi h_hi only
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
              const int k4 = 2 * half + k2;
              const uint32_t boff = (uint32_t)(half * TC_BLK + k2 * 256);
              umma_ts_f16_split(d_acc, a0 + 8u * k4, lo0 + (boff >> 4), desc_hi, idesc, (half | k2) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {  // W_lo h_hi, W_hi h_lo, W_hi h_hi (small terms first)
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int k4 = 2 * half + k2;
                const uint32_t boff = (uint32_t)(half * TC_BLK + k2 * 256 + (pass == 1 ? TC_BLKP : 0));
                umma_ts_f16_split(d_acc, a0 + (pass == 0 ? 128u : 0u) + 8u * k4, lo0 + (boff >> 4), desc_hi, idesc, (half | pass | k2) ? 1u : 0u);
              }
            }
          }
          if (half == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb_u) : "memory");
        }
        __syncwarp();
      }
    }
    // Nobody may exit while peers can still write into its shared memory: wait for the last h blocks to land.
    mbar_wait(&h_bar[(F & 1) * LSTM_CL + 2 * w], ((F - 1) >> 1) & 1);
    mbar_wait(&h_bar[(F & 1) * LSTM_CL + 2 * w + 1], ((F - 1) >> 1) & 1);
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3;    // TMEM lane quarter = gate
    const int ch = warp >> 2;  // column half: slots HS ch .. HS ch + HS - 1
    const uint32_t t_ld = tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)(HS * ch);
    const uint32_t gcol = (uint32_t)(dir * 4 * H + q * H + rank * UPC + lane);
    uint32_t vmask = 0;  // bit s: slot HS ch + s holds a real item
#pragma unroll
    for (int s2 = 0; s2 < HS; ++s2)
      if ((HS * ch + s2) < NBg && (b0 + HS * ch + s2) < B) vmask |= 1u << s2;
    const uint32_t row0 = (uint32_t)(b0 + HS * ch) * (uint32_t)F;  // G row of slot HS ch at t = 0
    float gq[HS];
    auto load_g = [&](int st) {
#pragma unroll
      for (int s2 = 0; s2 < HS; ++s2) {
        gq[s2] = 0.f;
        if (st < F && (vmask >> s2 & 1u) && !(dbg & 8)) {
          const uint32_t tq = (uint32_t)(dir ? F - 1 - st : st);
          gq[s2] = __ldg(G + (size_t)((row0 + (uint32_t)s2 * (uint32_t)F + tq) * (uint32_t)ldg + gcol));
        }
      }
    };
    load_g(0);
    const float sc = (q == 2) ? 2.0f : 1.0f;  // tanh(x) = 2 sigmoid(2x) - 1 for the cell gate
    // cell-phase role: one thread per (unit pair, slot cs + 16 jj)
    const int up = tid & 15, cs = tid >> 4;  // units 2 up, 2 up + 1; slots cs, cs + 16, ...
    const uint32_t ccol = (uint32_t)(dir * H + rank * UPC + 2 * up);
    float c0[SPT], c1[SPT];
#pragma unroll
    for (int jj = 0; jj < SPT; ++jj) c0[jj] = c1[jj] = 0.f;
    // sender lane i pushes this CTA's block to CTA (rank + i) mod 8: rotated, so every destination's eight blocks arrive in turn
    const uint32_t dst_cta = (rank + (uint32_t)lane) & 7u;
    const uint32_t dst_h = mapa_u32(smem_u32(h_buf) + rank * TC_BLK, dst_cta);
    const uint32_t dst_bar = mapa_u32(smem_u32(&h_bar[rank]), dst_cta);
    for (int step = 0; step < F; ++step) {
      const int cur = step & 1, nxt = cur ^ 1;
      const uint32_t tt = (uint32_t)(dir ? F - 1 - step : step);
      mbar_wait(mma_bar, step & 1);
      tc_fence_after();
      float acc[HS];
      {
        uint32_t v[TC_ISSUERS][HS];
#pragma unroll
        for (int a = 0; a < TC_ISSUERS; ++a) {
          if (HS == 8) tmem_ld8(t_ld + (uint32_t)(a * N), reinterpret_cast<uint32_t(&)[8]>(v[a]));
          else tmem_ld16(t_ld + (uint32_t)(a * N), reinterpret_cast<uint32_t(&)[16]>(v[a]));
        }
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int s2 = 0; s2 < HS; ++s2) {
          float t = __uint_as_float(v[0][s2]);
#pragma unroll
          for (int a = 1; a < TC_ISSUERS; ++a) t += __uint_as_float(v[a][s2]);
          acc[s2] = t;
        }
      }
      float* acol = act + (q * N + HS * ch) * UPC + lane;  // [gate][slot][unit]: conflict-free both ways
#pragma unroll
      for (int s2 = 0; s2 < HS; ++s2) {
        const float sg = lean_sigmoid(sc * (acc[s2] + gq[s2]));
        acol[s2 * UPC] = (q == 2) ? 2.0f * sg - 1.0f : sg;
      }
      named_bar_sync(bar1, 256);
      // ---- (unit pair, slot): c = f c + i g ; h = o tanh(c) ----
      uint8_t* stg = stage + nxt * TC_BLK;
      float h0[SPT], h1[SPT];
      uint32_t hh[SPT], hl[SPT];
#pragma unroll
      for (int jj = 0; jj < SPT; ++jj) {
        const int sl = cs + 16 * jj;
        const float* ap = act + sl * UPC + 2 * up;
        const float2 gi = *reinterpret_cast<const float2*>(ap);
        const float2 gf = *reinterpret_cast<const float2*>(ap + N * UPC);
        const float2 gg = *reinterpret_cast<const float2*>(ap + 2 * N * UPC);
        const float2 go = *reinterpret_cast<const float2*>(ap + 3 * N * UPC);
        c0[jj] = gf.x * c0[jj] + gi.x * gg.x;
        c1[jj] = gf.y * c1[jj] + gi.y * gg.y;
        h0[jj] = go.x * (2.0f * lean_sigmoid(2.0f * c0[jj]) - 1.0f);
        h1[jj] = go.y * (2.0f * lean_sigmoid(2.0f * c1[jj]) - 1.0f);
        float r0, r1;
        hh[jj] = pack_hi2(h0[jj], h1[jj], r0, r1);
        hl[jj] = pack2(r0, r1);
        const uint32_t st_off = (uint32_t)(((sl >> 3) * 4 + (up >> 2)) * 128 + (sl & 7) * 16 + (up & 3) * 4);  // core-matrix layout
        *reinterpret_cast<uint32_t*>(stg + st_off) = hh[jj];
        *reinterpret_cast<uint32_t*>(stg + TC_BLKP + st_off) = hl[jj];
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy (async proxy) reads
      named_bar_sync(bar2, 256);
      if (warp == 0 && lane < LSTM_CL) {
        const uint32_t xbytes = (dbg & 1) ? TC_BLKP : TC_BLK;
        mbar_arrive_expect_tx(&h_bar[nxt * LSTM_CL + lane], xbytes);  // lane s arms the local barrier of source CTA s
        bulk_s2cluster(dst_h + (uint32_t)(nxt * LSTM_CL * TC_BLK), smem_u32(stg), xbytes, dst_bar + (uint32_t)(nxt * LSTM_CL * sizeof(uint64_t)));
      }
      load_g(step + 1);  // next step's input projections: in flight during the exchange, never in front of the proxy fence
#pragma unroll
      for (int jj = 0; jj < SPT; ++jj) {  // layer output to HBM: off the critical path
        const int sl = cs + 16 * jj;
        if (sl < NBg && (b0 + sl) < B) {
          const uint32_t row = (uint32_t)(b0 + sl) * (uint32_t)F + tt;
          if (Hout) *reinterpret_cast<float2*>(Hout + (size_t)(row * (uint32_t)ldh + ccol)) = make_float2(h0[jj], h1[jj]);
          if (Hhi) {
            const size_t o = (size_t)(row * (uint32_t)ldhs + ccol);
            *reinterpret_cast<uint32_t*>(Hhi + o) = hh[jj];
            *reinterpret_cast<uint32_t*>(Hlo + o) = hl[jj];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  if (warp == 8 && grp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int tc_debug_flags() {  // timing experiments (wrong results on purpose): see `dbg` in the kernel
  static const int f = [] { const char* e = getenv("RFX_LSTM_TC_DEBUG"); return e ? (atoi(e) & 0xff) << 8 : 0; }();
  return f;
}

// two 16-slot groups per CTA: one cluster of 8 SMs per direction serves 32 slots, like <32>, with the groups overlapping each other
static int launch_tc_dual(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs, int B,
                          int F, int slots, cudaStream_t stream) {
  constexpr int N = TC_N;
  constexpr int BLK = 2 * N * 64;
  constexpr size_t group_smem = (size_t)2 * LSTM_CL * BLK + 2 * BLK + (size_t)4 * N * 32 * 4 + 256;
  const size_t smem = 2 * group_smem + 128;
  auto kern = lstm_rec_tc_kernel<N, 2>;
  RFX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nb = (slots > 0 && slots < 2 * N) ? slots : 2 * N;
  dim3 grid(LSTM_CL, ceil_div(B, nb), 2);
  kern<<<grid, 2 * TC_THREADS, smem, stream>>>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, nb, (get_matmul_precision() == 1 ? 1 : 0) | tc_debug_flags());
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int N>
static int launch_tc_n(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs, int B,
                       int F, int slots, cudaStream_t stream) {
  constexpr int BLK = 2 * N * 64;
  const size_t smem = (size_t)2 * LSTM_CL * BLK + 2 * BLK + (size_t)4 * N * 32 * 4 + 256 + 128;
  auto kern = lstm_rec_tc_kernel<N>;
  RFX_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nb = (slots > 0 && slots < N) ? slots : N;
  dim3 grid(LSTM_CL, ceil_div(B, nb), 2);
  kern<<<grid, TC_THREADS, smem, stream>>>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, nb, (get_matmul_precision() == 1 ? 1 : 0) | tc_debug_flags());
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_tc(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs, int B,
                     int F, int slots, cudaStream_t stream) {
  RFX_REQUIRE((long long)B * F * (long long)ldg < (1ll << 32) && (long long)B * F * (long long)std::max(ldh, ldhs) < (1ll << 32),
              "lstm: tensors too large for 32-bit element offsets");
  RFX_REQUIRE((!Hout || ldh % 2 == 0) && (!Hhi || ldhs % 2 == 0), "lstm: output row strides must be even");
  if (slots > TC_N) {
    // 32 slots per cluster: two interleaved 16-slot groups (default) or one 32-wide machine (RFX_LSTM_TC32_SINGLE=1)
    static const bool single = [] { const char* e = getenv("RFX_LSTM_TC32_SINGLE"); return e && atoi(e) != 0; }();
    if (!single) return launch_tc_dual(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
    return launch_tc_n<TC_N_MAX>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
  }
  return launch_tc_n<TC_N>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
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

static int g_lstm_impl = 0;  // 0 = tensor-core (mma.sync bf16x3), 1 = fp32 FFMA, 2 = tcgen05 (A from TMEM, 16 slots per cluster)
void lstm_set_impl(int impl) { g_lstm_impl = impl; }
int lstm_get_impl() { return g_lstm_impl; }

int lstm_clusters_for(int B, int slots) {
  const int s = slots <= 0 ? LSTM_SLOTS : (slots < TC_N_MAX ? slots : TC_N_MAX);
  return 2 * ceil_div(B, s);
}

int launch_lstm_layer(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo, int ldhs,
                      int B, int F, int H, cudaStream_t stream) {
  return launch_lstm_layer_slots(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, H, 0, stream);
}

int launch_lstm_layer_slots(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo,
                            int ldhs, int B, int F, int H, int slots, cudaStream_t stream) {
  return launch_lstm_layer_impl(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, H, -1, slots, stream);
}

int launch_lstm_layer_impl(const float* G, int ldg, const float* Whh, float* Hout, int ldh, __nv_bfloat16* Hhi, __nv_bfloat16* Hlo,
                           int ldhs, int B, int F, int H, int impl, int slots, cudaStream_t stream) {
  const int g_lstm_impl = impl >= 0 ? impl : rfx::lstm_get_impl();  // shadows the process-wide default below
  RFX_REQUIRE(H == 192 || H == 256 || H == 384, "lstm: hidden size per direction must be 192, 256 or 384");
  RFX_REQUIRE(B > 0 && F > 0, "lstm: positive sizes");
  RFX_REQUIRE(((uintptr_t)Whh & 15) == 0, "lstm: W_hh must be 16-byte aligned");
  RFX_REQUIRE(Hout || (Hhi && Hlo), "lstm: no output given");
  if (g_lstm_impl == 2 && H == LSTM_H) return launch_tc(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
  if (g_lstm_impl == 0 || H != LSTM_H) {
    RFX_REQUIRE(slots <= LSTM_SLOTS, "lstm: more than 8 slots per cluster needs the tcgen05 kernel (impl 2)");
    if (H == 256) return launch_mma<256, false>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
    if (H == 192) return launch_mma<192, false>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
    return launch_mma<384, true>(G, ldg, Whh, Hout, ldh, Hhi, Hlo, ldhs, B, F, slots, stream);
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
