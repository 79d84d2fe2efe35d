// TCN training path: forward that keeps every block's output, and the backward pass (remfx/tcn.py:48-59,126-130 under
// torch autograd, as the reference's Lightning step differentiates it: remfx/models.py:217-220 -> loss.backward()).
//
//   y_n = PReLU_c(z_n) + r_n,   z_n = b + sum_j W_j x_n[t + j d],   r_n = W_res x_n[t + off]           (block n, x_{n+1} = y_n)
//   out = tanh(w_o . y_last + b_o)
//
// Backward of one block n >= 1, given dy = dL/dy_n ([B][Lo][C] fp32):
//   1. z_n is RECOMPUTED from the saved input planes (gemm2, 7 taps, bias epilogue) -- only the split-bf16 block outputs are
//      kept by the forward, so training holds 4 bytes / activation instead of 8;
//   2. tcn_act_bwd_kernel:  dz = dy * (z > 0 ? 1 : slope_c)  -> split planes G[y=0] = dz, G[y=1] = dy;  db_c += sum dz,
//      dslope_c += sum_{z<=0} dy z;
//   3. tcn_wgrad_kernel:    dW_j[co][ci] = sum_{b,t} dz[b,t,co] x_n[b,t+jd,ci],  dW_res = sum dy x_n[.. + off]: contraction over
//      TIME, both operands time-major in HBM ([t][c]) -> mma.sync.m16n8k16 bf16x3 with ldmatrix.trans fragments, cp.async
//      3-stage ring, split over (item, time chunk) with fp32 atomics into a [K+1][C][C] accumulator;
//   4. input gradient = the forward engine (gemm2, tcgen05) on the TRANSPOSED packed weights with negated tap offsets; the
//      residual tap reads dy instead of dz by addressing G as a 2-row (y) tensor:  dx[s] = sum_j W_j^T dz[s - jd] + W_res^T dy[s - off].
// Block 0 (1 input channel) and the tanh tail are SIMT kernels.  Reductions use fp32 atomics: results are reproducible to
// rounding, not bit-for-bit.
#include "tcn_internal.h"

int rfx_encode_tiled_bf16(void* map, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                          const unsigned* box, int swizzle128);  // gemm2.cu

namespace rfx {

__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
    pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  lo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}
__device__ __forceinline__ void join8(const uint4& h, const uint4& l, float (&v)[8]) {
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw[i]));
    const float2 lf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&lw[i]));
    v[2 * i] = hf.x + lf.x;
    v[2 * i + 1] = hf.y + lf.y;
  }
}

// Per-channel sums of a CTA whose threads each own 8 consecutive channels: shared-memory atomics, then one global atomic
// per channel and CTA.  `sm` holds nvec * C floats, zeroed by the caller before the first use.
template <int NVEC>
__device__ __forceinline__ void cta_channel_sums(float* sm, int C, int c0, const float (&acc)[NVEC][8], float* const (&dst)[NVEC]) {
#pragma unroll
  for (int v = 0; v < NVEC; ++v)
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&sm[v * C + c0 + i], acc[v][i]);
  __syncthreads();
  for (int i = threadIdx.x; i < NVEC * C; i += blockDim.x) {
    const int v = i / C;
    if (dst[v]) atomicAdd(dst[v] + (i - v * C), sm[i]);
  }
}

// ---- tail backward: du = dout (1 - out^2);  dy[b][t][c] = du w_o[c];  dw_o[c] += du y_last[b][t][c];  db_o += du ----
__global__ void __launch_bounds__(256) tcn_tail_bwd_kernel(const float* __restrict__ out, const float* __restrict__ dout, long long o_bs, int L, int C,
                                                           const __nv_bfloat16* __restrict__ yhi, const __nv_bfloat16* __restrict__ ylo, long long y_bs,
                                                           const float* __restrict__ w, float* __restrict__ dy, float* __restrict__ dw,
                                                           float* __restrict__ db, int rows_per_cta) {
  extern __shared__ float sm[];
  const int groups = C / 8, rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups, c0 = g * 8;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C + 1; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();
  float acc[1][8] = {};
  float wv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) wv[i] = w[c0 + i];
  float dbl = 0.0f;
  const int t_begin = blockIdx.x * rows_per_cta, t_end = min(L, t_begin + rows_per_cta);
  if (r < rows) {
    for (int t = t_begin + r; t < t_end; t += rows) {
      const float o = out[(size_t)b * o_bs + t];
      const float du = dout[(size_t)b * o_bs + t] * (1.0f - o * o);
      const size_t off = (size_t)b * y_bs + (size_t)t * C + c0;
      float yv[8];
      join8(*reinterpret_cast<const uint4*>(yhi + off), *reinterpret_cast<const uint4*>(ylo + off), yv);
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0][i] = fmaf(du, yv[i], acc[0][i]);
        d[i] = du * wv[i];
      }
      *reinterpret_cast<float4*>(dy + off) = make_float4(d[0], d[1], d[2], d[3]);
      *reinterpret_cast<float4*>(dy + off + 4) = make_float4(d[4], d[5], d[6], d[7]);
      if (g == 0) dbl += du;
    }
  }
  if (g == 0 && r < rows) atomicAdd(&sm[C], dbl);
  float* const dst[1] = {dw};
  cta_channel_sums<1>(sm, C, c0, acc, dst);  // (threads with r >= rows add zeros)
  if (threadIdx.x == 0) atomicAdd(db, sm[C]);
}

// ---- PReLU backward + operand split: G[y=0] = dz, G[y=1] = dy as split planes; bias / slope gradients ----
__global__ void __launch_bounds__(256) tcn_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, long long f_bs, int L, int C,
                                                          const float* __restrict__ slope, __nv_bfloat16* __restrict__ G, long long g_bs,
                                                          long long g_ldy, long long g_plane, float* __restrict__ dbias,
                                                          float* __restrict__ dslope, int rows_per_cta) {
  extern __shared__ float sm[];
  const int groups = C / 8, rows = blockDim.x / groups;
  const int g = threadIdx.x % groups, r = threadIdx.x / groups, c0 = g * 8;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();
  float acc[2][8] = {};
  float sl[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sl[i] = slope[c0 + i];
  const int t_begin = blockIdx.x * rows_per_cta, t_end = min(L, t_begin + rows_per_cta);
  if (r < rows) {
    for (int t = t_begin + r; t < t_end; t += rows) {
      const size_t off = (size_t)b * f_bs + (size_t)t * C + c0;
      const float4 d0 = *reinterpret_cast<const float4*>(dy + off), d1 = *reinterpret_cast<const float4*>(dy + off + 4);
      const float4 z0 = *reinterpret_cast<const float4*>(z + off), z1 = *reinterpret_cast<const float4*>(z + off + 4);
      const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
      float dz[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pos = zv[i] > 0.0f;  // torch's prelu backward: x > 0 ? g : w g;  dw += x > 0 ? 0 : x g
        dz[i] = pos ? dv[i] : dv[i] * sl[i];
        acc[0][i] += dz[i];
        acc[1][i] += pos ? 0.0f : dv[i] * zv[i];
      }
      uint4 hi, lo;
      const size_t go = (size_t)b * g_bs + (size_t)t * C + c0;
      split8(dz, hi, lo);
      *reinterpret_cast<uint4*>(G + go) = hi;
      *reinterpret_cast<uint4*>(G + g_plane + go) = lo;
      split8(dv, hi, lo);
      *reinterpret_cast<uint4*>(G + g_ldy + go) = hi;
      *reinterpret_cast<uint4*>(G + g_plane + g_ldy + go) = lo;
    }
  }
  float* const dst[2] = {dbias, dslope};
  cta_channel_sums<2>(sm, C, c0, acc, dst);
}

// ---- block 0 backward (one input channel): thread = output channel, CTA = a range of time steps of one item ----
// z = b_c + sum_j w[c][j] x[t + j d];  dz = dy prelu'(z);  dw[c][j] += dz x[t + j d];  dwres[c] += dy x[t + off];  db, dslope.
__global__ void __launch_bounds__(256) tcn_first_bwd_kernel(const float* __restrict__ x, long long x_bs, const float* __restrict__ dy, long long f_bs,
                                                            int L1, int C, int K, int dil, int res_off, const float* __restrict__ w,
                                                            const float* __restrict__ bias, const float* __restrict__ slope, float* __restrict__ dw,
                                                            float* __restrict__ dwres, float* __restrict__ dbias, float* __restrict__ dslope,
                                                            int rows_per_cta) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const int b = blockIdx.y;
  float wk[16], sw[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    wk[j] = j < K ? w[c * K + j] : 0.0f;
    sw[j] = 0.0f;
  }
  const float bc = bias[c], sl = slope[c];
  float sres = 0.0f, sb = 0.0f, ss = 0.0f;
  const int t_begin = blockIdx.x * rows_per_cta, t_end = min(L1, t_begin + rows_per_cta);
  const float* xr = x + (size_t)b * x_bs;
  for (int t = t_begin; t < t_end; ++t) {
    const float g = dy[(size_t)b * f_bs + (size_t)t * C + c];
    float xs[16];
    float zc = bc;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      xs[j] = j < K ? xr[t + j * dil] : 0.0f;
      zc = fmaf(wk[j], xs[j], zc);
    }
    const bool pos = zc > 0.0f;
    const float dz = pos ? g : g * sl;
#pragma unroll
    for (int j = 0; j < 16; ++j) sw[j] = fmaf(dz, xs[j], sw[j]);
    sres = fmaf(g, xr[t + res_off], sres);
    sb += dz;
    ss += pos ? 0.0f : g * zc;
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < K) atomicAdd(dw + c * K + j, sw[j]);
  atomicAdd(dwres + c, sres);
  atomicAdd(dbias + c, sb);
  atomicAdd(dslope + c, ss);
}

// ---- weight gradient: contraction over time on mma.sync (bf16x3) ----
constexpr int WG_BM = 128;      // output channels (co) per CTA
constexpr int WG_BN = 128;      // input channels (ci) per CTA
constexpr int WG_BK = 32;       // time steps per pipeline stage
constexpr int WG_LD = 136;      // padded shared-memory row (elements): 272 bytes -> conflict-free ldmatrix
constexpr int WG_STAGES = 3;
constexpr int WG_TILE = WG_BK * WG_LD;                 // elements of one plane tile
constexpr int WG_STAGE_ELEMS = 4 * WG_TILE;            // G hi, G lo, X hi, X lo
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_ELEMS * 2;  // bytes

struct WgParams {
  const __nv_bfloat16* g;  // gradient planes: element (b, y, t, co) at g[b * g_bs + y * g_ldy + t * C + co], lo plane at + g_plane
  long long g_bs, g_ldy, g_plane;
  const __nv_bfloat16* x;  // block input planes: (b, t, ci) at x[b * x_bs + t * C + ci], lo plane at + x_plane
  long long x_bs, x_plane;
  int C, K, Lo, nchunks, tchunk;
  int off[16];             // x time offset of tap j (j = K: residual tap, which pairs with the y = 1 rows of g)
  float* dW;               // [K + 1][C][C] (tap, co, ci), accumulated with atomics
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 2) tcn_wgrad_kernel(const WgParams p) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  __nv_bfloat16* sm = reinterpret_cast<__nv_bfloat16*>(wg_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = p.C;
  const int tiles = (C + WG_BM - 1) / WG_BM;
  int bx = blockIdx.x;
  const int tn = bx % tiles; bx /= tiles;
  const int tm = bx % tiles;
  const int tap = bx / tiles;
  const int b = blockIdx.y / p.nchunks, chunk = blockIdx.y % p.nchunks;
  const int t_begin = chunk * p.tchunk, t_end = min(p.Lo, t_begin + p.tchunk);
  const int co0 = tm * WG_BM, ci0 = tn * WG_BN;
  const __nv_bfloat16* gsrc = p.g + (size_t)b * p.g_bs + (tap == p.K ? p.g_ldy : 0);
  const __nv_bfloat16* xsrc = p.x + (size_t)b * p.x_bs + (long long)p.off[tap] * C;
  const int iters = (t_end - t_begin + WG_BK - 1) / WG_BK;

  auto load_stage = [&](int it, int stage) {
    const int t0 = t_begin + it * WG_BK;
    __nv_bfloat16* st = sm + (size_t)stage * WG_STAGE_ELEMS;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int idx = tid + h * 256;  // 512 16-byte pieces per plane tile: 32 rows x 16
      const int r = idx >> 4, cc = (idx & 15) * 8;
      const int t = t0 + r;
      const bool row_ok = t < t_end;
      const bool gok = row_ok && (co0 + cc < C), xok = row_ok && (ci0 + cc < C);
      const __nv_bfloat16* gp = gok ? gsrc + (size_t)t * C + co0 + cc : p.g;
      const __nv_bfloat16* xp = xok ? xsrc + (size_t)t * C + ci0 + cc : p.x;
      const uint32_t d = smem_u32(st + r * WG_LD + cc);
      cp_async16(d, gp, gok);
      cp_async16(d + WG_TILE * 2, gok ? gp + p.g_plane : p.g, gok);
      cp_async16(d + 2 * WG_TILE * 2, xp, xok);
      cp_async16(d + 3 * WG_TILE * 2, xok ? xp + p.x_plane : p.x, xok);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.0f;

  for (int s = 0; s < WG_STAGES - 1; ++s) {
    if (s < iters) load_stage(s, s);
    cp_async_commit();
  }
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps: 64 co x 32 ci each
  const int lj = lane >> 3, lr = lane & 7;
  // ldmatrix.trans source rows are time steps.  A (co x t) fragment order: (m lo, k lo), (m hi, k lo), (m lo, k hi), (m hi, k hi);
  // B (t x ci) fragments for two n-tiles: (k lo, n0), (k hi, n0), (k lo, n1), (k hi, n1).
  const int a_row = (lj >> 1) * 8 + lr, a_col = wm * 64 + (lj & 1) * 8;
  const int b_row = (lj & 1) * 8 + lr, b_col = wn * 32 + (lj >> 1) * 8;

  for (int it = 0; it < iters; ++it) {
    cp_async_wait<WG_STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + WG_STAGES - 1;
      if (nx < iters) load_stage(nx, nx % WG_STAGES);
      cp_async_commit();
    }
    const __nv_bfloat16* st = sm + (size_t)(it % WG_STAGES) * WG_STAGE_ELEMS;
    const uint32_t g_hi = smem_u32(st), g_lo = g_hi + WG_TILE * 2, x_hi = g_hi + 2 * WG_TILE * 2, x_lo = g_hi + 3 * WG_TILE * 2;
#pragma unroll
    for (int kk = 0; kk < WG_BK; kk += 16) {
      uint32_t bh[2][4], bl[2][4];
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        const uint32_t o = (uint32_t)(((kk + b_row) * WG_LD + b_col + np * 16) * 2);
        ldsm_x4_t(x_hi + o, bh[np]);
        ldsm_x4_t(x_lo + o, bl[np]);
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        uint32_t ah[4], al[4];
        const uint32_t o = (uint32_t)(((kk + a_row) * WG_LD + a_col + mt * 16) * 2);
        ldsm_x4_t(g_hi + o, ah);
        ldsm_x4_t(g_lo + o, al);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int np = nt >> 1, q = (nt & 1) * 2;
          mma16816(acc[mt][nt], al, bh[np][q], bh[np][q + 1]);
          mma16816(acc[mt][nt], ah, bl[np][q], bl[np][q + 1]);
          mma16816(acc[mt][nt], ah, bh[np][q], bh[np][q + 1]);
        }
      }
    }
  }
  cp_async_wait<0>();

  float* dst = p.dW + (size_t)tap * C * C;
  const int gq = lane >> 2, qq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int co = co0 + wm * 64 + mt * 16 + gq;
      const int ci = ci0 + wn * 32 + nt * 8 + qq * 2;
      if (ci < C) {  // C is a multiple of 8: ci + 1 < C as well
        if (co < C) {
          atomicAdd(dst + (size_t)co * C + ci, acc[mt][nt][0]);
          atomicAdd(dst + (size_t)co * C + ci + 1, acc[mt][nt][1]);
        }
        if (co + 8 < C) {
          atomicAdd(dst + (size_t)(co + 8) * C + ci, acc[mt][nt][2]);
          atomicAdd(dst + (size_t)(co + 8) * C + ci + 1, acc[mt][nt][3]);
        }
      }
    }
}

// ---- weight gradient on tcgen05 (C = 256): MN-major shared-memory operands, both planes of a box in one TMA load ----
// Contraction over TIME with both operands stored [t][c]: for the MMA that is "MN-major" (the M / N index is the contiguous one).
// A TMA box of {64 channels, 32 time steps, 1 item, 2 planes} with SWIZZLE_128B lands in shared memory as the canonical MN-major
// SW128 layout: 128-byte rows = 64 channels of one time step, 8 time steps per 1024-byte swizzle atom (SBO = 1024 B), the next
// block of 64 channels one box further (LBO = box bytes); instruction descriptor bits 15 / 16 = MN-major A / B.
// Roles are swapped relative to the maths so that the epilogue's atomics coalesce: M = ci (x tile, 2 halves of 128), N = co (g tile,
// 256), D[ci][co] in TMEM (2 x 256 columns = all 512); a thread (= TMEM lane = ci) adds 32 consecutive co columns to dW[co][ci].
// One CTA = (tap, time chunk, item); warp 0 TMA producer, warp 1 MMA issuer (bf16x3: lo*hi + hi*lo + hi*hi), warps 2-5 epilogue.
// Brought up as tools/wgrad_tc_probe.cu (profiles/r2/wgrad_tc_probe.log: 2.5e-6 vs the fp32 sums, 0.45 ms against 2.32 ms for the
// mma.sync kernel at 1 x 249868).
constexpr int WT_C = 256, WT_BK = 32, WT_STAGES = 3;
constexpr int WT_BOX = 64 * WT_BK * 2 * 2;   // 64 channels x 32 steps x (hi, lo) = 8 KB
constexpr int WT_PLANE = 64 * WT_BK * 2;     // lo plane offset inside a box
constexpr int WT_STAGE = 8 * WT_BOX;         // 4 x-boxes (ci) + 4 g-boxes (co)
constexpr int WT_SMEM = WT_STAGES * WT_STAGE + 1024 + 256;

struct WtParams {
  int L;          // valid time steps of g (rows beyond are TMA zero fill)
  int tchunk, nchunks;
  int K;          // conv taps (tap K = residual, reads the dy rows through mapG1)
  int off[16];
  float* dW;      // [K + 1][C][C]
};
struct alignas(64) TmapBytes { unsigned char b[128]; };

__device__ __forceinline__ void wt_tma_load_4d(void* dst, const void* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t wt_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1) tcn_wgrad_tc_kernel(const __grid_constant__ TmapBytes mapG0, const __grid_constant__ TmapBytes mapG1,
                                                              const __grid_constant__ TmapBytes mapX, const WtParams p) {
  extern __shared__ uint8_t wt_smem_raw[];
  const uint32_t raw = smem_u32(wt_smem_raw);
  uint8_t* smem = wt_smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WT_STAGES * WT_STAGE);
  uint64_t* empty_bar = full_bar + WT_STAGES;
  uint64_t* tfull_bar = empty_bar + WT_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
  const int t_begin = chunk * p.tchunk;
  const int t_end = min(p.L, t_begin + p.tchunk);
  const int iters = (t_end - t_begin + WT_BK - 1) / WT_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const void* mg = tap == p.K ? (const void*)&mapG1 : (const void*)&mapG0;
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], WT_STAGE);
        uint8_t* st = smem + s * WT_STAGE;
        const int t0 = t_begin + it * WT_BK;
        for (int cb = 0; cb < 4; ++cb) wt_tma_load_4d(st + cb * WT_BOX, &mapX, cb * 64, t0 + p.off[tap], b, 0, &full_bar[s]);   // x: ci blocks
        for (int cb = 0; cb < 4; ++cb) wt_tma_load_4d(st + (4 + cb) * WT_BOX, mg, cb * 64, t0, b, 0, &full_bar[s]);             // g: co blocks
        if (++s == WT_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && iters > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256) | (1u << 15) | (1u << 16);   // M = 128 (ci half), N = 256 (co), MN-major A and B
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t xb = smem_u32(smem + s * WT_STAGE), gb = xb + 4 * WT_BOX;
#pragma unroll
        for (int kk = 0; kk < WT_BK / 16; ++kk) {
          const uint32_t ko = kk * 2048;  // 16 time steps = two 1024-byte atoms
          const uint64_t g_hi = wt_desc_mn_sw128(gb + ko, WT_BOX), g_lo = wt_desc_mn_sw128(gb + WT_PLANE + ko, WT_BOX);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t x_hi = wt_desc_mn_sw128(xb + h * 2 * WT_BOX + ko, WT_BOX), x_lo = wt_desc_mn_sw128(xb + h * 2 * WT_BOX + WT_PLANE + ko, WT_BOX);
            const uint32_t d = tmem + h * 256;
            const uint32_t first = (it == 0 && kk == 0) ? 0u : 1u;
            umma_f16(d, x_lo, g_hi, idesc, first);
            umma_f16(d, x_hi, g_lo, idesc, 1u);
            umma_f16(d, x_hi, g_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (++s == WT_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else if (iters > 0) {
    // epilogue: warp w may only read TMEM lanes [32 (w % 4), +32)
    const int q = warp & 3;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* dst = p.dW + (size_t)tap * WT_C * WT_C;
    for (int h = 0; h < 2; ++h) {
      const int ci = h * 128 + q * 32 + lane;
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 256 + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)(cc * 32 + j) * WT_C + ci, __uint_as_float(v[j]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// plain fp32 FFMA form of the same contraction (cross-check only, rfx_tcn_set_wgrad_impl(1)): thread = (tap, co, ci)
__global__ void tcn_wgrad_simt_kernel(const WgParams p, int B) {
  const int C = p.C;
  const long long total = (long long)(p.K + 1) * C * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C), co = (int)((i / C) % C), tap = (int)(i / ((long long)C * C));
    float acc = 0.0f;
    for (int b = 0; b < B; ++b) {
      const __nv_bfloat16* g = p.g + (size_t)b * p.g_bs + (tap == p.K ? p.g_ldy : 0) + co;
      const __nv_bfloat16* x = p.x + (size_t)b * p.x_bs + (long long)p.off[tap] * C + ci;
      for (int t = 0; t < p.Lo; ++t) {
        const float gv = __bfloat162float(g[(size_t)t * C]) + __bfloat162float(g[p.g_plane + (size_t)t * C]);
        const float xv = __bfloat162float(x[(size_t)t * C]) + __bfloat162float(x[p.x_plane + (size_t)t * C]);
        acc = fmaf(gv, xv, acc);
      }
    }
    p.dW[i] = acc;
  }
}

// dWcat [K+1][co][ci] -> conv1.weight.grad [co][ci][K] and res.weight.grad [co][ci]
__global__ void tcn_scatter_wgrad_kernel(const float* __restrict__ dW, int C, int K, float* __restrict__ gconv, float* __restrict__ gres) {
  const long long total = (long long)(K + 1) * C * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C), co = (int)((i / C) % C), tap = (int)(i / ((long long)C * C));
    if (tap < K) gconv[((size_t)co * C + ci) * K + tap] = dW[i];
    else gres[(size_t)co * C + ci] = dW[i];
  }
}

// WcatT [ci][tap * C + co] = tap < K ? conv1.weight[co][ci][tap] : res.weight[co][ci]
__global__ void tcn_gather_wt_kernel(const float* __restrict__ wconv, const float* __restrict__ wres, int C, int K, float* __restrict__ wcat) {
  const long long total = (long long)C * (K + 1) * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % C);
    const int tap = (int)((i / C) % (K + 1));
    const int ci = (int)(i / ((long long)C * (K + 1)));
    wcat[i] = tap < K ? wconv[((size_t)co * C + ci) * K + tap] : wres[(size_t)co * C + ci];
  }
}

namespace {

int g_tcn_wgrad_impl = 0;  // 0 = tcgen05 bf16x3 when C == 256, else mma.sync (product); 1 = fp32 SIMT cross-check; 2 = mma.sync bf16x3

struct TrainLayout {
  size_t plane_bytes;  // one bf16 plane of a saved block output
  size_t f_bytes;      // one fp32 [B][L1][C] buffer
  size_t saved_off, f_off[2], z_off, g_off, dw_off, total;
};
TrainLayout train_layout(const rfx_tcn* h, int B, long long T) {
  TrainLayout l{};
  const int C = h->cfg.channel_width, K = h->cfg.kernel_size, NBk = h->cfg.nblocks;
  const size_t elems = (size_t)B * (size_t)tcn_len_after(h, T, 1) * C;
  l.plane_bytes = tcn_plane_bytes(h, B, T);
  l.f_bytes = align_up(elems * 4, 256);
  size_t o = 0;
  l.saved_off = o; o += (size_t)NBk * 2 * l.plane_bytes;
  l.f_off[0] = o; o += l.f_bytes;
  l.f_off[1] = o; o += l.f_bytes;
  l.z_off = o; o += l.f_bytes;
  l.g_off = o; o += 4 * l.plane_bytes;  // (hi, lo) x (dz, dy)
  l.dw_off = o; o += align_up((size_t)(K + 1) * C * C * 4, 256);
  l.total = o;
  return l;
}

int ensure_transposed(rfx_tcn* h, cudaStream_t s) {
  if (h->transposed_ready) return 0;
  const int C = h->cfg.channel_width, K = h->cfg.kernel_size, NBk = h->cfg.nblocks;
  if ((int)h->wsplit_t.size() != NBk) {
    for (auto& b : h->wsplit_t) b.release();
    h->wsplit_t.assign(NBk, TcnBuf());
  }
  h->wpack_t.assign(NBk, SplitW());
  if (NBk <= 1) { h->transposed_ready = true; return 0; }
  if (h->wcat.alloc((size_t)C * (K + 1) * C)) return 1;  // stream-ordered reuse of the finalize staging buffer
  const int BN = g2_choose_bn(C);
  for (int n = 1; n < NBk; ++n) {
    const std::string p = "process_blocks." + std::to_string(n);
    tcn_gather_wt_kernel<<<148 * 4, 256, 0, s>>>(tcn_param(h, p + ".conv1.weight"), tcn_param(h, p + ".res.weight"), C, K, h->wcat.p);
    RFX_CHECK_CUDA(cudaGetLastError());
    if (h->wsplit_t[n].alloc(split_weight_elems(C, (K + 1) * C, BN))) return 1;
    int rc = pack_split_weights(h->wcat.p, (long long)(K + 1) * C, C, (K + 1) * C, BN, reinterpret_cast<__nv_bfloat16*>(h->wsplit_t[n].p), &h->wpack_t[n], s);
    if (rc) return rc;
  }
  h->transposed_ready = true;
  return 0;
}

int rows_per_cta_for(long long L, int B) {
  // about four waves of CTAs over the chip, at least 64 rows each
  const long long want = (148ll * 4 + B - 1) / B;
  long long rows = (L + want - 1) / want;
  if (rows < 64) rows = 64;
  return (int)rows;
}

}  // namespace
}  // namespace rfx

using namespace rfx;

extern "C" {

size_t rfx_tcn_train_workspace_bytes(const rfx_tcn_t* h, int B, long long T) {
  if (!h || B <= 0 || tcn_len_after(h, T, h->cfg.nblocks) <= 0) return 0;
  return train_layout(h, B, T).total;
}

int rfx_tcn_forward_train(rfx_tcn_t* h, const float* x, int B, long long T, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(h && x && out && workspace, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_tcn_finalize has not been called since the last parameter load");
  RFX_REQUIRE(B > 0 && tcn_len_after(h, T, h->cfg.nblocks) > 0, "input shorter than the receptive field");
  RFX_REQUIRE(T < (1ll << 31), "T too large");
  RFX_REQUIRE(workspace_bytes >= rfx_tcn_train_workspace_bytes(h, B, T), "workspace too small (rfx_tcn_train_workspace_bytes)");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const TrainLayout l = train_layout(h, B, T);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  std::vector<__nv_bfloat16*> outs(h->cfg.nblocks);
  for (int n = 0; n < h->cfg.nblocks; ++n) outs[n] = reinterpret_cast<__nv_bfloat16*>(ws + l.saved_off + (size_t)n * 2 * l.plane_bytes);
  return tcn_run_forward(h, x, B, T, out, outs.data(), (long long)(l.plane_bytes / 2), (cudaStream_t)stream);
}

int rfx_tcn_backward(rfx_tcn_t* h, const float* x, const float* out, const float* dout, int B, long long T, const char* const* keys,
                     float* const* grads, int nkeys, void* workspace, size_t workspace_bytes, void* stream) {
  RFX_REQUIRE(h && x && out && dout && workspace && keys && grads, "null argument");
  RFX_REQUIRE(h->finalized, "rfx_tcn_finalize has not been called since the last parameter load");
  const int C = h->cfg.channel_width, K = h->cfg.kernel_size, NBk = h->cfg.nblocks;
  const long long Lout = tcn_len_after(h, T, NBk);
  RFX_REQUIRE(B > 0 && Lout > 0, "input shorter than the receptive field");
  RFX_REQUIRE(T < (1ll << 31) && B <= 65535, "T or B too large");
  RFX_REQUIRE(workspace_bytes >= rfx_tcn_train_workspace_bytes(h, B, T), "workspace too small (rfx_tcn_train_workspace_bytes)");
  RFX_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  std::map<std::string, float*> gmap;
  for (int i = 0; i < nkeys; ++i) {
    RFX_REQUIRE(keys[i] && grads[i], "null gradient key / pointer");
    gmap[keys[i]] = grads[i];
  }
  auto grad = [&](const std::string& k) -> float* {
    auto it = gmap.find(k);
    return it == gmap.end() ? nullptr : it->second;
  };
  // every parameter must have a destination; all of them are overwritten (zeroed here, accumulated by the kernels)
  for (int n = 0; n < NBk; ++n) {
    const std::string p = "process_blocks." + std::to_string(n);
    const int cin = n == 0 ? 1 : C;
    const std::pair<const char*, size_t> need[4] = {{".conv1.weight", (size_t)C * cin * K}, {".conv1.bias", (size_t)C}, {".res.weight", (size_t)C * cin},
                                                    {".relu.weight", (size_t)C}};
    for (const auto& kv : need) {
      float* g = grad(p + kv.first);
      if (!g) { set_error("rfx_tcn_backward: no gradient buffer for '" + p + kv.first + "'"); return 2; }
      RFX_CHECK_CUDA(cudaMemsetAsync(g, 0, kv.second * sizeof(float), s));
    }
  }
  float* g_ow = grad("output.weight");
  float* g_ob = grad("output.bias");
  RFX_REQUIRE(g_ow && g_ob, "no gradient buffer for output.weight / output.bias");
  RFX_CHECK_CUDA(cudaMemsetAsync(g_ow, 0, (size_t)C * sizeof(float), s));
  RFX_CHECK_CUDA(cudaMemsetAsync(g_ob, 0, sizeof(float), s));
  int rc;
  if ((rc = ensure_transposed(h, s))) return rc;

  const TrainLayout l = train_layout(h, B, T);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const long long plane_elems = (long long)(l.plane_bytes / 2);
  auto saved = [&](int n) { return reinterpret_cast<__nv_bfloat16*>(ws + l.saved_off + (size_t)n * 2 * l.plane_bytes); };
  float* F[2] = {reinterpret_cast<float*>(ws + l.f_off[0]), reinterpret_cast<float*>(ws + l.f_off[1])};
  float* Z = reinterpret_cast<float*>(ws + l.z_off);
  __nv_bfloat16* G = reinterpret_cast<__nv_bfloat16*>(ws + l.g_off);
  float* dWcat = reinterpret_cast<float*>(ws + l.dw_off);
  const long long L1 = tcn_len_after(h, T, 1);
  const long long bs = L1 * C;
  // G: element (plane, b, y, t, c) at plane * g_plane + b * g_bs + y * g_ldy + t * C + c
  const long long g_ldy = bs, g_bs = 2 * bs, g_plane = 2 * plane_elems;
  const bool simt_wgrad = g_tcn_wgrad_impl == 1;
  const bool tc_wgrad = g_tcn_wgrad_impl == 0 && C == WT_C && K + 1 <= 16 && B <= 65535;
  if (tc_wgrad) RFX_CHECK_CUDA(cudaFuncSetAttribute(tcn_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
  const int groups = C / 8;
  const int ew_threads = (256 / groups) * groups;  // whole time rows per CTA
  RFX_CHECK_CUDA(cudaFuncSetAttribute(tcn_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));

  int cur = 0;
  // tail
  {
    const int rows = rows_per_cta_for(Lout, B);
    dim3 grid((unsigned)((Lout + rows - 1) / rows), B);
    tcn_tail_bwd_kernel<<<grid, ew_threads, (C + 1) * sizeof(float), s>>>(out, dout, Lout, (int)Lout, C, saved(NBk - 1), saved(NBk - 1) + plane_elems, bs,
                                                                         tcn_param(h, "output.weight"), F[cur], g_ow, g_ob, rows);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  for (int n = NBk - 1; n >= 1; --n) {
    const int d = tcn_dilation_of(h, n);
    const long long Lin = tcn_len_after(h, T, n), Lo = Lin - (long long)(K - 1) * d;
    const std::string p = "process_blocks." + std::to_string(n);
    // 1. z_n = b + sum_j W_j x_n[t + j d]
    {
      G2Problem pr;
      pr.A.hi = saved(n - 1); pr.A.rows = Lin; pr.A.ld = C; pr.A.batch_stride = bs; pr.A.plane_stride = plane_elems;
      pr.W = h->wpack[n];
      pr.M = (int)Lo; pr.N = C; pr.batch = B; pr.Ktap = C; pr.taps = K;
      for (int j = 0; j < K; ++j) pr.row_off[j] = j * d;
      pr.Cf = Z; pr.ldcf = C; pr.bscf = bs;
      pr.epi.t1 = tcn_param(h, p + ".conv1.bias");
      if ((rc = launch_gemm2(pr, s))) return rc;
    }
    // 2. dz, operand planes, bias / slope gradients
    {
      const int rows = rows_per_cta_for(Lo, B);
      dim3 grid((unsigned)((Lo + rows - 1) / rows), B);
      tcn_act_bwd_kernel<<<grid, ew_threads, 2 * C * sizeof(float), s>>>(F[cur], Z, bs, (int)Lo, C, tcn_param(h, p + ".relu.weight"), G, g_bs, g_ldy,
                                                                        g_plane, grad(p + ".conv1.bias"), grad(p + ".relu.weight"), rows);
      RFX_CHECK_CUDA(cudaGetLastError());
    }
    // 3. weight gradients
    {
      WgParams wp{};
      wp.g = G; wp.g_bs = g_bs; wp.g_ldy = g_ldy; wp.g_plane = g_plane;
      wp.x = saved(n - 1); wp.x_bs = bs; wp.x_plane = plane_elems;
      wp.C = C; wp.K = K; wp.Lo = (int)Lo;
      for (int j = 0; j < K; ++j) wp.off[j] = j * d;
      wp.off[K] = tcn_res_off(h, d);
      wp.dW = dWcat;
      const int tiles = ceil_div(C, WG_BM);
      const int per_t = (K + 1) * tiles * tiles;
      if (simt_wgrad) {
        tcn_wgrad_simt_kernel<<<148 * 8, 256, 0, s>>>(wp, B);
      } else if (tc_wgrad) {
        // tensor maps of this block's operands: {channel, time, item, plane}; rows beyond the valid length are zero-filled by the TMA unit
        TmapBytes mg0, mg1, mx;
        const unsigned box[4] = {64, (unsigned)WT_BK, 1, 2};
        const unsigned long long gd[4] = {(unsigned long long)C, (unsigned long long)Lo, (unsigned long long)B, 2};
        const unsigned long long gst[3] = {(unsigned long long)C * 2, (unsigned long long)g_bs * 2, (unsigned long long)g_plane * 2};
        const unsigned long long xd[4] = {(unsigned long long)C, (unsigned long long)Lin, (unsigned long long)B, 2};
        const unsigned long long xst[3] = {(unsigned long long)C * 2, (unsigned long long)bs * 2, (unsigned long long)plane_elems * 2};
        if ((rc = rfx_encode_tiled_bf16(&mg0, G, 4, gd, gst, box, 1))) return rc;
        if ((rc = rfx_encode_tiled_bf16(&mg1, G + g_ldy, 4, gd, gst, box, 1))) return rc;
        if ((rc = rfx_encode_tiled_bf16(&mx, saved(n - 1), 4, xd, xst, box, 1))) return rc;
        WtParams tp{};
        tp.L = (int)Lo; tp.K = K;
        for (int j = 0; j <= K; ++j) tp.off[j] = wp.off[j];
        long long want = (148ll * 2 + (long long)(K + 1) * B - 1) / ((long long)(K + 1) * B);  // about two waves of CTAs
        if (want < 1) want = 1;
        long long tchunk = ((Lo + want - 1) / want + WT_BK - 1) / WT_BK * WT_BK;
        if (tchunk < 8 * WT_BK) tchunk = 8 * WT_BK;
        tp.tchunk = (int)tchunk;
        tp.nchunks = (int)((Lo + tchunk - 1) / tchunk);
        tp.dW = dWcat;
        RFX_CHECK_CUDA(cudaMemsetAsync(dWcat, 0, (size_t)(K + 1) * C * C * sizeof(float), s));
        tcn_wgrad_tc_kernel<<<dim3(K + 1, tp.nchunks, B), 192, WT_SMEM, s>>>(mg0, mg1, mx, tp);
      } else {
        // time chunks: about three waves of CTAs at two CTAs per SM, whole stages each
        long long want = (148ll * 6 + (long long)per_t * B - 1) / ((long long)per_t * B);
        if (want < 1) want = 1;
        long long tchunk = (Lo + want - 1) / want;
        tchunk = (tchunk + WG_BK - 1) / WG_BK * WG_BK;
        if (tchunk < 8 * WG_BK) tchunk = 8 * WG_BK;
        wp.tchunk = (int)tchunk;
        wp.nchunks = (int)((Lo + tchunk - 1) / tchunk);
        RFX_CHECK_CUDA(cudaMemsetAsync(dWcat, 0, (size_t)(K + 1) * C * C * sizeof(float), s));
        dim3 grid(per_t, (unsigned)(B * wp.nchunks));
        tcn_wgrad_kernel<<<grid, 256, WG_SMEM, s>>>(wp);
      }
      RFX_CHECK_CUDA(cudaGetLastError());
      tcn_scatter_wgrad_kernel<<<148 * 4, 256, 0, s>>>(dWcat, C, K, grad(p + ".conv1.weight"), grad(p + ".res.weight"));
      RFX_CHECK_CUDA(cudaGetLastError());
    }
    // 4. dL/dx_n = dL/dy_{n-1}
    {
      G2Problem pr;
      pr.A.hi = G; pr.A.rows = Lo; pr.A.rows_y = 2; pr.A.ld = C; pr.A.ld_y = g_ldy; pr.A.batch_stride = g_bs; pr.A.plane_stride = g_plane;
      pr.W = h->wpack_t[n];
      pr.M = (int)Lin; pr.N = C; pr.batch = B; pr.Ktap = C; pr.taps = K + 1;
      for (int j = 0; j < K; ++j) pr.row_off[j] = -j * d;
      pr.row_off[K] = -tcn_res_off(h, d);
      pr.row_off_y[K] = 1;
      pr.Cf = F[cur ^ 1]; pr.ldcf = C; pr.bscf = bs;
      if ((rc = launch_gemm2(pr, s))) return rc;
    }
    cur ^= 1;
  }
  // block 0
  {
    const int d = tcn_dilation_of(h, 0);
    const int rows = rows_per_cta_for(L1, B);
    dim3 grid((unsigned)((L1 + rows - 1) / rows), B);
    tcn_first_bwd_kernel<<<grid, C, 0, s>>>(x, T, F[cur], bs, (int)L1, C, K, d, tcn_res_off(h, d), tcn_param(h, "process_blocks.0.conv1.weight"),
                                            tcn_param(h, "process_blocks.0.conv1.bias"), tcn_param(h, "process_blocks.0.relu.weight"),
                                            grad("process_blocks.0.conv1.weight"), grad("process_blocks.0.res.weight"),
                                            grad("process_blocks.0.conv1.bias"), grad("process_blocks.0.relu.weight"), rows);
    RFX_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int rfx_tcn_set_wgrad_impl(int impl) {
  RFX_REQUIRE(impl >= 0 && impl <= 2, "impl 0 (tcgen05 bf16x3 when C = 256, else mma.sync), 1 (fp32 SIMT cross-check) or 2 (mma.sync bf16x3)");
  g_tcn_wgrad_impl = impl;
  return 0;
}

int rfx_tcn_backward_launches_per_call(const rfx_tcn_t* h) { return h ? 2 + 5 * (h->cfg.nblocks - 1) : 0; }

}  // extern "C"
