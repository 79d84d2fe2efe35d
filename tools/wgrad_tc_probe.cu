// PROBE (not part of libremfx_b200.so, NOT YET RUN ON HARDWARE): the TCN weight-gradient contraction on tcgen05 with
// MN-major shared-memory operands, next to the fp32 SIMT form of the same sums.  DESIGN.md 4.7 names this as the next step
// for `tcn_wgrad_kernel` (mma.sync, 2.32 ms per block at 1 x 262144); this file is the stand-alone bring-up harness for it.
//
//   dW[tap][co][ci] = sum_t g[t][co] * x[t + off[tap]][ci]        g, x: time-major split-bf16 planes ([t][c], hi / lo)
//
// Contraction over TIME with both operands stored [t][c]: for the MMA that is "MN-major" (the M / N index is the
// contiguous one).  A TMA box of {64 channels, BK time steps} with SWIZZLE_128B lands in shared memory exactly as the
// canonical MN-major SW128 layout of cute::UMMA (mma_traits_sm100.hpp:  Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units): 128-byte rows = 64 channels of one time step, 8 time steps per 1024-byte swizzle atom (SBO = 1024 B to the
// next 8 time steps), the next block of 64 channels one box further (LBO = box bytes).  The instruction descriptor sets
// a_major = b_major = MN (bits 15 / 16).
//
// Roles are swapped relative to the maths so that the epilogue's atomics coalesce: M = ci (x tile, 2 halves of 128),
// N = co (g tile, 256), D[ci][co] in TMEM (2 x 256 columns = all 512); a thread (= TMEM lane = ci) adds its 32 consecutive co
// columns to dW[co][ci], i.e. every warp-wide atomic instruction covers 32 consecutive floats.
// One CTA = (tap, chunk of time); warp 0 TMA producer, warp 1 MMA issuer (bf16x3: lo*hi + hi*lo + hi*hi), warps 2-5 epilogue.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -o /tmp/wgrad_tc_probe tools/wgrad_tc_probe.cu -lcuda
//   /tmp/wgrad_tc_probe [L=249868] [dilation=1]
// Prints the max relative error against the SIMT sums (computed on a 4096-step slice so it finishes quickly) and the time
// per launch at the full length.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../remfx_b200/csrc/common.cuh"

namespace rfx {
void set_error(const std::string&) {}
const char* get_error() { return ""; }
}  // namespace rfx
using namespace rfx;

constexpr int C = 256;        // channels (co = ci = 256)
constexpr int KT = 7;         // conv taps (+ 1 residual tap)
constexpr int BK = 32;        // time steps per stage
constexpr int BOX = 64 * BK * 2 * 2;          // one TMA box: 64 channels x BK steps x (hi, lo) = 8 KB
constexpr int PLANE = 64 * BK * 2;            // lo plane offset inside a box (4 KB)
constexpr int STAGE = 8 * BOX;                // 4 x-boxes (ci) + 4 g-boxes (co) = 64 KB
constexpr int STAGES = 3;
constexpr int SMEM = STAGES * STAGE + 1024 + 256;

struct Params {
  int L;          // valid time steps of g (rows beyond are TMA zero fill)
  int tchunk;     // time steps per CTA (multiple of BK)
  int nchunks;
  int off[KT + 1];
  float* dW;      // [KT + 1][C][C]
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

// MN-major SWIZZLE_128B operand: start address, LBO = bytes between 64-channel blocks, SBO = 1024 (8 time steps)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX,
                                                          const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap = blockIdx.x, chunk = blockIdx.y;
  const int t_begin = chunk * p.tchunk;
  const int t_end = min(p.L, t_begin + p.tchunk);
  const int iters = (t_end - t_begin + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
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
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE);
        uint8_t* st = smem + s * STAGE;
        const int t0 = t_begin + it * BK;
        for (int cb = 0; cb < 4; ++cb) tma_load_3d(st + cb * BOX, &mapX, cb * 64, t0 + p.off[tap], 0, &full_bar[s]);       // x: ci blocks
        for (int cb = 0; cb < 4; ++cb) tma_load_3d(st + (4 + cb) * BOX, &mapG, cb * 64, t0, 0, &full_bar[s]);              // g: co blocks
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // M = 128 (ci half), N = 256 (co), both operands MN-major
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256) | (1u << 15) | (1u << 16);
      int s = 0; uint32_t ph = 0;
      for (int it = 0; it < iters; ++it) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t xb = smem_u32(smem + s * STAGE), gb = xb + 4 * BOX;
#pragma unroll
        for (int kk = 0; kk < BK / 16; ++kk) {
          const uint32_t ko = kk * 2048;  // 16 time steps = two 1024-byte atoms
          const uint64_t g_hi = umma_desc_mn_sw128(gb + ko, BOX), g_lo = umma_desc_mn_sw128(gb + PLANE + ko, BOX);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t x_hi = umma_desc_mn_sw128(xb + h * 2 * BOX + ko, BOX), x_lo = umma_desc_mn_sw128(xb + h * 2 * BOX + PLANE + ko, BOX);
            const uint32_t d = tmem + h * 256;
            const uint32_t first = (it == 0 && kk == 0) ? 0u : 1u;
            umma_f16(d, x_lo, g_hi, idesc, first);
            umma_f16(d, x_hi, g_lo, idesc, 1u);
            umma_f16(d, x_hi, g_hi, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else {
    // epilogue: warp w may only read TMEM lanes [32 (w % 4), +32)
    const int q = warp & 3;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* dst = p.dW + (size_t)tap * C * C;
    for (int h = 0; h < 2; ++h) {
      const int ci = h * 128 + q * 32 + lane;
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + h * 256 + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)(cc * 32 + j) * C + ci, __uint_as_float(v[j]));
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// fp32 form of the same sums on rows [0, Lcheck)
__global__ void wgrad_simt_kernel(const __nv_bfloat16* g, const __nv_bfloat16* x, size_t g_plane, size_t x_plane, int Lcheck, Params p, float* out) {
  const long long total = (long long)(KT + 1) * C * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C), co = (int)((i / C) % C), tap = (int)(i / ((long long)C * C));
    float acc = 0.0f;
    for (int t = 0; t < Lcheck; ++t) {
      const float gv = __bfloat162float(g[(size_t)t * C + co]) + __bfloat162float(g[g_plane + (size_t)t * C + co]);
      const size_t xr = (size_t)(t + p.off[tap]) * C + ci;
      acc = fmaf(gv, __bfloat162float(x[xr]) + __bfloat162float(x[x_plane + xr]), acc);
    }
    out[i] = acc;
  }
}

__global__ void fill_kernel(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float v = ((int)(h & 0xFFFF) - 32768) / 32768.0f;
    split_bf16(v, hi[i], lo[i]);
  }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static int make_map(CUtensorMap* m, void* base, size_t rows, size_t plane_elems) {
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)plane_elems * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)BK, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
  return 0;
}

static int run(int L, int d, int Lcheck, bool timing) {
  Params p{};
  p.L = L;
  for (int j = 0; j < KT; ++j) p.off[j] = j * d;
  p.off[KT] = (KT - 1) * d / 2;
  const int Lin = L + (KT - 1) * d;
  const size_t gp = (size_t)L * C, xp = (size_t)Lin * C;
  __nv_bfloat16 *g, *x;
  float *dW, *ref;
  CK(cudaMalloc(&g, 2 * gp * 2)); CK(cudaMalloc(&x, 2 * xp * 2));
  CK(cudaMalloc(&dW, (size_t)(KT + 1) * C * C * 4)); CK(cudaMalloc(&ref, (size_t)(KT + 1) * C * C * 4));
  fill_kernel<<<1024, 256>>>(g, g + gp, gp, 1u);
  fill_kernel<<<1024, 256>>>(x, x + xp, xp, 2u);
  CUtensorMap mg, mx;
  if (make_map(&mg, g, L, gp) || make_map(&mx, x, Lin, xp)) return 1;
  // about two waves of CTAs
  long long want = (148 * 2 + KT) / (KT + 1);
  long long tchunk = ((L + want - 1) / want + BK - 1) / BK * BK;
  p.tchunk = (int)tchunk;
  p.nchunks = (int)((L + tchunk - 1) / tchunk);
  p.dW = dW;
  CK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid(KT + 1, p.nchunks);
  CK(cudaMemset(dW, 0, (size_t)(KT + 1) * C * C * 4));
  wgrad_tc_kernel<<<grid, 192, SMEM>>>(mg, mx, p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (Lcheck > 0) {
    wgrad_simt_kernel<<<148 * 8, 256>>>(g, x, gp, xp, Lcheck, p, ref);
    CK(cudaDeviceSynchronize());
    std::vector<float> a((size_t)(KT + 1) * C * C), b(a.size());
    CK(cudaMemcpy(a.data(), dW, a.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), ref, b.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); ++i) { num += (double)(a[i] - b[i]) * (a[i] - b[i]); den += (double)b[i] * b[i]; }
    printf("L=%d d=%d grid=(%d,%d): rel-RMS vs SIMT = %.3e\n", L, d, grid.x, grid.y, sqrt(num / den));
  }
  if (timing) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) wgrad_tc_kernel<<<grid, 192, SMEM>>>(mg, mx, p);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * (KT + 1) * C * (double)C * L;
    printf("L=%d d=%d: %.3f ms per launch, %.1f TF/s fp32-equivalent (%.1f bf16-equivalent)\n", L, d, ms / reps, flop / (ms / reps * 1e-3) / 1e12,
           3 * flop / (ms / reps * 1e-3) / 1e12);
  }
  cudaFree(g); cudaFree(x); cudaFree(dW); cudaFree(ref);
  return 0;
}

int main(int argc, char** argv) {
  const int L = argc > 1 ? atoi(argv[1]) : 249868;
  const int d = argc > 2 ? atoi(argv[2]) : 1;
  if (cuInit(0) != CUDA_SUCCESS) { printf("cuInit failed\n"); return 1; }
  CK(cudaFree(0));
  if (run(4096 + 5, d, 4096 + 5, false)) return 1;   // correctness on a short, ragged length (whole sum checked)
  if (run(L, d, 0, true)) return 1;                   // timing at the benchmark length
  return 0;
}
