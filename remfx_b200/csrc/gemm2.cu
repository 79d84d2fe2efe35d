// gemm2: TMA-fed, warp-specialised, persistent tcgen05 GEMM / implicit-GEMM engine with fp32-grade accuracy.
//
//   C[b, y, x, n] = epilogue( sum_{tap} sum_{k} A[b, y + dy[tap], x + dx[tap], k] * W[n, tap * Ktap + k] )
//
// Plain dense layers use one "tap" on a 1-D row space (Y = 1); the TCN dilated Conv1d (remfx/tcn.py:28-36,
// 48-59) uses 7 taps at row offsets j * dilation plus an 8th tap for the 1x1 residual at the centre offset;
// 3x3 Conv2d layers (Cnn14, remfx/classifier.py:240-256) use 9 taps on a 2-D pixel space.  Activations are
// channel-last, so this is an im2col-free implicit GEMM: each tap is one TMA box of (xt x yt) pixels at a
// shifted origin, and zero padding comes from the TMA unit's out-of-bounds fill (signed coordinates).
//
// Numerics ("bf16x3"): every fp32 value v lives in HBM as TWO bf16 planes (hi = bf16(v), lo = bf16(v - hi)),
// 4 bytes per element like fp32; per K-step three MMAs (lo*hi, hi*lo, hi*hi) accumulate into one fp32
// TMEM accumulator.  The dropped lo*lo term is O(2^-16): results meet the reference's 1e-4 rel-RMS gate
// where a single bf16 or TF32 pass does not.
//
// Structure (one CTA per SM, persistent over output tiles, 576 threads):
//   warp 0    TMA producer: one 4-D tensor-map load per operand per stage (hi+lo planes in one box),
//             SWIZZLE_128B, mbarrier complete_tx; K / M / N tails are zero-filled by the TMA unit
//   warp 1    MMA issuer: a single thread issues tcgen05.mma (M=128, N=BN, K=16) from smem descriptors,
//             tcgen05.commit releases smem stages and publishes finished accumulators
//   warps 2-17 epilogue (four warps per TMEM lane quarter, each draining a quarter of the columns): tcgen05.ld the
//             accumulator (TMEM is double-buffered: 2 x BN columns, so the epilogue of tile i overlaps the
//             main loop of tile i+1), fused per-column affine(s) + activation, then either fp32 rows or
//             re-split bf16 hi/lo planes for the next layer
// DUAL mode (TCN block): the last tap accumulates into a second TMEM region and is added AFTER the
// activation:  out = PReLU(conv + bias) + W_res x  (remfx/tcn.py:50-57); BN = 256, no TMEM double buffering.
#include "kernels.h"

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace rfx {

constexpr int G2_BM = 128;
constexpr int G2_BK = 64;
constexpr int G2_A_STAGE = 2 * G2_BM * G2_BK * 2;  // hi + lo planes of the A tile (32 KB)

// ------------------------------------------------------------------------------------------------
// Tensor maps (driver entry point fetched at run time: no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 5-D bf16 map over split planes: dims {cols, X, Y, batch, 2 planes}; box {64, xt, yt, 1, 2}.
static int make_split_map(CUtensorMap* map, const void* base, long long cols, long long X, long long Y, long long batch, long long ldx,
                          long long ldy, long long batch_stride, long long plane_stride, int xt, int yt) {
  EncodeTiledFn fn = encode_fn();
  RFX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable");
  RFX_REQUIRE(((uintptr_t)base & 15) == 0 && (ldx % 8) == 0 && (ldy % 8) == 0 && (batch_stride % 8) == 0 && (plane_stride % 8) == 0,
              "split-bf16 operand must be 16-byte aligned with strides that are multiples of 8 elements");
  RFX_REQUIRE(xt >= 1 && yt >= 1 && xt <= 256 && yt <= 256, "tile extents");
  cuuint64_t dims[5] = {(cuuint64_t)cols, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)batch, 2};
  // a stride of 0 is not accepted by the encoder: degenerate dimensions get any valid stride
  const long long ldy_e = Y > 1 ? ldy : ldx * X, bs_e = batch > 1 ? batch_stride : (ldy_e * (Y > 1 ? Y : 1));
  cuuint64_t strides[4] = {(cuuint64_t)ldx * 2, (cuuint64_t)(ldy_e > 0 ? ldy_e : 8) * 2, (cuuint64_t)(bs_e > 0 ? bs_e : 8) * 2,
                           (cuuint64_t)plane_stride * 2};
  cuuint32_t box[5] = {(cuuint32_t)G2_BK, (cuuint32_t)xt, (cuuint32_t)yt, 1, 2};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 1;
  }
  return 0;
}

}  // namespace rfx
// Generic bf16 tiled tensor-map encoder for the other translation units (tcn_bwd.cu): rank <= 5, strides in BYTES for dims 1..rank-1.
int rfx_encode_tiled_bf16(void* map, const void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                          const unsigned* box, int swizzle128) {
  using namespace rfx;
  EncodeTiledFn fn = encode_fn();
  RFX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is unavailable");
  RFX_REQUIRE(rank >= 1 && rank <= 5, "tensor map rank");
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 1;
  }
  return 0;
}
namespace rfx {

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// ------------------------------------------------------------------------------------------------
// fp32 -> split bf16 planes (weights at load time; also a stand-alone op for tests)
// ------------------------------------------------------------------------------------------------
__global__ void split_rows_kernel(const float* __restrict__ src, long long ld_src, int rows, int cols, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long ld_dst, int rows_pad, int cols_pad) {
  const long long total = (long long)rows_pad * cols_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_pad), c = (int)(i % cols_pad);
    const float v = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.0f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[(size_t)r * ld_dst + c] = h;
    lo[(size_t)r * ld_dst + c] = l;
  }
}

int launch_split_rows(const float* src, long long ld_src, int rows, int cols, __nv_bfloat16* hi, __nv_bfloat16* lo, long long ld_dst,
                      int rows_pad, int cols_pad, cudaStream_t stream) {
  const long long total = (long long)rows_pad * cols_pad;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  split_rows_kernel<<<blocks, 256, 0, stream>>>(src, ld_src, rows, cols, hi, lo, ld_dst, rows_pad, cols_pad);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
struct G2Params {
  int X, Y;        // valid output extent per batch item (Y = 1 for plain row spaces)
  int xt, yt;      // pixel tile (xt * yt = 128)
  int tiles_x;     // m_tiles = tiles_x * tiles_y
  int N;           // valid output columns
  int batch;       // batch items (grid tiles = batch * m_tiles * n_tiles)
  int m_tiles, n_tiles;
  int chunk;       // > 0: CTA c owns the tiles [c * chunk, (c + 1) * chunk) and the tiles of an item run y-fastest (launches with fused
                   // GroupNorm statistics: consecutive tiles share their statistics segment); 0: tiles strided over the grid, x-fastest
  // shared-memory plan (host-chosen): [resident W: w_res_bytes][stages x stage_bytes][barriers]
  int stages;      // ring depth (2..G2_MAX_STAGES)
  int stage_bytes; // A tile (+ W tile when W is not resident); multiple of 1024
  int b_bytes;     // one W k-block: 2 planes x n_box rows x 64 columns bf16
  int n_box;       // W rows per TMA box (= BN, or the 16-rounded N when a single n-tile covers the output)
  int w_res_bytes; // > 0: all KB k-blocks of W stay in shared memory for the whole kernel (loaded once); needs n_tiles == 1
  int fast;        // 1 = bf16-fast: issue the hi*hi pass only (set_matmul_precision)
  int dbg_skip;    // timing experiments only (RFX_G2_DEBUG_SKIP): 1 = load W only for the first k-block of a tile, 2 = same for A
  int ktap;        // valid K per tap (the last 64-wide block of a tap may be partial: its dead 16-wide steps are skipped)
  int kb_per_tap;  // 64-wide K blocks per tap
  int taps;
  int dx[16], dy[16];  // A origin offset of each tap
  // Tap groups (ngroups > 0; pixel tiles of one row, yt == 1): the taps of a group differ by an x offset only, so ONE A box of
  // a_rows = 128 + 8 consecutive pixels serves them all -- each tap's MMAs read it from a start address shifted by whole 128-byte
  // rows.  A 3x3 conv loads 3 boxes per k-block instead of 9.
  int ngroups;
  int g_start[17];     // group g = taps g_taps[g_start[g] .. g_start[g + 1])
  int g_taps[16];
  int g_dx0[16], g_dy[16];   // A origin offset of the group's box
  int a_plane;         // bytes of one plane of the group's A box (a_rows * 128); the lo plane follows the hi plane
  long long ldcy_f, ldcy_s;  // output y strides (elements) for the fp32 / split outputs
  // outputs: fp32 (Cf) and/or split planes (Chi/Clo); row stride ld*, batch stride bs* (elements)
  float* Cf;
  long long ldcf, bscf;
  __nv_bfloat16* Chi;
  __nv_bfloat16* Clo;
  long long ldcs, bscs;
  // epilogue: v = acc * s1[n] + t1[n]; v = act(v) with optional per-column PReLU slope; v = v * s2[n] + t2[n] (if post) ...
  const float* s1;
  const float* t1;
  const float* s2;
  const float* t2;
  const float* slope;  // ACT_PRELU: per-column negative slope
  int act;
  int cf_pre;    // 1: the fp32 output receives the PRE-activation (after scale / shift), the split output the activated value
  int cst;       // 1: fp32 rows are staged through shared memory (GN instantiations; see G2Row)
  int cf_accum;  // 1: the fp32 output accumulates (Cf += v; ACT_NONE without GLU only) -- the input-gradient GEMMs of residual branches
  // fused GroupNorm statistics of the (post-bias) output: accum[(seg * G + g) * 2 + {0,1}] += (sum, sum of squares)
  double* gn_acc;
  int gn_G, gn_per_x, gn_cpg, gn_cmod;  // group of column n = ((n % cmod) / cpg); seg = per_x ? b * X + px : b
};

// Apply the epilogue to 8 consecutive columns [n, n+8) held in v[0..7].  Null vectors act as 1 / 0, which is
// exact in fp32 (v * 1 + 0 == v), so the result equals the conditional form while the inner loop stays branch-free.
__device__ __forceinline__ void load8(const float* vec, int n, float fill, bool aligned, float (&o)[8]) {
  if (!vec) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fill;
  } else if (aligned) {
    const float4 a = *reinterpret_cast<const float4*>(vec + n);
    const float4 b = *reinterpret_cast<const float4*>(vec + n + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = vec[n + i];
  }
}

template <int ACT>
__device__ __forceinline__ float g2_act(float v, float slope) {
  if (ACT == ACT_TANH) return tanhf(v);
  if (ACT == ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == ACT_SIGMOID) return sigmoidf_acc(v);
  if (ACT == ACT_PRELU) return v >= 0.0f ? v : v * slope;
  if (ACT == ACT_GELU) return gelu_fast(v);  // erf form (torch F.gelu default)
  return v;  // ACT_NONE; ACT_GLU_PAIR is resolved by the caller (needs pairs of columns)
}

struct G2Row {  // where this thread's output row lives
  float* cf;
  __nv_bfloat16* chi;
  __nv_bfloat16* clo;
  bool cf_vec, cs_vec, vec_al;
  // GN instantiations only: fp32 rows leave through a per-warp staging block in shared memory so that every store instruction
  // writes whole 128-byte lines (a thread owns a ROW: written directly, one instruction scatters 16-byte pieces over 32 lines and
  // the L1 / L2 request rate -- not HBM -- bounds the thin layers).  cst = the warp's [32][33] floats, rowp = its 32 row pointers
  // (null for rows outside the output), lane = this thread's row.  Null cst = direct stores.
  float* cst;
  float* const* rowp;
  int lane;
  bool ok;
};
constexpr int G2_CST_WARP_BYTES = 32 * 33 * 4 + 32 * 8;  // staging block + row pointers of one epilogue warp

// One full 32-column chunk [nb, nb + 32) of this thread's row, activation known at compile time: straight-line code, so the
// compiler hoists every scale / bias load to the top and the instruction cache only holds the variant in use.
// gs / gss accumulate the fused GroupNorm statistics when GN is set.
template <int ACT, bool DUAL, bool GN>
__device__ __forceinline__ void g2_chunk(const uint32_t (&v)[32], const uint32_t (&v2)[32], int nb, const G2Params& p, const G2Row& r, float& gs,
                                         float& gss) {
  float o[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
  if (p.s1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sc[8];
      load8(p.s1, nb + 8 * j, 1.0f, r.vec_al, sc);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[8 * j + i] *= sc[i];
    }
  }
  if (p.t1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sh[8];
      load8(p.t1, nb + 8 * j, 0.0f, r.vec_al, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[8 * j + i] += sh[i];
    }
  }
  if (p.s2 || p.t2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sc[8], sh[8];
      load8(p.s2, nb + 8 * j, 1.0f, r.vec_al, sc);
      load8(p.t2, nb + 8 * j, 0.0f, r.vec_al, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[8 * j + i] = fmaf(o[8 * j + i], sc[i], sh[i]);
    }
  }
  if (ACT != ACT_NONE && !DUAL && p.cf_pre && r.cf) {  // training forward: keep the pre-activation beside the activated split output
    if (r.cf_vec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(r.cf + nb + 4 * i) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) r.cf[nb + i] = o[i];
    }
  }
  if (ACT == ACT_PRELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float sl[8];
      load8(p.slope, nb + 8 * j, 0.0f, r.vec_al, sl);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[8 * j + i] = g2_act<ACT_PRELU>(o[8 * j + i], sl[i]);
    }
  } else if (ACT != ACT_NONE && ACT != ACT_GLU_PAIR) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = g2_act<ACT>(o[i], 0.0f);
  }
  if (DUAL) {
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] += __uint_as_float(v2[i]);
  }
  if (GN && r.ok) {
#pragma unroll
    for (int i = 0; i < 32; ++i) { gs += o[i]; gss = fmaf(o[i], o[i], gss); }
  }
  if (ACT == ACT_GLU_PAIR) {  // (value, gate) column pairs -> 16 output columns starting at nb / 2
    float g[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] = o[2 * i] * sigmoidf_fast(o[2 * i + 1]);
    const int c0 = nb >> 1;
    if (r.cf && !p.cf_pre) {
      if (r.cf_vec) {
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(r.cf + c0 + 4 * i) = make_float4(g[4 * i], g[4 * i + 1], g[4 * i + 2], g[4 * i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) r.cf[c0 + i] = g[i];
      }
    }
    if (r.chi) {
      uint32_t ph[8], pl[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
        const float2 hf = __bfloat1622float2(h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(g[2 * i] - hf.x, g[2 * i + 1] - hf.y);
        ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
        pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      if (r.cs_vec) {  // 16 columns = 32 bytes per plane; c0 is a multiple of 16
        *reinterpret_cast<uint4*>(r.chi + c0) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(r.chi + c0 + 8) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
        *reinterpret_cast<uint4*>(r.clo + c0) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        *reinterpret_cast<uint4*>(r.clo + c0 + 8) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<uint32_t*>(r.chi + c0 + 2 * i) = ph[i];
          *reinterpret_cast<uint32_t*>(r.clo + c0 + 2 * i) = pl[i];
        }
      }
    }
    return;
  }
  if (GN && r.cst) {  // warp-collective: every lane stages its row (16-byte pieces XOR-swizzled by the row: conflict-free both ways)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(r.cst + r.lane * 32 + ((j ^ (r.lane & 7)) << 2)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    __syncwarp();
    const int pc = r.lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {  // one instruction = 4 rows x 128 contiguous bytes
      const int row = it * 4 + (r.lane >> 3);
      float* rp = r.rowp[row];
      if (rp) {
        float4 q = *reinterpret_cast<const float4*>(r.cst + row * 32 + ((pc ^ (row & 7)) << 2));
        float4* dst = reinterpret_cast<float4*>(rp + nb + 4 * pc);
        if (p.cf_accum) { const float4 old = *dst; q.x += old.x; q.y += old.y; q.z += old.z; q.w += old.w; }
        *dst = q;
      }
    }
    __syncwarp();
  } else if (r.cf && !(ACT != ACT_NONE && p.cf_pre)) {
    if (ACT == ACT_NONE && p.cf_accum) {
      if (r.cf_vec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 q = *reinterpret_cast<const float4*>(r.cf + nb + 4 * i);
          o[4 * i] += q.x; o[4 * i + 1] += q.y; o[4 * i + 2] += q.z; o[4 * i + 3] += q.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] += r.cf[nb + i];
      }
    }
    if (r.cf_vec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(r.cf + nb + 4 * i) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) r.cf[nb + i] = o[i];
    }
  }
  if (r.chi) {
    uint32_t ph[16], pl[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
      const float2 hf = __bfloat1622float2(h2);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(o[2 * i] - hf.x, o[2 * i + 1] - hf.y);
      ph[i] = *reinterpret_cast<const uint32_t*>(&h2);
      pl[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    if (r.cs_vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<uint4*>(r.chi + nb + 8 * i) = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
        *reinterpret_cast<uint4*>(r.clo + nb + 8 * i) = make_uint4(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        r.chi[nb + i] = __ushort_as_bfloat16((unsigned short)((ph[i >> 1] >> (16 * (i & 1))) & 0xffffu));
        r.clo[nb + i] = __ushort_as_bfloat16((unsigned short)((pl[i >> 1] >> (16 * (i & 1))) & 0xffffu));
      }
    }
  }
}

// Ragged chunks (N % 32 != 0, or an activation without a specialised path): groups of 8 columns with per-column bounds
// predicates and run-time switches.  Warp-collective (tcgen05.ld): every lane must call it; only lanes with row_ok store.
template <bool DUAL, bool GN>
__device__ __forceinline__ void g2_chunk_ragged(uint32_t taddr, uint32_t taddr2, int nb, const G2Params& p, const G2Row& r, bool row_ok, float& gs,
                                             float& gss) {
  const int ncol = min(32, p.N - nb);
  const bool staged = GN && r.cst && p.act != ACT_GLU_PAIR;  // warp-uniform
#pragma unroll 1
  for (int g8 = 0; g8 < ncol; g8 += 8) {
    uint32_t a1[8], a2[8];
    tmem_ld8(taddr + g8, a1);
    if (DUAL) tmem_ld8(taddr2 + g8, a2);
    tmem_ld_wait();
    if (!row_ok && !staged) continue;
    const int n8 = nb + g8;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = min(n8 + i, p.N - 1);  // clamped: out-of-range columns are computed but never stored
      float val = fmaf(__uint_as_float(a1[i]), p.s1 ? p.s1[n] : 1.0f, p.t1 ? p.t1[n] : 0.0f);
      if (p.s2 || p.t2) val = fmaf(val, p.s2 ? p.s2[n] : 1.0f, p.t2 ? p.t2[n] : 0.0f);
      o[i] = val;
    }
    const bool pre = p.cf_pre && p.act != ACT_NONE;
    if (pre && r.cf) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n8 + i < p.N) r.cf[n8 + i] = o[i];
    }
    switch (p.act) {
      case ACT_TANH:
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = tanhf(o[i]);
        break;
      case ACT_RELU:
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.0f);
        break;
      case ACT_SIGMOID:
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sigmoidf_acc(o[i]);
        break;
      case ACT_PRELU:
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = o[i] >= 0.0f ? o[i] : o[i] * p.slope[min(n8 + i, p.N - 1)];
        break;
      case ACT_GELU:
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = gelu_fast(o[i]);
        break;
      default: break;
    }
    if (DUAL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] += __uint_as_float(a2[i]);
    }
    if (p.gn_acc && row_ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (n8 + i < p.N) { gs += o[i]; gss = fmaf(o[i], o[i], gss); }
    }
    if (staged) {  // [32 rows][33]: scalar writes by row and the row-major read-out below are both conflict-free
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (g8 + i < ncol) r.cst[r.lane * 33 + g8 + i] = o[i];
      continue;
    }
    if (p.act == ACT_GLU_PAIR) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (n8 + 2 * i + 1 < p.N) {
          const float gv = o[2 * i] * sigmoidf_fast(o[2 * i + 1]);
          const int c = (n8 >> 1) + i;
          if (r.cf && !pre) r.cf[c] = gv;
          if (r.chi) {
            __nv_bfloat16 h, l;
            split_bf16(gv, h, l);
            r.chi[c] = h;
            r.clo[c] = l;
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (n8 + i < p.N) {
        if (r.cf && !pre) r.cf[n8 + i] = p.cf_accum ? r.cf[n8 + i] + o[i] : o[i];
        if (r.chi) {
          __nv_bfloat16 h, l;
          split_bf16(o[i], h, l);
          r.chi[n8 + i] = h;
          r.clo[n8 + i] = l;
        }
      }
    }
  }
  if (staged) {  // the warp's 32 x ncol block in row-major order: consecutive lanes write consecutive addresses
    __syncwarp();
    const int total = 32 * ncol;
    for (int idx = r.lane; idx < total; idx += 32) {
      const int row = idx / ncol, col = idx - row * ncol;
      float* rp = r.rowp[row];
      if (rp) {
        const float v = r.cst[row * 33 + col];
        rp[nb + col] = p.cf_accum ? rp[nb + col] + v : v;
      }
    }
    __syncwarp();
  }
}

// warp 0 TMA, warp 1 MMA, then the epilogue warps: 16 (8 in DUAL mode, whose second accumulator costs 32 more registers)
constexpr int g2_epi_warps(bool dual) { return dual ? 8 : 16; }
constexpr int g2_threads(bool dual) { return 64 + 32 * g2_epi_warps(dual); }
// TWO = the "two CTAs per SM" plan of thin layers (BN = 128, small K): 8 epilogue warps, at most half the registers and shared
// memory, 2 x 256 TMEM columns.  Those layers are bound by the latency of their per-tile chain (load -> MMA -> drain), not by
// any throughput; a second resident CTA doubles the tiles in flight.
constexpr int g2_epi_warps2(bool dual, bool two) { return (dual || two) ? 8 : 16; }
constexpr int g2_threads2(bool dual, bool two) { return 64 + 32 * g2_epi_warps2(dual, two); }

constexpr int G2_MAX_STAGES = 6;

template <int BN, bool DUAL, bool GN, bool TWO>   // GN: fused GroupNorm statistics (chunked tile schedule, running fp64 sums) -- its own instantiation so
                                        // that the plain kernels carry none of that state in their register-tight epilogue
__global__ void __launch_bounds__(g2_threads2(DUAL, TWO), TWO ? 2 : 1)
    gemm2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const G2Params p) {
  const int STAGES = p.stages;
  const uint32_t STAGE_BYTES = (uint32_t)p.stage_bytes;
  const bool w_res = p.w_res_bytes > 0;
  constexpr uint32_t IDESC = umma_idesc_bf16(G2_BM, BN);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* wres = smem;                 // resident W k-blocks (w_res)
  uint8_t* ring = smem + p.w_res_bytes;  // operand stages
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G2_MAX_STAGES;
  uint64_t* wfull_bar = empty_bar + G2_MAX_STAGES;  // resident W has landed
  uint64_t* tfull_bar = wfull_bar + 1;              // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.kb_per_tap * p.taps;
  const int tiles_per_batch = p.m_tiles * p.n_tiles;
  const int total_tiles = p.batch * tiles_per_batch;
  // tile schedule of this CTA (identical in the three roles): my_tiles tiles, the i-th one being tile_at(i)
  const int my_tiles = GN ? max(0, min(p.chunk, total_tiles - (int)blockIdx.x * p.chunk))
                          : ((int)blockIdx.x < total_tiles ? (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
  auto tile_at = [&](int i) { return GN ? (int)blockIdx.x * p.chunk + i : (int)blockIdx.x + i * (int)gridDim.x; };
  const int tiles_y = p.m_tiles / p.tiles_x;
  auto tile_x = [&](int mt) { return GN ? mt / tiles_y : mt % p.tiles_x; };
  auto tile_y = [&](int mt) { return GN ? mt % tiles_y : mt / p.tiles_x; };

  // Epilogue warps whose column slice lies beyond N in every tile (single n-tile launches with N < BN) have nothing to drain: they
  // stay out of the tile loop instead of spinning on the accumulator barriers beside the warps that work.
  constexpr int EPI_CH = BN / (8 * g2_epi_warps2(DUAL, TWO));                      // 32-column chunks per epilogue warp
  const int epi_slices = (DUAL || p.n_tiles > 1) ? g2_epi_warps2(DUAL, TWO) / 4    // column slices (of EPI_CH chunks) with work
                                                 : min(g2_epi_warps2(DUAL, TWO) / 4, (p.N + EPI_CH * 32 - 1) / (EPI_CH * 32));
  const int epi_active = 4 * epi_slices;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(wfull_bar, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], epi_active);  // one arrive per epilogue warp that has columns to drain
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      prefetch_tmap(&mapA);
      prefetch_tmap(&mapW);
      if (w_res) {  // the whole (small) weight matrix: once per CTA
        mbar_arrive_expect_tx(wfull_bar, (uint32_t)(KB * p.b_bytes));
        for (int kb = 0; kb < KB; ++kb) tma_load_5d(wres + (size_t)kb * p.b_bytes, &mapW, kb * G2_BK, 0, 0, 0, 0, wfull_bar);
      }
      int s = 0;
      uint32_t ph = 0;  // ring position and its phase bit
      for (int ti = 0; ti < my_tiles; ++ti) {
        const int tile = tile_at(ti);
        const int b = tile / tiles_per_batch;
        const int r = tile % tiles_per_batch;
        const int mt = r / p.n_tiles;
        const int x0 = tile_x(mt) * p.xt, y0 = tile_y(mt) * p.yt;
        const int n0 = (r % p.n_tiles) * BN;
        if (p.ngroups > 0) {  // one stage = (tap group, k-block): the group's A box + (W not resident) one W box per tap of the group
          for (int g = 0; g < p.ngroups; ++g) {
            const int nt = p.g_start[g + 1] - p.g_start[g];
            for (int kbt = 0; kbt < p.kb_per_tap; ++kbt) {
              mbar_wait(&empty_bar[s], ph ^ 1);
              mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * p.a_plane + (w_res ? 0 : nt * p.b_bytes)));
              uint8_t* st = ring + s * STAGE_BYTES;
              tma_load_5d(st, &mapA, kbt * G2_BK, x0 + p.g_dx0[g], y0 + p.g_dy[g], b, 0, &full_bar[s]);
              if (!w_res)
                for (int j = 0; j < nt; ++j) {
                  const int tap = p.g_taps[p.g_start[g] + j];
                  tma_load_5d(st + 2 * p.a_plane + j * p.b_bytes, &mapW, (tap * p.kb_per_tap + kbt) * G2_BK, n0, 0, 0, 0, &full_bar[s]);
                }
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          }
          continue;
        }
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          const bool ldA = !(p.dbg_skip == 2 && kb > 0), ldW = !w_res && !(p.dbg_skip == 1 && kb > 0);
          mbar_arrive_expect_tx(&full_bar[s], (ldA ? G2_A_STAGE : 0) + (ldW ? p.b_bytes : 0));
          const int tap = kb / p.kb_per_tap;
          const int kc = (kb % p.kb_per_tap) * G2_BK;
          uint8_t* st = ring + s * STAGE_BYTES;
          if (ldA) tma_load_5d(st, &mapA, kc, x0 + p.dx[tap], y0 + p.dy[tap], b, 0, &full_bar[s]);
          if (ldW) tma_load_5d(st + G2_A_STAGE, &mapW, kb * G2_BK, n0, 0, 0, 0, &full_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0, local = 0;
      uint32_t ph = 0;
      if (w_res) {
        mbar_wait(wfull_bar, 0);
        tc_fence_after();
      }
      for (; local < my_tiles; ++local) {
        const int tile = tile_at(local);
        const int as = DUAL ? 0 : (local & 1);
        const int aphase = DUAL ? (local & 1) : ((local >> 1) & 1);
        mbar_wait(&tempty_bar[as], aphase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const int kb_res = DUAL ? (p.taps - 1) * p.kb_per_tap : KB;  // first k-block of the residual tap
        // narrow outputs: shrink the MMA N to the valid columns of this tile (multiple of 16) -- less smem operand traffic
        const int n0t = ((tile % tiles_per_batch) % p.n_tiles) * BN;
        const int n_eff = min(BN, ((p.N - n0t) + 15) & ~15);
        const uint32_t idesc = umma_idesc_bf16(G2_BM, n_eff);
        if (p.ngroups > 0) {
          const uint32_t d_tmem = tmem_base + as * BN;
          uint32_t acc = 0u;  // the tile's first MMA overwrites the accumulator
          for (int g = 0; g < p.ngroups; ++g) {
            const int nt = p.g_start[g + 1] - p.g_start[g];
            for (int kbt = 0; kbt < p.kb_per_tap; ++kbt) {
              mbar_wait(&full_bar[s], ph);
              tc_fence_after();
              const uint32_t a_hi = smem_u32(ring + s * STAGE_BYTES);
              const uint32_t a_lo = a_hi + (uint32_t)p.a_plane;
              const int kvalid = p.ktap - kbt * G2_BK;
              const int nks = kvalid >= G2_BK ? G2_BK / 16 : (kvalid + 15) >> 4;
              for (int j = 0; j < nt; ++j) {
                const int tap = p.g_taps[p.g_start[g] + j];
                const uint32_t shift = (uint32_t)(p.dx[tap] - p.g_dx0[g]) * 128u;   // whole rows inside the group's box
                const uint32_t b_hi = w_res ? smem_u32(wres) + (uint32_t)((tap * p.kb_per_tap + kbt) * p.b_bytes)
                                            : a_hi + 2u * (uint32_t)p.a_plane + (uint32_t)(j * p.b_bytes);
                const uint32_t b_lo = b_hi + (uint32_t)p.n_box * (G2_BK * 2);
#pragma unroll
                for (int ks = 0; ks < G2_BK / 16; ++ks) {
                  if (ks >= nks) break;
                  const uint32_t ko = ks * 32;
                  // start address shifted by whole rows, descriptor otherwise unchanged: the swizzle is a function of the absolute
                  // shared-memory address on both the TMA and the MMA side (measured: setting the descriptor's base-offset field to the
                  // row phase gives wrong products, leaving it 0 is bit-compatible with per-tap boxes)
                  const uint64_t dah = umma_desc_sw128(a_hi + shift + ko), dal = umma_desc_sw128(a_lo + shift + ko);
                  const uint64_t dbh = umma_desc_sw128(b_hi + ko), dbl = umma_desc_sw128(b_lo + ko);
                  if (p.fast) {
                    umma_f16(d_tmem, dah, dbh, idesc, acc);
                  } else {
                    umma_f16(d_tmem, dal, dbh, idesc, acc);
                    umma_f16(d_tmem, dah, dbl, idesc, 1u);
                    umma_f16(d_tmem, dah, dbh, idesc, 1u);
                  }
                  acc = 1u;
                }
              }
              umma_commit(&empty_bar[s]);
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          }
          umma_commit(&tfull_bar[as]);
          continue;
        }
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t d_tmem = tmem_base + (DUAL ? (kb >= kb_res ? BN : 0) : as * BN);
          const bool first_kb = (kb == 0) || (kb == kb_res);
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(ring + s * STAGE_BYTES);
          const uint32_t a_lo = a_hi + G2_BM * G2_BK * 2;
          const uint32_t b_hi = w_res ? smem_u32(wres) + (uint32_t)(kb * p.b_bytes) : a_hi + G2_A_STAGE;
          const uint32_t b_lo = b_hi + (uint32_t)p.n_box * (G2_BK * 2);
          const int kvalid = p.ktap - (kb % p.kb_per_tap) * G2_BK;  // valid K columns in this block (rest is TMA zero fill)
          const int nks = kvalid >= G2_BK ? G2_BK / 16 : (kvalid + 15) >> 4;
#pragma unroll
          for (int ks = 0; ks < G2_BK / 16; ++ks) {
            if (ks >= nks) break;
            const uint32_t ko = ks * 32;
            const uint64_t dah = umma_desc_sw128(a_hi + ko), dal = umma_desc_sw128(a_lo + ko);
            const uint64_t dbh = umma_desc_sw128(b_hi + ko), dbl = umma_desc_sw128(b_lo + ko);
            if (p.fast) {
              umma_f16(d_tmem, dah, dbh, idesc, (!first_kb || ks > 0) ? 1u : 0u);
            } else {
              umma_f16(d_tmem, dal, dbh, idesc, (!first_kb || ks > 0) ? 1u : 0u);
              umma_f16(d_tmem, dah, dbl, idesc, 1u);
              umma_f16(d_tmem, dah, dbh, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which slice of the tile's columns this warp drains
    constexpr int CH = BN / (8 * g2_epi_warps2(DUAL, TWO));  // 32-column chunks per warp
    const bool vec_al = (((uintptr_t)p.s1 | (uintptr_t)p.t1 | (uintptr_t)p.s2 | (uintptr_t)p.t2 | (uintptr_t)p.slope) & 15) == 0;
    // Fused GroupNorm statistics: a thread's (sum, sum of squares) run in fp64 across the CTA's tiles for as long as the
    // (segment, group) they belong to stays the same, and reach the global accumulators only when it changes (chunked schedule:
    // once or twice per CTA instead of once per warp and tile -- the atomics on one item's two addresses were what bound the thin
    // GroupNorm layers).  per_x segments: per thread; per-item segments: warp-reduced first.
    double gsd = 0.0, gssd = 0.0;
    int gkey = -1;   // seg * G + group of the running sums (host: segments * G < 2^31)
    auto gn_flush = [&]() {
      if (gkey >= 0) {
        double* dst = p.gn_acc + (size_t)gkey * 2;
        if (p.gn_per_x) {
          if (gsd != 0.0 || gssd != 0.0) { atomicAdd(dst, gsd); atomicAdd(dst + 1, gssd); }
        } else {
          double ws = gsd, wss = gssd;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            ws += __shfl_xor_sync(0xffffffffu, ws, o);
            wss += __shfl_xor_sync(0xffffffffu, wss, o);
          }
          if (lane == 0) { atomicAdd(dst, ws); atomicAdd(dst + 1, wss); }
        }
      }
      gsd = 0.0; gssd = 0.0;
    };
    for (int local = 0; local < (half < epi_slices ? my_tiles : 0); ++local) {
      const int tile = tile_at(local);
      const int b = tile / tiles_per_batch;
      const int r = tile % tiles_per_batch;
      const int mt = r / p.n_tiles;
      const int n0 = (r % p.n_tiles) * BN;
      const int as = DUAL ? 0 : (local & 1);
      const int aphase = DUAL ? (local & 1) : ((local >> 1) & 1);
      mbar_wait_backoff(&tfull_bar[as], aphase);   // (the drain is not latency-critical: a sleeping warp leaves the issue slots to the others)
      tc_fence_after();
      const int rt = q * 32 + lane;  // row inside the tile = TMEM lane
      const int px = tile_x(mt) * p.xt + rt % p.xt;
      const int py = tile_y(mt) * p.yt + rt / p.xt;
      const uint32_t trow = tmem_base + (DUAL ? 0 : as * BN) + ((uint32_t)(q * 32) << 16);
      const bool row_ok = px < p.X && py < p.Y;
      float* cf = p.Cf ? p.Cf + (size_t)b * p.bscf + (size_t)py * p.ldcy_f + (size_t)px * p.ldcf : nullptr;
      __nv_bfloat16* chi = p.Chi ? p.Chi + (size_t)b * p.bscs + (size_t)py * p.ldcy_s + (size_t)px * p.ldcs : nullptr;
      __nv_bfloat16* clo = p.Clo ? p.Clo + (size_t)b * p.bscs + (size_t)py * p.ldcy_s + (size_t)px * p.ldcs : nullptr;
      G2Row row;
      row.cf = cf; row.chi = chi; row.clo = clo; row.vec_al = vec_al;
      row.lane = lane; row.ok = row_ok; row.cst = nullptr; row.rowp = nullptr;
      if (GN && p.cst) {
        uint8_t* blk = ring + STAGES * STAGE_BYTES + 256 + (size_t)(warp - 2) * G2_CST_WARP_BYTES;
        row.cst = reinterpret_cast<float*>(blk);
        float** rp = reinterpret_cast<float**>(blk + 32 * 33 * 4);
        __syncwarp();   // the previous tile's read-out is complete
        rp[lane] = row_ok ? cf : nullptr;
        __syncwarp();
        row.rowp = rp;
      }
      row.cf_vec = cf && ((p.ldcf & 3) == 0) && ((p.bscf & 3) == 0) && ((p.ldcy_f & 3) == 0) && (((uintptr_t)p.Cf & 15) == 0);
      row.cs_vec = chi && ((p.ldcs & 7) == 0) && ((p.bscs & 7) == 0) && ((p.ldcy_s & 7) == 0) && (((uintptr_t)p.Chi & 15) == 0) &&
                   (((uintptr_t)p.Clo & 15) == 0);
      const int gseg = p.gn_per_x ? b * p.X + min(px, p.X - 1) : b;
#pragma unroll 1
      for (int cc = 0; cc < CH; ++cc) {
        const int c = half * CH + cc;
        const int nb = n0 + c * 32;
        float gs = 0.0f, gss = 0.0f;  // this chunk's sums of the thread's row
        if (GN && nb < p.N) {  // the group id is chunk-uniform, the segment tile-uniform (per_x: per thread, changing for all lanes at once)
          const int knew = gseg * p.gn_G + (nb % p.gn_cmod) / p.gn_cpg;
          if (knew != gkey) { gn_flush(); gkey = knew; }
        }
        if (nb >= p.N) continue;  // warp-uniform: nothing to drain
        const bool generic = (nb + 32 > p.N) || (DUAL && p.act != ACT_PRELU);  // warp-uniform
        if (generic) {
          g2_chunk_ragged<DUAL, GN>(trow + c * 32, trow + BN + c * 32, nb, p, row, row_ok, gs, gss);
          if (GN) { gsd += (double)gs; gssd += (double)gss; }
          continue;
        }
        uint32_t v[32];
        tmem_ld32(trow + c * 32, v);
        uint32_t v2[32];
        if (DUAL) tmem_ld32(trow + BN + c * 32, v2);
        tmem_ld_wait();
        if (!row_ok && !(GN && row.cst)) continue;   // (staged stores are warp-collective: every lane takes part)
        if (DUAL) {
          g2_chunk<ACT_PRELU, true, false>(v, v2, nb, p, row, gs, gss);
        } else if (GN) {
          g2_chunk<ACT_NONE, false, true>(v, v2, nb, p, row, gs, gss);
          gsd += (double)gs; gssd += (double)gss;
        } else {
          switch (p.act) {
            case ACT_TANH: g2_chunk<ACT_TANH, false, false>(v, v2, nb, p, row, gs, gss); break;
            case ACT_RELU: g2_chunk<ACT_RELU, false, false>(v, v2, nb, p, row, gs, gss); break;
            case ACT_SIGMOID: g2_chunk<ACT_SIGMOID, false, false>(v, v2, nb, p, row, gs, gss); break;
            case ACT_PRELU: g2_chunk<ACT_PRELU, false, false>(v, v2, nb, p, row, gs, gss); break;
            case ACT_GELU: g2_chunk<ACT_GELU, false, false>(v, v2, nb, p, row, gs, gss); break;
            case ACT_GLU_PAIR: g2_chunk<ACT_GLU_PAIR, false, false>(v, v2, nb, p, row, gs, gss); break;
            default: g2_chunk<ACT_NONE, false, false>(v, v2, nb, p, row, gs, gss); break;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
    if (GN) gn_flush();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
int g2_choose_bn(int N) { return (N >= 256 && N % 256 == 0) ? 256 : 128; }

int launch_gemm2(const G2Problem& pr, cudaStream_t stream) {
  RFX_REQUIRE(pr.A.hi && pr.W.hi, "null operand");
  RFX_REQUIRE(pr.taps >= 1 && pr.taps <= 16, "1..16 taps");
  // Ktap need not be a multiple of 64: W is packed with every tap padded to ceil64(Ktap) columns, and the TMA unit
  // zero-fills A columns >= Ktap.
  RFX_REQUIRE(pr.Cf || pr.Chi, "at least one output");
  const int BN = pr.W.BN;
  RFX_REQUIRE(BN == 128 || BN == 256, "weights must be packed with BN 128 or 256");
  G2Params p{};
  const int Yo = pr.My > 0 ? pr.My : 1;
  int xt = pr.xt > 0 ? pr.xt : G2_BM;
  RFX_REQUIRE(xt <= G2_BM && G2_BM % xt == 0, "pixel tile width must divide 128");
  p.X = pr.M; p.Y = Yo; p.xt = xt; p.yt = G2_BM / xt;
  p.tiles_x = ceil_div(pr.M, p.xt);
  p.N = pr.N; p.batch = pr.batch;
  p.m_tiles = p.tiles_x * ceil_div(Yo, p.yt);
  p.n_tiles = ceil_div(pr.N, BN);
  p.kb_per_tap = ceil_div(pr.Ktap, G2_BK);
  p.ktap = pr.Ktap;
  {
    static const int dbg = [] { const char* e = getenv("RFX_G2_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
    p.dbg_skip = dbg;
  }
  p.taps = pr.taps;
  p.fast = get_matmul_precision() == 1 ? 1 : 0;
  for (int i = 0; i < pr.taps; ++i) { p.dx[i] = pr.row_off[i]; p.dy[i] = pr.row_off_y[i]; }
  p.Cf = pr.Cf; p.ldcf = pr.ldcf; p.bscf = pr.bscf; p.ldcy_f = pr.ldcf_y;
  p.Chi = pr.Chi; p.Clo = pr.Clo; p.ldcs = pr.ldcs; p.bscs = pr.bscs; p.ldcy_s = pr.ldcs_y;
  p.s1 = pr.epi.s1; p.t1 = pr.epi.t1; p.s2 = pr.epi.s2; p.t2 = pr.epi.t2; p.slope = pr.epi.slope; p.act = pr.epi.act;
  p.cf_accum = pr.cf_accum ? 1 : 0;
  p.cf_pre = pr.cf_pre_act ? 1 : 0;
  RFX_REQUIRE(!pr.cf_pre_act || (pr.Cf && pr.Chi && !pr.dual && !pr.cf_accum && !pr.gn_acc), "pre-activation output: needs both outputs, no dual / accumulate / statistics");
  RFX_REQUIRE(!pr.cf_accum || (pr.Cf && !pr.Chi && pr.epi.act == ACT_NONE && !pr.gn_acc), "accumulating output: fp32 only, no activation, no fused statistics");
  p.gn_acc = pr.gn_acc; p.gn_G = pr.gn_G > 0 ? pr.gn_G : 1; p.gn_per_x = pr.gn_per_x;
  p.gn_cmod = pr.gn_cmod > 0 ? pr.gn_cmod : pr.N;
  p.gn_cpg = p.gn_cmod / p.gn_G;
  if (pr.gn_acc) {
    RFX_REQUIRE(p.gn_cmod % p.gn_G == 0 && (p.gn_G == 1 || p.gn_cpg % 32 == 0), "fused GroupNorm statistics need 32-column aligned groups");
    RFX_REQUIRE(pr.epi.act == ACT_NONE && !pr.dual, "fused GroupNorm statistics are taken on the linear (bias-only) output");
    RFX_REQUIRE((long long)pr.batch * (pr.gn_per_x ? pr.M : 1) * p.gn_G < (1ll << 31), "fused GroupNorm statistics: too many segments");
  }
  RFX_REQUIRE(pr.W.Kpad >= p.kb_per_tap * p.taps * G2_BK, "packed weight K extent too small for taps * Ktap");
  // Tap groups (see G2Params): pixel tiles of one row, taps sorted by (dy, dx), a group = equal dy and x offsets within 8 pixels
  constexpr int G2_GROUP_SPAN = 8;
  p.ngroups = 0;
  int max_group = 1;
  {
    // opt-in: parity-green (tests under RFX_G2_TAPGROUPS=1) but no net gain on Hybrid Demucs -- see DESIGN 4.5
    static const bool allow = [] { const char* e = getenv("RFX_G2_TAPGROUPS"); return e && atoi(e) != 0; }();
    const int span = G2_GROUP_SPAN;
    if (allow && !pr.dual && p.yt == 1 && pr.taps > 1) {
      int order[16];
      for (int i = 0; i < pr.taps; ++i) order[i] = i;
      std::sort(order, order + pr.taps, [&](int a, int b) { return p.dy[a] != p.dy[b] ? p.dy[a] < p.dy[b] : p.dx[a] < p.dx[b]; });
      int ng = 0;
      for (int i = 0; i < pr.taps; ++i) {
        const int t = order[i];
        if (ng > 0 && p.dy[t] == p.g_dy[ng - 1] && p.dx[t] - p.g_dx0[ng - 1] <= span) {
          max_group = std::max(max_group, i + 1 - p.g_start[ng - 1]);
        } else {
          p.g_start[ng] = i; p.g_dx0[ng] = p.dx[t]; p.g_dy[ng] = p.dy[t];
          ++ng;
        }
        p.g_taps[i] = t;
      }
      p.g_start[ng] = pr.taps;
      if (max_group > 1) p.ngroups = ng;
    }
  }
  const int a_rows = p.ngroups > 0 ? G2_BM + G2_GROUP_SPAN : G2_BM;
  p.a_plane = a_rows * G2_BK * 2;
  const int a_stage = 2 * p.a_plane;   // = G2_A_STAGE without tap groups
  // shared-memory plan: narrow single-tile outputs fetch only the rows they need; small weight matrices stay resident
  constexpr int SMEM_CAP = 227 * 1024 - 2048;  // barriers + 1024-byte alignment slack
  const int KB = p.kb_per_tap * p.taps;
  p.n_box = p.n_tiles == 1 ? std::min(BN, ceil_div(pr.N, 16) * 16) : BN;
  p.b_bytes = 2 * p.n_box * G2_BK * 2;
  const long long w_total = (long long)KB * p.b_bytes;
  const bool resident = !pr.dual && p.n_tiles == 1 && w_total + 3 * a_stage <= SMEM_CAP;
  p.w_res_bytes = resident ? (int)w_total : 0;
  p.stage_bytes = a_stage + (resident ? 0 : (p.ngroups > 0 ? max_group : 1) * p.b_bytes);
  p.stages = std::min(G2_MAX_STAGES, (SMEM_CAP - p.w_res_bytes) / p.stage_bytes);
  if (p.ngroups > 0 && p.stages < 2) {  // the grouped stages do not fit: back to one tap per stage
    p.ngroups = 0;
    p.a_plane = G2_BM * G2_BK * 2;
    p.stage_bytes = G2_A_STAGE + (resident ? 0 : p.b_bytes);
    p.stages = std::min(G2_MAX_STAGES, (SMEM_CAP - p.w_res_bytes) / p.stage_bytes);
  }
  RFX_REQUIRE(p.stages >= 2, "gemm2: operand stages do not fit in shared memory");
  CUtensorMap mapA, mapW;
  int rc;
  if ((rc = make_split_map(&mapA, pr.A.hi, pr.Ktap, pr.A.rows, pr.A.rows_y > 0 ? pr.A.rows_y : 1, pr.batch, pr.A.ld, pr.A.ld_y,
                           pr.A.batch_stride, pr.A.plane_stride, p.ngroups > 0 ? G2_BM + G2_GROUP_SPAN : p.xt, p.yt)))
    return rc;
  if ((rc = make_split_map(&mapW, pr.W.hi, pr.W.Kpad, pr.W.Npad, 1, 1, pr.W.Kpad, 0, 0, (long long)pr.W.Npad * pr.W.Kpad, p.n_box, 1))) return rc;
  const int total = p.batch * p.m_tiles * p.n_tiles;
  {  // RFX_G2_TRACE=1: one line per launch (shape, tiling, shared-memory plan) to match against an ncu launch list
    static const bool trace = [] { const char* e = getenv("RFX_G2_TRACE"); return e && atoi(e) != 0; }();
    if (trace)
      fprintf(stderr, "g2trace batch=%d Y=%d X=%d N=%d Ktap=%d taps=%d BN=%d xt=%d tiles=%d n_tiles=%d resident=%d stages=%d act=%d gn=%d fp32out=%d split=%d groups=%d\n",
              pr.batch, Yo, pr.M, pr.N, pr.Ktap, pr.taps, BN, xt, total, p.n_tiles, resident ? 1 : 0, p.stages, pr.epi.act, pr.gn_acc ? 1 : 0,
              pr.Cf ? 1 : 0, pr.Chi ? 1 : 0, p.ngroups);
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = total < sms ? total : sms;
  if (pr.max_ctas > 0 && grid > pr.max_ctas) grid = pr.max_ctas;
  // Plans beyond the default (one CTA per SM, direct stores), both for launches with a resident W:
  //   two  -- thin layers (at most three k-blocks per tile, many tiles): two CTAs per SM, each with half the shared memory
  //   cst  -- fp32 rows staged through shared memory (GN instantiations; see G2Row), when the staging block fits beside at least
  //           two operand stages
  // Preference: two + cst, one + cst, two, default.
  bool two = false;
  p.cst = 0;
  int cst_bytes = 0;
  {
    static const bool allow_two = [] { const char* e = getenv("RFX_G2_TWO"); return !(e && atoi(e) == 0); }();
    static const bool allow_cst = [] { const char* e = getenv("RFX_G2_CST"); return !(e && atoi(e) == 0); }();
    constexpr int HALF_CAP = (227 * 1024) / 2 - 1024 - 1280;   // per-CTA dynamic shared memory when two CTAs share an SM
    const bool aligned = (pr.ldcf % 4 == 0) && (pr.bscf % 4 == 0) && (pr.ldcf_y % 4 == 0) && (((uintptr_t)pr.Cf & 15) == 0);
    const bool can_two = allow_two && BN == 128 && !pr.dual && resident && pr.max_ctas == 0 && KB <= 3 && total >= 4 * sms;
    const bool can_cst = allow_cst && pr.gn_acc && pr.Cf && !pr.Chi && !pr.dual && aligned && resident && pr.epi.act != ACT_GLU_PAIR;
    const int need2 = g2_epi_warps2(false, true) * G2_CST_WARP_BYTES, need1 = g2_epi_warps2(false, false) * G2_CST_WARP_BYTES;
    if (can_two && can_cst && p.w_res_bytes + 2 * p.stage_bytes + need2 <= HALF_CAP) {
      two = true; p.cst = 1; cst_bytes = need2;
      p.stages = std::min(4, (HALF_CAP - p.w_res_bytes - need2) / p.stage_bytes);
    } else if (can_cst && p.w_res_bytes + 2 * p.stage_bytes + need1 <= SMEM_CAP) {
      p.cst = 1; cst_bytes = need1;
      p.stages = std::min(p.stages, (SMEM_CAP - p.w_res_bytes - need1) / p.stage_bytes);
    } else if (can_two && p.w_res_bytes + 2 * p.stage_bytes <= HALF_CAP) {
      two = true;
      p.stages = std::min(4, (HALF_CAP - p.w_res_bytes) / p.stage_bytes);
    }
    if (two) grid = std::min(total, 2 * sms);
  }
  p.chunk = pr.gn_acc ? ceil_div(total, grid) : 0;
  const int smem = p.w_res_bytes + p.stages * p.stage_bytes + 1024 + 256 + cst_bytes;
  if (pr.dual) {
    RFX_REQUIRE(BN == 256 && pr.N <= 256 * p.n_tiles && pr.taps >= 2, "dual-accumulator mode needs BN = 256 and >= 2 taps");
    RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<256, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    gemm2_kernel<256, true, false, false><<<grid, g2_threads(true), smem, stream>>>(mapA, mapW, p);
  } else if (BN == 256) {
    if (pr.gn_acc) {
      RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<256, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      gemm2_kernel<256, false, true, false><<<grid, g2_threads(false), smem, stream>>>(mapA, mapW, p);
    } else {
      RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<256, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      gemm2_kernel<256, false, false, false><<<grid, g2_threads(false), smem, stream>>>(mapA, mapW, p);
    }
  } else if (two) {
    if (pr.gn_acc) {
      RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<128, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      gemm2_kernel<128, false, true, true><<<grid, g2_threads2(false, true), smem, stream>>>(mapA, mapW, p);
    } else {
      RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<128, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      gemm2_kernel<128, false, false, true><<<grid, g2_threads2(false, true), smem, stream>>>(mapA, mapW, p);
    }
  } else {
    if (pr.gn_acc) {
      RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<128, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      gemm2_kernel<128, false, true, false><<<grid, g2_threads(false), smem, stream>>>(mapA, mapW, p);
    } else {
      RFX_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<128, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      gemm2_kernel<128, false, false, false><<<grid, g2_threads(false), smem, stream>>>(mapA, mapW, p);
    }
  }
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int g_matmul_precision = 0;
void set_matmul_precision(int mode) { g_matmul_precision = mode == 1 ? 1 : 0; }
int get_matmul_precision() { return g_matmul_precision; }

size_t split_weight_elems(int N, int K, int BN) {
  return (size_t)(ceil_div(N, BN) * BN) * (size_t)(ceil_div(K, G2_BK) * G2_BK);
}

int pack_split_weights(const float* W, long long ldw, int N, int K, int BN, __nv_bfloat16* dst, SplitW* out, cudaStream_t stream) {
  RFX_REQUIRE(BN == 128 || BN == 256, "BN must be 128 or 256");
  out->N = N; out->K = K; out->BN = BN;
  out->Npad = ceil_div(N, BN) * BN;
  out->Kpad = ceil_div(K, G2_BK) * G2_BK;
  out->hi = dst;
  out->lo = dst + (size_t)out->Npad * out->Kpad;
  return launch_split_rows(W, ldw, N, K, dst, dst + (size_t)out->Npad * out->Kpad, out->Kpad, out->Npad, out->Kpad, stream);
}

}  // namespace rfx
