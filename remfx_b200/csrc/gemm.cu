// Plain fp32 FFMA GEMM  C = epilogue(A . W^T)  with the same epilogue contract as the tensor-core engine
// (gemm2.cu).  It is NOT on the product path: it exists as an on-device cross-check for the tcgen05 kernels
// (tests/test_gpu_gemm_lstm.py) and as a debugging aid (rfx_gemm impl = 1).
#include "kernels.h"

namespace rfx {

// ------------------------------------------------------------------------------------------------
// Epilogue
// ------------------------------------------------------------------------------------------------
struct EpiDev {
  const float* s1;
  const float* t1;
  const float* s2;
  const float* t2;
  int act;
};

__device__ __forceinline__ float apply_epi(float v, int n, const EpiDev& e) {
  if (e.s1) v *= e.s1[n];
  if (e.t1) v += e.t1[n];
  if (e.s2) v *= e.s2[n];
  if (e.t2) v += e.t2[n];
  if (e.act == ACT_TANH) v = tanhf(v);
  else if (e.act == ACT_RELU) v = fmaxf(v, 0.0f);
  else if (e.act == ACT_SIGMOID) v = sigmoidf_acc(v);
  return v;
}

// ------------------------------------------------------------------------------------------------
// fp32 FFMA GEMM (cross-check path)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int lda, int M, const float* __restrict__ W, int ldw,
                                                        int N, int K, float* __restrict__ C, int ldc, EpiDev epi) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  const int lr = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + lk + e;
      const int m = m0 + lr, n = n0 + lr;
      As[lk + e][lr] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.0f;
      Ws[lk + e][lr] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[(size_t)m * ldc + n] = apply_epi(acc[i][j], n, epi);
    }
  }
}

int launch_gemm_simt(const float* A, int lda, int M, const float* W, int ldw, int N, int K, float* C, int ldc, const Epilogue& e,
                     cudaStream_t stream) {
  RFX_REQUIRE(M > 0 && N > 0 && K > 0, "positive sizes");
  dim3 grid(ceil_div(M, 64), ceil_div(N, 64));
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(A, lda, M, W, ldw, N, K, C, ldc, EpiDev{e.s1, e.t1, e.s2, e.t2, e.act});
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rfx
