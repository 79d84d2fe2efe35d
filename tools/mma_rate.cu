// Micro-benchmark: legacy mma.sync (HMMA) throughput on this GPU, for sizing the LSTM recurrence.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void hmma_loop(int iters, float* out, long long* cycles) {
  uint32_t a0 = threadIdx.x, a1 = 0x3f803f80, a2 = 0x3f803f80, a3 = 0x3f803f80, b0 = 0x3f803f80, b1 = 0x3f803f80;
  float c[8][4] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void hmma_chain(int iters, float* out, long long* cycles) {  // dependent chain: latency
  uint32_t a0 = threadIdx.x, a1 = 0x3f803f80, a2 = 0x3f803f80, a3 = 0x3f803f80, b0 = 0x3f803f80, b1 = 0x3f803f80;
  float c[4] = {};
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c[0] + c[1] + c[2] + c[3];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void ffma_loop(int iters, float* out, long long* cycles) {
  float c[16]; float a = threadIdx.x * 1e-3f, b = 1.0001f;
  for (int j = 0; j < 16; ++j) c[j] = j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) c[j] = fmaf(c[j], b, a);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 16; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int warps : {1, 2, 4, 8, 16}) {
    long long h[148];
    hmma_loop<<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize();
    hmma_loop<<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double macs = (double)iters * 8 * warps * 16 * 8 * 16;
    printf("HMMA m16n8k16 bf16: %2d warps/SM: %lld cycles -> %.0f MAC/clk/SM (%.2f cyc per mma per SM)\n", warps, h[0], macs / h[0], (double)h[0] / (iters * 8.0 * warps));
    ffma_loop<<<148, warps * 32>>>(iters, out, cyc); cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("FFMA              : %2d warps/SM: %lld cycles -> %.1f FMA lanes/clk/SM\n", warps, h[0], (double)iters * 16 * warps * 32 / h[0]);
  }
  long long h[148];
  hmma_chain<<<148, 32>>>(iters, out, cyc); cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("HMMA dependent chain latency: %.1f cycles\n", (double)h[0] / iters);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
