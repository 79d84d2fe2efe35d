// L5: the optimiser half of the training step -- global gradient norm, clip-by-norm and AdamW fused over ONE flat fp32 bucket.
//
// Reference behaviour (remfx/models.py:185-206 + cfg/config.yaml:110-120): torch.optim.AdamW(lr 1e-4, betas (0.95, 0.999),
// eps 1e-6, weight_decay 1e-3) stepped every batch, Lightning `gradient_clip_val: 10.0` (= torch.nn.utils.clip_grad_norm_:
// coef = min(1, max_norm / (||g||_2 + 1e-6)) over ALL parameters), fp32, no accumulation.
//
// Layout: parameters, gradients and both moments live in four flat, 16-byte aligned fp32 buffers of the same length (the
// host side re-points every nn.Parameter / .grad at a view of them), so a whole-model step is three launches:
//   memset(acc) -> grad_sumsq_kernel (fp64 accumulate) -> adamw_kernel (reads the norm, clips, updates p/m/v in one pass).
// Both kernels are pure streaming: 4 B/elem read for the norm, 16 B read + 12 B write per element for the update.
#include "common.cuh"
#include "kernels.h"
#include "../../include/remfx_b200.h"

namespace rfx {

__global__ void __launch_bounds__(512) grad_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ acc) {
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float s0 = 0.0f, s1 = 0.0f;
  double d = 0.0;
  int cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s0 = fmaf(v.x, v.x, fmaf(v.y, v.y, s0));
    s1 = fmaf(v.z, v.z, fmaf(v.w, v.w, s1));
    if (++cnt == 64) { d += (double)s0 + (double)s1; s0 = s1 = 0.0f; cnt = 0; }  // bound the fp32 run length
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; s0 = fmaf(v, v, s0); }
  d += (double)s0 + (double)s1;
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  __shared__ double part[16];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) atomicAdd(acc, t);
  }
}

struct AdamWArgs {
  float* p; const float* g; float* m; float* v;
  long long n;
  float lr, beta1, beta2, eps, wd;
  float decay, step_size, bc2_sqrt;  // 1 - lr * wd, lr / (1 - beta1^t), sqrt(1 - beta2^t): formed in fp64 on the host like torch does
  float grad_scale;          // 1 / world_size after a sum all-reduce (1 otherwise)
  float max_norm;            // <= 0: no clipping
  const double* sumsq;       // sum of squares of the UNSCALED gradient (device), required when max_norm > 0
  float* norm_out;           // optional: the (scaled) total norm, as clip_grad_norm_ returns it
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a, float coef) {
  // torch/optim/adamw.py (_single_tensor_adam with decoupled decay): same operation order
  g *= coef;
  p *= a.decay;
  m = m + (g - m) * (1.0f - a.beta1);          // exp_avg.lerp_(grad, 1 - beta1)
  v = v * a.beta2 + (1.0f - a.beta2) * g * g;  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
}

__global__ void __launch_bounds__(512) adamw_kernel(AdamWArgs a) {
  float coef = a.grad_scale;
  if (a.max_norm > 0.0f) {
    const float total = (float)sqrt(*a.sumsq) * a.grad_scale;
    coef *= fminf(a.max_norm / (total + 1e-6f), 1.0f);
    if (a.norm_out && blockIdx.x == 0 && threadIdx.x == 0) *a.norm_out = total;
  }
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.g) + i);
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    adamw_one(p.x, g.x, m.x, v.x, a, coef);
    adamw_one(p.y, g.y, m.y, v.y, a, coef);
    adamw_one(p.z, g.z, m.z, v.z, a, coef);
    adamw_one(p.w, g.w, m.w, v.w, a, coef);
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    adamw_one(a.p[i], a.g[i], a.m[i], a.v[i], a, coef);
  }
}

static int optim_grid(long long n, int threads) {
  const long long want = (n / 4 + threads - 1) / threads;
  const long long cap = 148LL * 4;  // four resident CTAs of 512 threads per SM
  return (int)std::max<long long>(1, std::min<long long>(want, cap));
}

}  // namespace rfx

using namespace rfx;

extern "C" {

size_t rfx_optim_workspace_bytes(void) { return 256; }

int rfx_grad_sumsq(const float* grad, long long n, void* workspace, int accumulate, void* stream) {
  RFX_REQUIRE(grad && workspace, "null argument");
  RFX_REQUIRE(n > 0, "positive size");
  RFX_REQUIRE(((uintptr_t)grad & 15) == 0 && ((uintptr_t)workspace & 7) == 0, "gradient bucket must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (!accumulate) RFX_CHECK_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double), s));
  grad_sumsq_kernel<<<optim_grid(n, 512), 512, 0, s>>>(grad, n, reinterpret_cast<double*>(workspace));
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int rfx_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, float max_norm, const void* workspace, float* total_norm,
                   void* stream) {
  RFX_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null argument");
  RFX_REQUIRE(n > 0 && step >= 1, "positive size and 1-based step");
  RFX_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, "buckets must be 16-byte aligned");
  RFX_REQUIRE(max_norm <= 0.0f || workspace, "clipping needs the workspace rfx_grad_sumsq filled");
  AdamWArgs a{};
  a.p = param; a.g = grad; a.m = exp_avg; a.v = exp_avg_sq; a.n = n;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.wd = weight_decay;
  a.decay = (float)(1.0 - (double)lr * (double)weight_decay);
  a.step_size = (float)((double)lr / (1.0 - pow((double)beta1, step)));
  a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  a.grad_scale = grad_scale; a.max_norm = max_norm;
  a.sumsq = reinterpret_cast<const double*>(workspace);
  a.norm_out = total_norm;
  adamw_kernel<<<optim_grid(n, 512), 512, 0, (cudaStream_t)stream>>>(a);
  RFX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
