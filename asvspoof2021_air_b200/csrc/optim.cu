// Optimiser steps on flat fp32 buffers (HBM-bound, one pass):
//   air_adam_l2_step : torch.optim.Adam with coupled L2 weight decay (main_train.py:175,408)
//   air_sgd_step     : torch.optim.SGD for the OC-Softmax centre (main_train.py:272,409)
// `grad_scale` folds the 1/world_size of the data-parallel gradient average into the step.
#include <algorithm>
#include "common.cuh"

namespace air_optim {

__global__ void adam_l2_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                               float* __restrict__ v, long long n, float lr_over_bc1, float inv_sqrt_bc2,
                               float beta1, float beta2, float eps, float wd, float grad_scale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(wd, pi, g[i] * grad_scale);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] = pi - lr_over_bc1 * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, long long n, float lr, float grad_scale) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    p[i] = fmaf(-lr * grad_scale, g[i], p[i]);
}

}  // namespace air_optim

extern "C" int air_adam_l2_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, float grad_scale,
                                cudaStream_t stream) {
  if (!p || !g || !m || !v || n <= 0 || step < 1) return AIR_ERR_ARG;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  air_optim::adam_l2_kernel<<<blocks, 256, 0, stream>>>(p, g, m, v, n, static_cast<float>(lr / bc1),
                                                        static_cast<float>(1.0 / sqrt(bc2)), beta1, beta2, eps,
                                                        weight_decay, grad_scale);
  return air_launch_status();
}

extern "C" int air_sgd_step(float* p, const float* g, long long n, float lr, float grad_scale, cudaStream_t stream) {
  if (!p || !g || n <= 0) return AIR_ERR_ARG;
  const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
  air_optim::sgd_kernel<<<blocks, 256, 0, stream>>>(p, g, n, lr, grad_scale);
  return air_launch_status();
}
