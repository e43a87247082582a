// Pointwise stages of the adversarial channel-classifier head:
//   x -> GRL -> Linear -> Dropout(0.3) -> ReLU -> Linear -> ReLU -> CrossEntropy        model.py:976-1023,
//   main_train.py:251,377-403,420-453.  The two Linear layers run on air_linear_fwd / air_linear_bwd (csrc/head.cu);
// this file holds what sits between and after them.  Everything is fp32 on (batch, <= 128 + classes) rows: a few
// hundred KB per step, HBM- / launch-latency bound by construction.
//
// STATUS: compiled and checked against the oracle on the CPU side only; the GPU parity tests
// (tests/test_adv_gpu.py) have not run on hardware yet and are marked accordingly.
#include "common.cuh"

namespace air_adv {

// Counter-based keep decision (the reference draws torch's Philox stream; parity is defined for a GIVEN mask, the
// generated one only has to be Bernoulli(1 - p), reproducible from (seed, element)).
__device__ __forceinline__ uint32_t mix(uint64_t v) {
  v ^= v >> 33; v *= 0xff51afd7ed558ccdull;
  v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ull;
  v ^= v >> 33;
  return (uint32_t)(v >> 32);
}

// y = relu(x * keep / (1 - p)); keep is read (generate == 0) or drawn and written (generate != 0).
__global__ void dropout_relu_fwd_kernel(const float* __restrict__ x, uint8_t* __restrict__ keep, float* __restrict__ y,
                                        long long n, float p, int generate, unsigned long long seed) {
  const float scale = 1.f / (1.f - p);
  const uint32_t thresh = (uint32_t)((double)p * 4294967296.0);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint8_t k;
    if (generate) {
      k = mix(seed * 0x9e3779b97f4a7c15ull + (unsigned long long)i) >= thresh ? 1 : 0;
      keep[i] = k;
    } else {
      k = keep[i];
    }
    const float v = k ? x[i] * scale : 0.f;
    y[i] = v > 0.f ? v : 0.f;
  }
}

// dx = dy * [x * keep > 0] * keep / (1 - p)
__global__ void dropout_relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                        const uint8_t* __restrict__ keep, float* __restrict__ dx, long long n, float p) {
  const float scale = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dx[i] = (keep[i] && x[i] > 0.f) ? dy[i] * scale : 0.f;
}

// One warp per row: logits = relu(z); loss += -log softmax[label] / B; dz = grad_scale * (softmax - onehot) / B * [z > 0];
// correct += (argmax == label) with torch.max's first-maximum rule.  loss / correct are accumulated with atomics:
// loss in fixed-point-free fp32 would depend on the order, so each row's term goes through a double atomicAdd.
__global__ void relu_ce_kernel(const float* __restrict__ z, const long long* __restrict__ labels, int B, int C,
                               float grad_scale, double* __restrict__ loss_sum, int* __restrict__ correct,
                               float* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  const float* zr = z + (long long)row * C;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = fmaxf(zr[c], 0.f);
    if (v > mx) { mx = v; arg = c; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(fmaxf(zr[c], 0.f) - mx);
  se = warp_sum(se);
  const int lab = (int)labels[row];
  const float inv = 1.f / se, invB = 1.f / (float)B;
  if (dz) {
    float* dr = dz + (long long)row * C;
    for (int c = lane; c < C; c += 32) {
      const float zc = zr[c];
      const float sm = expf(fmaxf(zc, 0.f) - mx) * inv;
      dr[c] = zc > 0.f ? grad_scale * (sm - (c == lab ? 1.f : 0.f)) * invB : 0.f;
    }
  }
  if (lane == 0) {
    // a label outside [0, C) is the caller's error (torch raises): no out-of-bounds read, the loss turns NaN
    const float ll = (lab >= 0 && lab < C) ? fmaxf(zr[lab], 0.f) - mx - logf(se) : NAN;
    if (loss_sum) atomicAdd(loss_sum, -(double)ll * (double)invB);
    if (correct && arg == lab) atomicAdd(correct, 1);
  }
}

}  // namespace air_adv

using namespace air_adv;

static int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  return (int)(g > 148 * 8 ? 148 * 8 : (g < 1 ? 1 : g));
}

extern "C" int air_dropout_relu_fwd(const float* x, unsigned char* keep, float* y, long long n, float p, int generate,
                                    unsigned long long seed, cudaStream_t stream) {
  if (!x || !keep || !y || n < 0 || !(p >= 0.f && p < 1.f)) return AIR_ERR_ARG;
  if (n == 0) return AIR_OK;
  dropout_relu_fwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, keep, y, n, p, generate, seed);
  return air_launch_status();
}

extern "C" int air_dropout_relu_bwd(const float* dy, const float* x, const unsigned char* keep, float* dx, long long n,
                                    float p, cudaStream_t stream) {
  if (!dy || !x || !keep || !dx || n < 0 || !(p >= 0.f && p < 1.f)) return AIR_ERR_ARG;
  if (n == 0) return AIR_OK;
  dropout_relu_bwd_kernel<<<grid_for(n, 256), 256, 0, stream>>>(dy, x, keep, dx, n, p);
  return air_launch_status();
}

// loss_sum (double) and correct (int) are ACCUMULATED: zero them before the first call of a step.
extern "C" int air_relu_ce_fwd_bwd(const float* z, const long long* labels, int B, int C, float grad_scale,
                                   double* loss_sum, int* correct, float* dz, cudaStream_t stream) {
  if (!z || !labels || B < 1 || C < 1) return AIR_ERR_ARG;
  const int rows_per_block = 8;
  relu_ce_kernel<<<(B + rows_per_block - 1) / rows_per_block, 32 * rows_per_block, 0, stream>>>(
      z, labels, B, C, grad_scale, loss_sum, correct, dz);
  return air_launch_status();
}
