// Train/eval BatchNorm (+ReLU) forward and backward on channels-last bf16 activations.
//
// Replaces nn.BatchNorm2d / nn.BatchNorm1d + F.relu (resnet.py:54-69,132,141,175-182;
// ecapa_tdnn.py:40,51,57,68-88,113,161) and their autograd backward.  HBM-bound: every kernel
// streams [M][C] rows with 16-byte loads (8 channels per thread) and reduces per channel in
// registers -> shared memory -> one fp64 atomic per channel per CTA.
//
//   order 0 ("pre-activation", ResNet):  y = relu(bn(x))
//   order 1 (ECAPA: conv -> ReLU -> BN): y = bn(x) with x = relu(conv) already applied upstream;
//            the backward additionally masks the result with (x > 0).
#include <algorithm>
#include "common.cuh"

namespace air_bn {

constexpr int THREADS = 256;

// ---------------------------------------------------------------------------------------------
// per-channel sum / sum of squares:  sums[0..C) += sum_m x[m][c], sums[C..2C) += sum_m x^2
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS) bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                            long long M, int C, double* __restrict__ sums) {
  extern __shared__ float sh[];                 // [2][THREADS][8]
  const int cpr = C >> 3;                       // 16-byte chunks per row
  const int rows_per_it = THREADS / cpr;        // C <= 2048
  const int tc = threadIdx.x % cpr, tr = threadIdx.x / cpr;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (tr < rows_per_it) {
    for (long long m = static_cast<long long>(blockIdx.x) * rows_per_it + tr; m < M;
         m += static_cast<long long>(gridDim.x) * rows_per_it) {
      const bf16x8 v = *reinterpret_cast<const bf16x8*>(x + m * ld + tc * 8);
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
  float* ss = sh;
  float* sq = sh + THREADS * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) { ss[threadIdx.x * 8 + i] = s[i]; sq[threadIdx.x * 8 + i] = q[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += THREADS) {
    const int chunk = c >> 3, e = c & 7;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < rows_per_it; ++r) { a += ss[(r * cpr + chunk) * 8 + e]; b += sq[(r * cpr + chunk) * 8 + e]; }
    atomicAdd(&sums[c], static_cast<double>(a));
    atomicAdd(&sums[C + c], static_cast<double>(b));
  }
}

// ---------------------------------------------------------------------------------------------
// y = [relu](gamma * (x - mean) * invstd + beta).  training: batch statistics from `sums`
// (biased variance), block 0 saves mean/invstd and updates the running statistics
// (momentum, unbiased variance).  eval: running statistics.
// ---------------------------------------------------------------------------------------------
struct ApplyParams {
  const __nv_bfloat16* x; long long x_ld; __nv_bfloat16* y; long long y_ld; long long M; int C;
  const double* sums; const float* gamma; const float* beta; float eps; int relu; int training;
  float* save_mean; float* save_invstd; float* running_mean; float* running_var; float momentum;
  const __nv_bfloat16* add; long long add_ld; __nv_bfloat16* y2; long long y2_ld;   // optional: y2 = y + add
};

__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const ApplyParams p) {
  extern __shared__ float sh[];                 // scale[C], shift[C]
  float* scale = sh;
  float* shift = sh + p.C;
  for (int c = threadIdx.x; c < p.C; c += THREADS) {
    float mean, invstd;
    if (p.training) {
      const double mu = p.sums[c] / static_cast<double>(p.M);
      double var = p.sums[p.C + c] / static_cast<double>(p.M) - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = static_cast<float>(mu);
      invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
      if (blockIdx.x == 0) {
        if (p.save_mean) { p.save_mean[c] = mean; p.save_invstd[c] = invstd; }
        if (p.running_mean) {
          const double unb = p.M > 1 ? var * static_cast<double>(p.M) / static_cast<double>(p.M - 1) : var;
          p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
          p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * static_cast<float>(unb);
        }
      }
    } else {
      mean = p.running_mean[c];
      invstd = rsqrtf(p.running_var[c] + p.eps);
    }
    const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
    scale[c] = g * invstd;
    shift[c] = b - mean * g * invstd;
  }
  __syncthreads();
  const int cpr = p.C >> 3;
  const long long total = p.M * cpr;
  for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * THREADS) {
    const long long m = i / cpr;
    const int c0 = static_cast<int>(i - m * cpr) * 8;
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(p.x + m * p.x_ld + c0), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = fmaf(f[k], scale[c0 + k], shift[c0 + k]);
      if (p.relu) f[k] = fmaxf(f[k], 0.f);
    }
    const bf16x8 yv = pack8(f);
    *reinterpret_cast<bf16x8*>(p.y + m * p.y_ld + c0) = yv;
    if (p.y2) {                                   // Res2 branch input: sp_{i} + spx[i+1] (ecapa_tdnn.py:77-80)
      float a[8];
      unpack8(yv, f);                             // the consumer adds the ROUNDED branch output
      unpack8(*reinterpret_cast<const bf16x8*>(p.add + m * p.add_ld + c0), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += a[k];
      *reinterpret_cast<bf16x8*>(p.y2 + m * p.y2_ld + c0) = pack8(f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1: rsum[0..C) += sum g, rsum[C..2C) += sum g * xhat
//   order 0: g = dy * (gamma*xhat + beta > 0)        order 1: g = dy
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  const __nv_bfloat16* dy; long long dy_ld; const __nv_bfloat16* x; long long x_ld;
  const __nv_bfloat16* add; long long add_ld;          // optional extra gradient added to dx
  __nv_bfloat16* dx; long long dx_ld; long long M; int C; int order;
  const float* mean; const float* invstd; const float* gamma; const float* beta;
  double* rsum; float* dgamma; float* dbeta;
  float* dbias;                                        // optional: dbias[c] += sum_m dx[m][c] (bias of the producing conv)
};

__global__ void __launch_bounds__(THREADS) bn_bwd_reduce_kernel(const BwdParams p) {
  extern __shared__ float sh[];                 // [2][THREADS][8]
  const int cpr = p.C >> 3;
  const int rows_per_it = THREADS / cpr;
  const int tc = threadIdx.x % cpr, tr = threadIdx.x / cpr;
  float s[8], q[8], mu[8], is[8], ga[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (tr < rows_per_it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = tc * 8 + i;
      mu[i] = p.mean[c]; is[i] = p.invstd[c]; ga[i] = p.gamma ? p.gamma[c] : 1.f; be[i] = p.beta ? p.beta[c] : 0.f;
    }
    for (long long m = static_cast<long long>(blockIdx.x) * rows_per_it + tr; m < p.M;
         m += static_cast<long long>(gridDim.x) * rows_per_it) {
      float g[8], xv[8];
      unpack8(*reinterpret_cast<const bf16x8*>(p.dy + m * p.dy_ld + tc * 8), g);
      unpack8(*reinterpret_cast<const bf16x8*>(p.x + m * p.x_ld + tc * 8), xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xh = (xv[i] - mu[i]) * is[i];
        float gi = g[i];
        if (p.order == 0 && !(fmaf(ga[i], xh, be[i]) > 0.f)) gi = 0.f;
        s[i] += gi; q[i] = fmaf(gi, xh, q[i]);
      }
    }
  }
  float* ss = sh;
  float* sq = sh + THREADS * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) { ss[threadIdx.x * 8 + i] = s[i]; sq[threadIdx.x * 8 + i] = q[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += THREADS) {
    const int chunk = c >> 3, e = c & 7;
    float a = 0.f, b = 0.f;
    for (int r = 0; r < rows_per_it; ++r) { a += ss[(r * cpr + chunk) * 8 + e]; b += sq[(r * cpr + chunk) * 8 + e]; }
    atomicAdd(&p.rsum[c], static_cast<double>(a));
    atomicAdd(&p.rsum[p.C + c], static_cast<double>(b));
  }
}

// backward, pass 2: dx = gamma*invstd*(g - sum_g/M - xhat*sum_gx/M) [* (x > 0) for order 1] [+ add]
// block 0 also writes dgamma = sum g*xhat, dbeta = sum g (accumulating into the gradient buffer).
__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const BwdParams p) {
  extern __shared__ float sh[];                 // k1[C] = gamma*invstd, k2[C] = sum_g/M, k3[C] = sum_gx/M, mean, invstd, gamma, beta
  float* k1 = sh; float* k2 = sh + p.C; float* k3 = sh + 2 * p.C;
  float* smu = sh + 3 * p.C; float* sis = sh + 4 * p.C; float* sga = sh + 5 * p.C; float* sbe = sh + 6 * p.C;
  for (int c = threadIdx.x; c < p.C; c += THREADS) {
    const float ga = p.gamma ? p.gamma[c] : 1.f;
    const double sg = p.rsum[c], sgx = p.rsum[p.C + c];
    k1[c] = ga * p.invstd[c];
    k2[c] = static_cast<float>(sg / static_cast<double>(p.M));
    k3[c] = static_cast<float>(sgx / static_cast<double>(p.M));
    smu[c] = p.mean[c]; sis[c] = p.invstd[c]; sga[c] = ga; sbe[c] = p.beta ? p.beta[c] : 0.f;
    if (blockIdx.x == 0) {
      if (p.dgamma) p.dgamma[c] += static_cast<float>(sgx);
      if (p.dbeta) p.dbeta[c] += static_cast<float>(sg);
    }
  }
  __syncthreads();
  const int cpr = p.C >> 3;
  const long long total = p.M * cpr;
  float bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bs[k] = 0.f;
  int my_c0 = -1;
  for (long long i = static_cast<long long>(blockIdx.x) * THREADS + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * THREADS) {
    const long long m = i / cpr;
    const int c0 = static_cast<int>(i - m * cpr) * 8;
    my_c0 = c0;                       // constant per thread when (gridDim.x * THREADS) % cpr == 0 (launcher guarantees it with dbias)
    float g[8], xv[8], ad[8];
    unpack8(*reinterpret_cast<const bf16x8*>(p.dy + m * p.dy_ld + c0), g);
    unpack8(*reinterpret_cast<const bf16x8*>(p.x + m * p.x_ld + c0), xv);
    if (p.add) unpack8(*reinterpret_cast<const bf16x8*>(p.add + m * p.add_ld + c0), ad);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      const float xh = (xv[k] - smu[c]) * sis[c];
      float gi = g[k];
      if (p.order == 0 && !(fmaf(sga[c], xh, sbe[c]) > 0.f)) gi = 0.f;
      float d = k1[c] * (gi - k2[c] - xh * k3[c]);
      if (p.order == 1 && !(xv[k] > 0.f)) d = 0.f;
      if (p.add) d += ad[k];
      g[k] = d;
      bs[k] += d;
    }
    *reinterpret_cast<bf16x8*>(p.dx + m * p.dx_ld + c0) = pack8(g);
  }
  if (p.dbias) {
    // block-level reduction first (one atomic per channel per CTA): thread t owns chunk
    // ((blockIdx.x * THREADS + t) % cpr) for the whole loop (launcher guarantees (gridDim.x * THREADS) % cpr == 0)
    float* red = sh + 7 * p.C;                  // [THREADS][8]
    (void)my_c0;
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = bs[k];
    __syncthreads();
    const int first = static_cast<int>((static_cast<long long>(blockIdx.x) * THREADS) % cpr);
    for (int c = threadIdx.x; c < p.C; c += THREADS) {
      const int chunk = c >> 3, e = c & 7;
      int t0 = chunk - first; if (t0 < 0) t0 += cpr;
      float a = 0.f;
      for (int t = t0; t < THREADS; t += cpr) a += red[t * 8 + e];
      atomicAdd(&p.dbias[c], a);
    }
  }
}

static int grid_for(long long work_items, int per_block, int num_sms) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms > 0 ? num_sms : 148) * 8;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace air_bn

using namespace air_bn;

static bool bn_args_ok(long long M, int C, long long ld) {
  return M > 0 && C >= 8 && C % 8 == 0 && C <= 2048 && ld % 8 == 0 && (THREADS % (C / 8) == 0 || C / 8 > THREADS ? (C / 8 <= THREADS) : true);
}

extern "C" int air_bn_stats(const void* x, long long x_ld, long long M, int C, double* sums, int num_sms, cudaStream_t stream) {
  if (!x || !sums || !bn_args_ok(M, C, x_ld)) return AIR_ERR_ARG;
  const int rows_per_it = THREADS / (C / 8);
  if (rows_per_it < 1) return AIR_ERR_UNSUPPORTED;
  const int grid = grid_for(M, rows_per_it * 16, num_sms);
  bn_stats_kernel<<<grid, THREADS, 2 * THREADS * 8 * sizeof(float), stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), x_ld, M, C, sums);
  return air_launch_status();
}

extern "C" int air_bn_apply_add(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                                const double* sums, const float* gamma, const float* beta, float eps, int relu,
                                int training, float* save_mean, float* save_invstd, float* running_mean,
                                float* running_var, float momentum, const void* add, long long add_ld, void* y2,
                                long long y2_ld, int num_sms, cudaStream_t stream) {
  if (!x || !y || !bn_args_ok(M, C, x_ld) || y_ld % 8 != 0) return AIR_ERR_ARG;
  if (training ? !sums : (!running_mean || !running_var)) return AIR_ERR_ARG;
  if (y2 && (!add || add_ld % 8 != 0 || y2_ld % 8 != 0)) return AIR_ERR_ARG;
  ApplyParams p{reinterpret_cast<const __nv_bfloat16*>(x), x_ld, reinterpret_cast<__nv_bfloat16*>(y), y_ld, M, C,
                sums, gamma, beta, eps, relu, training, save_mean, save_invstd, running_mean, running_var, momentum,
                reinterpret_cast<const __nv_bfloat16*>(add), add_ld, reinterpret_cast<__nv_bfloat16*>(y2), y2_ld};
  const int grid = grid_for(M * (C / 8), THREADS * 8, num_sms);
  bn_apply_kernel<<<grid, THREADS, 2 * C * sizeof(float), stream>>>(p);
  return air_launch_status();
}

extern "C" int air_bn_apply(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                            const double* sums, const float* gamma, const float* beta, float eps, int relu,
                            int training, float* save_mean, float* save_invstd, float* running_mean,
                            float* running_var, float momentum, int num_sms, cudaStream_t stream) {
  return air_bn_apply_add(x, x_ld, y, y_ld, M, C, sums, gamma, beta, eps, relu, training, save_mean, save_invstd,
                          running_mean, running_var, momentum, nullptr, 0, nullptr, 0, num_sms, stream);
}

extern "C" int air_bn_bwd_bias(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                               void* dx, long long dx_ld, long long M, int C, int order,
                               const float* mean, const float* invstd, const float* gamma, const float* beta,
                               double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, cudaStream_t stream);

extern "C" int air_bn_bwd(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                          void* dx, long long dx_ld, long long M, int C, int order,
                          const float* mean, const float* invstd, const float* gamma, const float* beta,
                          double* rsum, float* dgamma, float* dbeta, int num_sms, cudaStream_t stream) {
  return air_bn_bwd_bias(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum, dgamma,
                         dbeta, nullptr, num_sms, stream);
}

extern "C" int air_bn_bwd_bias(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                               void* dx, long long dx_ld, long long M, int C, int order,
                               const float* mean, const float* invstd, const float* gamma, const float* beta,
                               double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, cudaStream_t stream) {
  if (!dy || !x || !dx || !mean || !invstd || !rsum || !bn_args_ok(M, C, x_ld)) return AIR_ERR_ARG;
  if (dy_ld % 8 != 0 || dx_ld % 8 != 0 || (add && add_ld % 8 != 0)) return AIR_ERR_ARG;
  BwdParams p{reinterpret_cast<const __nv_bfloat16*>(dy), dy_ld, reinterpret_cast<const __nv_bfloat16*>(x), x_ld,
              reinterpret_cast<const __nv_bfloat16*>(add), add_ld, reinterpret_cast<__nv_bfloat16*>(dx), dx_ld, M, C, order,
              mean, invstd, gamma, beta, rsum, dgamma, dbeta, dbias};
  const int rows_per_it = THREADS / (C / 8);
  if (rows_per_it < 1) return AIR_ERR_UNSUPPORTED;
  bn_bwd_reduce_kernel<<<grid_for(M, rows_per_it * 16, num_sms), THREADS, 2 * THREADS * 8 * sizeof(float), stream>>>(p);
  int agrid = grid_for(M * (C / 8), THREADS * 8, num_sms);
  if (dbias) {                         // every thread must stay on one 8-channel chunk: (grid * THREADS) % (C/8) == 0
    const int cpr = C / 8;
    int q = cpr;                       // smallest multiple of cpr / gcd(cpr, THREADS)
    { int a = cpr, b = THREADS; while (b) { int t = a % b; a = b; b = t; } q = cpr / a; }
    agrid = std::max(q, agrid / q * q);
  }
  bn_bwd_apply_kernel<<<agrid, THREADS, (7 * C + (dbias ? THREADS * 8 : 0)) * sizeof(float), stream>>>(p);
  return air_launch_status();
}
