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
template <typename T>
__global__ void __launch_bounds__(THREADS) bn_stats_kernel(const T* __restrict__ x, long long ld,
                                                            long long M, int C, double* __restrict__ sums) {
  extern __shared__ float sh[];                 // [2][THREADS][8]
  const int cpr = C >> 3;                       // 16-byte chunks per row
  const int rows_per_it = THREADS / cpr;        // C <= 2048
  const int tc = threadIdx.x % cpr, tr = threadIdx.x / cpr;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (tr < rows_per_it) {
    const long long step = static_cast<long long>(gridDim.x) * rows_per_it;
    long long m = static_cast<long long>(blockIdx.x) * rows_per_it + tr;
    for (; m + 3 * step < M; m += 4 * step) {             // four independent 16-byte loads in flight per thread
      V8<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldv8(x + (m + u * step) * ld + tc * 8);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    for (; m < M; m += step) {
      const V8<T> v = ldv8(x + m * ld + tc * 8);
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
template <typename T>
struct ApplyParams {
  const T* x; long long x_ld; T* y; long long y_ld; long long M; int C;
  const double* sums; const float* gamma; const float* beta; float eps; int relu; int training;
  float* save_mean; float* save_invstd; float* running_mean; float* running_var; float momentum;
  const T* add; long long add_ld; T* y2; long long y2_ld;   // optional: y2 = y + add
};

template <typename T>
__global__ void __launch_bounds__(THREADS) bn_apply_kernel(const ApplyParams<T> p) {
  // thread -> fixed 8-channel chunk tc (scale / shift live in registers), rows strided over the grid
  const int cpr = p.C >> 3;
  const int rows_per_it = THREADS / cpr;
  const int tc = threadIdx.x % cpr, tr = threadIdx.x / cpr;
  if (tr >= rows_per_it) return;
  float scale[8], shift[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = tc * 8 + k;
    float mean, invstd;
    if (p.training) {
      const double mu = p.sums[c] / static_cast<double>(p.M);
      double var = p.sums[p.C + c] / static_cast<double>(p.M) - mu * mu;
      if (var < 0.0) var = 0.0;
      mean = static_cast<float>(mu);
      invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
      if (blockIdx.x == 0 && tr == 0) {
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
    const float g = p.gamma ? p.gamma[c] : 1.f, be = p.beta ? p.beta[c] : 0.f;
    scale[k] = g * invstd;
    shift[k] = be - mean * g * invstd;
  }
  // Rows are walked from the END of the tensor: the kernel that produced x (a persistent conv grid, items in ascending
  // order) wrote its tail last, so that part is still in the 126 MB L2; and this kernel's own last writes are then the HEAD
  // of y, which the consumer (a conv walking its items in ascending order) reads first.
  const long long step = static_cast<long long>(gridDim.x) * rows_per_it;
  long long mi = static_cast<long long>(blockIdx.x) * rows_per_it + tr;      // mirrored row index: row = M - 1 - mi
  const long long last = p.M - 1;
  const int c0 = tc * 8;
  if (!p.y2) {
    for (; mi + 3 * step < p.M; mi += 4 * step) {          // four independent loads in flight before the first store
      V8<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldv8(p.x + (last - (mi + u * step)) * p.x_ld + c0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          f[k] = fmaf(f[k], scale[k], shift[k]);
          if (p.relu) f[k] = fmaxf(f[k], 0.f);
        }
        st8(p.y + (last - (mi + u * step)) * p.y_ld + c0, f);
      }
    }
  }
  for (; mi < p.M; mi += step) {
    const long long m = last - mi;
    float f[8];
    unpack8(ldv8(p.x + m * p.x_ld + c0), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = fmaf(f[k], scale[k], shift[k]);
      if (p.relu) f[k] = fmaxf(f[k], 0.f);
    }
    V8<T> yv;
    packv(f, yv);
    stv8(p.y + m * p.y_ld + c0, yv);
    if (p.y2) {                                   // Res2 branch input: sp_{i} + spx[i+1] (ecapa_tdnn.py:77-80)
      float a[8];
      unpack8(yv, f);                             // the consumer adds the ROUNDED branch output
      unpack8(ldv8(p.add + m * p.add_ld + c0), a);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += a[k];
      st8(p.y2 + m * p.y2_ld + c0, f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, pass 1: rsum[0..C) += sum g, rsum[C..2C) += sum g * xhat
//   order 0: g = dy * (gamma*xhat + beta > 0)        order 1: g = dy
// ---------------------------------------------------------------------------------------------
template <typename T>
struct BwdParams {
  const T* dy; long long dy_ld; const T* x; long long x_ld;
  const T* add; long long add_ld;          // optional extra gradient added to dx
  T* dx; long long dx_ld; long long M; int C; int order;
  const float* mean; const float* invstd; const float* gamma; const float* beta;
  double* rsum; float* dgamma; float* dbeta;
  float* dbias;                                        // optional: dbias[c] += sum_m dx[m][c] (bias of the producing conv)
};

template <typename T>
__global__ void __launch_bounds__(THREADS) bn_bwd_reduce_kernel(const BwdParams<T> p) {
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
    const long long step = static_cast<long long>(gridDim.x) * rows_per_it;
    long long m = static_cast<long long>(blockIdx.x) * rows_per_it + tr;
    constexpr int U = 4;                                   // 2 x U independent 16-byte loads in flight per thread
    while (m < p.M) {
      V8<T> gv[U], xq[U];
      int n = 0;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long mm = m + u * step;
        if (mm < p.M) {
          // rows from the END first: dy was just written by a data-gradient kernel whose tail is still in L2, and the
          // second pass (bn_bwd_apply, ascending) then finds the head of dy / x that this pass read last
          const long long row = p.M - 1 - mm;
          gv[u] = ldv8(p.dy + row * p.dy_ld + tc * 8);
          xq[u] = ldv8(p.x + row * p.x_ld + tc * 8);
          n = u + 1;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u < n) {
          float g[8], xv[8];
          unpack8(gv[u], g);
          unpack8(xq[u], xv);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xh = (xv[i] - mu[i]) * is[i];
            float gi = g[i];
            if (p.order == 0 && !(fmaf(ga[i], xh, be[i]) > 0.f)) gi = 0.f;
            s[i] += gi; q[i] = fmaf(gi, xh, q[i]);
          }
        }
      }
      m += U * step;
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
template <typename T>
__global__ void __launch_bounds__(THREADS) bn_bwd_apply_kernel(const BwdParams<T> p) {
  // thread -> fixed 8-channel chunk tc; the per-channel constants live in registers:
  //   xh = x * a + b2 (a = invstd, b2 = -mean * invstd),  dx = gi * k1 - (c2 + xh * c3),  c2 = k1 * sum_g / M, c3 = k1 * sum_gx / M
  extern __shared__ float sh[];                 // [THREADS][8] (dbias reduction only)
  const int cpr = p.C >> 3;
  const int rows_per_it = THREADS / cpr;
  const int tc = threadIdx.x % cpr, tr = threadIdx.x / cpr;
  const bool active = tr < rows_per_it;
  float a[8], b2[8], k1[8], c2[8], c3[8], ga[8], be[8], bs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = tc * 8 + k;
    const float gam = p.gamma ? p.gamma[c] : 1.f;
    const double sg = p.rsum[c], sgx = p.rsum[p.C + c];
    const float is = p.invstd[c], mu = p.mean[c];
    a[k] = is; b2[k] = -mu * is; k1[k] = gam * is;
    c2[k] = k1[k] * static_cast<float>(sg / static_cast<double>(p.M));
    c3[k] = k1[k] * static_cast<float>(sgx / static_cast<double>(p.M));
    ga[k] = gam; be[k] = p.beta ? p.beta[c] : 0.f; bs[k] = 0.f;
    if (blockIdx.x == 0 && tr == 0) {
      if (p.dgamma) p.dgamma[c] += static_cast<float>(sgx);
      if (p.dbeta) p.dbeta[c] += static_cast<float>(sg);
    }
  }
  if (active) {
    const long long step = static_cast<long long>(gridDim.x) * rows_per_it;
    const int c0 = tc * 8;
    for (long long m0 = static_cast<long long>(blockIdx.x) * rows_per_it + tr; m0 < p.M; m0 += 2 * step) {
      // two rows' loads (up to 6 x 16 bytes) are issued before the first store
      V8<T> gq[2], xq[2], aq[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long m = m0 + u * step;
        if (m < p.M) {
          gq[u] = ldv8(p.dy + m * p.dy_ld + c0);
          xq[u] = ldv8(p.x + m * p.x_ld + c0);
          if (p.add) aq[u] = ldv8(p.add + m * p.add_ld + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long m = m0 + u * step;
        if (m >= p.M) continue;
        float g[8], xv[8], ad[8];
        unpack8(gq[u], g);
        unpack8(xq[u], xv);
        if (p.add) unpack8(aq[u], ad);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = fmaf(xv[k], a[k], b2[k]);
          float gi = g[k];
          if (p.order == 0 && !(fmaf(ga[k], xh, be[k]) > 0.f)) gi = 0.f;
          float d = fmaf(gi, k1[k], -fmaf(xh, c3[k], c2[k]));
          if (p.order == 1 && !(xv[k] > 0.f)) d = 0.f;
          if (p.add) d += ad[k];
          g[k] = d;
          bs[k] += d;
        }
        st8(p.dx + m * p.dx_ld + c0, g);
      }
    }
  }
  if (p.dbias) {
    // block-level reduction first (one atomic per channel per CTA): thread t owns chunk t % cpr
#pragma unroll
    for (int k = 0; k < 8; ++k) sh[threadIdx.x * 8 + k] = active ? bs[k] : 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += THREADS) {
      const int chunk = c >> 3, e = c & 7;
      float acc = 0.f;
      for (int t = chunk; t < THREADS; t += cpr) acc += sh[t * 8 + e];
      atomicAdd(&p.dbias[c], acc);
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

template <typename T>
static int bn_stats_impl(const void* x, long long x_ld, long long M, int C, double* sums, int num_sms, cudaStream_t stream) {
  if (!x || !sums || !bn_args_ok(M, C, x_ld)) return AIR_ERR_ARG;
  const int rows_per_it = THREADS / (C / 8);
  if (rows_per_it < 1) return AIR_ERR_UNSUPPORTED;
  const int grid = grid_for(M, rows_per_it * 16, num_sms);
  bn_stats_kernel<T><<<grid, THREADS, 2 * THREADS * 8 * sizeof(float), stream>>>(reinterpret_cast<const T*>(x), x_ld, M, C, sums);
  return air_launch_status();
}

template <typename T>
static int bn_apply_impl(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                         const double* sums, const float* gamma, const float* beta, float eps, int relu,
                         int training, float* save_mean, float* save_invstd, float* running_mean,
                         float* running_var, float momentum, const void* add, long long add_ld, void* y2,
                         long long y2_ld, int num_sms, cudaStream_t stream) {
  if (!x || !y || !bn_args_ok(M, C, x_ld) || y_ld % 8 != 0) return AIR_ERR_ARG;
  if (training ? !sums : (!running_mean || !running_var)) return AIR_ERR_ARG;
  if (y2 && (!add || add_ld % 8 != 0 || y2_ld % 8 != 0)) return AIR_ERR_ARG;
  ApplyParams<T> p{reinterpret_cast<const T*>(x), x_ld, reinterpret_cast<T*>(y), y_ld, M, C,
                   sums, gamma, beta, eps, relu, training, save_mean, save_invstd, running_mean, running_var, momentum,
                   reinterpret_cast<const T*>(add), add_ld, reinterpret_cast<T*>(y2), y2_ld};
  const int rows_per_it = THREADS / (C / 8);
  if (rows_per_it < 1) return AIR_ERR_UNSUPPORTED;
  const int grid = grid_for(M, rows_per_it * 8, num_sms);
  bn_apply_kernel<T><<<grid, THREADS, 0, stream>>>(p);
  return air_launch_status();
}

template <typename T>
static int bn_bwd_impl(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                       void* dx, long long dx_ld, long long M, int C, int order,
                       const float* mean, const float* invstd, const float* gamma, const float* beta,
                       double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, cudaStream_t stream) {
  if (!dy || !x || !dx || !mean || !invstd || !rsum || !bn_args_ok(M, C, x_ld)) return AIR_ERR_ARG;
  if (dy_ld % 8 != 0 || dx_ld % 8 != 0 || (add && add_ld % 8 != 0)) return AIR_ERR_ARG;
  BwdParams<T> p{reinterpret_cast<const T*>(dy), dy_ld, reinterpret_cast<const T*>(x), x_ld,
                 reinterpret_cast<const T*>(add), add_ld, reinterpret_cast<T*>(dx), dx_ld, M, C, order,
                 mean, invstd, gamma, beta, rsum, dgamma, dbeta, dbias};
  const int rows_per_it = THREADS / (C / 8);
  if (rows_per_it < 1) return AIR_ERR_UNSUPPORTED;
  bn_bwd_reduce_kernel<T><<<grid_for(M, rows_per_it * 16, num_sms), THREADS, 2 * THREADS * 8 * sizeof(float), stream>>>(p);
  const int agrid = grid_for(M, rows_per_it * 8, num_sms);
  bn_bwd_apply_kernel<T><<<agrid, THREADS, (dbias ? THREADS * 8 : 0) * sizeof(float), stream>>>(p);
  return air_launch_status();
}

extern "C" int air_bn_stats(const void* x, long long x_ld, long long M, int C, double* sums, int num_sms, cudaStream_t stream) {
  return bn_stats_impl<__nv_bfloat16>(x, x_ld, M, C, sums, num_sms, stream);
}
extern "C" int air_bn_stats_f32(const void* x, long long x_ld, long long M, int C, double* sums, int num_sms, cudaStream_t stream) {
  return bn_stats_impl<float>(x, x_ld, M, C, sums, num_sms, stream);
}

extern "C" int air_bn_apply_add(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                                const double* sums, const float* gamma, const float* beta, float eps, int relu,
                                int training, float* save_mean, float* save_invstd, float* running_mean,
                                float* running_var, float momentum, const void* add, long long add_ld, void* y2,
                                long long y2_ld, int num_sms, cudaStream_t stream) {
  return bn_apply_impl<__nv_bfloat16>(x, x_ld, y, y_ld, M, C, sums, gamma, beta, eps, relu, training, save_mean, save_invstd,
                                      running_mean, running_var, momentum, add, add_ld, y2, y2_ld, num_sms, stream);
}
extern "C" int air_bn_apply_add_f32(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                                    const double* sums, const float* gamma, const float* beta, float eps, int relu,
                                    int training, float* save_mean, float* save_invstd, float* running_mean,
                                    float* running_var, float momentum, const void* add, long long add_ld, void* y2,
                                    long long y2_ld, int num_sms, cudaStream_t stream) {
  return bn_apply_impl<float>(x, x_ld, y, y_ld, M, C, sums, gamma, beta, eps, relu, training, save_mean, save_invstd,
                              running_mean, running_var, momentum, add, add_ld, y2, y2_ld, num_sms, stream);
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
                               double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, cudaStream_t stream) {
  return bn_bwd_impl<__nv_bfloat16>(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum,
                                    dgamma, dbeta, dbias, num_sms, stream);
}
extern "C" int air_bn_bwd_bias_f32(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                                   void* dx, long long dx_ld, long long M, int C, int order,
                                   const float* mean, const float* invstd, const float* gamma, const float* beta,
                                   double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, cudaStream_t stream) {
  return bn_bwd_impl<float>(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum,
                            dgamma, dbeta, dbias, num_sms, stream);
}

extern "C" int air_bn_bwd(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                          void* dx, long long dx_ld, long long M, int C, int order,
                          const float* mean, const float* invstd, const float* gamma, const float* beta,
                          double* rsum, float* dgamma, float* dbeta, int num_sms, cudaStream_t stream) {
  return air_bn_bwd_bias(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum, dgamma,
                         dbeta, nullptr, num_sms, stream);
}
