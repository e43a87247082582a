// ECAPA-TDNN (Res2Net2) non-GEMM stages on channels-last bf16 activations (B, T, C):
//   SE block (ecapa_tdnn.py:15-29), residual add (:93), context statistics and attentive
//   statistics pooling (:168-186), the small fp32 BatchNorm1d layers that act on (B, C) rows
//   (SE bottleneck BN, bn5, bn7: :22,147,150,188,195) and helpers (channel-slice copy, column sums).
// All kernels are HBM-bound: 16-byte loads of 8 channels per thread, time reductions through
// registers -> shared memory, fp32 math.
#include <algorithm>
#include "common.cuh"

namespace air_ecapa {

constexpr int TG = 8;          // time groups per CTA
constexpr int CT = 32;         // chunk threads per CTA (32 x 8 = 256 channels)
constexpr int THREADS = TG * CT;

// ---------------------------------------------------------------------------------------------
// mean_t and (optionally) sqrt(clamp(unbiased var_t, clampv)) over time.   grid (ceil(C/256), B)
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(THREADS) time_stats_kernel(const AT* __restrict__ x, long long ld, int T, int C,
                                                              float* __restrict__ mean_out, float* __restrict__ std_out, float clampv) {
  __shared__ float sh[2][TG][CT * 8 + 8];
  const int b = blockIdx.y, ct = threadIdx.x % CT, tg = threadIdx.x / CT;
  const int c0 = (blockIdx.x * CT + ct) * 8;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  if (c0 < C) {
    const AT* p = x + static_cast<long long>(b) * T * ld + c0;
    for (int t = tg; t < T; t += TG) {
      float f[8];
      ld8(p + t * ld, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { sh[0][tg][ct * 8 + i] = s[i]; sh[1][tg][ct * 8 + i] = q[i]; }
  __syncthreads();
  const int c = blockIdx.x * CT * 8 + threadIdx.x;
  if (threadIdx.x < CT * 8 && c < C) {
    float a = 0.f, d = 0.f;
#pragma unroll
    for (int g = 0; g < TG; ++g) { a += sh[0][g][threadIdx.x]; d += sh[1][g][threadIdx.x]; }
    const float mean = a / T;
    mean_out[static_cast<long long>(b) * C + c] = clampv < 0.f ? a : mean;     // clampv < 0: plain sum over time
    if (std_out) {
      const float var = (d - a * mean) / (T - 1);           // torch.var default: unbiased
      std_out[static_cast<long long>(b) * C + c] = sqrtf(fmaxf(var, clampv));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Attentive statistics pooling forward (ecapa_tdnn.py:182-186):
//   w = softmax_t(e);  mu = sum_t x w;  sg = sqrt(clamp(sum_t x^2 w - mu^2, 1e-4))
// e, x: (B, T, C) bf16.  out (B, 2C) = [mu | sg]; saved per (b, c): max_t e, sum_t exp(e - max), q = sum x^2 w.
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(THREADS) asp_fwd_kernel(const AT* __restrict__ e, long long e_ld,
                                                           const AT* __restrict__ x, long long x_ld, int T, int C,
                                                           float* __restrict__ out, float* __restrict__ smax, float* __restrict__ ssum,
                                                           float* __restrict__ sq) {
  __shared__ float sh[4][TG][CT * 8 + 8];
  const int b = blockIdx.y, ct = threadIdx.x % CT, tg = threadIdx.x / CT;
  const int c0 = (blockIdx.x * CT + ct) * 8;
  float m[8], s[8], sx[8], sxx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = -INFINITY; s[i] = 0.f; sx[i] = 0.f; sxx[i] = 0.f; }
  if (c0 < C) {
    const AT* pe = e + static_cast<long long>(b) * T * e_ld + c0;
    const AT* px = x + static_cast<long long>(b) * T * x_ld + c0;
    for (int t = tg; t < T; t += TG) {
      float fe[8], fx[8];
      ld8(pe + t * e_ld, fe);
      ld8(px + t * x_ld, fx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mn = fmaxf(m[i], fe[i]);
        const float sc = __expf(m[i] - mn), w = __expf(fe[i] - mn);
        s[i] = s[i] * sc + w;
        sx[i] = sx[i] * sc + w * fx[i];
        sxx[i] = sxx[i] * sc + w * fx[i] * fx[i];
        m[i] = mn;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sh[0][tg][ct * 8 + i] = m[i]; sh[1][tg][ct * 8 + i] = s[i]; sh[2][tg][ct * 8 + i] = sx[i]; sh[3][tg][ct * 8 + i] = sxx[i];
  }
  __syncthreads();
  const int c = blockIdx.x * CT * 8 + threadIdx.x;
  if (threadIdx.x < CT * 8 && c < C) {
    float mm = -INFINITY;
#pragma unroll
    for (int g = 0; g < TG; ++g) mm = fmaxf(mm, sh[0][g][threadIdx.x]);
    float a = 0.f, ax = 0.f, axx = 0.f;
#pragma unroll
    for (int g = 0; g < TG; ++g) {
      const float sc = __expf(sh[0][g][threadIdx.x] - mm);
      a += sh[1][g][threadIdx.x] * sc; ax += sh[2][g][threadIdx.x] * sc; axx += sh[3][g][threadIdx.x] * sc;
    }
    const float mu = ax / a, q = axx / a;
    const long long o = static_cast<long long>(b) * C + c;
    out[static_cast<long long>(b) * 2 * C + c] = mu;
    out[static_cast<long long>(b) * 2 * C + C + c] = sqrtf(fmaxf(q - mu * mu, 1e-4f));
    smax[o] = mm; ssum[o] = a; sq[o] = q;
  }
}

// ---------------------------------------------------------------------------------------------
// Pooling backward, elementwise over (B, T, C) given per-(b, c) constants:
//   de = p (dp - <p, dp>),  dp = x dmu' + x^2 dq,  <p, dp> = mu dmu' + q dq
//   dx = p dmu' + 2 x p dq                                   (attentive statistics, direct path)
//      + dmean / T + dstd (x - mean) / ((T-1) std) [var > clamp]  (context statistics, :169-172)
// with dq = dsg / (2 sg) [q - mu^2 > 1e-4], dmu' = dmu - 2 mu dq.
// ---------------------------------------------------------------------------------------------
template <typename AT>
struct AspBwd {
  const AT* e; long long e_ld; const AT* x; long long x_ld;
  const float* out; const float* dout; const float* smax; const float* ssum; const float* sq;
  const float* cmean; const float* cstd; const float* dcmean; const float* dcstd; float clampv;
  AT* de; long long de_ld; AT* dx; long long dx_ld; int T, C;
};

template <typename AT>
__global__ void __launch_bounds__(THREADS) asp_bwd_kernel(const AspBwd<AT> p) {
  const int b = blockIdx.y, ct = threadIdx.x % CT, tg = threadIdx.x / CT;
  const int c0 = (blockIdx.x * CT + ct) * 8;
  if (c0 >= p.C) return;
  float mx[8], rs[8], dmu[8], dq[8], dot[8], k0[8], k1[8], cm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long o = static_cast<long long>(b) * p.C + c0 + i;
    const float mu = p.out[static_cast<long long>(b) * 2 * p.C + c0 + i];
    const float sg = p.out[static_cast<long long>(b) * 2 * p.C + p.C + c0 + i];
    const float q = p.sq[o];
    const float g_mu = p.dout[static_cast<long long>(b) * 2 * p.C + c0 + i];
    const float g_sg = p.dout[static_cast<long long>(b) * 2 * p.C + p.C + c0 + i];
    dq[i] = (q - mu * mu > 1e-4f) ? g_sg / (2.f * sg) : 0.f;
    dmu[i] = g_mu - 2.f * mu * dq[i];
    dot[i] = mu * dmu[i] + q * dq[i];
    mx[i] = p.smax[o]; rs[i] = 1.f / p.ssum[o];
    const float sd = p.cstd[o];
    cm[i] = p.cmean[o];
    k0[i] = p.dcmean[o] / p.T;
    k1[i] = (sd * sd > p.clampv * 1.000001f) ? p.dcstd[o] / ((p.T - 1) * sd) : 0.f;
  }
  const long long base = static_cast<long long>(b) * p.T;
  for (int t = tg; t < p.T; t += TG) {
    float fe[8], fx[8], oe[8], ox[8];
    ld8(p.e + (base + t) * p.e_ld + c0, fe);
    ld8(p.x + (base + t) * p.x_ld + c0, fx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float pw = __expf(fe[i] - mx[i]) * rs[i];
      const float dp = fx[i] * dmu[i] + fx[i] * fx[i] * dq[i];
      oe[i] = pw * (dp - dot[i]);
      ox[i] = pw * dmu[i] + 2.f * fx[i] * pw * dq[i] + k0[i] + k1[i] * (fx[i] - cm[i]);
    }
    st8(p.de + (base + t) * p.de_ld + c0, oe);
    st8(p.dx + (base + t) * p.dx_ld + c0, ox);
  }
}

// dx = (dx + dmean / T + dstd (x - mean) / ((T-1) std) [var > clamp]) * (x > 0):
// context-statistics backward (ecapa_tdnn.py:169-172) fused with the ReLU mask of layer4 (:166).
template <typename AT>
__global__ void __launch_bounds__(THREADS) ctx_bwd_mask_kernel(const AT* __restrict__ x, long long x_ld,
                                                                const float* __restrict__ cmean, const float* __restrict__ cstd,
                                                                const float* __restrict__ dcmean, const float* __restrict__ dcstd,
                                                                float clampv, AT* __restrict__ dx, long long dx_ld,
                                                                int T, int C) {
  const int b = blockIdx.y, ct = threadIdx.x % CT, tg = threadIdx.x / CT;
  const int c0 = (blockIdx.x * CT + ct) * 8;
  if (c0 >= C) return;
  float k0[8], k1[8], cm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long o = static_cast<long long>(b) * C + c0 + i;
    const float sd = cstd[o];
    cm[i] = cmean[o];
    k0[i] = dcmean[o] / T;
    k1[i] = (sd * sd > clampv * 1.000001f) ? dcstd[o] / ((T - 1) * sd) : 0.f;
  }
  const long long base = static_cast<long long>(b) * T;
  for (int t = tg; t < T; t += TG) {
    float fx[8], fd[8];
    ld8(x + (base + t) * x_ld + c0, fx);
    ld8(dx + (base + t) * dx_ld + c0, fd);
#pragma unroll
    for (int i = 0; i < 8; ++i) fd[i] = fx[i] > 0.f ? fd[i] + k0[i] + k1[i] * (fx[i] - cm[i]) : 0.f;
    st8(dx + (base + t) * dx_ld + c0, fd);
  }
}

// ---------------------------------------------------------------------------------------------
// SE gate application + residual:  out = x * g[b, c] + res        (ecapa_tdnn.py:27-28, :93)
// backward pieces:  dg[b, c] = sum_t dout * x;   dx = dout * g + ds[b, c] / T  [* (mask > 0)]
// ---------------------------------------------------------------------------------------------
template <typename AT>
__global__ void __launch_bounds__(256) scale_residual_kernel(const AT* __restrict__ x, long long x_ld,
                                                              const float* __restrict__ g, const AT* __restrict__ res,
                                                              long long res_ld, AT* __restrict__ out, long long out_ld,
                                                              long long M, int T, int C) {
  // 32-bit index arithmetic (the launcher checks M * C / 8 < 2^31): the two 64-bit divisions per 16-byte chunk of the first
  // version cost more instructions than the load, the FMA and the store together
  const uint32_t cpr = static_cast<uint32_t>(C) >> 3;
  const uint32_t total = static_cast<uint32_t>(M) * cpr;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t mi = i / cpr;
    const int c0 = static_cast<int>(i - mi * cpr) * 8;
    const long long m = mi, b = mi / static_cast<uint32_t>(T);
    float f[8], r[8];
    ld8(x + m * x_ld + c0, f);
    if (res) ld8(res + m * res_ld + c0, r);
    const float4 g0 = *reinterpret_cast<const float4*>(g + b * C + c0), g1 = *reinterpret_cast<const float4*>(g + b * C + c0 + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], gg[k], res ? r[k] : 0.f);
    st8(out + m * out_ld + c0, f);
  }
}

template <typename AT>
__global__ void __launch_bounds__(THREADS) se_dgate_kernel(const AT* __restrict__ dout, long long d_ld,
                                                            const AT* __restrict__ x, long long x_ld, int T, int C,
                                                            float* __restrict__ dg) {
  __shared__ float sh[TG][CT * 8 + 8];
  const int b = blockIdx.y, ct = threadIdx.x % CT, tg = threadIdx.x / CT;
  const int c0 = (blockIdx.x * CT + ct) * 8;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (c0 < C) {
    const long long base = static_cast<long long>(b) * T;
    for (int t = tg; t < T; t += TG) {
      float fd[8], fx[8];
      ld8(dout + (base + t) * d_ld + c0, fd);
      ld8(x + (base + t) * x_ld + c0, fx);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = fmaf(fd[i], fx[i], s[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[tg][ct * 8 + i] = s[i];
  __syncthreads();
  const int c = blockIdx.x * CT * 8 + threadIdx.x;
  if (threadIdx.x < CT * 8 && c < C) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < TG; ++g) a += sh[g][threadIdx.x];
    dg[static_cast<long long>(b) * C + c] = a;
  }
}

template <typename AT>
__global__ void __launch_bounds__(256) se_apply_bwd_kernel(const AT* __restrict__ dout, long long d_ld,
                                                            const float* __restrict__ g, const float* __restrict__ ds,
                                                            AT* __restrict__ dx, long long dx_ld, long long M, int T, int C) {
  const uint32_t cpr = static_cast<uint32_t>(C) >> 3;            // 32-bit index arithmetic, see scale_residual_kernel
  const uint32_t total = static_cast<uint32_t>(M) * cpr;
  const float invT = 1.f / T;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t mi = i / cpr;
    const int c0 = static_cast<int>(i - mi * cpr) * 8;
    const long long m = mi, b = mi / static_cast<uint32_t>(T);
    float f[8];
    ld8(dout + m * d_ld + c0, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], g[b * C + c0 + k], ds[b * C + c0 + k] * invT);
    st8(dx + m * dx_ld + c0, f);
  }
}

// ---------------------------------------------------------------------------------------------
// fp32 BatchNorm1d over (M, C) rows, M = batch (SE bottleneck BN(128), bn5(3072), bn7(2)).
// relu_in != 0: the input is relu'd first (Conv -> ReLU -> BN order of ecapa_tdnn.py:19-22).
// One thread per channel (M <= a few thousand rows), rows added in order (the summation order is part of the parity
// contract at B = 4); the row loops are unrolled so that eight rows' loads are in flight.  The training forms remain
// latency chains over M (0.02 ms forward, 0.07 ms backward at B = 256: 2 CTAs, HBM-latency batches) -- staging the
// (M, 64) column block in shared memory first is the next step.
// ---------------------------------------------------------------------------------------------
__global__ void bn1d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int M, int C, int relu_in,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int training,
                                float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                float* __restrict__ running_mean, float* __restrict__ running_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    float s = 0.f;
#pragma unroll 8
    for (int m = 0; m < M; ++m) { float v = x[static_cast<long long>(m) * C + c]; if (relu_in) v = fmaxf(v, 0.f); s += v; }
    mean = s / M;
    float q = 0.f;
#pragma unroll 8
    for (int m = 0; m < M; ++m) { float v = x[static_cast<long long>(m) * C + c]; if (relu_in) v = fmaxf(v, 0.f); v -= mean; q = fmaf(v, v, q); }
    const float var = q / M;
    invstd = rsqrtf(var + eps);
    if (save_mean) { save_mean[c] = mean; save_invstd[c] = invstd; }
    if (running_mean) {
      const float unb = M > 1 ? q / (M - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  } else {
    mean = running_mean[c];
    invstd = rsqrtf(running_var[c] + eps);
  }
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
#pragma unroll 8
  for (int m = 0; m < M; ++m) {
    float v = x[static_cast<long long>(m) * C + c];
    if (relu_in) v = fmaxf(v, 0.f);
    y[static_cast<long long>(m) * C + c] = (v - mean) * invstd * g + b;
  }
}

// Eval mode: the running-statistics affine has no reduction, so it runs one thread per element (the per-channel kernel above
// walks the batch serially: 0.09 - 0.2 ms per call at the scoring batch of 1024).  Same expression, same results.
__global__ void bn1d_eval_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int C, int relu_in,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 const float* __restrict__ running_mean, const float* __restrict__ running_var) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const float mean = running_mean[c], invstd = rsqrtf(running_var[c] + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float v = x[i];
    if (relu_in) v = fmaxf(v, 0.f);
    y[i] = (v - mean) * invstd * g + b;
  }
}

__global__ void bn1d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int M, int C,
                                int relu_in, const float* __restrict__ gamma, const float* __restrict__ mean,
                                const float* __restrict__ invstd, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = mean[c], is = invstd[c], g = gamma ? gamma[c] : 1.f;
  float sg = 0.f, sgx = 0.f;
#pragma unroll 8
  for (int m = 0; m < M; ++m) {
    float v = x[static_cast<long long>(m) * C + c]; if (relu_in) v = fmaxf(v, 0.f);
    const float d = dy[static_cast<long long>(m) * C + c];
    sg += d; sgx = fmaf(d, (v - mu) * is, sgx);
  }
  if (dgamma) dgamma[c] += sgx;
  if (dbeta) dbeta[c] += sg;
  const float k2 = sg / M, k3 = sgx / M;
#pragma unroll 8
  for (int m = 0; m < M; ++m) {
    const float raw = x[static_cast<long long>(m) * C + c];
    const float v = relu_in ? fmaxf(raw, 0.f) : raw;
    float d = g * is * (dy[static_cast<long long>(m) * C + c] - k2 - (v - mu) * is * k3);
    if (relu_in && !(raw > 0.f)) d = 0.f;
    dx[static_cast<long long>(m) * C + c] = d;
  }
}

__global__ void sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = 1.f / (1.f + __expf(-x[i]));
}
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
    dx[i] = dy[i] * y[i] * (1.f - y[i]);
}

// channel-slice copy / masked copy: dst[m][c] = src[m][c] * (mask[m][c] > 0 if mask)
template <typename AT>
__global__ void __launch_bounds__(256) copy_channels_kernel(const AT* __restrict__ src, long long s_ld,
                                                             const AT* __restrict__ mask, long long m_ld,
                                                             AT* __restrict__ dst, long long d_ld, long long M, int C) {
  const uint32_t cpr = static_cast<uint32_t>(C) >> 3;            // 32-bit index arithmetic, see scale_residual_kernel
  const uint32_t total = static_cast<uint32_t>(M) * cpr;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t mi = i / cpr;
    const int c0 = static_cast<int>(i - mi * cpr) * 8;
    const long long m = mi;
    V8<AT> v = ldv8(src + m * s_ld + c0);
    if (mask) {
      float f[8], k[8];
      unpack8(v, f);
      ld8(mask + m * m_ld + c0, k);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = k[j] > 0.f ? f[j] : 0.f;
      packv(f, v);
    }
    stv8(dst + m * d_ld + c0, v);
  }
}

// out[c] += sum_m x[m][c]   (conv bias gradients)
template <typename AT>
__global__ void __launch_bounds__(256) colsum_kernel(const AT* __restrict__ x, long long ld, long long M, int C,
                                                      float* __restrict__ out) {
  extern __shared__ float sh[];                 // [256][8]
  const int cpr = C >> 3;
  const int rows_per_it = 256 / cpr;
  const int tc = threadIdx.x % cpr, tr = threadIdx.x / cpr;
  float s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  if (tr < rows_per_it) {
    for (long long m = static_cast<long long>(blockIdx.x) * rows_per_it + tr; m < M; m += static_cast<long long>(gridDim.x) * rows_per_it) {
      float f[8];
      ld8(x + m * ld + tc * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[threadIdx.x * 8 + i] = s[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int chunk = c >> 3, e = c & 7;
    float a = 0.f;
    for (int r = 0; r < rows_per_it; ++r) a += sh[(r * cpr + chunk) * 8 + e];
    atomicAdd(&out[c], a);
  }
}

static int ew_grid(long long items, int per_block) {
  long long b = (items + per_block - 1) / per_block;
  return static_cast<int>(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace air_ecapa

using namespace air_ecapa;

template <typename AT>
static int time_stats_fwd_impl(const void* x, long long x_ld, int B, int T, int C, float* mean_out, float* std_out,
                                  float clampv, cudaStream_t stream) {
  if (!x || !mean_out || B <= 0 || T < 2 || C % 8 != 0 || x_ld % 8 != 0) return AIR_ERR_ARG;
  dim3 grid((C + CT * 8 - 1) / (CT * 8), B);
  time_stats_kernel<AT><<<grid, THREADS, 0, stream>>>(reinterpret_cast<const AT*>(x), x_ld, T, C, mean_out, std_out, clampv);
  return air_launch_status();
}
extern "C" int air_time_stats_fwd(const void* x, long long x_ld, int B, int T, int C, float* mean_out, float* std_out,
                                  float clampv, cudaStream_t stream) {
  return time_stats_fwd_impl<__nv_bfloat16>(x, x_ld, B, T, C, mean_out, std_out, clampv, stream);
}
extern "C" int air_time_stats_fwd_f32(const void* x, long long x_ld, int B, int T, int C, float* mean_out, float* std_out,
                                  float clampv, cudaStream_t stream) {
  return time_stats_fwd_impl<float>(x, x_ld, B, T, C, mean_out, std_out, clampv, stream);
}

template <typename AT>
static int ecapa_asp_fwd_impl(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                                 float* out, float* save_max, float* save_sum, float* save_q, cudaStream_t stream) {
  if (!e || !x || !out || !save_max || !save_sum || !save_q || B <= 0 || T < 1 || C % 8 != 0 || e_ld % 8 != 0 || x_ld % 8 != 0)
    return AIR_ERR_ARG;
  dim3 grid((C + CT * 8 - 1) / (CT * 8), B);
  asp_fwd_kernel<AT><<<grid, THREADS, 0, stream>>>(reinterpret_cast<const AT*>(e), e_ld,
                                               reinterpret_cast<const AT*>(x), x_ld, T, C, out, save_max, save_sum, save_q);
  return air_launch_status();
}
extern "C" int air_ecapa_asp_fwd(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                                 float* out, float* save_max, float* save_sum, float* save_q, cudaStream_t stream) {
  return ecapa_asp_fwd_impl<__nv_bfloat16>(e, e_ld, x, x_ld, B, T, C, out, save_max, save_sum, save_q, stream);
}
extern "C" int air_ecapa_asp_fwd_f32(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                                 float* out, float* save_max, float* save_sum, float* save_q, cudaStream_t stream) {
  return ecapa_asp_fwd_impl<float>(e, e_ld, x, x_ld, B, T, C, out, save_max, save_sum, save_q, stream);
}

template <typename AT>
static int ecapa_asp_bwd_impl(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                                 const float* out, const float* dout, const float* save_max, const float* save_sum,
                                 const float* save_q, const float* ctx_mean, const float* ctx_std, const float* dctx_mean,
                                 const float* dctx_std, float clampv, void* de, long long de_ld, void* dx, long long dx_ld,
                                 cudaStream_t stream) {
  if (!e || !x || !out || !dout || !save_max || !save_sum || !save_q || !ctx_mean || !ctx_std || !dctx_mean || !dctx_std || !de || !dx)
    return AIR_ERR_ARG;
  if (B <= 0 || T < 2 || C % 8 != 0 || (e_ld | x_ld | de_ld | dx_ld) % 8 != 0) return AIR_ERR_ARG;
  AspBwd<AT> p{reinterpret_cast<const AT*>(e), e_ld, reinterpret_cast<const AT*>(x), x_ld, out, dout,
           save_max, save_sum, save_q, ctx_mean, ctx_std, dctx_mean, dctx_std, clampv,
           reinterpret_cast<AT*>(de), de_ld, reinterpret_cast<AT*>(dx), dx_ld, T, C};
  dim3 grid((C + CT * 8 - 1) / (CT * 8), B);
  asp_bwd_kernel<AT><<<grid, THREADS, 0, stream>>>(p);
  return air_launch_status();
}
extern "C" int air_ecapa_asp_bwd(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                                 const float* out, const float* dout, const float* save_max, const float* save_sum,
                                 const float* save_q, const float* ctx_mean, const float* ctx_std, const float* dctx_mean,
                                 const float* dctx_std, float clampv, void* de, long long de_ld, void* dx, long long dx_ld,
                                 cudaStream_t stream) {
  return ecapa_asp_bwd_impl<__nv_bfloat16>(e, e_ld, x, x_ld, B, T, C, out, dout, save_max, save_sum, save_q, ctx_mean, ctx_std, dctx_mean, dctx_std, clampv, de, de_ld, dx, dx_ld, stream);
}
extern "C" int air_ecapa_asp_bwd_f32(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                                 const float* out, const float* dout, const float* save_max, const float* save_sum,
                                 const float* save_q, const float* ctx_mean, const float* ctx_std, const float* dctx_mean,
                                 const float* dctx_std, float clampv, void* de, long long de_ld, void* dx, long long dx_ld,
                                 cudaStream_t stream) {
  return ecapa_asp_bwd_impl<float>(e, e_ld, x, x_ld, B, T, C, out, dout, save_max, save_sum, save_q, ctx_mean, ctx_std, dctx_mean, dctx_std, clampv, de, de_ld, dx, dx_ld, stream);
}

template <typename AT>
static int ctx_stats_bwd_mask_impl(const void* x, long long x_ld, int B, int T, int C, const float* ctx_mean,
                                      const float* ctx_std, const float* dctx_mean, const float* dctx_std, float clampv,
                                      void* dx, long long dx_ld, cudaStream_t stream) {
  if (!x || !ctx_mean || !ctx_std || !dctx_mean || !dctx_std || !dx || B <= 0 || T < 2 || C % 8 != 0 || (x_ld | dx_ld) % 8 != 0)
    return AIR_ERR_ARG;
  dim3 grid((C + CT * 8 - 1) / (CT * 8), B);
  ctx_bwd_mask_kernel<AT><<<grid, THREADS, 0, stream>>>(reinterpret_cast<const AT*>(x), x_ld, ctx_mean, ctx_std,
                                                    dctx_mean, dctx_std, clampv, reinterpret_cast<AT*>(dx), dx_ld, T, C);
  return air_launch_status();
}
extern "C" int air_ctx_stats_bwd_mask(const void* x, long long x_ld, int B, int T, int C, const float* ctx_mean,
                                      const float* ctx_std, const float* dctx_mean, const float* dctx_std, float clampv,
                                      void* dx, long long dx_ld, cudaStream_t stream) {
  return ctx_stats_bwd_mask_impl<__nv_bfloat16>(x, x_ld, B, T, C, ctx_mean, ctx_std, dctx_mean, dctx_std, clampv, dx, dx_ld, stream);
}
extern "C" int air_ctx_stats_bwd_mask_f32(const void* x, long long x_ld, int B, int T, int C, const float* ctx_mean,
                                      const float* ctx_std, const float* dctx_mean, const float* dctx_std, float clampv,
                                      void* dx, long long dx_ld, cudaStream_t stream) {
  return ctx_stats_bwd_mask_impl<float>(x, x_ld, B, T, C, ctx_mean, ctx_std, dctx_mean, dctx_std, clampv, dx, dx_ld, stream);
}

template <typename AT>
static int scale_residual_fwd_impl(const void* x, long long x_ld, const float* gate, const void* res, long long res_ld,
                                      void* out, long long out_ld, int B, int T, int C, cudaStream_t stream) {
  if (!x || !gate || !out || B <= 0 || T <= 0 || C % 8 != 0 || (x_ld | out_ld) % 8 != 0 || (res && res_ld % 8 != 0)) return AIR_ERR_ARG;
  const long long M = static_cast<long long>(B) * T;
  if (M * (C / 8) >= 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;            // 32-bit chunk index in the kernel
  scale_residual_kernel<AT><<<ew_grid(M * (C / 8), 256 * 4), 256, 0, stream>>>(
      reinterpret_cast<const AT*>(x), x_ld, gate, reinterpret_cast<const AT*>(res), res_ld,
      reinterpret_cast<AT*>(out), out_ld, M, T, C);
  return air_launch_status();
}
extern "C" int air_scale_residual_fwd(const void* x, long long x_ld, const float* gate, const void* res, long long res_ld,
                                      void* out, long long out_ld, int B, int T, int C, cudaStream_t stream) {
  return scale_residual_fwd_impl<__nv_bfloat16>(x, x_ld, gate, res, res_ld, out, out_ld, B, T, C, stream);
}
extern "C" int air_scale_residual_fwd_f32(const void* x, long long x_ld, const float* gate, const void* res, long long res_ld,
                                      void* out, long long out_ld, int B, int T, int C, cudaStream_t stream) {
  return scale_residual_fwd_impl<float>(x, x_ld, gate, res, res_ld, out, out_ld, B, T, C, stream);
}

template <typename AT>
static int se_dgate_impl(const void* dout, long long d_ld, const void* x, long long x_ld, int B, int T, int C, float* dgate,
                            cudaStream_t stream) {
  if (!dout || !x || !dgate || B <= 0 || T <= 0 || C % 8 != 0 || (d_ld | x_ld) % 8 != 0) return AIR_ERR_ARG;
  dim3 grid((C + CT * 8 - 1) / (CT * 8), B);
  se_dgate_kernel<AT><<<grid, THREADS, 0, stream>>>(reinterpret_cast<const AT*>(dout), d_ld,
                                                reinterpret_cast<const AT*>(x), x_ld, T, C, dgate);
  return air_launch_status();
}
extern "C" int air_se_dgate(const void* dout, long long d_ld, const void* x, long long x_ld, int B, int T, int C, float* dgate,
                            cudaStream_t stream) {
  return se_dgate_impl<__nv_bfloat16>(dout, d_ld, x, x_ld, B, T, C, dgate, stream);
}
extern "C" int air_se_dgate_f32(const void* dout, long long d_ld, const void* x, long long x_ld, int B, int T, int C, float* dgate,
                            cudaStream_t stream) {
  return se_dgate_impl<float>(dout, d_ld, x, x_ld, B, T, C, dgate, stream);
}

template <typename AT>
static int se_apply_bwd_impl(const void* dout, long long d_ld, const float* gate, const float* dmean, void* dx, long long dx_ld,
                                int B, int T, int C, cudaStream_t stream) {
  if (!dout || !gate || !dmean || !dx || B <= 0 || T <= 0 || C % 8 != 0 || (d_ld | dx_ld) % 8 != 0) return AIR_ERR_ARG;
  const long long M = static_cast<long long>(B) * T;
  if (M * (C / 8) >= 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;            // 32-bit chunk index in the kernel
  se_apply_bwd_kernel<AT><<<ew_grid(M * (C / 8), 256 * 4), 256, 0, stream>>>(
      reinterpret_cast<const AT*>(dout), d_ld, gate, dmean, reinterpret_cast<AT*>(dx), dx_ld, M, T, C);
  return air_launch_status();
}
extern "C" int air_se_apply_bwd(const void* dout, long long d_ld, const float* gate, const float* dmean, void* dx, long long dx_ld,
                                int B, int T, int C, cudaStream_t stream) {
  return se_apply_bwd_impl<__nv_bfloat16>(dout, d_ld, gate, dmean, dx, dx_ld, B, T, C, stream);
}
extern "C" int air_se_apply_bwd_f32(const void* dout, long long d_ld, const float* gate, const float* dmean, void* dx, long long dx_ld,
                                int B, int T, int C, cudaStream_t stream) {
  return se_apply_bwd_impl<float>(dout, d_ld, gate, dmean, dx, dx_ld, B, T, C, stream);
}

extern "C" int air_bn1d_f32_fwd(const float* x, float* y, int M, int C, int relu_in, const float* gamma, const float* beta,
                                float eps, int training, float* save_mean, float* save_invstd, float* running_mean,
                                float* running_var, float momentum, cudaStream_t stream) {
  if (!x || !y || M <= 0 || C <= 0) return AIR_ERR_ARG;
  if (!training && (!running_mean || !running_var)) return AIR_ERR_ARG;
  if (!training) {
    const long long n = static_cast<long long>(M) * C;
    const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8));
    bn1d_eval_kernel<<<blocks, 256, 0, stream>>>(x, y, n, C, relu_in, gamma, beta, eps, running_mean, running_var);
    return air_launch_status();
  }
  bn1d_fwd_kernel<<<(C + 63) / 64, 64, 0, stream>>>(x, y, M, C, relu_in, gamma, beta, eps, training, save_mean, save_invstd,
                                                     running_mean, running_var, momentum);
  return air_launch_status();
}

extern "C" int air_bn1d_f32_bwd(const float* dy, const float* x, float* dx, int M, int C, int relu_in, const float* gamma,
                                const float* mean, const float* invstd, float* dgamma, float* dbeta, cudaStream_t stream) {
  if (!dy || !x || !dx || !mean || !invstd || M <= 0 || C <= 0) return AIR_ERR_ARG;
  bn1d_bwd_kernel<<<(C + 63) / 64, 64, 0, stream>>>(dy, x, dx, M, C, relu_in, gamma, mean, invstd, dgamma, dbeta);
  return air_launch_status();
}

extern "C" int air_sigmoid_fwd(const float* x, float* y, long long n, cudaStream_t stream) {
  if (!x || !y || n <= 0) return AIR_ERR_ARG;
  sigmoid_fwd_kernel<<<ew_grid(n, 256), 256, 0, stream>>>(x, y, n);
  return air_launch_status();
}

extern "C" int air_sigmoid_bwd(const float* dy, const float* y, float* dx, long long n, cudaStream_t stream) {
  if (!dy || !y || !dx || n <= 0) return AIR_ERR_ARG;
  sigmoid_bwd_kernel<<<ew_grid(n, 256), 256, 0, stream>>>(dy, y, dx, n);
  return air_launch_status();
}

template <typename AT>
static int copy_channels_impl(const void* src, long long s_ld, const void* mask, long long m_ld, void* dst, long long d_ld,
                                 long long M, int C, cudaStream_t stream) {
  if (!src || !dst || M <= 0 || C % 8 != 0 || (s_ld | d_ld) % 8 != 0 || (mask && m_ld % 8 != 0)) return AIR_ERR_ARG;
  if (M * (C / 8) >= 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;            // 32-bit chunk index in the kernel
  copy_channels_kernel<AT><<<ew_grid(M * (C / 8), 256 * 4), 256, 0, stream>>>(
      reinterpret_cast<const AT*>(src), s_ld, reinterpret_cast<const AT*>(mask), m_ld,
      reinterpret_cast<AT*>(dst), d_ld, M, C);
  return air_launch_status();
}
extern "C" int air_copy_channels(const void* src, long long s_ld, const void* mask, long long m_ld, void* dst, long long d_ld,
                                 long long M, int C, cudaStream_t stream) {
  return copy_channels_impl<__nv_bfloat16>(src, s_ld, mask, m_ld, dst, d_ld, M, C, stream);
}
extern "C" int air_copy_channels_f32(const void* src, long long s_ld, const void* mask, long long m_ld, void* dst, long long d_ld,
                                 long long M, int C, cudaStream_t stream) {
  return copy_channels_impl<float>(src, s_ld, mask, m_ld, dst, d_ld, M, C, stream);
}

template <typename AT>
static int colsum_bf16_impl(const void* x, long long ld, long long M, int C, float* out, cudaStream_t stream) {
  if (!x || !out || M <= 0 || C % 8 != 0 || C > 2048 || ld % 8 != 0) return AIR_ERR_ARG;
  const int rows_per_it = 256 / (C / 8);
  if (rows_per_it < 1) return AIR_ERR_UNSUPPORTED;
  long long blocks = (M + rows_per_it * 16 - 1) / (rows_per_it * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  colsum_kernel<AT><<<static_cast<int>(blocks), 256, 256 * 8 * sizeof(float), stream>>>(
      reinterpret_cast<const AT*>(x), ld, M, C, out);
  return air_launch_status();
}
extern "C" int air_colsum_bf16(const void* x, long long ld, long long M, int C, float* out, cudaStream_t stream) {
  return colsum_bf16_impl<__nv_bfloat16>(x, ld, M, C, out, stream);
}
extern "C" int air_colsum_f32(const void* x, long long ld, long long M, int C, float* out, cudaStream_t stream) {
  return colsum_bf16_impl<float>(x, ld, M, C, out, stream);
}
