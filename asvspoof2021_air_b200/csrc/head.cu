// Pooling / embedding head / loss kernels (fp32 math on small tensors; HBM- or latency-bound).
//
//   air_selfattn_pool_{fwd,bwd} : SelfAttention.forward, resnet.py:23-46 (+ autograd backward)
//   air_linear_{fwd,bwd}        : nn.Linear fc / fc_mu / fc6 / fc7 (resnet.py:142-143,187-189;
//                                 ecapa_tdnn.py:148-149,190-192)
//   air_ocsoftmax_fwd_bwd       : OCSoftmax / AngularIsoLoss forward + analytic backward
//                                 (loss.py:187-206 == :73-97) and the logged CE (main_train.py:355-357)
#include <algorithm>
#include "common.cuh"

namespace air_head {

// Counter-based Gaussian noise standing in for `1e-5*torch.randn(...)` (resnet.py:38): the same
// value is regenerated in the backward from (seed, element index), so nothing is stored.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ float gauss_noise(long long seed, long long idx) {
  if (seed < 0) return 0.f;
  const uint32_t lo = static_cast<uint32_t>(idx), hi = static_cast<uint32_t>(idx >> 32);
  const uint32_t s0 = static_cast<uint32_t>(seed), s1 = static_cast<uint32_t>(seed >> 32);
  const uint32_t a = mix32(lo ^ mix32(hi + 0x9e3779b9U) ^ s0);
  const uint32_t b = mix32(a + 0x85ebca6bU + s1);
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777217.0f);   // (0, 1)
  const float u2 = static_cast<float>(b >> 8) * (1.0f / 16777216.0f);
  return 1e-5f * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// ------------------------------------------------------------------------------------------
// SelfAttention pooling.  One CTA per utterance, blockDim.x == C (<= 1024, multiple of 32).
//   x (B,T,C) bf16;  att (C);  stats (B,2C) = [sum_t x*p | unbiased std_t(x*p + noise)]
//   saved: p (B,T) softmax weights, th (B,T) tanh(x . att)
// ------------------------------------------------------------------------------------------
template <typename AT>
__global__ void selfattn_pool_fwd_kernel(const AT* __restrict__ x, const float* __restrict__ att,
                                         float* __restrict__ stats, float* __restrict__ p_out, float* __restrict__ th_out,
                                         int T, int C, long long seed) {
  extern __shared__ float sh[];           // sw[T], red[32]
  float* sw = sh;
  float* red = sh + T;
  const int b = blockIdx.x, c = threadIdx.x, warp = c >> 5, lane = c & 31, nw = C >> 5;
  const AT* xb = x + static_cast<long long>(b) * T * C;
  for (int t = warp; t < T; t += nw) {
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) acc = fmaf(ld1(xb + t * C + k), att[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) sw[t] = tanhf(acc);
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int t = c; t < T; t += C) mx = fmaxf(mx, sw[t]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int t = c; t < T; t += C) se += expf(sw[t] - mx);
  se = block_sum(se, red);
  __syncthreads();
  for (int t = c; t < T; t += C) {
    const float th = sw[t];
    const float pv = expf(th - mx) / se;
    th_out[b * T + t] = th;
    p_out[b * T + t] = pv;
    sw[t] = pv;
  }
  __syncthreads();
  float sum = 0.f, sumn = 0.f;
  for (int t = 0; t < T; ++t) {
    const float v = ld1(xb + t * C + c) * sw[t];
    sum += v;
    sumn += v + gauss_noise(seed, (static_cast<long long>(b) * T + t) * C + c);
  }
  const float mean = sumn / T;
  float ss = 0.f;
  for (int t = 0; t < T; ++t) {
    const float v = ld1(xb + t * C + c) * sw[t] + gauss_noise(seed, (static_cast<long long>(b) * T + t) * C + c) - mean;
    ss = fmaf(v, v, ss);
  }
  stats[static_cast<long long>(b) * 2 * C + c] = sum;
  stats[static_cast<long long>(b) * 2 * C + C + c] = sqrtf(ss / (T - 1));
}

template <typename AT>
__global__ void selfattn_pool_bwd_kernel(const AT* __restrict__ x, const float* __restrict__ att,
                                         const float* __restrict__ p_in, const float* __restrict__ th_in,
                                         const float* __restrict__ stats, const float* __restrict__ dstats,
                                         AT* __restrict__ dx, float* __restrict__ datt,
                                         int T, int C, long long seed) {
  extern __shared__ float sh[];           // sp[T], sdp[T], sdw[T], dav[C], kk[C], mn[C], red[32]
  float* sp = sh; float* sdp = sh + T; float* sdw = sh + 2 * T;
  float* dav = sh + 3 * T; float* kk = dav + C; float* mn = kk + C; float* red = mn + C;
  const int b = blockIdx.x, c = threadIdx.x, warp = c >> 5, lane = c & 31, nw = C >> 5;
  const AT* xb = x + static_cast<long long>(b) * T * C;
  for (int t = c; t < T; t += C) sp[t] = p_in[b * T + t];
  __syncthreads();
  {
    float sumn = 0.f;
    for (int t = 0; t < T; ++t)
      sumn += ld1(xb + t * C + c) * sp[t] + gauss_noise(seed, (static_cast<long long>(b) * T + t) * C + c);
    const float sd = stats[static_cast<long long>(b) * 2 * C + C + c];
    dav[c] = dstats[static_cast<long long>(b) * 2 * C + c];
    // d std / d v_t = (v_t - mean) / ((T-1) std); a dead channel (std == 0) gets zero gradient
    kk[c] = sd > 0.f ? dstats[static_cast<long long>(b) * 2 * C + C + c] / ((T - 1) * sd) : 0.f;
    mn[c] = sumn / T;
  }
  __syncthreads();
  for (int t = warp; t < T; t += nw) {            // dp[t] = sum_c dweighted[t][c] * x[t][c]
    float acc = 0.f;
    for (int k = lane; k < C; k += 32) {
      const float xv = ld1(xb + t * C + k);
      const float v = xv * sp[t] + gauss_noise(seed, (static_cast<long long>(b) * T + t) * C + k);
      acc = fmaf(dav[k] + kk[k] * (v - mn[k]), xv, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) sdp[t] = acc;
  }
  __syncthreads();
  float dot = 0.f;
  for (int t = c; t < T; t += C) dot = fmaf(sp[t], sdp[t], dot);
  dot = block_sum(dot, red);
  __syncthreads();
  for (int t = c; t < T; t += C) {
    const float th = th_in[b * T + t];
    sdw[t] = sp[t] * (sdp[t] - dot) * (1.f - th * th);     // softmax bwd then tanh bwd
  }
  __syncthreads();
  const float a = att[c];
  float da = 0.f;
  for (int t = 0; t < T; ++t) {
    const float xv = ld1(xb + t * C + c);
    const float v = xv * sp[t] + gauss_noise(seed, (static_cast<long long>(b) * T + t) * C + c);
    const float dwgt = dav[c] + kk[c] * (v - mn[c]);
    st1(dx + (static_cast<long long>(b) * T + t) * C + c, dwgt * sp[t] + sdw[t] * a);
    da = fmaf(sdw[t], xv, da);
  }
  atomicAdd(&datt[c], da);
}

// ------------------------------------------------------------------------------------------
// fp32 linear layer  y = x W^T + b   (x (M,K), W (N,K))
// ------------------------------------------------------------------------------------------
// Small fp32 GEMM with arbitrary element strides (the fully connected layers: fc / fc_mu resnet.py:142-143, fc6 / fc7
// and the SE bottleneck ecapa_tdnn.py:19-23,148-149):  C[i][j] (+)= sum_l A(i,l) * B(l,j) (+ bias[j]) (+ rowsum)
//   A(i,l) = A[i*sai + l*sal],  B(l,j) = B[l*sbl + j*sbj],  C[i][j] = C[i*ldc + j]
// 32 x 32 output tile per 256-thread CTA (each thread 4 consecutive rows of one column), 32-deep K tiles through shared
// memory held in the accumulator type: the A tile is stored l-major so that a thread's four row values are one 16-byte
// (two for fp64) broadcast load per l, and the fp32 -> fp64 conversion of the long reductions happens once per loaded
// element instead of once per multiply.  Every output is still the sequential fma chain over l = 0 .. L-1.
// The next K tile is fetched into registers while the current one is multiplied (a CTA's chain of K tiles is latency-,
// not throughput-bound: 64 CTAs x 96 tiles for fc6 at B = 256).  Split-K form (part != nullptr, fp64 only): CTA z takes
// K tiles [z * l_chunk, (z + 1) * l_chunk) and writes its fp64 partial sums to part[z][i][j]; splitk_reduce_kernel adds
// them in z order -- deterministic, and the fp64 partials keep the result within an ulp of fp32 of the single chain.
template <typename AccT, bool COLSUM>
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, long long sai, long long sal,
                                                            const float* __restrict__ B, long long sbl, long long sbj,
                                                            float* __restrict__ C, long long ldc, int I, int J, int L,
                                                            const float* __restrict__ bias, int accumulate,
                                                            float* __restrict__ colsum_of_a, double* __restrict__ part,
                                                            int l_chunk) {
  __shared__ __align__(16) AccT sa[32][36];                          // sa[l][i]
  __shared__ AccT sb[32][33];                                        // sb[l][j]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // ty 0..7: rows 4 ty .. 4 ty + 3
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int lbeg = part ? blockIdx.z * l_chunk : 0, lend = part ? min(L, lbeg + l_chunk) : L;
  // AccT = double for long reductions (fc6: K = 3072): the 1-D BatchNorms that follow (bn5 over B rows) amplify the
  // summation-order noise of an fp32 reduction; float otherwise
  AccT acc[4] = {0, 0, 0, 0};
  float csum[4] = {0.f, 0.f, 0.f, 0.f};
  // tile coordinates of this thread's loads: lanes follow the contiguous index of each operand
  int ai[4], al[4], bl[4], bj[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int u = ty + 8 * r;                                         // the slow index of this load
    if (sal == 1) { ai[r] = u; al[r] = tx; } else { ai[r] = tx; al[r] = u; }
    if (sbj == 1) { bl[r] = u; bj[r] = tx; } else { bl[r] = tx; bj[r] = u; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int l0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gi = i0 + ai[r], gl = l0 + al[r];
      ra[r] = (gi < I && gl < lend) ? A[gi * sai + gl * sal] : 0.f;
      const int hl = l0 + bl[r], gj = j0 + bj[r];
      rb[r] = (hl < lend && gj < J) ? B[hl * sbl + gj * sbj] : 0.f;
    }
  };
  if (lbeg < lend) fetch(lbeg);
  for (int l0 = lbeg; l0 < lend; l0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      sa[al[r]][ai[r]] = static_cast<AccT>(ra[r]);
      sb[bl[r]][bj[r]] = static_cast<AccT>(rb[r]);
    }
    __syncthreads();
    if (l0 + 32 < lend) fetch(l0 + 32);
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      const AccT bv = sb[l][tx];
      AccT av[4];
      if constexpr (sizeof(AccT) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&sa[l][4 * ty]);
        av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
      } else {
        const double2 t0 = *reinterpret_cast<const double2*>(&sa[l][4 * ty]);
        const double2 t1 = *reinterpret_cast<const double2*>(&sa[l][4 * ty + 2]);
        av[0] = t0.x; av[1] = t0.y; av[2] = t1.x; av[3] = t1.y;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r] = fma(av[r], bv, acc[r]);
        if constexpr (COLSUM) csum[r] += static_cast<float>(av[r]);
      }
    }
    __syncthreads();
  }
  const int j = j0 + tx;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + 4 * ty + r;
    if (i < I && j < J) {
      if (part) {
        part[(static_cast<long long>(blockIdx.z) * I + i) * J + j] = static_cast<double>(acc[r]);
      } else {
        float v = static_cast<float>(acc[r] + (bias ? static_cast<AccT>(bias[j]) : static_cast<AccT>(0)));
        float* cp = C + i * ldc + j;
        *cp = accumulate ? *cp + v : v;
      }
    }
    // row sums of A (the bias gradient when A = dy^T): written once per row by the j-tile 0, column 0 thread
    if constexpr (COLSUM) { if (blockIdx.x == 0 && tx == 0 && i < I) colsum_of_a[i] += csum[r]; }
  }
}

__global__ void splitk_reduce_kernel(const double* __restrict__ part, int splits, long long n, int J, const float* __restrict__ bias,
                                     float* __restrict__ C, long long ldc) {
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += part[z * n + e];
    const long long i = e / J;
    const int j = static_cast<int>(e - i * J);
    C[i * ldc + j] = static_cast<float>(s + (bias ? static_cast<double>(bias[j]) : 0.0));
  }
}

static int launch_sgemm(const float* A, long long sai, long long sal, const float* B, long long sbl, long long sbj,
                        float* C, long long ldc, int I, int J, int L, const float* bias, int accumulate, float* colsum,
                        cudaStream_t stream) {
  dim3 grid((J + 31) / 32, (I + 31) / 32);
#define AIR_SGEMM(T, CS) sgemm_strided_kernel<T, CS><<<grid, 256, 0, stream>>>(A, sai, sal, B, sbl, sbj, C, ldc, I, J, L, bias, accumulate, colsum, nullptr, 0)
  if (L >= 2048) { if (colsum) AIR_SGEMM(double, true); else AIR_SGEMM(double, false); }
  else { if (colsum) AIR_SGEMM(float, true); else AIR_SGEMM(float, false); }
#undef AIR_SGEMM
  return air_launch_status();
}

// ------------------------------------------------------------------------------------------
// OC-Softmax forward + backward (single CTA, 1024 threads: the kernel sits between the forward and the backward pass, nothing
// overlaps it, and at 256 threads it took 0.12 ms for 256 x 256 values).
//   loss = mean_i softplus(alpha * m_i),  m_i = r_real - s_i (label 0) | s_i - r_fake (label 1) | s_i
//   s_i = <x_i/|x_i|, w/|w|>, score_i = -s_i.   softplus: beta 1, threshold 20.
//   dfeat = grad_scale * dloss/dx, dcenter += grad_scale * dloss/dcenter.
// Optional CE over `logits` (B, ncls) for logging only.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) ocsoftmax_kernel(const float* __restrict__ x, const long long* __restrict__ labels,
                                                         const float* __restrict__ center, int B, int D,
                                                         float r_real, float r_fake, float alpha, float grad_scale,
                                                         float* __restrict__ loss_out, float* __restrict__ score,
                                                         float* __restrict__ dfeat, float* __restrict__ dcenter,
                                                         const float* __restrict__ logits, int ncls, float* __restrict__ ce_out) {
  extern __shared__ float sh[];           // wn[D], s[B], coef[B], rinv[B], red[32], part[blockDim / D][D]
  float* wn = sh; float* ss = sh + D; float* coef = ss + B; float* rinv = coef + B; float* red = rinv + B; float* part = red + 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  float wsq = 0.f;
  for (int d = tid; d < D; d += blockDim.x) wsq = fmaf(center[d], center[d], wsq);
  wsq = block_sum(wsq, red);
  const float wnorm = fmaxf(sqrtf(wsq), 1e-12f);
  for (int d = tid; d < D; d += blockDim.x) wn[d] = center[d] / wnorm;
  __syncthreads();
  float lsum = 0.f;
  for (int i = warp; i < B; i += nw) {
    float xx = 0.f, xw = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = x[static_cast<long long>(i) * D + d]; xx = fmaf(v, v, xx); xw = fmaf(v, wn[d], xw); }
    xx = warp_sum(xx); xw = warp_sum(xw);
    if (lane == 0) {
      const float xn = fmaxf(sqrtf(xx), 1e-12f);
      const float s = xw / xn;
      const long long lab = labels ? labels[i] : 0;
      float m = s, sign = 1.f;
      if (lab == 0) { m = r_real - s; sign = -1.f; } else if (lab == 1) { m = s - r_fake; }
      const float z = alpha * m;
      const float sp = z > 20.f ? z : log1pf(expf(z));
      const float dsp = z > 20.f ? 1.f : 1.f / (1.f + expf(-z));
      lsum += sp;
      ss[i] = s; rinv[i] = 1.f / xn;
      coef[i] = grad_scale * sign * alpha * dsp / B;          // dloss/ds_i
      if (score) score[i] = -s;
    }
  }
  lsum = block_sum(lsum, red);
  if (tid == 0 && loss_out) loss_out[0] = lsum / B;
  __syncthreads();
  if (dfeat) {
    for (long long j = tid; j < static_cast<long long>(B) * D; j += blockDim.x) {
      const int i = static_cast<int>(j / D), d = static_cast<int>(j - static_cast<long long>(i) * D);
      const float xh = x[j] * rinv[i];
      dfeat[j] = coef[i] * (wn[d] - ss[i] * xh) * rinv[i];
    }
  }
  if (dcenter) {
    // the rows are dealt to P = blockDim / D thread groups (fixed assignment, partial sums combined in a fixed order: the
    // result does not depend on timing); D > blockDim falls back to one group striding over d
    const int P = blockDim.x >= static_cast<unsigned>(D) ? static_cast<int>(blockDim.x) / D : 1;
    if (P > 1) {
      const int d = tid % D, g = tid / D;
      if (g < P) {
        float acc = 0.f;
        for (int i = g; i < B; i += P) acc = fmaf(coef[i], x[static_cast<long long>(i) * D + d] * rinv[i] - ss[i] * wn[d], acc);
        part[g * D + d] = acc;
      }
      __syncthreads();
      if (tid < D) {
        float acc = 0.f;
        for (int g2 = 0; g2 < P; ++g2) acc += part[g2 * D + tid];
        dcenter[tid] += acc / wnorm;
      }
    } else {
      for (int d = tid; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < B; ++i) acc = fmaf(coef[i], x[static_cast<long long>(i) * D + d] * rinv[i] - ss[i] * wn[d], acc);
        dcenter[d] += acc / wnorm;
      }
    }
  }
  if (logits && ce_out) {
    float cs = 0.f;
    for (int i = tid; i < B; i += blockDim.x) {
      float mx = -INFINITY;
      for (int k = 0; k < ncls; ++k) mx = fmaxf(mx, logits[i * ncls + k]);
      float se = 0.f;
      for (int k = 0; k < ncls; ++k) se += expf(logits[i * ncls + k] - mx);
      const long long lab = labels ? labels[i] : 0;
      cs += (mx + logf(se)) - logits[i * ncls + (lab >= 0 && lab < ncls ? lab : 0)];
    }
    cs = block_sum(cs, red);
    if (tid == 0) ce_out[0] = cs / B;
  }
}

}  // namespace air_head

using namespace air_head;

template <typename AT>
static int pool_fwd_impl(const void* x, const float* att, float* stats, float* p_out, float* th_out,
                         int B, int T, int C, long long noise_seed, cudaStream_t stream) {
  if (!x || !att || !stats || !p_out || !th_out || B <= 0 || T < 2 || C % 32 != 0 || C > 1024) return AIR_ERR_ARG;
  selfattn_pool_fwd_kernel<AT><<<B, C, (T + 32) * sizeof(float), stream>>>(
      reinterpret_cast<const AT*>(x), att, stats, p_out, th_out, T, C, noise_seed);
  return air_launch_status();
}

template <typename AT>
static int pool_bwd_impl(const void* x, const float* att, const float* p_in, const float* th_in,
                         const float* stats, const float* dstats, void* dx, float* datt,
                         int B, int T, int C, long long noise_seed, cudaStream_t stream) {
  if (!x || !att || !p_in || !th_in || !stats || !dstats || !dx || !datt || B <= 0 || T < 2 || C % 32 != 0 || C > 1024)
    return AIR_ERR_ARG;
  selfattn_pool_bwd_kernel<AT><<<B, C, (3 * T + 3 * C + 32) * sizeof(float), stream>>>(
      reinterpret_cast<const AT*>(x), att, p_in, th_in, stats, dstats, reinterpret_cast<AT*>(dx), datt, T, C, noise_seed);
  return air_launch_status();
}

extern "C" int air_selfattn_pool_fwd(const void* x, const float* att, float* stats, float* p_out, float* th_out,
                                     int B, int T, int C, long long noise_seed, cudaStream_t stream) {
  return pool_fwd_impl<__nv_bfloat16>(x, att, stats, p_out, th_out, B, T, C, noise_seed, stream);
}
extern "C" int air_selfattn_pool_fwd_f32(const void* x, const float* att, float* stats, float* p_out, float* th_out,
                                         int B, int T, int C, long long noise_seed, cudaStream_t stream) {
  return pool_fwd_impl<float>(x, att, stats, p_out, th_out, B, T, C, noise_seed, stream);
}
extern "C" int air_selfattn_pool_bwd(const void* x, const float* att, const float* p_in, const float* th_in,
                                     const float* stats, const float* dstats, void* dx, float* datt,
                                     int B, int T, int C, long long noise_seed, cudaStream_t stream) {
  return pool_bwd_impl<__nv_bfloat16>(x, att, p_in, th_in, stats, dstats, dx, datt, B, T, C, noise_seed, stream);
}
extern "C" int air_selfattn_pool_bwd_f32(const void* x, const float* att, const float* p_in, const float* th_in,
                                         const float* stats, const float* dstats, void* dx, float* datt,
                                         int B, int T, int C, long long noise_seed, cudaStream_t stream) {
  return pool_bwd_impl<float>(x, att, p_in, th_in, stats, dstats, dx, datt, B, T, C, noise_seed, stream);
}

extern "C" int air_linear_fwd(const float* x, const float* W, const float* bias, float* y, int M, int N, int K,
                              cudaStream_t stream) {
  if (!x || !W || !y || M <= 0 || N <= 0 || K <= 0) return AIR_ERR_ARG;
  // y[m][n] = sum_k x[m][k] W[n][k] + b[n]
  return launch_sgemm(x, K, 1, W, 1, K, y, N, M, N, K, bias, 0, nullptr, stream);
}

// The same product with the K dimension dealt to `splits` CTAs per output tile (fc6 of ECAPA, K = 3072: one CTA chain
// of 96 K tiles took 0.30 ms whatever the batch).  fp64 accumulation and fp64 partial sums in `scratch`
// (>= splits * M * N doubles), added in split order: deterministic.
extern "C" int air_linear_fwd_splitk(const float* x, const float* W, const float* bias, float* y, int M, int N, int K,
                                     double* scratch, int splits, cudaStream_t stream) {
  if (!x || !W || !y || !scratch || M <= 0 || N <= 0 || K <= 0 || splits < 1 || splits > 64) return AIR_ERR_ARG;
  const int tiles = (K + 31) / 32;
  const int l_chunk = ((tiles + splits - 1) / splits) * 32;
  const int used = (K + l_chunk - 1) / l_chunk;                      // splits that own at least one K tile
  dim3 grid((N + 31) / 32, (M + 31) / 32, used);
  sgemm_strided_kernel<double, false><<<grid, 256, 0, stream>>>(x, K, 1, W, 1, K, y, N, M, N, K, nullptr, 0, nullptr, scratch, l_chunk);
  int st = air_launch_status();
  if (st != AIR_OK) return st;
  const long long n = static_cast<long long>(M) * N;
  splitk_reduce_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 148 * 8)), 256, 0, stream>>>(scratch, used, n, N, bias, y, N);
  return air_launch_status();
}

// y (+)= x W[:, slice]^T + b with a row stride on W (column slice of a wider weight, e.g. the
// mean / std blocks of ECAPA's attention.0.weight (128, 4608), ecapa_tdnn.py:139,173-175)
extern "C" int air_linear_fwd_ld(const float* x, const float* W, long long ldw, const float* bias, float* y, int M, int N, int K,
                                 int accumulate, cudaStream_t stream) {
  if (!x || !W || !y || M <= 0 || N <= 0 || K <= 0 || ldw < K) return AIR_ERR_ARG;
  return launch_sgemm(x, K, 1, W, 1, ldw, y, N, M, N, K, bias, accumulate, nullptr, stream);
}

extern "C" int air_linear_bwd_ld(const float* x, const float* W, long long ldw, const float* dy, float* dx, float* dW, float* db,
                                 int M, int N, int K, cudaStream_t stream) {
  if (!x || !W || !dy || M <= 0 || N <= 0 || K <= 0 || ldw < K) return AIR_ERR_ARG;
  int st = AIR_OK;
  // dx[m][k] = sum_n dy[m][n] W[n][k]
  if (dx) st = launch_sgemm(dy, N, 1, W, ldw, 1, dx, K, M, K, N, nullptr, 0, nullptr, stream);
  // dW[n][k] += sum_m dy[m][n] x[m][k];  db[n] += sum_m dy[m][n]
  if (dW && st == AIR_OK) st = launch_sgemm(dy, 1, N, x, K, 1, dW, ldw, N, K, M, nullptr, 1, db, stream);
  return st;
}

extern "C" int air_linear_bwd(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db,
                              int M, int N, int K, cudaStream_t stream) {
  return air_linear_bwd_ld(x, W, K, dy, dx, dW, db, M, N, K, stream);
}

extern "C" int air_ocsoftmax_fwd_bwd(const float* x, const long long* labels, const float* center, int B, int D,
                                     float r_real, float r_fake, float alpha, float grad_scale,
                                     float* loss, float* score, float* dfeat, float* dcenter,
                                     const float* logits, int ncls, float* ce, cudaStream_t stream) {
  if (!x || !center || B <= 0 || D <= 0) return AIR_ERR_ARG;
  const int threads = 1024;
  const size_t smem = (static_cast<size_t>(D) + 3 * static_cast<size_t>(B) + 32 + (threads >= D ? static_cast<size_t>(threads / D) * D : 0)) * sizeof(float);
  if (smem > 200 * 1024) return AIR_ERR_UNSUPPORTED;          // B <= ~17 000 at D = 256; callers chunk above that
  if (smem > 48 * 1024) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(ocsoftmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return static_cast<int>(e);
      attr_done = true;
    }
  }
  ocsoftmax_kernel<<<1, threads, smem, stream>>>(x, labels, center, B, D, r_real, r_fake, alpha, grad_scale,
                                              loss, score, dfeat, dcenter, logits, ncls, ce);
  return air_launch_status();
}
