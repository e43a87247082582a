// ResNet stem: conv1 = nn.Conv2d(1, 16, kernel_size=(9,3), stride=(3,1), padding=(1,1), bias=False)
// (resnet.py:131,176).  Cin = 1 / K = 27 / N = 16 is not a tensor-core shape: direct CUDA-core
// kernels, forward and weight gradient (no data gradient: the LFCC input needs none).
//   x  (B, H, W)        bf16   (H = 60 cepstral rows, W = 750 frames; NHWC with C = 1)
//   y  (B, Ho, Wo, 16)  bf16
//   w  (16, kh*kw)      fp32   GEMM layout [Cout][kh][kw][1]
#include <algorithm>
#include "common.cuh"

namespace air_stem {

constexpr int CO = 16;
constexpr int MAX_TAPS = 32;

struct StemParams {
  const __nv_bfloat16* x; int B, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw;
  const float* w; __nv_bfloat16* y; const __nv_bfloat16* dy; float* dw; long long M;
};

__global__ void __launch_bounds__(256) stem_fwd_kernel(const StemParams p) {
  __shared__ float sw[CO * MAX_TAPS];
  const int taps = p.kh * p.kw;
  for (int i = threadIdx.x; i < CO * taps; i += blockDim.x) sw[i] = p.w[i];
  __syncthreads();
  for (long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; m < p.M;
       m += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(m % p.Wo); const long long t = m / p.Wo;
    const int ho = static_cast<int>(t % p.Ho), b = static_cast<int>(t / p.Ho);
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = 0.f;
    const __nv_bfloat16* img = p.x + static_cast<long long>(b) * p.H * p.W;
    for (int i = 0; i < p.kh; ++i) {
      const int hi = ho * p.sh - p.ph + i;
      if (hi < 0 || hi >= p.H) continue;
      for (int j = 0; j < p.kw; ++j) {
        const int wi = wo * p.sw - p.pw + j;
        if (wi < 0 || wi >= p.W) continue;
        const float xv = bf2f(img[hi * p.W + wi]);
        const int tp = i * p.kw + j;
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[c] = fmaf(xv, sw[c * taps + tp], acc[c]);
      }
    }
    bf16x8* op = reinterpret_cast<bf16x8*>(p.y + m * CO);
    op[0] = pack8(acc);
    op[1] = pack8(acc + 8);
  }
}

// dw[co][tap] += sum_m dy[m][co] * x[pix(m, tap)]
constexpr int WG_THREADS = 512;
constexpr int WG_PIX = 128;
__global__ void __launch_bounds__(WG_THREADS) stem_wgrad_kernel(const StemParams p) {
  __shared__ float sdy[WG_PIX][CO];
  __shared__ float sx[WG_PIX][MAX_TAPS + 1];
  const int taps = p.kh * p.kw;
  const int co = threadIdx.x / taps, tp = threadIdx.x - co * taps;
  const bool active = threadIdx.x < CO * taps;
  float acc = 0.f;
  const long long chunks = (p.M + WG_PIX - 1) / WG_PIX;
  for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const long long m0 = ch * WG_PIX;
    for (int i = threadIdx.x; i < WG_PIX * CO; i += WG_THREADS) {
      const int pp = i / CO, c = i - pp * CO;
      const long long m = m0 + pp;
      sdy[pp][c] = m < p.M ? bf2f(p.dy[m * CO + c]) : 0.f;
    }
    for (int i = threadIdx.x; i < WG_PIX * taps; i += WG_THREADS) {
      const int pp = i / taps, t2 = i - pp * taps;
      const long long m = m0 + pp;
      float v = 0.f;
      if (m < p.M) {
        const int wo = static_cast<int>(m % p.Wo); const long long t = m / p.Wo;
        const int ho = static_cast<int>(t % p.Ho), b = static_cast<int>(t / p.Ho);
        const int ki = t2 / p.kw, kj = t2 - ki * p.kw;
        const int hi = ho * p.sh - p.ph + ki, wi = wo * p.sw - p.pw + kj;
        if (hi >= 0 && hi < p.H && wi >= 0 && wi < p.W) v = bf2f(p.x[(static_cast<long long>(b) * p.H + hi) * p.W + wi]);
      }
      sx[pp][t2] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll 8
      for (int pp = 0; pp < WG_PIX; ++pp) acc = fmaf(sdy[pp][co], sx[pp][tp], acc);
    }
    __syncthreads();
  }
  if (active) atomicAdd(&p.dw[co * taps + tp], acc);
}

}  // namespace air_stem

using namespace air_stem;

static int stem_fill(StemParams& p, const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw) {
  if (!x || B <= 0 || kh * kw > MAX_TAPS) return AIR_ERR_ARG;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.B = B; p.H = H; p.W = W;
  p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw;
  p.Ho = (H + 2 * ph - kh) / sh + 1; p.Wo = (W + 2 * pw - kw) / sw + 1;
  p.M = static_cast<long long>(B) * p.Ho * p.Wo;
  return AIR_OK;
}

extern "C" int air_stem_conv_fwd(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                                 const float* w, int Cout, void* y, cudaStream_t stream) {
  StemParams p{};
  if (int e = stem_fill(p, x, B, H, W, kh, kw, sh, sw, ph, pw)) return e;
  if (!w || !y || Cout != CO) return AIR_ERR_UNSUPPORTED;
  p.w = w; p.y = reinterpret_cast<__nv_bfloat16*>(y);
  const int blocks = static_cast<int>(std::min<long long>((p.M + 255) / 256, 148 * 16));
  stem_fwd_kernel<<<blocks, 256, 0, stream>>>(p);
  return air_launch_status();
}

extern "C" int air_stem_conv_wgrad(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                                   const void* dy, int Cout, float* dw, cudaStream_t stream) {
  StemParams p{};
  if (int e = stem_fill(p, x, B, H, W, kh, kw, sh, sw, ph, pw)) return e;
  if (!dy || !dw || Cout != CO || CO * kh * kw > WG_THREADS) return AIR_ERR_UNSUPPORTED;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.dw = dw;
  const long long chunks = (p.M + WG_PIX - 1) / WG_PIX;
  const int blocks = static_cast<int>(std::min<long long>(chunks, 148 * 2));
  stem_wgrad_kernel<<<blocks, WG_THREADS, 0, stream>>>(p);
  return air_launch_status();
}
