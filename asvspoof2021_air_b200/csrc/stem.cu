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

template <typename T>
struct StemParams {
  const T* x; int B, H, W, Ho, Wo, kh, kw, sh, sw, ph, pw;
  const float* w; T* y; const T* dy; float* dw; long long M;
};

// Forward: a thread computes 4 consecutive output columns x 16 channels; a weight vector (16 channels of one tap,
// 4 x LDS.128) feeds 64 FMAs, an input row segment of 6 samples is loaded once for the 3 horizontal taps.
constexpr int FW_PIX = 4;
template <typename T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const StemParams<T> p) {
  __shared__ __align__(16) float swt[MAX_TAPS * CO];          // [tap][co]
  const int taps = p.kh * p.kw;
  for (int i = threadIdx.x; i < CO * taps; i += blockDim.x) { const int c = i / taps, t = i - c * taps; swt[t * CO + c] = p.w[i]; }
  __syncthreads();
  const int WQ = (p.Wo + FW_PIX - 1) / FW_PIX;
  const long long total = static_cast<long long>(p.B) * p.Ho * WQ;
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < total;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wq = static_cast<int>(q % WQ); const long long t = q / WQ;
    const int ho = static_cast<int>(t % p.Ho), b = static_cast<int>(t / p.Ho);
    const int wo0 = wq * FW_PIX;
    float acc[FW_PIX][CO];
#pragma unroll
    for (int px = 0; px < FW_PIX; ++px)
#pragma unroll
      for (int c = 0; c < CO; ++c) acc[px][c] = 0.f;
    const T* img = p.x + static_cast<long long>(b) * p.H * p.W;
    for (int i = 0; i < p.kh; ++i) {
      const int hi = ho * p.sh - p.ph + i;
      if (hi < 0 || hi >= p.H) continue;
      const T* row = img + static_cast<long long>(hi) * p.W;
      float xin[FW_PIX + 2];
#pragma unroll
      for (int u = 0; u < FW_PIX + 2; ++u) {
        const int wi = wo0 - p.pw + u;
        xin[u] = (wi >= 0 && wi < p.W) ? ld1(row + wi) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const float4* wv = reinterpret_cast<const float4*>(swt + (i * 3 + j) * CO);
        const float4 w0 = wv[0], w1 = wv[1], w2 = wv[2], w3 = wv[3];
        const float wl[CO] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int px = 0; px < FW_PIX; ++px)
#pragma unroll
          for (int c = 0; c < CO; ++c) acc[px][c] = fmaf(xin[px + j], wl[c], acc[px][c]);
      }
    }
    const long long m0 = (static_cast<long long>(b) * p.Ho + ho) * p.Wo + wo0;
#pragma unroll
    for (int px = 0; px < FW_PIX; ++px) {
      if (wo0 + px < p.Wo) {
        st8(p.y + (m0 + px) * CO, acc[px]);
        st8(p.y + (m0 + px) * CO + 8, acc[px] + 8);
      }
    }
  }
}

// dw[co][tap] += sum_m dy[m][co] * x[pix(m, tap)].  One warp per kernel row i; a block walks whole output rows (b, ho), so
// the pixel decode is per row, not per pixel (the first version spent most of its instructions on two 64-bit modulo
// operations per pixel); a lane takes the pixels wo = lane, lane + 32, ... of the row, two per iteration: it reads the 16
// gradients of each pixel (2 x 16 B) and the 3 input samples of row i once for 48 FMAs.  Partial sums stay in registers
// over all rows of the block and are reduced by warp shuffles at the end.
template <typename T>
__global__ void __launch_bounds__(320) stem_wgrad_kernel(const StemParams<T> p) {
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;      // blockDim.x = 32 * kh
  float acc[3][CO];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[j][c] = 0.f;
  const int rows = p.B * p.Ho;
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const int b = r / p.Ho, ho = r - b * p.Ho;
    const int hi = ho * p.sh - p.ph + i;
    if (hi < 0 || hi >= p.H) continue;                            // warp-uniform
    const T* row = p.x + (static_cast<long long>(b) * p.H + hi) * p.W;
    const T* dyr = p.dy + static_cast<long long>(r) * p.Wo * CO;
    for (int wo = lane; wo < p.Wo; wo += 64) {
      const int wo2 = wo + 32;
      const bool two = wo2 < p.Wo;
      float g0[CO], g1[CO], x0[3], x1[3];
      ld8(dyr + static_cast<long long>(wo) * CO, g0); ld8(dyr + static_cast<long long>(wo) * CO + 8, g0 + 8);
      if (two) { ld8(dyr + static_cast<long long>(wo2) * CO, g1); ld8(dyr + static_cast<long long>(wo2) * CO + 8, g1 + 8); }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int wi = wo - p.pw + j, wj = wo2 - p.pw + j;
        x0[j] = (wi >= 0 && wi < p.W) ? ld1(row + wi) : 0.f;
        x1[j] = (two && wj >= 0 && wj < p.W) ? ld1(row + wj) : 0.f;
      }
      if (!two) {
#pragma unroll
        for (int c = 0; c < CO; ++c) g1[c] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[j][c] = fmaf(g1[c], x1[j], fmaf(g0[c], x0[j], acc[j][c]));
    }
  }
  const int taps = p.kh * 3;
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const float v = warp_sum(acc[j][c]);
      if (lane == 0) atomicAdd(&p.dw[c * taps + i * 3 + j], v);
    }
}

}  // namespace air_stem

using namespace air_stem;

template <typename T>
static int stem_fill(StemParams<T>& p, const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw) {
  if (!x || B <= 0 || kh * kw > MAX_TAPS) return AIR_ERR_ARG;
  p.x = reinterpret_cast<const T*>(x); p.B = B; p.H = H; p.W = W;
  p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw;
  p.Ho = (H + 2 * ph - kh) / sh + 1; p.Wo = (W + 2 * pw - kw) / sw + 1;
  p.M = static_cast<long long>(B) * p.Ho * p.Wo;
  return AIR_OK;
}

template <typename T>
static int stem_fwd_impl(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                         const float* w, int Cout, void* y, cudaStream_t stream) {
  StemParams<T> p{};
  if (int e = stem_fill(p, x, B, H, W, kh, kw, sh, sw, ph, pw)) return e;
  if (!w || !y || Cout != CO || kw != 3 || sw != 1) return AIR_ERR_UNSUPPORTED;
  p.w = w; p.y = reinterpret_cast<T*>(y);
  const long long quads = static_cast<long long>(B) * p.Ho * ((p.Wo + FW_PIX - 1) / FW_PIX);
  const int blocks = static_cast<int>(std::min<long long>((quads + 255) / 256, 148 * 16));
  stem_fwd_kernel<T><<<blocks, 256, 0, stream>>>(p);
  return air_launch_status();
}

template <typename T>
static int stem_wgrad_impl(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                           const void* dy, int Cout, float* dw, cudaStream_t stream) {
  StemParams<T> p{};
  if (int e = stem_fill(p, x, B, H, W, kh, kw, sh, sw, ph, pw)) return e;
  if (!dy || !dw || Cout != CO || kw != 3 || sw != 1 || kh > 10) return AIR_ERR_UNSUPPORTED;
  p.dy = reinterpret_cast<const T*>(dy); p.dw = dw;
  if (static_cast<long long>(B) * p.Ho > 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;
  const int blocks = static_cast<int>(std::min<long long>(static_cast<long long>(B) * p.Ho, 148 * 2));      // two resident blocks per SM (96 registers x 288 threads)
  stem_wgrad_kernel<T><<<blocks, 32 * kh, 0, stream>>>(p);
  return air_launch_status();
}

extern "C" int air_stem_conv_fwd(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                                 const float* w, int Cout, void* y, cudaStream_t stream) {
  return stem_fwd_impl<__nv_bfloat16>(x, B, H, W, kh, kw, sh, sw, ph, pw, w, Cout, y, stream);
}
extern "C" int air_stem_conv_fwd_f32(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                                     const float* w, int Cout, void* y, cudaStream_t stream) {
  return stem_fwd_impl<float>(x, B, H, W, kh, kw, sh, sw, ph, pw, w, Cout, y, stream);
}
extern "C" int air_stem_conv_wgrad(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                                   const void* dy, int Cout, float* dw, cudaStream_t stream) {
  return stem_wgrad_impl<__nv_bfloat16>(x, B, H, W, kh, kw, sh, sw, ph, pw, dy, Cout, dw, stream);
}
extern "C" int air_stem_conv_wgrad_f32(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                                       const void* dy, int Cout, float* dw, cudaStream_t stream) {
  return stem_wgrad_impl<float>(x, B, H, W, kh, kw, sh, sw, ph, pw, dy, Cout, dw, stream);
}
