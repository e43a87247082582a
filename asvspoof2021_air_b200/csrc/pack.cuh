// Element mappings of the two weight-packing layouts (fp32 master weights -> bf16 tensor-core operand images),
// shared by the per-tensor kernels (conv_gemm.cu, conv_patch.cu) and the batched job kernel (pack.cu).
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace air_pack {

// UMMA "column of rows" tiles of the generic implicit-GEMM kernel: destination order [n_tile][kb][chunk][n_local][e]
//   value(n, k) = src[n*sn + (k / inner)*so + (k % inner)*si]
__device__ __forceinline__ void gemm_pack_elem(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long i,
                                               int N, int K, int KB, int block_n, long long sn, int inner, long long so,
                                               long long si) {
  const int e = static_cast<int>(i & 7);
  long long t = i >> 3;
  const int n_local = static_cast<int>(t % block_n); t /= block_n;
  const int c = static_cast<int>(t & 7); t >>= 3;
  const int kb = static_cast<int>(t % KB);
  const int n_tile = static_cast<int>(t / KB);
  const int n = n_tile * block_n + n_local;
  const int k = kb * 64 + c * 8 + e;
  float v = 0.f;
  if (k < K && n < N) v = src[n * sn + static_cast<long long>(k / inner) * so + static_cast<long long>(k % inner) * si];
  dst[i] = f2bf(v);
}

// pre-swizzled [N rows][CB channels] K-major slices of the patch kernel, one per (channel block, tap)
//   mode 0: w is [N][taps][C], value = w[n][tap][ch];  mode 1: w is [C][taps][N], value = w[ch][taps-1-tap][n]
__device__ __forceinline__ void patch_pack_elem(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, long long i,
                                                int C, int N, int CB, int taps, int mode) {
  const uint32_t mask = CB == 64 ? 7u : (CB == 32 ? 3u : 1u);
  const int k = static_cast<int>(i % CB);
  long long t = i / CB;
  const int n = static_cast<int>(t % N); t /= N;
  const int tap = static_cast<int>(t % taps);
  const int cb = static_cast<int>(t / taps);
  const int ch = cb * CB + k;
  const float v = mode == 0 ? w[(static_cast<long long>(n) * taps + tap) * C + ch]
                            : w[(static_cast<long long>(ch) * taps + (taps - 1 - tap)) * N + n];
  const uint32_t off = tc05::swizzle_offset(static_cast<uint32_t>(n) * CB * 2 + (k >> 3) * 16, mask) + (k & 7) * 2;
  dst[(static_cast<long long>(cb) * taps + tap) * N * CB + (off >> 1)] = f2bf(v);
}

}  // namespace air_pack
