// Element mappings of the two weight-packing layouts (fp32 master weights -> bf16 tensor-core operand images),
// shared by the per-tensor kernels (conv_gemm.cu, conv_patch.cu) and the batched job kernel (pack.cu).
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace air_pack {

// SWIZZLE_128B K-major tiles of the generic implicit-GEMM kernel: destination = [n_tile][kb] tiles of
// [block_n rows][64 K elements] (128 B per row), 16-byte chunk c of row r stored at chunk position c ^ (r & 7)
//   value(n, k) = src[n*sn + (k / inner)*so + (k % inner)*si]
__device__ __forceinline__ void gemm_pack_elem(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long i,
                                               int N, int K, int KB, int block_n, long long sn, int inner, long long so,
                                               long long si) {
  const int kk = static_cast<int>(i & 63);               // logical element index: [n_tile][kb][n_local][kk]
  long long t = i >> 6;
  const int n_local = static_cast<int>(t % block_n); t /= block_n;
  const int kb = static_cast<int>(t % KB);
  const int n_tile = static_cast<int>(t / KB);
  const int n = n_tile * block_n + n_local;
  const int k = kb * 64 + kk;
  float v = 0.f;
  if (k < K && n < N) v = src[n * sn + static_cast<long long>(k / inner) * so + static_cast<long long>(k % inner) * si];
  const long long tile = (static_cast<long long>(n_tile) * KB + kb) * block_n * 64;
  dst[tile + n_local * 64 + ((((kk >> 3) ^ (n_local & 7)) << 3) | (kk & 7))] = f2bf(v);
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
