// fp32 parity mode: an fp32 tensor as a sum of bf16 terms for the bf16 tensor-core kernels.
//
//   x = hi + lo + O(2^-17 |x|),   hi = bf16(x),  lo = bf16(x - hi)
//   x * w ~= hi_x hi_w + lo_x hi_w + hi_x lo_w     (the dropped lo_x lo_w term is 2^-16 relative)
//
// The three products share one accumulator when the terms are CONCATENATED along the contraction axis: a convolution
// over [hi_x | lo_x | hi_x] (3 C channels) with weights [hi_w | hi_w | lo_w] is an ordinary bf16 convolution with three
// times the input channels, so the tcgen05 kernels run unchanged and only their epilogue stores fp32 (AIR_CONV_F32_OUT).
// This kernel writes such a concatenation:  out[m][t * C + c] = term_t(x[m][c]),  term = hi when bit t of `lo_mask` is
// clear, lo when it is set.  out is bf16 (activations, gradients) or fp32 holding the same bf16-exact values (weights:
// the packing kernels read fp32 master weights).
#include "common.cuh"

namespace air_split {

template <typename TO>
__global__ void __launch_bounds__(256) split_terms_kernel(const float* __restrict__ x, long long x_ld, long long M, int C,
                                                          TO* __restrict__ out, long long out_ld, int nterms, unsigned lo_mask) {
  const int cpr = C >> 3;
  const long long total = M * cpr;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / cpr;
    const int c0 = static_cast<int>(i - m * cpr) << 3;
    float v[8], hi[8], lo[8];
    ld8(x + m * x_ld + c0, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      hi[k] = __bfloat162float(__float2bfloat16_rn(v[k]));
      lo[k] = __bfloat162float(__float2bfloat16_rn(v[k] - hi[k]));
    }
    for (int t = 0; t < nterms; ++t) st8(out + m * out_ld + static_cast<long long>(t) * C + c0, ((lo_mask >> t) & 1u) ? lo : hi);
  }
}

}  // namespace air_split

// x: fp32 [M][x_ld] (C channels, C % 8 == 0, 16-byte aligned rows); out: [M][out_ld] bf16 (out_f32 == 0) or fp32, nterms * C
// channels written per row.
extern "C" int air_split_terms(const float* x, long long x_ld, long long M, int C, void* out, long long out_ld, int out_f32,
                               int nterms, unsigned lo_mask, cudaStream_t stream) {
  if (!x || !out || M <= 0 || C <= 0 || nterms < 1 || nterms > 8) return AIR_ERR_ARG;
  if (C % 8 != 0 || x_ld % 4 != 0 || out_ld % 8 != 0 || x_ld < C || out_ld < static_cast<long long>(nterms) * C) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) return AIR_ERR_UNSUPPORTED;
  const long long total = M * (C / 8);
  const int blocks = static_cast<int>(total < 256LL * 148 * 8 ? (total + 255) / 256 : 148 * 8);
  if (out_f32)
    air_split::split_terms_kernel<float><<<blocks, 256, 0, stream>>>(x, x_ld, M, C, static_cast<float*>(out), out_ld, nterms, lo_mask);
  else
    air_split::split_terms_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(x, x_ld, M, C, static_cast<__nv_bfloat16*>(out),
                                                                             out_ld, nterms, lo_mask);
  return air_launch_status();
}
