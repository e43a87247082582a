// Batched weight packing: ONE launch re-packs every convolution weight of a model (fp32 master weights -> bf16
// tensor-core operand images) from a device-resident job table, instead of one small launch per tensor and layout
// (the ResNet has 70 of them per optimiser step).
//   job record = 16 x int64: [kind, src, dst, total, a0 .. a11]
//     kind 0 (generic implicit-GEMM tiles): a = N, K, KB, block_n, sn, inner, so, si
//     kind 1 (patch-kernel slices):         a = C, N, CB, taps, mode
#include <algorithm>
#include "common.cuh"
#include "pack.cuh"

namespace air_pack {

constexpr int REC = 16;

__global__ void __launch_bounds__(256) pack_jobs_kernel(const long long* __restrict__ jobs) {
  const long long* j = jobs + static_cast<long long>(blockIdx.y) * REC;
  const int kind = static_cast<int>(j[0]);
  const float* src = reinterpret_cast<const float*>(j[1]);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(j[2]);
  const long long total = j[3];
  const long long step = static_cast<long long>(gridDim.x) * blockDim.x;
  if (kind == 0) {
    const int N = static_cast<int>(j[4]), K = static_cast<int>(j[5]), KB = static_cast<int>(j[6]), bn = static_cast<int>(j[7]);
    const long long sn = j[8]; const int inner = static_cast<int>(j[9]); const long long so = j[10], si = j[11];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += step)
      gemm_pack_elem(src, dst, i, N, K, KB, bn, sn, inner, so, si);
  } else {
    const int C = static_cast<int>(j[4]), N = static_cast<int>(j[5]), CB = static_cast<int>(j[6]);
    const int taps = static_cast<int>(j[7]), mode = static_cast<int>(j[8]);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total; i += step)
      patch_pack_elem(src, dst, i, C, N, CB, taps, mode);
  }
}

}  // namespace air_pack

extern "C" int air_conv_block_n(int N);
extern "C" int air_conv3x3_patch_supported(int C, int N, int H, int W);
extern "C" int air_conv_patch_cb(int C);

// Fill one host-side job record for the generic implicit-GEMM packing (same arguments as air_conv_pack_weights_ld).
extern "C" int air_pack_job_gemm(long long* rec, const float* w, long long w_ld, void* dst, int N, int K, int mode,
                                 int Cin, int Cout, int taps) {
  if (!rec || !w || !dst || w_ld < static_cast<long long>(taps) * Cin) return AIR_ERR_ARG;
  const int bn = air_conv_block_n(N);
  if (bn == 0) return AIR_ERR_UNSUPPORTED;
  const int KB = (K + 63) / 64;
  long long sn, so, si; int inner;
  if (mode == 0) { if (N != Cout || K != taps * Cin) return AIR_ERR_ARG; sn = w_ld; inner = K; so = 0; si = 1; }
  else if (mode == 1) { if (N != Cin || K != taps * Cout) return AIR_ERR_ARG; sn = 1; inner = Cout; so = Cin; si = w_ld; }
  else return AIR_ERR_ARG;
  for (int i = 0; i < air_pack::REC; ++i) rec[i] = 0;
  rec[0] = 0; rec[1] = reinterpret_cast<long long>(w); rec[2] = reinterpret_cast<long long>(dst);
  rec[3] = static_cast<long long>(N) * KB * 64;
  rec[4] = N; rec[5] = K; rec[6] = KB; rec[7] = bn; rec[8] = sn; rec[9] = inner; rec[10] = so; rec[11] = si;
  return AIR_OK;
}

// Fill one host-side job record for the patch-kernel packing (same arguments as air_conv_patch_pack_weights).
extern "C" int air_pack_job_patch(long long* rec, const float* w, void* dst, int C, int N, int taps, int mode) {
  if (!rec || !w || !dst || !air_conv3x3_patch_supported(C, N, 1, 1) || (mode != 0 && mode != 1) || taps < 1 || taps > 9) return AIR_ERR_ARG;
  for (int i = 0; i < air_pack::REC; ++i) rec[i] = 0;
  rec[0] = 1; rec[1] = reinterpret_cast<long long>(w); rec[2] = reinterpret_cast<long long>(dst);
  rec[3] = static_cast<long long>(taps) * C * N;
  rec[4] = C; rec[5] = N; rec[6] = air_conv_patch_cb(C); rec[7] = taps; rec[8] = mode;
  return AIR_OK;
}

// jobs: DEVICE array of njobs records; max_total: the largest `total` among them (sizes the grid).
extern "C" int air_pack_jobs(const long long* jobs, int njobs, long long max_total, cudaStream_t stream) {
  if (!jobs || njobs <= 0 || max_total <= 0) return AIR_ERR_ARG;
  const int bx = static_cast<int>(std::min<long long>((max_total + 255) / 256, 64));
  dim3 grid(bx, njobs);
  air_pack::pack_jobs_kernel<<<grid, 256, 0, stream>>>(jobs);
  return air_launch_status();
}
