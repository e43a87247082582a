// Host-side construction of TMA tensor maps (CUtensorMap) without a link-time dependency on libcuda:
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint, so the library still loads on a
// machine without a driver (the CPU-only ABI tests) and fails loudly only when a kernel is launched.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace air_tmap {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 channels-last activation tensor [B][H][W][ld >= C] viewed as the 4-D tensor (C, W, H, B) with a
// (box_c, box_w, box_h, 1) box.  swizzle_bytes = box_c * 2 in {32, 64, 128} (the shared-memory row), or 0 (none).
// Returns 0 on success, a CUresult (> 0) or -1 otherwise.
static inline int make_act_tmap(CUtensorMap* tm, const void* base, long long ld, int B, int H, int W, int C,
                                int box_c, int box_w, int box_h, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ld) * 2 * W,
                           static_cast<cuuint64_t>(ld) * 2 * W * H};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                        : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return static_cast<int>(r);
}

// The same 4-D view with explicit element strides between W neighbours / H neighbours / images: a strided sub-image of
// a channels-last tensor (e.g. every second pixel of every second row, starting at `base`) is an ordinary tensor map.
static inline int make_act_tmap_strided(CUtensorMap* tm, const void* base, long long sw, long long sh, long long sb,
                                        int B, int H, int W, int C, int box_c, int box_w, int box_h, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(sw) * 2, static_cast<cuuint64_t>(sh) * 2, static_cast<cuuint64_t>(sb) * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                         : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                         : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return static_cast<int>(r);
}

// bf16 row-major matrix [rows][ld >= cols] viewed as the 2-D tensor (cols, rows) with a (64, box_rows) SWIZZLE_128B box:
// the A operand of a 1x1 / stride-1 convolution (rows = pixels, cols = channels) needs no gather at all.
static inline int make_mat_tmap(CUtensorMap* tm, const void* base, long long ld, long long rows, int cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return static_cast<int>(r);
}

// bf16 row-major matrix [rows][ld >= cols] as the 2-D tensor (cols, rows) with a (32 columns, 32 rows) SWIZZLE_64B box: the
// shared-memory tile one epilogue warp stores (rows / columns outside the matrix are clipped by the TMA unit).
static inline int make_out_tmap(CUtensorMap* tm, const void* base, long long ld, long long rows, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return static_cast<int>(r);
}

}  // namespace air_tmap
