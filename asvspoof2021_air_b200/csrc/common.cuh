// Shared device/host helpers for the air_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "air_b200.h"   // status codes; and every extern "C" definition is checked against its public declaration

// Every non-zero status leaves a per-thread record (status, source line) behind for air_last_error_string() (capi.cu):
// inside the kernel sources the two argument-error codes expand to a call that notes where they were returned.
extern "C" int air_internal_note_status(int status, const char* file, int line);
#undef AIR_ERR_ARG
#define AIR_ERR_ARG air_internal_note_status(-1, __FILE__, __LINE__)
#undef AIR_ERR_UNSUPPORTED
#define AIR_ERR_UNSUPPORTED air_internal_note_status(-2, __FILE__, __LINE__)
#undef AIR_ERR_DRIVER
#define AIR_ERR_DRIVER air_internal_note_status(-7, __FILE__, __LINE__)

// Launch-error check used by every entry point: asynchronous, no device sync.
static inline int air_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? AIR_OK : air_internal_note_status((int)e, nullptr, 0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < nw) ? red[l] : -INFINITY;
  r = warp_max(r);
  return r;
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }

// 8 bf16 <-> 8 floats through ONE 16-byte vector access.  The payload is a uint4 so that loads / stores of a
// bf16x8 compile to LDG.128 / STG.128 (a struct of four __nv_bfloat162 is split into 32-bit accesses).
struct __align__(16) bf16x8 { uint4 u; };
__device__ __forceinline__ float2 bf2x_to_f2(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f2_to_bf2x(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
  float2 t;
  t = bf2x_to_f2(p.u.x); f[0] = t.x; f[1] = t.y;
  t = bf2x_to_f2(p.u.y); f[2] = t.x; f[3] = t.y;
  t = bf2x_to_f2(p.u.z); f[4] = t.x; f[5] = t.y;
  t = bf2x_to_f2(p.u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
  p.u.x = f2_to_bf2x(f[0], f[1]); p.u.y = f2_to_bf2x(f[2], f[3]);
  p.u.z = f2_to_bf2x(f[4], f[5]); p.u.w = f2_to_bf2x(f[6], f[7]);
  return p;
}

// ---------------------------------------------------------------------------------------------------------------
// Activation storage type T = __nv_bfloat16 (the product path) or float (the fp32 parity mode):
// the HBM-bound kernels are templates over T and move 8 channels per access either way (16 B of bf16, 2 x 16 B of fp32).
// ---------------------------------------------------------------------------------------------------------------
struct __align__(16) f32x8 { float4 a, b; };
template <typename T> struct V8sel;
template <> struct V8sel<__nv_bfloat16> { using type = bf16x8; };
template <> struct V8sel<float> { using type = f32x8; };
template <typename T> using V8 = typename V8sel<T>::type;

__device__ __forceinline__ void unpack8(const f32x8& p, float* f) {
  f[0] = p.a.x; f[1] = p.a.y; f[2] = p.a.z; f[3] = p.a.w; f[4] = p.b.x; f[5] = p.b.y; f[6] = p.b.z; f[7] = p.b.w;
}
__device__ __forceinline__ bf16x8 ldv8(const __nv_bfloat16* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ f32x8 ldv8(const float* p) {
  f32x8 v;
  v.a = reinterpret_cast<const float4*>(p)[0]; v.b = reinterpret_cast<const float4*>(p)[1];
  return v;
}
__device__ __forceinline__ void packv(const float* f, bf16x8& out) { out = pack8(f); }
__device__ __forceinline__ void packv(const float* f, f32x8& out) {
  out.a = make_float4(f[0], f[1], f[2], f[3]); out.b = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void stv8(__nv_bfloat16* p, const bf16x8& v) { *reinterpret_cast<bf16x8*>(p) = v; }
__device__ __forceinline__ void stv8(float* p, const f32x8& v) {
  reinterpret_cast<float4*>(p)[0] = v.a; reinterpret_cast<float4*>(p)[1] = v.b;
}
// 8 floats -> storage (rounded for bf16)
template <typename T> __device__ __forceinline__ void st8(T* p, const float* f) { V8<T> v; packv(f, v); stv8(p, v); }
// 8 stored values -> floats
template <typename T> __device__ __forceinline__ void ld8(const T* p, float* f) { unpack8(ldv8(p), f); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
