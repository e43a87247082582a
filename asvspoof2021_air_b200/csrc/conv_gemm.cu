// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Replaces the cuDNN calls behind nn.Conv2d (resnet.py:56-60,131,140) and nn.Conv1d
// (ecapa_tdnn.py:19-23,39,50,56,111-118,139-145) for forward, data-gradient and weight-gradient.
// Activations are channels-last bf16 ([B][H][W][C], 1-D convs have H = 1); weights are kept in
// GEMM layout [Cout][kh][kw][Cin].
//
//   fprop : out[m, n]  = sum_k  A[m, k] * Wp[n, k]      m = (b, ho, wo), k = (tap, ci)
//   dgrad : dx [m, n]  = sum_k  A'[m, k] * Wd[n, k]     m = (b, h, w),   k = (tap, co), n = ci
//   wgrad : dW [n, k] += sum_m  dy[m, n] * A[m, k]      (conv_wgrad.cu)
//
// A is never materialised: 128 producer threads gather one 128-byte im2col row each per K block
// with zero-filling 16-byte cp.async, straight into the UMMA "column of rows" layout (tc05.cuh).
// Pre-packed weight tiles arrive with one bulk copy per K block on the TMA engine.  One elected
// thread issues tcgen05.mma (M = 128, N = block_n <= 256, K = 16 x 4 per stage); accumulators are
// double-buffered in TMEM so the epilogue warps (TMEM -> registers -> bias/residual/ReLU -> bf16 ->
// HBM) overlap the next tile's main loop.  The kernel is persistent: grid = #SMs, tiles strided.
#include <algorithm>
#include "common.cuh"
#include "tc05.cuh"
#include "pack.cuh"

namespace air_gemm {
using namespace tc05;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB
constexpr int A_CHUNK_STRIDE = BLOCK_M * 16;              // bytes between 8-element K chunks
constexpr int NUM_A_THREADS = 128;
constexpr int THREADS = 320;                              // 4 gather warps, TMA warp, MMA warp, 4 epilogue warps

struct ConvParams {
  const __nv_bfloat16* a; long long a_ld;   // gather source, elements between consecutive pixels
  int H, W, C;                              // gather-source geometry, C channels per tap (multiple of 8)
  int Ho, Wo;                               // pixel grid enumerated by m
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int mode;                                 // 0: fprop gather; 1: transposed gather (dgrad)
  const __nv_bfloat16* wpk;                 // packed weights [n_tile][k_block][8 chunks][block_n][8]
  int N, K, KB, block_n, n_tiles;
  long long M; int m_tiles;
  __nv_bfloat16* out; long long out_ld;
  const float* bias; const __nv_bfloat16* res; long long res_ld; int relu;
  int bias_rows;                            // 0: bias[n]; > 0: bias[(m / bias_rows) * N + n] (per-utterance bias)
  __nv_bfloat16* out2; long long out2_ld;   // optional second output: the accumulator (+bias) WITHOUT the residual
  int stages; int flags;
};

__global__ void __launch_bounds__(THREADS, 1) conv_gemm_kernel(const ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.block_n) * BLOCK_K * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + S * b_stage_bytes);
  uint64_t* full = bars;               // [S]  128 gather arrivals + 1 expect_tx arrival
  uint64_t* empty = bars + S;          // [S]  tcgen05.commit
  uint64_t* tfull = bars + 2 * S;      // [2]  accumulator ready
  uint64_t* tempty = bars + 2 * S + 2; // [2]  accumulator drained (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  uint32_t ncols = 32;
  while (ncols < 2u * p.block_n) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], NUM_A_THREADS + 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp < 4) {
    // ===================== A gather producers: one im2col row per thread =====================
    const int r = threadIdx.x;
    uint32_t stage = 0, phase = 0;
    const uint32_t dst_row = smem_u32(sA) + r * 16;
    const bool fast = (p.C % BLOCK_K) == 0;          // a K block never straddles two taps
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      const long long m = static_cast<long long>(m_tile) * BLOCK_M + r;
      const bool row_ok = m < p.M;
      int wo = 0, ho = 0, bb = 0;
      if (row_ok) { wo = static_cast<int>(m % p.Wo); const long long t = m / p.Wo; ho = static_cast<int>(t % p.Ho); bb = static_cast<int>(t / p.Ho); }
      const int hbase = p.mode == 0 ? ho * p.sh - p.ph : ho + p.ph;
      const int wbase = p.mode == 0 ? wo * p.sw - p.pw : wo + p.pw;
      const __nv_bfloat16* img = p.a + static_cast<long long>(bb) * p.H * p.W * p.a_ld;
      int tap = 0, ci0 = 0;                          // fast path running position
      for (int kb = 0; kb < p.KB; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t dst = dst_row + stage * A_STAGE_BYTES;
        if (fast) {
          const int khi = tap / p.kw, kwi = tap - khi * p.kw;
          int hi, wi; bool ok = row_ok && (kb * BLOCK_K < p.K);
          if (p.mode == 0) { hi = hbase + khi * p.dh; wi = wbase + kwi * p.dw; }
          else {
            const int hn = hbase - khi * p.dh, wn = wbase - kwi * p.dw;
            ok = ok && hn >= 0 && wn >= 0 && (hn % p.sh) == 0 && (wn % p.sw) == 0;
            hi = hn / p.sh; wi = wn / p.sw;
          }
          ok = ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
          const __nv_bfloat16* src = ok ? img + (static_cast<long long>(hi) * p.W + wi) * p.a_ld + ci0 : p.a;
          const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
          for (int c = 0; c < 8; ++c) cp_async16(dst + c * A_CHUNK_STRIDE, src + (ok ? c * 8 : 0), nb);
          ci0 += BLOCK_K;
          if (ci0 >= p.C) { ci0 = 0; ++tap; }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int k = kb * BLOCK_K + c * 8;
            const int tp = k / p.C, ci = k - tp * p.C;
            const int khi = tp / p.kw, kwi = tp - khi * p.kw;
            int hi, wi; bool ok = row_ok && k < p.K;
            if (p.mode == 0) { hi = hbase + khi * p.dh; wi = wbase + kwi * p.dw; }
            else {
              const int hn = hbase - khi * p.dh, wn = wbase - kwi * p.dw;
              ok = ok && hn >= 0 && wn >= 0 && (hn % p.sh) == 0 && (wn % p.sw) == 0;
              hi = hn / p.sh; wi = wn / p.sw;
            }
            ok = ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const __nv_bfloat16* src = ok ? img + (static_cast<long long>(hi) * p.W + wi) * p.a_ld + ci : p.a;
            cp_async16(dst + c * A_CHUNK_STRIDE, src, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(&full[stage]);
        if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ===================== weight tiles: one bulk copy per K block =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const __nv_bfloat16* wsrc = p.wpk + static_cast<long long>(n_tile) * p.KB * p.block_n * BLOCK_K;
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], b_stage_bytes);
          bulk_g2s(smem_u32(sB) + stage * b_stage_bytes, wsrc + static_cast<long long>(kb) * p.block_n * BLOCK_K,
                   b_stage_bytes, &full[stage]);
          if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer (warp-uniform loop: descriptors in uniform registers, elected lane issues) =====================
    const bool leader = elect_one();
    const uint32_t idesc = instr_desc_bf16(BLOCK_M, p.block_n, 0, 0);
    const uint32_t b_chunk = static_cast<uint32_t>(p.block_n) * 16;
    uint32_t a_lbo = A_CHUNK_STRIDE, a_sbo = 128, b_lbo = b_chunk, b_sbo = 128;
    if (p.flags & 1) { a_lbo = 128; a_sbo = A_CHUNK_STRIDE; b_lbo = 128; b_sbo = b_chunk; }   // debug: swapped roles
    const uint32_t a_hi = (a_sbo >> 4) | (1u << 14), b_hi = (b_sbo >> 4) | (1u << 14);
    const uint32_t a_lbo16 = (a_lbo >> 4) << 16, b_lbo16 = (b_lbo >> 4) << 16;
    uint32_t stage = 0, phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * p.block_n;
      for (int kb = 0; kb < p.KB; ++kb) {
        mbar_wait(&full[stage], phase);
        fence_after_sync();
        const uint32_t a0 = (((smem_u32(sA) + stage * A_STAGE_BYTES) >> 4) & 0x3FFF) | a_lbo16;
        const uint32_t b0 = (((smem_u32(sB) + stage * b_stage_bytes) >> 4) & 0x3FFF) | b_lbo16;
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
          const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) | (a0 + kk * ((2 * A_CHUNK_STRIDE) >> 4));
          const uint64_t bd = (static_cast<uint64_t>(b_hi) << 32) | (b0 + kk * ((2 * b_chunk) >> 4));
          if (leader) mma_bf16(d_tmem, ad, bd, idesc, (kb | kk) != 0);
        }
        if (leader) mma_commit(&empty[stage]);
        if (kb == p.KB - 1 && leader) mma_commit(&tfull[acc]);
        if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_tile = tile / p.n_tiles, n_tile = tile % p.n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      fence_after_sync();
      const long long m = static_cast<long long>(m_tile) * BLOCK_M + row;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (m < p.M) {
          const int n0 = n_tile * p.block_n + c0;
          if (p.bias) {
            const float* bp = p.bias + n0 + (p.bias_rows > 0 ? (m / p.bias_rows) * p.N : 0);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += __ldg(bp + i);
          }
          if (p.out2) {
            bf16x8* o2 = reinterpret_cast<bf16x8*>(p.out2 + m * p.out2_ld + n0);
            o2[0] = pack8(v);
            o2[1] = pack8(v + 8);
          }
          if (p.res) {
            const bf16x8* rp = reinterpret_cast<const bf16x8*>(p.res + m * p.res_ld + n0);
            float rf[16];
            unpack8(rp[0], rf); unpack8(rp[1], rf + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rf[i];
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          bf16x8* op = reinterpret_cast<bf16x8*>(p.out + m * p.out_ld + n0);
          op[0] = pack8(v);
          op[1] = pack8(v + 8);
        }
      }
      fence_before_sync();
      mbar_arrive(&tempty[acc]);
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 5) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
}

// ------------------------------------------------------------------------------------------
// Weight packing: fp32 master weights in GEMM layout -> bf16 UMMA tiles.
//   value(n, k) = src[n*sn + (k / inner)*so + (k % inner)*si]
//   fprop: n = co, k = tap*Cin + ci : sn = K, inner = K, so = 0, si = 1
//   dgrad: n = ci, k = tap*Cout + co: sn = 1, inner = Cout, so = Cin, si = taps*Cin
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                    int N, int K, int KB, int block_n, long long sn, int inner, long long so, long long si) {
  // (for a column slice of a wider matrix, e.g. attention.0's W_x = W[:, :1536], sn is the full row stride)
  const long long total = static_cast<long long>(N) * KB * BLOCK_K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    air_pack::gemm_pack_elem(src, dst, i, N, K, KB, block_n, sn, inner, so, si);
}

static int pick_block_n(int N) {
  if (N <= 256) return N;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

}  // namespace air_gemm

using namespace air_gemm;

extern "C" int air_conv_block_n(int N) { return (N % 16 == 0) ? pick_block_n(N) : 0; }

// number of bf16 elements of the packed weight buffer for an (N, K) GEMM
extern "C" long long air_conv_packed_elems(int N, int K) {
  const int KB = (K + BLOCK_K - 1) / BLOCK_K;
  return static_cast<long long>(N) * KB * BLOCK_K;
}

extern "C" int air_conv_pack_weights_ld(const float* w, long long w_ld, void* dst, int N, int K, int mode, int Cin, int Cout,
                                        int taps, cudaStream_t stream);

extern "C" int air_conv_pack_weights(const float* w, void* dst, int N, int K, int mode, int Cin, int Cout, int taps,
                                     cudaStream_t stream) {
  return air_conv_pack_weights_ld(w, static_cast<long long>(taps) * Cin, dst, N, K, mode, Cin, Cout, taps, stream);
}

// w_ld: elements between consecutive output-channel rows of w (== taps*Cin for a dense tensor)
extern "C" int air_conv_pack_weights_ld(const float* w, long long w_ld, void* dst, int N, int K, int mode, int Cin, int Cout,
                                        int taps, cudaStream_t stream) {
  // mode 0: fprop pack of w[Cout][taps][Cin] (N = Cout, K = taps*Cin)
  // mode 1: dgrad pack of the same tensor      (N = Cin,  K = taps*Cout)
  if (!w || !dst || w_ld < static_cast<long long>(taps) * Cin) return AIR_ERR_ARG;
  const int bn = air_conv_block_n(N);
  if (bn == 0) return AIR_ERR_UNSUPPORTED;
  const int KB = (K + BLOCK_K - 1) / BLOCK_K;
  long long sn, so, si; int inner;
  if (mode == 0) { if (N != Cout || K != taps * Cin) return AIR_ERR_ARG; sn = w_ld; inner = K; so = 0; si = 1; }
  else if (mode == 1) { if (N != Cin || K != taps * Cout) return AIR_ERR_ARG; sn = 1; inner = Cout; so = Cin; si = w_ld; }
  else return AIR_ERR_ARG;
  const long long total = static_cast<long long>(N) * KB * BLOCK_K;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 2048));
  pack_weights_kernel<<<blocks, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(dst), N, K, KB, bn, sn, inner, so, si);
  return air_launch_status();
}

extern "C" int air_conv_gemm_bf16_ex(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                     int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                     const void* wpk, int N, int K, void* out, long long out_ld,
                                     const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                                     void* out2, long long out2_ld, int num_sms, int flags, cudaStream_t stream);

// Generic implicit-GEMM launch (fprop when mode == 0, dgrad when mode == 1).
extern "C" int air_conv_gemm_bf16(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                  int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                  const void* wpk, int N, int K, void* out, long long out_ld,
                                  const float* bias, const void* res, long long res_ld, int relu,
                                  int num_sms, int flags, cudaStream_t stream) {
  return air_conv_gemm_bf16_ex(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K, out, out_ld,
                               bias, 0, res, res_ld, relu, nullptr, 0, num_sms, flags, stream);
}

extern "C" int air_conv_gemm_bf16_ex(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                     int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                     const void* wpk, int N, int K, void* out, long long out_ld,
                                     const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                                     void* out2, long long out2_ld, int num_sms, int flags, cudaStream_t stream) {
  if (!a || !wpk || !out || B <= 0) return AIR_ERR_ARG;
  if (C % 8 != 0 || a_ld % 8 != 0 || out_ld % 8 != 0 || (res && res_ld % 8 != 0) || (out2 && out2_ld % 8 != 0)) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out2) & 15) || bias_rows < 0) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(wpk) |
       reinterpret_cast<uintptr_t>(res)) & 15) return AIR_ERR_UNSUPPORTED;
  const int bn = air_conv_block_n(N);
  if (bn == 0) return AIR_ERR_UNSUPPORTED;
  ConvParams p;
  p.a = reinterpret_cast<const __nv_bfloat16*>(a); p.a_ld = a_ld; p.H = H; p.W = W; p.C = C; p.Ho = Ho; p.Wo = Wo;
  p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw; p.dh = dh; p.dw = dw; p.mode = mode;
  p.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk); p.N = N; p.K = K; p.KB = (K + BLOCK_K - 1) / BLOCK_K;
  p.block_n = bn; p.n_tiles = N / bn;
  p.M = static_cast<long long>(B) * Ho * Wo; p.m_tiles = static_cast<int>((p.M + BLOCK_M - 1) / BLOCK_M);
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_ld = out_ld; p.bias = bias;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.res_ld = res_ld; p.relu = relu; p.flags = flags;
  p.bias_rows = bias_rows; p.out2 = reinterpret_cast<__nv_bfloat16*>(out2); p.out2_ld = out2_ld;
  const int stage_bytes = A_STAGE_BYTES + bn * BLOCK_K * 2;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return AIR_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + (2 * stages + 4) * 8 + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  conv_gemm_kernel<<<grid, THREADS, smem, stream>>>(p);
  return air_launch_status();
}
