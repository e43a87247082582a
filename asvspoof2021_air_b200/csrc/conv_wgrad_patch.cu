// Weight gradient of a 3x3 / stride-1 / pad-1 convolution from shared-memory resident patches
// (TMA + tcgen05 + TMEM); the backward-weights companion of conv_patch.cu.
//
//   dW[co][i][j][ci] += sum_{b,h,w} x[b, h+i-1, w+j-1, ci] * dy[b, h, w, co]          (autograd of resnet.py:56-60)
//
// The reduction runs over pixels, so both UMMA operands are MN-major: a shared-memory row is one pixel
// (the GEMM K index) holding 64 channels (128 B, SWIZZLE_128B) -- exactly the image a TMA box load of a
// channels-last activation produces.  A work item is R = 2 output rows x 128 pixels of one image:
//   * its (R+2) x 130 pixel patch of x (64 input channels)  -> one TMA load, out-of-bounds = zero padding
//   * its R x 128 pixel tile of dy (64 output channels)     -> one TMA load, out-of-bounds = zero contribution
// and the nine taps are nine shifted windows of the x patch (descriptor start address + pixels * 128 B).
// Two taps share one M = 128 instruction: the second 64 rows of the A operand are the SAME patch shifted by the
// distance between the two taps, expressed through the descriptor's leading byte offset.  So an item costs
// R * 8 * 5 instructions of 128 x 64 x 16, the nine 64 x 64 accumulators (five M = 128 blocks, 320 TMEM columns)
// stay in TMEM for ALL items of the CTA, and each CTA adds its partial sums to the fp32 gradient once at the end.
//
// A CTA works on one (64 input channels, 64 output channels) job and a share of the items; grid = #SMs.
// Roles (256 threads): warp 0 TMA producer, warp 2 MMA issuer (warp-uniform loop, elected lane issues),
// warps 4-7 final TMEM -> red.global.add.f32.
#include <algorithm>
#include "common.cuh"
#include "tc05.cuh"
#include "tmap.cuh"

namespace air_wpatch {
using namespace tc05;

constexpr int TW = 128, PW = TW + 2, R = 2, PR = R + 2, PPIX = PR * PW;
constexpr int THREADS = 256;
constexpr int STAGES = 2;
constexpr uint32_t X_BYTES = PPIX * 128;            // 66 560 = 65 * 1024
constexpr uint32_t DY_BYTES = R * TW * 128;         // 32 768
constexpr uint32_t STAGE_BYTES = X_BYTES + DY_BYTES;
constexpr int NACC = 5;                             // tap pairs (0,1) (2,3) (4,5) (6,7) (8,-)
constexpr int ACC_COLS = 64;

struct WParams {
  int B, H, W, C, N;
  float* dw; long long dw_ld;
  int NCB, NNB, WT, HP;
  uint32_t items;                                   // B * HP * WT
  int parts;                                        // CTAs per job
};

__global__ void __launch_bounds__(THREADS, 1) conv3x3_wgrad_patch_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                         const __grid_constant__ CUtensorMap tmdy,
                                                                         const WParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                 // [STAGES] expect_tx
  uint64_t* empty = bars + STAGES;       // [STAGES] tcgen05.commit
  uint64_t* tfull = bars + 2 * STAGES;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int jobs = p.NCB * p.NNB;
  const int job = blockIdx.x % jobs, part = blockIdx.x / jobs;
  const int cb = job % p.NCB, nb = job / p.NCB;
  const bool active = part < p.parts;                // trailing CTAs (grid % jobs) have no work
  const uint32_t my_items = active ? (p.items > static_cast<uint32_t>(part) ? (p.items - part + p.parts - 1) / p.parts : 0u) : 0u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmdy);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t WT = p.WT, HP = p.HP;
      for (uint32_t k = 0; k < my_items; ++k) {
        const uint32_t item = part + k * p.parts;
        const uint32_t wt = item % WT, r1 = item / WT;
        const int w0 = static_cast<int>(wt) * TW, h0 = static_cast<int>(r1 % HP) * R, b = static_cast<int>(r1 / HP);
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t dst = sbase + stage * STAGE_BYTES;
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        tma_load_4d(dst, &tmx, cb * 64, w0 - 1, h0 - 1, b, &full[stage]);
        tma_load_4d(dst + X_BYTES, &tmdy, nb * 64, w0, h0, b, &full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    const bool leader = elect_one();
    const uint32_t idesc = instr_desc_bf16(128, ACC_COLS, 1, 1);
    // descriptor high word: SBO = 1024 B (next 8 pixels), version 1, SWIZZLE_128B
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    uint32_t stage = 0, phase = 0;
    for (uint32_t k = 0; k < my_items; ++k) {
      mbar_wait(&full[stage], phase);
      fence_after_sync();
      const uint32_t x16 = ((sbase + stage * STAGE_BYTES) >> 4) & 0x3FFF;      // 16-byte units; one pixel row = 8 units
      const uint32_t d16 = x16 + (X_BYTES >> 4);
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll 2
        for (int ks = 0; ks < TW / 16; ++ks) {
          const uint32_t b_lo = (d16 + static_cast<uint32_t>(r * TW + ks * 16) * 8) | (1u << 16);
          const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | b_lo;
          const uint32_t xrow = x16 + static_cast<uint32_t>(r * PW + ks * 16) * 8;
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            const int t0 = 2 * a, t1 = (2 * a + 1 < 9) ? 2 * a + 1 : 2 * a;
            const int o0 = (t0 / 3) * PW + (t0 % 3), o1 = (t1 / 3) * PW + (t1 % 3);
            // LBO = distance between the windows of the two taps (in 16-byte units, 8 per pixel)
            const uint32_t a_lo = (xrow + static_cast<uint32_t>(o0) * 8) | (static_cast<uint32_t>((o1 - o0) * 8) << 16);
            const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | a_lo;
            if (leader) mma_bf16(tmem_base + a * ACC_COLS, ad, bd, idesc, (k | static_cast<uint32_t>(r) | static_cast<uint32_t>(ks)) != 0);
          }
        }
      }
      if (leader) mma_commit(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (leader) mma_commit(tfull);
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------- final reduction: TMEM -> fp32 red.global.add into dW[co][tap][ci] ----------------
    if (my_items > 0) {
      mbar_wait(tfull, 0);
      fence_after_sync();
      const int q = warp & 3;
      const int ci = cb * 64 + (q & 1) * 32 + lane;
      for (int a = 0; a < NACC; ++a) {
        const int tap = 2 * a + (q >> 1);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * ACC_COLS;
        for (int c0 = 0; c0 < ACC_COLS; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (tap < 9) {
            float* dst = p.dw + static_cast<long long>(nb * 64 + c0) * p.dw_ld + static_cast<long long>(tap) * p.C + ci;
#pragma unroll
            for (int i = 0; i < 16; ++i) atomicAdd(dst + static_cast<long long>(i) * p.dw_ld, v[i]);
          }
        }
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 2) { fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace air_wpatch

using namespace air_wpatch;

extern "C" int air_conv3x3_wgrad_patch_supported(int C, int N) {
  return (C >= 64 && C % 64 == 0 && N >= 64 && N % 64 == 0 && (C / 64) * (N / 64) <= 148) ? 1 : 0;
}

// dw_out: fp32 [N][dw_ld >= 9*C] in GEMM layout [Cout][tap][Cin], accumulated in place (caller zeroes).
extern "C" int air_conv3x3_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                            const void* dy, long long dy_ld, int N,
                                            float* dw_out, long long dw_ld, int num_sms, cudaStream_t stream) {
  if (!x || !dy || !dw_out || B <= 0 || H < 1 || W < 1) return AIR_ERR_ARG;
  if (!air_conv3x3_wgrad_patch_supported(C, N)) return AIR_ERR_UNSUPPORTED;
  if (x_ld % 8 != 0 || dy_ld % 8 != 0 || dw_ld < 9LL * C) return AIR_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) return AIR_ERR_UNSUPPORTED;
  WParams p;
  p.B = B; p.H = H; p.W = W; p.C = C; p.N = N; p.dw = dw_out; p.dw_ld = dw_ld;
  p.NCB = C / 64; p.NNB = N / 64; p.WT = (W + TW - 1) / TW; p.HP = (H + R - 1) / R;
  const long long items = static_cast<long long>(B) * p.HP * p.WT;
  if (items > 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;
  p.items = static_cast<uint32_t>(items);
  if (num_sms <= 0) num_sms = 148;
  const int jobs = p.NCB * p.NNB;
  if (jobs > num_sms) return AIR_ERR_UNSUPPORTED;
  p.parts = static_cast<int>(std::min<long long>(num_sms / jobs, items));
  CUtensorMap tmx, tmdy;
  int tr = air_tmap::make_act_tmap(&tmx, x, x_ld, B, H, W, C, 64, PW, PR, 128);
  if (tr == 0) tr = air_tmap::make_act_tmap(&tmdy, dy, dy_ld, B, H, W, N, 64, TW, R, 128);
  if (tr != 0) return tr < 0 ? AIR_ERR_UNSUPPORTED : 10000 + tr;
  const size_t smem = 1024 + static_cast<size_t>(STAGES) * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  conv3x3_wgrad_patch_kernel<<<jobs * p.parts, THREADS, smem, stream>>>(tmx, tmdy, p);
  return air_launch_status();
}
