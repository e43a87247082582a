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

constexpr int TW = 128, PW = TW + 2, R = 2, PR = R + 2, PPIX = PR * PW;      // R: default rows per item (p.R: 2 or 3)
constexpr int NACC = 5;                             // wide 3x3: tap pairs (0,1) (2,3) (4,5) (6,7) (8,-); narrow 3x3: 3 tap rows
constexpr int THREADS = 256;
constexpr int STAGES = 2;
constexpr uint32_t X_BYTES = PPIX * 128;            // 66 560 = 65 * 1024
constexpr uint32_t DY_BYTES = R * TW * 128;         // 32 768
constexpr uint32_t STAGE_BYTES = X_BYTES + DY_BYTES;
constexpr int ACC_COLS = 64;

struct WParams {
  int B, H, W, C, N;
  float* dw; long long dw_ld;
  int NCB, NNB, WT, HP;
  uint32_t items;                                   // B * HP * WT
  int parts;                                        // CTAs per job
  // operand geometry: "wide" = 64 input channels per pixel row (SWIZZLE_128B, two 64-row groups per M = 128),
  // "narrow" = 16 input channels (SWIZZLE_32B, eight 16-row groups = eight consecutive pixel shifts per M = 128)
  int narrow, ktaps, org_h, org_w;                  // taps of the layer; patch origin relative to the item's first pixel
  int tw;                                           // pixels per row segment of an item (K of the MMAs): 128, or 96 when
                                                    // that covers W with less padding (W = 94 / 188: 73 % -> 98 % useful)
  int pw;                                           // patch width in pixels (tw + 2, or tw + 8 for dilated 1-D taps)
  int R;                                            // rows per item: 2, or 3 when H = 3 and the wider stage fits (tw = 96)
  int xr;                                           // x patch rows actually loaded: R + largest tap row offset
  int nbw;                                          // 64-channel dy images per job (1, or 2 = 128-column accumulators when <= 4 accumulators)
  uint32_t dy_img_bytes;                            // one dy image [R][tw][64 channels] in shared memory (1024-aligned)
  uint32_t x_bytes, stage_bytes;
  int nacc; int acc_off[NACC]; int acc_lbo[NACC];   // per accumulator: window offset / group distance, in patch pixels
  int acc_tap[NACC][8];                             // tap of each M row group (-1: not a real tap)
};

struct WCtx {
  const WParams& p;
  uint32_t sbase, tmem_base;
  uint64_t *full, *empty, *tfull;
  uint32_t my_items;
};

// NA = accumulators (0: run-time p.nacc), KS = K = 16 steps per row (0: run-time p.tw / 16)
template <int NA, int KS, bool NARROW>
__device__ __forceinline__ void wgrad_mma_role(const WCtx& c) {
  const WParams& p = c.p;
  const bool leader = elect_one();
  const int nacc = NA > 0 ? NA : p.nacc, ksteps = KS > 0 ? KS : p.tw / 16, rows = p.R;
  const int acc_cols = ACC_COLS * p.nbw;                 // 64, or 128: two dy images chained through the B descriptor's LBO
  const uint32_t idesc = instr_desc_bf16(128, acc_cols, 1, 1);
  // descriptor high words: SBO = next 8 pixels, version 1, swizzle mode (A: 128 B or 32 B rows; B: always 128 B rows)
  const uint32_t b_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  const uint32_t a_hi = NARROW ? ((256u >> 4) | (1u << 14) | (6u << 29)) : b_hi;
  constexpr uint32_t upp = NARROW ? 2u : 8u;              // 16-byte units per patch pixel
  // per-accumulator A descriptor without the (stage, row, K step) offset: window offset in the low word, LBO above it
  uint64_t abase[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a)
    abase[a] = (static_cast<uint64_t>(a_hi) << 32) |
               ((static_cast<uint32_t>(p.acc_off[a]) * upp) | ((static_cast<uint32_t>(p.acc_lbo[a]) * upp) << 16));
  const uint64_t bbase = (static_cast<uint64_t>(b_hi) << 32) | ((p.nbw > 1 ? (p.dy_img_bytes >> 4) : 1u) << 16);
  const uint32_t stage16 = p.stage_bytes >> 4, xb16 = p.x_bytes >> 4;
  const uint32_t xrow16 = static_cast<uint32_t>(p.pw) * upp, drow16 = static_cast<uint32_t>(p.tw) * 8u;
  const uint32_t s16 = (c.sbase >> 4) & 0x3FFF;
  uint32_t stage = 0, phase = 0;
  for (uint32_t k = 0; k < c.my_items; ++k) {
    mbar_wait(&c.full[stage], phase);
    fence_after_sync();
    const uint32_t x16 = s16 + stage * stage16;
    const uint32_t d16 = x16 + xb16;
    for (int r = 0; r < rows; ++r) {
      const uint64_t bd0 = bbase + (d16 + static_cast<uint32_t>(r) * drow16);
      const uint32_t xr = x16 + static_cast<uint32_t>(r) * xrow16;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < (KS > 0 ? KS : 8); ++ks) {
          if (KS == 0 && ks >= ksteps) break;
#pragma unroll
          for (int a = 0; a < NACC; ++a) {
            if (a < nacc)
              mma_bf16(c.tmem_base + a * acc_cols, abase[a] + (xr + static_cast<uint32_t>(ks) * 16u * upp),
                       bd0 + static_cast<uint32_t>(ks) * 128u, idesc, (k | static_cast<uint32_t>(r) | static_cast<uint32_t>(ks)) != 0);
          }
        }
      }
    }
    if (leader) mma_commit(&c.empty[stage]);
    if (++stage == STAGES) { stage = 0; phase ^= 1; }
  }
  if (leader) mma_commit(c.tfull);
}

__global__ void __launch_bounds__(THREADS, 1) conv3x3_wgrad_patch_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                         const __grid_constant__ CUtensorMap tmdy,
                                                                         const WParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + STAGES * p.stage_bytes);
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
        const int w0 = static_cast<int>(wt) * p.tw, h0 = static_cast<int>(r1 % HP) * p.R, b = static_cast<int>(r1 / HP);
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t dst = sbase + stage * p.stage_bytes;
        mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(p.xr * p.pw) * (p.narrow ? 32u : 128u) +
                                                static_cast<uint32_t>(p.nbw) * static_cast<uint32_t>(p.R * p.tw) * 128u);
        tma_load_4d(dst, &tmx, p.narrow ? 0 : cb * 64, w0 + p.org_w, h0 + p.org_h, b, &full[stage]);
        for (int j = 0; j < p.nbw; ++j)
          tma_load_4d(dst + p.x_bytes + j * p.dy_img_bytes, &tmdy, (nb * p.nbw + j) * 64, w0, h0, b, &full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 2) {
    // MMA issuer, specialised on (accumulators, K steps per row, operand geometry): straight-line descriptors
    const WCtx c{p, sbase, tmem_base, full, empty, tfull, my_items};
    const int ks = p.tw / 16;
    if (!p.narrow && p.nacc == 5 && ks == 8) wgrad_mma_role<5, 8, false>(c);
    else if (!p.narrow && p.nacc == 5 && ks == 6) wgrad_mma_role<5, 6, false>(c);
    else if (!p.narrow && p.nacc == 2 && ks == 8) wgrad_mma_role<2, 8, false>(c);
    else if (!p.narrow && p.nacc == 2 && ks == 6) wgrad_mma_role<2, 6, false>(c);
    else if (!p.narrow && p.nacc == 1 && ks == 8) wgrad_mma_role<1, 8, false>(c);
    else if (!p.narrow && p.nacc == 1 && ks == 6) wgrad_mma_role<1, 6, false>(c);
    else if (p.narrow) wgrad_mma_role<0, 0, true>(c);
    else wgrad_mma_role<0, 0, false>(c);
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------- final reduction: TMEM -> fp32 red.global.add into dW[co][tap][ci] ----------------
    if (my_items > 0) {
      mbar_wait(tfull, 0);
      fence_after_sync();
      const int q = warp & 3;
      const int m = q * 32 + lane;                       // M row of the accumulator = TMEM lane
      const int grp = p.narrow ? (m >> 4) : (m >> 6);
      const int ci = p.narrow ? (m & 15) : (cb * 64 + (m & 63));
      for (int a = 0; a < p.nacc; ++a) {
        const int tap = p.acc_tap[a][grp];
        const int acc_cols = ACC_COLS * p.nbw;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * acc_cols;
        for (int c0 = 0; c0 < acc_cols; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (tap >= 0) {
            float* dst = p.dw + static_cast<long long>(nb * acc_cols + c0) * p.dw_ld + static_cast<long long>(tap) * p.C + ci;
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

static bool wpatch_ok(int C, int N) {
  return (C == 16 || (C >= 64 && C % 64 == 0)) && N >= 64 && N % 64 == 0 && (C == 16 ? 1 : C / 64) * (N / 64) <= 148;
}

extern "C" int air_conv3x3_wgrad_patch_supported(int C, int N) { return wpatch_ok(C, N) ? 1 : 0; }

// General launcher: explicit tap list.  Tap t reads the x window that starts at patch pixel (tap_dr[t], tap_dc[t]) and
// accumulates into dw_out[co][t][ci] (GEMM layout [Cout][ntaps][Cin], fp32, caller zeroes).  Wide mode (C % 64 == 0) pairs
// consecutive taps in one M = 128 instruction; narrow mode (C == 16) needs runs of taps that are consecutive in dc.
// x view: by default the (B, H, W) grid of dy with pixel stride x_ld; `xv` describes a strided sub-image instead (stride-2
// layers: one parity class of the input).  tap_id (optional): index of tap t in the layer's [Cout][taps][Cin] gradient.
struct XView { const void* base; long long sw, sh, sb; int H, W; };

static int launch_wgrad_patch(const void* x, long long x_ld, int B, int H, int W, int C, const void* dy, long long dy_ld, int N,
                              int ntaps, const int* tap_dr, const int* tap_dc, int org_h, int org_w,
                              float* dw_out, long long dw_ld, int num_sms, cudaStream_t stream,
                              const XView* xv = nullptr, const int* tap_id = nullptr, int dw_taps = 0) {
  if (!x || !dy || !dw_out || B <= 0 || H < 1 || W < 1 || ntaps < 1 || ntaps > 9) return AIR_ERR_ARG;
  if (dw_taps <= 0) dw_taps = ntaps;
  if (!wpatch_ok(C, N)) return AIR_ERR_UNSUPPORTED;
  if (x_ld % 8 != 0 || dy_ld % 8 != 0 || dw_ld < static_cast<long long>(dw_taps) * C) return AIR_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) return AIR_ERR_UNSUPPORTED;
  WParams p;
  p.B = B; p.H = H; p.W = W; p.C = C; p.N = N; p.dw = dw_out; p.dw_ld = dw_ld;
  p.narrow = (C == 16) ? 1 : 0; p.ktaps = ntaps; p.org_h = org_h; p.org_w = org_w;
  int max_dc = 0;
  for (int t = 0; t < ntaps; ++t) {
    if (tap_dr[t] < 0 || tap_dr[t] > PR - R || tap_dc[t] < 0 || tap_dc[t] > 8) return AIR_ERR_ARG;
    max_dc = std::max(max_dc, tap_dc[t]);
  }
  p.tw = ((W + 95) / 96) * 96 < ((W + TW - 1) / TW) * TW ? 96 : TW;
  p.pw = p.tw + (max_dc <= 2 ? 2 : 8);
  // three rows per item when that wastes fewer rows (H = 3: 75 % -> 100 % useful) and two such stages fit
  p.R = R;
  if ((H + 2) / 3 * 3 < (H + 1) / 2 * 2 && p.tw == 96 && !p.narrow) p.R = 3;
  int max_dr = 0;
  for (int t = 0; t < ntaps; ++t) { if (tap_dr[t] > 2) return AIR_ERR_ARG; max_dr = std::max(max_dr, tap_dr[t]); }
  p.xr = p.R + max_dr;
  // 128-column accumulators (two dy images per job, chained through the B descriptor's leading byte offset): an M = 128,
  // N = 64 instruction reads 4 KB + 2 KB of operands for 32 cycles of tensor work, N = 128 reads 4 KB + 4 KB for 64, and the
  // x patch is loaded once per 128 output channels instead of once per 64.  Needs <= 4 accumulators (512 TMEM columns: the
  // 1 / 2 / 4-tap parity classes of the stride-2 layers, 1x1 layers) and two stages in shared memory (96-pixel tiles if needed).
  static const int wide_env = [] { const char* e = getenv("AIR_WGRAD_WIDE"); return (e && e[0] == '0') ? 0 : 1; }();
  p.nbw = 1;
  auto x_bytes_of = [&](int tw) { return (static_cast<uint32_t>(p.xr * (tw + (max_dc <= 2 ? 2 : 8))) * 128u + 1023u) / 1024u * 1024u; };
  auto dy_img_of = [&](int tw) { return (static_cast<uint32_t>(p.R * tw) * 128u + 1023u) / 1024u * 1024u; };
  auto fits2 = [&](int tw) { return 1024 + STAGES * static_cast<size_t>(x_bytes_of(tw) + 2 * dy_img_of(tw)) + 64 <= 227 * 1024; };
  if (wide_env && !p.narrow && (ntaps + 1) / 2 <= 4 && N % 128 == 0 && p.R == R) {
    if (fits2(p.tw)) p.nbw = 2;
    else if (fits2(96)) { p.nbw = 2; p.tw = 96; p.pw = p.tw + (max_dc <= 2 ? 2 : 8); }
  }
  p.NCB = p.narrow ? 1 : C / 64; p.NNB = N / (64 * p.nbw); p.WT = (W + p.tw - 1) / p.tw; p.HP = (H + p.R - 1) / p.R;
  p.x_bytes = (static_cast<uint32_t>(p.xr * p.pw) * (p.narrow ? 32u : 128u) + 1023u) / 1024u * 1024u;
  p.dy_img_bytes = p.nbw > 1 ? dy_img_of(p.tw)
                             : (p.R == R ? DY_BYTES : (static_cast<uint32_t>(p.R * p.tw) * 128u + 1023u) / 1024u * 1024u);
  p.stage_bytes = p.x_bytes + static_cast<uint32_t>(p.nbw) * p.dy_img_bytes;
  if (1024 + STAGES * static_cast<size_t>(p.stage_bytes) + 64 > 227 * 1024) return AIR_ERR_UNSUPPORTED;
  for (int a = 0; a < NACC; ++a) { p.acc_off[a] = 0; p.acc_lbo[a] = 0; for (int g = 0; g < 8; ++g) p.acc_tap[a][g] = -1; }
  int off[9];
  for (int t = 0; t < ntaps; ++t) off[t] = tap_dr[t] * p.pw + tap_dc[t];
  p.nacc = 0;
  if (p.narrow) {
    // eight 16-row groups = eight consecutive pixel shifts: one accumulator per run of taps with consecutive offsets
    int t = 0;
    while (t < ntaps) {
      if (p.nacc == NACC) return AIR_ERR_UNSUPPORTED;
      const int a = p.nacc++;
      p.acc_off[a] = off[t]; p.acc_lbo[a] = 1;
      int g = 0;
      p.acc_tap[a][g++] = t++;
      while (t < ntaps && g < 8 && off[t] == off[t - 1] + 1) p.acc_tap[a][g++] = t++;
    }
  } else {
    // two 64-row groups = two taps per instruction (the second one through the leading byte offset)
    for (int t = 0; t < ntaps; t += 2) {
      if (p.nacc == NACC) return AIR_ERR_UNSUPPORTED;
      const int a = p.nacc++;
      p.acc_off[a] = off[t]; p.acc_tap[a][0] = t;
      if (t + 1 < ntaps) {
        if (off[t + 1] < off[t]) return AIR_ERR_ARG;                 // offsets must ascend (LBO is unsigned)
        p.acc_lbo[a] = off[t + 1] - off[t]; p.acc_tap[a][1] = t + 1;
      }                                                              // else: lbo = 0, the second group duplicates the first
    }
  }
  if (tap_id) {                                                        // accumulators write the layer's real tap slots
    for (int a = 0; a < p.nacc; ++a)
      for (int g = 0; g < 8; ++g)
        if (p.acc_tap[a][g] >= 0) {
          const int id = tap_id[p.acc_tap[a][g]];
          if (id < 0 || id >= dw_taps) return AIR_ERR_ARG;
          p.acc_tap[a][g] = id;
        }
  }
  const long long items = static_cast<long long>(B) * p.HP * p.WT;
  if (items > 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;
  p.items = static_cast<uint32_t>(items);
  if (num_sms <= 0) num_sms = 148;
  const int jobs = p.NCB * p.NNB;
  if (jobs > num_sms) return AIR_ERR_UNSUPPORTED;
  p.parts = static_cast<int>(std::min<long long>(num_sms / jobs, items));
  CUtensorMap tmx, tmdy;
  int tr;
  if (xv) tr = air_tmap::make_act_tmap_strided(&tmx, xv->base, xv->sw, xv->sh, xv->sb, B, xv->H, xv->W, C, p.narrow ? 16 : 64,
                                               p.pw, p.xr, p.narrow ? 32 : 128);
  else tr = p.narrow ? air_tmap::make_act_tmap(&tmx, x, x_ld, B, H, W, C, 16, p.pw, p.xr, 32)
                     : air_tmap::make_act_tmap(&tmx, x, x_ld, B, H, W, C, 64, p.pw, p.xr, 128);
  if (tr == 0) tr = air_tmap::make_act_tmap(&tmdy, dy, dy_ld, B, H, W, N, 64, p.tw, p.R, 128);
  if (tr != 0) return tr < 0 ? AIR_ERR_DRIVER : 10000 + tr;
  const size_t smem = 1024 + static_cast<size_t>(STAGES) * p.stage_bytes + (2 * STAGES + 1) * 8 + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  conv3x3_wgrad_patch_kernel<<<jobs * p.parts, THREADS, smem, stream>>>(tmx, tmdy, p);
  return air_launch_status();
}

// dw_out: fp32 [N][dw_ld >= k*k*C] in GEMM layout [Cout][tap][Cin], accumulated in place (caller zeroes).
// k = 3: 3x3 / stride 1 / pad 1;  k = 1: 1x1 / stride 1 / pad 0.  C = 16 or a multiple of 64; N a multiple of 64.
extern "C" int air_conv_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                         const void* dy, long long dy_ld, int N, int k,
                                         float* dw_out, long long dw_ld, int num_sms, cudaStream_t stream) {
  if (k != 3 && k != 1) return AIR_ERR_ARG;
  int dr[9], dc[9];
  for (int t = 0; t < k * k; ++t) { dr[t] = t / k; dc[t] = t % k; }
  return launch_wgrad_patch(x, x_ld, B, H, W, C, dy, dy_ld, N, k * k, dr, dc, k == 3 ? -1 : 0, k == 3 ? -1 : 0,
                            dw_out, dw_ld, num_sms, stream);
}

// 1-D convolution over W (H rows are independent sequences), kernel k <= 5, dilation d, "same" padding d*(k-1)/2, with
// d*(k-1) <= 8: the dilated Conv1d of the Res2 branches (ecapa_tdnn.py:50).  dw_out [N][k][C].
extern "C" int air_conv1d_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                           const void* dy, long long dy_ld, int N, int k, int d,
                                           float* dw_out, long long dw_ld, int num_sms, cudaStream_t stream) {
  if (k < 1 || k > 5 || (k & 1) == 0 || d < 1 || d * (k - 1) > 8) return AIR_ERR_UNSUPPORTED;
  int dr[9], dc[9];
  for (int t = 0; t < k; ++t) { dr[t] = 0; dc[t] = t * d; }
  return launch_wgrad_patch(x, x_ld, B, H, W, C, dy, dy_ld, N, k, dr, dc, 0, -d * (k - 1) / 2, dw_out, dw_ld, num_sms, stream);
}

extern "C" int air_conv3x3_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                            const void* dy, long long dy_ld, int N,
                                            float* dw_out, long long dw_ld, int num_sms, cudaStream_t stream) {
  return air_conv_wgrad_patch_bf16(x, x_ld, B, H, W, C, dy, dy_ld, N, 3, dw_out, dw_ld, num_sms, stream);
}

// Weight gradient of a k x k (k = 3 / pad 1, or k = 1 / pad 0) STRIDE-2 convolution (the first convolution and the
// shortcut of a down-sampling block, resnet.py:56-60):
//     dW[co][i][j][ci] += sum_{b,ho,wo} x[b, 2 ho + i - p, 2 wo + j - p, ci] * dy[b, ho, wo, co]
// The input pixels a tap touches all share the tap's row / column parity, so the layer splits into one stride-1 problem per
// parity class of x (a strided sub-image = an ordinary TMA tensor map) with only that class's taps: 4 / 2 / 2 / 1 taps for
// 3x3, the single tap for 1x1.  No zero-stuffed gather, no im2col: x is read once, dy once per class.
//   x (B, H, W, C) with pixel stride x_ld;  dy (B, Ho, Wo, N) with pixel stride dy_ld;  dw_out fp32 [N][dw_ld >= k*k*C].
extern "C" int air_conv_s2_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                            const void* dy, long long dy_ld, int Ho, int Wo, int N, int k,
                                            float* dw_out, long long dw_ld, int num_sms, cudaStream_t stream) {
  if (k != 3 && k != 1) return AIR_ERR_UNSUPPORTED;
  if (!x || !dy || !dw_out || B <= 0 || H < 1 || W < 1) return AIR_ERR_ARG;
  const int pad = k == 3 ? 1 : 0;
  if (Ho != (H + 2 * pad - k) / 2 + 1 || Wo != (W + 2 * pad - k) / 2 + 1) return AIR_ERR_ARG;
  if (C == 16) return AIR_ERR_UNSUPPORTED;                             // narrow mode needs runs of adjacent taps
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  for (int pr = 0; pr < 2; ++pr) {
    for (int pc = 0; pc < 2; ++pc) {
      // taps (i, j) whose input row 2 ho + i - pad has parity pr (columns likewise); sub-image row of tap i for output
      // row ho: (2 ho + i - pad - pr) / 2 = ho + (i - pad - pr) / 2  ->  patch origin -1 when some tap reaches back
      int dr[9], dc[9], id[9], nt = 0, org_h = 0, org_w = 0;
      for (int i = 0; i < k; ++i) if (((i - pad - pr) & 1) == 0 && (i - pad - pr) / 2 < 0) org_h = -1;
      for (int j = 0; j < k; ++j) if (((j - pad - pc) & 1) == 0 && (j - pad - pc) / 2 < 0) org_w = -1;
      for (int i = 0; i < k; ++i) {
        if ((i - pad - pr) & 1) continue;
        for (int j = 0; j < k; ++j) {
          if ((j - pad - pc) & 1) continue;
          dr[nt] = (i - pad - pr) / 2 - org_h; dc[nt] = (j - pad - pc) / 2 - org_w; id[nt] = i * k + j;
          ++nt;
        }
      }
      if (nt == 0) continue;
      const int Hs = (H - pr + 1) / 2, Ws = (W - pc + 1) / 2;             // rows / columns of this parity class
      if (Hs < 1 || Ws < 1) continue;
      XView xv{xb + (static_cast<long long>(pr) * W + pc) * x_ld, 2 * x_ld, 2 * static_cast<long long>(W) * x_ld,
               static_cast<long long>(H) * W * x_ld, Hs, Ws};
      const int st = launch_wgrad_patch(x, x_ld, B, Ho, Wo, C, dy, dy_ld, N, nt, dr, dc, org_h, org_w, dw_out, dw_ld, num_sms,
                                        stream, &xv, id, k * k);
      if (st != 0) return st;
    }
  }
  return AIR_OK;
}
