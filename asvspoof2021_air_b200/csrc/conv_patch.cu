// 3x3 / stride-1 / pad-1 convolution with a shared-memory resident input patch (TMA + tcgen05 + TMEM).
//
// The generic implicit-GEMM kernel (conv_gemm.cu) gathers every activation once per tap (9x) through the
// LSU, which makes the N = 64 / 128 layers of the ResNet (resnet.py:56-60, layer1 / layer2) gather bound.
// Here a work item is R = 2 output rows x 128 pixels of one image; its (R+2) x 130 pixel input patch
// (one block of <= 64 channels at a time) is brought into shared memory ONCE by a single TMA box load
// (out-of-bounds rows / columns are zero filled by the TMA unit = the padding of the convolution) and
// the nine taps are nine shifted windows of it: the patch image is "one swizzled row per pixel", so moving
// the UMMA operand by one pixel is +row_bytes on the descriptor start address, by one image row +130 rows.
//
//   out[b, h, w, n] = sum_{i,j,c} a[b, h+i-1, w+j-1, c] * Wp[n][i][j][c]   (+ res) (ReLU)
//
// Forward uses Wp = W; the data gradient of the same layer is the same kernel on dy with the taps
// flipped and (ci, co) swapped, which is done once in the weight packing (mode 1).
// Roles (256 threads): warp 0 issues the patch TMA loads, warp 1 streams pre-swizzled weight slices (one
// bulk copy per (channel block, tap); all slices stay resident when they fit), warp 2 issues tcgen05.mma
// (M = 128 pixels, N = Cout, K = 16 per instruction), warps 4-7 drain the double-buffered TMEM accumulators.
// Persistent grid.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "tc05.cuh"
#include "tmap.cuh"

namespace air_patch {
using namespace tc05;

constexpr int TW = 128;                 // output pixels per row segment (UMMA M)
constexpr int PW = TW + 2;              // patch width
constexpr int R = 2;                    // output rows per work item
constexpr int PR = R + 2;               // patch rows
constexpr int PPIX = PR * PW;           // 520 patch pixels
constexpr int THREADS = 256;
constexpr int PSTAGES = 2;

struct PatchParams {
  int B, H, W, C;
  const __nv_bfloat16* wpk; int N;
  __nv_bfloat16* out; long long out_ld;
  const __nv_bfloat16* res; long long res_ld; int relu;
  int CB, NCB, WT, HP;                  // channels per block (16 / 32 / 64), #blocks, W tiles, row pairs
  int row_bytes, layout;                // CB * 2; UMMA layout type (6 / 4 / 2)
  uint32_t pstage_bytes, bslot_bytes, bslot_stride;
  int nb_slots, resident;               // weight ring
  int acc_stages;                       // 1 or 2
  int dbg;                              // AIR_PATCH_DBG: 1 no stores, 2 no MMAs, 4 no patch loads (timing experiments only)
  long long items;
};

__global__ void __launch_bounds__(THREADS, 1) conv3x3_patch_kernel(const __grid_constant__ CUtensorMap tmap, const PatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle patterns are anchored at 1024 B
  const uint32_t sP = sbase;
  const uint32_t sB = sbase + PSTAGES * p.pstage_bytes;
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + PSTAGES * p.pstage_bytes + static_cast<size_t>(p.nb_slots) * p.bslot_stride);
  uint64_t* full_p = bars;                          // [PSTAGES] expect_tx (TMA)
  uint64_t* empty_p = bars + PSTAGES;               // [PSTAGES] tcgen05.commit
  uint64_t* full_b = bars + 2 * PSTAGES;            // [nb_slots] expect_tx
  uint64_t* empty_b = full_b + p.nb_slots;          // [nb_slots] tcgen05.commit
  uint64_t* tfull = empty_b + p.nb_slots;           // [2]
  uint64_t* tempty = tfull + 2;                     // [2] 128 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(p.acc_stages * R * p.N)) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full_p[s], 1); mbar_init(&empty_p[s], 1); }
    for (int s = 0; s < p.nb_slots; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
  }
  if (warp == 2) tmem_alloc(tmem_slot, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nslices = p.NCB * 9;

  if (warp == 0) {
    // ===================== patch loads: one TMA box per (item, channel block) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const uint32_t items32 = static_cast<uint32_t>(p.items), WT = p.WT, HP = p.HP;
      for (uint32_t item = blockIdx.x; item < items32; item += gridDim.x) {
        const uint32_t wt = item % WT, r1 = item / WT;            // coordinates are ready BEFORE the slot frees up
        const int c1 = static_cast<int>(wt) * TW - 1, c2 = static_cast<int>(r1 % HP) * R - 1, c3 = static_cast<int>(r1 / HP);
        for (int cb = 0; cb < p.NCB; ++cb) {
          mbar_wait(&empty_p[stage], phase ^ 1);
          if (p.dbg & 4) { mbar_arrive(&full_p[stage]); if (++stage == PSTAGES) { stage = 0; phase ^= 1; } continue; }
          mbar_arrive_expect_tx(&full_p[stage], static_cast<uint32_t>(PPIX) * p.row_bytes);
          tma_load_4d(sP + stage * p.pstage_bytes, &tmap, cb * p.CB, c1, c2, c3, &full_p[stage]);
          if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== weight slices =====================
    if (lane == 0) {
      const long long slice_elems = static_cast<long long>(p.N) * p.CB;
      if (p.resident) {
        for (int s = 0; s < nslices; ++s) {
          mbar_arrive_expect_tx(&full_b[s], p.bslot_bytes);
          bulk_g2s(sB + s * p.bslot_stride, p.wpk + s * slice_elems, p.bslot_bytes, &full_b[s]);
        }
      } else {
        uint32_t slot = 0, phase = 0;
        for (uint32_t item = blockIdx.x; item < static_cast<uint32_t>(p.items); item += gridDim.x) {
          for (int s = 0; s < nslices; ++s) {
            mbar_wait(&empty_b[slot], phase ^ 1);
            mbar_arrive_expect_tx(&full_b[slot], p.bslot_bytes);
            bulk_g2s(sB + slot * p.bslot_stride, p.wpk + s * slice_elems, p.bslot_bytes, &full_b[slot]);
            if (++slot == static_cast<uint32_t>(p.nb_slots)) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs the (warp-uniform) loop so that descriptors live in uniform registers; one elected
    // lane issues tcgen05.mma / tcgen05.commit.
    const bool leader = elect_one();
    const uint32_t idesc = instr_desc_bf16(TW, p.N, 0, 0);
    const uint32_t rb16 = static_cast<uint32_t>(p.row_bytes) >> 4;                  // row stride in 16-byte units
    const uint32_t desc_hi = ((8u * p.row_bytes) >> 4) | (1u << 14) | (static_cast<uint32_t>(p.layout) << 29);
    const int KK = p.CB >> 4;
    const uint32_t items32 = static_cast<uint32_t>(p.items);
    uint32_t pstage = 0, pphase = 0, slot = 0, bphase = 0;
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < items32; item += gridDim.x, ++it) {
      const uint32_t hp = (item / p.WT) % p.HP;
      const int rows = min(R, p.H - static_cast<int>(hp) * R);
      const uint32_t acc = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      fence_after_sync();
      const uint32_t d0 = tmem_base + acc * R * p.N;
      for (int cb = 0; cb < p.NCB; ++cb) {
        mbar_wait(&full_p[pstage], pphase);
        fence_after_sync();
        const uint32_t a_lo = (((sP + pstage * p.pstage_bytes) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ti = tap / 3, tj = tap - ti * 3;
          uint32_t b0;
          if (p.resident) {
            const int s = cb * 9 + tap;
            if (it == 0) { mbar_wait(&full_b[s], 0); fence_after_sync(); }
            b0 = sB + s * p.bslot_stride;
          } else {
            mbar_wait(&full_b[slot], bphase);
            fence_after_sync();
            b0 = sB + slot * p.bslot_stride;
          }
          const uint32_t b_lo = ((b0 >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
          for (int j = 0; j < R; ++j) {
            if (j < rows) {
              const uint32_t arow = a_lo + static_cast<uint32_t>((j + ti) * PW + tj) * rb16;
              for (int kk = 0; kk < KK; ++kk) {
                const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | (arow + kk * 2);
                const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + kk * 2);
                if (leader && !(p.dbg & 2)) mma_bf16(d0 + j * p.N, ad, bd, idesc, (cb | tap | kk) != 0);
              }
            }
          }
          if (!p.resident) {
            if (leader) mma_commit(&empty_b[slot]);
            if (++slot == static_cast<uint32_t>(p.nb_slots)) { slot = 0; bphase ^= 1; }
          }
        }
        if (leader) mma_commit(&empty_p[pstage]);
        if (++pstage == PSTAGES) { pstage = 0; pphase ^= 1; }
      }
      if (leader) mma_commit(&tfull[acc]);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int it = 0;
    const uint32_t items32 = static_cast<uint32_t>(p.items);
    for (uint32_t item = blockIdx.x; item < items32; item += gridDim.x, ++it) {
      const int wt = static_cast<int>(item % static_cast<uint32_t>(p.WT));
      const uint32_t r1 = item / static_cast<uint32_t>(p.WT);
      const int hp = static_cast<int>(r1 % static_cast<uint32_t>(p.HP)), b = static_cast<int>(r1 / static_cast<uint32_t>(p.HP));
      const int rows = min(R, p.H - hp * R);
      const int acc = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(&tfull[acc], acc_phase);
      fence_after_sync();
      const int w = wt * TW + m;
      for (int j = 0; j < ((p.dbg & 16) ? 0 : rows); ++j) {
        const long long pixel = (static_cast<long long>(b) * p.H + hp * R + j) * p.W + w;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (acc * R + j) * p.N;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          float v[16];
          if (!(p.dbg & 8)) tmem_ld16(taddr + c0, v);
          if (w < p.W && !(p.dbg & 1)) {
            if (p.res) {
              const bf16x8* rp = reinterpret_cast<const bf16x8*>(p.res + pixel * p.res_ld + c0);
              float rf[16];
              unpack8(rp[0], rf); unpack8(rp[1], rf + 8);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += rf[i];
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            bf16x8* op = reinterpret_cast<bf16x8*>(p.out + pixel * p.out_ld + c0);
            op[0] = pack8(v);
            op[1] = pack8(v + 8);
          }
        }
      }
      fence_before_sync();
      mbar_arrive(&tempty[acc]);
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 2) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
}

// Weight packing for the patch kernel: one pre-swizzled [N rows][CB channels] K-major image per (cb, tap), exactly
// the bytes a TMA load with the matching swizzle would have produced, so one bulk copy per slice suffices:
//   element (n, k) of slice (cb, tap) at byte  swizzle(n * CB*2 + (k / 8) * 16) + (k % 8) * 2
//   mode 0 (fprop): w is [N = Cout][9][C = Cin];   value = w[n][tap][ch]
//   mode 1 (dgrad): w is [C = Cout][9][N = Cin];   value = w[ch][8 - tap][n]   (flipped taps, swapped channels)
__global__ void pack3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int C, int N, int CB, int mode) {
  const long long total = 9LL * C * N;
  const uint32_t mask = CB == 64 ? 7u : (CB == 32 ? 3u : 1u);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % CB);
    long long t = i / CB;
    const int n = static_cast<int>(t % N); t /= N;
    const int tap = static_cast<int>(t % 9);
    const int cb = static_cast<int>(t / 9);
    const int ch = cb * CB + k;
    const float v = mode == 0 ? w[(static_cast<long long>(n) * 9 + tap) * C + ch]
                              : w[(static_cast<long long>(ch) * 9 + (8 - tap)) * N + n];
    const uint32_t off = swizzle_offset(static_cast<uint32_t>(n) * CB * 2 + (k >> 3) * 16, mask) + (k & 7) * 2;
    dst[(static_cast<long long>(cb) * 9 + tap) * N * CB + (off >> 1)] = f2bf(v);
  }
}

}  // namespace air_patch

using namespace air_patch;

static int patch_cb(int C) { return C <= 64 ? C : 64; }

// 1 when (C, N, W) can run on the patch kernel
extern "C" int air_conv3x3_patch_supported(int C, int N, int H, int W) {
  if (!(C == 16 || C == 32 || (C >= 64 && C % 64 == 0))) return 0;
  if (N % 16 != 0 || N > 256 || N < 16) return 0;
  if ((N * patch_cb(C) * 2) % 256 != 0) return 0;
  if (H < 1 || W < 1) return 0;
  return 1;
}

extern "C" int air_conv3x3_pack_weights(const float* w, void* dst, int C, int N, int mode, cudaStream_t stream) {
  if (!w || !dst || !air_conv3x3_patch_supported(C, N, 1, 1) || (mode != 0 && mode != 1)) return AIR_ERR_ARG;
  const long long total = 9LL * C * N;
  pack3x3_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 2048)), 256, 0, stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(dst), C, N, patch_cb(C), mode);
  return air_launch_status();
}

extern "C" int air_conv3x3_patch_bf16(const void* a, long long a_ld, int B, int H, int W, int C,
                                      const void* wpk, int N, void* out, long long out_ld,
                                      const void* res, long long res_ld, int relu, int num_sms, cudaStream_t stream) {
  if (!a || !wpk || !out || B <= 0) return AIR_ERR_ARG;
  if (!air_conv3x3_patch_supported(C, N, H, W)) return AIR_ERR_UNSUPPORTED;
  if (a_ld % 8 != 0 || out_ld % 8 != 0 || (res && res_ld % 8 != 0)) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(wpk) |
       reinterpret_cast<uintptr_t>(res)) & 15) return AIR_ERR_UNSUPPORTED;
  PatchParams p;
  p.B = B; p.H = H; p.W = W; p.C = C;
  p.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk); p.N = N;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_ld = out_ld;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.res_ld = res_ld; p.relu = relu;
  p.CB = patch_cb(C); p.NCB = C / p.CB; p.WT = (W + TW - 1) / TW; p.HP = (H + R - 1) / R;
  p.row_bytes = p.CB * 2; p.layout = p.CB == 64 ? 2 : (p.CB == 32 ? 4 : 6);
  p.items = static_cast<long long>(B) * p.HP * p.WT;
  p.acc_stages = (2 * R * N <= 512) ? 2 : 1;
  { const char* e = getenv("AIR_PATCH_DBG"); p.dbg = e ? atoi(e) : 0; }
  p.pstage_bytes = static_cast<uint32_t>((PPIX * p.row_bytes + 1023) / 1024 * 1024);
  p.bslot_bytes = static_cast<uint32_t>(N * p.row_bytes);
  p.bslot_stride = (p.bslot_bytes + 1023u) / 1024u * 1024u;
  const int budget = 225 * 1024 - PSTAGES * static_cast<int>(p.pstage_bytes) - 2048;
  int slots = budget / static_cast<int>(p.bslot_stride);
  const int nslices = p.NCB * 9;
  if (slots >= nslices) { slots = nslices; p.resident = 1; } else { p.resident = 0; if (slots > 12) slots = 12; }
  if (slots < 2) return AIR_ERR_UNSUPPORTED;
  p.nb_slots = slots;
  const size_t smem = 1024 + static_cast<size_t>(PSTAGES) * p.pstage_bytes + static_cast<size_t>(slots) * p.bslot_stride +
                      (2 * PSTAGES + 2 * slots + 4) * 8 + 16;
  CUtensorMap tm;
  const int tr = air_tmap::make_act_tmap(&tm, a, a_ld, B, H, W, C, p.CB, PW, PR, p.row_bytes);
  if (tr != 0) return tr < 0 ? AIR_ERR_UNSUPPORTED : 10000 + tr;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int grid = static_cast<int>(std::min<long long>(p.items, num_sms));
  conv3x3_patch_kernel<<<grid, THREADS, smem, stream>>>(tm, p);
  return air_launch_status();
}
