// 3x3 / stride-1 / pad-1 convolution with a shared-memory resident input patch (tcgen05, TMEM).
//
// The generic implicit-GEMM kernel (conv_gemm.cu) gathers every activation once per tap (9x) from
// L2, which makes the N = 64 / 128 layers of the ResNet (resnet.py:56-60, layer1 / layer2) L2-bandwidth
// bound.  Here a work item is R = 2 output rows x 128 pixels of one image; its (R+2) x 130 pixel input
// patch (64 channels at a time) is brought into shared memory ONCE and the nine taps are nine
// shifted windows of it: in the "column of rows" operand image (tc05.cuh) moving the matrix by one pixel
// is a +16 B change of the UMMA descriptor start address, by one image row +130*16 B.
//
//   out[b, h, w, n] = sum_{i,j,c} a[b, h+i-1, w+j-1, c] * Wp[n][i][j][c]   (+ res) (ReLU)
//
// Forward uses Wp = W; the data gradient of the same layer is the same kernel on dy with the taps
// flipped and (ci, co) swapped, which is done once in the weight packing (mode 1).
// Roles (320 threads): warps 0-3 gather the patch (zero-filling 16-byte cp.async, image borders are
// the zero fill), warp 4 streams pre-packed weight slices (one bulk copy per (channel block, tap); all
// slices stay resident when they fit), warp 5 issues tcgen05.mma (M = 128 pixels, N = Cout, K = 16 per
// instruction), warps 6-9 drain the R double-buffered TMEM accumulators.  Persistent grid.
#include <algorithm>
#include "common.cuh"
#include "tc05.cuh"

namespace air_patch {
using namespace tc05;

constexpr int TW = 128;                 // output pixels per row segment (UMMA M)
constexpr int PW = TW + 2;              // patch width
constexpr int R = 2;                    // output rows per work item
constexpr int PR = R + 2;               // patch rows
constexpr int PPIX = PR * PW;           // 520 patch pixels
constexpr int CH = (PPIX + 1) * 16;     // bytes between 8-channel chunks (+1 row: conflict-free cp.async stores)
constexpr int THREADS = 320;
constexpr int NGATHER = 128;
constexpr int PSTAGES = 2;

struct PatchParams {
  const __nv_bfloat16* a; long long a_ld; int B, H, W, C;
  const __nv_bfloat16* wpk; int N;
  __nv_bfloat16* out; long long out_ld;
  const __nv_bfloat16* res; long long res_ld; int relu;
  int CB, NCB, WT, HP;                  // channels per block (<= 64), #blocks, W tiles, row pairs
  int nb_slots, resident;               // weight ring
  int acc_stages;                       // 1 or 2
  long long items;
};

__global__ void __launch_bounds__(THREADS, 1) conv3x3_patch_kernel(const PatchParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = p.CB >> 3;
  const uint32_t pstage_bytes = static_cast<uint32_t>(nchunk) * CH;
  const uint32_t bslot_bytes = static_cast<uint32_t>(p.N) * p.CB * 2;
  uint8_t* sP = smem;
  uint8_t* sB = smem + PSTAGES * pstage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + static_cast<size_t>(p.nb_slots) * bslot_bytes);
  uint64_t* full_p = bars;                          // [PSTAGES] 128 gather arrivals
  uint64_t* empty_p = bars + PSTAGES;               // [PSTAGES] tcgen05.commit
  uint64_t* full_b = bars + 2 * PSTAGES;            // [nb_slots] expect_tx
  uint64_t* empty_b = full_b + p.nb_slots;          // [nb_slots] tcgen05.commit
  uint64_t* tfull = empty_b + p.nb_slots;           // [2]
  uint64_t* tempty = tfull + 2;                     // [2] 128 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(p.acc_stages * R * p.N)) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full_p[s], NGATHER); mbar_init(&empty_p[s], 1); }
    for (int s = 0; s < p.nb_slots; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nslices = p.NCB * 9;

  if (warp < 4) {
    // ===================== patch gather =====================
    const int t = threadIdx.x;
    uint32_t stage = 0, phase = 0;
    for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int wt = static_cast<int>(item % p.WT);
      const long long r1 = item / p.WT;
      const int hp = static_cast<int>(r1 % p.HP), b = static_cast<int>(r1 / p.HP);
      const int h0 = hp * R - 1, w0 = wt * TW - 1;
      const __nv_bfloat16* img = p.a + static_cast<long long>(b) * p.H * p.W * p.a_ld;
      for (int cb = 0; cb < p.NCB; ++cb) {
        mbar_wait(&empty_p[stage], phase ^ 1);
        const uint32_t dst0 = smem_u32(sP) + stage * pstage_bytes;
        const int total = PPIX * nchunk;
        // element e -> chunk = e % nchunk, pixel = e / nchunk: consecutive lanes read consecutive 16-byte
        // chunks of one pixel (coalesced), and write rows of different chunk columns (conflict-free, CH odd*16)
        int e = t;
        int c = e % nchunk, pix = e / nchunk;
        int pr = pix / PW, pp = pix - pr * PW;
        const int dpix = NGATHER / nchunk, dc = NGATHER % nchunk;      // per-iteration increments
        for (; e < total; e += NGATHER) {
          const int h = h0 + pr, w = w0 + pp;
          const bool ok = h >= 0 && h < p.H && w >= 0 && w < p.W;
          const __nv_bfloat16* src = ok ? img + (static_cast<long long>(h) * p.W + w) * p.a_ld + cb * p.CB + c * 8 : p.a;
          cp_async16(dst0 + c * CH + pix * 16, src, ok ? 16u : 0u);
          c += dc; pix += dpix; pp += dpix;
          if (c >= nchunk) { c -= nchunk; ++pix; ++pp; }
          while (pp >= PW) { pp -= PW; ++pr; }
        }
        cp_async_arrive_noinc(&full_p[stage]);
        if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 4) {
    // ===================== weight slices =====================
    if (lane == 0) {
      if (p.resident) {
        for (int s = 0; s < nslices; ++s) {
          mbar_arrive_expect_tx(&full_b[s], bslot_bytes);
          bulk_g2s(smem_u32(sB) + s * bslot_bytes, p.wpk + static_cast<long long>(s) * p.N * p.CB, bslot_bytes, &full_b[s]);
        }
      } else {
        uint32_t slot = 0, phase = 0;
        for (long long item = blockIdx.x; item < p.items; item += gridDim.x) {
          for (int s = 0; s < nslices; ++s) {
            mbar_wait(&empty_b[slot], phase ^ 1);
            mbar_arrive_expect_tx(&full_b[slot], bslot_bytes);
            bulk_g2s(smem_u32(sB) + slot * bslot_bytes, p.wpk + static_cast<long long>(s) * p.N * p.CB, bslot_bytes, &full_b[slot]);
            if (++slot == static_cast<uint32_t>(p.nb_slots)) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = instr_desc_bf16(TW, p.N, 0, 0);
      const uint32_t b_chunk = static_cast<uint32_t>(p.N) * 16;
      const int KK = p.CB >> 4;
      uint32_t pstage = 0, pphase = 0, slot = 0, bphase = 0;
      int it = 0;
      for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const long long r1 = item / p.WT;
        const int hp = static_cast<int>(r1 % p.HP);
        const int rows = min(R, p.H - hp * R);
        const int acc = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t acc_phase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        fence_after_sync();
        for (int cb = 0; cb < p.NCB; ++cb) {
          mbar_wait(&full_p[pstage], pphase);
          fence_after_sync();
          const uint32_t a0 = smem_u32(sP) + pstage * pstage_bytes;
          for (int tap = 0; tap < 9; ++tap) {
            const int ti = tap / 3, tj = tap - ti * 3;
            uint32_t b0;
            if (p.resident) {
              const int s = cb * 9 + tap;
              if (it == 0) { mbar_wait(&full_b[s], 0); fence_after_sync(); }
              b0 = smem_u32(sB) + s * bslot_bytes;
            } else {
              mbar_wait(&full_b[slot], bphase);
              fence_after_sync();
              b0 = smem_u32(sB) + slot * bslot_bytes;
            }
            for (int j = 0; j < rows; ++j) {
              const uint32_t d_tmem = tmem_base + (acc * R + j) * p.N;
              const uint32_t arow = a0 + ((j + ti) * PW + tj) * 16;
              for (int kk = 0; kk < KK; ++kk) {
                const uint64_t ad = smem_desc(arow + kk * 2 * CH, CH, 128);
                const uint64_t bd = smem_desc(b0 + kk * 2 * b_chunk, b_chunk, 128);
                mma_bf16(d_tmem, ad, bd, idesc, (cb | tap | kk) != 0);
              }
            }
            if (!p.resident) {
              mma_commit(&empty_b[slot]);
              if (++slot == static_cast<uint32_t>(p.nb_slots)) { slot = 0; bphase ^= 1; }
            }
          }
          mma_commit(&empty_p[pstage]);
          if (++pstage == PSTAGES) { pstage = 0; pphase ^= 1; }
        }
        mma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    int it = 0;
    for (long long item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int wt = static_cast<int>(item % p.WT);
      const long long r1 = item / p.WT;
      const int hp = static_cast<int>(r1 % p.HP), b = static_cast<int>(r1 / p.HP);
      const int rows = min(R, p.H - hp * R);
      const int acc = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(&tfull[acc], acc_phase);
      fence_after_sync();
      const int w = wt * TW + m;
      for (int j = 0; j < rows; ++j) {
        const long long pixel = (static_cast<long long>(b) * p.H + hp * R + j) * p.W + w;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (acc * R + j) * p.N;
        for (int c0 = 0; c0 < p.N; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (w < p.W) {
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
  if (warp == 5) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
}

// Weight packing for the patch kernel: dst[((cb*9 + tap)*nchunk + c) * N*8 + n*8 + e] = value(n, tap, cb*CB + c*8 + e)
//   mode 0 (fprop): w is [N = Cout][9][C = Cin];   value = w[n][tap][ch]
//   mode 1 (dgrad): w is [C = Cout][9][N = Cin];   value = w[ch][8 - tap][n]   (flipped taps, swapped channels)
__global__ void pack3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int C, int N, int CB, int mode) {
  const long long total = 9LL * C * N;
  const int nchunk = CB >> 3;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i & 7);
    long long t = i >> 3;
    const int n = static_cast<int>(t % N); t /= N;
    const int c = static_cast<int>(t % nchunk); t /= nchunk;
    const int tap = static_cast<int>(t % 9);
    const int cb = static_cast<int>(t / 9);
    const int ch = cb * CB + c * 8 + e;
    const float v = mode == 0 ? w[(static_cast<long long>(n) * 9 + tap) * C + ch]
                              : w[(static_cast<long long>(ch) * 9 + (8 - tap)) * N + n];
    dst[i] = f2bf(v);
  }
}

}  // namespace air_patch

using namespace air_patch;

static int patch_cb(int C) { return C <= 64 ? C : 64; }

// 1 when (C, N, W) can run on the patch kernel
extern "C" int air_conv3x3_patch_supported(int C, int N, int H, int W) {
  if (C % 16 != 0 || (C > 64 && C % 64 != 0)) return 0;
  if (N % 16 != 0 || N > 256) return 0;
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
  p.a = reinterpret_cast<const __nv_bfloat16*>(a); p.a_ld = a_ld; p.B = B; p.H = H; p.W = W; p.C = C;
  p.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk); p.N = N;
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_ld = out_ld;
  p.res = reinterpret_cast<const __nv_bfloat16*>(res); p.res_ld = res_ld; p.relu = relu;
  p.CB = patch_cb(C); p.NCB = C / p.CB; p.WT = (W + TW - 1) / TW; p.HP = (H + R - 1) / R;
  p.items = static_cast<long long>(B) * p.HP * p.WT;
  p.acc_stages = (2 * R * N <= 512) ? 2 : 1;
  const int pstage_bytes = (p.CB / 8) * CH;
  const int bslot_bytes = N * p.CB * 2;
  const int budget = 222 * 1024 - PSTAGES * pstage_bytes - 1024;
  int slots = budget / bslot_bytes;
  const int nslices = p.NCB * 9;
  if (slots >= nslices) { slots = nslices; p.resident = 1; } else { p.resident = 0; if (slots > 12) slots = 12; }
  if (slots < 2) return AIR_ERR_UNSUPPORTED;
  p.nb_slots = slots;
  const size_t smem = static_cast<size_t>(PSTAGES) * pstage_bytes + static_cast<size_t>(slots) * bslot_bytes +
                      (2 * PSTAGES + 2 * slots + 4) * 8 + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int grid = static_cast<int>(std::min<long long>(p.items, num_sms));
  conv3x3_patch_kernel<<<grid, THREADS, smem, stream>>>(p);
  return air_launch_status();
}
