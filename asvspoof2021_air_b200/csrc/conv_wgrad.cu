// Weight-gradient of the implicit-GEMM convolution on tcgen05 tensor cores.
//
//   dW[n, kx] += sum_{pixels m} dy[m, n] * im2col(x)[m, kx]        n = co, kx = (tap, ci)
//
// The reduction runs over pixels, so both operands are "MN-major" for the tensor core: a shared-memory row is
// one pixel (the GEMM K index) holding 64 consecutive kx (or output channels) = 128 B, SWIZZLE_128B; a 128-wide
// M (or N) extent is two such 8 KB images, chained through the descriptor's leading byte offset.  Eight consecutive
// lanes copy the eight 16-byte chunks of one pixel row, so a cp.async instruction reads four contiguous 128-byte
// segments of global memory and writes four consecutive shared-memory rows.  Tile: 256 (kx) x block_n (co) accumulated in TMEM as two M = 128
// halves; the pixel range is split across CTAs (split-K) and reduced with fp32 red.global.add into
// the caller-zeroed gradient buffer (GEMM layout [Cout][taps*Cin], fp32).
#include <algorithm>
#include "common.cuh"
#include "tc05.cuh"
#include "tmap.cuh"

namespace air_wgrad {
using namespace tc05;

constexpr int TILE_M = 256;              // kx per work item (two UMMA M=128 halves)
constexpr int PIX = 64;                  // pixels per pipeline stage (GEMM K block)
constexpr int GROUP_BYTES = PIX * 128;           // one [64 pixels][64 elements] SWIZZLE_128B image = 8 KB
constexpr int A_HALF_BYTES = 2 * GROUP_BYTES;    // 128 kx = two images = 16 KB
constexpr int THREADS = 256;             // 4 gather warps, 1 MMA warp, (1 idle), ... see roles below

struct WgradParams {
  const __nv_bfloat16* x; long long x_ld; int H, W, C;       // forward input (gather source)
  const __nv_bfloat16* dy; long long dy_ld; int Ho, Wo, N;   // output gradient, N = Cout
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int Kx;                                 // taps * C
  float* dw_out; long long dw_ld;         // [N][dw_ld >= Kx] fp32, accumulated atomically
  long long M;                            // pixels = B*Ho*Wo
  int m_tiles, n_tiles, splits, kb_total, kb_per_split, block_n;
  int stages, flags;
  int use_tma;                            // 1x1 / stride 1: x and dy tiles are plain 2-D boxes of [pixels][channels] matrices
};

struct ChunkInfo { int hoff, woff, ci, ok; };

__global__ void __launch_bounds__(THREADS, 1) conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                const __grid_constant__ CUtensorMap tmdy, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzle patterns are anchored at 1024 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int nbg = (p.block_n + 63) / 64;                                        // 64-channel images of dy per stage
  const uint32_t b_bytes = static_cast<uint32_t>(nbg) * GROUP_BYTES;
  const uint32_t stage_bytes = 2 * A_HALF_BYTES + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* full = bars;            // [S] 128 gather arrivals
  uint64_t* empty = bars + S;       // [S] tcgen05.commit
  uint64_t* tfull = bars + 2 * S;   // [1] accumulators ready
  uint64_t* tempty = bars + 2 * S + 1;   // [1] accumulators drained (128 epilogue threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2);
  ChunkInfo* cinfo = reinterpret_cast<ChunkInfo*>(bars + 2 * S + 4);   // [32] per work item

  uint32_t ncols = 32;
  while (ncols < 2u * p.block_n) ncols <<= 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 128); mbar_init(&empty[s], 1); }
    mbar_init(tfull, 1); mbar_init(tempty, 128);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int total_items = p.m_tiles * p.n_tiles * p.splits;

  // Roles: warps 0-3 gather (and, after the main loop of an item, act as the epilogue: they own
  // TMEM lane quadrants 0-3), warp 4 issues the MMAs.  Warps 5-7 idle (kept so that blockDim is
  // a multiple of 128 for the quadrant mapping).
  uint32_t stage = 0, phase = 0;
  int it = 0;
  for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
    const int split = item % p.splits;
    const int nt = (item / p.splits) % p.n_tiles;
    const int mt = item / (p.splits * p.n_tiles);
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int m0 = mt * TILE_M, n0 = nt * p.block_n;
    const int halves = (p.Kx - m0 > 128) ? 2 : 1;

    if (threadIdx.x < 32) {             // per-item chunk table: kx = m0 + 8c -> (tap offsets, ci)
      const int kx = m0 + 8 * threadIdx.x;
      ChunkInfo ci; ci.ok = kx < p.Kx;
      const int tp = ci.ok ? kx / p.C : 0;
      ci.ci = ci.ok ? kx - tp * p.C : 0;
      const int khi = tp / p.kw, kwi = tp - khi * p.kw;
      ci.hoff = khi * p.dh - p.ph; ci.woff = kwi * p.dw - p.pw;
      cinfo[threadIdx.x] = ci;
    }
    __syncthreads();

    if (warp < 4) {
      const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;              // chunk of a 64-element group, first pixel row
      const uint32_t rowoff = r0 * 128 + ((c ^ (r0 & 7)) << 4);           // (r0 + 16 i) & 7 == r0 & 7
      const int nag = 2 * halves;                                         // 64-kx images of x to fill
      if (p.use_tma) {
        // 1x1 / stride-1 layer: [64 pixels][64 channels] TMA boxes, no gather.  Thread 0 issues them (its arrival carries
        // the transaction bytes); the other gather threads only keep the barrier's arrival count.
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (threadIdx.x == 0) {
            const uint32_t s0 = smem_u32(smem) + stage * stage_bytes;
            mbar_arrive_expect_tx(&full[stage], static_cast<uint32_t>(nag + nbg) * GROUP_BYTES);
            for (int g = 0; g < nag; ++g) tma_load_2d(s0 + g * GROUP_BYTES, &tmx, m0 + g * 64, kb * PIX, &full[stage]);
            for (int g = 0; g < nbg; ++g) tma_load_2d(s0 + 2 * A_HALF_BYTES + g * GROUP_BYTES, &tmdy, n0 + g * 64, kb * PIX, &full[stage]);
          } else {
            mbar_arrive(&full[stage]);
          }
          if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
        }
      } else {
      // pixel coordinates of this thread's four rows: decoded once per item, then advanced by PIX per stage
      int rw[PIX / 16], rh[PIX / 16], rb[PIX / 16];
#pragma unroll
      for (int i = 0; i < PIX / 16; ++i) {
        const long long m = static_cast<long long>(kb0) * PIX + r0 + 16 * i;
        rw[i] = static_cast<int>(m % p.Wo); const long long t = m / p.Wo;
        rh[i] = static_cast<int>(t % p.Ho); rb[i] = static_cast<int>(t / p.Ho);
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t sbase = smem_u32(smem) + stage * stage_bytes + rowoff;
#pragma unroll
        for (int i = 0; i < PIX / 16; ++i) {
          const long long m = static_cast<long long>(kb) * PIX + r0 + 16 * i;
          const bool row_ok = m < p.M;
          const int wo = rw[i], ho = rh[i], bb = rb[i];
          rw[i] += PIX;                                                   // next stage: the same row is PIX pixels further
          while (rw[i] >= p.Wo) { rw[i] -= p.Wo; if (++rh[i] == p.Ho) { rh[i] = 0; ++rb[i]; } }
          const int hs = ho * p.sh, ws = wo * p.sw;
          const __nv_bfloat16* img = p.x + static_cast<long long>(bb) * p.H * p.W * p.x_ld;
          const uint32_t drow = sbase + i * (16 * 128);
          for (int g = 0; g < nag; ++g) {                                 // x: chunk c of each 64-kx group
            const ChunkInfo ci = cinfo[g * 8 + c];
            const int hi = hs + ci.hoff, wi = ws + ci.woff;
            const bool ok = row_ok && ci.ok && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            const __nv_bfloat16* src = ok ? img + (static_cast<long long>(hi) * p.W + wi) * p.x_ld + ci.ci : p.x;
            cp_async16(drow + g * GROUP_BYTES, src, ok ? 16u : 0u);
          }
          const __nv_bfloat16* dsrc = p.dy + m * p.dy_ld + n0 + c * 8;
          for (int g = 0; g < nbg; ++g) {                                 // dy: chunk c of each 64-channel group
            const bool ok = row_ok && (g * 64 + c * 8) < p.block_n;
            cp_async16(drow + 2 * A_HALF_BYTES + g * GROUP_BYTES, ok ? dsrc + g * 64 : p.dy, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(&full[stage]);
        if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
      }
      }
      // ---------------- epilogue (same warps): TMEM -> red.global.add ----------------
      mbar_wait(tfull, it & 1);
      fence_after_sync();
      const int q = warp & 3;
      for (int h = 0; h < halves; ++h) {
        const int kx = m0 + h * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * p.block_n;
        for (int c0 = 0; c0 < p.block_n; c0 += 16) {
          float v[16];
          tmem_ld16(taddr + c0, v);
          if (kx < p.Kx) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              atomicAdd(p.dw_out + static_cast<long long>(n0 + c0 + i) * p.dw_ld + kx, v[i]);
          }
        }
      }
      fence_before_sync();
      mbar_arrive(tempty);
    } else if (warp == 4) {
      // warp-uniform issue loop: descriptors in uniform registers, one elected lane issues
      const bool leader = elect_one();
      const uint32_t idesc = instr_desc_bf16(128, p.block_n, 1, 1);
      // MN-major SWIZZLE_128B: SBO = next 8 pixel rows (1024 B), LBO = next 64-element image (8 KB), version 1
      const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29), lbo16 = (static_cast<uint32_t>(GROUP_BYTES) >> 4) << 16;
      mbar_wait(tempty, (it & 1) ^ 1);
      fence_after_sync();
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full[stage], phase);
        fence_after_sync();
        const uint32_t s0 = (((smem_u32(smem) + stage * stage_bytes) >> 4) & 0x3FFF) | lbo16;
#pragma unroll
        for (int kk = 0; kk < PIX / 16; ++kk) {                            // 16 pixel rows = 2048 B per K step
          const uint64_t bd = (static_cast<uint64_t>(d_hi) << 32) | (s0 + ((2 * A_HALF_BYTES + kk * 2048) >> 4));
          for (int h = 0; h < halves; ++h) {
            const uint64_t ad = (static_cast<uint64_t>(d_hi) << 32) | (s0 + ((h * A_HALF_BYTES + kk * 2048) >> 4));
            if (leader) mma_bf16(tmem_base + h * p.block_n, ad, bd, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
        }
        if (leader) mma_commit(&empty[stage]);
        if (kb == kb1 - 1 && leader) mma_commit(tfull);
        if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
    }
    __syncthreads();      // cinfo reuse + keeps the item sequence aligned across roles
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 4) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
}

}  // namespace air_wgrad

using namespace air_wgrad;

extern "C" int air_conv_wgrad_bf16_ld(const void* x, long long x_ld, int B, int H, int W, int C,
                                      const void* dy, long long dy_ld, int Ho, int Wo, int N,
                                      int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                                      float* dw_out, long long dw_ld, int num_sms, int flags, cudaStream_t stream);

extern "C" int air_conv_wgrad_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                   const void* dy, long long dy_ld, int Ho, int Wo, int N,
                                   int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                                   float* dw_out, int num_sms, int flags, cudaStream_t stream) {
  return air_conv_wgrad_bf16_ld(x, x_ld, B, H, W, C, dy, dy_ld, Ho, Wo, N, kh, kw, sh, sw, ph, pw, dh, dw, dw_out,
                                static_cast<long long>(kh) * kw * C, num_sms, flags, stream);
}

// dw_ld: elements between consecutive output-channel rows of dw_out (a column slice of a wider gradient)
extern "C" int air_conv_wgrad_bf16_ld(const void* x, long long x_ld, int B, int H, int W, int C,
                                      const void* dy, long long dy_ld, int Ho, int Wo, int N,
                                      int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                                      float* dw_out, long long dw_ld, int num_sms, int flags, cudaStream_t stream) {
  if (!x || !dy || !dw_out || B <= 0) return AIR_ERR_ARG;
  if (C % 8 != 0 || x_ld % 8 != 0 || dy_ld % 8 != 0 || N % 16 != 0) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) return AIR_ERR_UNSUPPORTED;
  int bn = N <= 256 ? N : (N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 0)));
  if (bn == 0) return AIR_ERR_UNSUPPORTED;
  WgradParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(x); p.x_ld = x_ld; p.H = H; p.W = W; p.C = C;
  p.dy = reinterpret_cast<const __nv_bfloat16*>(dy); p.dy_ld = dy_ld; p.Ho = Ho; p.Wo = Wo; p.N = N;
  p.kh = kh; p.kw = kw; p.sh = sh; p.sw = sw; p.ph = ph; p.pw = pw; p.dh = dh; p.dw = dw;
  p.Kx = kh * kw * C; p.dw_out = dw_out; p.dw_ld = dw_ld;
  if (dw_ld < p.Kx) return AIR_ERR_ARG; p.M = static_cast<long long>(B) * Ho * Wo;
  p.block_n = bn; p.n_tiles = N / bn; p.m_tiles = (p.Kx + TILE_M - 1) / TILE_M;
  p.kb_total = static_cast<int>((p.M + PIX - 1) / PIX);
  if (num_sms <= 0) num_sms = 148;
  const int base_items = p.m_tiles * p.n_tiles;
  int splits = (2 * num_sms + base_items - 1) / base_items;          // ~2 work items per SM
  splits = std::max(1, std::min(splits, (p.kb_total + 7) / 8));      // at least 8 K blocks per split
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.flags = flags;
  const int stage_bytes = 2 * A_HALF_BYTES + ((bn + 63) / 64) * GROUP_BYTES;
  int stages = (196 * 1024) / stage_bytes;
  if (stages > 6) stages = 6;
  if (stages < 2) return AIR_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smem = 1024 + static_cast<size_t>(stages) * stage_bytes + (2 * stages + 4) * 8 + 32 * sizeof(ChunkInfo) + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  const int total_items = base_items * p.splits;
  const int grid = std::min(total_items, num_sms);
  CUtensorMap tmx{}, tmdy{};
  p.use_tma = 0;
  if (kh == 1 && kw == 1 && sh == 1 && sw == 1 && ph == 0 && pw == 0 && Ho == H && Wo == W && p.M < 0x7fffffffLL) {
    if (air_tmap::make_mat_tmap(&tmx, x, x_ld, p.M, C, PIX) == 0 && air_tmap::make_mat_tmap(&tmdy, dy, dy_ld, p.M, N, PIX) == 0)
      p.use_tma = 1;                                         // otherwise: the gather path
  }
  conv_wgrad_kernel<<<grid, THREADS, smem, stream>>>(tmx, tmdy, p);
  return air_launch_status();
}
