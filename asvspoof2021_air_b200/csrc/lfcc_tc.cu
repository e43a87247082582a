// Fused LFCC front-end on the tensor cores (tcgen05 + TMEM), the fast path of air_lfcc_fwd.
//
// Same contract as csrc/lfcc.cu (feature_extraction.py:93-138 + dataset.py:66-79,513-528 + main_train.py:338,347-348),
// different arithmetic for the 512-point spectrum: the radix FFT on CUDA cores (FP32-issue bound, 9 % of the HBM
// roofline) is replaced by a real DFT as a GEMM.  With the window centre n = 160 as origin the 320 windowed samples
// a[n] of a frame fold into an even and an odd part,
//     e[m] = a[160+m] + a[160-m],  o[m] = a[160+m] - a[160-m]   (m = 1..159),   e[0] = a[160],  o[0] = 0,
//     Re X_k = sum_m e[m] cos(2 pi k m / 512) + a[0] cos(5 pi k / 8)
//     Im X_k = sum_m o[m] sin(2 pi k m / 512) - a[0] sin(5 pi k / 8)          (sign irrelevant for |X_k|^2)
// i.e. two [frames x 160] x [160 x 255] products (bins 0 and 256 carry no filterbank weight).  16-bit tensor cores
// with fp32 accumulation reach the fp32 parity bar through a 3-term split of x = x_hi + x_lo and w = w_hi + w_lo,
//     x*w ~ x_hi*w_hi  (fp16 x fp16: 11 + 11 significant bits)
//         + x_lo*w_b   (bf16 x bf16: the residual of x keeps fp32's exponent range, so quiet recordings and tiny residuals
//                       do not underflow; 8 bits of w = bf16(w) are plenty against a term that is 2^-11 of the product)
//         + x_hi*w_lo  (fp16 x fp16: w_lo = w - w_hi sits in fp16's subnormal range, absolute precision 2^-24)
// (kind::f16 takes fp16 or bf16, but both operands of one instruction must have the same format: mixing them is an
// illegal instruction on sm_100a).  Worst deviation 1.4e-6 * (|ref| + 1) on the golden waves, 1.7e-5 on a speech-like
// frame with bands 60 dB under its peak, 1.5e-6 on the same frame 100 dB quieter (a bf16 / bf16 split: 7e-6 and 1.0e-4; an
// fp16 / fp16 split: 3e-4 on quiet input) -- tolerance 1e-4, the reference's own fp32 arithmetic 9e-6.
//
// A tile is 128 consecutive frames (UMMA M) of the flattened (utterance, frame) sequence, 124 of which produce
// output (2 halo frames each side for the delta-deltas).  Per tile, per CTA (persistent, grid = #SMs):
//   prep warps (8)   : wave -> pre-emphasis -> window -> fold -> bf16 hi/lo -> SWIZZLE_64B K-major A images in smem
//   producer warp    : streams the pre-swizzled DFT matrices (w_hi fp16 / w bf16 / w_lo fp16, 8 KB chunks) through a shared-memory ring
//   MMA warp         : 120 x tcgen05.mma 128x128x16 per tile; Re/Im x two 128-bin halves = 4 accumulators (512 TMEM columns)
//   epilogue warps(4): TMEM -> |X|^2 -> sparse triangular filterbank (2 filters per bin, compile-time structure,
//                      run-time weights) -> log10 -> 20x20 DCT -> cepstra in smem -> deltas, pad/crop map, layout, dtype
#include <algorithm>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc05.cuh"

namespace air_lfcc_tc {
using namespace tc05;

constexpr int NF = 20, FL = 320, FS = 160;
constexpr int TM = 128, TOUT = 124, HALO = 2;
constexpr int KBLK = 5;                         // 160 = 5 x 32 folded samples
constexpr uint32_t CHUNK = 128 * 64;            // one [128 rows][32 bf16] SWIZZLE_64B image
constexpr int NSLOT = 5;
constexpr int PREP_WARPS = 16;
constexpr int THREADS = 32 * (6 + PREP_WARPS);  // warps 0-3 epilogue, 4 MMA, 5 producer, 6.. prep
constexpr int CEP_LD = 21;

// fp32 table: window[320], filterbank weights [256][2] (bin k -> weight of filter floor(21k/256)-1 and floor(21k/256)), DCT [20][20]
constexpr int OFF_WIN = 0, OFF_FBW = 320, OFF_DCT = OFF_FBW + 512, TBL_FLOATS = OFF_DCT + 400;

constexpr uint32_t SM_A = 0;                                   // [4 matrices][5 kb] chunks
constexpr uint32_t SM_B = SM_A + 4 * KBLK * CHUNK;             // ring
constexpr uint32_t SM_CEP = SM_B + NSLOT * CHUNK;
constexpr uint32_t SM_A0 = SM_CEP + TM * CEP_LD * 4;
constexpr uint32_t SM_TBL = SM_A0 + 2 * TM * 4;      // a0 is double-buffered by tile parity
constexpr uint32_t SM_ROW = SM_TBL + TBL_FLOATS * 4;     // per-row (utterance, frame) of the current tile
constexpr uint32_t SM_BAR = SM_ROW + 2 * TM * 4;
constexpr uint32_t SM_TOTAL = SM_BAR + 32 * 8;

struct Params {
  const float* wave; long long ldw; const int* lengths; int L; int B;
  const float* tbl; const __nv_bfloat16* wmat;
  void* out; long long sb, sj, sd; int out_bf16; int time_minor;
  int Tout; int feat_len; int pad_mode; const int* start;
  float preemph;
  int flat;                 // 1: all utterances have T frames, tiles run over the flattened (b, t) sequence
  int T;                    // frames per utterance (flat mode)
  int segs;                 // tiles per utterance (per-utterance mode)
  int tiles;
};

// cos / sin (5 pi j / 8), j = k mod 16: the a[0] sample (window position 0, distance 160 from the fold centre)
__device__ constexpr float C160[16] = {1.f, -0.38268343236508977f, -0.70710678118654752f, 0.92387953251128674f,
                                       0.f, -0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                       -1.f, 0.38268343236508977f, 0.70710678118654752f, -0.92387953251128674f,
                                       0.f, 0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
__device__ constexpr float S160[16] = {0.f, 0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f,
                                       1.f, -0.38268343236508977f, -0.70710678118654752f, 0.92387953251128674f,
                                       0.f, -0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                       -1.f, 0.38268343236508977f, 0.70710678118654752f, -0.92387953251128674f};

struct UttInfo { int T, first, tend, shift, rep; int len; };

__device__ __forceinline__ UttInfo utt_info(const Params& p, int b) {
  UttInfo u;
  int len = p.lengths ? p.lengths[b] : p.L;
  len = max(0, min(len, p.L));
  u.len = len; u.T = 1 + len / FS; u.first = 0; u.tend = u.T; u.shift = 0; u.rep = 0x40000000;
  if (p.feat_len > 0) {
    if (u.T > p.feat_len) { int f = p.start ? p.start[b] : 0; f = max(0, min(f, u.T - p.feat_len));
                            u.first = f; u.tend = f + p.feat_len; u.shift = -f; }
    else if (p.pad_mode == 2) u.rep = u.T;                  // repeat: j = t + m*T
    else if (p.pad_mode == 3) u.shift = p.feat_len - u.T;   // silence is PREPENDED (dataset.py:528)
  }
  return u;
}

// row r of tile -> (utterance, frame); false when the row is outside every utterance
__device__ __forceinline__ bool row_bt(const Params& p, int tile, int r, int& b, int& t) {
  if (p.flat) {
    const int F = tile * TOUT - HALO + r;                  // B * T < 2^31 is checked by the launcher
    if (F < 0 || F >= p.B * p.T) return false;
    b = static_cast<int>(static_cast<uint32_t>(F) / static_cast<uint32_t>(p.T)); t = F - b * p.T;
    return true;
  }
  b = tile / p.segs;
  const UttInfo u = utt_info(p, b);
  t = u.first + (tile - b * p.segs) * TOUT - HALO + r;
  return t >= 0 && t < u.T;
}

template <int H>
__device__ __forceinline__ void epilogue_half(uint32_t tcol, float a0, const float2* __restrict__ fbw, float (&fbv)[NF + 2]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float re[16], im[16];
    tmem_ld16(tcol + H * 256 + 16 * c, re);
    tmem_ld16(tcol + H * 256 + 128 + 16 * c, im);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const int k = 1 + 128 * H + 16 * c + n;             // compile-time after unrolling
      if (k <= 255) {
        const float r = fmaf(a0, C160[k & 15], re[n]);
        const float q = fmaf(-a0, S160[k & 15], im[n]);
        const float P = fmaf(q, q, r * r);                  // |X_k|^2   (feature_extraction.py:113)
        const float2 w = fbw[k];
        const int fh = (21 * k) >> 8;                       // bin k feeds filters fh-1 and fh only
        fbv[fh] = fmaf(w.x, P, fbv[fh]);
        fbv[fh + 1] = fmaf(w.y, P, fbv[fh + 1]);
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) lfcc_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));
  float* s_cep = reinterpret_cast<float*>(gen + SM_CEP);
  float* s_a0 = reinterpret_cast<float*>(gen + SM_A0);
  float* s_tbl = reinterpret_cast<float*>(gen + SM_TBL);
  int* s_rowb = reinterpret_cast<int*>(gen + SM_ROW);
  int* s_rowt = s_rowb + TM;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + SM_BAR);
  uint64_t* a_full = bars;             // [2] e / o images written (PREP_WARPS*32 arrivals)
  uint64_t* a_empty = bars + 2;        // [2] tcgen05.commit
  uint64_t* b_full = bars + 4;         // [NSLOT]
  uint64_t* b_empty = b_full + NSLOT;  // [NSLOT]
  uint64_t* tfull = b_empty + NSLOT;   // [2] per 128-bin half
  uint64_t* tempty = tfull + 2;        // [2] 128 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  for (int i = threadIdx.x; i < TBL_FLOATS; i += THREADS) s_tbl[i] = __ldg(p.tbl + i);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], PREP_WARPS * 32); mbar_init(&a_empty[s], 1);
                                  mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 128); }
    for (int s = 0; s < NSLOT; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 6) {
    // ===================== prep: wave -> folded bf16 hi/lo operand images =====================
    const int pw = warp - 6;
    const float* s_win = s_tbl + OFF_WIN;
    float wp[KBLK], wm[KBLK];                             // window at 160 + m and 160 - m for this lane's five m
#pragma unroll
    for (int i = 0; i < KBLK; ++i) { wp[i] = s_win[160 + 32 * i + lane]; wm[i] = s_win[160 - 32 * i - lane]; }
    const float w0 = s_win[0];
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        mbar_wait(&a_empty[part], (it & 1) ^ 1);
        uint8_t* img_hi = gen + SM_A + (part * 2 + 0) * KBLK * CHUNK;
        uint8_t* img_lo = gen + SM_A + (part * 2 + 1) * KBLK * CHUNK;
        for (int fr = pw; fr < TM; fr += PREP_WARPS) {
          int b = 0, t = 0;
          const bool valid = row_bt(p, tile, fr, b, t);
          const int len = valid ? (p.lengths ? max(0, min(p.lengths[b], p.L)) : p.L) : 0;
          const float* w = p.wave + static_cast<long long>(valid ? b : 0) * p.ldw;
          const int s0 = FS * (t - 1);
          float xp[KBLK], xq[KBLK], xm[KBLK], xn[KBLK], x0, x00;
          const bool interior = s0 >= 1 && s0 + FL <= len;            // every sample of the frame (and its predecessor) exists
          if (interior) {
            const float* f = w + s0 + 160;
#pragma unroll
            for (int i = 0; i < KBLK; ++i) {
              const int m = 32 * i + lane;
              xp[i] = __ldg(f + m); xq[i] = __ldg(f + m - 1); xm[i] = __ldg(f - m); xn[i] = __ldg(f - m - 1);
            }
            x0 = __ldg(w + s0); x00 = __ldg(w + s0 - 1);
          } else {
            auto ld = [&](int idx) -> float { return (idx >= 0 && idx < len) ? __ldg(w + idx) : 0.f; };
#pragma unroll
            for (int i = 0; i < KBLK; ++i) {
              const int m = 32 * i + lane;
              const int pp = s0 + 160 + m, pm = s0 + 160 - m;
              const float a = ld(pp), a1 = ld(pp - 1), c = ld(pm), c1 = ld(pm - 1);
              // y[n] = x[n] - 0.97 x[n-1] is zero outside [0, len): fold the range test into the operands
              xp[i] = (pp >= 0 && pp < len) ? a : 0.f; xq[i] = (pp >= 0 && pp < len) ? a1 : 0.f;
              xm[i] = (pm >= 0 && pm < len) ? c : 0.f; xn[i] = (pm >= 0 && pm < len) ? c1 : 0.f;
            }
            x0 = (s0 >= 0 && s0 < len) ? ld(s0) : 0.f; x00 = (s0 >= 0 && s0 < len) ? ld(s0 - 1) : 0.f;
          }
          const uint32_t rowoff = static_cast<uint32_t>(fr) * 64 + lane * 2;
          const uint32_t soff = rowoff ^ (((rowoff >> 7) & 3u) << 4);           // SWIZZLE_64B
#pragma unroll
          for (int i = 0; i < KBLK; ++i) {
            const int m = 32 * i + lane;
            // y[n] = x[n] - 0.97 x[n-1]   (feature_extraction.py:105-106; centred zero padding of stft)
            const float yp = fmaf(-p.preemph, xq[i], xp[i]);
            const float ym = fmaf(-p.preemph, xn[i], xm[i]);
            const float ap = yp * wp[i], am = ym * wm[i];
            float v;
            if (part == 0) v = (m == 0) ? ap : ap + am; else v = (m == 0) ? 0.f : ap - am;
            const __half hi = __float2half_rn(v);                       // 11 significant bits
            const __nv_bfloat16 lo = f2bf(v - __half2float(hi));         // the residual, fp32 exponent range
            *reinterpret_cast<__half*>(img_hi + i * CHUNK + soff) = hi;
            *reinterpret_cast<__nv_bfloat16*>(img_lo + i * CHUNK + soff) = lo;
          }
          if (part == 0 && lane == 0) s_a0[(it & 1) * TM + fr] = fmaf(-p.preemph, x00, x0) * w0;
        }
        fence_proxy_async();
        mbar_arrive(&a_full[part]);
      }
    }
  } else if (warp == 5) {
    // ===================== DFT matrix chunks: 40 x 8 KB per tile, in MMA order =====================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int s = 0; s < 2 * 2 * KBLK * 3; ++s) {
          mbar_wait(&b_empty[slot], phase ^ 1);
          mbar_arrive_expect_tx(&b_full[slot], CHUNK);
          bulk_g2s(sbase + SM_B + slot * CHUNK, p.wmat + static_cast<long long>(s) * (CHUNK / 2), CHUNK, &b_full[slot]);
          if (++slot == NSLOT) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
    const bool leader = elect_one();
    const bool committer = leader;
    // operand formats per term (both operands of an instruction share one): fp16 x fp16 for the hi images, bf16 x bf16 for x_lo
    const uint32_t idesc_ff = instr_desc_f16(128, 128, 0, 0, 0, 0), idesc_bb = instr_desc_f16(128, 128, 0, 0, 1, 1);
    const uint32_t desc_hi = (512u >> 4) | (1u << 14) | (4u << 29);      // SBO = 8 rows x 64 B, version 1, SWIZZLE_64B
    const uint32_t a16 = ((sbase + SM_A) >> 4) & 0x3FFF, b16 = ((sbase + SM_B) >> 4) & 0x3FFF;
    uint32_t slot = 0, bphase = 0, it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        mbar_wait(&a_full[part], it & 1);
        fence_after_sync();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (part == 0) { mbar_wait(&tempty[h], (it & 1) ^ 1); fence_after_sync(); }
          const uint32_t d = tmem_base + h * 256 + part * 128;
#pragma unroll
          for (int kb = 0; kb < KBLK; ++kb) {
            const uint32_t ahi = a16 + ((part * 2 + 0) * KBLK + kb) * (CHUNK >> 4);
            const uint32_t alo = a16 + ((part * 2 + 1) * KBLK + kb) * (CHUNK >> 4);
            // three chunks per K block, each used by one term: w_hi (fp16) x x_hi, bf16(w) x x_lo, w_lo (fp16) x x_hi
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              mbar_wait(&b_full[slot], bphase);
              fence_after_sync();
              const uint32_t bch = b16 + slot * (CHUNK >> 4);
              const uint32_t ach = term == 1 ? alo : ahi;
              const uint32_t idesc = term == 1 ? idesc_bb : idesc_ff;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t ad = (static_cast<uint64_t>(desc_hi) << 32) | ((ach + ks * 2) | (1u << 16));
                const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | ((bch + ks * 2) | (1u << 16));
                if (leader) mma_bf16(d, ad, bd, idesc, (term | kb | ks) != 0);
              }
              if (committer) mma_commit(&b_empty[slot]);
              if (++slot == NSLOT) { slot = 0; bphase ^= 1; }
            }
          }
          if (part == 1 && committer) mma_commit(&tfull[h]);
        }
        if (committer) mma_commit(&a_empty[part]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: one thread per frame (TMEM lane) =====================
    const int row = threadIdx.x;                         // warps 0-3 <-> TMEM lane quadrants 0-3
    const float2* s_fbw = reinterpret_cast<const float2*>(s_tbl + OFF_FBW);
    const float* s_dct = s_tbl + OFF_DCT;
    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      float fbv[NF + 2];
#pragma unroll
      for (int f = 0; f < NF + 2; ++f) fbv[f] = 0.f;
      mbar_wait(&tfull[0], it & 1);
      fence_after_sync();
      // a0 was written (double-buffered by tile parity) before the prep warps released a_full[0], which the MMA warp
      // acquired before issuing the MMAs whose completion tfull[0] tracks
      const float a0 = s_a0[(it & 1) * TM + row];
      epilogue_half<0>(tcol, a0, s_fbw, fbv);
      fence_before_sync();
      mbar_arrive(&tempty[0]);
      mbar_wait(&tfull[1], it & 1);
      fence_after_sync();
      epilogue_half<1>(tcol, a0, s_fbw, fbv);
      fence_before_sync();
      mbar_arrive(&tempty[1]);
      // log10 + DCT-II ortho (feature_extraction.py:116-120)
      float fbe[NF];
#pragma unroll
      for (int f = 0; f < NF; ++f) fbe[f] = log10f(fbv[f + 1] + 1.1920928955078125e-07f);
#pragma unroll 4
      for (int k = 0; k < NF; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int f = 0; f < NF; ++f) acc = fmaf(fbe[f], s_dct[k * NF + f], acc);
        s_cep[row * CEP_LD + k] = acc;
      }
      { int b = -1, t = 0; if (!row_bt(p, tile, row, b, t)) b = -1; s_rowb[row] = b; s_rowt[row] = t; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- deltas (replicate edges, feature_extraction.py:41-58), pad/crop scatter, layout, dtype ----
      // Output frames are rows HALO .. HALO+TOUT-1.  In per-utterance mode all rows share one utterance; in flat mode
      // the utterance of each output row is recomputed (tiles may straddle two utterances).
      int ub = -1; UttInfo u{};
      const int total = TOUT * NF;
      for (int i = row; i < total; i += 128) {
        int fo, k;
        if (p.time_minor) { k = i / TOUT; fo = i - k * TOUT; } else { fo = i / NF; k = i - fo * NF; }
        const int r = HALO + fo;
        const int b = s_rowb[r], t = s_rowt[r];
        if (b < 0) continue;
        if (b != ub) { u = utt_info(p, b); ub = b; }
        if (t < u.first || t >= u.tend) continue;
        const int T = u.T;
        const int tp = min(t + 1, T - 1), tm = max(t - 1, 0);
        const float* c = s_cep + k;
        const float cpp = c[(r + min(tp + 1, T - 1) - t) * CEP_LD], cp = c[(r + tp - t) * CEP_LD], c0 = c[r * CEP_LD];
        const float cm = c[(r + tm - t) * CEP_LD], cmm = c[(r + max(tm - 1, 0) - t) * CEP_LD];
        const float cpm = c[(r + max(tp - 1, 0) - t) * CEP_LD], cmp = c[(r + min(tm + 1, T - 1) - t) * CEP_LD];
        // delta[t] = c[t+1] - c[t-1] (replicate edges), delta-delta = delta of delta   (feature_extraction.py:41-58,130-133)
        const float v3[3] = {c0, cp - cm, (cpp - cpm) - (cmp - cmm)};
        for (int j = t + u.shift; j < p.Tout; j += u.rep) {
          const long long off = static_cast<long long>(b) * p.sb + static_cast<long long>(j) * p.sj + static_cast<long long>(k) * p.sd;
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            const long long o2 = off + static_cast<long long>(part * NF) * p.sd;
            if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[o2] = f2bf(v3[part]);
            else reinterpret_cast<float*>(p.out)[o2] = v3[part];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 4) { fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace air_lfcc_tc

extern "C" int air_lfcc_tc_table_floats() { return air_lfcc_tc::TBL_FLOATS; }
// 16-bit elements of the DFT operand table: [Re/Im][half][kb][w_hi fp16 / w bf16 / w_lo fp16] chunks of [128 bins][32 samples]
extern "C" int air_lfcc_tc_wmat_elems() { return 2 * 2 * air_lfcc_tc::KBLK * 3 * (air_lfcc_tc::CHUNK / 2); }

extern "C" int air_lfcc_fill(const int* lengths, int B, int L, void* out, long long sb, long long sj, long long sd,
                             int out_bf16, int Tout, int feat_len, int pad_mode, const float* silence, cudaStream_t stream);

extern "C" int air_lfcc_tc_fwd(const float* wave, long long ldw, const int* lengths, int B, int L,
                               const float* table, const void* wmat, void* out, long long sb, long long sj, long long sd,
                               int out_bf16, int Tout, int feat_len, int pad_mode, const int* start,
                               const float* silence, float preemph, int num_sms, cudaStream_t stream) {
  using namespace air_lfcc_tc;
  if (!wave || !table || !wmat || !out || B <= 0 || L < 0 || Tout <= 0) return AIR_ERR_ARG;
  if (pad_mode < 0 || pad_mode > 3) return AIR_ERR_ARG;
  if (pad_mode == 3 && feat_len > 0 && !silence) return AIR_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(wmat) & 15) return AIR_ERR_UNSUPPORTED;
  const int Tmax = 1 + L / FS;
  if (feat_len > 0 && Tout != feat_len) return AIR_ERR_ARG;
  if (feat_len <= 0 && Tout < Tmax) return AIR_ERR_ARG;
  Params p;
  p.wave = wave; p.ldw = ldw; p.lengths = lengths; p.L = L; p.B = B; p.tbl = table;
  p.wmat = reinterpret_cast<const __nv_bfloat16*>(wmat);
  p.out = out; p.sb = sb; p.sj = sj; p.sd = sd; p.out_bf16 = out_bf16; p.time_minor = (sj == 1);
  p.Tout = Tout; p.feat_len = feat_len > 0 ? feat_len : 0; p.pad_mode = feat_len > 0 ? pad_mode : 0;
  p.start = start; p.preemph = preemph;
  // flat tiling needs identical frame counts and no crop
  p.flat = (lengths == nullptr && !(p.feat_len > 0 && Tmax > p.feat_len)) ? 1 : 0;
  p.T = Tmax;
  const int need = (p.feat_len > 0 && Tmax > p.feat_len) ? p.feat_len : Tmax;
  p.segs = (need + TOUT - 1) / TOUT;
  const long long tiles = p.flat ? (static_cast<long long>(B) * Tmax + TOUT - 1) / TOUT : static_cast<long long>(B) * p.segs;
  if (tiles > 0x7fffffffLL || static_cast<long long>(B) * Tmax > 0x7ffffff0LL) return AIR_ERR_UNSUPPORTED;
  p.tiles = static_cast<int>(tiles);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(lfcc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int grid = static_cast<int>(std::min<long long>(tiles, num_sms));
  lfcc_tc_kernel<<<grid, THREADS, SM_TOTAL + 1024, stream>>>(p);
  int st = air_launch_status();
  if (st == 0 && p.feat_len > 0 && (pad_mode == 1 || pad_mode == 3))
    st = air_lfcc_fill(lengths, B, L, out, sb, sj, sd, out_bf16, Tout, feat_len, pad_mode, silence, stream);
  return st;
}
