// Shifted-window convolution from a shared-memory resident input patch (TMA + tcgen05 + TMEM).
//
// The generic implicit-GEMM kernel (conv_gemm.cu) gathers every activation once per tap (9x for a 3x3) through
// the LSU.  Here a work item is R = 2 rows x 128 columns of an output pixel grid of one image; the
// (R+2) x 130 pixel input patch it needs (one block of <= 64 channels at a time) is brought into shared memory
// ONCE by a single TMA box load (out-of-bounds rows / columns are zero filled by the TMA unit = the zero padding of
// the convolution) and every tap is a shifted window of it: the patch image is "one swizzled row per pixel", so
// moving the UMMA A operand by one pixel is +row_bytes on the descriptor start address, by one patch row +130 rows.
//
//   out[b, oh(g), ow(g'), n] = sum_{t < ntaps} sum_c a[b, g + org_h + dr[t], g' + org_w + dc[t], c] * Wp[slice[t]][n][c]
//                              (+ res) (ReLU),          oh(g) = g * osh + oph,  ow(g') = g' * osw + opw
//
// With the tap table of a 3x3 / stride 1 / pad 1 layer (org = -1, dr/dc = 0..2, os = 1) this is the forward
// convolution of resnet.py:56-60 (mode-0 weights) or its data gradient (mode-1 weights: flipped taps, swapped
// channels).  The data gradient of a STRIDE-2 3x3 layer is four such launches, one per output parity class
// (oph, opw): each class only sees the 1 / 2 / 2 / 4 taps that are structurally non-zero for it, so no
// zero-stuffed gather is ever made; a 1x1 layer is the single-tap case.
//
// Roles (384 threads): warp 0 issues the patch TMA loads, warp 1 streams pre-swizzled weight slices (one
// bulk copy per (channel block, tap); all slices stay resident when they fit), warp 2 issues tcgen05.mma
// (M = 128 pixels, N = output channels, K = 16 per instruction; the whole warp runs the warp-uniform loop so
// descriptors live in uniform registers, one elected lane issues), warps 4-11 drain the double-buffered TMEM
// accumulators: warp w owns TMEM lanes 32 (w % 4) .. +31 (one pixel per lane) and the 32-channel blocks of parity
// (w - 4) / 4, so every scheduler holds two epilogue warps whose global-load / TMEM / shuffle latencies overlap
// (ncu of the four-warp version: the epilogue warps issued 21 % of the time and bounded the N = 64 layers).
// Persistent grid.
#include <algorithm>
#include "common.cuh"
#include "tc05.cuh"
#include "tmap.cuh"
#include "pack.cuh"

namespace air_patch {
using namespace tc05;

constexpr int TW = 128;                 // output columns per row segment (UMMA M)
constexpr int PW = TW + 2;              // patch width
constexpr int R = 2;                    // output rows per work item
constexpr int PR = R + 2;               // patch rows of a 3-row kernel (p.pr = R + the largest row offset of the taps)
constexpr int PW_MAX = TW + 8;          // widest patch (dilated 1-D convolutions)
constexpr int THREADS = 384;            // 4 role warps + 2 x 4 epilogue warps
constexpr int MAX_PSTAGES = 4;           // patch ring: 2 stages of a 64-channel block or 4 of a 32-channel block
constexpr int MAX_TAPS = 9;
constexpr uint32_t STAGE_TILE = 2048;    // one epilogue warp's output tile: 32 pixels x 32 channels bf16

struct PatchParams {
  int B, GH, GW;                        // images, output pixel grid enumerated by the items
  int OH, OW, osh, osw, oph, opw;       // output tensor geometry: grid (g, g') -> pixel (g*osh + oph, g'*osw + opw)
  int org_h, org_w;                     // patch origin relative to (first grid row, first grid column) of the item
  int pw;                               // patch width in pixels: TW + 2 (3x3) or TW + 8 (dilated 1-D taps up to +8)
  int C, N;
  const __nv_bfloat16* wpk; int wtaps;  // packed weights: [C/CB][wtaps] slices of [N][CB]
  void* out; long long out_ld;          // bf16, or fp32 when f32 != 0 (out, res and out2 share the type)
  const void* res; long long res_ld; int relu;
  const float* bias;                    // optional per-channel bias (ECAPA convs, ecapa_tdnn.py:39,50)
  void* out2; long long out2_ld;        // optional second output: accumulator (+bias) WITHOUT the residual
  const float* post_scale; const float* post_shift;   // optional per-channel affine applied after the ReLU, BEFORE out2 / res
  int f32;                              // fp32 parity mode: outputs / residual stored as float (operands stay bf16 splits)
  double* stats;                        // optional: stats[n] += sum of the stored outputs, stats[N + n] += sum of squares
                                        // (the batch statistics of the BatchNorm that consumes `out`, bn_stats fused)
  int CB, NCB, WT, HP;                  // channels per block (16 / 32 / 64), #blocks, column tiles, row pairs
  int row_bytes, layout;                // CB * 2; UMMA layout type (6 / 4 / 2)
  uint32_t pstage_bytes, bslot_bytes, bslot_stride;
  int pstages;                          // patch ring depth
  int pr;                               // patch rows actually loaded: R + max tap row offset (2 for 1-D / 1x1 layers)
  int stage_out;                        // 1 / 2: bf16 outputs leave through 1 / 2 shared-memory tiles per epilogue warp + TMA stores
  int nb_slots, resident;               // weight ring
  int acc_stages;                       // 1 or 2
  int ntaps; int tap_off[MAX_TAPS]; int tap_slice[MAX_TAPS];     // window offset in patch pixels, weight slice
  // flat mode (flat_h > 0): the images of the batch are treated as ONE image of B * flat_h rows, so that a work item may
  // pair the last row of an image with the first row of the next one (odd heights: 5 rows = 3 pairs per image -> 2.5);
  // the taps that would cross the image boundary are skipped per row: tap row offset tap_dr[t] is applied to image row h
  // only when 0 <= h + org_h + tap_dr[t] < flat_h.  (Otherwise every tap applies and the TMA zero fill is the padding.)
  int flat_h; int tap_dr[MAX_TAPS];
  uint32_t items;
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const bf16x8& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.u.x), "r"(v.u.y), "r"(v.u.z), "r"(v.u.w) : "memory");
}

// residual of one 32-column (NC = 32) or 16-column chunk of a pixel: 16-byte loads, issued ahead of use
template <int NC>
__device__ __forceinline__ void load_res(const PatchParams& p, long long pixel, int c0, bool valid, bf16x8 (&rv)[4]) {
  if (p.res != nullptr && valid && !p.f32) {
    const bf16x8* rp = reinterpret_cast<const bf16x8*>(static_cast<const __nv_bfloat16*>(p.res) + pixel * p.res_ld + c0);
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) rv[i] = rp[i];
  }
}

// Transpose-reduce over the warp: every lane holds 32 values v[0..31] (its pixel's channels); on return lane l holds
// sum over the 32 lanes of v[l] (31 shuffles: each stage exchanges half of the remaining values).
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// bf16 epilogue arithmetic of NC accumulator columns of one (valid) pixel, in the order the layer needs:
//   default:       v += bias;  out2 <- v;  v += res;  ReLU
//   affine_first:  v += bias;  ReLU;  v = v * post_scale + post_shift;  out2 <- v;  v = round_bf16(v) + res
// (affine_first = conv -> ReLU -> eval-mode BatchNorm folded into the epilogue, ecapa_tdnn.py:73-83: out2 is the branch
// output, out = branch output + next split is the next branch's input.)
template <int NC>
__device__ __forceinline__ void epilogue_math(const PatchParams& p, long long pixel, int c0, const bf16x8 (&rv)[4], float (&v)[NC]) {
  if (p.bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(p.bias + c0);
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      const float4 b4 = __ldg(bp + i);
      v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
    }
  }
  if (p.post_scale != nullptr) {
    const float4* sp = reinterpret_cast<const float4*>(p.post_scale + c0);
    const float4* tp = reinterpret_cast<const float4*>(p.post_shift + c0);
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) {
      const float4 s4 = __ldg(sp + i), t4 = __ldg(tp + i);
      const float sc[4] = {s4.x, s4.y, s4.z, s4.w}, sh[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = v[4 * i + e];
        if (p.relu) t = fmaxf(t, 0.f);
        v[4 * i + e] = fmaf(t, sc[e], sh[e]);
      }
    }
  }
  if (p.out2 != nullptr) {
    bf16x8* o2 = reinterpret_cast<bf16x8*>(static_cast<__nv_bfloat16*>(p.out2) + pixel * p.out2_ld + c0);
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
      const bf16x8 pk = pack8(v + i * 8);
      o2[i] = pk;
      if (p.post_scale != nullptr) unpack8(pk, v + i * 8);       // the consumer adds the ROUNDED branch output
    }
  }
  if (p.res != nullptr) {
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
      float rf[8];
      unpack8(rv[i], rf);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i * 8 + e] += rf[e];
    }
  }
  if (p.relu && p.post_scale == nullptr) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.f);
  }
}

// One chunk of NC = 32 / 16 accumulator columns of a pixel: (+bias) (-> out2) (+res) (ReLU) -> bf16 -> out.  With
// keep_vals the array v holds, on return, the STORED (bf16-rounded) values as floats, zeros for an invalid pixel.
// stage != 0 (NC = 32, bf16): the values go to the warp's shared-memory tile ([32 pixels][32 channels], SWIZZLE_64B) and the
// caller issues one TMA store per tile.  A direct store is 16 bytes per lane into 32 different 128-byte lines = 32 L1
// wavefronts per instruction, and those wavefronts share the L1 data pipe with the tensor core's operand reads (ncu: LSU
// 63 % + tensor 42 % of the pipe on the N = 64 layers); the tile costs 4 conflict-free wavefronts per instruction.
template <int NC>
__device__ __forceinline__ void epilogue_chunk(const PatchParams& p, uint32_t taddr, long long pixel, int c0, bool valid,
                                               const bf16x8 (&rv)[4], float (&v)[NC], bool keep_vals, uint32_t stage = 0) {
  if (NC == 32) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v);
  if (p.f32) {
    // fp32 parity mode: the same epilogue with float storage (no rounding point between the layers)
    if (valid) {
      if (p.bias != nullptr) {
        const float4* bp = reinterpret_cast<const float4*>(p.bias + c0);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
          const float4 b4 = __ldg(bp + i);
          v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
        }
      }
      if (p.out2 != nullptr) {
        float4* o2 = reinterpret_cast<float4*>(static_cast<float*>(p.out2) + pixel * p.out2_ld + c0);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) o2[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      if (p.res != nullptr) {
        const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.res) + pixel * p.res_ld + c0);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
          const float4 r4 = rp[i];
          v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
        }
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pixel * p.out_ld + c0);
#pragma unroll
      for (int i = 0; i < NC / 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else if (keep_vals) {
#pragma unroll
      for (int i = 0; i < NC; ++i) v[i] = 0.f;
    }
    return;
  }
  if (valid) epilogue_math<NC>(p, pixel, c0, rv, v);
  if (NC == 32 && stage != 0) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) { if (p.stage_out == 2) bulk_wait_read_1(); else bulk_wait_read_all(); }   // this tile's previous store has left shared memory
    __syncwarp();
    const uint32_t row = stage + static_cast<uint32_t>(lane) * 64u, sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
      const bf16x8 pk = pack8(v + i * 8);
      st_shared_v4(row + ((static_cast<uint32_t>(i) ^ sw) << 4), pk);
      if (keep_vals) {
        if (valid) unpack8(pk, v + i * 8);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[i * 8 + e] = 0.f;
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    return;
  }
  if (valid) {
    bf16x8* op = reinterpret_cast<bf16x8*>(static_cast<__nv_bfloat16*>(p.out) + pixel * p.out_ld + c0);
#pragma unroll
    for (int i = 0; i < NC / 8; ++i) {
      const bf16x8 pk = pack8(v + i * 8);
      op[i] = pk;
      if (keep_vals) unpack8(pk, v + i * 8);
    }
  } else if (keep_vals) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = 0.f;
  }
}

struct MmaCtx {
  const PatchParams& p;
  uint32_t sP, sB, tmem_base;
  uint64_t *full_p, *empty_p, *full_b, *empty_b, *tfull, *tempty;
};

// The MMA warp.  NT = taps (0: run-time p.ntaps), KKT = K = 16 steps per channel block (0: run-time), RES = all weight slices
// resident (waited for once) or streamed through the slot ring, FLAT = flat row mode (per-row tap masks; compiled out otherwise).
template <int NT, int KKT, bool RES, bool FLAT>
__device__ __forceinline__ void mma_role(const MmaCtx& c) {
  const PatchParams& p = c.p;
  const bool leader = elect_one();
  const int ntaps = NT > 0 ? NT : p.ntaps;
  const int KK = KKT > 0 ? KKT : (p.CB >> 4);
  const int NCB = p.NCB, N = p.N, GH = p.GH, acc_stages = p.acc_stages, nb_slots = p.nb_slots;
  const uint32_t WT = p.WT, HP = p.HP, items = p.items, PSTAGES = static_cast<uint32_t>(p.pstages);
  const uint32_t pstage16 = p.pstage_bytes >> 4, bslot16 = p.bslot_stride >> 4;
  const uint32_t idesc = instr_desc_bf16(TW, N, 0, 0);
  const uint32_t rb16 = static_cast<uint32_t>(p.row_bytes) >> 4;                  // row stride in 16-byte units
  // descriptor high word: SBO = 8 rows, version 1, swizzle mode
  const uint64_t desc_hi = static_cast<uint64_t>(((8u * p.row_bytes) >> 4) | (1u << 14) | (static_cast<uint32_t>(p.layout) << 29)) << 32;
  const uint32_t row16 = static_cast<uint32_t>(p.pw) * rb16;                      // one patch row
  uint32_t tap16[NT > 0 ? NT : MAX_TAPS];                                         // window offsets in 16-byte units
#pragma unroll
  for (int t = 0; t < (NT > 0 ? NT : MAX_TAPS); ++t) tap16[t] = static_cast<uint32_t>(p.tap_off[t]) * rb16;
  const uint32_t sP16 = ((c.sP >> 4) & 0x3FFF) | (1u << 16), sB16 = ((c.sB >> 4) & 0x3FFF) | (1u << 16);
  if (RES) {                                      // every slice is loaded exactly once: wait for all of them up front
    for (int s = 0; s < NCB * ntaps; ++s) mbar_wait(&c.full_b[s], 0);
    fence_after_sync();
  }
  uint32_t pstage = 0, pphase = 0, slot = 0, bphase = 0, it = 0;
  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x, ++it) {
    const uint32_t hp = (item / WT) % HP;
    const bool two = GH - static_cast<int>(hp) * R >= 2;
    uint32_t m0 = 7u, m1 = 7u;                    // per-row masks of the tap row offsets that apply (flat mode)
    if (FLAT) {
      const int h0 = static_cast<int>(hp * R) % p.flat_h, h1 = static_cast<int>(hp * R + 1) % p.flat_h;
      m0 = 0u; m1 = 0u;
#pragma unroll
      for (int dr = 0; dr < 3; ++dr) {
        if (h0 + p.org_h + dr >= 0 && h0 + p.org_h + dr < p.flat_h) m0 |= 1u << dr;
        if (h1 + p.org_h + dr >= 0 && h1 + p.org_h + dr < p.flat_h) m1 |= 1u << dr;
      }
    }
    const uint32_t acc = acc_stages == 2 ? (it & 1) : 0;
    const uint32_t acc_phase = acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
    mbar_wait(&c.tempty[acc], acc_phase ^ 1);
    fence_after_sync();
    const uint32_t d0 = c.tmem_base + acc * R * N, d1 = d0 + N;
    uint32_t acc0 = 0u, acc1 = 0u;
    for (int cb = 0; cb < NCB; ++cb) {
      mbar_wait(&c.full_p[pstage], pphase);
      fence_after_sync();
      const uint32_t a_lo = sP16 + pstage * pstage16;
      const uint32_t b_cb = sB16 + static_cast<uint32_t>(cb * ntaps) * bslot16;
#pragma unroll
      for (int t = 0; t < (NT > 0 ? NT : MAX_TAPS); ++t) {
        if (NT == 0 && t >= ntaps) break;
        uint32_t b_lo;
        if (RES) {
          b_lo = b_cb + static_cast<uint32_t>(t) * bslot16;
        } else {
          mbar_wait(&c.full_b[slot], bphase);
          fence_after_sync();
          b_lo = sB16 + slot * bslot16;
        }
        const uint64_t ad0 = desc_hi | (a_lo + tap16[t]), ad1 = ad0 + row16, bd = desc_hi | b_lo;
        const uint32_t drt = FLAT ? static_cast<uint32_t>(p.tap_dr[t]) : 0u;
        const bool do0 = !FLAT || ((m0 >> drt) & 1u), do1 = two && (!FLAT || ((m1 >> drt) & 1u));
        // acc0 / acc1: has this accumulator row been written in this item yet?  (A skipped tap contributes the zeros the TMA
        // fill would have contributed, so flat mode gives bit-identical sums.)
        if (leader) {
          if (KKT > 0) {
            if (do0) {
#pragma unroll
              for (int kk = 0; kk < KKT; ++kk) mma_bf16(d0, ad0 + 2 * kk, bd + 2 * kk, idesc, kk != 0 ? 1u : (FLAT ? acc0 : (t != 0 ? 1u : static_cast<uint32_t>(cb != 0))));
            }
            if (do1) {
#pragma unroll
              for (int kk = 0; kk < KKT; ++kk) mma_bf16(d1, ad1 + 2 * kk, bd + 2 * kk, idesc, kk != 0 ? 1u : (FLAT ? acc1 : (t != 0 ? 1u : static_cast<uint32_t>(cb != 0))));
            }
          } else {
            if (do0)
              for (int kk = 0; kk < KK; ++kk) mma_bf16(d0, ad0 + 2 * kk, bd + 2 * kk, idesc, kk != 0 ? 1u : acc0);
            if (do1)
              for (int kk = 0; kk < KK; ++kk) mma_bf16(d1, ad1 + 2 * kk, bd + 2 * kk, idesc, kk != 0 ? 1u : acc1);
          }
          if (!RES) mma_commit(&c.empty_b[slot]);
        }
        if (do0) acc0 = 1u;
        if (do1) acc1 = 1u;
        if (!RES) { if (++slot == static_cast<uint32_t>(nb_slots)) { slot = 0; bphase ^= 1; } }
      }
      if (leader) mma_commit(&c.empty_p[pstage]);
      if (++pstage == PSTAGES) { pstage = 0; pphase ^= 1; }
    }
    if (leader) mma_commit(&c.tfull[acc]);
  }
}

__global__ void __launch_bounds__(THREADS, 1) conv_patch_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmout,
                                                                const PatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;      // swizzle patterns are anchored at 1024 B
  const uint32_t sP = sbase;
  const uint32_t PSTAGES = static_cast<uint32_t>(p.pstages);
  const uint32_t sB = sbase + PSTAGES * p.pstage_bytes;
  uint8_t* gen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sS = sB + static_cast<uint32_t>(p.nb_slots) * p.bslot_stride;       // [8 epilogue warps][2 KB] output tiles
  const uint32_t stage_total = static_cast<uint32_t>(p.stage_out) * 8u * STAGE_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + PSTAGES * p.pstage_bytes + static_cast<size_t>(p.nb_slots) * p.bslot_stride + stage_total);
  uint64_t* full_p = bars;                          // [PSTAGES] expect_tx (TMA)
  uint64_t* empty_p = bars + PSTAGES;               // [PSTAGES] tcgen05.commit
  uint64_t* full_b = bars + 2 * PSTAGES;            // [nb_slots] expect_tx
  uint64_t* empty_b = full_b + p.nb_slots;          // [nb_slots] tcgen05.commit
  uint64_t* tfull = empty_b + p.nb_slots;           // [2]
  uint64_t* tempty = tfull + 2;                     // [2] 256 epilogue threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* s_stats = p.stats ? reinterpret_cast<float*>(tmem_slot + 4) : nullptr;      // [4 TMEM lane quarters][2][N] partial sums (each quarter's two warps own disjoint channel blocks)

  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(p.acc_stages * R * p.N)) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < PSTAGES; ++s) { mbar_init(&full_p[s], 1); mbar_init(&empty_p[s], 1); }
    for (int s = 0; s < p.nb_slots; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 256); }
    fence_barrier_init();
    tma_prefetch_desc(&tmap);
    if (p.stage_out) tma_prefetch_desc(&tmout);
  }
  if (warp == 2) tmem_alloc(tmem_slot, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const int nslices = p.NCB * p.ntaps;
  const uint32_t WT = p.WT, HP = p.HP;

  if (warp == 0) {
    // ===================== patch loads: one TMA box per (item, channel block) =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t item = blockIdx.x; item < p.items; item += gridDim.x) {
        const uint32_t wt = item % WT, r1 = item / WT;            // coordinates are ready BEFORE the slot frees up
        const int c1 = static_cast<int>(wt) * TW + p.org_w, c2 = static_cast<int>(r1 % HP) * R + p.org_h;
        const int c3 = static_cast<int>(r1 / HP);
        for (int cb = 0; cb < p.NCB; ++cb) {
          mbar_wait(&empty_p[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_p[stage], static_cast<uint32_t>(p.pr * p.pw) * p.row_bytes);
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
          const int cb = s / p.ntaps, t = s - cb * p.ntaps;
          mbar_arrive_expect_tx(&full_b[s], p.bslot_bytes);
          bulk_g2s(sB + s * p.bslot_stride, p.wpk + (cb * p.wtaps + p.tap_slice[t]) * slice_elems, p.bslot_bytes, &full_b[s]);
        }
      } else {
        uint32_t slot = 0, phase = 0;
        for (uint32_t item = blockIdx.x; item < p.items; item += gridDim.x) {
          for (int s = 0; s < nslices; ++s) {
            const int cb = s / p.ntaps, t = s - cb * p.ntaps;
            mbar_wait(&empty_b[slot], phase ^ 1);
            mbar_arrive_expect_tx(&full_b[slot], p.bslot_bytes);
            bulk_g2s(sB + slot * p.bslot_stride, p.wpk + (cb * p.wtaps + p.tap_slice[t]) * slice_elems, p.bslot_bytes, &full_b[slot]);
            if (++slot == static_cast<uint32_t>(p.nb_slots)) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
    // Specialised on (taps, K steps per channel block, resident weights) so that the tap / row / K loops are straight-line
    // code with immediate offsets: the generic loop costs ~18 instructions per MMA (ncu: 81-110 cycles between MMAs of
    // 32-64 cycles), which bounded every layer with N <= 128.
    const int KK = p.CB >> 4;
    const MmaCtx c{p, sP, sB, tmem_base, full_p, empty_p, full_b, empty_b, tfull, tempty};
    if (p.flat_h > 0) {                               // odd-height 3x3 layers (layer 2 / layer 3 geometries)
      if (p.resident && KK == 4 && p.ntaps == 9) mma_role<9, 4, true, true>(c);
      else if (!p.resident && KK == 4 && p.ntaps == 9) mma_role<9, 4, false, true>(c);
      else if (p.resident) mma_role<0, 0, true, true>(c);
      else mma_role<0, 0, false, true>(c);
    }
    else if (p.resident && KK == 4 && p.ntaps == 9) mma_role<9, 4, true, false>(c);
    else if (p.resident && KK == 4 && p.ntaps == 4) mma_role<4, 4, true, false>(c);
    else if (p.resident && KK == 4 && p.ntaps == 3) mma_role<3, 4, true, false>(c);
    else if (p.resident && KK == 4 && p.ntaps == 2) mma_role<2, 4, true, false>(c);
    else if (p.resident && KK == 4 && p.ntaps == 1) mma_role<1, 4, true, false>(c);
    else if (p.resident && KK == 1 && p.ntaps == 9) mma_role<9, 1, true, false>(c);
    else if (p.resident && KK == 1 && p.ntaps == 1) mma_role<1, 1, true, false>(c);
    else if (!p.resident && KK == 4 && p.ntaps == 9) mma_role<9, 4, false, false>(c);
    else if (!p.resident && KK == 4 && p.ntaps == 4) mma_role<4, 4, false, false>(c);
    else if (!p.resident && KK == 4 && p.ntaps == 2) mma_role<2, 4, false, false>(c);
    else if (p.resident) mma_role<0, 0, true, false>(c);
    else mma_role<0, 0, false, false>(c);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3, set = (warp - 4) >> 2;      // TMEM lane quarter; parity of the 32-channel blocks this warp drains
    const int m = q * 32 + lane;
    const uint32_t tile0 = p.stage_out ? sS + static_cast<uint32_t>(warp - 4) * static_cast<uint32_t>(p.stage_out) * STAGE_TILE : 0u;
    uint32_t tsel = 0;                             // which of the warp's (1 or 2) tiles the next unit uses
    // fused BatchNorm statistics: lane l accumulates channel (32 cb + l) of its warp's pixels in registers, in a fixed
    // order (bit-reproducible run to run); combined per CTA in shared memory and across CTAs with fp64 atomics
    float acc_s[4], acc_q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc_s[u] = 0.f; acc_q[u] = 0.f; }
    uint32_t it = 0;
    for (uint32_t item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int wt = static_cast<int>(item % WT);
      const uint32_t r1 = item / WT;
      const int hp = static_cast<int>(r1 % HP), b = static_cast<int>(r1 / HP);
      const int rows = min(R, p.GH - hp * R);
      const uint32_t acc = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = p.acc_stages == 2 ? ((it >> 1) & 1) : (it & 1);
      const int g = wt * TW + m;
      const int ow = g * p.osw + p.opw;
      const bool col_ok = g < p.GW && ow < p.OW;
      const int cpr = p.N >> 5, tail = p.N & 31;                       // full 32-column chunks per row, 16-column tail
      static_assert(R == 2, "row select below assumes two rows per item");
      long long pixr[R]; bool valr[R];
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int oh = (hp * R + j) * p.osh + p.oph;
        valr[j] = col_ok && j < rows && oh < p.OH;
        pixr[j] = (static_cast<long long>(b) * p.OH + oh) * p.OW + ow;
      }
      const long long pix0 = pixr[0], pix1 = pixr[1];
      const bool val0 = valr[0], val1 = valr[1];
      // column blocks of 32 channels (this warp: blocks set, set + 2, ...); both rows of a block are processed together
      // so that the fused BatchNorm statistics need one warp transpose-reduce per block instead of one per (row, block).
      // The residual of the first block is requested before the accumulator is even complete, the one of block i + 2
      // while block i is processed.
      bf16x8 ra[4], rb[4];
      if (set < cpr) { load_res<32>(p, pix0, set << 5, val0, ra); if (rows > 1) load_res<32>(p, pix1, set << 5, val1, rb); }
      mbar_wait(&tfull[acc], acc_phase);
      fence_after_sync();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * R * p.N;
      const bool st = s_stats != nullptr;
      for (int cb = set; cb < cpr; cb += 2) {
        const int c0 = cb << 5;
        float v0[32], v1[32];
        uint32_t my_tile = tile0 + tsel * STAGE_TILE;
        epilogue_chunk<32>(p, tacc, pix0, c0, val0, ra, v0, st, my_tile);
        if (my_tile && lane == 0) { tma_store_4d(&tmout, my_tile, c0, wt * TW + q * 32, hp * R, b); bulk_commit_group(); }
        if (p.stage_out == 2) tsel ^= 1u;
        if (rows > 1) {
          my_tile = tile0 + tsel * STAGE_TILE;
          epilogue_chunk<32>(p, tacc + p.N, pix1, c0, val1, rb, v1, st, my_tile);
          if (my_tile && lane == 0) { tma_store_4d(&tmout, my_tile, c0, wt * TW + q * 32, hp * R + 1, b); bulk_commit_group(); }
          if (p.stage_out == 2) tsel ^= 1u;
        }
        if (cb + 2 < cpr) { load_res<32>(p, pix0, c0 + 64, val0, ra); if (rows > 1) load_res<32>(p, pix1, c0 + 64, val1, rb); }
        if (st) {                                   // warp-uniform: every lane takes part in the shuffles
          if (rows > 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { const float t = v1[i]; v1[i] = fmaf(v0[i], v0[i], t * t); v0[i] += t; }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v1[i] = v0[i] * v0[i];
          }
          const float cs = warp_column_sums(v0, lane);
          const float cq = warp_column_sums(v1, lane);
#pragma unroll
          for (int u = 0; u < 4; ++u) if (u == (cb >> 1)) { acc_s[u] += cs; acc_q[u] += cq; }
        }
      }
      if (tail && set < rows) {                     // 16-column tail: row `set` of the item
        float vt[16];
        const long long pixt = set == 0 ? pix0 : pix1;
        const bool valt = set == 0 ? val0 : val1;
        load_res<16>(p, pixt, cpr << 5, valt, ra);
        epilogue_chunk<16>(p, tacc + set * p.N, pixt, cpr << 5, valt, ra, vt, false);
      }
      fence_before_sync();
      mbar_arrive(&tempty[acc]);
    }
    if (tile0 && lane == 0) bulk_wait_all();            // every output tile of this warp has been written
    if (s_stats != nullptr) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cb = 2 * u + set;
        if (cb * 32 < p.N) {
          s_stats[q * 2 * p.N + cb * 32 + lane] = acc_s[u];
          s_stats[q * 2 * p.N + p.N + cb * 32 + lane] = acc_q[u];
        }
      }
    }
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 2) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
  if (s_stats) {
    for (int i = threadIdx.x; i < 2 * p.N; i += THREADS) {
      const float t = ((s_stats[i] + s_stats[2 * p.N + i]) + s_stats[4 * p.N + i]) + s_stats[6 * p.N + i];
      atomicAdd(p.stats + i, static_cast<double>(t));
    }
  }
}

// Weight packing for the patch kernel: one pre-swizzled [N rows][CB channels] K-major image per (cb, tap), exactly
// the bytes a TMA load with the matching swizzle would have produced, so one bulk copy per slice suffices:
//   element (n, k) of slice (cb, tap) at byte  swizzle(n * CB*2 + (k / 8) * 16) + (k % 8) * 2
//   mode 0 (fprop): w is [N = Cout][taps][C = Cin];   value = w[n][tap][ch]
//   mode 1 (dgrad): w is [C = Cout][taps][N = Cin];   value = w[ch][taps - 1 - tap][n]   (flipped taps, swapped channels)
__global__ void pack_patch_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int C, int N, int CB, int taps, int mode) {
  const long long total = static_cast<long long>(taps) * C * N;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    air_pack::patch_pack_elem(w, dst, i, C, N, CB, taps, mode);
}

}  // namespace air_patch

using namespace air_patch;

// channels per block.  AIR_PATCH_CB32=1 splits 64-channel inputs into two 32-channel blocks (SWIZZLE_64B rows, a patch ring of
// four half-size stages = three loads in flight per SM instead of one): measured SLOWER on the N = 64 layers (0.52 vs 0.39 ms,
// profiles/experiments), so it is an experiment switch only.
static int patch_cb(int C) {
  static const int split64 = [] { const char* e = getenv("AIR_PATCH_CB32"); return (e && e[0] == '1') ? 1 : 0; }();
  if (C == 64 && split64) return 32;
  return C <= 64 ? C : 64;
}
extern "C" int air_conv_patch_cb(int C) { return patch_cb(C); }

// 1 when (C, N) can run on the patch kernel
extern "C" int air_conv3x3_patch_supported(int C, int N, int H, int W) {
  if (!(C == 16 || C == 32 || (C >= 64 && C % 64 == 0))) return 0;
  if (N % 16 != 0 || N > 256 || N < 16) return 0;
  if ((N * patch_cb(C) * 2) % 256 != 0) return 0;
  if (H < 1 || W < 1) return 0;
  return 1;
}

// taps = 9 (3x3) or 1 (1x1); dst holds taps*C*N bf16
extern "C" int air_conv_patch_pack_weights(const float* w, void* dst, int C, int N, int taps, int mode, cudaStream_t stream) {
  if (!w || !dst || !air_conv3x3_patch_supported(C, N, 1, 1) || (mode != 0 && mode != 1) || taps < 1 || taps > MAX_TAPS) return AIR_ERR_ARG;
  const long long total = static_cast<long long>(taps) * C * N;
  pack_patch_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 2048)), 256, 0, stream>>>(
      w, reinterpret_cast<__nv_bfloat16*>(dst), C, N, patch_cb(C), taps, mode);
  return air_launch_status();
}

extern "C" int air_conv3x3_pack_weights(const float* w, void* dst, int C, int N, int mode, cudaStream_t stream) {
  return air_conv_patch_pack_weights(w, dst, C, N, 9, mode, stream);
}

extern "C" int air_conv_patch_taps_ex2_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                            const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                            const void* res, long long res_ld, int relu, const float* bias,
                                            void* out2, long long out2_ld, double* stats,
                                            int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                            int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                            int flags, int num_sms, cudaStream_t stream);
extern "C" int air_conv_patch_taps_ex3_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                            const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                            const void* res, long long res_ld, int relu, const float* bias,
                                            const float* post_scale, const float* post_shift,
                                            void* out2, long long out2_ld, double* stats,
                                            int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                            int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                            int flags, int num_sms, cudaStream_t stream);

// The general entry point: explicit tap table and output pixel mapping (see the formula at the top of this file).
//   a: (B, Hin, Win, C) channels-last bf16;  out / res: (B, OH, OW, N);  item grid GH x GW;
//   tap t reads the window that starts at patch pixel (tap_dr[t], tap_dc[t]) (0 <= dr, dc <= 2) and uses weight slice
//   tap_slice[t] (< wtaps) of every channel block.
extern "C" int air_conv_patch_taps_ex_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                           const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                           const void* res, long long res_ld, int relu, const float* bias,
                                           void* out2, long long out2_ld, double* stats,
                                           int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                           int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                           int num_sms, cudaStream_t stream);

extern "C" int air_conv_patch_taps_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                        const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                        const void* res, long long res_ld, int relu,
                                        int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                        int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                        int num_sms, cudaStream_t stream) {
  return air_conv_patch_taps_ex_bf16(a, a_ld, B, Hin, Win, C, wpk, wtaps, N, out, out_ld, OH, OW, res, res_ld, relu, nullptr,
                                     nullptr, 0, nullptr, GH, GW, org_h, org_w, osh, osw, oph, opw, ntaps, tap_dr, tap_dc,
                                     tap_slice, num_sms, stream);
}

// as air_conv_patch_taps_bf16 plus a per-channel fp32 bias, a second output without the residual, and column offsets
// tap_dc up to 8 (the patch is then 136 pixels wide): the dilated k = 3 1-D convolutions of ecapa_tdnn.py:50
extern "C" int air_conv_patch_taps_ex_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                           const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                           const void* res, long long res_ld, int relu, const float* bias,
                                           void* out2, long long out2_ld, double* stats,
                                           int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                           int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                           int num_sms, cudaStream_t stream) {
  return air_conv_patch_taps_ex2_bf16(a, a_ld, B, Hin, Win, C, wpk, wtaps, N, out, out_ld, OH, OW, res, res_ld, relu, bias,
                                      out2, out2_ld, stats, GH, GW, org_h, org_w, osh, osw, oph, opw, ntaps, tap_dr, tap_dc,
                                      tap_slice, 0, num_sms, stream);
}

// as air_conv_patch_taps_ex_bf16 plus `flags`: AIR_CONV_F32_OUT = out / res / out2 are float tensors (fp32 parity mode: the
// operands are bf16 split terms concatenated along the channels, the accumulator is stored unrounded)
extern "C" int air_conv_patch_taps_ex2_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                            const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                            const void* res, long long res_ld, int relu, const float* bias,
                                            void* out2, long long out2_ld, double* stats,
                                            int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                            int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                            int flags, int num_sms, cudaStream_t stream) {
  return air_conv_patch_taps_ex3_bf16(a, a_ld, B, Hin, Win, C, wpk, wtaps, N, out, out_ld, OH, OW, res, res_ld, relu, bias,
                                      nullptr, nullptr, out2, out2_ld, stats, GH, GW, org_h, org_w, osh, osw, oph, opw, ntaps,
                                      tap_dr, tap_dc, tap_slice, flags, num_sms, stream);
}

// as air_conv_patch_taps_ex2_bf16 plus an optional per-channel affine applied right after the ReLU (bf16 storage only):
//   t = relu(acc + bias) * post_scale + post_shift;  out2 <- t;  out <- round_bf16(t) + res
// = conv -> ReLU -> eval-mode BatchNorm with the running-statistics affine folded in (ecapa_tdnn.py:73-83: out2 is the
// branch output, out the next branch's input); without out2 / res simply out <- t.
extern "C" int air_conv_patch_taps_ex3_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                            const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                            const void* res, long long res_ld, int relu, const float* bias,
                                            const float* post_scale, const float* post_shift,
                                            void* out2, long long out2_ld, double* stats,
                                            int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                            int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                            int flags, int num_sms, cudaStream_t stream) {
  if (!a || !wpk || !out || B <= 0 || !tap_dr || !tap_dc || !tap_slice) return AIR_ERR_ARG;
  if ((post_scale == nullptr) != (post_shift == nullptr)) return AIR_ERR_ARG;
  if (post_scale && ((flags & AIR_CONV_F32_OUT) || stats || (N % 32) != 0)) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(post_scale) | reinterpret_cast<uintptr_t>(post_shift)) & 15) return AIR_ERR_UNSUPPORTED;
  if (stats && (N % 32) != 0) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(out2)) & 15) return AIR_ERR_UNSUPPORTED;
  if (out2 && out2_ld % 8 != 0) return AIR_ERR_UNSUPPORTED;
  if (ntaps < 1 || ntaps > MAX_TAPS || wtaps < 1 || GH < 1 || GW < 1 || osh < 1 || osw < 1 || oph < 0 || opw < 0) return AIR_ERR_ARG;
  if (!air_conv3x3_patch_supported(C, N, Hin, Win)) return AIR_ERR_UNSUPPORTED;
  if (a_ld % 8 != 0 || out_ld % 8 != 0 || (res && res_ld % 8 != 0)) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(wpk) |
       reinterpret_cast<uintptr_t>(res)) & 15) return AIR_ERR_UNSUPPORTED;
  PatchParams p;
  p.B = B; p.GH = GH; p.GW = GW; p.OH = OH; p.OW = OW; p.osh = osh; p.osw = osw; p.oph = oph; p.opw = opw;
  p.org_h = org_h; p.org_w = org_w; p.C = C; p.N = N;
  p.wpk = reinterpret_cast<const __nv_bfloat16*>(wpk); p.wtaps = wtaps;
  p.out = out; p.out_ld = out_ld; p.f32 = (flags & AIR_CONV_F32_OUT) ? 1 : 0;
  p.res = res; p.res_ld = res_ld; p.relu = relu;
  p.ntaps = ntaps;
  for (int t = 0; t < MAX_TAPS; ++t) { p.tap_off[t] = 0; p.tap_slice[t] = 0; }
  int max_dc = 0;
  int max_dr = 0;
  for (int t = 0; t < ntaps; ++t) { max_dc = std::max(max_dc, tap_dc[t]); max_dr = std::max(max_dr, tap_dr[t]); }
  if (max_dr < 0 || max_dr > PR - R) return AIR_ERR_ARG;
  p.pr = R + max_dr;
  p.pw = max_dc <= PW - TW ? PW : PW_MAX;
  p.bias = bias; p.out2 = out2; p.out2_ld = out2_ld; p.stats = stats;
  p.post_scale = post_scale; p.post_shift = post_shift;
  // flat mode: a plain stride-1, same-size layer whose taps cross rows, odd image height, more than one image
  static const int flat_env = [] { const char* e = getenv("AIR_PATCH_FLAT"); return (e && e[0] == '0') ? 0 : 1; }();
  const bool flat = flat_env && B > 1 && (GH & 1) && max_dr > 0 && osh == 1 && osw == 1 && oph == 0 && opw == 0 && GH == OH &&
                    GW == OW && Hin == GH && Win == GW && org_h <= 0 && org_h + max_dr >= 0 &&
                    static_cast<long long>(B) * GH < 0x7fffffffLL;
  p.flat_h = flat ? GH : 0;
  for (int t = 0; t < MAX_TAPS; ++t) p.tap_dr[t] = 0;
  for (int k = 0; k < ntaps; ++k) {
    const int t = k;
    if (tap_dr[t] < 0 || tap_dr[t] > PR - R || tap_dc[t] < 0 || tap_dc[t] > p.pw - TW || tap_slice[t] < 0 || tap_slice[t] >= wtaps)
      return AIR_ERR_ARG;
    p.tap_off[k] = tap_dr[t] * p.pw + tap_dc[t];
    p.tap_slice[k] = tap_slice[t];
    p.tap_dr[k] = tap_dr[t];
  }
  if (flat) {                                         // one image of B * GH rows (same addresses: the images are contiguous)
    Hin = B * GH; GH = B * GH; OH = GH; B = 1;
    p.B = 1; p.GH = GH; p.OH = OH;
  }
  p.CB = patch_cb(C); p.NCB = C / p.CB; p.WT = (GW + TW - 1) / TW; p.HP = (GH + R - 1) / R;
  p.row_bytes = p.CB * 2; p.layout = p.CB == 64 ? 2 : (p.CB == 32 ? 4 : 6);
  const long long items = static_cast<long long>(B) * p.HP * p.WT;
  if (items > 0x7fffffffLL) return AIR_ERR_UNSUPPORTED;
  p.items = static_cast<uint32_t>(items);
  p.acc_stages = (2 * R * N <= 512) ? 2 : 1;
  p.pstage_bytes = static_cast<uint32_t>((p.pr * p.pw * p.row_bytes + 1023) / 1024 * 1024);
  p.bslot_bytes = static_cast<uint32_t>(N * p.row_bytes);
  p.bslot_stride = (p.bslot_bytes + 1023u) / 1024u * 1024u;
  p.pstages = p.pstage_bytes <= 36u * 1024u ? MAX_PSTAGES : 2;
  const int PSTAGES = p.pstages;
  const int nslices = p.NCB * ntaps;
  const int fixed = 3 * 1024 + (stats ? 4 * 2 * N * static_cast<int>(sizeof(float)) : 0);     // alignment, barriers, statistics
  auto plan = [&](int staging, int& slots, int& resident) {
    const int budget = 227 * 1024 - PSTAGES * static_cast<int>(p.pstage_bytes) - fixed - staging;
    slots = budget / static_cast<int>(p.bslot_stride);
    if (slots >= nslices) { slots = nslices; resident = 1; } else { resident = 0; if (slots > 12) slots = 12; }
  };
  int slots, resident, slots_s, resident_s;
  plan(0, slots, resident);
  // bf16 outputs leave through shared-memory tiles + TMA stores when the 16 KB fit without giving up resident weights
  static const int stage_env = [] { const char* e = getenv("AIR_PATCH_STAGE_OUT"); return (e && e[0] == '0') ? 0 : 1; }();
  p.stage_out = 0;
  if (stage_env && !p.f32 && N % 32 == 0) {
    plan(8 * static_cast<int>(STAGE_TILE), slots_s, resident_s);
    if (resident_s == resident && (slots_s == slots || slots_s >= 3)) {
      p.stage_out = 1; const int slots1 = slots_s;
      // a second tile per warp (the store of one drains while the next is filled) when it costs no weight slot
      plan(16 * static_cast<int>(STAGE_TILE), slots_s, resident_s);
      if (resident_s == resident && slots_s == slots1) p.stage_out = 2;
      slots = slots1;
    }
  }
  if (slots < 2 && nslices > 1) return AIR_ERR_UNSUPPORTED;
  p.nb_slots = slots; p.resident = resident;
  const size_t smem = 1024 + static_cast<size_t>(PSTAGES) * p.pstage_bytes + static_cast<size_t>(slots) * p.bslot_stride +
                      static_cast<size_t>(p.stage_out) * 8 * STAGE_TILE +
                      (2 * PSTAGES + 2 * slots + 4) * 8 + 32 + (stats ? 4 * 2 * static_cast<size_t>(N) * sizeof(float) : 0);
  CUtensorMap tm, tmo;
  int tr = air_tmap::make_act_tmap(&tm, a, a_ld, B, Hin, Win, C, p.CB, p.pw, p.pr, p.row_bytes);
  if (tr == 0 && p.stage_out) {
    // output view of the item grid: grid point (g, g') is pixel (g * osh + oph, g' * osw + opw); grid points whose pixel
    // lies outside the tensor are outside the view, so the TMA store clips them
    const int GWv = std::min(GW, (OW - opw + osw - 1) / osw), GHv = std::min(GH, (OH - oph + osh - 1) / osh);
    if (GWv < 1 || GHv < 1) return AIR_ERR_ARG;
    const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(out) + (static_cast<long long>(oph) * OW + opw) * out_ld;
    tr = air_tmap::make_act_tmap_strided(&tmo, ob, static_cast<long long>(osw) * out_ld, static_cast<long long>(osh) * OW * out_ld,
                                         static_cast<long long>(OH) * OW * out_ld, B, GHv, GWv, N, 32, 32, 1, 64);
  } else if (tr == 0) {
    tmo = tm;
  }
  if (tr != 0) return tr < 0 ? AIR_ERR_DRIVER : 10000 + tr;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int grid = static_cast<int>(std::min<long long>(items, num_sms));
  conv_patch_kernel<<<grid, THREADS, smem, stream>>>(tm, tmo, p);
  return air_launch_status();
}

// 3x3 / stride 1 / pad 1 (forward with mode-0 weights, data gradient with mode-1 weights)
extern "C" int air_conv3x3_patch_bf16(const void* a, long long a_ld, int B, int H, int W, int C,
                                      const void* wpk, int N, void* out, long long out_ld,
                                      const void* res, long long res_ld, int relu, int num_sms, cudaStream_t stream) {
  int dr[9], dc[9], sl[9];
  for (int t = 0; t < 9; ++t) { dr[t] = t / 3; dc[t] = t % 3; sl[t] = t; }
  return air_conv_patch_taps_bf16(a, a_ld, B, H, W, C, wpk, 9, N, out, out_ld, H, W, res, res_ld, relu,
                                  H, W, -1, -1, 1, 1, 0, 0, 9, dr, dc, sl, num_sms, stream);
}

// 3x3 / stride 1 / pad 1 forward that also accumulates the per-channel sum / sum of squares of its (stored) output into
// stats[0..N) / stats[N..2N) (fp64, caller zeroes): the batch statistics of the BatchNorm that follows (resnet.py:65-68)
extern "C" int air_conv3x3_patch_stats_bf16(const void* a, long long a_ld, int B, int H, int W, int C,
                                            const void* wpk, int N, void* out, long long out_ld,
                                            const void* res, long long res_ld, int relu, double* stats,
                                            int num_sms, cudaStream_t stream) {
  int dr[9], dc[9], sl[9];
  for (int t = 0; t < 9; ++t) { dr[t] = t / 3; dc[t] = t % 3; sl[t] = t; }
  return air_conv_patch_taps_ex_bf16(a, a_ld, B, H, W, C, wpk, 9, N, out, out_ld, H, W, res, res_ld, relu, nullptr, nullptr, 0,
                                     stats, H, W, -1, -1, 1, 1, 0, 0, 9, dr, dc, sl, num_sms, stream);
}

// Data gradient of a k x k (k = 3, pad 1 or k = 1, pad 0) / STRIDE-2 convolution: dy (B, Ho, Wo, Cout) -> dx (B, H, W, Cin),
// one launch per output parity class with only its structurally non-zero taps.  wpk: mode-1 packed weights (k*k taps).
// k = 3: every pixel of dx is written (+ res).  k = 1: only the even-even pixels are written (dx = res + contribution there);
// the other pixels of dx are left untouched, so pass res = dx to accumulate into an existing gradient.
extern "C" int air_conv_s2_dgrad_patch_bf16(const void* dy, long long dy_ld, int B, int Ho, int Wo, int Cout,
                                            const void* wpk, int k, int Cin, void* dx, long long dx_ld, int H, int W,
                                            const void* res, long long res_ld, int num_sms, cudaStream_t stream) {
  return air_conv_s2_dgrad_patch_ex_bf16(dy, dy_ld, B, Ho, Wo, Cout, wpk, k, Cin, dx, dx_ld, H, W, res, res_ld, 0, num_sms, stream);
}

// as above plus `flags` (AIR_CONV_F32_OUT: dx / res are float tensors, fp32 parity mode)
extern "C" int air_conv_s2_dgrad_patch_ex_bf16(const void* dy, long long dy_ld, int B, int Ho, int Wo, int Cout,
                                               const void* wpk, int k, int Cin, void* dx, long long dx_ld, int H, int W,
                                               const void* res, long long res_ld, int flags, int num_sms, cudaStream_t stream) {
  if (k != 3 && k != 1) return AIR_ERR_UNSUPPORTED;
  for (int ph = 0; ph < 2; ++ph) {
    for (int pw = 0; pw < 2; ++pw) {
      if (k == 1 && (ph | pw)) continue;
      const int GH = (H - ph + 1) / 2, GW = (W - pw + 1) / 2;
      if (GH < 1 || GW < 1) continue;
      int dr[9], dc[9], sl[9], nt = 0;
      if (k == 1) { dr[0] = 0; dc[0] = 0; sl[0] = 0; nt = 1; }
      else {
        // forward: y[ho][wo] += x[2ho + i - 1][2wo + j - 1] * W[i][j]  =>  dx[h][w] += dy[(h+1-i)/2][(w+1-j)/2] * W[i][j]
        for (int i = 0; i < 3; ++i) {
          if ((ph + 1 - i) & 1) continue;
          for (int j = 0; j < 3; ++j) {
            if ((pw + 1 - j) & 1) continue;
            dr[nt] = (ph + 1 - i) / 2; dc[nt] = (pw + 1 - j) / 2;
            sl[nt] = 8 - (3 * i + j);                       // mode-1 packing stores original tap t at slice 8 - t
            ++nt;
          }
        }
      }
      const int st = air_conv_patch_taps_ex2_bf16(dy, dy_ld, B, Ho, Wo, Cout, wpk, k * k, Cin, dx, dx_ld, H, W, res, res_ld, 0,
                                                  nullptr, nullptr, 0, nullptr, GH, GW, 0, 0, 2, 2, ph, pw, nt, dr, dc, sl,
                                                  flags, num_sms, stream);
      if (st != 0) return st;
    }
  }
  return AIR_OK;
}
