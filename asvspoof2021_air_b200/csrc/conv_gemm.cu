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
// A is never materialised: 256 producer threads (eight gather warps) gather the 128 x 64 im2col tile of a K block with
// zero-filling 16-byte cp.async straight into the SWIZZLE_128B K-major operand image (one 128-byte row per output pixel).
// Eight consecutive lanes copy the eight 16-byte chunks of ONE row, so a warp instruction reads four contiguous
// 128-byte segments of global memory and writes four consecutive shared-memory rows (the first one-row-per-lane
// mapping cost 64 LSU wavefronts per instruction and made the gather the bottleneck).
// Pre-packed, pre-swizzled weight tiles arrive with one bulk copy per K block on the TMA engine.  One elected
// thread issues tcgen05.mma (M = 128, N = block_n <= 256, K = 16 x 4 per stage); accumulators are
// double-buffered in TMEM so the epilogue warps (TMEM -> registers -> bias/residual/ReLU -> bf16 ->
// HBM) overlap the next tile's main loop.  The kernel is persistent: grid = #SMs, tiles strided.
#include <algorithm>
#include "common.cuh"
#include "tc05.cuh"
#include "pack.cuh"
#include "tmap.cuh"

namespace air_gemm {
using namespace tc05;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB
constexpr int GATHER_WARPS = 8;                           // ncu: with four, each gather warp issued one instruction per ~4.5 cycles
                                                          // all tile long (a single warp's limit) while the tensor pipe idled at 18 %
constexpr int NUM_A_THREADS = 32 * GATHER_WARPS;
constexpr int ROW_STEP = NUM_A_THREADS / 8;               // thread t: chunk t & 7 of rows (t >> 3) + ROW_STEP i
constexpr int ROWS_PER_THREAD = 128 / ROW_STEP;
constexpr uint32_t STAGE_TILE = 2048;                     // one epilogue warp's output tile: 32 rows x 32 channels bf16
constexpr int W_LOAD = GATHER_WARPS, W_MMA = GATHER_WARPS + 1;      // warp indices of the weight / TMA loader and the MMA issuer
constexpr int THREADS = 32 * (GATHER_WARPS + 6);          // gather warps, loader warp, MMA warp, 4 epilogue warps

struct ConvParams {
  const __nv_bfloat16* a; long long a_ld;   // gather source, elements between consecutive pixels
  int H, W, C;                              // gather-source geometry, C channels per tap (multiple of 8)
  int Ho, Wo;                               // pixel grid enumerated by m
  int kh, kw, sh, sw, ph, pw, dh, dw;
  int mode;                                 // 0: fprop gather; 1: transposed gather (dgrad)
  const __nv_bfloat16* wpk;                 // packed weights [n_tile][k_block][8 chunks][block_n][8]
  int N, K, KB, block_n, n_tiles;
  long long M; int m_tiles;
  void* out; long long out_ld;              // bf16, or fp32 when flags & AIR_CONV_F32_OUT (out, res, out2 share the type)
  const float* bias; const void* res; long long res_ld; int relu;
  int bias_rows;                            // 0: bias[n]; > 0: bias[(m / bias_rows) * N + n] (per-utterance bias)
  void* out2; long long out2_ld;            // optional second output: the accumulator (+bias) WITHOUT the residual
  const float* post_scale; const float* post_shift;   // optional per-channel affine AFTER the ReLU (eval-mode BatchNorm of conv -> ReLU -> BN)
  int stages; int flags;
  int use_tma;                              // 1x1 / stride 1: the A tile is a plain 2-D box of the activation matrix
  int stage_out;                            // 1 / 2: bf16 outputs leave through 1 / 2 shared-memory tiles per epilogue warp + TMA stores
  double* stats;                            // optional (staged path only): stats[n] += sum of the stored outputs, stats[N + n] += sum of
                                            // squares -- the batch statistics of the BatchNorm that consumes `out` (bn_stats fused)
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, const bf16x8& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.u.x), "r"(v.u.y), "r"(v.u.z), "r"(v.u.w) : "memory");
}

// Transpose-reduce over the warp: every lane holds 32 values v[0..31] (its row's columns); on return lane l holds the sum over
// the 32 lanes of v[l] (31 shuffles: each stage exchanges half of the remaining values).  As in conv_patch.cu.
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

// bias / second output / residual / ReLU / affine of 16 accumulator columns of row m (the part of the epilogue before the store)
__device__ __forceinline__ void epilogue_math16(const ConvParams& p, long long m, int n0, bool f32, float (&v)[16]) {
  if (p.bias) {                                  // 16 consecutive floats, 16-byte aligned (checked by the launcher)
    const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + (p.bias_rows > 0 ? (m / p.bias_rows) * p.N : 0));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b4 = __ldg(bp + i);
      v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
    }
  }
  if (p.out2) {
    if (f32) {
      float4* o2 = reinterpret_cast<float4*>(static_cast<float*>(p.out2) + m * p.out2_ld + n0);
#pragma unroll
      for (int i = 0; i < 4; ++i) o2[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
      bf16x8* o2 = reinterpret_cast<bf16x8*>(static_cast<__nv_bfloat16*>(p.out2) + m * p.out2_ld + n0);
      o2[0] = pack8(v);
      o2[1] = pack8(v + 8);
    }
  }
  if (p.res) {
    if (f32) {
      const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.res) + m * p.res_ld + n0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 r4 = rp[i];
        v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
      }
    } else {
      const bf16x8* rp = reinterpret_cast<const bf16x8*>(static_cast<const __nv_bfloat16*>(p.res) + m * p.res_ld + n0);
      float rf[16];
      unpack8(rp[0], rf); unpack8(rp[1], rf + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += rf[i];
    }
  }
  if (p.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (p.post_scale) {
    const float4* sp = reinterpret_cast<const float4*>(p.post_scale + n0);
    const float4* tp = reinterpret_cast<const float4*>(p.post_shift + n0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 s4 = __ldg(sp + i), t4 = __ldg(tp + i);
      v[4 * i] = fmaf(v[4 * i], s4.x, t4.x); v[4 * i + 1] = fmaf(v[4 * i + 1], s4.y, t4.y);
      v[4 * i + 2] = fmaf(v[4 * i + 2], s4.z, t4.z); v[4 * i + 3] = fmaf(v[4 * i + 3], s4.w, t4.w);
    }
  }
}

// Epilogue of one warp: TMEM lane quadrant q, columns [c_begin, c_end) of every tile -> bias / residual / ReLU / affine ->
// bf16 -> HBM.  In TMA mode the four (otherwise idle) gather warps take the upper half of the columns.
// stage_tile != 0: bf16 outputs leave 32 columns at a time through this warp's shared-memory tile ([32 rows][32 channels],
// SWIZZLE_64B) and one TMA store -- a direct store is 16 bytes per lane into 32 different 128-byte lines, 32 wavefronts of the
// L1 data pipe the tensor core reads its operands through (the K <= 512 layers were bound by it, not by the tensor pipe).
__device__ __forceinline__ void epilogue_role(const ConvParams& p, uint32_t tmem_base, uint64_t* tfull, uint64_t* tempty,
                                              int q, int lane, int c_begin, int c_end, int total_tiles,
                                              uint32_t stage_tile, const CUtensorMap* tmo) {
  const int row = q * 32 + lane;
  const bool f32 = (p.flags & AIR_CONV_F32_OUT) != 0;       // fp32 parity mode: float storage, no rounding point
  const uint32_t ssw = static_cast<uint32_t>((lane >> 1) & 3);
  uint32_t tsel = 0;                               // which of the warp's (1 or 2) tiles the next 32-column group uses
  float acc_s[8], acc_q[8];                        // fused statistics (n_tiles == 1): lane l <-> column c_begin + 32 g + l
#pragma unroll
  for (int u = 0; u < 8; ++u) { acc_s[u] = 0.f; acc_q[u] = 0.f; }
  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int m_tile = tile / p.n_tiles, n_tile = tile % p.n_tiles;
    const int acc = it & 1;
    const uint32_t acc_phase = (it >> 1) & 1;
    mbar_wait(&tfull[acc], acc_phase);
    fence_after_sync();
    const long long m = static_cast<long long>(m_tile) * BLOCK_M + row;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * p.block_n;
    int c0 = c_begin;
    if (stage_tile != 0) {
      for (; c0 + 32 <= c_end; c0 += 32) {
        float va[16], vb[16];
        tmem_ld16(taddr + c0, va);
        tmem_ld16(taddr + c0 + 16, vb);
        const int n0 = n_tile * p.block_n + c0;
        if (m < p.M) { epilogue_math16(p, m, n0, false, va); epilogue_math16(p, m, n0 + 16, false, vb); }
        const uint32_t tile = stage_tile + tsel * STAGE_TILE, srow = tile + static_cast<uint32_t>(lane) * 64u;
        if (lane == 0) { if (p.stage_out == 2) bulk_wait_read_1(); else bulk_wait_read_all(); }   // this tile's previous store has left shared memory
        __syncwarp();
        const bf16x8 k0 = pack8(va), k1 = pack8(va + 8), k2 = pack8(vb), k3 = pack8(vb + 8);
        st_shared_v4(srow + ((0u ^ ssw) << 4), k0);
        st_shared_v4(srow + ((1u ^ ssw) << 4), k1);
        st_shared_v4(srow + ((2u ^ ssw) << 4), k2);
        st_shared_v4(srow + ((3u ^ ssw) << 4), k3);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tma_store_2d(tmo, tile, n0, m_tile * BLOCK_M + q * 32); bulk_commit_group(); }
        if (p.stage_out == 2) tsel ^= 1u;
        if (p.stats != nullptr) {                   // warp-uniform: per-column sums of the STORED values of this warp's 32 rows
          float sv[32], sq[32];
          unpack8(k0, sv); unpack8(k1, sv + 8); unpack8(k2, sv + 16); unpack8(k3, sv + 24);
#pragma unroll
          for (int i = 0; i < 32; ++i) { sv[i] = (m < p.M) ? sv[i] : 0.f; sq[i] = sv[i] * sv[i]; }
          const float cs = warp_column_sums(sv, lane), cq = warp_column_sums(sq, lane);
          if (p.n_tiles == 1) {                     // every tile has the same columns: keep the sums in registers until the end
            const int g = (c0 - c_begin) >> 5;
#pragma unroll
            for (int u = 0; u < 8; ++u) if (u == g) { acc_s[u] += cs; acc_q[u] += cq; }
          } else {
            atomicAdd(p.stats + n0 + lane, static_cast<double>(cs));
            atomicAdd(p.stats + p.N + n0 + lane, static_cast<double>(cq));
          }
        }
      }
    }
    for (; c0 < c_end; c0 += 16) {
      float v[16];
      tmem_ld16(taddr + c0, v);
      if (m < p.M) {
        const int n0 = n_tile * p.block_n + c0;
        epilogue_math16(p, m, n0, f32, v);
        if (f32) {
          float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + m * p.out_ld + n0);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          bf16x8* op = reinterpret_cast<bf16x8*>(static_cast<__nv_bfloat16*>(p.out) + m * p.out_ld + n0);
          op[0] = pack8(v);
          op[1] = pack8(v + 8);
        }
      }
    }
    fence_before_sync();
    mbar_arrive(&tempty[acc]);
  }
  if (stage_tile != 0 && lane == 0) bulk_wait_all();
  if (p.stats != nullptr && p.n_tiles == 1 && stage_tile != 0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n0 = c_begin + 32 * u;
      if (n0 + 32 <= c_end) {
        atomicAdd(p.stats + n0 + lane, static_cast<double>(acc_s[u]));
        atomicAdd(p.stats + p.N + n0 + lane, static_cast<double>(acc_q[u]));
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) conv_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_o,
                                                               const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // swizzle patterns are anchored at 1024 B
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t b_stage_bytes = static_cast<uint32_t>(p.block_n) * BLOCK_K * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + S * A_STAGE_BYTES;
  uint8_t* sS = sB + S * b_stage_bytes;      // [8 epilogue warps][2 KB] output tiles (stage_out)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + static_cast<uint32_t>(p.stage_out) * 8 * STAGE_TILE);
  uint64_t* full = bars;               // [S]  128 gather arrivals + 1 expect_tx arrival
  uint64_t* empty = bars + S;          // [S]  tcgen05.commit
  uint64_t* tfull = bars + 2 * S;      // [2]  accumulator ready
  uint64_t* tempty = bars + 2 * S + 2; // [2]  accumulator drained (128 epilogue threads; 256 in TMA mode)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  uint32_t ncols = 32;
  while (ncols < 2u * p.block_n) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], p.use_tma ? 1 : NUM_A_THREADS + 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], p.use_tma ? 256 : 128); }
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, ncols);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp < GATHER_WARPS) {
    // ===================== A gather producers: 8 lanes per im2col row =====================
    const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;            // chunk of the K block, first row
    uint32_t stage = 0, phase = 0;
    // destination of row r0 + ROW_STEP i, chunk c:  row * 128 + ((c ^ (row & 7)) << 4); (r0 + ROW_STEP i) & 7 == r0 & 7
    const uint32_t dst0 = smem_u32(sA) + r0 * 128 + ((c ^ (r0 & 7)) << 4);
    if (p.use_tma) {
      // 1x1 / stride-1 layer: no gather (the loader warp issues one TMA box per K block); the first four of these warps own
      // TMEM lane quadrants 0-3 too, so they drain the upper half of the accumulator columns (the other four stay idle)
      if (warp < 4) {
        const int c_begin = ((p.block_n / 16 + 1) / 2) * 16;
        epilogue_role(p, tmem_base, tfull, tempty, warp & 3, lane, c_begin, p.block_n, total_tiles,
                      p.stage_out ? smem_u32(sS) + static_cast<uint32_t>(4 + (warp & 3)) * static_cast<uint32_t>(p.stage_out) * STAGE_TILE : 0u, &tma_o);
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles;
      // per-row pixel decode, once per tile: gather base (h, w) and the element offset of that pixel
      int hb[ROWS_PER_THREAD], wb[ROWS_PER_THREAD]; long long rbase[ROWS_PER_THREAD]; uint32_t okmask = 0;
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) {
        const long long m = static_cast<long long>(m_tile) * BLOCK_M + r0 + ROW_STEP * i;
        int wo = 0, ho = 0, bb = 0;
        if (m < p.M) { okmask |= 1u << i; wo = static_cast<int>(m % p.Wo); const long long t = m / p.Wo; ho = static_cast<int>(t % p.Ho); bb = static_cast<int>(t / p.Ho); }
        hb[i] = p.mode == 0 ? ho * p.sh - p.ph : ho + p.ph;
        wb[i] = p.mode == 0 ? wo * p.sw - p.pw : wo + p.pw;
        rbase[i] = ((static_cast<long long>(bb) * p.H + hb[i]) * p.W + wb[i]) * p.a_ld;      // may point outside: only used when in range
      }
      const bool unit = p.mode == 0 || (p.sh == 1 && p.sw == 1);     // (h, w) = base + signed tap offset, no divisibility test
      const int sgn = p.mode == 0 ? 1 : -1;
      for (int kb = 0; kb < p.KB; ++kb) {
        // this thread's 8 K elements of the block: one tap, 8 consecutive channels
        const int k = kb * BLOCK_K + c * 8;
        const int tp = k / p.C, ci = k - tp * p.C;
        const int khi = tp / p.kw, kwi = tp - khi * p.kw;
        const bool k_ok = k < p.K;
        const int dhh = sgn * khi * p.dh, dww = sgn * kwi * p.dw;
        const long long delta = (static_cast<long long>(dhh) * p.W + dww) * p.a_ld + ci;
        mbar_wait(&empty[stage], phase ^ 1);
        const uint32_t dst = dst0 + stage * A_STAGE_BYTES;
        if (unit) {
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            const int hi = hb[i] + dhh, wi = wb[i] + dww;
            const bool ok = k_ok && ((okmask >> i) & 1u) && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            cp_async16(dst + i * (ROW_STEP * 128), ok ? p.a + rbase[i] + delta : p.a, ok ? 16u : 0u);
          }
        } else {
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            const int hn = hb[i] + dhh, wn = wb[i] + dww;          // strided data gradient: only multiples of the stride hit an output pixel
            bool ok = k_ok && ((okmask >> i) & 1u) && hn >= 0 && wn >= 0 && (hn % p.sh) == 0 && (wn % p.sw) == 0;
            const int hi = hn / p.sh, wi = wn / p.sw;
            ok = ok && hi < p.H && wi < p.W;
            const long long off = (rbase[i] / p.a_ld - (static_cast<long long>(hb[i]) * p.W + wb[i]) + static_cast<long long>(hi) * p.W + wi) * p.a_ld + ci;
            cp_async16(dst + i * (ROW_STEP * 128), ok ? p.a + off : p.a, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(&full[stage]);
        if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_LOAD) {
    // ===================== weight tiles: one bulk copy per K block =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const __nv_bfloat16* wsrc = p.wpk + static_cast<long long>(n_tile) * p.KB * p.block_n * BLOCK_K;
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], b_stage_bytes + (p.use_tma ? A_STAGE_BYTES : 0));
          if (p.use_tma)       // the A tile is a plain [128 pixels][64 channels] box of the activation matrix
            tma_load_2d(smem_u32(sA) + stage * A_STAGE_BYTES, &tma_a, kb * BLOCK_K, (tile / p.n_tiles) * BLOCK_M, &full[stage]);
          bulk_g2s(smem_u32(sB) + stage * b_stage_bytes, wsrc + static_cast<long long>(kb) * p.block_n * BLOCK_K,
                   b_stage_bytes, &full[stage]);
          if (++stage == static_cast<uint32_t>(S)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (warp-uniform loop: descriptors in uniform registers, elected lane issues) =====================
    const bool leader = elect_one();
    const uint32_t idesc = instr_desc_bf16(BLOCK_M, p.block_n, 0, 0);
    const uint32_t d_hi = (1024u >> 4) | (1u << 14) | (2u << 29);       // SBO = 8 rows x 128 B, version 1, SWIZZLE_128B
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
        const uint32_t a0 = (((smem_u32(sA) + stage * A_STAGE_BYTES) >> 4) & 0x3FFF) | (1u << 16);
        const uint32_t b0 = (((smem_u32(sB) + stage * b_stage_bytes) >> 4) & 0x3FFF) | (1u << 16);
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / 16; ++kk) {
          const uint64_t ad = (static_cast<uint64_t>(d_hi) << 32) | (a0 + kk * 2);       // +32 B per 16 K elements
          const uint64_t bd = (static_cast<uint64_t>(d_hi) << 32) | (b0 + kk * 2);
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
    const int c_end = p.use_tma ? ((p.block_n / 16 + 1) / 2) * 16 : p.block_n;
    epilogue_role(p, tmem_base, tfull, tempty, warp & 3, lane, 0, c_end, total_tiles,
                  p.stage_out ? smem_u32(sS) + static_cast<uint32_t>(warp & 3) * static_cast<uint32_t>(p.stage_out) * STAGE_TILE : 0u, &tma_o);
  }

  fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) { fence_after_sync(); tmem_dealloc(tmem_base, ncols); }
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

extern "C" int air_conv_gemm_bf16_affine(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                         int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                         const void* wpk, int N, int K, void* out, long long out_ld,
                                         const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                                         void* out2, long long out2_ld, const float* post_scale, const float* post_shift,
                                         int num_sms, int flags, cudaStream_t stream);

extern "C" int air_conv_gemm_bf16_ex(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                     int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                     const void* wpk, int N, int K, void* out, long long out_ld,
                                     const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                                     void* out2, long long out2_ld, int num_sms, int flags, cudaStream_t stream) {
  return air_conv_gemm_bf16_affine(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K, out, out_ld,
                                   bias, bias_rows, res, res_ld, relu, out2, out2_ld, nullptr, nullptr, num_sms, flags, stream);
}

// as _ex, plus an optional per-channel affine applied after the ReLU: out = relu(acc + bias) * post_scale[n] + post_shift[n]
static int launch_conv_gemm(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                            int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                            const void* wpk, int N, int K, void* out, long long out_ld,
                            const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                            void* out2, long long out2_ld, const float* post_scale, const float* post_shift,
                            double* stats, int num_sms, int flags, cudaStream_t stream);

extern "C" int air_conv_gemm_bf16_affine(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                         int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                         const void* wpk, int N, int K, void* out, long long out_ld,
                                         const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                                         void* out2, long long out2_ld, const float* post_scale, const float* post_shift,
                                         int num_sms, int flags, cudaStream_t stream) {
  return launch_conv_gemm(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K, out, out_ld, bias,
                          bias_rows, res, res_ld, relu, out2, out2_ld, post_scale, post_shift, nullptr, num_sms, flags, stream);
}

// air_conv_gemm_bf16 whose epilogue also accumulates the per-channel sum / sum of squares of the stored (bf16) output into
// stats[0..N) / stats[N..2N) (fp64, caller zeroes): the batch statistics of the BatchNorm that follows (resnet.py:65-68).
// bf16 storage and N a multiple of 32 (per tile a multiple of 32 columns) only.
extern "C" int air_conv_gemm_bf16_stats(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                                        int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                                        const void* wpk, int N, int K, void* out, long long out_ld,
                                        const float* bias, const void* res, long long res_ld, int relu, double* stats,
                                        int num_sms, int flags, cudaStream_t stream) {
  if (!stats) return AIR_ERR_ARG;
  return launch_conv_gemm(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K, out, out_ld, bias,
                          0, res, res_ld, relu, nullptr, 0, nullptr, nullptr, stats, num_sms, flags, stream);
}

static int launch_conv_gemm(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                            int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                            const void* wpk, int N, int K, void* out, long long out_ld,
                            const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                            void* out2, long long out2_ld, const float* post_scale, const float* post_shift,
                            double* stats, int num_sms, int flags, cudaStream_t stream) {
  if (!a || !wpk || !out || B <= 0 || ((post_scale == nullptr) != (post_shift == nullptr))) return AIR_ERR_ARG;
  if (C % 8 != 0 || a_ld % 8 != 0 || out_ld % 8 != 0 || (res && res_ld % 8 != 0) || (out2 && out2_ld % 8 != 0)) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out2) & 15) || bias_rows < 0) return AIR_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(post_scale) | reinterpret_cast<uintptr_t>(post_shift)) & 15) return AIR_ERR_UNSUPPORTED;
  if (bias_rows > 0 && (N % 4) != 0) return AIR_ERR_UNSUPPORTED;
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
  p.out = out; p.out_ld = out_ld; p.bias = bias;
  p.res = res; p.res_ld = res_ld; p.relu = relu; p.flags = flags;
  p.bias_rows = bias_rows; p.out2 = out2; p.out2_ld = out2_ld;
  p.post_scale = post_scale; p.post_shift = post_shift; p.stats = stats;
  const int stage_bytes = A_STAGE_BYTES + bn * BLOCK_K * 2;
  int stages = (198 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return AIR_ERR_UNSUPPORTED;
  p.stages = stages;
  static const int stage_env = [] { const char* e = getenv("AIR_GEMM_STAGE_OUT"); return (e && e[0] == '0') ? 0 : 1; }();
  p.stage_out = (stage_env && !(flags & AIR_CONV_F32_OUT) && bn % 32 == 0 && p.M < 0x7fffffffLL) ? 1 : 0;
  if (p.stage_out) {
    // a second tile per epilogue warp (the store of one drains while the next is filled): for the price of one pipeline stage
    // when there are at least six, for free when it fits anyway
    auto total = [&](int st, int tiles) { return 1024 + static_cast<size_t>(st) * stage_bytes + tiles * 8 * STAGE_TILE + (2 * st + 4) * 8 + 16; };
    if (total(stages, 2) <= 227 * 1024) p.stage_out = 2;
    else if (stages >= 6 && total(stages - 1, 2) <= 227 * 1024) { p.stage_out = 2; p.stages = --stages; }
  }
  if (stats && !p.stage_out) return AIR_ERR_UNSUPPORTED;
  const size_t smem = 1024 + static_cast<size_t>(stages) * stage_bytes + static_cast<size_t>(p.stage_out) * 8 * STAGE_TILE + (2 * stages + 4) * 8 + 16;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  if (num_sms <= 0) num_sms = 148;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int grid = total_tiles < num_sms ? total_tiles : num_sms;
  CUtensorMap tm{};
  p.use_tma = 0;
  if (kh == 1 && kw == 1 && sh == 1 && sw == 1 && ph == 0 && pw == 0 && Ho == H && Wo == W && p.M < 0x7fffffffLL) {
    if (air_tmap::make_mat_tmap(&tm, a, a_ld, p.M, C, BLOCK_M) == 0) p.use_tma = 1;     // otherwise: the gather path
  }
  if (stats && p.use_tma && (((bn / 16 + 1) / 2) * 16) % 32 != 0) return AIR_ERR_UNSUPPORTED;     // a 16-column remainder would bypass the staged path
  CUtensorMap tmo = tm;
  if (p.stage_out && air_tmap::make_out_tmap(&tmo, out, out_ld, p.M, N) != 0) return AIR_ERR_DRIVER;
  conv_gemm_kernel<<<grid, THREADS, smem, stream>>>(tm, tmo, p);
  return air_launch_status();
}
