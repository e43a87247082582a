// Detection-error curve, EER and min t-DCF on device.
//
// Replaces eval_metrics.py:19-46 (compute_det_curve / compute_eer) and :141-160 (the t-DCF curve) of the reference,
// which run on numpy after a D2H copy of every score (call sites main_train.py:662-663, score_fusion.py:117-118).
//
// The reference sorts concat(target, nontarget) with a STABLE mergesort, so inside a group of equal scores all
// targets come before all nontargets.  Here the scores become order-preserving u64 images of their fp64 values
// (fp32 scores convert exactly), the class rides along as a one-byte payload, and a STABLE LSD radix sort (8-bit
// digits: histogram, scan, rank-preserving scatter) of the array in the same concat order reproduces that order
// exactly.  fp32 scores have 29 zero low mantissa bits as fp64, so their sort starts at bit 24 (5 passes, not 8).
// The curve itself is integer work (a running count of nontargets) followed by two IEEE fp64 divisions per point,
// so frr / far / EER / t-DCF are BIT-IDENTICAL to numpy's.  No fused multiply-add is allowed in the t-DCF line
// (numpy rounds the two products separately), hence the explicit __dmul_rn / __dadd_rn.
//
// Everything here is HBM- / launch-latency bound (n <= a few hundred thousand trials: 9 bytes per trial and pass).
#include "common.cuh"
#include <math.h>

namespace air_det {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS;        // 2048 keys per CTA
constexpr int RADIX = 256;

__device__ __forceinline__ unsigned long long ordered_bits(double s, int negate) {
  unsigned long long u = (unsigned long long)__double_as_longlong(s);
  const unsigned long long SIGN = 0x8000000000000000ull;
  if (negate) u ^= SIGN;
  if ((u & ~SIGN) == 0ull) u = 0ull;                                   // -0.0 and +0.0 tie in the reference's comparison
  if ((u & ~SIGN) > 0x7ff0000000000000ull) u = 0x7ff8000000000000ull;  // NaN sorts last, like numpy
  return (u & SIGN) ? ~u : (u | SIGN);
}
__device__ __forceinline__ double score_of_key(unsigned long long o) {
  const unsigned long long SIGN = 0x8000000000000000ull;
  return __longlong_as_double((long long)((o & SIGN) ? (o & ~SIGN) : ~o));
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
det_keys_kernel(const T* __restrict__ tar, long long n_tar, const T* __restrict__ non, long long n_non,
                int negate, unsigned long long* __restrict__ keys, uint8_t* __restrict__ cls) {
  const long long n = n_tar + n_non;
  for (long long i = blockIdx.x * (long long)THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * THREADS) {
    const bool is_non = i >= n_tar;
    keys[i] = ordered_bits((double)(is_non ? non[i - n_tar] : tar[i]), negate);
    cls[i] = is_non ? 1 : 0;
  }
}

// tile element e of warp w, round j, lane l  ->  tile_base + w*(32*ITEMS) + j*32 + l   (order-preserving per warp)
__device__ __forceinline__ long long tile_index(long long tile, int w, int j, int lane) {
  return tile * TILE + w * (32 * ITEMS) + j * 32 + lane;
}

__global__ void __launch_bounds__(THREADS)
radix_hist_kernel(const unsigned long long* __restrict__ keys, long long n, int shift, int ntiles,
                  uint32_t* __restrict__ counts) {
  __shared__ uint32_t hist[RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const long long i = tile_index(blockIdx.x, w, j, lane);
    if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & (RADIX - 1)], 1u);
  }
  __syncthreads();
  counts[(long long)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];     // digit-major for the scan
}

// In-place exclusive scan of a[0..len) by ONE CTA; a[len] receives the total.
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(uint32_t* __restrict__ a, long long len) {
  __shared__ uint32_t part[1024];
  const int t = threadIdx.x;
  const long long per = (len + 1023) / 1024;
  const long long lo = (long long)t * per, hi = lo + per < len ? lo + per : len;
  uint32_t s = 0;
  for (long long i = lo; i < hi; ++i) s += a[i];
  part[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {                 // Hillis-Steele inclusive scan of the 1024 partial sums
    const uint32_t v = t >= o ? part[t - o] : 0u;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  uint32_t run = part[t] - s;
  for (long long i = lo; i < hi; ++i) {
    const uint32_t v = a[i];
    a[i] = run;
    run += v;
  }
  if (t == 1023) a[len] = part[1023];
}

__global__ void __launch_bounds__(THREADS)
radix_scatter_kernel(const unsigned long long* __restrict__ in, const uint8_t* __restrict__ cls_in,
                     unsigned long long* __restrict__ out, uint8_t* __restrict__ cls_out, long long n,
                     int shift, int ntiles, const uint32_t* __restrict__ offsets) {
  __shared__ uint32_t cnt[WARPS][RADIX];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < WARPS; ++k) cnt[k][threadIdx.x] = 0;
  __syncthreads();
  unsigned long long key[ITEMS];
  uint32_t rank[ITEMS];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const long long i = tile_index(blockIdx.x, w, j, lane);
    const bool valid = i < n;
    key[j] = valid ? in[i] : 0ull;
    const uint32_t d = valid ? ((uint32_t)(key[j] >> shift) & (RADIX - 1)) : 0xffffu;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t prev = 0;
    if (valid && lane == leader) {
      prev = cnt[w][d];
      cnt[w][d] = prev + __popc(peers);
    }
    prev = __shfl_sync(0xffffffffu, prev, leader);
    rank[j] = prev + __popc(peers & lt);
    __syncwarp();
  }
  __syncthreads();
  {                                                    // thread d: where each warp's run of digit d starts
    uint32_t run = offsets[(long long)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
    for (int k = 0; k < WARPS; ++k) {
      const uint32_t c = cnt[k][threadIdx.x];
      cnt[k][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const long long i = tile_index(blockIdx.x, w, j, lane);
    if (i < n) {
      const uint32_t dst = cnt[w][(uint32_t)(key[j] >> shift) & (RADIX - 1)] + rank[j];
      out[dst] = key[j];
      cls_out[dst] = cls_in[i];
    }
  }
}

__global__ void __launch_bounds__(THREADS)
tile_nontarget_count_kernel(const uint8_t* __restrict__ cls, long long n, uint32_t* __restrict__ tile_non) {
  __shared__ uint32_t total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t c = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const long long i = tile_index(blockIdx.x, w, j, lane);
    c += __popc(__ballot_sync(0xffffffffu, i < n && cls[i] != 0));
  }
  if (lane == 0) atomicAdd(&total, c);                 // integer: order-free
  __syncthreads();
  if (threadIdx.x == 0) tile_non[blockIdx.x] = total;
}

// One operating point of the curve and the argmin bookkeeping (np.argmin: first minimum, first NaN wins).
struct Best {
  double v; long long idx; double frr, far;
};
__device__ __forceinline__ bool better(const Best& a, const Best& b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an != bn) return an;
  if (an || a.v == b.v) return a.idx < b.idx;
  return a.v < b.v;
}
__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best c;
    c.v = __shfl_xor_sync(0xffffffffu, b.v, o);
    c.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
    c.frr = __shfl_xor_sync(0xffffffffu, b.frr, o);
    c.far = __shfl_xor_sync(0xffffffffu, b.far, o);
    if (better(c, b)) b = c;
  }
  return b;
}
__device__ __forceinline__ Best block_best(Best b, Best* sm) {      // sm: >= 32 entries; result valid in warp 0
  b = warp_best(b);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = b;
  __syncthreads();
  Best r = sm[lane < nw ? lane : 0];
  return warp_best(r);
}

__device__ __forceinline__ double tdcf_norm(double c1, double c2, double frr, double far) {
  // eval_metrics.py:157-160: (C1 * Pmiss_cm + C2 * Pfa_cm) / np.minimum(C1, C2), every operation rounded on its own
  return __ddiv_rn(__dadd_rn(__dmul_rn(c1, frr), __dmul_rn(c2, far)), fmin(c1, c2));
}

// Curve point k (1-based, k = i + 1 for sorted position i): frr = #targets among the first k / n_tar,
// far = #nontargets NOT among the first k / n_non  (eval_metrics.py:31-35).
__global__ void __launch_bounds__(THREADS)
det_curve_kernel(const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ cls, long long n,
                 long long n_tar, long long n_non,
                 const uint32_t* __restrict__ tile_non_before, double c1, double c2, int want_tdcf,
                 double* __restrict__ frr_out, double* __restrict__ far_out, double* __restrict__ thr_out,
                 double* __restrict__ tdcf_out, Best* __restrict__ tile_eer, Best* __restrict__ tile_tdcf) {
  __shared__ uint32_t wtot[WARPS];
  __shared__ Best sm[32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t le = 0xffffffffu >> (31 - lane);
  unsigned long long key[ITEMS];
  uint32_t incl[ITEMS];
  uint32_t run = 0;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const long long i = tile_index(blockIdx.x, w, j, lane);
    key[j] = i < n ? keys[i] : 0ull;
    const uint32_t b = __ballot_sync(0xffffffffu, i < n && cls[i] != 0);
    incl[j] = run + __popc(b & le);
    run += __popc(b);
  }
  if (lane == 0) wtot[w] = run;
  __syncthreads();
  uint32_t before = tile_non_before[blockIdx.x];
  for (int k = 0; k < w; ++k) before += wtot[k];
  const double dt = (double)n_tar, dn = (double)n_non;
  Best be, bt;
  be.v = INFINITY; be.idx = 0x7fffffffffffffffll; be.frr = 0.0; be.far = 0.0;
  bt = be;
  if (blockIdx.x == 0 && threadIdx.x == 0) {           // point 0: frr 0, far 1, threshold = lowest score - 0.001
    be.v = 1.0; be.idx = 0; be.frr = 0.0; be.far = 1.0;
    bt = be;
    bt.v = tdcf_norm(c1, c2, 0.0, 1.0);
    if (frr_out) frr_out[0] = 0.0;
    if (far_out) far_out[0] = 1.0;
    if (thr_out) thr_out[0] = score_of_key(keys[0]) - 0.001;
    if (tdcf_out) tdcf_out[0] = bt.v;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const long long i = tile_index(blockIdx.x, w, j, lane);
    if (i < n) {
      const long long non_seen = (long long)before + incl[j];
      const double frr = __ddiv_rn((double)(i + 1 - non_seen), dt);
      const double far = __ddiv_rn((double)(n_non - non_seen), dn);
      Best c;
      c.v = fabs(frr - far); c.idx = i + 1; c.frr = frr; c.far = far;
      if (better(c, be)) be = c;
      if (frr_out) frr_out[i + 1] = frr;
      if (far_out) far_out[i + 1] = far;
      if (thr_out) thr_out[i + 1] = score_of_key(key[j]);
      if (want_tdcf) {
        c.v = tdcf_norm(c1, c2, frr, far);
        if (better(c, bt)) bt = c;
        if (tdcf_out) tdcf_out[i + 1] = c.v;
      }
    }
  }
  be = block_best(be, sm);
  if (threadIdx.x == 0) tile_eer[blockIdx.x] = be;
  if (want_tdcf) {
    bt = block_best(bt, sm);
    if (threadIdx.x == 0) tile_tdcf[blockIdx.x] = bt;
  }
}

// out[0..9] = eer, eer threshold, frr, far, min t-DCF, its threshold, n_tar, n_non, eer index, t-DCF index
__global__ void __launch_bounds__(1024)
det_final_kernel(const unsigned long long* __restrict__ keys, long long n_tar, long long n_non, int ntiles,
                 const Best* __restrict__ tile_eer, const Best* __restrict__ tile_tdcf, int want_tdcf,
                 double* __restrict__ out) {
  __shared__ Best sm[32];
  Best be, bt;
  be.v = INFINITY; be.idx = 0x7fffffffffffffffll; be.frr = 0.0; be.far = 0.0;
  bt = be;
  for (int i = threadIdx.x; i < ntiles; i += blockDim.x) {
    const Best c = tile_eer[i];
    if (better(c, be)) be = c;
    if (want_tdcf) {
      const Best d = tile_tdcf[i];
      if (better(d, bt)) bt = d;
    }
  }
  be = block_best(be, sm);
  if (want_tdcf) bt = block_best(bt, sm);
  if (threadIdx.x == 0) {
    const double s0 = score_of_key(keys[0]) - 0.001;
    out[0] = (be.frr + be.far) / 2.0;                  // np.mean((frr[i], far[i])), eval_metrics.py:44
    out[1] = be.idx == 0 ? s0 : score_of_key(keys[be.idx - 1]);
    out[2] = be.frr;
    out[3] = be.far;
    out[4] = want_tdcf ? bt.v : nan("");
    out[5] = want_tdcf ? (bt.idx == 0 ? s0 : score_of_key(keys[bt.idx - 1])) : nan("");
    out[6] = (double)n_tar;
    out[7] = (double)n_non;
    out[8] = (double)be.idx;
    out[9] = want_tdcf ? (double)bt.idx : -1.0;
  }
}

// counts of scores >= threshold and < threshold (eval_metrics.py:7-8,14), compared as fp64 like numpy
template <typename T>
__global__ void __launch_bounds__(THREADS)
threshold_count_kernel(const T* __restrict__ s, long long n, double thr, unsigned long long* __restrict__ out) {
  __shared__ unsigned long long ge_lt[2];
  if (threadIdx.x < 2) ge_lt[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long ge = 0, lt = 0;
  for (long long i = blockIdx.x * (long long)THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * THREADS) {
    const double v = (double)s[i];
    ge += v >= thr;
    lt += v < thr;
  }
  for (int o = 16; o > 0; o >>= 1) {
    ge += __shfl_xor_sync(0xffffffffu, ge, o);
    lt += __shfl_xor_sync(0xffffffffu, lt, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&ge_lt[0], ge); atomicAdd(&ge_lt[1], lt); }
  __syncthreads();
  if (threadIdx.x < 2) atomicAdd(&out[threadIdx.x], ge_lt[threadIdx.x]);
}

struct Workspace {
  long long ntiles, keys_a, keys_b, cls_a, cls_b, counts, tile_non, tile_eer, tile_tdcf, total;
};
static Workspace layout(long long n) {
  Workspace w;
  auto up = [](long long v) { return (v + 255) / 256 * 256; };
  w.ntiles = (n + TILE - 1) / TILE;
  long long o = 0;
  w.keys_a = o; o += up(n * 8);
  w.keys_b = o; o += up(n * 8);
  w.cls_a = o; o += up(n);
  w.cls_b = o; o += up(n);
  w.counts = o; o += up((RADIX * w.ntiles + 1) * 4);
  w.tile_non = o; o += up((w.ntiles + 1) * 4);
  w.tile_eer = o; o += up(w.ntiles * (long long)sizeof(Best));
  w.tile_tdcf = o; o += up(w.ntiles * (long long)sizeof(Best));
  w.total = o;
  return w;
}

// fp32 scores are exact fp64 values with 29 zero low mantissa bits: their keys agree below bit 24 whenever they
// agree above it, so the three lowest digit passes cannot reorder anything and are skipped.
static int first_pass(bool is_f64) { return is_f64 ? 0 : 3; }

template <typename T>
static int det_curve(const T* target, long long n_tar, const T* nontarget, long long n_non, int negate, double c1,
                     double c2, int want_tdcf, void* workspace, long long workspace_bytes, double* frr, double* far,
                     double* thresholds, double* tdcf, double* out, cudaStream_t stream) {
  const long long n = n_tar + n_non;
  if (n_tar < 0 || n_non < 0 || n < 1 || n > 0x7fffffffll || !workspace || !out) return AIR_ERR_ARG;
  if ((n_tar && !target) || (n_non && !nontarget)) return AIR_ERR_ARG;
  const Workspace w = layout(n);
  if (workspace_bytes < w.total) return AIR_ERR_ARG;
  char* base = (char*)workspace;
  unsigned long long* ka = (unsigned long long*)(base + w.keys_a);
  unsigned long long* kb = (unsigned long long*)(base + w.keys_b);
  uint8_t* ca = (uint8_t*)(base + w.cls_a);
  uint8_t* cb = (uint8_t*)(base + w.cls_b);
  uint32_t* counts = (uint32_t*)(base + w.counts);
  uint32_t* tile_non = (uint32_t*)(base + w.tile_non);
  Best* tile_eer = (Best*)(base + w.tile_eer);
  Best* tile_tdcf = (Best*)(base + w.tile_tdcf);
  const int ntiles = (int)w.ntiles;
  int grid = (int)((n + THREADS - 1) / THREADS);
  if (grid > 148 * 8) grid = 148 * 8;
  det_keys_kernel<T><<<grid, THREADS, 0, stream>>>(target, n_tar, nontarget, n_non, negate, ka, ca);
  for (int p = first_pass(sizeof(T) == 8); p < 8; ++p) {
    radix_hist_kernel<<<ntiles, THREADS, 0, stream>>>(ka, n, 8 * p, ntiles, counts);
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(counts, (long long)RADIX * ntiles);
    radix_scatter_kernel<<<ntiles, THREADS, 0, stream>>>(ka, ca, kb, cb, n, 8 * p, ntiles, counts);
    unsigned long long* t = ka; ka = kb; kb = t;
    uint8_t* u = ca; ca = cb; cb = u;
  }
  tile_nontarget_count_kernel<<<ntiles, THREADS, 0, stream>>>(ca, n, tile_non);
  exclusive_scan_kernel<<<1, 1024, 0, stream>>>(tile_non, ntiles);
  det_curve_kernel<<<ntiles, THREADS, 0, stream>>>(ka, ca, n, n_tar, n_non, tile_non, c1, c2, want_tdcf, frr, far,
                                                   thresholds, tdcf, tile_eer, tile_tdcf);
  det_final_kernel<<<1, 1024, 0, stream>>>(ka, n_tar, n_non, ntiles, tile_eer, tile_tdcf, want_tdcf, out);
  return air_launch_status();
}

template <typename T>
static int threshold_counts(const T* scores, long long n, double threshold, unsigned long long* counts_ge_lt,
                            cudaStream_t stream) {
  if (n < 0 || !counts_ge_lt || (n && !scores)) return AIR_ERR_ARG;
  cudaError_t e = cudaMemsetAsync(counts_ge_lt, 0, 16, stream);
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return AIR_OK;
  int grid = (int)((n + THREADS - 1) / THREADS);
  if (grid > 148 * 4) grid = 148 * 4;
  threshold_count_kernel<T><<<grid, THREADS, 0, stream>>>(scores, n, threshold, counts_ge_lt);
  return air_launch_status();
}

}  // namespace air_det

using namespace air_det;

extern "C" int air_det_workspace_bytes(long long n, long long* bytes) {
  if (n < 1 || !bytes) return AIR_ERR_ARG;
  *bytes = layout(n).total;
  return AIR_OK;
}

extern "C" int air_det_launches(long long n, int is_f64) {
  return n < 1 ? 0 : 1 + 3 * (8 - first_pass(is_f64 != 0)) + 4;
}

extern "C" int air_det_curve_f32(const float* target, long long n_tar, const float* nontarget, long long n_non,
                                 int negate, double c1, double c2, int want_tdcf, void* workspace,
                                 long long workspace_bytes, double* frr, double* far, double* thresholds,
                                 double* tdcf, double* out, cudaStream_t stream) {
  return det_curve<float>(target, n_tar, nontarget, n_non, negate, c1, c2, want_tdcf, workspace, workspace_bytes, frr,
                          far, thresholds, tdcf, out, stream);
}

extern "C" int air_det_curve_f64(const double* target, long long n_tar, const double* nontarget, long long n_non,
                                 int negate, double c1, double c2, int want_tdcf, void* workspace,
                                 long long workspace_bytes, double* frr, double* far, double* thresholds,
                                 double* tdcf, double* out, cudaStream_t stream) {
  return det_curve<double>(target, n_tar, nontarget, n_non, negate, c1, c2, want_tdcf, workspace, workspace_bytes, frr,
                           far, thresholds, tdcf, out, stream);
}

extern "C" int air_det_threshold_counts_f32(const float* scores, long long n, double threshold,
                                            unsigned long long* counts_ge_lt, cudaStream_t stream) {
  return threshold_counts<float>(scores, n, threshold, counts_ge_lt, stream);
}

extern "C" int air_det_threshold_counts_f64(const double* scores, long long n, double threshold,
                                            unsigned long long* counts_ge_lt, cudaStream_t stream) {
  return threshold_counts<double>(scores, n, threshold, counts_ge_lt, stream);
}
