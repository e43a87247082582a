// Fused LFCC front-end for sm_100a.
//
// One kernel replaces feature_extraction.py:93-138 of the reference (pre-emphasis, centred
// zero-padded framing, periodic Hamming-320, 512-pt rFFT, |.|^2, 20 linear triangular filters,
// log10, ortho DCT-II, delta, delta-delta) AND the crop / pad policy of dataset.py:66-79,513-528
// AND the layout change of main_train.py:338,347-348: each wave sample is read from HBM once
// and the features are written once, already in the layout/dtype the first conv consumes.
//
// Work split: a CTA owns `fseg` consecutive LFCC frames of one utterance (+2 halo frames each
// side for the delta-deltas).  A half-warp (16 lanes x 16 registers) computes one frame:
//   z[n] = y[2n] + i*y[2n+1]  (n < 160; the frame's 320 windowed samples, zero padded to 512)
//   256-point complex FFT as 16x16 Cooley-Tukey (radix-16 in registers, one smem transpose),
//   real-FFT split X[k] = Ze[k] + W512^k Zo[k] through a lane mirror shuffle, power,
//   sparse filterbank (each filter = one lane, ~24 FMAs), log10, 20x20 DCT.
// The window sits at offset 96 inside the 512 buffer in the reference (torch.stft centring); that is
// a pure phase factor and drops out of |X[k]|^2, so the samples are placed at offset 0 here.
#include "common.cuh"

namespace air_lfcc {

constexpr int NF = 20;                 // filters == cepstral coefficients
constexpr int FL = 320, FS = 160;      // window / hop (n_fft = 512)
constexpr int THREADS = 128;
constexpr int HW_PER_CTA = THREADS / 16;
constexpr int MAX_FRAMES = 132;        // computed frames per CTA (fseg + 4 halo) upper bound

// packed constant table (floats), built on the host from the module's registered buffers
constexpr int OFF_WIN = 0;                      // 320: Hamming window
constexpr int OFF_TW1 = OFF_WIN + FL;           // 256 float2: W256^(l*kj) at [kj*16+l]
constexpr int OFF_TW2 = OFF_TW1 + 512;          // 256 float2: (cos, sin)(2*pi*k/512)
constexpr int OFF_FBW = OFF_TW2 + 512;          // NF*33: filter weights, row stride 33
constexpr int OFF_FBS = OFF_FBW + NF * 33;      // NF int: first bin of each filter
constexpr int OFF_FBC = OFF_FBS + NF;           // NF int: number of bins (<= 32)
constexpr int OFF_DCT = OFF_FBC + NF;           // NF*21: DCT matrix W[k][f], row stride 21
constexpr int TBL_FLOATS = OFF_DCT + NF * 21;

struct Params {
  const float* wave; long long ldw; const int* lengths; int L; int B;
  const float* tbl;
  void* out; long long sb, sj, sd; int out_bf16; int time_minor;
  int Tout; int feat_len; int pad_mode; const int* start; int fseg;
  const float* silence;   // 60 floats (pad_mode 3 only)
  float preemph;          // 0.97 (with_emphasis) or 0
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// forward 4-point DFT; Z2/Z3: inputs x2 / x3 are known zeros (pruned first stage)
template <bool Z2, bool Z3>
__device__ __forceinline__ void dft4(float2 x0, float2 x1, float2 x2, float2 x3,
                                     float2& X0, float2& X1, float2& X2, float2& X3) {
  float2 t0, t1, t2, t3;
  if (Z2) { t0 = x0; t1 = x0; } else { t0 = cadd(x0, x2); t1 = csub(x0, x2); }
  if (Z3) { t2 = x1; t3 = x1; } else { t2 = cadd(x1, x3); t3 = csub(x1, x3); }
  X0 = cadd(t0, t2);
  X2 = csub(t0, t2);
  float2 m = mul_mi(t3);
  X1 = cadd(t1, m);
  X3 = csub(t1, m);
}

// In-place forward 16-point DFT, natural order in and out.  NZ = number of leading non-zero inputs.
template <int NZ>
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  float2 T[4][4];
  const float2 zero = make_float2(0.f, 0.f);
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) {
    const bool z2 = (n1 + 8) >= NZ, z3 = (n1 + 12) >= NZ;
    float2 x0 = v[n1], x1 = v[n1 + 4];
    float2 x2 = z2 ? zero : v[n1 + 8], x3 = z3 ? zero : v[n1 + 12];
    if (z2 && z3) dft4<true, true>(x0, x1, x2, x3, T[n1][0], T[n1][1], T[n1][2], T[n1][3]);
    else if (z3)  dft4<false, true>(x0, x1, x2, x3, T[n1][0], T[n1][1], T[n1][2], T[n1][3]);
    else          dft4<false, false>(x0, x1, x2, x3, T[n1][0], T[n1][1], T[n1][2], T[n1][3]);
  }
  // twiddles W16^(n1*k2) = exp(-2*pi*i*n1*k2/16)
  T[1][1] = cmul(T[1][1], make_float2(C1, -S1));
  T[1][2] = cmul(T[1][2], make_float2(R2, -R2));
  T[1][3] = cmul(T[1][3], make_float2(S1, -C1));
  T[2][1] = cmul(T[2][1], make_float2(R2, -R2));
  T[2][2] = mul_mi(T[2][2]);
  T[2][3] = cmul(T[2][3], make_float2(-R2, -R2));
  T[3][1] = cmul(T[3][1], make_float2(S1, -C1));
  T[3][2] = cmul(T[3][2], make_float2(-R2, -R2));
  T[3][3] = cmul(T[3][3], make_float2(-C1, S1));
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2)
    dft4<false, false>(T[0][k2], T[1][k2], T[2][k2], T[3][k2], v[k2], v[4 + k2], v[8 + k2], v[12 + k2]);
}

__global__ void __launch_bounds__(THREADS) lfcc_kernel(const Params p) {
  __shared__ __align__(16) float s_tbl[TBL_FLOATS];
  __shared__ __align__(16) float2 s_x[HW_PER_CTA][16 * 17];
  __shared__ float s_fbe[HW_PER_CTA][NF + 4];
  __shared__ float s_c[MAX_FRAMES][NF];

  const int b = blockIdx.y;
  int len = p.lengths ? p.lengths[b] : p.L;
  len = max(0, min(len, p.L));
  const int T = 1 + len / FS;

  // which source frames this utterance must produce (crop when T > feat_len)
  int first = 0, tend = T, shift = 0, rep = 0x40000000;
  if (p.feat_len > 0) {
    if (T > p.feat_len) { first = p.start ? p.start[b] : 0; first = max(0, min(first, T - p.feat_len));
                          tend = first + p.feat_len; shift = -first; }
    else if (p.pad_mode == 2) rep = T;                    // repeat: j = t + m*T
    else if (p.pad_mode == 3) shift = p.feat_len - T;     // silence is PREPENDED (dataset.py:528)
  }
  const int t0 = first + blockIdx.x * p.fseg;
  if (t0 >= tend) return;
  const int t1 = min(t0 + p.fseg, tend);
  const int fbase = max(t0 - 2, 0), fend = min(t1 + 2, T);
  const int nfr = fend - fbase;

  for (int i = threadIdx.x; i < TBL_FLOATS; i += THREADS) s_tbl[i] = __ldg(p.tbl + i);
  __syncthreads();
  const float2* s_win = reinterpret_cast<const float2*>(s_tbl + OFF_WIN);
  const float2* s_tw1 = reinterpret_cast<const float2*>(s_tbl + OFF_TW1);
  const float2* s_tw2 = reinterpret_cast<const float2*>(s_tbl + OFF_TW2);
  const float* s_fbw = s_tbl + OFF_FBW;
  const int* s_fbs = reinterpret_cast<const int*>(s_tbl + OFF_FBS);
  const int* s_fbc = reinterpret_cast<const int*>(s_tbl + OFF_FBC);
  const float* s_dct = s_tbl + OFF_DCT;

  const int hw = threadIdx.x >> 4, l = threadIdx.x & 15;
  const unsigned FULL = 0xffffffffu;
  const float* __restrict__ w = p.wave + (long long)b * p.ldw;
  const bool vec_ok = ((p.ldw & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.wave) & 7) == 0);
  float2* S = s_x[hw];
  float* Pbuf = reinterpret_cast<float*>(S);

  const int iters = (nfr + HW_PER_CTA - 1) / HW_PER_CTA;
  for (int it = 0; it < iters; ++it) {
    const int fl = it * HW_PER_CTA + hw;
    const bool valid = fl < nfr;
    const int t = fbase + min(fl, nfr - 1);
    const int s0 = FS * (t - 1);

    // ---- load 320 samples (10 float2 per lane), pre-emphasis, window ----
    float2 raw[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int idx = s0 + 2 * (l + 16 * j);
      float a = 0.f, c = 0.f;
      if (idx >= 0 && idx + 1 < len && vec_ok) {
        const float2 tv = __ldg(reinterpret_cast<const float2*>(w + idx));
        a = tv.x; c = tv.y;
      } else {
        if (idx >= 0 && idx < len) a = __ldg(w + idx);
        if (idx + 1 >= 0 && idx + 1 < len) c = __ldg(w + idx + 1);
      }
      raw[j] = make_float2(a, c);
    }
    float xfirst = 0.f;
    if (l == 0 && s0 - 1 >= 0 && s0 - 1 < len) xfirst = __ldg(w + s0 - 1);
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const int idx = s0 + 2 * (l + 16 * j);
      const float up = __shfl_up_sync(FULL, raw[j].y, 1, 16);
      const float wrap = __shfl_sync(FULL, raw[j > 0 ? j - 1 : 0].y, 15, 16);
      const float xm1 = (l > 0) ? up : (j > 0 ? wrap : xfirst);
      // y[n] = x[n] - 0.97 x[n-1]; zero outside [0, len)   (feature_extraction.py:105-106)
      float y0 = (idx >= 0 && idx < len) ? fmaf(-p.preemph, xm1, raw[j].x) : 0.f;
      float y1 = (idx + 1 >= 0 && idx + 1 < len) ? fmaf(-p.preemph, raw[j].x, raw[j].y) : 0.f;
      const float2 wv = s_win[l + 16 * j];
      v[j] = make_float2(y0 * wv.x, y1 * wv.y);
    }
#pragma unroll
    for (int j = 10; j < 16; ++j) v[j] = make_float2(0.f, 0.f);

    // ---- 256-point complex FFT = 16 x 16 ----
    fft16<10>(v);                                   // over j (n = l + 16 j) -> kj
#pragma unroll
    for (int kj = 1; kj < 16; ++kj) v[kj] = cmul(v[kj], s_tw1[kj * 16 + l]);
#pragma unroll
    for (int kj = 0; kj < 16; ++kj) S[l * 17 + kj] = v[kj];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = S[i * 17 + l];
    __syncwarp();
    fft16<16>(v);                                   // over l -> v[r] = Z[q + 16 r], q = lane
    // ---- real-FFT split + power:  k = q + 16 r pairs with 256 - k (lane (16-q)&15, reg 15-r) ----
    const int pl = (16 - l) & 15;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float mx = __shfl_sync(FULL, v[15 - r].x, pl, 16);
      const float my = __shfl_sync(FULL, v[15 - r].y, pl, 16);
      const float2 own = v[(16 - r) & 15];
      const float2 m = (l == 0) ? own : make_float2(mx, my);
      const float2 A = v[r];
      const float sx = A.x + m.x, sy = A.y - m.y;     // A + conj(M)
      const float dx = A.x - m.x, dy = A.y + m.y;     // A - conj(M)
      const float2 tw = s_tw2[l + 16 * r];            // (cos, sin)(2 pi k / 512)
      const float re = sx - tw.y * dx + tw.x * dy;
      const float im = sy - tw.y * dy - tw.x * dx;
      Pbuf[l + 16 * r] = 0.25f * (re * re + im * im); // |X[k]|^2   (feature_extraction.py:113)
    }
    __syncwarp();
    // ---- sparse triangular filterbank + log10 (feature_extraction.py:116-117) ----
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int f = l + 16 * pass;
      if (f < NF) {
        const int st = s_fbs[f], cn = s_fbc[f];
        float acc = 0.f;
        for (int i = 0; i < cn; ++i) acc = fmaf(Pbuf[st + i], s_fbw[f * 33 + i], acc);
        s_fbe[hw][f] = log10f(acc + 1.1920928955078125e-07f);
      }
    }
    __syncwarp();
    // ---- DCT-II ortho: c[k] = sum_f fbe[f] W[k][f]  (feature_extraction.py:120) ----
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int k = l + 16 * pass;
      if (k < NF) {
        float acc = 0.f;
#pragma unroll
        for (int f = 0; f < NF; ++f) acc = fmaf(s_fbe[hw][f], s_dct[k * 21 + f], acc);
        if (valid) s_c[fl][k] = acc;
      }
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- deltas (replicate edges, feature_extraction.py:41-58), pad/crop scatter, layout, dtype ----
  const int nout = t1 - t0;
  const int total = nout * 3 * NF;
  for (int i = threadIdx.x; i < total; i += THREADS) {
    int fo, d;
    if (p.time_minor) { d = i / nout; fo = i - d * nout; } else { fo = i / (3 * NF); d = i - fo * 3 * NF; }
    const int t = t0 + fo;
    const int part = d / NF, k = d - part * NF;
    const int tp = min(t + 1, T - 1), tm = max(t - 1, 0);
    float val;
    if (part == 0) val = s_c[t - fbase][k];
    else if (part == 1) val = s_c[tp - fbase][k] - s_c[tm - fbase][k];
    else {
      const float dp = s_c[min(tp + 1, T - 1) - fbase][k] - s_c[max(tp - 1, 0) - fbase][k];
      const float dm = s_c[min(tm + 1, T - 1) - fbase][k] - s_c[max(tm - 1, 0) - fbase][k];
      val = dp - dm;
    }
    for (int j = t + shift; j < p.Tout; j += rep) {
      const long long off = (long long)b * p.sb + (long long)j * p.sj + (long long)d * p.sd;
      if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[off] = f2bf(val);
      else reinterpret_cast<float*>(p.out)[off] = val;
    }
  }
}

// rows that no source frame maps to: zero tail (pad_mode 1) or silence head (pad_mode 3)
__global__ void lfcc_fill_kernel(const Params p) {
  const int b = blockIdx.y;
  int len = p.lengths ? p.lengths[b] : p.L;
  len = max(0, min(len, p.L));
  const int T = 1 + len / FS;
  if (T >= p.feat_len) return;
  const int npad = p.feat_len - T;
  const int total = npad * 3 * NF;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int jr, d;
    if (p.time_minor) { d = i / npad; jr = i - d * npad; } else { jr = i / (3 * NF); d = i - jr * 3 * NF; }
    const int j = (p.pad_mode == 3) ? jr : T + jr;
    const float val = (p.pad_mode == 3) ? p.silence[d] : 0.f;
    const long long off = (long long)b * p.sb + (long long)j * p.sj + (long long)d * p.sd;
    if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[off] = f2bf(val);
    else reinterpret_cast<float*>(p.out)[off] = val;
  }
}

}  // namespace air_lfcc

// rows that no source frame maps to (zero tail / silence head); shared with the tensor-core path (lfcc_tc.cu)
extern "C" int air_lfcc_fill(const int* lengths, int B, int L, void* out, long long sb, long long sj, long long sd,
                             int out_bf16, int Tout, int feat_len, int pad_mode, const float* silence, cudaStream_t stream) {
  using namespace air_lfcc;
  if (!out || B <= 0 || feat_len <= 0 || (pad_mode != 1 && pad_mode != 3)) return AIR_ERR_ARG;
  Params p{};
  p.lengths = lengths; p.L = L; p.B = B; p.out = out; p.sb = sb; p.sj = sj; p.sd = sd; p.out_bf16 = out_bf16;
  p.time_minor = (sj == 1); p.Tout = Tout; p.feat_len = feat_len; p.pad_mode = pad_mode; p.silence = silence;
  dim3 g2(8, B);
  lfcc_fill_kernel<<<g2, 256, 0, stream>>>(p);
  return air_launch_status();
}

extern "C" int air_lfcc_table_floats() { return air_lfcc::TBL_FLOATS; }

extern "C" int air_lfcc_fwd(const float* wave, long long ldw, const int* lengths, int B, int L,
                            const float* table, void* out, long long sb, long long sj, long long sd,
                            int out_bf16, int Tout, int feat_len, int pad_mode, const int* start,
                            const float* silence, float preemph, int fseg, cudaStream_t stream) {
  using namespace air_lfcc;
  if (!wave || !table || !out || B <= 0 || L < 0 || Tout <= 0) return AIR_ERR_ARG;
  if (pad_mode < 0 || pad_mode > 3) return AIR_ERR_ARG;
  if (pad_mode == 3 && feat_len > 0 && !silence) return AIR_ERR_ARG;
  const int Tmax = 1 + L / FS;
  if (feat_len > 0 && Tout != feat_len) return AIR_ERR_ARG;
  if (feat_len <= 0 && Tout < Tmax) return AIR_ERR_ARG;
  if (fseg <= 0) fseg = 60;
  if (fseg + 4 > MAX_FRAMES) return AIR_ERR_ARG;
  Params p;
  p.wave = wave; p.ldw = ldw; p.lengths = lengths; p.L = L; p.B = B; p.tbl = table;
  p.out = out; p.sb = sb; p.sj = sj; p.sd = sd; p.out_bf16 = out_bf16; p.time_minor = (sj == 1);
  p.Tout = Tout; p.feat_len = feat_len > 0 ? feat_len : 0; p.pad_mode = feat_len > 0 ? pad_mode : 0;
  p.start = start; p.fseg = fseg; p.silence = silence; p.preemph = preemph;
  const int need = (feat_len > 0 && Tmax > feat_len) ? feat_len : Tmax;
  dim3 grid((need + fseg - 1) / fseg, B);
  lfcc_kernel<<<grid, THREADS, 0, stream>>>(p);
  if (p.feat_len > 0 && (pad_mode == 1 || pad_mode == 3)) {
    dim3 g2(8, B);
    lfcc_fill_kernel<<<g2, 256, 0, stream>>>(p);
  }
  return air_launch_status();
}
