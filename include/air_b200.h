/* air_b200 -- C ABI of the B200-native hot path of yzyouzhang/ASVspoof2021_AIR.
 *
 * The reference has no FFI: its boundary is Python.  Each entry point below
 * replaces the arithmetic behind one reference call site (cited per function, paths relative
 * to the reference root) and is what the Python drop-in classes in asvspoof2021_air_b200/ bind
 * through ctypes.  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller
 *     (inputs, outputs, saved tensors, workspaces); the library never allocates or frees
 *     user-visible memory and keeps no per-call state;
 *   - every call is asynchronous on `stream` (no device synchronisation inside) and re-entrant
 *     across streams;
 *   - return value: 0 on success, < 0 for argument errors (AIR_ERR_*), > 0 = cudaError_t.
 * The air_audio_* group at the end is host code (file decoding into caller-owned HOST memory, no stream).
 */
#ifndef AIR_B200_H
#define AIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* air_stream_t; /* == cudaStream_t */

#define AIR_OK 0
#define AIR_ERR_ARG (-1)
#define AIR_ERR_UNSUPPORTED (-2)
/* host-side audio ingest only (air_audio_*): */
#define AIR_ERR_IO (-3)
#define AIR_ERR_FORMAT (-4)
#define AIR_ERR_CHECKSUM (-5)
#define AIR_ERR_NOMEM (-6)
/* the CUDA driver entry point needed to encode a TMA tensor map could not be resolved (no driver loaded) */
#define AIR_ERR_DRIVER (-7)

/* Library version (major*10000 + minor*100 + patch). */
int air_version(void);
/* Static description of a status code returned by any entry point (AIR_ERR_* name and meaning, or the CUDA runtime's
 * text for a positive cudaError_t).  The reference reports errors as Python exceptions; the
 * binding raises AirError with this text. */
const char* air_status_string(int status);
/* The most recent non-zero status returned to the CALLING THREAD by a device entry point, with the source line of the
 * check that rejected the call ("status -2 at conv_patch.cu:447: AIR_ERR_UNSUPPORTED ...").  The pointer stays valid
 * until the next call of this function on the same thread. */
const char* air_last_error_string(void);

/* ---------------------------------------------------------------------------------------------
 * LFCC front-end.  Replaces LFCC.forward (feature_extraction.py:93-138: pre-emphasis, torch.stft
 * framing, power, linear filterbank, log10, LinearDCT, delta x2), the crop/pad policy of
 * dataset.py:66-79,513-528 and the layout change of main_train.py:338,347-348 in one kernel.
 *
 *   wave     (B, L) float32, row stride `ldw` elements
 *   lengths  optional int32[B] valid samples per row (NULL: all L); T_b = 1 + len_b/160
 *   table    air_lfcc_table_floats() float32 constants packed by the host
 *            (asvspoof2021_air_b200/lfcc_tables.py, from the module's lfcc_fb / l_dct.weight)
 *   out      element (b, j, d) at out[b*sb + j*sj + d*sd], float32 or bf16 (out_bf16 != 0);
 *            j < Tout output frames, d < 60 = [c | delta | delta-delta]
 *   feat_len 0: no crop/pad, Tout >= 1 + L/160, frame j = LFCC frame j.
 *            >0: Tout == feat_len; utterances with T_b > feat_len are cropped at start[b]
 *            (NULL: 0), shorter ones are padded per pad_mode: 1 zero-append, 2 repeat
 *            (j <- j mod T_b), 3 silence-PREPEND with `silence` (60 float32).
 *   preemph  pre-emphasis coefficient: 0.97f (with_emphasis=True) or 0
 *   fseg     LFCC frames per CTA (0: default)
 * --------------------------------------------------------------------------------------------- */
int air_lfcc_table_floats(void);
int air_lfcc_fwd(const float* wave, long long ldw, const int* lengths, int B, int L,
                 const float* table, void* out, long long sb, long long sj, long long sd,
                 int out_bf16, int Tout, int feat_len, int pad_mode, const int* start,
                 const float* silence, float preemph, int fseg, air_stream_t stream);

/* Tensor-core path of the same front-end (csrc/lfcc_tc.cu): the 512-point spectrum of feature_extraction.py:109-113 as
 * a folded real DFT on tcgen05 (bf16 hi/lo 3-term split, fp32 accumulation in TMEM; deviation <= 1e-5*(|ref|+1)).
 * Same arguments as air_lfcc_fwd except the tables: `table` = air_lfcc_tc_table_floats() floats (window, per-bin
 * filter weight pairs, DCT), `wmat` = air_lfcc_tc_wmat_elems() bf16 (pre-swizzled DFT operand chunks); both are built
 * by lfcc_tables.pack_tc_table / pack_tc_dft from the module's registered buffers. */
int air_lfcc_tc_table_floats(void);
int air_lfcc_tc_wmat_elems(void);
int air_lfcc_tc_fwd(const float* wave, long long ldw, const int* lengths, int B, int L,
                    const float* table, const void* wmat, void* out, long long sb, long long sj, long long sd,
                    int out_bf16, int Tout, int feat_len, int pad_mode, const int* start,
                    const float* silence, float preemph, int num_sms, air_stream_t stream);
/* rows of a padded output that no source frame maps to (zero tail / silence head; dataset.py:513-528) */
int air_lfcc_fill(const int* lengths, int B, int L, void* out, long long sb, long long sj, long long sd,
                  int out_bf16, int Tout, int feat_len, int pad_mode, const float* silence, air_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM).
 * Replaces nn.Conv2d / nn.Conv1d forward and data-gradient (cuDNN in the reference:
 * resnet.py:56-60,131,140; ecapa_tdnn.py:19-23,39,50,56,111-118,139-145).  Activations are
 * channels-last bf16; weights are fp32 in GEMM layout [Cout][kh][kw][Cin] and are packed to bf16
 * UMMA tiles by air_conv_pack_weights (mode 0: fprop operand, N=Cout, K=taps*Cin;
 * mode 1: dgrad operand, N=Cin, K=taps*Cout).
 *
 * air_conv_gemm_bf16: out[m, n] = sum_k gather(a)[m, k] * W[n, k] (+bias[n]) (+res[m, n]) (ReLU)
 *   mode 0 (fprop): m enumerates (b, ho, wo) of the OUTPUT grid (Ho, Wo); the gather reads pixel
 *                   (ho*sh - ph + i*dh, wo*sw - pw + j*dw) of the (H, W, C) source, zero outside.
 *   mode 1 (dgrad): m enumerates (b, h, w) of the INPUT grid passed as (Ho, Wo); the gather reads
 *                   dy at ((h + ph - i*dh)/sh, (w + pw - j*dw)/sw) of the (H, W, C) = dy grid
 *                   when divisible and in range.
 *   a_ld / out_ld / res_ld: elements between consecutive pixels (channel slices are allowed).
 *   Constraints: C, a_ld, out_ld, res_ld multiples of 8; N multiple of 16; 16-byte aligned bases (bias included).
 *   1x1 / stride-1 layers take a TMA path (the A tile is a plain 2-D box of the [pixels][channels] matrix).
 *   `flags` is reserved and must be 0.
 * --------------------------------------------------------------------------------------------- */
int air_conv_block_n(int N);
long long air_conv_packed_elems(int N, int K);
int air_conv_pack_weights(const float* w, void* dst, int N, int K, int mode, int Cin, int Cout, int taps,
                          air_stream_t stream);
int air_conv_gemm_bf16(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                       int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                       const void* wpk, int N, int K, void* out, long long out_ld,
                       const float* bias, const void* res, long long res_ld, int relu,
                       int num_sms, int flags, air_stream_t stream);

/* air_conv_gemm_bf16 whose epilogue also adds the per-channel sum / sum of squares of the stored bf16 output to
 * stats[0..N) / stats[N..2N) (fp64, caller-zeroed): the batch statistics of the BatchNorm that follows (resnet.py:65-68) */
int air_conv_gemm_bf16_stats(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                             int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                             const void* wpk, int N, int K, void* out, long long out_ld,
                             const float* bias, const void* res, long long res_ld, int relu, double* stats,
                             int num_sms, int flags, air_stream_t stream);

/* Extended forms used by the ECAPA path:
 *   air_conv_pack_weights_ld : `w` rows are `w_ld` elements apart (a column slice of a wider weight, e.g. the
 *                              x-block of attention.0.weight (128, 4608), ecapa_tdnn.py:139,173-175)
 *   air_conv_gemm_bf16_ex    : bias_rows > 0 selects a per-utterance bias bias[(m / bias_rows) * N + n]
 *                              (the folded mean/std columns of attention.0); out2 (optional) receives the
 *                              accumulator (+bias) WITHOUT the residual (Res2 branch gradients, :77-80). */
int air_conv_pack_weights_ld(const float* w, long long w_ld, void* dst, int N, int K, int mode, int Cin, int Cout,
                             int taps, air_stream_t stream);
int air_conv_gemm_bf16_ex(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                          int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                          const void* wpk, int N, int K, void* out, long long out_ld,
                          const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                          void* out2, long long out2_ld, int num_sms, int flags, air_stream_t stream);
/* as _ex plus an optional per-channel affine after the ReLU (out = relu(acc + bias) * post_scale[n] + post_shift[n]):
 * eval-mode BatchNorm of the conv -> ReLU -> BN blocks of ecapa_tdnn.py:67-69,87-89,156-158 folded into the conv epilogue */
int air_conv_gemm_bf16_affine(const void* a, long long a_ld, int B, int H, int W, int C, int Ho, int Wo,
                              int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw, int mode,
                              const void* wpk, int N, int K, void* out, long long out_ld,
                              const float* bias, int bias_rows, const void* res, long long res_ld, int relu,
                              void* out2, long long out2_ld, const float* post_scale, const float* post_shift,
                              int num_sms, int flags, air_stream_t stream);

/* 3x3 / stride 1 / pad 1 convolution with a shared-memory resident input patch (csrc/conv_patch.cu): one TMA box
 * load brings a (2+2) x (128+2) pixel patch (zero padding = TMA out-of-bounds fill) and the nine taps are shifted
 * UMMA descriptor windows of it, so every activation crosses L2 -> SM about twice instead of nine times.  Used for
 * the ResNet 3x3 layers (resnet.py:56-60), forward (mode-0 weights) and data gradient (mode-1 weights: flipped
 * taps, swapped channels).
 *   a (B,H,W,C) channels-last bf16, out / res (B,H,W,N); C in {16, 32} or C % 64 == 0; N % 16 == 0, N <= 256.
 *   w for air_conv3x3_pack_weights: fp32 GEMM layout [Cout][3][3][Cin]; mode 0: (C, N) = (Cin, Cout),
 *   mode 1: (C, N) = (Cout, Cin); dst holds 9*C*N bf16. */
int air_conv3x3_patch_supported(int C, int N, int H, int W);
/* channels per block of the packed patch-kernel weight slices for C input channels (16 / 32 / 64) */
int air_conv_patch_cb(int C);
int air_conv3x3_pack_weights(const float* w, void* dst, int C, int N, int mode, air_stream_t stream);
int air_conv3x3_patch_bf16(const void* a, long long a_ld, int B, int H, int W, int C,
                           const void* wpk, int N, void* out, long long out_ld,
                           const void* res, long long res_ld, int relu, int num_sms, air_stream_t stream);
int air_conv3x3_patch_stats_bf16(const void* a, long long a_ld, int B, int H, int W, int C,
                                 const void* wpk, int N, void* out, long long out_ld,
                                 const void* res, long long res_ld, int relu, double* stats,
                                 int num_sms, air_stream_t stream);

/* Batched weight packing (csrc/pack.cu): one launch re-packs every convolution weight of a model from a
 * device-resident table of 16 x int64 job records.  air_pack_job_* fill ONE host-side record with the same arguments
 * as air_conv_pack_weights_ld / air_conv_patch_pack_weights; the caller uploads the table once and calls air_pack_jobs
 * after every optimiser step.  max_total = largest packed element count among the jobs (record field 3). */
int air_pack_job_gemm(long long* rec, const float* w, long long w_ld, void* dst, int N, int K, int mode,
                      int Cin, int Cout, int taps);
int air_pack_job_patch(long long* rec, const float* w, void* dst, int C, int N, int taps, int mode);
int air_pack_jobs(const long long* jobs, int njobs, long long max_total, air_stream_t stream);

/* General form of the patch kernel: explicit tap table and output pixel mapping
 *   out[b, g*osh+oph, g'*osw+opw, n] = sum_t sum_c a[b, g+org_h+tap_dr[t], g'+org_w+tap_dc[t], c] * Wp[tap_slice[t]][n][c] (+res)(ReLU)
 * over the item grid g < GH, g' < GW (0 <= tap_dr, tap_dc <= 2; out-of-range reads are zero).  wpk holds wtaps slices
 * per channel block (air_conv_patch_pack_weights, taps = 9 or 1). */
int air_conv_patch_pack_weights(const float* w, void* dst, int C, int N, int taps, int mode, air_stream_t stream);
int air_conv_patch_taps_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                             const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                             const void* res, long long res_ld, int relu,
                             int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                             int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                             int num_sms, air_stream_t stream);
/* as above plus a per-channel fp32 bias, a second output WITHOUT the residual (out2), and column offsets tap_dc up to 8
 * (the patch is then 136 pixels wide): the dilated k = 3 Conv1d of the Res2 branches, ecapa_tdnn.py:50, forward
 * (bias, ReLU) and data gradient (mode-1 weights, res / out2).  `stats` (optional, N % 32 == 0): fp64 [2N], the
 * per-channel sum and sum of squares of the stored output are ADDED to it -- the batch statistics of the BatchNorm that
 * consumes `out` (air_bn_stats fused into the conv epilogue; resnet.py:65-68). */
int air_conv_patch_taps_ex_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                const void* res, long long res_ld, int relu, const float* bias,
                                void* out2, long long out2_ld, double* stats,
                                int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                int num_sms, air_stream_t stream);
/* Data gradient of a stride-2 layer (k = 3 / pad 1: resnet.py:56 conv1 of layer2-4.0; k = 1 / pad 0: the shortcut,
 * resnet.py:60-61) decomposed by output parity so that only structurally non-zero taps are multiplied.
 * dy (B,Ho,Wo,Cout) -> dx (B,H,W,Cin); wpk = mode-1 packed weights with k*k taps.  k = 1 writes only the even-even
 * pixels (pass res = dx to accumulate into an existing gradient). */
int air_conv_s2_dgrad_patch_bf16(const void* dy, long long dy_ld, int B, int Ho, int Wo, int Cout,
                                 const void* wpk, int k, int Cin, void* dx, long long dx_ld, int H, int W,
                                 const void* res, long long res_ld, int num_sms, air_stream_t stream);

/* Weight gradient of the same 3x3 / stride 1 / pad 1 layers from shared-memory resident patches
 * (csrc/conv_wgrad_patch.cu): dW[co][i][j][ci] += sum_{b,h,w} x[b,h+i-1,w+j-1,ci] * dy[b,h,w,co], accumulated with fp32
 * atomics into the caller-zeroed dw_out ([Cout][dw_ld >= 9*Cin] fp32, GEMM layout).  C % 64 == 0, N % 64 == 0. */
int air_conv3x3_wgrad_patch_supported(int C, int N);
/* general form: k = 3 (3x3 / stride 1 / pad 1) or k = 1 (1x1 / stride 1 / pad 0); C = 16 (SWIZZLE_32B operand, eight
 * pixel shifts per M = 128 instruction) or a multiple of 64; N a multiple of 64; dw_out [N][dw_ld >= k*k*C]. */
int air_conv_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                              const void* dy, long long dy_ld, int N, int k,
                              float* dw_out, long long dw_ld, int num_sms, air_stream_t stream);
/* weight gradient of a 1-D convolution over W (kernel k <= 5 odd, dilation d, "same" padding, d*(k-1) <= 8; H rows are
 * independent sequences): dw_out [N][k][C] fp32, accumulated (ecapa_tdnn.py:50 and its autograd) */
/* stride-2 k x k (k = 3 pad 1 / k = 1 pad 0) weight gradient as one stride-1 patch problem per input parity class
 * (strided TMA sub-images, only the class's taps): x (B,H,W,C), dy (B,Ho,Wo,N), dw_out fp32 [N][dw_ld >= k*k*C] */
int air_conv_s2_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                 const void* dy, long long dy_ld, int Ho, int Wo, int N, int k,
                                 float* dw_out, long long dw_ld, int num_sms, air_stream_t stream);
int air_conv1d_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                const void* dy, long long dy_ld, int N, int k, int d,
                                float* dw_out, long long dw_ld, int num_sms, air_stream_t stream);
int air_conv3x3_wgrad_patch_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                                 const void* dy, long long dy_ld, int N,
                                 float* dw_out, long long dw_ld, int num_sms, air_stream_t stream);

/* Weight gradient (cuDNN wgrad in the reference's autograd): dW[n][kx] += sum_m dy[m][n] *
 * im2col(x)[m][kx], kx = (tap, ci), accumulated with fp32 atomics into the caller-zeroed
 * `dw_out` ([Cout][kh*kw*Cin] fp32, GEMM layout).  x: forward input (B,H,W,C) channels-last bf16,
 * dy: (B,Ho,Wo,N) channels-last bf16; pixel strides x_ld / dy_ld in elements. */
int air_conv_wgrad_bf16(const void* x, long long x_ld, int B, int H, int W, int C,
                        const void* dy, long long dy_ld, int Ho, int Wo, int N,
                        int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                        float* dw_out, int num_sms, int flags, air_stream_t stream);

/* as above with dw_out rows `dw_ld` elements apart (gradient of a column slice of a wider weight) */
int air_conv_wgrad_bf16_ld(const void* x, long long x_ld, int B, int H, int W, int C,
                           const void* dy, long long dy_ld, int Ho, int Wo, int N,
                           int kh, int kw, int sh, int sw, int ph, int pw, int dh, int dw,
                           float* dw_out, long long dw_ld, int num_sms, int flags, air_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm (+ReLU) on channels-last bf16 (nn.BatchNorm2d/1d + F.relu: resnet.py:54-69,132,141;
 * ecapa_tdnn.py:40,51,57,113 and their autograd backward).  eps / momentum as nn.BatchNorm.
 *   air_bn_stats : sums[0..C) += sum_m x, sums[C..2C) += sum_m x^2 (fp64, caller-zeroed)
 *   air_bn_apply : y = [relu](gamma*(x-mean)*invstd+beta); training != 0 uses `sums` (biased var),
 *                  saves mean/invstd and updates running stats (unbiased var); else running stats
 *   air_bn_bwd   : order 0 (y = relu(bn(x))): dx from dy = dL/dy;  order 1 (y = bn(x), x = relu(.)):
 *                  dx additionally masked by x > 0.  `add` (optional) is added to dx.  dgamma/dbeta are
 *                  accumulated (+=).  rsum: 2C fp64 workspace, caller-zeroed.
 * --------------------------------------------------------------------------------------------- */
int air_bn_stats(const void* x, long long x_ld, long long M, int C, double* sums, int num_sms, air_stream_t stream);
int air_bn_apply(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                 const double* sums, const float* gamma, const float* beta, float eps, int relu,
                 int training, float* save_mean, float* save_invstd, float* running_mean,
                 float* running_var, float momentum, int num_sms, air_stream_t stream);
int air_bn_bwd(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
               void* dx, long long dx_ld, long long M, int C, int order,
               const float* mean, const float* invstd, const float* gamma, const float* beta,
               double* rsum, float* dgamma, float* dbeta, int num_sms, air_stream_t stream);

/* air_bn_apply_add: air_bn_apply that also writes y2 = y + add (the Res2 branch input sp + spx[i+1],
 * ecapa_tdnn.py:77-80).  air_bn_bwd_bias: air_bn_bwd that also accumulates dbias[c] += sum_m dx[m][c], the
 * bias gradient of the convolution that produced x (Conv1d has bias=True throughout ecapa_tdnn.py). */
int air_bn_apply_add(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                     const double* sums, const float* gamma, const float* beta, float eps, int relu,
                     int training, float* save_mean, float* save_invstd, float* running_mean,
                     float* running_var, float momentum, const void* add, long long add_ld, void* y2,
                     long long y2_ld, int num_sms, air_stream_t stream);
int air_bn_bwd_bias(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                    void* dx, long long dx_ld, long long M, int C, int order,
                    const float* mean, const float* invstd, const float* gamma, const float* beta,
                    double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, air_stream_t stream);

/* ResNet stem conv (Cin = 1, Cout = 16; resnet.py:131,176): direct CUDA-core forward / wgrad. */
int air_stem_conv_fwd(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                      const float* w, int Cout, void* y, air_stream_t stream);
int air_stem_conv_wgrad(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                        const void* dy, int Cout, float* dw, air_stream_t stream);

/* SelfAttention pooling (resnet.py:23-46): x (B,T,C) bf16 -> stats (B,2C) fp32 = [avg | std];
 * saves softmax weights p (B,T) and tanh scores th (B,T).  noise_seed >= 0 adds the reference's
 * 1e-5*N(0,1) noise (counter-based, regenerated in the backward); < 0 disables it.
 * Backward: dx (B,T,C) bf16, datt (C) accumulated. */
int air_selfattn_pool_fwd(const void* x, const float* att, float* stats, float* p_out, float* th_out,
                          int B, int T, int C, long long noise_seed, air_stream_t stream);
int air_selfattn_pool_bwd(const void* x, const float* att, const float* p_in, const float* th_in,
                          const float* stats, const float* dstats, void* dx, float* datt,
                          int B, int T, int C, long long noise_seed, air_stream_t stream);

/* fp32 nn.Linear: y = x W^T + b;  backward: dx (optional), dW += , db += (optional). */
int air_linear_fwd(const float* x, const float* W, const float* bias, float* y, int M, int N, int K,
                   air_stream_t stream);
int air_linear_bwd(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db,
                   int M, int N, int K, air_stream_t stream);

/* air_linear_fwd with the K dimension dealt to `splits` (1..64) CTAs per output tile: fp64 accumulation, fp64 partial sums
 * in `scratch` (>= splits * M * N doubles, device), added in split order (deterministic).  For long reductions over few
 * outputs (fc6 of ecapa_tdnn.py:148, K = 3072), where a single CTA's chain of K tiles bounds the call. */
int air_linear_fwd_splitk(const float* x, const float* W, const float* bias, float* y, int M, int N, int K,
                          double* scratch, int splits, air_stream_t stream);

/* as above with W rows `ldw` elements apart; accumulate != 0 adds into y */
int air_linear_fwd_ld(const float* x, const float* W, long long ldw, const float* bias, float* y, int M, int N, int K,
                      int accumulate, air_stream_t stream);
int air_linear_bwd_ld(const float* x, const float* W, long long ldw, const float* dy, float* dx, float* dW, float* db,
                      int M, int N, int K, air_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ECAPA-TDNN (Res2Net2) non-GEMM stages on channels-last bf16 (B, T, C) activations
 * (ecapa_tdnn.py:15-29 SEModule, :64-95 Bottle2neck, :152-198 Res2Net2.forward) + autograd backward.
 *   air_time_stats_fwd     : mean_t (and sqrt(clamp(unbiased var_t, clampv)) if std_out) -> (B, C) fp32
 *                            (SE squeeze :25; context statistics :169-172)
 *   air_ecapa_asp_fwd      : w = softmax_t(e); out (B,2C) = [sum_t x w | sqrt(clamp(sum_t x^2 w - mu^2, 1e-4))]
 *                            (:176-186); saves max_t e, sum_t exp(e - max), q = sum_t x^2 w per (b, c)
 *   air_ecapa_asp_bwd      : de (softmax backward) and dx = attentive-statistics direct path + context-statistics path
 *   air_scale_residual_fwd : out = x * gate[b, c] + res                      (:27-28, :93)
 *   air_se_dgate           : dgate[b, c] = sum_t dout * x
 *   air_se_apply_bwd       : dx = dout * gate[b, c] + dmean[b, c] / T
 *   air_bn1d_f32_fwd/bwd   : fp32 BatchNorm1d over (M = batch, C) rows; relu_in applies ReLU to the input first
 *   air_sigmoid_fwd/bwd    : SE gate
 *   air_copy_channels      : dst = src [* (mask > 0)] on channel slices
 *   air_colsum_bf16        : out[c] += sum_m x[m][c] (conv bias gradients)
 * --------------------------------------------------------------------------------------------- */
int air_time_stats_fwd(const void* x, long long x_ld, int B, int T, int C, float* mean_out, float* std_out,
                       float clampv, air_stream_t stream);
int air_ecapa_asp_fwd(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                      float* out, float* save_max, float* save_sum, float* save_q, air_stream_t stream);
int air_ecapa_asp_bwd(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C,
                      const float* out, const float* dout, const float* save_max, const float* save_sum,
                      const float* save_q, const float* ctx_mean, const float* ctx_std, const float* dctx_mean,
                      const float* dctx_std, float clampv, void* de, long long de_ld, void* dx, long long dx_ld,
                      air_stream_t stream);
/* dx = (dx + dctx_mean/T + dctx_std (x - mean)/((T-1) std) [var > clampv]) * (x > 0): context-statistics backward
 * (ecapa_tdnn.py:169-172) fused with the ReLU mask of layer4 (:166).  air_time_stats_fwd with clampv < 0 and
 * std_out == NULL writes the plain sum over time instead of the mean. */
int air_ctx_stats_bwd_mask(const void* x, long long x_ld, int B, int T, int C, const float* ctx_mean,
                           const float* ctx_std, const float* dctx_mean, const float* dctx_std, float clampv,
                           void* dx, long long dx_ld, air_stream_t stream);
int air_scale_residual_fwd(const void* x, long long x_ld, const float* gate, const void* res, long long res_ld,
                           void* out, long long out_ld, int B, int T, int C, air_stream_t stream);
int air_se_dgate(const void* dout, long long d_ld, const void* x, long long x_ld, int B, int T, int C, float* dgate,
                 air_stream_t stream);
int air_se_apply_bwd(const void* dout, long long d_ld, const float* gate, const float* dmean, void* dx, long long dx_ld,
                     int B, int T, int C, air_stream_t stream);
int air_bn1d_f32_fwd(const float* x, float* y, int M, int C, int relu_in, const float* gamma, const float* beta,
                     float eps, int training, float* save_mean, float* save_invstd, float* running_mean,
                     float* running_var, float momentum, air_stream_t stream);
int air_bn1d_f32_bwd(const float* dy, const float* x, float* dx, int M, int C, int relu_in, const float* gamma,
                     const float* mean, const float* invstd, float* dgamma, float* dbeta, air_stream_t stream);
int air_sigmoid_fwd(const float* x, float* y, long long n, air_stream_t stream);
int air_sigmoid_bwd(const float* dy, const float* y, float* dx, long long n, air_stream_t stream);
int air_copy_channels(const void* src, long long s_ld, const void* mask, long long m_ld, void* dst, long long d_ld,
                      long long M, int C, air_stream_t stream);
int air_colsum_bf16(const void* x, long long ld, long long M, int C, float* out, air_stream_t stream);

/* OCSoftmax / AngularIsoLoss forward + analytic backward (loss.py:187-206 == :73-97) and the logged
 * CrossEntropy (main_train.py:355-357).  labels int64 (NULL: all 0).  Outputs (each optional):
 * loss (1), score (B) = -cos, dfeat (B,D) = grad_scale*dloss/dx, dcenter (D) += grad_scale*dloss/dc,
 * ce (1) from logits (B,ncls). */
int air_ocsoftmax_fwd_bwd(const float* x, const long long* labels, const float* center, int B, int D,
                          float r_real, float r_fake, float alpha, float grad_scale,
                          float* loss, float* score, float* dfeat, float* dcenter,
                          const float* logits, int ncls, float* ce, air_stream_t stream);

/* Optimiser steps on flat fp32 buffers: Adam with coupled L2 (main_train.py:175,408; `step` 1-based)
 * and SGD (main_train.py:272,409).  grad_scale multiplies the gradient (1/world_size under DP). */
int air_adam_l2_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, float grad_scale,
                     air_stream_t stream);
int air_sgd_step(float* p, const float* g, long long n, float lr, float grad_scale, air_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Detection-error curve, EER and min t-DCF of countermeasure scores (csrc/det.cu).  Replaces
 * eval_metrics.py:19-46 (compute_det_curve, compute_eer) and :141-172 (the t-DCF curve of compute_tDCF);
 * reference call sites main_train.py:662-663, score_fusion.py:117-118, evaluate_tDCF_asvspoof19.py:45-62.
 * target / nontarget: fp32 (_f32) or fp64 (_f64) device arrays (bona fide / spoof scores; the reference
 * feeds fp32 model scores and fp64 values parsed from score files).  negate != 0 scores -s instead
 * (the reference's `other_eer`).  c1, c2: the t-DCF weights C1, C2 (eval_metrics.py:160-162), used when
 * want_tdcf != 0.  Optional curve outputs, each n_tar + n_non + 1 doubles or NULL: frr, far, thresholds,
 * tdcf (normalised).  out: 10 doubles = EER, EER threshold, frr and far at that point, min normalised
 * t-DCF, its threshold, n_tar, n_non, argmin index of |frr - far|, argmin index of the t-DCF.
 * Results are bit-identical to numpy's on the fp64 values of the scores.  workspace: caller-owned,
 * air_det_workspace_bytes(n_tar + n_non) bytes.  air_det_launches: kernels launched per call.
 * air_det_threshold_counts_*: counts_ge_lt[0] = #(s >= threshold), [1] = #(s < threshold), the sums of
 * obtain_asv_error_rates (eval_metrics.py:4-16). */
int air_det_workspace_bytes(long long n, long long* bytes);
int air_det_launches(long long n, int is_f64);
int air_det_curve_f32(const float* target, long long n_tar, const float* nontarget, long long n_non,
                      int negate, double c1, double c2, int want_tdcf, void* workspace,
                      long long workspace_bytes, double* frr, double* far, double* thresholds,
                      double* tdcf, double* out, air_stream_t stream);
int air_det_curve_f64(const double* target, long long n_tar, const double* nontarget, long long n_non,
                      int negate, double c1, double c2, int want_tdcf, void* workspace,
                      long long workspace_bytes, double* frr, double* far, double* thresholds,
                      double* tdcf, double* out, air_stream_t stream);
int air_det_threshold_counts_f32(const float* scores, long long n, double threshold,
                                 unsigned long long* counts_ge_lt, air_stream_t stream);
int air_det_threshold_counts_f64(const double* scores, long long n, double threshold,
                                 unsigned long long* counts_ge_lt, air_stream_t stream);


/* ---------------------------------------------------------------------------------------------
 * Host-side audio ingest (csrc/audio_io.cpp, no CUDA): FLAC (decoder written from RFC 9639; CRC-8 /
 * CRC-16 checked, MD5 of the audio on request) and RIFF/WAVE (PCM 8/16/24/32, float32) -> float32
 * mono samples, scaled by 2^-(bits-1), channels averaged.  Replaces the reference's per-item
 * librosa.load / soundfile.read (raw_dataset.py:20-28,61-66).  flags bit 0: verify the FLAC MD5.
 * Status: 0, -1 argument, -2 unsupported stream, -3 I/O, -4 malformed stream, -5 checksum mismatch,
 * -6 out of memory.  No exception crosses the boundary; damaged input never crashes the process
 * (fuzzed under ASan / UBSan, scripts/fuzz_audio.py).
 * air_audio_decode_batch_f32: n files -> rows of a caller-owned (pinned) matrix with row stride ld,
 * zero-padded / truncated to ld, on `threads` host threads (<= 0: all cores); lengths[i] = the file's
 * own length, status[i] its code; returns the first non-zero status. */
int air_audio_info(const char* path, int* sample_rate, int* channels, int* bits, long long* frames);
int air_audio_decode_f32(const char* path, float* out, long long capacity, long long* frames,
                         int* sample_rate, int flags);
int air_audio_decode_i32(const char* path, int* out, long long capacity, long long* frames, int* channels,
                         int* bits, int* sample_rate, int flags);
int air_audio_decode_batch_f32(const char* const* paths, int n, float* out, long long ld, int* lengths,
                               int* sample_rates, int* status, int threads, int flags);
/* rows of a (pinned) float matrix from a packed int16 corpus: row i = blob[offsets[i] .. + min(lengths[i], ld)) / 32768,
 * zero-padded to ld (asvspoof2021_air_b200/data.py PackedWaves; decode once, train many epochs).  blob_samples = size of
 * the mapping in samples: every (offset, length) pair is checked against it (AIR_ERR_ARG). */
int air_audio_gather_i16_f32(const short* blob, long long blob_samples, const long long* offsets, const int* lengths, int n,
                             float* out, long long ld, int threads);


/* ---------------------------------------------------------------------------------------------
 * Adversarial channel-classifier head (csrc/adv.cu): the stages between / after the two nn.Linear of
 * ChannelClassifier (model.py:1002-1023) and its CrossEntropyLoss (main_train.py:251,377-403,420-453).
 *   air_dropout_relu_fwd : y = relu(x * keep / (1 - p)); keep (bytes) is read, or drawn and written when generate != 0
 *   air_dropout_relu_bwd : dx = dy * [x * keep > 0] * keep / (1 - p)
 *   air_relu_ce_fwd_bwd  : logits = relu(z) (B, C); *loss_sum += mean CE; *correct += #(argmax == label) (first maximum);
 *                          dz = grad_scale * dCE/dz (optional).  loss_sum / correct accumulate: zero them first.
 * Not yet validated on hardware (round 1 ended without GPU time for tests/test_adv_gpu.py). */
int air_dropout_relu_fwd(const float* x, unsigned char* keep, float* y, long long n, float p, int generate,
                         unsigned long long seed, air_stream_t stream);
int air_dropout_relu_bwd(const float* dy, const float* x, const unsigned char* keep, float* dx, long long n,
                         float p, air_stream_t stream);
int air_relu_ce_fwd_bwd(const float* z, const long long* labels, int B, int C, float grad_scale,
                        double* loss_sum, int* correct, float* dz, air_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 parity mode.  BASELINE.json asks for logits / scores within 1e-3 of the reference's fp32
 * CPU path; a network with bf16 storage between its layers cannot get there (the fp32 and bf16-storage ORACLES differ by
 * 1e-2).  In this mode activations and gradients are stored as float, and every convolution still runs on the same
 * tcgen05 kernels: its fp32 operands are written as bf16 split terms concatenated along the contraction axis
 * (air_split_terms: x = hi + lo, x*w ~= hi*hi + lo*hi + hi*lo with one fp32 accumulator in TMEM), and the epilogue
 * stores the accumulator unrounded (AIR_CONV_F32_OUT).  The *_f32 entry points are the float-storage instances of the
 * HBM-bound kernels above: same arguments, activation pointers are float*.
 * --------------------------------------------------------------------------------------------- */
#define AIR_CONV_F32_OUT 1 /* `flags` bit of the conv entry points: out / res / out2 are float tensors */
int air_split_terms(const float* x, long long x_ld, long long M, int C, void* out, long long out_ld, int out_f32,
                    int nterms, unsigned lo_mask, air_stream_t stream);
int air_conv_patch_taps_ex2_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                 const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                 const void* res, long long res_ld, int relu, const float* bias,
                                 void* out2, long long out2_ld, double* stats,
                                 int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                 int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                 int flags, int num_sms, air_stream_t stream);
/* as air_conv_patch_taps_ex2_bf16 plus an optional per-channel affine applied right after the ReLU (bf16 storage only):
 * t = relu(acc + bias) * post_scale + post_shift; out2 <- t; out <- round_bf16(t) + res -- conv -> ReLU -> eval-mode
 * BatchNorm1d with the running-statistics affine folded in (ecapa_tdnn.py:73-83; generate_score.py's forward pass) */
int air_conv_patch_taps_ex3_bf16(const void* a, long long a_ld, int B, int Hin, int Win, int C,
                                 const void* wpk, int wtaps, int N, void* out, long long out_ld, int OH, int OW,
                                 const void* res, long long res_ld, int relu, const float* bias,
                                 const float* post_scale, const float* post_shift,
                                 void* out2, long long out2_ld, double* stats,
                                 int GH, int GW, int org_h, int org_w, int osh, int osw, int oph, int opw,
                                 int ntaps, const int* tap_dr, const int* tap_dc, const int* tap_slice,
                                 int flags, int num_sms, air_stream_t stream);
int air_conv_s2_dgrad_patch_ex_bf16(const void* dy, long long dy_ld, int B, int Ho, int Wo, int Cout,
                                    const void* wpk, int k, int Cin, void* dx, long long dx_ld, int H, int W,
                                    const void* res, long long res_ld, int flags, int num_sms, air_stream_t stream);
int air_bn_stats_f32(const void* x, long long x_ld, long long M, int C, double* sums, int num_sms, air_stream_t stream);
int air_bn_apply_add_f32(const void* x, long long x_ld, void* y, long long y_ld, long long M, int C,
                         const double* sums, const float* gamma, const float* beta, float eps, int relu,
                         int training, float* save_mean, float* save_invstd, float* running_mean,
                         float* running_var, float momentum, const void* add, long long add_ld, void* y2,
                         long long y2_ld, int num_sms, air_stream_t stream);
int air_bn_bwd_bias_f32(const void* dy, long long dy_ld, const void* x, long long x_ld, const void* add, long long add_ld,
                        void* dx, long long dx_ld, long long M, int C, int order,
                        const float* mean, const float* invstd, const float* gamma, const float* beta,
                        double* rsum, float* dgamma, float* dbeta, float* dbias, int num_sms, air_stream_t stream);
int air_stem_conv_fwd_f32(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                          const float* w, int Cout, void* y, air_stream_t stream);
int air_stem_conv_wgrad_f32(const void* x, int B, int H, int W, int kh, int kw, int sh, int sw, int ph, int pw,
                            const void* dy, int Cout, float* dw, air_stream_t stream);
int air_selfattn_pool_fwd_f32(const void* x, const float* att, float* stats, float* p_out, float* th_out,
                              int B, int T, int C, long long noise_seed, air_stream_t stream);
int air_selfattn_pool_bwd_f32(const void* x, const float* att, const float* p_in, const float* th_in,
                              const float* stats, const float* dstats, void* dx, float* datt,
                              int B, int T, int C, long long noise_seed, air_stream_t stream);

/* float-storage instances of the ECAPA non-GEMM stages (same arguments as the bf16 entry points above) */
int air_time_stats_fwd_f32(const void* x, long long x_ld, int B, int T, int C, float* mean_out, float* std_out, float
    clampv, air_stream_t stream);
int air_ecapa_asp_fwd_f32(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C, float*
    out, float* save_max, float* save_sum, float* save_q, air_stream_t stream);
int air_ecapa_asp_bwd_f32(const void* e, long long e_ld, const void* x, long long x_ld, int B, int T, int C, const
    float* out, const float* dout, const float* save_max, const float* save_sum, const float* save_q, const float*
    ctx_mean, const float* ctx_std, const float* dctx_mean, const float* dctx_std, float clampv, void* de, long long
    de_ld, void* dx, long long dx_ld, air_stream_t stream);
int air_ctx_stats_bwd_mask_f32(const void* x, long long x_ld, int B, int T, int C, const float* ctx_mean, const float*
    ctx_std, const float* dctx_mean, const float* dctx_std, float clampv, void* dx, long long dx_ld, air_stream_t
    stream);
int air_scale_residual_fwd_f32(const void* x, long long x_ld, const float* gate, const void* res, long long res_ld,
    void* out, long long out_ld, int B, int T, int C, air_stream_t stream);
int air_se_dgate_f32(const void* dout, long long d_ld, const void* x, long long x_ld, int B, int T, int C, float*
    dgate, air_stream_t stream);
int air_se_apply_bwd_f32(const void* dout, long long d_ld, const float* gate, const float* dmean, void* dx, long long
    dx_ld, int B, int T, int C, air_stream_t stream);
int air_copy_channels_f32(const void* src, long long s_ld, const void* mask, long long m_ld, void* dst, long long
    d_ld, long long M, int C, air_stream_t stream);
int air_colsum_f32(const void* x, long long ld, long long M, int C, float* out, air_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AIR_B200_H */
