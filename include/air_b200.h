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

/* Library version (major*10000 + minor*100 + patch). */
int air_version(void);

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

#ifdef __cplusplus
}
#endif
#endif /* AIR_B200_H */
