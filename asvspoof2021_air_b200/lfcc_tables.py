"""Host-side constant tables for the fused LFCC kernel (csrc/lfcc.cu).

The filterbank and DCT matrix are built with the same torch calls as the reference
(feature_extraction.py:77-87, utils_dsp.py:147-176,234-244) so that the registered buffers
`lfcc_fb` / `l_dct.weight` are bit-identical to the reference module's; the kernel consumes a
packed float32 table derived from them (sparse filter rows, twiddles in float64 -> float32).
"""
import math

import numpy as np
import torch

NF = 20
FL, FS, FN = 320, 160, 512
OFF_WIN = 0
OFF_TW1 = OFF_WIN + FL
OFF_TW2 = OFF_TW1 + 512
OFF_FBW = OFF_TW2 + 512
OFF_FBS = OFF_FBW + NF * 33
OFF_FBC = OFF_FBS + NF
OFF_DCT = OFF_FBC + NF
TBL_FLOATS = OFF_DCT + NF * 21


def trimf(x, params):
    """MATLAB-style triangular membership, strict inequalities (feature_extraction.py:16-39)."""
    a, b, c = params
    if a > b or b > c:
        raise ValueError("trimf(x, [a, b, c]) requires a<=b<=c")
    y = torch.zeros_like(x, dtype=torch.float32)
    if a < b:
        index = (a < x) & (x < b)
        y[index] = (x[index] - a) / (b - a)
    if b < c:
        index = (b < x) & (x < c)
        y[index] = (c - x[index]) / (c - b)
    y[x == b] = 1
    return y


def linear_filterbank(fn, sr, filter_num):
    """(fn//2+1, filter_num) float32 (feature_extraction.py:77-86)."""
    f = (sr / 2) * torch.linspace(0, 1, fn // 2 + 1)
    filter_bands = torch.linspace(min(f), max(f), filter_num + 2)
    filter_bank = torch.zeros([fn // 2 + 1, filter_num])
    for idx in range(filter_num):
        filter_bank[:, idx] = trimf(f, [filter_bands[idx], filter_bands[idx + 1], filter_bands[idx + 2]])
    return filter_bank


def dct_ortho_matrix(n):
    """LinearDCT(n, 'dct', norm='ortho').weight: DCT-II of the identity through an FFT, float32
    (utils_dsp.py:147-176 with torch.rfft(v, 1, onesided=False) == view_as_real(fft(v)))."""
    x = torch.eye(n)
    v = torch.cat([x[:, ::2], x[:, 1::2].flip([1])], dim=1)
    Vc = torch.view_as_real(torch.fft.fft(v, dim=-1))
    k = -torch.arange(n, dtype=x.dtype)[None, :] * np.pi / (2 * n)
    W_r, W_i = torch.cos(k), torch.sin(k)
    V = Vc[:, :, 0] * W_r - Vc[:, :, 1] * W_i
    V[:, 0] /= np.sqrt(n) * 2
    V[:, 1:] /= np.sqrt(n / 2) * 2
    V = 2 * V
    return V.t().contiguous()


def pack_table(lfcc_fb: torch.Tensor, dct_weight: torch.Tensor, fl=FL) -> torch.Tensor:
    """Pack window, FFT twiddles, sparse filterbank rows and the DCT matrix for csrc/lfcc.cu."""
    fb = lfcc_fb.detach().cpu().float().numpy()
    assert fb.shape == (FN // 2 + 1, NF) and tuple(dct_weight.shape) == (NF, NF) and fl == FL
    tbl = np.zeros(TBL_FLOATS, dtype=np.float32)
    tbl[OFF_WIN:OFF_WIN + FL] = torch.hamming_window(FL).numpy()        # periodic, feature_extraction.py:110
    kj = np.arange(16)[:, None].astype(np.float64)
    l = np.arange(16)[None, :].astype(np.float64)
    ang = -2.0 * math.pi * kj * l / 256.0
    tw1 = np.stack([np.cos(ang), np.sin(ang)], axis=-1)                 # [kj][l] -> W256^(l*kj)
    tbl[OFF_TW1:OFF_TW1 + 512] = tw1.reshape(-1).astype(np.float32)
    k = np.arange(256, dtype=np.float64)
    tw2 = np.stack([np.cos(2 * math.pi * k / 512), np.sin(2 * math.pi * k / 512)], axis=-1)
    tbl[OFF_TW2:OFF_TW2 + 512] = tw2.reshape(-1).astype(np.float32)
    starts = np.zeros(NF, dtype=np.int32)
    counts = np.zeros(NF, dtype=np.int32)
    for f in range(NF):
        nz = np.nonzero(fb[:, f])[0]
        assert len(nz) > 0 and nz[0] >= 1 and nz[-1] <= 255, "bins 0/256 must carry no weight"
        starts[f], counts[f] = nz[0], nz[-1] - nz[0] + 1
        assert counts[f] <= 32
        tbl[OFF_FBW + f * 33:OFF_FBW + f * 33 + counts[f]] = fb[nz[0]:nz[-1] + 1, f]
    tbl[OFF_FBS:OFF_FBS + NF] = starts.view(np.float32)
    tbl[OFF_FBC:OFF_FBC + NF] = counts.view(np.float32)
    d = dct_weight.detach().cpu().float().numpy()
    for kk in range(NF):
        tbl[OFF_DCT + kk * 21:OFF_DCT + kk * 21 + NF] = d[kk]
    return torch.from_numpy(tbl)


# ---------------------------------------------------------------------------------------------
# tables of the tensor-core path (csrc/lfcc_tc.cu)
# ---------------------------------------------------------------------------------------------
TC_OFF_WIN, TC_OFF_FBW, TC_OFF_DCT = 0, 320, 320 + 512
TC_TBL_FLOATS = TC_OFF_DCT + NF * NF
TC_KBLK, TC_CHUNK_ELEMS = 5, 128 * 32


def tc_filter_structure_ok(lfcc_fb) -> bool:
    """The kernel hard-codes WHICH filters a bin feeds (floor(21k/256)-1 and floor(21k/256), the support of the
    reference's trimf bands, feature_extraction.py:77-85) and takes the weights from the module's buffer."""
    fb = lfcc_fb.detach().cpu().float().numpy()
    if fb.shape != (FN // 2 + 1, NF) or np.count_nonzero(fb[0]) or np.count_nonzero(fb[256]):
        return False
    for k in range(1, 256):
        fh = (21 * k) >> 8
        if not set(np.nonzero(fb[k])[0].tolist()) <= {fh - 1, fh}:
            return False
    return True


def pack_tc_table(lfcc_fb: torch.Tensor, dct_weight: torch.Tensor) -> torch.Tensor:
    """fp32 table: periodic Hamming window (feature_extraction.py:110), per-bin filter weight pairs, DCT matrix."""
    assert tc_filter_structure_ok(lfcc_fb)
    fb = lfcc_fb.detach().cpu().float().numpy()
    tbl = np.zeros(TC_TBL_FLOATS, dtype=np.float32)
    tbl[TC_OFF_WIN:TC_OFF_WIN + FL] = torch.hamming_window(FL).numpy()
    for k in range(1, 256):
        fh = (21 * k) >> 8
        if fh - 1 >= 0:
            tbl[TC_OFF_FBW + 2 * k] = fb[k, fh - 1]
        if fh < NF:
            tbl[TC_OFF_FBW + 2 * k + 1] = fb[k, fh]
    tbl[TC_OFF_DCT:TC_OFF_DCT + NF * NF] = dct_weight.detach().cpu().float().numpy().reshape(-1)
    return torch.from_numpy(tbl)


def pack_tc_dft() -> torch.Tensor:
    """3-term operand split of the folded real-DFT matrices cos / sin (2 pi k m / 512), k = 1..256, m = 0..159, as pre-swizzled
    (SWIZZLE_64B, K-major) [128 bins][32 samples] operand chunks in MMA order [Re/Im][half][kb][term]:
        term 0: w_hi = fp16(w)            (multiplies x_hi, fp16 x fp16)
        term 1: w_b  = bf16(w)            (multiplies x_lo, bf16 x bf16)
        term 2: w_lo = fp16(w - w_hi)     (multiplies x_hi, fp16 x fp16; fp16 subnormals: absolute precision 2^-24)
    The returned tensor is 16-bit storage typed bf16: the chunks of terms 0 and 2 hold fp16 BIT PATTERNS."""
    k = np.arange(1, 257, dtype=np.float64)[:, None]
    m = np.arange(160, dtype=np.float64)[None, :]
    ang = 2.0 * math.pi * k * m / 512.0
    out = torch.zeros(2, 2, TC_KBLK, 3, TC_CHUNK_ELEMS, dtype=torch.bfloat16)
    n = np.arange(128)[:, None]
    kk = np.arange(32)[None, :]
    off = n * 64 + kk * 2
    idx = torch.from_numpy(((off ^ (((off >> 7) & 3) << 4)) >> 1).reshape(-1).astype(np.int64))
    for part, mat in enumerate((np.cos(ang), np.sin(ang))):
        full = torch.from_numpy(mat.astype(np.float32))
        hi16 = full.to(torch.float16)
        lo16 = (torch.from_numpy(mat) - hi16.double()).float().to(torch.float16)
        terms = (hi16.view(torch.bfloat16), full.to(torch.bfloat16), lo16.view(torch.bfloat16))
        for h in range(2):
            for kb in range(TC_KBLK):
                for which, src in enumerate(terms):
                    blk = src[128 * h:128 * h + 128, 32 * kb:32 * kb + 32].reshape(-1)
                    out[part, h, kb, which, idx] = blk
    return out.reshape(-1)
