"""float64 numpy emulation of the lane/register algorithm of csrc/lfcc.cu (index math only).

Host-side check that the 16x16 Cooley-Tukey split, the twiddle table layout, the mirror-lane
real-FFT split and the sparse filterbank tables are right before any GPU time is spent."""
import numpy as np

from asvspoof2021_air_b200 import lfcc_tables as lt


def fft16(v):
    """v (..., 16) complex, natural order -> natural order; same radix-4x4 split as the kernel."""
    W16 = np.exp(-2j * np.pi * np.arange(16) / 16)
    T = np.zeros(v.shape[:-1] + (4, 4), dtype=complex)
    for n1 in range(4):
        x = [v[..., n1 + 4 * n2] for n2 in range(4)]
        t0, t1, t2, t3 = x[0] + x[2], x[0] - x[2], x[1] + x[3], x[1] - x[3]
        T[..., n1, 0] = t0 + t2
        T[..., n1, 2] = t0 - t2
        T[..., n1, 1] = t1 + (-1j) * t3
        T[..., n1, 3] = t1 - (-1j) * t3
    for n1 in range(4):
        for k2 in range(4):
            T[..., n1, k2] *= W16[(n1 * k2) % 16]
    out = np.zeros_like(v)
    for k2 in range(4):
        x = [T[..., n1, k2] for n1 in range(4)]
        t0, t1, t2, t3 = x[0] + x[2], x[0] - x[2], x[1] + x[3], x[1] - x[3]
        out[..., k2] = t0 + t2
        out[..., 8 + k2] = t0 - t2
        out[..., 4 + k2] = t1 + (-1j) * t3
        out[..., 12 + k2] = t1 - (-1j) * t3
    return out


def frame_cepstrum(y_frame, tbl):
    """y_frame: 320 pre-emphasised samples (zeros outside the utterance) -> 20 cepstra."""
    tbl = np.asarray(tbl, dtype=np.float32)
    win = tbl[lt.OFF_WIN:lt.OFF_WIN + 320].astype(np.float64)
    tw1 = tbl[lt.OFF_TW1:lt.OFF_TW1 + 512].astype(np.float64).reshape(16, 16, 2)
    tw2 = tbl[lt.OFF_TW2:lt.OFF_TW2 + 512].astype(np.float64).reshape(256, 2)
    yw = y_frame * win
    z = np.zeros(256, dtype=complex)
    z[:160] = yw[0::2] + 1j * yw[1::2]
    v = np.zeros((16, 16), dtype=complex)          # [lane l][reg j] = z[l + 16 j]
    for l in range(16):
        for j in range(16):
            v[l, j] = z[l + 16 * j]
    v = fft16(v)                                    # [l][kj]
    for l in range(16):
        for kj in range(16):
            v[l, kj] *= tw1[kj, l, 0] + 1j * tw1[kj, l, 1]
    v = v.T.copy()                                  # smem transpose: lane q=kj holds [i=l]
    v = fft16(v)                                    # v[q][r] = Z[q + 16 r]
    P = np.zeros(256)
    for q in range(16):
        pl = (16 - q) & 15
        for r in range(16):
            m = v[q, (16 - r) & 15] if q == 0 else v[pl, 15 - r]
            A = v[q, r]
            sx, sy = A.real + m.real, A.imag - m.imag
            dx, dy = A.real - m.real, A.imag + m.imag
            c, s = tw2[q + 16 * r]
            re = sx - s * dx + c * dy
            im = sy - s * dy - c * dx
            P[q + 16 * r] = 0.25 * (re * re + im * im)
    starts = tbl[lt.OFF_FBS:lt.OFF_FBS + lt.NF].view(np.int32)
    counts = tbl[lt.OFF_FBC:lt.OFF_FBC + lt.NF].view(np.int32)
    fbe = np.zeros(lt.NF)
    for f in range(lt.NF):
        wts = tbl[lt.OFF_FBW + f * 33:lt.OFF_FBW + f * 33 + counts[f]].astype(np.float64)
        fbe[f] = np.log10(np.dot(P[starts[f]:starts[f] + counts[f]], wts) + 1.1920928955078125e-07)
    dct = np.stack([tbl[lt.OFF_DCT + k * 21:lt.OFF_DCT + k * 21 + lt.NF] for k in range(lt.NF)]).astype(np.float64)
    return dct @ fbe, P


def lfcc_emulated(wave, tbl):
    """wave (L,) -> (T, 60) following the kernel's framing / delta rules."""
    x = np.asarray(wave, dtype=np.float64)
    L = len(x)
    y = x.copy()
    y[1:] = x[1:] - 0.97 * x[:-1]
    T = 1 + L // 160
    c = np.zeros((T, 20))
    for t in range(T):
        fr = np.zeros(320)
        s0 = 160 * (t - 1)
        lo, hi = max(s0, 0), min(s0 + 320, L)
        if hi > lo:
            fr[lo - s0:hi - s0] = y[lo:hi]
        c[t], _ = frame_cepstrum(fr, tbl)
    cl = lambda u: min(max(u, 0), T - 1)
    out = np.zeros((T, 60))
    for t in range(T):
        tp, tm = cl(t + 1), cl(t - 1)
        out[t, :20] = c[t]
        out[t, 20:40] = c[tp] - c[tm]
        out[t, 40:] = (c[cl(tp + 1)] - c[cl(tp - 1)]) - (c[cl(tm + 1)] - c[cl(tm - 1)])
    return out


# ---------------------------------------------------------------------------------------------------
# tensor-core path (csrc/lfcc_tc.cu): folded real DFT with the 3-term split x_hi*w_hi (fp16) + x_lo*bf16(w) (bf16) + x_hi*w_lo
# (fp16), emulated from the PACKED tables
# ---------------------------------------------------------------------------------------------------
def _bf16(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def _f16(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.float16).float().numpy()


def unpack_tc_dft(wmat):
    """Inverse of lfcc_tables.pack_tc_dft: -> (C_hi, C_b, C_lo, S_hi, S_b, S_lo), each (256 bins k=1..256, 160 samples)
    float32.  Terms 0 / 2 of the 16-bit table hold fp16 bit patterns (w_hi, w_lo), term 1 bf16 values (bf16(w))."""
    import torch
    w16 = wmat.reshape(2, 2, lt.TC_KBLK, 3, lt.TC_CHUNK_ELEMS)
    w = np.zeros(w16.shape, dtype=np.float32)
    for term in (0, 2):
        w[:, :, :, term] = w16[:, :, :, term].contiguous().view(torch.float16).float().numpy()
    w[:, :, :, 1] = w16[:, :, :, 1].float().numpy()
    n = np.arange(128)[:, None]
    kk = np.arange(32)[None, :]
    off = n * 64 + kk * 2
    idx = (off ^ (((off >> 7) & 3) << 4)) >> 1
    mats = np.zeros((2, 3, 256, 160), dtype=np.float32)                        # [part][term][k][m]
    for part in range(2):
        for h in range(2):
            for kb in range(lt.TC_KBLK):
                for which in range(3):
                    mats[part, which, 128 * h:128 * h + 128, 32 * kb:32 * kb + 32] = w[part, h, kb, which][idx]
    return mats[0, 0], mats[0, 1], mats[0, 2], mats[1, 0], mats[1, 1], mats[1, 2]


def lfcc_tc_emulate(wave, tbl, wmat, preemph=0.97):
    """(B, L) float32 -> (B, T, 20) cepstra with the arithmetic of the tensor-core kernel (fp32 accumulation emulated
    by float64 sums rounded once, which only makes this emulation slightly MORE exact than the device)."""
    tbl = np.asarray(tbl, dtype=np.float32)
    win = tbl[lt.TC_OFF_WIN:lt.TC_OFF_WIN + 320]
    fbw = tbl[lt.TC_OFF_FBW:lt.TC_OFF_FBW + 512].reshape(256, 2)
    dct = tbl[lt.TC_OFF_DCT:lt.TC_OFF_DCT + 400].reshape(20, 20)
    Chi, Cb, Clo, Shi, Sb, Slo = (m.astype(np.float64) for m in unpack_tc_dft(wmat))
    wave = np.asarray(wave, dtype=np.float32)
    B, L = wave.shape
    T = 1 + L // 160
    y = wave.copy()
    y[:, 1:] = wave[:, 1:] - np.float32(preemph) * wave[:, :-1]
    ypad = np.zeros((B, 160 * (T + 2)), np.float32)
    ypad[:, 160:160 + L] = y
    kbin = np.arange(1, 257)
    c160 = np.cos(5 * np.pi * (kbin % 16) / 8).astype(np.float32)
    s160 = np.sin(5 * np.pi * (kbin % 16) / 8).astype(np.float32)
    mm = np.arange(1, 160)
    out = np.zeros((B, T, 20), np.float32)
    for t in range(T):
        a = ypad[:, 160 * t:160 * t + 320] * win
        e = np.zeros((B, 160), np.float32)
        o = np.zeros((B, 160), np.float32)
        e[:, 0] = a[:, 160]
        e[:, 1:] = a[:, 160 + mm] + a[:, 160 - mm]
        o[:, 1:] = a[:, 160 + mm] - a[:, 160 - mm]
        ehi, ohi = _f16(e), _f16(o)                               # hi parts fp16, residuals bf16 (csrc/lfcc_tc.cu)
        elo, olo = _bf16(e - ehi), _bf16(o - ohi)
        re = (ehi @ Chi.T + elo @ Cb.T + ehi @ Clo.T).astype(np.float32) + a[:, :1] * c160
        im = (ohi @ Shi.T + olo @ Sb.T + ohi @ Slo.T).astype(np.float32) - a[:, :1] * s160
        P = re * re + im * im
        fb = np.zeros((B, 22), np.float32)
        for k in range(1, 256):
            fh = (21 * k) >> 8
            fb[:, fh] += fbw[k, 0] * P[:, k - 1]
            fb[:, fh + 1] += fbw[k, 1] * P[:, k - 1]
        fbe = np.log10(fb[:, 1:21] + np.float32(1.1920929e-07))
        out[:, t] = fbe @ dct.T
    return out
