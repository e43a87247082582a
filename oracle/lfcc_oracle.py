"""CPU oracle (float64 numpy) for the LFCC front-end and the crop/pad index policy.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg as the *checker*; the product path
(asvspoof2021_air_b200/) never imports anything from oracle/.

Parity status: the reference ships no golden vectors for this path, so the
oracle is pinned against outputs of the reference's own code run in the authoring container
(oracle/ref_shim.py + oracle/make_golden.py -> tests/golden/lfcc_golden.npz).

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
import numpy as np
import torch

FLT_EPS = float(np.finfo(np.float32).eps)  # torch.finfo(torch.float32).eps, feature_extraction.py:117


def trimf(x, a, b, c):
    """Triangular membership function, MATLAB style, strict inequalities.

    feature_extraction.py:16-39.  x: float32 tensor; a,b,c float32 scalars (tensors).
    Arithmetic stays in float32 torch so the constants are bit-identical to the reference's."""
    y = torch.zeros_like(x, dtype=torch.float32)
    if a < b:
        idx = (a < x) & (x < b)
        y[idx] = (x[idx] - a) / (b - a)
    if b < c:
        idx = (b < x) & (x < c)
        y[idx] = (c - x[idx]) / (c - b)
    y[x == b] = 1
    return y


def linear_filterbank(fn=512, sr=16000, filter_num=20):
    """(fn//2+1, filter_num) float32 triangular linear-frequency filterbank.

    feature_extraction.py:77-86."""
    f = (sr / 2) * torch.linspace(0, 1, fn // 2 + 1)
    bands = torch.linspace(float(f.min()), float(f.max()), filter_num + 2)
    fb = torch.zeros(fn // 2 + 1, filter_num)
    for i in range(filter_num):
        fb[:, i] = trimf(f, bands[i], bands[i + 1], bands[i + 2])
    return fb.numpy()


def dct_matrix(n=20):
    """Ortho DCT-II matrix W (n,n) such that y = x @ W.T, analytic float64.

    utils_dsp.py:147-176 computes the same matrix through an FFT of the identity
    (LinearDCT.reset_parameters, utils_dsp.py:234-244); the two agree to ~4e-8."""
    k = np.arange(n)[:, None].astype(np.float64)
    m = np.arange(n)[None, :].astype(np.float64)
    w = np.cos(np.pi * (2 * m + 1) * k / (2 * n)) * np.sqrt(2.0 / n)
    w[0, :] *= np.sqrt(0.5)
    return w


def hamming_periodic(fl=320):
    """torch.hamming_window(fl) default periodic=True (feature_extraction.py:110)."""
    n = np.arange(fl, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2 * np.pi * n / fl)


def num_frames(length, fs=160):
    """torch.stft(center=True): T = 1 + L // hop (feature_extraction.py:109)."""
    return 1 + length // fs


def delta(x):
    """out[t] = x[t+1] - x[t-1] with replicate edges, no 1/2 factor (feature_extraction.py:41-58)."""
    xp = np.concatenate([x[:, :1], x, x[:, -1:]], axis=1)
    return xp[:, 2:] - xp[:, :-2]


def lfcc(wave, fl=320, fs=160, fn=512, sr=16000, filter_num=20,
         with_emphasis=True, with_delta=True, fb=None, dct=None):
    """wave (B, L) -> (B, 1 + L//fs, 3*filter_num) float64.  feature_extraction.py:93-138.

    with_energy is False at every reference call site (dataset.py:13, preprocess.py:237) and is
    not restated."""
    x = np.asarray(wave, dtype=np.float64)
    B, L = x.shape
    if with_emphasis:                                   # :105-106 (RHS evaluated first: non-recursive)
        y = x.copy()
        y[:, 1:] = x[:, 1:] - 0.97 * x[:, :-1]
    else:
        y = x
    T = num_frames(L, fs)
    pad = fn // 2                                       # center=True, pad_mode="constant" :109-111
    yp = np.zeros((B, L + 2 * pad), dtype=np.float64)
    yp[:, pad:pad + L] = y
    win = np.zeros(fn, dtype=np.float64)                # window centred inside the n_fft buffer
    off = (fn - fl) // 2
    win[off:off + fl] = hamming_periodic(fl)
    idx = (np.arange(T) * fs)[:, None] + np.arange(fn)[None, :]
    frames = yp[:, idx] * win                           # (B, T, fn)
    spec = np.fft.rfft(frames, axis=-1)                 # (B, T, fn//2+1)
    power = spec.real ** 2 + spec.imag ** 2             # :113
    if fb is None:
        fb = linear_filterbank(fn, sr, filter_num)
    fbe = np.log10(power @ fb.astype(np.float64) + FLT_EPS)   # :116-117
    if dct is None:
        dct = dct_matrix(filter_num)
    c = fbe @ dct.T                                     # :120  (nn.Linear: x @ W.T)
    if not with_delta:
        return c
    d = delta(c)                                        # :130-133
    dd = delta(d)
    return np.concatenate([c, d, dd], axis=2)


# ---------------------------------------------------------------------------------------
# crop / pad policy between LFCC and the model (integer index work: must be bit-exact)
# ---------------------------------------------------------------------------------------
PAD_NONE, PAD_ZERO, PAD_REPEAT, PAD_SILENCE = 0, 1, 2, 3
SRC_ZERO, SRC_SILENCE = -1, -2


def frame_index_map(T, feat_len, padding="repeat", startp=0):
    """int64 (feat_len,) map: output frame j takes LFCC frame map[j];
    SRC_ZERO (-1) = zero frame, SRC_SILENCE (-2) = the silence LFCC vector.

    dataset.py:66-79 (policy), :513-528 (padding_Tensor, repeat_padding_Tensor,
    silence_padding_Tensor -- NB silence is PREPENDED, :528)."""
    j = np.arange(feat_len, dtype=np.int64)
    if T > feat_len:                                    # :67-69  startp = np.random.randint(T - feat_len)
        assert 0 <= startp < T - feat_len
        return startp + j
    if T == feat_len:
        return j
    if padding == "zero":                               # :513-517
        return np.where(j < T, j, SRC_ZERO)
    if padding == "repeat":                             # :519-522
        return j % T
    if padding == "silence":                            # :524-528
        npad = feat_len - T
        return np.where(j < npad, SRC_SILENCE, j - npad)
    raise ValueError("Padding should be zero or repeat!")


def silence_vector(**kw):
    """LFCC frame 0 of 3200 zeros (dataset.py:13-16)."""
    return lfcc(np.zeros((1, 3200)), **kw)[0, 0]


def apply_frame_map(feat, fmap, silence=None):
    """feat (B,T,D) -> (B,len(fmap),D) following frame_index_map."""
    B, T, D = feat.shape
    out = np.zeros((B, len(fmap), D), dtype=feat.dtype)
    for j, s in enumerate(fmap):
        if s >= 0:
            out[:, j] = feat[:, s]
        elif s == SRC_SILENCE:
            out[:, j] = silence
    return out
