"""CPU study behind DESIGN.md section 4.2 (d): how far can the tensor-core LFCC drop split terms?

Emulates the folded-DFT arithmetic of csrc/lfcc_tc.cu with different operand splits and reports the worst cepstral
deviation |a - b| / (|b| + 1) against the float64 oracle (the parity bar is 1e-4) on white noise, the golden edge
signals and a speech-like signal (strong low-frequency harmonics, weak high band).
    python scripts/lfcc_split_study.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from asvspoof2021_air_b200 import lfcc_tables as lt  # noqa: E402
from oracle import lfcc_oracle as lo  # noqa: E402


def rnd(x, kind):
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return (t.to(torch.bfloat16) if kind == "bf16" else t.to(torch.float16)).float().numpy().astype(np.float64)


def cepstra(wave, scheme):
    kind, terms = scheme
    fb = lt.linear_filterbank(512, 16000, 20).numpy().astype(np.float64)
    dct = lt.dct_ortho_matrix(20).numpy().astype(np.float64)
    win = torch.hamming_window(320).numpy().astype(np.float32)
    k = np.arange(1, 257, dtype=np.float64)[:, None]
    m = np.arange(160, dtype=np.float64)[None, :]
    C, S = np.cos(2 * np.pi * k * m / 512), np.sin(2 * np.pi * k * m / 512)
    Chi, Shi = rnd(C, kind), rnd(S, kind)
    Clo, Slo = rnd(C - Chi, kind), rnd(S - Shi, kind)
    wave = np.asarray(wave, dtype=np.float32)
    B, L = wave.shape
    T = 1 + L // 160
    y = wave.copy()
    y[:, 1:] = wave[:, 1:] - np.float32(0.97) * wave[:, :-1]
    ypad = np.zeros((B, 160 * (T + 2)), np.float32)
    ypad[:, 160:160 + L] = y
    kb = np.arange(1, 257)
    c160, s160 = np.cos(5 * np.pi * (kb % 16) / 8), np.sin(5 * np.pi * (kb % 16) / 8)
    mm = np.arange(1, 160)
    out = np.zeros((B, T, 20))
    for t in range(T):
        a = (ypad[:, 160 * t:160 * t + 320] * win).astype(np.float64)
        e = np.zeros((B, 160)); o = np.zeros((B, 160))
        e[:, 0] = a[:, 160]
        e[:, 1:] = a[:, 160 + mm] + a[:, 160 - mm]
        o[:, 1:] = a[:, 160 + mm] - a[:, 160 - mm]
        ehi, ohi = rnd(e, kind), rnd(o, kind)
        elo, olo = rnd(e - ehi, kind), rnd(o - ohi, kind)
        re, im = ehi @ Chi.T, ohi @ Shi.T
        if terms >= 2:
            re, im = re + elo @ Chi.T, im + olo @ Shi.T
        if terms >= 3:
            re, im = re + ehi @ Clo.T, im + ohi @ Slo.T
        if terms >= 4:
            re, im = re + elo @ Clo.T, im + olo @ Slo.T
        re = re.astype(np.float32) + a[:, :1] * c160
        im = im.astype(np.float32) - a[:, :1] * s160
        P = (re * re + im * im)[:, :255]
        fbe = np.log10(P @ fb[1:256] + 1.1920929e-07)
        out[:, t] = fbe @ dct.T
    return out


def signals(L=16000):
    rng = np.random.RandomState(0)
    n = np.arange(L)
    sq = np.where((n // 40) % 2 == 0, 1.0, -1.0)
    imp = np.zeros(L); imp[0] = 1.0
    speech = sum(0.3 / h * np.sin(2 * np.pi * 140 * h * n / 16000) for h in range(1, 9)) + 3e-4 * rng.randn(L)
    return {"white noise 0.1": 0.1 * rng.randn(L), "square wave +-1": sq, "impulse": imp, "speech-like (harmonics + -60 dB noise)": speech}


if __name__ == "__main__":
    schemes = {"bf16 4 terms (+x_lo*w_lo)": ("bf16", 4), "bf16 3 terms (kernel)": ("bf16", 3), "bf16 2 terms (no x_hi*w_lo)": ("bf16", 2),
               "fp16 2 terms (no x_hi*w_lo)": ("fp16", 2), "fp16 1 term": ("fp16", 1)}
    sig = signals()
    print("%-42s" % "worst |a-b|/(|b|+1) of the 20 cepstra" + "".join("%30s" % s for s in schemes))
    for name, w in sig.items():
        ref = lo.lfcc(w[None].astype(np.float32))[:, :, :20]
        row = []
        for s in schemes.values():
            c = cepstra(w[None], s)
            row.append(np.max(np.abs(c - ref) / (np.abs(ref) + 1)))
        print("%-42s" % name + "".join("%30.2e" % v for v in row))
