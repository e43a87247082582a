"""Waveform sources for the fused path.  The reference trains from pre-extracted LFCC `.pt` files
(dataset.py:18-85) produced offline by preprocess.py:232-245; here raw waves go straight to the device
and LFCC / crop / pad happen inside the step, so a source only has to deliver
(waves (B, Lmax) float32, lengths (B,), labels (B,), names, crop starts (B,)) batches.

  SyntheticWaves : seeded white-noise utterances (the benchmark / smoke configuration)
  WaveFolder     : a folder of .wav (PCM16, stdlib `wave`) / .npy files + a protocol text file with lines
                   `utt_id label` or ASVspoof-style `spk utt_id - attack label` (label: bonafide / spoof)
Crop policy for utterances longer than feat_len frames follows dataset.py:66-69: start ~ np.random.randint.
"""
import os
import wave

import numpy as np
import torch

HOP = 160


def _crop_starts(lengths, feat_len, rng):
    frames = 1 + np.asarray(lengths) // HOP
    return np.array([rng.randint(int(t) - feat_len) if t > feat_len else 0 for t in frames], dtype=np.int32)


class SyntheticWaves:
    def __init__(self, n_utts, length=64000, seed=0, feat_len=750):
        self.n, self.length, self.seed, self.feat_len = int(n_utts), int(length), int(seed), feat_len

    def __len__(self):
        return self.n

    def batch(self, indices):
        idx = np.asarray(indices, dtype=np.int64)
        waves = torch.empty(len(idx), self.length)
        labels = torch.empty(len(idx), dtype=torch.long)
        for j, i in enumerate(idx):
            g = torch.Generator().manual_seed(self.seed * 1000003 + int(i))
            labels[j] = int(i) & 1                               # alternate bonafide (0) / spoof (1)
            # the two classes differ in spectral tilt so that training has something to learn
            w = torch.randn(self.length, generator=g)
            if labels[j] == 1:
                w[1:] = 0.7 * w[1:] + 0.3 * w[:-1]
            waves[j] = 0.1 * w
        names = ["SYN_%07d" % int(i) for i in idx]
        return waves, torch.full((len(idx),), self.length, dtype=torch.int32), labels, names, None


class WaveFolder:
    def __init__(self, folder, protocol, feat_len=750, seed=0):
        self.folder, self.feat_len = folder, feat_len
        self.items = []
        with open(protocol) as f:
            for line in f:
                p = line.split()
                if not p:
                    continue
                utt, lab = (p[1], p[-1]) if len(p) >= 4 else (p[0], p[-1] if len(p) > 1 else "bonafide")
                self.items.append((utt, 0 if lab == "bonafide" else 1))
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.items)

    def _read(self, utt):
        base = os.path.join(self.folder, utt)
        if os.path.exists(base + ".npy"):
            a = np.load(base + ".npy")
            return (a.astype(np.float32) / 32768.0) if a.dtype == np.int16 else a.astype(np.float32)
        with wave.open(base + ".wav", "rb") as w:
            assert w.getsampwidth() == 2, "PCM16 only"
            a = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
            if w.getnchannels() > 1:
                a = a.reshape(-1, w.getnchannels())[:, 0]
        return a.astype(np.float32) / 32768.0

    def batch(self, indices):
        arrs = [self._read(self.items[i][0]) for i in indices]
        lens = np.array([len(a) for a in arrs], dtype=np.int32)
        waves = torch.zeros(len(arrs), int(lens.max()))
        for j, a in enumerate(arrs):
            waves[j, :len(a)] = torch.from_numpy(a)
        labels = torch.tensor([self.items[i][1] for i in indices], dtype=torch.long)
        start = torch.from_numpy(_crop_starts(lens, self.feat_len, self.rng))
        return waves, torch.from_numpy(lens), labels, [self.items[i][0] for i in indices], start
