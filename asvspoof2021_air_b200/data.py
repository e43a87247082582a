"""Waveform sources for the fused path.  The reference trains from pre-extracted LFCC `.pt` files
(dataset.py:18-85) produced offline by preprocess.py:232-245; here raw waves go straight to the device
and LFCC / crop / pad happen inside the step, so a source only has to deliver
(waves (B, Lmax) float32, lengths (B,), labels (B,), names, crop starts (B,)) batches.

  SyntheticWaves : seeded white-noise utterances (the benchmark / smoke configuration)
  WaveFolder     : a folder of .flac / .wav / .npy files + a protocol text file with lines
                   `utt_id label` or ASVspoof-style `spk utt_id - attack label` (label: bonafide / spoof);
                   FLAC and WAV are decoded by the native batch decoder (csrc/audio_io.cpp, host threads, straight
                   into pinned rows) -- the ASVspoof corpora ship as 16 kHz FLAC (raw_dataset.py:20-28,61-66)
  Prefetcher     : decodes / collates batch i+1.. on a host thread and copies it to the device on a copy stream while
                   step i runs (the DataLoader(num_workers) + .to(device) of main_train.py:226-242,338-348)
Crop policy for utterances longer than feat_len frames follows dataset.py:66-69: start ~ np.random.randint.
"""
import os
import queue
import threading

import numpy as np
import torch

from . import audio_io

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
    EXTS = (".flac", ".wav", ".npy")

    def __init__(self, folder, protocol, feat_len=750, seed=0, threads=0, verify=False):
        self.folder, self.feat_len, self.threads, self.verify = folder, feat_len, threads, verify
        self.items = []
        with open(protocol) as f:
            for line in f:
                p = line.split()
                if not p:
                    continue
                utt, lab = (p[1], p[-1]) if len(p) >= 4 else (p[0], p[-1] if len(p) > 1 else "bonafide")
                self.items.append((utt, 0 if lab == "bonafide" else 1))
        present = set(os.listdir(folder))
        self.paths = {}
        for utt, _ in self.items:
            ext = next((e for e in self.EXTS if utt + e in present), None)
            if ext is None:
                raise FileNotFoundError("%s: no %s file for utterance %s" % (folder, " / ".join(self.EXTS), utt))
            self.paths[utt] = os.path.join(folder, utt + ext)
        self._frames = {}
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.items)

    def frames(self, utt):
        """Length in samples, from the container header (cached)."""
        n = self._frames.get(utt)
        if n is None:
            p = self.paths[utt]
            n = int(np.load(p, mmap_mode="r").shape[0]) if p.endswith(".npy") else audio_io.info(p)[3]
            self._frames[utt] = n
        return n

    def batch(self, indices, pinned=None):
        utts = [self.items[i][0] for i in indices]
        lens = np.array([self.frames(u) for u in utts], dtype=np.int32)
        pinned = torch.cuda.is_available() if pinned is None else pinned
        waves = torch.zeros(len(utts), int(lens.max()), pin_memory=pinned)
        coded = [j for j, u in enumerate(utts) if not self.paths[u].endswith(".npy")]
        if coded:                                                       # one native call, host threads, GIL released
            rows, got = audio_io.decode_batch([self.paths[utts[j]] for j in coded], waves.shape[1],
                                              out=waves if len(coded) == len(utts) else None, threads=self.threads,
                                              verify=self.verify)
            assert got.tolist() == lens[coded].tolist()
            if len(coded) != len(utts):
                waves[coded] = rows
        for j, u in enumerate(utts):
            if self.paths[u].endswith(".npy"):
                a = np.load(self.paths[u])
                a = (a.astype(np.float32) / 32768.0) if a.dtype == np.int16 else a.astype(np.float32)
                waves[j, :len(a)] = torch.from_numpy(a)
        labels = torch.tensor([self.items[i][1] for i in indices], dtype=torch.long)
        start = torch.from_numpy(_crop_starts(lens, self.feat_len, self.rng))
        return waves, torch.from_numpy(lens), labels, utts, start


class Batch(tuple):
    """(waves, lengths, labels, names, start) plus `.labels_host`; `lengths` is None when no row is shorter than the
    batch matrix (nothing for the kernels to mask)."""
    labels_host = None


class Prefetcher:
    """Iterate `source.batch(idx)` over `index_batches` with up to `depth` batches decoded ahead on a host thread
    and, when `device` is a CUDA device, already on their way to it on a copy stream.  Yields Batch tuples
    (waves, lengths, labels, names, start) -- tensors on `device` -- in order.  An exception in the worker is
    re-raised in the consumer."""

    def __init__(self, source, index_batches, depth=2, device=None):
        self.source, self.batches, self.device = source, list(index_batches), device
        self.q = queue.Queue(maxsize=max(1, depth))
        self.copy_stream = torch.cuda.Stream(device) if device is not None else None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        try:
            for idx in self.batches:
                waves, lens, labels, names, start = self.source.batch(idx)
                if int(lens.min()) == waves.shape[1]:
                    lens = None
                host = (waves, lens, labels, start)
                ev = None
                if self.device is not None:
                    with torch.cuda.stream(self.copy_stream):
                        moved = [t.to(self.device, non_blocking=True) if t is not None else None for t in host]
                        ev = torch.cuda.Event()
                        ev.record(self.copy_stream)
                else:
                    moved = list(host)
                b = Batch((moved[0], moved[1], moved[2], names, moved[3]))
                b.labels_host = labels
                self.q.put((b, ev, host))                              # `host` keeps the pinned source alive
            self.q.put(None)
        except BaseException as e:                                      # surfaced by __iter__
            self.q.put(e)

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            if isinstance(item, BaseException):
                raise item
            b, ev, _host = item
            if ev is not None:
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                for t in (b[0], b[1], b[2], b[4]):
                    if t is not None:
                        t.record_stream(cur)
            yield b
