"""Waveform sources for the fused path.  The reference trains from pre-extracted LFCC `.pt` files
(dataset.py:18-85) produced offline by preprocess.py:232-245; here raw waves go straight to the device
and LFCC / crop / pad happen inside the step, so a source only has to deliver
(waves (B, Lmax) float32, lengths (B,), labels (B,), names, crop starts (B,)) batches.

  SyntheticWaves : seeded white-noise utterances (the benchmark / smoke configuration)
  WaveFolder     : a folder of .flac / .wav / .npy files + a protocol text file with lines
                   `utt_id label` or ASVspoof-style `spk utt_id - attack label` (label: bonafide / spoof);
                   FLAC and WAV are decoded by the native batch decoder (csrc/audio_io.cpp, host threads, straight
                   into pinned rows) -- the ASVspoof corpora ship as 16 kHz FLAC (raw_dataset.py:20-28,61-66)
  AugWaveFolder  : originals + channel-augmented copies with channel / device class labels (--ADV_AUG)
  PackedWaves    : a folder decoded once by pack_folder() into one memory-mapped int16 file; batches are gathered and
                   converted by a native threaded copy (full-speed epochs without re-decoding FLAC)
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
    PIN = None                       # None: pin batch buffers when a CUDA device is present

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
        if pinned is None:
            pinned = torch.cuda.is_available() if self.PIN is None else self.PIN
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


def pack_folder(source, prefix, batch=256):
    """Decode a WaveFolder ONCE into a packed int16 corpus: `<prefix>.i16` (all utterances back to back) and
    `<prefix>.json` (names, labels, offsets, lengths, extra per-item classes).  The native decoder delivers about 700
    four-second FLAC files per second and core while one GPU consumes ~11 000 utterances per second, so every epoch
    after the first should read this file (PackedWaves) instead of the FLAC folder.  16-bit sources only: the float
    samples are mapped back to the exact integers they came from."""
    import json
    offsets, lengths, pos = [], [], 0
    with open(prefix + ".i16", "wb") as f:
        for lo in range(0, len(source), batch):
            idx = list(range(lo, min(lo + batch, len(source))))
            item = source.batch(idx, pinned=False)
            waves, lens = item[0], item[1]
            q = torch.round(waves * 32768.0)
            if float((q - waves * 32768.0).abs().max()) > 1e-3 or float(q.abs().max()) > 32768:
                raise ValueError("pack_folder: the source is not 16-bit PCM (samples are not multiples of 2^-15)")
            q = q.clamp(-32768, 32767).to(torch.int16).numpy()
            for j, n in enumerate(lens.tolist()):
                f.write(q[j, :n].tobytes())
                offsets.append(pos)
                lengths.append(n)
                pos += n
    meta = {"names": [u for u, _ in source.items], "labels": [int(l) for _, l in source.items], "offsets": offsets,
            "lengths": lengths, "sample_rate": 16000, "samples": pos,
            "classes": [list(c) for c in getattr(source, "classes", [])], "kind": getattr(source, "kind", None),
            "n_ori": getattr(source, "n_ori", len(source))}
    with open(prefix + ".json", "w") as f:
        json.dump(meta, f)
    return meta


class PackedWaves:
    """A corpus written by pack_folder(): memory-mapped int16 samples, rows gathered and converted to float32 by the
    native threaded gather (csrc/audio_io.cpp air_audio_gather_i16_f32) straight into pinned memory.  Same batch()
    contract as WaveFolder / AugWaveFolder (a sixth element with the channel classes when the source had them)."""
    PIN = None

    def __init__(self, prefix, feat_len=750, seed=0, threads=0):
        import json
        with open(prefix + ".json") as f:
            m = json.load(f)
        self.names, self.labels = m["names"], np.asarray(m["labels"], dtype=np.int64)
        self.offsets = np.asarray(m["offsets"], dtype=np.int64)
        self.lengths = np.asarray(m["lengths"], dtype=np.int32)
        self.classes = np.asarray(m["classes"], dtype=np.int64) if m.get("classes") else None
        self.kind, self.n_ori = m.get("kind"), m.get("n_ori", len(self.names))
        if self.kind:
            self.channel_names, self.device_names = channel_tables(self.kind)
        self.blob = np.memmap(prefix + ".i16", dtype=np.int16, mode="r")
        if self.blob.shape[0] != m["samples"]:
            raise ValueError("%s.i16 holds %d samples, the index announces %d" % (prefix, self.blob.shape[0], m["samples"]))
        if len(self.offsets) != len(self.names) or len(self.lengths) != len(self.names):
            raise ValueError("%s.json: %d names, %d offsets, %d lengths" % (prefix, len(self.names), len(self.offsets), len(self.lengths)))
        bad = (self.offsets < 0) | (self.lengths < 0) | (self.offsets + self.lengths > self.blob.shape[0])
        if bad.any():                                        # a stale / corrupt index must not read outside the mapping
            raise ValueError("%s.json: item %d (offset %d, length %d) lies outside the %d samples of the corpus"
                             % (prefix, int(np.argmax(bad)), int(self.offsets[np.argmax(bad)]), int(self.lengths[np.argmax(bad)]),
                                self.blob.shape[0]))
        self.items = list(zip(self.names, self.labels.tolist()))
        self.feat_len, self.threads = feat_len, threads
        self.rng = np.random.RandomState(seed)

    def __len__(self):
        return len(self.names)

    def batch(self, indices, pinned=None):
        import ctypes
        from . import _lib
        idx = np.asarray(indices, dtype=np.int64)
        lens = np.ascontiguousarray(self.lengths[idx])
        offs = np.ascontiguousarray(self.offsets[idx])
        if pinned is None:
            pinned = torch.cuda.is_available() if self.PIN is None else self.PIN
        waves = torch.empty(len(idx), int(lens.max()), pin_memory=pinned)
        st = _lib.lib().air_audio_gather_i16_f32(ctypes.c_void_p(self.blob.ctypes.data), _lib.LL(self.blob.shape[0]),
                                                 offs.ctypes.data_as(ctypes.c_void_p),
                                                 lens.ctypes.data_as(ctypes.c_void_p), len(idx), ctypes.c_void_p(waves.data_ptr()),
                                                 _lib.LL(waves.stride(0)), int(self.threads))
        if st != 0:
            raise audio_io.AudioError("air_audio_gather_i16_f32 failed: status %d" % st)
        labels = torch.from_numpy(self.labels[idx])
        start = torch.from_numpy(_crop_starts(lens, self.feat_len, self.rng))
        out = (waves, torch.from_numpy(lens), labels, [self.names[i] for i in idx], start)
        if self.classes is not None and self.classes.size:
            ch = torch.from_numpy(self.classes[idx])
            out += (ch[:, 0] if ch.shape[1] == 1 else ch,)
        return out


def channel_tables(kind):
    """(channel names, device names or None) of an augmented set -- `kind` in LA / DF / LAPA / DFPA; the list position is
    the class index (dataset.py:122-138,207-228,345-346,407-413; stored in channel_tables.json)."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "channel_tables.json")) as f:
        t = json.load(f)[kind]
    return t["channel"], t.get("devices")


class AugWaveFolder(WaveFolder):
    """Original + channel-augmented waves of the --ADV_AUG branch (raw_dataset.py:149-300, dataset.py:105-183,190-277).

    `folder` holds the original utterances `<utt>.{flac,wav,npy}` named by the protocol; `aug_folder` holds augmented
    copies `<utt>_<channel>.<ext>` (LA / DF) or `<utt>_<channel>_<device>.<ext>` (LAPA / DFPA).  Items 0 .. n_ori-1 are
    the originals (channel `no_channel`, device ``""``), the rest the augmented files in sorted order -- the index
    layout main_train.py:226-231 samples its two half-batches from.  batch() returns a sixth element: the int64 class
    indices, (B,) or (B, 2) = [channel, device]."""

    def __init__(self, folder, aug_folder, protocol, kind, feat_len=750, seed=0, threads=0, verify=False):
        super().__init__(folder, protocol, feat_len, seed, threads, verify)
        self.kind = kind
        self.channel_names, self.device_names = channel_tables(kind)
        cidx = {n: i for i, n in enumerate(self.channel_names)}
        didx = {n: i for i, n in enumerate(self.device_names)} if self.device_names else None
        label_of = dict(self.items)
        self.n_ori = len(self.items)
        self.classes = [(cidx["no_channel"], didx[""]) if didx else (cidx["no_channel"],) for _ in self.items]
        for fn in sorted(os.listdir(aug_folder)):
            stem, ext = os.path.splitext(fn)
            if ext not in self.EXTS:
                continue
            parts = stem.split("_")
            ntail = 2 if didx else 1
            utt, tail = "_".join(parts[:-ntail]), parts[-ntail:]
            if utt not in label_of:
                raise KeyError("%s: %s is not in the protocol" % (aug_folder, utt))
            if tail[0] not in cidx or (didx and tail[1] not in didx):
                raise KeyError("%s: unknown channel / device in %s" % (aug_folder, fn))
            name = stem
            self.items.append((name, label_of[utt]))
            self.paths[name] = os.path.join(aug_folder, fn)
            self.classes.append((cidx[tail[0]], didx[tail[1]]) if didx else (cidx[tail[0]],))

    def batch(self, indices, pinned=None):
        out = super().batch(indices, pinned)
        ch = torch.tensor([self.classes[i] for i in indices], dtype=torch.long)
        return out + (ch[:, 0] if ch.shape[1] == 1 else ch,)


def half_batches(n_ori, n_total, batch_size, ratio, steps, rng):
    """Index lists of main_train.py:226-233,309-325: every step int(batch_size * ratio) originals and the rest augmented
    utterances, each drawn without replacement from its own reshuffled-when-exhausted permutation."""
    k_ori = int(batch_size * ratio)
    k_aug = batch_size - k_ori
    pools = [list(rng.permutation(n_ori)), list(n_ori + rng.permutation(n_total - n_ori))]

    def take(which, k):
        out = []
        while len(out) < k:
            if not pools[which]:
                pools[which] = list(rng.permutation(n_ori)) if which == 0 else list(n_ori + rng.permutation(n_total - n_ori))
            out.append(int(pools[which].pop()))
        return out
    return [take(0, k_ori) + (take(1, k_aug) if n_total > n_ori else []) for _ in range(steps)]


class Batch(tuple):
    """(waves, lengths, labels, names, start) plus `.labels_host` and `.channels` (device tensor, sources that carry
    channel labels only); `lengths` is None when no row is shorter than the batch matrix (nothing for the kernels to
    mask)."""
    labels_host = None
    channels = None


class Prefetcher:
    """Iterate `source.batch(idx)` over `index_batches` with up to `depth` batches decoded ahead on a host thread
    and, when `device` is a CUDA device, already on their way to it on a copy stream.  Yields Batch tuples
    (waves, lengths, labels, names, start) -- tensors on `device` -- in order.  An exception in the worker is
    re-raised in the consumer."""

    def __init__(self, source, index_batches, depth=2, device=None):
        if device is not None and torch.device(device).type != "cuda":
            device = None                                               # host consumer: no copy stream
        self.source, self.batches, self.device = source, list(index_batches), device
        self.q = queue.Queue(maxsize=max(1, depth))
        self.copy_stream = torch.cuda.Stream(device) if device is not None else None
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        try:
            for idx in self.batches:
                item = self.source.batch(idx)
                waves, lens, labels, names, start = item[:5]
                channels = item[5] if len(item) > 5 else None
                if int(lens.min()) == waves.shape[1]:
                    lens = None
                host = (waves, lens, labels, start, channels)
                ev = None
                if self.device is not None:
                    with torch.cuda.stream(self.copy_stream):
                        moved = [t.to(self.device, non_blocking=True) if t is not None else None for t in host]
                        ev = torch.cuda.Event()
                        ev.record(self.copy_stream)
                else:
                    moved = list(host)
                b = Batch((moved[0], moved[1], moved[2], names, moved[3]))
                b.labels_host, b.channels = labels, moved[4]
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
                for t in (b[0], b[1], b[2], b[4], b.channels):
                    if t is not None:
                        t.record_stream(cur)
            yield b


def _main(argv=None):
    """python -m asvspoof2021_air_b200.data pack --wave_dir DIR --protocol FILE --out PREFIX
                                            [--aug_wave_dir DIR --kind LA|DF|LAPA|DFPA] [--verify]"""
    import argparse
    ap = argparse.ArgumentParser(prog="python -m asvspoof2021_air_b200.data")
    ap.add_argument("command", choices=["pack"])
    ap.add_argument("--wave_dir", required=True)
    ap.add_argument("--protocol", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--aug_wave_dir", default=None)
    ap.add_argument("--kind", default=None, choices=[None, "LA", "DF", "LAPA", "DFPA"])
    ap.add_argument("--verify", action="store_true", help="check the MD5 signature of every FLAC file")
    a = ap.parse_args(argv)
    if a.aug_wave_dir:
        src = AugWaveFolder(a.wave_dir, a.aug_wave_dir, a.protocol, a.kind or "LA", verify=a.verify)
    else:
        src = WaveFolder(a.wave_dir, a.protocol, verify=a.verify)
    m = pack_folder(src, a.out)
    print("packed %d utterances, %.1f MB -> %s.i16" % (len(m["names"]), 2e-6 * m["samples"], a.out))


if __name__ == "__main__":
    _main()
