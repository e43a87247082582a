"""Audio ingest through the C ABI (csrc/audio_io.cpp): FLAC / WAV files -> float32 mono samples, one file or a
whole batch at a time on host threads, straight into pinned memory.

Replaces the reference's per-item `librosa.load(path, sr=16000)` / `soundfile.read` (raw_dataset.py:20-28,61-66):
integer PCM scaled by 2^-(bits-1), channels averaged to mono, no resampling (a file that is not at `expect_sr`
raises).  ctypes releases the GIL for the duration of a call, so a Python prefetch thread overlaps decoding with
the training step."""
import ctypes
import os

import numpy as np
import torch

from . import _lib

ERRORS = {-1: "bad argument", -2: "unsupported stream (sample size / channel count changes, exotic WAV coding)",
          -3: "cannot read the file", -4: "not a valid FLAC / WAV stream", -5: "checksum mismatch (CRC-8 / CRC-16 / MD5)", -6: "out of memory while decoding"}
VERIFY_MD5 = 1


class AudioError(_lib.AirError):
    pass


def _raise(status, path):
    raise AudioError("%s: %s (status %d)" % (path, ERRORS.get(status, "error"), status))


def info(path):
    """(sample_rate, channels, bits, frames) of a FLAC / WAV file."""
    sr, ch, bits, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
    st = _lib.lib().air_audio_info(os.fsencode(path), ctypes.byref(sr), ctypes.byref(ch), ctypes.byref(bits), ctypes.byref(n))
    if st != 0:
        _raise(st, path)
    return sr.value, ch.value, bits.value, n.value


def decode(path, verify=False):
    """-> (float32 mono samples, sample_rate)."""
    n = info(path)[3]
    out = np.empty(max(n, 1), dtype=np.float32)
    frames, sr = ctypes.c_longlong(), ctypes.c_int()
    st = _lib.lib().air_audio_decode_f32(os.fsencode(path), out.ctypes.data_as(ctypes.c_void_p), _lib.LL(n),
                                         ctypes.byref(frames), ctypes.byref(sr), VERIFY_MD5 if verify else 0)
    if st != 0:
        _raise(st, path)
    return out[:n], sr.value


def decode_int(path, verify=False):
    """-> (int32 samples (frames, channels) exactly as stored, bits, sample_rate)."""
    sr0, ch0, bits0, n = info(path)
    out = np.empty(max(n * ch0, 1), dtype=np.int32)
    frames, ch, bits, sr = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    st = _lib.lib().air_audio_decode_i32(os.fsencode(path), out.ctypes.data_as(ctypes.c_void_p), _lib.LL(out.size),
                                         ctypes.byref(frames), ctypes.byref(ch), ctypes.byref(bits), ctypes.byref(sr),
                                         VERIFY_MD5 if verify else 0)
    if st != 0:
        _raise(st, path)
    return out[:n * ch0].reshape(n, ch0), bits.value, sr.value


def decode_batch(paths, max_len, out=None, threads=0, verify=False, expect_sr=16000):
    """Decode len(paths) files into the rows of `out` ((>= n, max_len) float32, pinned when CUDA is present; made if
    None): row i = the first min(length_i, max_len) samples, zero-padded.  -> (out[:n], lengths int32 (n,))."""
    n = len(paths)
    if out is None:
        out = torch.empty(n, max_len, dtype=torch.float32, pin_memory=torch.cuda.is_available())
    assert out.dtype == torch.float32 and out.dim() == 2 and out.shape[0] >= n and out.shape[1] >= max_len
    assert out.stride(1) == 1
    arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
    lengths = np.zeros(n, dtype=np.int32)
    rates = np.zeros(n, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    st = _lib.lib().air_audio_decode_batch_f32(arr, n, ctypes.c_void_p(out.data_ptr()), _lib.LL(out.stride(0)),
                                               ip(lengths), ip(rates), ip(status), int(threads),
                                               VERIFY_MD5 if verify else 0)
    if st != 0:
        bad = int(np.nonzero(status)[0][0]) if status.any() else 0
        _raise(int(status[bad]) if status.any() else st, paths[bad] if n else "<batch>")
    if expect_sr and n and (rates != expect_sr).any():
        bad = int(np.nonzero(rates != expect_sr)[0][0])
        raise AudioError("%s is sampled at %d Hz, the path expects %d Hz (no resampler)" % (paths[bad], rates[bad], expect_sr))
    if out.shape[1] > max_len:
        out[:n, max_len:].zero_()
    return out[:n], torch.from_numpy(lengths)
