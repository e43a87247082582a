"""CPU tests of the audio ingest (csrc/audio_io.cpp): the FLAC decoder against (a) the worked examples of the format
specification (RFC 9639 appendix D.1 and D.3 -- streams produced by the reference libFLAC encoder, each carrying the
MD5 of its audio, which the decoder verifies) and (b) streams from tests/flac_writer.py that exercise every subframe
type, predictor order, Rice layout, channel decorrelation and header code; the WAV reader against scipy."""
import os
import struct

import numpy as np
import pytest
import scipy.io.wavfile
import torch

import flac_writer as fw
from asvspoof2021_air_b200 import audio_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RFC_D1 = ("664c614380000022100010000000 0f00000f0ac442f0000000013e84b41807dc690307586a3dad1a2e0f"
          "fff869180000bf0358fd03128baa9a")
RFC_D3 = ("664c6143800000221000100000001f00001f07d0007000000018f8f9e396f5cbcfc6dc807f9977906b32"
          "fff868020017e944004f6f313d1047d227cb6d090831452bdc2822228057a3")


def _write(tmp_path, name, data):
    p = str(tmp_path / name)
    with open(p, "wb") as f:
        f.write(data)
    return p


def _speechlike(n, rng, bits=16, channels=1):
    """AR(2) resonance + noise, scaled to ~1/4 full scale: predictable enough for LPC / fixed predictors."""
    x = np.zeros((n + 2, channels))
    e = rng.randn(n + 2, channels)
    for i in range(2, n + 2):
        x[i] = 1.6 * x[i - 1] - 0.8 * x[i - 2] + e[i]
    x = x[2:] / np.abs(x).max() * (1 << (bits - 3))
    return np.round(x).astype(np.int64)


def test_rfc9639_worked_examples_decode_and_pass_their_md5(tmp_path):
    p1 = _write(tmp_path, "d1.flac", bytes.fromhex(RFC_D1.replace(" ", "")))
    x, bits, sr = audio_io.decode_int(p1, verify=True)
    assert (bits, sr) == (16, 44100) and x.tolist() == [[25588, 10416]]
    y, _ = audio_io.decode(p1)
    assert y.tolist() == [np.float32((25588 / 32768 + 10416 / 32768) / 2)]                # mono = mean of channels
    p3 = _write(tmp_path, "d3.flac", bytes.fromhex(RFC_D3))
    x, bits, sr = audio_io.decode_int(p3, verify=True)
    assert (bits, sr) == (8, 32000) and audio_io.info(p3) == (32000, 1, 8, 24)
    assert x[:, 0].tolist() == [0, 79, 111, 78, 8, -61, -90, -68, -13, 42, 67, 53, 13, -27, -46, -38, -12, 14, 24, 19,
                                6, -4, -5, 0]
    # the checks are live: one flipped bit anywhere is caught by the CRC-8 / CRC-16 / MD5 it falls under
    raw = bytearray(bytes.fromhex(RFC_D3))
    for pos in (30, 44, 47, 60):                       # MD5 field, frame header, LPC subframe, residual
        bad = bytearray(raw)
        bad[pos] ^= 0x10
        with pytest.raises(audio_io.AudioError, match="status -[45]"):
            audio_io.decode_int(_write(tmp_path, "bad.flac", bytes(bad)), verify=True)


SUBS = {
    "fixed0": fw.Sub("fixed", 0), "fixed1": fw.Sub("fixed", 1, porder=2), "fixed2": fw.Sub("fixed", 2, porder=3),
    "fixed3": fw.Sub("fixed", 3, method=1), "fixed4": fw.Sub("fixed", 4, porder=4, method=1),
    "lpc2": fw.Sub("lpc", 2, coefs=[1638, -819], precision=12, shift=10, porder=1),
    "lpc8": fw.Sub("lpc", 8, coefs=[9000, -7000, 3000, -900, 200, -50, 10, -2], precision=15, shift=13, porder=2),
    "lpc32": fw.Sub("lpc", 32, coefs=[1200, -500] + [(-1) ** k * (30 - k) for k in range(30)], precision=12, shift=10),
    "verbatim": fw.Sub("verbatim"),
    "escape": fw.Sub("fixed", 2, porder=2, escape_parts=(1, 3)),
    "rice0": fw.Sub("fixed", 2, rice=0), "rice14": fw.Sub("fixed", 1, rice=14), "rice30": fw.Sub("fixed", 1, rice=30, method=1),
}


@pytest.mark.parametrize("name", sorted(SUBS))
def test_every_subframe_kind_round_trips_mono_16bit(tmp_path, name):
    rng = np.random.RandomState(len(name))
    x = _speechlike(4096 + 1152 + 100, rng)
    frames = [fw.FrameSpec(4096, [SUBS[name]]), fw.FrameSpec(1152, [SUBS[name]]), fw.FrameSpec(100, [fw.Sub("fixed", 2)])]
    if name in ("fixed4", "lpc2"):
        frames[1] = fw.FrameSpec(1152, [SUBS[name]], block_code=7)      # explicit 16-bit block size
    p = _write(tmp_path, name + ".flac", fw.encode_flac(x, 16, 16000, frames))
    y, bits, sr = audio_io.decode_int(p, verify=True)
    assert (bits, sr) == (16, 16000) and np.array_equal(y, x)
    f, sr = audio_io.decode(p, verify=True)
    assert f.dtype == np.float32 and np.array_equal(f, (x[:, 0] / 32768.0).astype(np.float32))


@pytest.mark.parametrize("assignment", ["independent", "left_side", "right_side", "mid_side"])
@pytest.mark.parametrize("bits", [8, 12, 16, 20, 24, 32])
def test_stereo_decorrelation_and_sample_sizes(tmp_path, assignment, bits):
    rng = np.random.RandomState(bits)
    x = _speechlike(700, rng, bits=bits, channels=2)
    x[:, 1] = x[:, 0] // 2 + rng.randint(-3, 4, 700)                    # correlated channels, odd sums for mid/side
    if bits == 32:
        x[5] = [(1 << 31) - 1, -(1 << 31)]                              # the side channel needs 33 bits
    subs = [fw.Sub("fixed", 2, porder=2, method=1), fw.Sub("lpc", 3, coefs=[7, -6, 2], precision=4, shift=2, method=1)]
    if bits == 32:                                                      # residuals must fit 32 bits: full-scale jumps go verbatim
        subs = [fw.Sub("verbatim"), fw.Sub("verbatim")]
    frames = [fw.FrameSpec(512, subs, assignment), fw.FrameSpec(188, subs, assignment)]
    p = _write(tmp_path, "st.flac", fw.encode_flac(x, bits, 48000, frames))
    y, b, sr = audio_io.decode_int(p, verify=True)
    assert (b, sr) == (bits, 48000) and np.array_equal(y, x)
    f, _ = audio_io.decode(p)
    scale = np.float32(1.0 / (1 << (bits - 1)))
    want = (x[:, 0].astype(np.float32) * scale + x[:, 1].astype(np.float32) * scale) / np.float32(2)
    assert np.array_equal(f, want)


def test_constant_wasted_bits_variable_blocking_and_header_codes(tmp_path):
    rng = np.random.RandomState(5)
    a = _speechlike(192, rng)
    silence = np.zeros((576, 1), np.int64)
    dc = np.full((256, 1), -1234, np.int64)
    coarse = _speechlike(300, rng) // 8 * 8                              # 3 wasted bits
    x = np.concatenate([a, silence, dc, coarse, a[:17]])
    frames = [fw.FrameSpec(192, [fw.Sub("fixed", 1)], variable=True, rate_code=0, size_code=0),
              fw.FrameSpec(576, [fw.Sub("constant")], variable=True, rate_code=12),
              fw.FrameSpec(256, [fw.Sub("constant")], variable=True, rate_code=13),
              fw.FrameSpec(300, [fw.Sub("lpc", 2, coefs=[1638, -819], precision=12, shift=10, wasted=3)], variable=True, rate_code=14),
              fw.FrameSpec(17, [fw.Sub("verbatim", wasted=0)], variable=True, block_code=6)]
    extra = [(4, b"\x00" * 40), (1, b"\x00" * 1000)]                    # a VORBIS_COMMENT-sized blob and PADDING
    data = fw.encode_flac(x, 16, 16000, frames, extra_blocks=extra)
    y, _, _ = audio_io.decode_int(_write(tmp_path, "v.flac", data), verify=True)
    assert np.array_equal(y, x)
    # a long stream: sample numbers beyond 2^31 in the coded number (7-byte form) do not matter to the decoder
    assert fw.utf8_number(0) == b"\x00" and fw.utf8_number(0x7ff) == b"\xdf\xbf" and len(fw.utf8_number(1 << 35)) == 7
    # unknown length + no MD5 (a streamed encode), ID3v2 tag in front, junk after the last frame of a known length
    data = fw.encode_flac(x, 16, 16000, frames, total=0, md5=False)
    assert np.array_equal(audio_io.decode_int(_write(tmp_path, "s.flac", data), verify=True)[0], x)
    id3 = b"ID3\x04\x00\x00" + bytes([0, 0, 1, 2]) + b"\x00" * 130
    data = fw.encode_flac(x, 16, 16000, frames, prefix=id3, suffix=b"TAG" + b"\x00" * 125)
    assert np.array_equal(audio_io.decode_int(_write(tmp_path, "t.flac", data), verify=True)[0], x)


def test_damaged_and_foreign_files_are_reported(tmp_path):
    rng = np.random.RandomState(9)
    x = _speechlike(1000, rng)
    good = fw.encode_flac(x, 16, 16000, [fw.FrameSpec(1000, [fw.Sub("fixed", 2, porder=1)])])
    with pytest.raises(audio_io.AudioError, match="status -3"):
        audio_io.decode(str(tmp_path / "missing.flac"))
    with pytest.raises(audio_io.AudioError, match="status -4"):
        audio_io.decode(_write(tmp_path, "t.flac", b"OggS" + good[4:]))
    with pytest.raises(audio_io.AudioError, match="status -4"):
        audio_io.decode(_write(tmp_path, "t.flac", good[:len(good) // 2]))       # truncated inside the frame
    with pytest.raises(audio_io.AudioError, match="status -4"):
        audio_io.decode(_write(tmp_path, "t.flac", good[:42]))                   # announces 1000 samples, has none
    bad = bytearray(good)
    bad[-20] ^= 1
    with pytest.raises(audio_io.AudioError, match="status -[45]"):
        audio_io.decode(_write(tmp_path, "t.flac", bytes(bad)))
    wrong_md5 = bytearray(good)
    wrong_md5[4 + 4 + 18] ^= 0xff
    p = _write(tmp_path, "t.flac", bytes(wrong_md5))
    assert np.array_equal(audio_io.decode_int(p, verify=False)[0], x)            # audio itself is intact
    with pytest.raises(audio_io.AudioError, match="status -5"):
        audio_io.decode_int(p, verify=True)


def test_wav_reader_against_scipy(tmp_path):
    rng = np.random.RandomState(3)
    for dtype, bits in ((np.int16, 16), (np.int32, 32), (np.uint8, 8), (np.float32, 32)):
        for ch in (1, 2):
            if dtype == np.float32:
                x = rng.uniform(-1, 1, (999, ch)).astype(np.float32)
            else:
                info = np.iinfo(dtype)
                x = rng.randint(info.min, info.max, (999, ch)).astype(dtype)
            p = str(tmp_path / "w.wav")
            scipy.io.wavfile.write(p, 16000, x[:, 0] if ch == 1 else x)
            f, sr = audio_io.decode(p)
            if dtype == np.float32:
                want = x
            elif dtype == np.uint8:
                want = (x.astype(np.float32) - 128) / np.float32(128)
            else:
                want = x.astype(np.float32) * np.float32(1.0 / (1 << (bits - 1)))
            want = want[:, 0] if ch == 1 else (want[:, 0] + want[:, 1]) / np.float32(2)
            assert sr == 16000 and np.array_equal(f, want.astype(np.float32)), (dtype, ch)
    # 24-bit PCM by hand (scipy cannot write it)
    v = rng.randint(-(1 << 23), 1 << 23, 50)
    body = b"".join(int(s & 0xffffff).to_bytes(3, "little") for s in v)
    hdr = b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 48000, 3, 24)
    p = _write(tmp_path, "w24.wav", hdr + b"LIST" + struct.pack("<I", 3) + b"abc\x00" + b"data" + struct.pack("<I", len(body)) + body)
    assert np.array_equal(audio_io.decode(p)[0], (v / float(1 << 23)).astype(np.float32))
    assert np.array_equal(audio_io.decode_int(p)[0][:, 0], v)


def test_batch_decode_into_padded_rows_on_threads(tmp_path):
    rng = np.random.RandomState(11)
    paths, waves = [], []
    for i in range(13):
        n = int(rng.randint(200, 3000))
        x = _speechlike(n, rng)
        p = str(tmp_path / ("u%02d" % i))
        if i % 3 == 2:
            scipy.io.wavfile.write(p + ".wav", 16000, x[:, 0].astype(np.int16))
            p += ".wav"
        else:
            blocks = [1024] * (n // 1024) + ([n % 1024] if n % 1024 else [])
            fw.write_flac(p + ".flac", x, 16, 16000, [fw.FrameSpec(b, [fw.Sub("fixed", 2, porder=0)]) for b in blocks])
            p += ".flac"
        paths.append(p)
        waves.append((x[:, 0] / 32768.0).astype(np.float32))
    for threads in (1, 4, 0):
        out = torch.full((16, 2100), 7.0)
        rows, lengths = audio_io.decode_batch(paths, 2048, out=out, threads=threads, verify=True)
        assert rows.shape == (13, 2100) and lengths.tolist() == [len(w) for w in waves]
        for i, w in enumerate(waves):
            k = min(len(w), 2048)
            assert np.array_equal(rows[i, :k].numpy(), w[:k]) and not rows[i, k:].any()
        assert (out[13:] == 7.0).all()                                  # rows beyond the batch are not touched
    with pytest.raises(audio_io.AudioError, match="u99"):
        audio_io.decode_batch(paths[:3] + [str(tmp_path / "u99.flac")], 2048)
    p8 = str(tmp_path / "r8k.wav")
    scipy.io.wavfile.write(p8, 8000, np.zeros(100, np.int16))
    with pytest.raises(audio_io.AudioError, match="8000 Hz"):
        audio_io.decode_batch([paths[0], p8], 2048)


def test_mutated_streams_never_crash_the_decoder(tmp_path):
    """400 damaged variants of valid streams: every one returns a status (scripts/fuzz_audio.py runs the same
    campaign, larger, under AddressSanitizer + UBSan)."""
    import subprocess
    import sys
    code = r'''
import sys, os, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import fuzz_audio as fz
from asvspoof2021_air_b200 import audio_io
rng = np.random.RandomState(3)
ok = bad = 0
for si, s in enumerate(fz.seeds(rng)):
    for t in range(80):
        p = os.path.join(%r, "m.bin")
        open(p, "wb").write(fz.mutate(s, t, rng))
        try:
            audio_io.decode(p, verify=True); ok += 1
        except audio_io.AudioError:
            bad += 1
print("done", ok, bad)
''' % (os.path.join(ROOT, "scripts"), os.path.join(ROOT, "tests"), ROOT, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("done"), r.stdout[-500:] + r.stderr[-2000:]
    ok, bad = (int(v) for v in r.stdout.split()[1:3])
    assert ok + bad == 400 and bad > 200
