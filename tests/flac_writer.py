"""A small FLAC ENCODER for tests only -- written from the format specification (RFC 9639), independently of the
decoder in asvspoof2021_air_b200/csrc/audio_io.cpp, so that the decoder can be exercised on every bitstream feature
(there is no FLAC encoder in this image).  It is not efficient and makes no attempt to choose good parameters: the
caller decides, per frame and per channel, which subframe type / predictor / Rice layout to emit.

    write_flac(path, samples[(n, channels) int], bits, sample_rate, frames=[FrameSpec...])
"""
import hashlib
import struct

import numpy as np


class BitWriter:
    def __init__(self):
        self.bits = []

    def put(self, value, n):
        value &= (1 << n) - 1
        self.bits.extend((value >> (n - 1 - i)) & 1 for i in range(n))

    def unary(self, q):
        self.bits.extend([0] * q)
        self.bits.append(1)

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def tobytes(self):
        assert len(self.bits) % 8 == 0
        b = np.packbits(np.array(self.bits, dtype=np.uint8))
        return b.tobytes()


def crc8(data):
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xff if c & 0x80 else (c << 1) & 0xff
    return c


def crc16(data):
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xffff if c & 0x8000 else (c << 1) & 0xffff
    return c


def utf8_number(v):
    if v < 0x80:
        return bytes([v])
    out, n = [], 0
    while True:
        n += 1
        out.append(0x80 | (v & 0x3f))
        v >>= 6
        if v < (1 << (6 - n)):
            break
    lead = ((0xff << (7 - n)) & 0xff) | v
    return bytes([lead] + out[::-1])


class Sub:
    """How to code one channel of one frame.  kind: constant | verbatim | fixed | lpc."""

    def __init__(self, kind="fixed", order=2, coefs=None, precision=12, shift=10, wasted=0, porder=0, rice=None,
                 method=0, escape_parts=()):
        self.kind, self.order, self.coefs, self.precision, self.shift = kind, order, coefs, precision, shift
        self.wasted, self.porder, self.rice, self.method, self.escape_parts = wasted, porder, rice, method, set(escape_parts)


class FrameSpec:
    def __init__(self, block, subs, assignment="independent", variable=False, block_code=None, rate_code=None,
                 size_code=None):
        self.block, self.subs, self.assignment, self.variable = block, subs, assignment, variable
        self.block_code, self.rate_code, self.size_code = block_code, rate_code, size_code


def _zigzag(r):
    return 2 * r if r >= 0 else -2 * r - 1


def _residual(bw, res, block, order, sub):
    parts = 1 << sub.porder
    assert block % parts == 0 or sub.porder == 0
    bw.put(sub.method, 2)
    bw.put(sub.porder, 4)
    pbits, esc = (4, 15) if sub.method == 0 else (5, 31)
    i = 0
    for p in range(parts):
        count = (block >> sub.porder) - (order if p == 0 else 0)
        chunk = res[i:i + count]
        i += count
        if p in sub.escape_parts:
            width = max([1] + [int(abs(int(r))).bit_length() + 1 for r in chunk]) if len(chunk) and any(chunk) else 0
            bw.put(esc, pbits)
            bw.put(width, 5)
            for r in chunk:
                if width:
                    bw.put(int(r), width)
            continue
        if sub.rice is not None:
            k = sub.rice
        else:
            mean = (sum(_zigzag(int(r)) for r in chunk) / max(1, len(chunk)))
            k = min(esc - 1, max(0, int(mean).bit_length() - 1))
        bw.put(k, pbits)
        for r in chunk:
            u = _zigzag(int(r))
            bw.unary(u >> k)
            if k:
                bw.put(u & ((1 << k) - 1), k)
    assert i == len(res)


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _subframe(bw, x, bps, sub):
    x = [int(v) for v in x]
    block = len(x)
    if sub.wasted:
        assert all(v % (1 << sub.wasted) == 0 for v in x)
        x = [v >> sub.wasted for v in x]
    code = {"constant": 0, "verbatim": 1}.get(sub.kind)
    if sub.kind == "fixed":
        code = 8 + sub.order
    elif sub.kind == "lpc":
        code = 31 + sub.order
    bw.put(0, 1)
    bw.put(code, 6)
    if sub.wasted:
        bw.put(1, 1)
        bw.unary(sub.wasted - 1)
    else:
        bw.put(0, 1)
    b = bps - sub.wasted
    if sub.kind == "constant":
        assert len(set(x)) == 1
        bw.put(x[0], b)
    elif sub.kind == "verbatim":
        for v in x:
            bw.put(v, b)
    else:
        order = sub.order
        coefs, shift = (FIXED[order], 0) if sub.kind == "fixed" else (list(sub.coefs), sub.shift)
        for v in x[:order]:
            bw.put(v, b)
        if sub.kind == "lpc":
            bw.put(sub.precision - 1, 4)
            bw.put(shift, 5)
            for c in coefs:
                assert -(1 << (sub.precision - 1)) <= c < (1 << (sub.precision - 1))
                bw.put(c, sub.precision)
        res = []
        for i in range(order, block):
            pred = sum(coefs[j] * x[i - 1 - j] for j in range(order)) >> shift
            res.append(x[i] - pred)
        _residual(bw, res, block, order, sub)


def _block_code(block):
    if block == 192:
        return 1
    for c in range(2, 6):
        if block == 576 << (c - 2):
            return c
    for c in range(8, 16):
        if block == 256 << (c - 8):
            return c
    return 6 if block <= 256 else 7


RATE_CODES = {88200: 1, 176400: 2, 192000: 3, 8000: 4, 16000: 5, 22050: 6, 24000: 7, 32000: 8, 44100: 9, 48000: 10, 96000: 11}
SIZE_CODES = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}


def encode_frame(x, bits, sample_rate, spec, number):
    """x: (block, channels) ints.  number: frame number (fixed blocking) or first sample number (variable)."""
    block, nch = x.shape
    assert block == spec.block
    bw = BitWriter()
    bw.put(0b11111111111110, 14)
    bw.put(0, 1)
    bw.put(1 if spec.variable else 0, 1)
    bcode = spec.block_code if spec.block_code is not None else _block_code(block)
    rcode = spec.rate_code if spec.rate_code is not None else RATE_CODES.get(sample_rate, 0)
    scode = spec.size_code if spec.size_code is not None else SIZE_CODES.get(bits, 0)
    bw.put(bcode, 4)
    bw.put(rcode, 4)
    ch = {"independent": nch - 1, "left_side": 8, "right_side": 9, "mid_side": 10}[spec.assignment]
    bw.put(ch, 4)
    bw.put(scode, 3)
    bw.put(0, 1)
    for byte in utf8_number(number):
        bw.put(byte, 8)
    if bcode == 6:
        bw.put(block - 1, 8)
    elif bcode == 7:
        bw.put(block - 1, 16)
    if rcode == 12:
        bw.put(sample_rate // 1000, 8)
    elif rcode == 13:
        bw.put(sample_rate, 16)
    elif rcode == 14:
        bw.put(sample_rate // 10, 16)
    bw.put(crc8(bw.tobytes()), 8)
    cols = [[int(v) for v in x[:, c]] for c in range(nch)]
    widths = [bits] * nch
    if spec.assignment != "independent":
        left, right = cols
        side = [a - b for a, b in zip(left, right)]
        if spec.assignment == "left_side":
            cols, widths = [left, side], [bits, bits + 1]
        elif spec.assignment == "right_side":
            cols, widths = [side, right], [bits + 1, bits]
        else:
            cols, widths = [[(a + b) >> 1 for a, b in zip(left, right)], side], [bits, bits + 1]
    for c in range(nch):
        _subframe(bw, cols[c], widths[c], spec.subs[c])
    bw.align()
    body = bw.tobytes()
    return body + struct.pack(">H", crc16(body))


def md5_of(samples, bits):
    nbytes = (bits + 7) // 8
    flat = np.asarray(samples, dtype=np.int64).reshape(-1)
    raw = b"".join(int(v & ((1 << (8 * nbytes)) - 1)).to_bytes(nbytes, "little") for v in flat)
    return hashlib.md5(raw).digest()


def streaminfo(samples, bits, sample_rate, blocks, frame_sizes, total=None, md5=True):
    n, nch = samples.shape
    total = n if total is None else total
    bw = BitWriter()
    bw.put(min(blocks), 16)
    bw.put(max(blocks), 16)
    bw.put(min(frame_sizes), 24)
    bw.put(max(frame_sizes), 24)
    bw.put(sample_rate, 20)
    bw.put(nch - 1, 3)
    bw.put(bits - 1, 5)
    bw.put(total, 36)
    body = bw.tobytes() + (md5_of(samples, bits) if md5 else bytes(16))
    assert len(body) == 34
    return body


def metadata_block(kind, body, last):
    return bytes([(0x80 if last else 0) | kind]) + len(body).to_bytes(3, "big") + body


def encode_flac(samples, bits, sample_rate, frames, total=None, md5=True, extra_blocks=(), prefix=b"", suffix=b""):
    samples = np.asarray(samples, dtype=np.int64)
    if samples.ndim == 1:
        samples = samples[:, None]
    out, pos = [], 0
    for i, spec in enumerate(frames):
        x = samples[pos:pos + spec.block]
        out.append(encode_frame(x, bits, sample_rate, spec, pos if spec.variable else i))
        pos += spec.block
    assert pos == samples.shape[0], (pos, samples.shape)
    info = streaminfo(samples, bits, sample_rate, [f.block for f in frames], [len(f) for f in out], total, md5)
    blocks = [(0, info)] + list(extra_blocks)
    meta = b"".join(metadata_block(k, b, i == len(blocks) - 1) for i, (k, b) in enumerate(blocks))
    return prefix + b"fLaC" + meta + b"".join(out) + suffix


def write_flac(path, samples, bits, sample_rate, frames, **kw):
    data = encode_flac(samples, bits, sample_rate, frames, **kw)
    with open(path, "wb") as f:
        f.write(data)
    return data
