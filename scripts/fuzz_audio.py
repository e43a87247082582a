"""Fuzz the native FLAC / WAV decoder (csrc/audio_io.cpp) under AddressSanitizer + UBSan.

Valid streams from tests/flac_writer.py, the RFC 9639 example and a scipy WAV are mutated (bit flips, truncation,
byte-run replacement, extreme bytes) and decoded by a sanitizer build of the decoder; any report or crash fails.
    python scripts/fuzz_audio.py [--cases 600]          # cases per seed stream
"""
import argparse
import io
import os
import subprocess
import sys
import tempfile

import numpy as np
import scipy.io.wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import flac_writer as fw  # noqa: E402

MAIN = r'''
#include <stdio.h>
#include <vector>
extern "C" int air_audio_decode_f32(const char*, float*, long long, long long*, int*, int);
extern "C" int air_audio_info(const char*, int*, int*, int*, long long*);
int main(int argc, char** argv) {
  std::vector<float> out(1 << 20);
  int hist[8] = {0};
  for (int i = 1; i < argc; ++i) {
    long long n = 0; int sr = 0, ch = 0, bits = 0;
    air_audio_info(argv[i], &sr, &ch, &bits, &n);
    int st = air_audio_decode_f32(argv[i], out.data(), (long long)out.size(), &n, &sr, 1);
    hist[st <= 0 && st >= -6 ? -st : 7]++;
  }
  printf("ok %d arg %d unsupported %d io %d format %d checksum %d nomem %d other %d\n", hist[0], hist[1], hist[2], hist[3],
         hist[4], hist[5], hist[6], hist[7]);
  return 0;
}
'''


def seeds(rng):
    def walk(n, ch=1):
        return np.round(np.cumsum(rng.randn(n, ch), axis=0) * 30).astype(np.int64).clip(-30000, 30000)
    lpc = fw.Sub("lpc", 8, coefs=[9000, -7000, 3000, -900, 200, -50, 10, -2], precision=15, shift=13, porder=3, method=1)
    out = [fw.encode_flac(walk(3000), 16, 16000, [fw.FrameSpec(1024, [fw.Sub("fixed", 2, porder=2)]), fw.FrameSpec(1024, [lpc]),
                                                   fw.FrameSpec(952, [fw.Sub("fixed", 1, escape_parts=(0,))])]),
           fw.encode_flac(walk(1200, 2), 16, 44100, [fw.FrameSpec(600, [fw.Sub("fixed", 2), fw.Sub("fixed", 2)], "mid_side"),
                                                      fw.FrameSpec(600, [fw.Sub("verbatim"), fw.Sub("fixed", 3)], "left_side", block_code=7)]),
           fw.encode_flac(np.zeros((4096, 1), np.int64), 16, 16000, [fw.FrameSpec(4096, [fw.Sub("constant")])], total=0, md5=False),
           bytes.fromhex("664c6143800000221000100000001f00001f07d0007000000018f8f9e396f5cbcfc6dc807f9977906b32"
                         "fff868020017e944004f6f313d1047d227cb6d090831452bdc2822228057a3")]
    b = io.BytesIO()
    scipy.io.wavfile.write(b, 16000, (rng.randn(500, 2) * 1000).astype(np.int16))
    out.append(b.getvalue())
    return out


def mutate(data, t, rng):
    d = bytearray(data)
    mode = t % 4
    if mode == 0:
        for _ in range(rng.randint(1, 4)):
            d[rng.randint(len(d))] ^= 1 << rng.randint(8)
    elif mode == 1:
        d = d[:rng.randint(1, len(d))]
    elif mode == 2:
        i = rng.randint(len(d))
        d[i:i + rng.randint(1, 8)] = bytes(rng.randint(0, 256, rng.randint(1, 8)).tolist())
    else:
        d[rng.randint(len(d))] = int(rng.choice([0, 0xff, 0x7f, 0x80]))
    return bytes(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=600)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    work = tempfile.mkdtemp(prefix="air_fuzz_")
    with open(os.path.join(work, "main.cpp"), "w") as f:
        f.write(MAIN)
    exe = os.path.join(work, "fuzz_asan")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "asvspoof2021_air_b200", "csrc", "audio_io.cpp"), os.path.join(work, "main.cpp"),
                    "-o", exe, "-lpthread"], check=True)
    rng = np.random.RandomState(args.seed)
    failed = 0
    for si, s in enumerate(seeds(rng)):
        paths = []
        for t in range(args.cases):
            p = os.path.join(work, "s%d_%05d" % (si, t))
            with open(p, "wb") as f:
                f.write(mutate(s, t, rng))
            paths.append(p)
        r = subprocess.run([exe] + paths, capture_output=True, text=True, timeout=1200)
        clean = r.returncode == 0 and "runtime error" not in r.stderr and "Sanitizer" not in r.stderr
        print("seed stream %d: %s  %s" % (si, "clean" if clean else "FAILED", r.stdout.strip()))
        if not clean:
            failed += 1
            print(r.stderr[-3000:])
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
