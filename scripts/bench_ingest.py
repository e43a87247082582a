"""Host-side ingest throughput: FLAC files -> pinned float32 rows through the native batch
decoder, by thread count; with a CUDA device also through data.Prefetcher onto the device (decode + H2D overlapped).
No GPU needed for the decode figures.
    python scripts/bench_ingest.py [--files 512] [--seconds 4] > profiles/rNN_ingest.json
The test stream is what libFLAC -5 typically emits for speech: 4096-sample blocks, 8th-order LPC, Rice partition
order 4 (written by tests/flac_writer.py; one file's bytes are reused under many names)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import flac_writer as fw  # noqa: E402
from asvspoof2021_air_b200 import audio_io, data  # noqa: E402


def make_folder(n_files, seconds):
    rng = np.random.RandomState(0)
    n = 16000 * seconds
    x = np.zeros(n + 2)
    e = rng.randn(n + 2)
    for i in range(2, n + 2):
        x[i] = 1.6 * x[i - 1] - 0.8 * x[i - 2] + e[i]
    x = np.round(x[2:] / np.abs(x).max() * 8000).astype(np.int64)
    blocks = [4096] * (n // 4096) + ([n % 4096] if n % 4096 else [])
    co = [13107, -6553, 0, 0, 0, 0, 0, 0]
    sub = lambda b: fw.Sub("lpc", 8, coefs=co, precision=15, shift=13, porder=4 if b == 4096 else 0)
    blob = fw.encode_flac(x, 16, 16000, [fw.FrameSpec(b, [sub(b)]) for b in blocks])
    d = tempfile.mkdtemp(prefix="air_ingest_")
    proto = []
    for i in range(n_files):
        with open(os.path.join(d, "LA_T_%07d.flac" % i), "wb") as f:
            f.write(blob)
        proto.append("LA_0000 LA_T_%07d - - %s" % (i, "bonafide" if i % 2 == 0 else "spoof"))
    with open(os.path.join(d, "proto.txt"), "w") as f:
        f.write("\n".join(proto) + "\n")
    return d, len(blob), n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=512)
    ap.add_argument("--seconds", type=int, default=4)
    ap.add_argument("--batch", type=int, default=256)
    args = ap.parse_args()
    folder, nbytes, n = make_folder(args.files, args.seconds)
    paths = [os.path.join(folder, "LA_T_%07d.flac" % i) for i in range(args.files)]
    out = torch.zeros(args.files, n, pin_memory=torch.cuda.is_available())
    audio_io.decode_batch(paths, n, out=out, threads=0)                  # warm the page cache and the buffers
    cores = os.cpu_count() or 1
    res = {"files": args.files, "seconds_per_file": args.seconds, "flac_bytes_per_file": nbytes,
           "compression": nbytes / (2.0 * n), "host_cores": cores, "decode": []}
    for th in sorted({1, 2, 4, 8, cores}):
        if th > cores:
            continue
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            audio_io.decode_batch(paths, n, out=out, threads=th)
            best = min(best, time.perf_counter() - t0)
        res["decode"].append({"threads": th, "utterances_per_s": args.files / best,
                              "x_realtime_per_thread": args.files * args.seconds / best / th})
    t0 = time.perf_counter()
    audio_io.decode_batch(paths, n, out=out, threads=0, verify=True)
    res["decode_with_md5"] = {"threads": cores, "utterances_per_s": args.files / (time.perf_counter() - t0)}
    if torch.cuda.is_available():
        src = data.WaveFolder(folder, os.path.join(folder, "proto.txt"))
        order = [list(range(lo, min(lo + args.batch, len(src)))) for lo in range(0, len(src), args.batch)] * 4
        dev = torch.device("cuda", 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        total = 0
        for waves, lengths, labels, names, start in data.Prefetcher(src, order, depth=2, device=dev):
            total += len(names)
            waves.sum()                                                  # consume on the compute stream
        torch.cuda.synchronize()
        res["prefetch_to_device"] = {"utterances_per_s": total / (time.perf_counter() - t0), "batch": args.batch}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
