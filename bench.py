#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl reference]

Workloads
  resnet_train  (default when the train engine is built) wave -> LFCC -> ResNet-18-OC fwd/bwd ->
                OC-Softmax -> Adam/SGD, B=256/GPU, bf16, synthetic 4 s @ 16 kHz waves
  ecapa_train   same with ECAPA-TDNN-512
  ecapa_score   generate_score.py inference, B=1024
  lfcc          the fused LFCC kernel alone, B=256 (HBM roofline of the LFCC kernel)
  resnet_adv    resnet_train with the --ADV_AUG channel classifier (second encoder forward per step)
The default invocation (no --workload) also attaches short runs of lfcc / ecapa_train / ecapa_score / resnet_adv as
`also` (value, ms, roofline fraction, clocks each).

One JSON line is printed by rank 0 (contract in the task statement): value = whole-job
utterances/s with inputs resident in HBM; e2e = same metric through the public Python API with
pinned-host inputs, H2D/D2H inside the timed region; roofline = dominant kernel vs the measured
peak in MEASURED_PEAKS.json; cpu_baseline = the oracle port of the reference's CPU path timed on
this box's host cores over a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WAVE_LEN = 64000
LFCC_BYTES_PER_UTT = 4 * WAVE_LEN + 4 * 401 * 60      # 352 240 B


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of a kernel, from the tracked summary of the round's
    `ncu --set full` captures (profiles/r02_ncu_metrics.json, written by scripts/gpu_ncu_r02.sh); None when absent."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_metrics.json")
    try:
        with open(path) as f:
            d = json.load(f).get(key) or {}
        return d.get("traffic_bytes")
    except (OSError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_MAX_CTAS", "2")          # see asvspoof2021_air_b200/parallel.py:init_from_env
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def timed(step_fn, steps, warmup, world):
    for i in range(warmup):
        step_fn(i)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step_fn(warmup + i)
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


# ----------------------------------------------------------------------------------------------
def _cpu_lfcc():
    """(callable, kind): the reference's own LFCC module from oracle/_ref when present, else the oracle port."""
    from oracle import ref_shim
    if ref_shim.use_copy_if_needed():
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = ref_shim.load("feature_extraction").LFCC(320, 160, 512, 16000, 20, with_energy=False)

        def run(w):
            with torch.no_grad():
                return mod(w.clone())                      # the reference pre-emphasises its input in place
        return run, "reference"
    from oracle import lfcc_torch
    return lfcc_torch.TorchLFCC(), "port"


def cpu_lfcc_baseline(seconds=10.0, batch=32):
    """The reference LFCC (feature_extraction.py:93-138, torch fp32, all host threads) on a bounded sample."""
    from oracle import state_spec as ss
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    m, kind = _cpu_lfcc()
    w = ss.seeded_waves(batch, WAVE_LEN, seed=0)
    m(w)
    t0, it = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds and it < 200:
        m(w)
        it += 1
    dt = time.perf_counter() - t0
    return {"value": batch * it / dt, "unit": "utterances/s", "cores": n, "kind": kind,
            "sample": "%d x LFCC of %d synthetic 4 s waves (feature_extraction.py:93-138, %s)" % (it, batch, kind)}


def lfcc_config(B, nbuf=4):
    return {"workload": "lfcc: fused wave->LFCC kernel, B=%d/GPU, 4 s @ 16 kHz, fp32 out (B,401,60)" % B,
            "batch_per_gpu": B, "l2": "inputs rotate over %d buffers (360 MB > L2)" % nbuf}


def det_config(n_tar, n_non):
    return {"workload": "det: EER of %d bona fide + %d spoof scores, both orientations (main_train.py:662-664)" % (n_tar, n_non),
            "l2": "a 256 MB buffer is rewritten before every timed step (per-step CUDA events, summed)"}


def run_lfcc(args, rank, world):
    from asvspoof2021_air_b200.feature_extraction import LFCC
    from asvspoof2021_air_b200.bench_train import _waves
    B = args.batch or 256
    mod = LFCC(320, 160, 512, 16000, 20).cuda()
    if mod.impl == "auto":
        # the tensor-core kernel (what LFCC.forward and the train step launch) at the contract's fp32 output; the fp32 FFT
        # kernel (AIR_LFCC_IMPL=fft, pinned by the Trainer's fp32 parity mode) is timed beside it as `fft_kernel`
        mod.impl = "tc"
    nbuf = 4                                   # 4 x (65.5 MB in + 24.6 MB out) = 360 MB > 126 MB L2
    waves = [_waves(B, rank * 16 + i).cuda() for i in range(nbuf)]
    outs = [torch.empty(B, 401, 60, device="cuda") for _ in range(nbuf)]
    launches = [0]

    def step(i):
        mod.extract(waves[i % nbuf], feat_len=0, layout="btd", dtype=torch.float32, out=outs[i % nbuf], fseg=args.fseg)
        launches[0] += 1

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    ms = timed(step, args.steps, args.warmup, world)
    clocks = sampler.stop()
    n_timed_launches = args.steps
    ms_fft = None
    if mod.impl == "tc" and "AIR_LFCC_IMPL" not in os.environ:
        mod.impl = "fft"
        ms_fft = timed(step, args.steps, args.warmup, world)
        mod.impl = "tc"
    # kernel time = step time here (one kernel per step, back to back on one stream)
    per_launch_s = ms / 1e3 / args.steps
    peaks = measured_peaks()
    achieved = B * LFCC_BYTES_PER_UTT / per_launch_s / 1e9

    # e2e: pinned host waves -> H2D -> kernel -> D2H of the features' checksum row
    host = [w.cpu().pin_memory() for w in waves[:2]]
    dev_in = torch.empty(B, WAVE_LEN, device="cuda")
    res_host = torch.empty(B, 60, pin_memory=True)

    def step_e2e(i):
        dev_in.copy_(host[i % 2], non_blocking=True)
        y = mod.extract(dev_in, feat_len=0, layout="btd", dtype=torch.float32, out=outs[i % nbuf], fseg=args.fseg)
        res_host.copy_(y[:, 200, :], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e = timed(step_e2e, args.steps, args.warmup, world)
    line = {
        "metric": "LFCC utterances/sec (4 s@16 kHz)", "value": world * B * args.steps / (ms / 1e3),
        "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": lfcc_config(B, nbuf),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch (profiles/r02_ncu_metrics.json): the 65.5 MB
                     # of waves are read once; most of the 24.6 MB of output is still L2-resident at kernel end
                     "traffic": ncu_traffic("lfcc_tc") if (mod.impl == "tc" and B == 256) else None, "peak_src": peaks["src"],
                     "kernel": ("air_lfcc_tc::lfcc_tc_kernel (tensor-core folded DFT)" if mod.impl == "tc"
                                else "air_lfcc::lfcc_kernel (radix FFT on CUDA cores)"),
                     "bytes_per_launch": B * LFCC_BYTES_PER_UTT,
                     "fft_kernel": None if ms_fft is None else {
                         "kernel": "air_lfcc::lfcc_kernel (radix FFT in fp32 on CUDA cores; AIR_LFCC_IMPL=fft, the Trainer's fp32 parity mode)",
                         "ms_per_launch": ms_fft / args.steps,
                         "frac": B * LFCC_BYTES_PER_UTT / (ms_fft / 1e3 / args.steps) / 1e9 / peaks["hbm_gbs"]}},
        "e2e": {"value": world * B * args.steps / (ms_e2e / 1e3), "unit": "utterances/s",
                "h2d_bytes_per_step": B * WAVE_LEN * 4, "d2h_bytes_per_step": B * 60 * 4},
        "gpu_launches": n_timed_launches, "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_lfcc_baseline()
    return line


# ----------------------------------------------------------------------------------------------
DET_TRIALS = (7355, 63882)          # ASVspoof 2019 LA eval: bona fide, spoof trials


def _det_scores(n_tar, n_non, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_tar, generator=g) + 1.0, torch.randn(n_non, generator=g) - 1.0


def cpu_det_baseline(n_tar, n_non, seconds=10.0):
    """numpy restatement of eval_metrics.compute_eer (both orientations, main_train.py:662-664), one host thread."""
    from oracle import metrics_oracle as mo
    t, n = (x.numpy() for x in _det_scores(n_tar, n_non, 0))
    mo.eer(t, n)
    t0, it = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds and it < 500:
        mo.eer(t, n)
        mo.eer(t, n, negate=True)
        it += 1
    dt = time.perf_counter() - t0
    return {"value": (n_tar + n_non) * it / dt, "unit": "trials/s", "cores": 1, "kind": "port",
            "sample": "%d x (EER + EER of negated scores) of %d + %d synthetic scores (numpy)" % (it, n_tar, n_non)}


def run_det(args, rank, world):
    """EER of one evaluation set, both score orientations, sorted / reduced on the GPU."""
    from asvspoof2021_air_b200 import eval_metrics as em, ops
    n_tar, n_non = DET_TRIALS if not args.batch else (args.batch // 8, args.batch - args.batch // 8)
    n = n_tar + n_non
    tar, non = (x.cuda() for x in _det_scores(n_tar, n_non, rank))
    ws = torch.empty(ops.det_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib_launches = _lib_launches()

    def step(i):
        a = em.det(tar, non, workspace=ws)
        b = em.det(tar, non, negate=True, workspace=ws)
        return a, b

    def timed_flushed(fn, sync_result):
        for i in range(args.warmup):
            fn(i)
        barrier(world)
        total = 0.0
        for i in range(args.steps):
            flush.zero_()                                  # 256 MB > 126 MB L2: every step starts cold
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(i)
            if sync_result:
                r = min(r[0].host()["eer"], r[1].host()["eer"])
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        barrier(world)
        return max_over_ranks(total, world)

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    l0 = lib_launches()
    ms = timed_flushed(step, False)
    per_step_launches = (lib_launches() - l0) // (args.steps + args.warmup)
    clocks = sampler.stop()
    host = [x.cpu().pin_memory() for x in (tar, non)]

    def step_e2e(i):
        t, m = host[0].cuda(non_blocking=True), host[1].cuda(non_blocking=True)
        return em.det(t, m, workspace=ws), em.det(t, m, negate=True, workspace=ws)

    ms_e2e = timed_flushed(step_e2e, True)
    peaks = measured_peaks()
    passes = 5
    algo = 2 * (n * 4 + 80)                                # per step: the scores read once per orientation + 10 doubles out
    sort_bytes = 2 * n * (4 + 9 + passes * (9 + 9 + 9) + 9 + 9)   # what the kernels move: keys+class per radix pass etc.
    per_step_s = ms / 1e3 / args.steps
    line = {
        "metric": "detection trials/sec (EER, both orientations)", "value": world * n * args.steps / (ms / 1e3),
        "unit": "trials/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 keys / u32 counts", "data": "synthetic",
        "config": det_config(n_tar, n_non),
        "roofline": {"bound": "hbm", "achieved": algo / per_step_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": algo / per_step_s / 1e9 / peaks["hbm_gbs"], "traffic": None, "peak_src": peaks["src"],
                     "kernel": "air_det:: radix sort (hist / scan / scatter x %d) + curve kernels" % passes,
                     "bytes_per_launch": algo, "moved_bytes_per_step": sort_bytes,
                     "moved_gbs": sort_bytes / per_step_s / 1e9,
                     "note": "launch-latency bound at this size: %d dependent launches per step" % per_step_launches},
        "e2e": {"value": world * n * args.steps / (ms_e2e / 1e3), "unit": "trials/s", "h2d_bytes_per_step": n * 4,
                "d2h_bytes_per_step": 160},
        "gpu_launches": per_step_launches * args.steps, "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_det_baseline(n_tar, n_non)
    return line


def _lib_launches():
    from asvspoof2021_air_b200 import _lib
    return lambda: _lib.LAUNCHES[0]


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port; /root/reference is not on
    the GPU box), all host threads, bounded sample per step."""
    if rank != 0:
        return None
    if args.workload == "lfcc":
        from oracle import state_spec as ss
        n = os.cpu_count() or 1
        torch.set_num_threads(n)
        B = 32
        m, kind = _cpu_lfcc()
        w = ss.seeded_waves(B, WAVE_LEN, seed=0)
        for _ in range(args.warmup):
            m(w)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m(w)
        dt = time.perf_counter() - t0
        v = B * args.steps / dt
        return {"impl": "reference", "metric": "LFCC utterances/sec (4 s@16 kHz)", "value": v, "unit": "utterances/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": lfcc_config(args.batch or 256),
                "cpu_baseline": {"value": v, "unit": "utterances/s", "cores": n, "kind": kind,
                                 "sample": "%d waves of the workload per step x %d steps (feature_extraction.LFCC, %s)"
                                           % (B, args.steps, kind)},
                "e2e": {"value": v, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.workload == "det":
        from oracle import metrics_oracle as mo
        n_tar, n_non = DET_TRIALS
        t, n = (x.numpy() for x in _det_scores(n_tar, n_non, 0))
        for _ in range(args.warmup):
            mo.eer(t, n)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            mo.eer(t, n)
            mo.eer(t, n, negate=True)
        dt = time.perf_counter() - t0
        v = (n_tar + n_non) * args.steps / dt
        return {"impl": "reference", "metric": "detection trials/sec (EER, both orientations)", "value": v,
                "unit": "trials/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": det_config(n_tar, n_non),
                "cpu_baseline": {"value": v, "unit": "trials/s", "cores": 1, "kind": "port",
                                 "sample": "%d + %d scores x %d steps (numpy restatement of eval_metrics.compute_eer)"
                                           % (n_tar, n_non, args.steps)},
                "e2e": {"value": v, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    from asvspoof2021_air_b200 import bench_train
    return bench_train.run_reference(args, rank, world)


ALSO = ("lfcc", "ecapa_train", "ecapa_score", "resnet_adv")


def also_results(args, rank, world):
    """Short runs of the other workloads of the path, attached to the default line as `also` so that they land in the
    driver's BENCH record (value, ms, roofline fraction, clocks of each; no CPU arm)."""
    import copy
    import gc
    out = {}
    for w in ALSO:
        a = copy.copy(args)
        a.workload, a.steps, a.warmup, a.no_cpu_baseline, a.batch = w, min(args.steps, 10), 3, True, 0
        try:
            if w == "lfcc":
                ln = run_lfcc(a, rank, world)
            else:
                from asvspoof2021_air_b200 import bench_train
                ln = bench_train.run(a, rank, world, helpers=sys.modules[__name__])
            r = ln["roofline"]
            out[w] = {"value": ln["value"], "unit": ln["unit"], "ms_per_step": ln["ms_per_step"], "steps": a.steps,
                      "e2e": ln["e2e"]["value"], "gpu_launches": ln["gpu_launches"],
                      "roofline": {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "kernel",
                                                         "whole_step_tflops", "conv_ms_per_step", "fft_kernel")
                                   if r.get(k) is not None},
                      "clocks": ln["clocks"], "workload": ln["config"]["workload"]}
            if ln.get("kernels"):
                out[w]["kernels_ms"] = {k: v["ms_per_step"] for k, v in ln["kernels"].items()}
        except Exception as e:                       # a failing side workload must not take the headline line with it
            out[w] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        gc.collect()
        torch.cuda.empty_cache()
    return out


def default_workload():
    try:
        from asvspoof2021_air_b200 import bench_train  # noqa: F401
        return "resnet_train"
    except ImportError:
        return "lfcc"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=None,
                    choices=[None, "lfcc", "resnet_train", "ecapa_train", "ecapa_score", "resnet_adv", "det"])
    ap.add_argument("--no-also", dest="no_also", action="store_true",
                    help="default invocation only: skip the short runs of the other workloads (`also` in the JSON line)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--fseg", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    default_invocation = args.workload is None
    if args.workload is None:
        args.workload = default_workload()

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        line = run_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    rank, world, _ = dist_setup(args.gpus)
    if args.workload == "lfcc":
        line = run_lfcc(args, rank, world)
    elif args.workload == "det":
        line = run_det(args, rank, world)
    else:
        from asvspoof2021_air_b200 import bench_train
        line = bench_train.run(args, rank, world, helpers=sys.modules[__name__])
    if default_invocation and world == 1 and not args.no_also:
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        line["also"] = also_results(args, rank, world)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
