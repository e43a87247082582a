"""Train-step / scoring workloads of bench.py (resnet_train, ecapa_train, ecapa_score).

A "step" is one pass of the hot path over one batch of synthetic 4 s @ 16 kHz waves:
wave -> fused LFCC -> net forward -> OC-Softmax -> backward -> [NCCL all-reduce] -> Adam/SGD
(SURVEY.md section 8d).  `value` has the waves resident in HBM; `e2e` goes through the public
Trainer API from pinned host memory (H2D of every step's waves and D2H of its loss inside the
timed region, double-buffered on a copy stream).  The roofline pass re-runs a few steps with CUDA
events around every C-ABI launch (ops.Profile) and reports the conv stack against the measured
bf16 peak; the unprofiled timed region is what `value` comes from.
"""
import os
import time

import torch

WAVE_LEN = 64000
# algorithmic conv/linear FLOPs per utterance of one train step
TRAIN_FLOPS_PER_UTT = {"resnet": 47.12e9, "ecapa": 22.86e9}
FWD_FLOPS_PER_UTT = {"resnet": 15.709e9, "ecapa": 7.698e9}


def _labels(B, seed):
    g = torch.Generator().manual_seed(seed + 7)
    lab = torch.randint(0, 2, (B,), generator=g)
    lab[0], lab[1] = 0, 1                   # both classes present
    return lab


def _waves(B, seed):
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(B, WAVE_LEN, generator=g)


NBUF = 2
NAMES = {"resnet": "ResNet-18-OC", "ecapa": "ECAPA-TDNN-512"}


def workload_config(workload, B, world, overlap=None):
    """The `config` object of the JSON line -- shared by this arm and `--impl reference` (which times a bounded sample of
    the SAME workload: the driver compares the two configs)."""
    arch = "resnet" if workload.startswith("resnet") else "ecapa"
    scoring = workload.endswith("_score")
    if workload == "resnet_adv":
        return {"workload": "resnet_adv: wave->LFCC->ResNet-18-OC fwd/bwd + OC-Softmax + gradient-reversed channel classifier "
                            "(--ADV_AUG --LA_aug, 60 classes) + Adam(L2)/SGD, then a SECOND encoder forward and the classifier's "
                            "own Adam step (main_train.py:377-453), B=%d/GPU, 4 s @ 16 kHz, feat_len 750" % B,
                "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                "l2": "per-step activations (> 4 GB) exceed the 126 MB L2; waves rotate over %d buffers" % NBUF,
                "streams": "weight-gradient kernels on a side stream"}
    if overlap is None:
        overlap = os.environ.get("AIR_OVERLAP_WGRAD", "1") != "0"
    return {"workload": "%s: wave->LFCC->%s %s + OC-Softmax%s, B=%d/GPU, 4 s @ 16 kHz, feat_len 750 (repeat pad)"
                        % (workload, NAMES[arch], "eval forward" if scoring else "fwd/bwd",
                           "" if scoring else " + Adam(L2)/SGD", B),
            "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "l2": "per-step activations (> 4 GB) exceed the 126 MB L2; waves rotate over %d buffers" % NBUF,
            "streams": "weight-gradient kernels on a side stream (overlap the HBM-bound BatchNorm backward); the "
                       "roofline / per-family pass runs serialised" if (overlap and not scoring) else "single stream"}


def run(args, rank, world, helpers):
    from . import _lib, ops
    from .trainer import Trainer
    arch = "resnet" if args.workload.startswith("resnet") else "ecapa"
    scoring = args.workload.endswith("_score")
    B = args.batch or (1024 if scoring else 256)
    pg = None
    if world > 1:
        import torch.distributed as dist
        pg = dist.group.WORLD
    tr = Trainer(arch=arch, device="cuda", process_group=pg, seed=688)
    nbuf = NBUF
    waves = [_waves(B, rank * 16 + i).cuda() for i in range(nbuf)]
    labels = _labels(B, rank).cuda()
    adv = args.workload == "resnet_adv"
    channels = None
    if adv:                                   # --ADV_AUG --LA_aug: one 60-class codec head (main_train.py:211-224)
        tr.attach_adversaries([60], lambda_=0.05, lr_d=1e-4, seed=688)
        channels = torch.randint(0, 60, (B,), generator=torch.Generator().manual_seed(rank + 3)).cuda()

    def run_step(w, i):
        if scoring:
            return tr.score_step(w)
        return tr.train_step(w, labels, channels=channels, step_seed=i)

    def step(i):
        run_step(waves[i % nbuf], i)

    step(0)                                   # allocate activations / pack weights outside the timing
    torch.cuda.synchronize()
    sampler = helpers.ClockSampler(torch.cuda.current_device())
    sampler.start()
    l0 = _lib.LAUNCHES[0]
    ms = helpers.timed(step, args.steps, args.warmup, world)
    launches = (_lib.LAUNCHES[0] - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop()
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: pinned host waves -> H2D (copy stream, double-buffered) -> step -> D2H loss ----
    host = [w.cpu().pin_memory() for w in waves]
    dev = [torch.empty_like(waves[0]) for _ in range(2)]
    res_host = torch.empty(B if scoring else 1, pin_memory=True)
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    main = torch.cuda.current_stream()

    def issue_copy(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            dev[i % 2].copy_(host[i % nbuf], non_blocking=True)
            ready[i % 2].record(copy_stream)

    for e in consumed:
        e.record(main)
    state = {"next": 0}

    def step_e2e(i):
        while state["next"] <= i + 1:          # this step's waves + prefetch of the next step's
            issue_copy(state["next"])
            state["next"] += 1
        main.wait_event(ready[i % 2])
        out = run_step(dev[i % 2], i)
        consumed[i % 2].record(main)
        res_host.copy_(out.reshape(-1), non_blocking=True)
        main.synchronize()                     # the user reads the loss / scores every step

    ms_e2e = helpers.timed(step_e2e, args.steps, args.warmup, world)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- roofline pass: per-family device time over a few profiled steps ----
    peaks = helpers.measured_peaks()
    nprof = 3
    overlap = tr.engine.overlap_wgrad
    tr.engine.overlap_wgrad = False          # per-launch times must not include a concurrently running kernel
    with ops.Profile() as prof:
        for i in range(nprof):
            step(i)
    tr.engine.overlap_wgrad = overlap
    fam = prof.summary()
    conv = {k: v for k, v in fam.items() if k.startswith("conv_")}
    conv_ms = sum(v["ms"] for v in conv.values())
    conv_flops = sum(v["flops"] for v in conv.values())
    total_ms = sum(v["ms"] for v in fam.values())
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    breakdown = {k: {"ms_per_step": round(v["ms"] / nprof, 4), "launches_per_step": v["launches"] // nprof,
                     **({"tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1)} if v["flops"] else {}),
                     **({"gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1)} if v["bytes"] else {})}
                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    lf = fam.get("lfcc")
    lfcc_roof = None
    if lf and lf["ms"] > 0:
        gbs = lf["bytes"] / (lf["ms"] / 1e3) / 1e9
        lfcc_roof = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                     "kernel": "air_lfcc_tc::lfcc_tc_kernel (fused wave -> padded bf16 model-layout LFCC, inside the step)",
                     "bytes_per_launch": lf["bytes"] / nprof, "ms_per_step": lf["ms"] / nprof}
    flops_per_utt = (FWD_FLOPS_PER_UTT if scoring else TRAIN_FLOPS_PER_UTT)[arch] + (FWD_FLOPS_PER_UTT[arch] if adv else 0.0)
    line = {
        "metric": "utterances/sec (4 s@16 kHz) %s" % ("scoring" if scoring else "train-step"),
        "value": value, "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args.workload, B, world, overlap),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None,
                     # ncu --set full of the current kernels (profiles/r02_ncu_metrics.json): the layer-1 3x3 fprop in its
                     # in-step form (residual + fused BatchNorm statistics) moves 1 295 MB of DRAM traffic for 1 327 MB
                     # algorithmic (x + residual + out), the layer-1 wgrad 890 MB for 885 MB -- no re-reads
                     "traffic": helpers.ncu_traffic("patch_l1_stats") if (arch == "resnet" and B == 256 and not scoring) else None,
                     "traffic_kernel": "air_patch::conv_patch_kernel, layer-1 3x3 fprop (+ residual, fused BN statistics)"
                                       if arch == "resnet" else None,
                     "peak_src": peaks["src"] + " (sustained)",
                     "kernel": "conv stack (all tcgen05): air_patch::conv_patch_kernel (3x3 / 1x1 fprop + dgrad, stride-2 dgrad "
                               "by parity), air_wpatch::conv3x3_wgrad_patch_kernel, air_gemm::conv_gemm_kernel, "
                               "air_wgrad::conv_wgrad_kernel",
                     "flops_per_step": conv_flops / nprof, "conv_ms_per_step": conv_ms / nprof,
                     "conv_share_of_step": conv_ms / total_ms if total_ms else None,
                     "whole_step_tflops": B * flops_per_utt / (ms / args.steps / 1e3) / 1e12},
        "roofline_lfcc": lfcc_roof,
        "kernels": breakdown,
        "e2e": {"value": e2e, "unit": "utterances/s", "h2d_bytes_per_step": B * WAVE_LEN * 4,
                "d2h_bytes_per_step": int(res_host.numel()) * 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
    }
    red = getattr(tr, "reducer", None)
    if red is not None and red.exposed_ms() is not None:
        line["exchange_exposed_ms_per_step"] = red.exposed_ms()      # AIR_DDP_TRACE=1: compute stream waiting for the all-reduce
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(arch, scoring, seconds=15.0)
    return line


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (torch-CPU LFCC feature_extraction.py:93-138 ->
# pad dataset.py:519-522 -> reference net + OC-Softmax + Adam/SGD, fp32), all host threads.
# ----------------------------------------------------------------------------------------------
class _CpuStep:
    def __init__(self, arch, scoring, B):
        from oracle import lfcc_torch, nets_oracle as no, state_spec as ss
        self.no, self.arch, self.scoring, self.B = no, arch, scoring, B
        self.lfcc = lfcc_torch.TorchLFCC()
        self.spec = ss.resnet_spec() if arch == "resnet" else ss.ecapa_spec()
        self.sd = ss.seeded_state(self.spec, 11)
        skip = ("fc_mu.",) if arch == "resnet" else ("fc7.", "bn7.")
        self.keys = [k for k in ss.trainable_keys(self.spec) if not k.startswith(skip)]
        for k in self.keys:
            self.sd[k].requires_grad_(not scoring)
        self.center = ss.seeded_center(256, 11).requires_grad_(not scoring)
        self.m = {k: torch.zeros_like(self.sd[k]) for k in self.keys}
        self.v = {k: torch.zeros_like(self.sd[k]) for k in self.keys}
        self.waves = ss.seeded_waves(B, WAVE_LEN, seed=0)
        self.labels = ss.seeded_labels(B, 0)
        self.step = 0

    def __call__(self):
        no = self.no
        y = self.lfcc(self.waves)                                         # (B,401,60)
        idx = torch.arange(750) % y.shape[1]                              # repeat pad, dataset.py:519-522
        y = y[:, idx]
        if self.arch == "resnet":
            x = y.unsqueeze(1).transpose(2, 3)
            fwd = lambda: no.resnet_forward(self.sd, x, not self.scoring, update_running=not self.scoring)
        else:
            x = y.transpose(1, 2)
            fwd = lambda: no.ecapa_forward(self.sd, x, not self.scoring, update_running=not self.scoring)
        if self.scoring:
            with torch.no_grad():
                feat, _ = fwd()
                return no.ocsoftmax(self.center, feat, torch.zeros(self.B, dtype=torch.long))[1]
        feat, logits = fwd()
        loss, _ = no.ocsoftmax(self.center, feat, self.labels, 0.9, 0.2, 20.0)
        no.cross_entropy(logits.detach(), self.labels)
        for k in self.keys:
            self.sd[k].grad = None
        self.center.grad = None
        loss.backward()
        self.step += 1
        with torch.no_grad():
            for k in self.keys:
                if self.sd[k].grad is not None:
                    no.adam_l2_step(self.sd[k], self.sd[k].grad, self.m[k], self.v[k], self.step, 5e-4)
            no.sgd_step(self.center, self.center.grad, 5e-4)
        return loss.detach()


class _RefStep:
    """The reference ITSELF: the unmodified modules of oracle/_ref (copied there by oracle/build_ref.sh; or the mounted
    tree) imported under oracle/ref_shim and wired as main_train.py:162-175,338-409 wires them -- feature_extraction.LFCC
    -> repeat pad (dataset.py:519-522) -> transpose (main_train.py:338,347-348) -> resnet.ResNet / ecapa_tdnn.Res2Net2 ->
    loss.AngularIsoLoss -> torch.optim.Adam(weight_decay 5e-4) + SGD on the centre; scoring as generate_score.py:84-119."""
    kind = "reference"

    def __init__(self, arch, scoring, B):
        import warnings
        from oracle import ref_shim, state_spec as ss
        if not ref_shim.use_copy_if_needed():
            raise ImportError("no reference modules (oracle/_ref is built by oracle/build_ref.sh where /root/reference exists)")
        fe, ls = ref_shim.load("feature_extraction"), ref_shim.load("loss")
        torch.manual_seed(688)                                             # main_train.py:26
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.lfcc = fe.LFCC(320, 160, 512, 16000, 20, with_energy=False)      # dataset.py:13
            if arch == "resnet":
                self.model = ref_shim.load("resnet").ResNet(3, 256, resnet_type="18", nclasses=2)
            else:
                ec = ref_shim.load("ecapa_tdnn")
                self.model = ec.Res2Net2(ec.Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60)
        self.loss = ls.AngularIsoLoss(256, r_real=0.9, r_fake=0.2, alpha=20.0)
        self.arch, self.scoring, self.B = arch, scoring, B
        self.opt = torch.optim.Adam(self.model.parameters(), lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0005)
        self.opt_c = torch.optim.SGD(self.loss.parameters(), lr=5e-4)
        self.waves = ss.seeded_waves(B, WAVE_LEN, seed=0)
        self.labels = ss.seeded_labels(B, 0)
        self.model.eval() if scoring else self.model.train()

    def __call__(self):
        with torch.no_grad():
            y = self.lfcc(self.waves.clone())                              # the reference pre-emphasises in place
        idx = torch.arange(750) % y.shape[1]
        x = y[:, idx].unsqueeze(1).transpose(2, 3)
        if self.arch == "ecapa":
            x = x.squeeze(1)
        if self.scoring:
            with torch.no_grad():
                feats, _ = self.model(x)
                return self.loss(feats, torch.zeros(self.B, dtype=torch.long))[1]
        feats, outputs = self.model(x)
        torch.nn.functional.cross_entropy(outputs.detach(), self.labels)   # the logged CE, main_train.py:355-357
        loss, _ = self.loss(feats, self.labels)
        self.opt.zero_grad()
        self.opt_c.zero_grad()
        loss.backward()
        self.opt.step()
        self.opt_c.step()
        return loss.detach()


def _cpu_step(arch, scoring, B):
    """The reference itself when its modules are present (oracle/_ref), else the oracle port."""
    try:
        return _RefStep(arch, scoring, B)
    except ImportError:
        st = _CpuStep(arch, scoring, B)
        st.kind = "port"
        return st


def cpu_baseline(arch, scoring, seconds=15.0, B=8):
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    st = _cpu_step(arch, scoring, B)
    st()
    t0, it = time.perf_counter(), 0
    while it < 2 or (time.perf_counter() - t0 < seconds and it < 50):
        st()
        it += 1
    dt = time.perf_counter() - t0
    return {"value": B * it / dt, "unit": "utterances/s", "cores": n, "kind": st.kind,
            "sample": "%d %s steps of B=%d synthetic 4 s waves (%s, torch-CPU fp32: LFCC -> %s -> OC-Softmax%s)"
                      % (it, "scoring" if scoring else "train", B, _KIND_TEXT[st.kind], arch, "" if scoring else " -> Adam/SGD")}


_KIND_TEXT = {"reference": "the reference's own modules from oracle/_ref under oracle/ref_shim",
              "port": "oracle port of the reference"}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    arch = "resnet" if args.workload.startswith("resnet") else "ecapa"
    scoring = args.workload.endswith("_score")
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    B = 8
    st = _cpu_step(arch, scoring, B)
    for _ in range(min(args.warmup, 2)):
        st()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st()
    dt = time.perf_counter() - t0
    v = B * args.steps / dt
    return {"impl": "reference", "metric": "utterances/sec (4 s@16 kHz) %s" % ("scoring" if scoring else "train-step"),
            "value": v, "unit": "utterances/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, args.batch or (1024 if scoring else 256), world),
            "cpu_baseline": {"value": v, "unit": "utterances/s", "cores": n, "kind": st.kind,
                             "sample": "%d steps x B=%d utterances of the workload per step (%s; fp32, %d host threads)"
                                       % (args.steps, B, _KIND_TEXT[st.kind], n)},
            "e2e": {"value": v, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
