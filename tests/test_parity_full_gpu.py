"""GPU parity AT THE BENCHMARK CONFIGURATIONS (BASELINE.json configs 2-5): Trainer.train_step at B = 256 for ResNet-OC
and ECAPA-TDNN and Trainer.score_step at B = 1024, against the fp32 oracle (oracle/nets_oracle.py, pinned on the
reference golden) run on this box's CPU on the same seeded waves and weights.

Two precisions of the same kernels:
  * "bf16"  the product path.  Asserted: loss within 1e-3; feat / logits / scores within 1.5 x the bf16 noise floor (the
            distance between the fp32 oracle and the SAME oracle with bf16 storage points, computed here at full size).
  * "fp32"  the parity mode.  Asserted: loss, feat, logits and scores within 1e-3 -- the north-star
            tolerance -- at full size.
Every measured deviation is appended to gpurun_out/parity_full.json (copied to profiles/r02_parity.json).
"""
import json
import os
import time

import pytest
import torch

from oracle import lfcc_torch, nets_oracle as no, state_spec as ss

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3
_CPU = {}


def _maxrel(a, b):
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _rel(a, b):
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "parity_full.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data[key] = value
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _inputs(arch, B, seed):
    spec = ss.resnet_spec() if arch == "resnet" else ss.ecapa_spec()
    return spec, ss.seeded_waves(B, 64000, seed=seed), ss.seeded_labels(B, seed)


def _cpu_forward(arch, B, seed, training):
    """fp32 oracle and bf16-storage oracle forward on the CPU (no autograd: forward quantities only)."""
    key = (arch, B, seed, training)
    if key in _CPU:
        return _CPU[key]
    torch.set_num_threads(os.cpu_count() or 1)
    spec, waves, labels = _inputs(arch, B, seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        y = lfcc_torch.TorchLFCC()(waves)                                 # the reference's LFCC arithmetic (fp32)
        y = y[:, torch.arange(750) % y.shape[1]]                          # repeat pad, dataset.py:519-522
        x = y.unsqueeze(1).transpose(2, 3).contiguous() if arch == "resnet" else y.transpose(1, 2).contiguous()
        fwd = no.resnet_forward if arch == "resnet" else no.ecapa_forward
        out = {}
        for name, bf in (("f32", False), ("bf16pt", True)):
            sd = ss.seeded_state(spec, 11)
            feat, logits = fwd(sd, x, training, bf16_points=bf)
            loss, score = no.ocsoftmax(ss.seeded_center(256, 11), feat, labels, 0.9, 0.2, 20.0)
            out[name] = dict(feat=feat, logits=logits, loss=float(loss), score=score)
    out["cpu_s"] = time.perf_counter() - t0
    _CPU[key] = out
    return out


def _devs(got, ref):
    return {"loss": abs(got["loss"] - ref["loss"]) / abs(ref["loss"]) if "loss" in got else None,
            "feat_max": _maxrel(got["feat"], ref["feat"]), "feat_norm": _rel(got["feat"], ref["feat"]),
            "logits_max": _maxrel(got["logits"], ref["logits"]), "logits_norm": _rel(got["logits"], ref["logits"]),
            "score_abs": float((got["score"].double().cpu() - ref["score"].double()).abs().max())}


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("arch", ["resnet", "ecapa"])
def test_train_step_at_batch_256(arch, precision):
    from asvspoof2021_air_b200.trainer import Trainer
    B, seed = 256, 21
    spec, waves, labels = _inputs(arch, B, seed)
    tr = Trainer(arch=arch, seed=5, precision=precision)
    tr.load_state(ss.seeded_state(spec, 11), ss.seeded_center(256, 11))
    loss = float(tr.train_step(waves.cuda(), labels.cuda()))
    eng = tr.engine
    logits = eng.mu if arch == "resnet" else eng.logits
    got = dict(loss=loss, feat=eng.feat.clone().cpu(), logits=logits.clone().cpu(), score=tr.score.clone().cpu())
    loss2 = float(tr.train_step(waves.cuda(), labels.cuda()))
    assert loss2 == loss2 and loss2 != loss                               # the optimiser step took effect
    del tr
    torch.cuda.empty_cache()
    cpu = _cpu_forward(arch, B, seed, True)
    dev, floor = _devs(got, cpu["f32"]), _devs(cpu["bf16pt"], cpu["f32"])
    _record("train_B256_%s_%s" % (arch, precision), {"vs_fp32_oracle": dev, "bf16_noise_floor": floor, "loss_gpu": loss,
                                                    "loss_oracle": cpu["f32"]["loss"], "oracle_cpu_s": round(cpu["cpu_s"], 1)})
    print(arch, precision, "B=256 vs fp32 oracle:", {k: "%.2e" % v for k, v in dev.items()},
          "| bf16 floor:", {k: "%.2e" % v for k, v in floor.items()})
    assert dev["loss"] <= TOL, dev
    if precision == "fp32":
        for k in ("feat_max", "logits_max", "score_abs"):
            assert dev[k] <= TOL, (k, dev)
    else:
        for k in ("feat_norm", "logits_norm", "score_abs"):
            assert dev[k] <= max(1.5 * floor[k], TOL), (k, dev[k], floor[k])


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_ecapa_scoring_at_batch_1024(precision):
    """BASELINE.json config 5: generate_score.py inference, ECAPA-512, batch 1024, scores vs reference within 1e-3."""
    from asvspoof2021_air_b200.trainer import Trainer
    arch, B, seed = "ecapa", 1024, 33
    spec, waves, labels = _inputs(arch, B, seed)
    sd = ss.seeded_state(spec, 11)
    tr = Trainer(arch=arch, seed=5, precision=precision)
    tr.load_state(sd, ss.seeded_center(256, 11))
    score = tr.score_step(waves.cuda())                                   # +cos, generate_score.py:117
    got = dict(feat=tr.engine.feat.clone().cpu(), logits=tr.engine.logits.clone().cpu(), score=(-score).clone().cpu())
    del tr
    torch.cuda.empty_cache()
    cpu = _cpu_forward(arch, B, seed, False)
    dev, floor = _devs(got, cpu["f32"]), _devs(cpu["bf16pt"], cpu["f32"])
    dev.pop("loss"), floor.pop("loss")
    _record("score_B1024_ecapa_%s" % precision, {"vs_fp32_oracle": dev, "bf16_noise_floor": floor,
                                                 "oracle_cpu_s": round(cpu["cpu_s"], 1)})
    print("ecapa scoring", precision, "B=1024 vs fp32 oracle:", {k: "%.2e" % v for k, v in dev.items()},
          "| bf16 floor:", {k: "%.2e" % v for k, v in floor.items()})
    if precision == "fp32":
        for k in ("feat_max", "logits_max", "score_abs"):
            assert dev[k] <= TOL, (k, dev)
    else:
        for k in ("feat_norm", "score_abs"):
            assert dev[k] <= max(1.5 * floor[k], 3e-3), (k, dev[k], floor[k])
