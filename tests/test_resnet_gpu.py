"""GPU parity of the ResNet-18-OC train / score step (LFCC features -> ResNet -> OC-Softmax ->
backward -> optimiser) through the drop-in modules.

Two references:
  * the bf16-point oracle (oracle/nets_oracle.py with bf16_points=True): fp32 arithmetic that rounds
    to bf16 exactly where the sm_100a path stores bf16;
  * the golden vectors produced by the UNMODIFIED reference modules in fp32 (tests/golden/
    nets_golden.npz).

Tolerances.  north_star asks for loss/logits within 1e-3 rel.  The LOSS meets that against both
references.  Embeddings / logits / deep gradients cannot: this 20-layer ReLU net amplifies rounding
noise (the reference's own fp32 arithmetic vs the same arithmetic with bf16 storage -- the two
ORACLES -- differ by 8e-3 on feat, 4e-2 on the logits and ~30 % on layer-1 gradients at B=4), so
for those the bar is "within the bf16 noise floor": the distance of the CUDA path to the bf16-point
oracle must not exceed the distance between the two oracles (x1.5 slack), with the floor computed
in the test.  Kernel-level exactness is pinned layer by layer in test_conv_gpu.py /
test_kernels_gpu.py, where no amplification is involved.
"""
import os

import numpy as np
import pytest
import torch

from oracle import lfcc_oracle as lo, nets_oracle as no, state_spec as ss

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().reshape(-1).cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _features(batch, seed):
    y = lo.lfcc(ss.seeded_waves(batch, 64000, seed=seed).numpy())
    y = lo.apply_frame_map(y, lo.frame_index_map(y.shape[1], 750, "repeat"))
    return torch.from_numpy(y).float()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "nets_golden.npz"))


@pytest.fixture(scope="module")
def run(gold):
    """One train-mode forward/backward of the drop-in ResNet + AngularIsoLoss on the golden inputs."""
    from asvspoof2021_air_b200.resnet import ResNet
    from asvspoof2021_air_b200.loss import AngularIsoLoss
    B, seed = int(gold["batch"]), int(gold["seed"])
    feats = _features(B, seed)
    x = feats.unsqueeze(1).transpose(2, 3).contiguous()             # (B,1,60,750)  main_train.py:338
    labels = torch.from_numpy(gold["labels"])
    spec = ss.resnet_spec()
    model = ResNet(3, 256, "18", nclasses=2).cuda()
    model.load_state_dict(ss.seeded_state(spec, 11))
    loss_mod = AngularIsoLoss(256, r_real=0.9, r_fake=0.2, alpha=20.0).cuda()
    with torch.no_grad():
        loss_mod.center.copy_(ss.seeded_center(256, 11))
    model.train()
    feat, logits = model(x.cuda())
    loss, score = loss_mod(feat, labels.cuda())
    loss.backward()
    torch.cuda.synchronize()
    # bf16-point oracle on the same inputs
    sd = ss.seeded_state(spec, 11)
    for k in ss.trainable_keys(spec):
        sd[k].requires_grad_(True)
    center = ss.seeded_center(256, 11).requires_grad_(True)
    ofeat, omu = no.resnet_forward(sd, x, True, bf16_points=True)
    oloss, oscore = no.ocsoftmax(center, ofeat, labels, 0.9, 0.2, 20.0)
    oloss.backward()
    # fp32 oracle (the reference arithmetic) on the same inputs: defines the bf16 noise floor
    sd32 = ss.seeded_state(spec, 11)
    for k in ss.trainable_keys(spec):
        sd32[k].requires_grad_(True)
    center32 = ss.seeded_center(256, 11).requires_grad_(True)
    ffeat, fmu = no.resnet_forward(sd32, x, True, bf16_points=False)
    floss, _ = no.ocsoftmax(center32, ffeat, labels, 0.9, 0.2, 20.0)
    floss.backward()
    f32 = dict(feat=ffeat.detach(), mu=fmu.detach(), loss=float(floss),
               grads={k: sd32[k].grad for k in ss.trainable_keys(spec)})
    return dict(f32=f32, model=model, loss_mod=loss_mod, feat=feat.detach().cpu(), logits=logits.detach().cpu(),
                loss=float(loss), score=score.detach().cpu(), x=x, labels=labels,
                o=dict(feat=ofeat.detach(), mu=omu.detach(), loss=float(oloss), score=oscore.detach(),
                       grads={k: sd[k].grad for k in ss.trainable_keys(spec)}, cgrad=center.grad))


def test_state_dict_keys_match_reference(run):
    spec = ss.resnet_spec()
    sd = run["model"].state_dict()
    assert list(sd.keys()) == [k for k, _, _ in spec]
    for k, shape, _ in spec:
        assert tuple(sd[k].shape) == tuple(shape), k


def test_forward_matches_bf16_point_oracle(run):
    o = run["o"]
    f = run["f32"]
    assert abs(run["loss"] - o["loss"]) <= 1e-3 * abs(o["loss"]), (run["loss"], o["loss"])
    assert abs(run["loss"] - f["loss"]) <= 1e-3 * abs(f["loss"]), (run["loss"], f["loss"])
    floor_feat, floor_mu = _rel(o["feat"], f["feat"]), _rel(o["mu"], f["mu"])
    print("bf16 noise floor (oracle fp32 vs oracle bf16 points): feat %.2e logits %.2e; ours vs bf16 oracle: %.2e %.2e"
          % (floor_feat, floor_mu, _rel(run["feat"], o["feat"]), _rel(run["logits"], o["mu"])))
    assert _rel(run["feat"], o["feat"]) <= max(1.5 * floor_feat, 1e-3)
    assert _rel(run["logits"], o["mu"]) <= max(1.5 * floor_mu, 1e-3)
    assert float((run["score"] - o["score"]).abs().max()) <= 1e-3


def test_forward_vs_reference_golden_fp32(run, gold):
    # bf16 activations (8-bit mantissa) through 20 conv layers vs the fp32 reference
    assert abs(run["loss"] - float(gold["resnet_loss"])) <= 1e-3 * abs(float(gold["resnet_loss"]))
    assert _rel(run["feat"], gold["resnet_feat"]) <= 2e-2
    assert _rel(run["logits"], gold["resnet_logits"]) <= 8e-2      # 2 logits/utt = small differences of feat
    assert float(np.abs(run["score"].numpy() - gold["resnet_score"]).max()) <= 2e-3


def test_gradients_match_bf16_point_oracle(run):
    model, o, f = run["model"], run["o"], run["f32"]
    ours, floor = [], []
    for k, p in model.named_parameters():
        g = o["grads"].get(k)
        if g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k      # fc_mu.* get no gradient
            continue
        assert p.grad is not None, k
        ours.append((_rel(p.grad, g), k))
        floor.append((_rel(g, f["grads"][k]), k))
        # every gradient within the bf16 noise floor of its own tensor (x1.5) or 2e-2
        assert ours[-1][0] <= max(1.5 * floor[-1][0], 2e-2), (k, ours[-1][0], floor[-1][0])
    ours.sort(reverse=True)
    floor.sort(reverse=True)
    print("worst grads (ours vs bf16 oracle):", ours[:4], "noise floor:", floor[:4])
    assert np.median([w[0] for w in ours]) <= np.median([w[0] for w in floor])
    # the shallow end of the backward pass is not amplified: tight tolerances
    named = dict(model.named_parameters())
    assert _rel(named["fc.bias"].grad, o["grads"]["fc.bias"]) <= 2e-3
    assert _rel(named["fc.weight"].grad, o["grads"]["fc.weight"]) <= 1e-2
    assert _rel(run["loss_mod"].center.grad, o["cgrad"]) <= 1e-2


def test_gradient_norms_vs_reference_golden(run, gold):
    keys = [str(k) for k in gold["resnet_grad_keys"]]
    norms = dict(zip(keys, gold["resnet_grad_norm"]))
    named = dict(run["model"].named_parameters())
    assert set(keys) == {k for k, p in named.items() if p.grad is not None and float(p.grad.abs().max()) > 0}
    o = run["o"]
    for k in keys:
        n = float(named[k].grad.double().norm())
        # bf16 noise floor of this norm: how far the bf16-point ORACLE (same arithmetic as the reference, bf16 storage
        # where the CUDA path stores bf16) lands from the fp32 reference; layer-1 BatchNorm gradients sit at 6-9 %
        floor = abs(float(o["grads"][k].double().norm()) - norms[k]) / norms[k]
        assert abs(n - norms[k]) <= max(8e-2, 1.5 * floor) * norms[k] + 1e-7, (k, n, norms[k], floor)


def test_running_stats_after_one_step(run, gold):
    sd = run["model"].state_dict()
    for k, s in zip(gold["resnet_running_keys"], gold["resnet_running_sum"]):
        got = float(sd[str(k)].double().sum())
        assert abs(got - s) <= 2e-2 * abs(s) + 1e-3, (k, got, s)
    assert int(sd["bn1.num_batches_tracked"]) == 1


def test_eval_mode_scores_vs_reference_golden(run, gold):
    model, loss_mod = run["model"], run["loss_mod"]
    model.eval()
    with torch.no_grad():
        feat, logits = model(run["x"].cuda())
        _, score = loss_mod(feat, torch.zeros(feat.shape[0], device="cuda"))
    model.train()
    assert _rel(feat, gold["resnet_eval_feat"]) <= 3e-2
    assert float(np.abs((-score).cpu().numpy() - gold["resnet_eval_score"]).max()) <= 3e-2


def test_trainer_step_from_raw_waves_matches_module_path():
    """Trainer.train_step (wave -> fused LFCC -> engine -> loss -> Adam/SGD) == the same step done
    with the oracle on CPU: loss of step 1 and of step 2 (i.e. after one optimiser update)."""
    from asvspoof2021_air_b200.trainer import Trainer
    B = 4
    waves = ss.seeded_waves(B, 64000, seed=3)
    labels = ss.seeded_labels(B, 3)
    tr = Trainer(arch="resnet", seed=5)
    spec = ss.resnet_spec()
    sd = ss.seeded_state(spec, 11)
    tr.load_state(sd, ss.seeded_center(256, 11))
    l1 = float(tr.train_step(waves.cuda(), labels.cuda()))
    l2 = float(tr.train_step(waves.cuda(), labels.cuda()))
    # oracle: same two steps in fp32 with bf16 rounding points
    x = _features(B, 3).unsqueeze(1).transpose(2, 3).contiguous()
    keys = [k for k in ss.trainable_keys(spec) if not k.startswith("fc_mu.")]
    for k in keys:
        sd[k].requires_grad_(True)
    center = ss.seeded_center(256, 11).requires_grad_(True)
    m = {k: torch.zeros_like(sd[k]) for k in keys}
    v = {k: torch.zeros_like(sd[k]) for k in keys}
    losses = []
    for step in (1, 2):
        feat, _ = no.resnet_forward(sd, x, True, bf16_points=True, update_running=True)
        loss, _ = no.ocsoftmax(center, feat, labels, 0.9, 0.2, 20.0)
        for k in keys:
            sd[k].grad = None
        center.grad = None
        loss.backward()
        losses.append(float(loss))
        with torch.no_grad():
            for k in keys:
                no.adam_l2_step(sd[k], sd[k].grad, m[k], v[k], step, 5e-4)
            no.sgd_step(center, center.grad, 5e-4)
    assert abs(l1 - losses[0]) <= 1e-3 * abs(losses[0]), (l1, losses[0])
    assert abs(l2 - losses[1]) <= 2e-2 * abs(losses[1]), (l2, losses[1])
    assert l2 != l1


def test_train_step_is_reproducible_run_to_run():
    """Two identical forward/backward passes: the forward is bit-identical (no floating-point atomics with a free
    order on the forward path -- the fused BatchNorm statistics accumulate in registers in a fixed order) and the
    gradients agree to fp32 atomic-order noise.  The data-parallel check (scripts/ddp_check.py) relies on it."""
    from asvspoof2021_air_b200 import ops
    from asvspoof2021_air_b200.trainer import Trainer
    B = 8
    tr = Trainer(arch="resnet", seed=100)
    st = tr.engine.store
    w, lab = ss.seeded_waves(B, 64000, seed=50).cuda(), ss.seeded_labels(B, 0).cuda()
    runs = []
    for _ in range(2):
        x0 = tr.features(w)
        feat, logits = tr.engine.forward(x0, training=True)
        dfeat, score = torch.empty(B, feat.shape[1], device="cuda"), torch.empty(B, device="cuda")
        tr.engine.zero_grad(); tr.center_grad.zero_()
        ops.ocsoftmax(feat, lab, tr.center, B, feat.shape[1], tr.r_real, tr.r_fake, tr.alpha, 1.0, tr.loss, score, dfeat,
                      tr.center_grad, logits, logits.shape[1], tr.ce)
        tr.engine.backward(dfeat)
        torch.cuda.synchronize()
        runs.append((feat.clone(), st.grads[:st.n_train].clone()))
    assert torch.equal(runs[0][0], runs[1][0])
    for name, (off, n, _) in st.offsets.items():
        if off >= st.n_train:
            continue
        a, b = runs[0][1][off:off + n].double(), runs[1][1][off:off + n].double()
        assert float((a - b).norm()) <= 1e-5 * float(a.norm()) + 1e-12, name
