"""GPU checks of the adversarial channel-classifier head (asvspoof2021_air_b200/adv.py, csrc/adv.cu) against the golden
vectors of the reference module (tests/golden/adv_golden.npz) and the numpy oracle.  Run as a script by
tests/test_adv_gpu.py IN A SUBPROCESS: these kernels have not been on hardware yet, and a fault must not take the CUDA
context of the rest of the GPU suite with it.  Prints `adv checks ok` on success."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from asvspoof2021_air_b200.adv import ChannelClassifier  # noqa: E402
from oracle import adv_oracle as ao  # noqa: E402


def close(a, b, rtol=1e-4, atol=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() <= atol + rtol * np.abs(b).max()


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "adv_golden.npz"))
    lam = float(g["lambda"])
    x = torch.from_numpy(g["x"]).cuda()
    for C in (60, 13):
        p = "c%d_" % C
        clf = ChannelClassifier(256, C, lam, device="cuda")
        sd = {"classifier.0.weight": g[p + "w1"], "classifier.0.bias": g[p + "b1"],
              "classifier.3.weight": g[p + "w2"], "classifier.3.bias": g[p + "b2"]}
        clf.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        labels = torch.from_numpy(g[p + "labels"]).cuda()
        keep = torch.from_numpy(g[p + "keep"]).cuda()
        # step 1: loss, accuracy count, gradient-reversed feature gradient (added into dfeat)
        dfeat = torch.zeros(16, 256, device="cuda")
        loss, correct = clf.head_loss_and_feat_grad(x, labels, dfeat, keep_mask=keep)
        assert abs(float(loss) - float(g[p + "loss"])) < 1e-5, (float(loss), float(g[p + "loss"]))
        assert int(correct) == int((g[p + "pred"] == g[p + "labels"]).sum())
        assert close(dfeat.cpu().numpy(), g[p + "dfeat"]), "dfeat"
        base = torch.ones(16, 256, device="cuda")
        clf.head_loss_and_feat_grad(x, labels, base, keep_mask=keep)
        assert close((base - 1).cpu().numpy(), g[p + "dfeat"], rtol=1e-3, atol=1e-6), "dfeat accumulates"
        # step 2: parameter gradients and one Adam(L2) step
        before = {k: v.clone() for k, v in clf.state_dict().items()}
        loss2, _ = clf.classifier_step(x, labels, lr=1e-4, keep_mask=keep)
        assert abs(float(loss2) - float(g[p + "loss"])) < 1e-5
        grads = {"classifier.0.weight": g[p + "dw1"], "classifier.0.bias": g[p + "db1"],
                 "classifier.3.weight": g[p + "dw2"], "classifier.3.bias": g[p + "db2"]}
        for k, ref in grads.items():
            assert close(clf._gviews[k].cpu().numpy(), ref), k
            gg = ref.astype(np.float64) + 0.0005 * before[k].cpu().numpy().astype(np.float64)      # coupled L2
            m, v = 0.1 * gg, 0.001 * gg * gg
            want = before[k].cpu().numpy() - 1e-4 * (m / 0.1) / (np.sqrt(v / 0.001) + 1e-8)
            assert close(clf.state_dict()[k].cpu().numpy(), want, rtol=1e-6, atol=1e-7), ("adam", k)
        # eval-style logits (no dropout): oracle with an all-ones mask and p = 0
        r = ao.forward_backward(g["x"], g[p + "labels"], *(clf.state_dict()[k].cpu().numpy() for k in sd), np.ones((16, 128)), lam, p=0.0)
        assert close(clf(x).cpu().numpy(), r["logits"], rtol=1e-5, atol=1e-5), "eval logits"
    # generated masks: Bernoulli(0.7), reproducible per seed, different across seeds
    clf = ChannelClassifier(256, 13, lam, device="cuda")
    xs = torch.randn(256, 256, device="cuda")
    lab = torch.randint(0, 13, (256,), device="cuda")
    d = torch.zeros(256, 256, device="cuda")
    clf.head_loss_and_feat_grad(xs, lab, d, seed=5)
    k5 = clf._work[256]["keep"].clone()
    clf.head_loss_and_feat_grad(xs, lab, d, seed=5)
    assert torch.equal(k5, clf._work[256]["keep"])
    clf.head_loss_and_feat_grad(xs, lab, d, seed=6)
    k6 = clf._work[256]["keep"]
    assert 0.68 < float(k5.float().mean()) < 0.72 and 0.3 < float((k5 != k6).float().mean()) < 0.55
    # against the oracle with the mask the kernel drew
    d.zero_()
    loss, correct = clf.head_loss_and_feat_grad(xs, lab, d, seed=5)
    r = ao.forward_backward(xs.cpu().numpy(), lab.cpu().numpy(), *(clf.state_dict()[k].cpu().numpy() for k in
                            ("classifier.0.weight", "classifier.0.bias", "classifier.3.weight", "classifier.3.bias")),
                            k5.cpu().numpy(), lam)
    assert abs(float(loss) - r["loss"]) < 1e-4 and close(d.cpu().numpy(), r["dfeat"], rtol=1e-3)
    assert int(correct) == int((r["pred"] == lab.cpu().numpy()).sum())
    # inside a train step: the reversed channel gradient reaches the encoder, the classifiers move, nothing is NaN
    from asvspoof2021_air_b200.trainer import Trainer
    from asvspoof2021_air_b200 import data
    waves, _, labels, _, _ = data.SyntheticWaves(4, length=32000, seed=2).batch([0, 1, 2, 3])
    grads = []
    for adv in (False, True):
        tr = Trainer(arch="resnet", device="cuda", seed=3)
        heads = tr.attach_adversaries([5, 3], lambda_=0.5, lr_d=1e-3, seed=4)
        before = [h.flat.clone() for h in heads]
        ch = torch.tensor([[0, 1], [4, 2], [2, 0], [1, 1]], device="cuda")
        l = tr.train_step(waves.cuda(), labels.cuda(), channels=ch if adv else None)
        assert float(l) == float(l)
        grads.append(tr.engine.store.grads.clone())
        moved = [not torch.equal(b, h.flat) for b, h in zip(before, heads)]
        assert moved == [adv, adv], moved
        if adv:
            assert all(float(s[0]) == float(s[0]) and 0 <= int(s[1]) <= 4 for s in tr.adv_stats)
    assert not torch.equal(grads[0], grads[1]) and bool(torch.isfinite(grads[1]).all())
    print("adv checks ok")


if __name__ == "__main__":
    main()
