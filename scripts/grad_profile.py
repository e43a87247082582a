"""Per-tensor gradient deviation of one train-mode forward/backward against the fp32 oracle's autograd (CPU), in layer
order -- localises where a backward pass leaves the reference's arithmetic.

    python scripts/grad_profile.py resnet fp32 [B]        (prints one line per parameter tensor)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lfcc_torch, nets_oracle as no, state_spec as ss  # noqa: E402
from asvspoof2021_air_b200 import ops  # noqa: E402
from asvspoof2021_air_b200.trainer import Trainer  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet"
precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
spec = ss.resnet_spec() if arch == "resnet" else ss.ecapa_spec()
waves, labels = ss.seeded_waves(B, 64000, seed=3), ss.seeded_labels(B, 3)

tr = Trainer(arch=arch, seed=5, precision=precision)
tr.load_state(ss.seeded_state(spec, 11), ss.seeded_center(256, 11))
eng = tr.engine
x0 = tr.features(waves.cuda())
feat, logits = eng.forward(x0, training=True)
dfeat, score = torch.empty_like(feat), torch.empty(B, device="cuda")
eng.zero_grad()
tr.center_grad.zero_()
ops.ocsoftmax(feat, labels.cuda(), tr.center, B, 256, 0.9, 0.2, 20.0, 1.0, tr.loss, score, dfeat, tr.center_grad, logits,
              logits.shape[1], tr.ce)
eng.backward(dfeat)
torch.cuda.synchronize()

y = lfcc_torch.TorchLFCC()(waves)
y = y[:, torch.arange(750) % y.shape[1]]
x = y.unsqueeze(1).transpose(2, 3).contiguous() if arch == "resnet" else y.transpose(1, 2).contiguous()
sd = ss.seeded_state(spec, 11)
keys = ss.trainable_keys(spec)
for k in keys:
    sd[k].requires_grad_(True)
center = ss.seeded_center(256, 11).requires_grad_(True)
fwd = no.resnet_forward if arch == "resnet" else no.ecapa_forward
ofeat, ologits = fwd(sd, x, True)
oloss, _ = no.ocsoftmax(center, ofeat, labels, 0.9, 0.2, 20.0)
oloss.backward()
print("# %s %s B=%d: loss %.7f (oracle %.7f), feat rel %.2e" % (arch, precision, B, float(tr.loss), float(oloss),
      float((feat.cpu().double() - ofeat.detach().double()).norm() / ofeat.detach().double().norm())))
print("# %-40s %10s %10s %12s" % ("parameter", "rel err", "norm dev", "ref norm"))
for k in keys:
    g = sd[k].grad
    if g is None:
        continue
    ours = eng.store.pt_view(k, eng.store.grads).detach().cpu().double()
    g = g.double()
    rel = float((ours - g).norm() / (g.norm() + 1e-30))
    nd = float(abs(ours.norm() - g.norm()) / (g.norm() + 1e-30))
    print("%-42s %10.2e %10.2e %12.4e" % (k, rel, nd, float(g.norm())))
