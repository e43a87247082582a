"""Run the same forward/backward twice on one GPU and report, per parameter, how much the gradients differ."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvspoof2021_air_b200 import ops
from asvspoof2021_air_b200.trainer import Trainer
from asvspoof2021_air_b200.bench_train import _waves, _labels

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
tr = Trainer(arch=arch, seed=100)
st = tr.engine.store
w, lab = _waves(B, 50).cuda(), _labels(B, 0).cuda()
grads = []
for rep in range(3):
    x0 = tr.features(w)
    feat, logits = tr.engine.forward(x0, training=True)
    tr.dfeat = torch.empty(B, feat.shape[1], device="cuda"); tr.score = torch.empty(B, device="cuda")
    tr.engine.zero_grad(); tr.center_grad.zero_()
    ops.ocsoftmax(feat, lab, tr.center, B, feat.shape[1], tr.r_real, tr.r_fake, tr.alpha, 1.0, tr.loss, tr.score, tr.dfeat,
                  tr.center_grad, logits, logits.shape[1], tr.ce)
    tr.engine.backward(tr.dfeat)
    torch.cuda.synchronize()
    grads.append((st.grads[:st.n_train].clone(), feat.clone(), float(tr.loss)))
print("loss", [g[2] for g in grads])
print("feat diff", float((grads[0][1] - grads[1][1]).abs().max()), float((grads[1][1] - grads[2][1]).abs().max()))
for name, (off, n, shape) in st.offsets.items():
    if off >= st.n_train:
        continue
    a, b, c = (g[0][off:off + n].double() for g in grads)
    d01 = float((a - b).norm() / (a.norm() + 1e-30)); d12 = float((b - c).norm() / (b.norm() + 1e-30))
    if max(d01, d12) > 1e-5:
        print("%-32s rel diff run0-1 %.3e  run1-2 %.3e" % (name, d01, d12))
print("done")
