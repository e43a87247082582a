"""Per-launch device times of one train step (default) or one scoring step (third argument `score`): CUDA events around
every C-ABI call.   python scripts/prof_step.py [B] [resnet|ecapa] [score]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvspoof2021_air_b200 import ops
from asvspoof2021_air_b200.trainer import Trainer
from asvspoof2021_air_b200.bench_train import _waves, _labels

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
arch = sys.argv[2] if len(sys.argv) > 2 else "resnet"
score = len(sys.argv) > 3 and sys.argv[3] == "score"
tr = Trainer(arch=arch, seed=688)
w, lab = _waves(B, 0).cuda(), _labels(B, 0).cuda()
step = (lambda: tr.score_step(w)) if score else (lambda: tr.train_step(w, lab))
for _ in range(2):
    step()
torch.cuda.synchronize()
with ops.Profile() as prof:
    step()
tot = 0.0
for fam, ms, fl, d in prof.per_call():
    tot += ms
    print("%-12s %8.3f ms %8.1f TF/s  %s" % (fam, ms, fl / ms / 1e9 if fl else 0.0, d))
print("total %.3f ms" % tot)
