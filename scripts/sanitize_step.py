"""One small train step (+ one scoring step) of a net, for compute-sanitizer (sanitizer row):

    compute-sanitizer --tool memcheck  python scripts/sanitize_step.py resnet
    compute-sanitizer --tool racecheck python scripts/sanitize_step.py ecapa

B = 2 keeps the instrumented run short; every kernel family of the step (LFCC, stem, tcgen05 convs incl. the TMA patch
kernels, BatchNorm, pooling, loss, optimiser) is launched.  Prints the loss so that a silent no-op cannot pass."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvspoof2021_air_b200.trainer import Trainer  # noqa: E402
from asvspoof2021_air_b200.bench_train import _waves, _labels  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
kw = {} if precision == "bf16" else {"precision": precision}
tr = Trainer(arch=arch, seed=688, **kw)
w, lab = _waves(B, 0).cuda(), _labels(B, 0).cuda()
loss = float(tr.train_step(w, lab))
score = tr.score_step(w)
torch.cuda.synchronize()
print("sanitize_step %s B=%d %s: loss %.6f, scores %s" % (arch, B, precision, loss, [round(float(s), 4) for s in score]))
