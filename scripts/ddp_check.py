"""Data-parallel check (run under torchrun, N >= 2 GPUs): the all-reduced gradient equals the mean of the
per-replica gradients (BatchNorm statistics stay per replica), replicas start from
identical parameters and stay bit-identical after optimiser steps."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from asvspoof2021_air_b200 import parallel                      # noqa: E402
from asvspoof2021_air_b200.trainer import Trainer               # noqa: E402
from asvspoof2021_air_b200.bench_train import _waves, _labels   # noqa: E402


def main(arch="resnet", B=8):
    rank, world, local = parallel.init_from_env()
    assert world >= 2
    tr = Trainer(arch=arch, process_group=dist.group.WORLD, seed=100 + rank)     # different seeds: broadcast must fix it
    st = tr.engine.store
    p0 = st.params.clone()
    gathered = [torch.empty_like(p0) for _ in range(world)]
    dist.all_gather(gathered, p0)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas do not start from identical parameters"
    w, lab = _waves(B, 50 + rank).cuda(), _labels(B, rank).cuda()
    # local gradient (no exchange): same trainer, hook disabled
    tr.engine.grad_hook = None
    red, tr.reducer = tr.reducer, None
    x0 = tr.features(w)
    feat, logits = tr.engine.forward(x0, training=True)
    tr.dfeat = torch.empty(B, feat.shape[1], device="cuda"); tr.score = torch.empty(B, device="cuda")
    tr.engine.zero_grad(); tr.center_grad.zero_()
    from asvspoof2021_air_b200 import ops
    ops.ocsoftmax(feat, lab, tr.center, B, feat.shape[1], tr.r_real, tr.r_fake, tr.alpha, 1.0, tr.loss, tr.score, tr.dfeat,
                  tr.center_grad, logits, logits.shape[1], tr.ce)
    tr.engine.backward(tr.dfeat)
    local_g = st.grads[:st.n_train].clone()
    all_g = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(all_g, local_g)
    mean_g = torch.stack(all_g).double().mean(0)
    # reduced gradient through the bucketed reducer (running stats were touched twice; irrelevant here)
    tr.reducer = red
    tr.engine.grad_hook = red.ready
    red.begin()
    tr.engine.zero_grad()
    tr.engine.forward(x0, training=True)
    tr.engine.backward(tr.dfeat)
    scale = red.finish(tr.center_grad)
    torch.cuda.synchronize()
    got = st.grads[:st.n_train].double() * scale
    err = float((got - mean_g).norm() / (mean_g.norm() + 1e-30))
    assert err < 1e-6, err
    # two full steps: replicas stay bit-identical
    for i in range(2):
        tr.train_step(w, lab)
    dist.all_gather(gathered, st.params)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "replicas diverged"
    assert not torch.equal(st.params, p0)
    if rank == 0:
        print("ddp_check ok: world %d arch %s reduce rel err %.2e" % (world, arch, err))
    dist.destroy_process_group()


if __name__ == "__main__":
    main(*(sys.argv[1:2] or ["resnet"]))
