"""GPU diagnostic: ResNet engine (bf16 tensor-core path) vs the fp32 oracle and the bf16-point oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import lfcc_oracle as lo, nets_oracle as no, state_spec as ss   # noqa: E402
from asvspoof2021_air_b200.resnet import ResNet                               # noqa: E402
from asvspoof2021_air_b200 import ops                                        # noqa: E402


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def main(B=4):
    torch.set_num_threads(os.cpu_count() or 1)
    feats = lo.apply_frame_map(lo.lfcc(ss.seeded_waves(B, 64000, seed=3).numpy()), lo.frame_index_map(401, 750, "repeat"))
    x = torch.from_numpy(feats).float().unsqueeze(1).transpose(2, 3).contiguous()       # (B,1,60,750)
    labels = ss.seeded_labels(B, 3)
    spec = ss.resnet_spec()
    results = {}
    for name, bfp in (("fp32", False), ("bf16pt", True)):
        sd = ss.seeded_state(spec, 11)
        for k in ss.trainable_keys(spec):
            sd[k].requires_grad_(True)
        center = ss.seeded_center(256, 11).requires_grad_(True)
        feat, mu = no.resnet_forward(sd, x, True, bf16_points=bfp)
        loss, score = no.ocsoftmax(center, feat, labels, 0.9, 0.2, 20.0)
        loss.backward()
        results[name] = dict(feat=feat.detach(), mu=mu.detach(), loss=float(loss), score=score.detach(),
                             grads={k: sd[k].grad for k in ss.trainable_keys(spec) if sd[k].grad is not None},
                             cgrad=center.grad)
    m = ResNet(3, 256, "18", nclasses=2).cuda()
    m.load_state_dict(ss.seeded_state(spec, 11))
    m.train()
    eng = m.engine
    xb = x[:, 0].to(torch.bfloat16).contiguous().cuda()
    feat, mu = eng.forward(xb, training=True)
    cen = ss.seeded_center(256, 11).cuda()
    lossb = torch.zeros(1, device="cuda"); score = torch.zeros(B, device="cuda"); dfeat = torch.zeros(B, 256, device="cuda")
    cgrad = torch.zeros(1, 256, device="cuda")
    ops.ocsoftmax(feat, labels.cuda(), cen, B, 256, 0.9, 0.2, 20.0, 1.0, lossb, score, dfeat, cgrad)
    eng.zero_grad()
    eng.backward(dfeat)
    torch.cuda.synchronize()
    for name in ("fp32", "bf16pt"):
        r = results[name]
        print("== vs %s oracle: loss %.6f (ours %.6f) feat rel %.3e mu rel %.3e score maxabs %.3e cgrad rel %.3e" % (
            name, r["loss"], float(lossb), rel(feat.cpu(), r["feat"]), rel(mu.cpu(), r["mu"]),
            float((score.cpu() - r["score"]).abs().max()), rel(cgrad.cpu(), r["cgrad"])))
        worst = []
        for k, g in r["grads"].items():
            ours = eng.store.pt_view(k, eng.store.grads).cpu()
            worst.append((rel(ours, g), k, float(g.norm())))
        worst.sort(reverse=True)
        print("   grads: median rel %.3e; worst: %s" % (np.median([w[0] for w in worst]), [(k, "%.2e" % e, "%.2e" % n) for e, k, n in worst[:6]]))
        print("   best: %s" % [(k, "%.2e" % e) for e, k, n in worst[-4:]])
    # intermediate activations vs bf16-point oracle would need hooks; print a few stats instead
    print("feat ours[0,:4]", feat[0, :4].tolist(), "oracle", results["fp32"]["feat"][0, :4].tolist())
    print("fp32-vs-bf16pt oracle: feat rel %.3e, loss %.6f vs %.6f" % (rel(results["bf16pt"]["feat"], results["fp32"]["feat"]),
                                                                       results["bf16pt"]["loss"], results["fp32"]["loss"]))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
