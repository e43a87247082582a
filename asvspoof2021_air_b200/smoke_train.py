"""smoke(): one tiny train step of the flagship path on cuda:0 (wave -> fused LFCC -> ResNet-18-OC
fwd/bwd -> OC-Softmax -> Adam/SGD), checked against the CPU oracle with the same bf16 rounding points.
The oracle is the checker here (test infrastructure), never the thing that runs the step."""
import torch


def run(B=2):
    from oracle import lfcc_oracle as lo, nets_oracle as no, state_spec as ss
    from .trainer import Trainer
    waves = ss.seeded_waves(B, 64000, seed=3)
    labels = ss.seeded_labels(B, 3)
    spec = ss.resnet_spec()
    tr = Trainer(arch="resnet", seed=5)
    tr.load_state(ss.seeded_state(spec, 11), ss.seeded_center(256, 11))
    loss = float(tr.train_step(waves.cuda(), labels.cuda()))
    loss2 = float(tr.train_step(waves.cuda(), labels.cuda()))
    y = lo.apply_frame_map(lo.lfcc(waves.numpy()), lo.frame_index_map(401, 750, "repeat"))
    x = torch.from_numpy(y).float().unsqueeze(1).transpose(2, 3).contiguous()
    feat, _ = no.resnet_forward(ss.seeded_state(spec, 11), x, True, bf16_points=True)
    want, _ = no.ocsoftmax(ss.seeded_center(256, 11), feat, labels, 0.9, 0.2, 20.0)
    want = float(want)
    assert abs(loss - want) <= 1e-3 * abs(want), "train-step loss %g vs oracle %g" % (loss, want)
    assert loss2 == loss2 and loss2 != loss, "optimiser step had no effect"
    print("smoke: ResNet-18-OC train step B=%d loss %.6f (oracle %.6f), after one Adam step %.6f" % (B, loss, want, loss2))
