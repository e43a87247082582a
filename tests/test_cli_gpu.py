"""GPU test of the two CLIs on synthetic waves: main_train.py writes the reference's artefacts (args.json,
train_loss.log `epoch\\tstep\\tloss`, dev_loss.log, whole-module checkpoints), generate_score.py reloads them and
writes `utt score[ label]` lines equal to scoring through the drop-in modules directly."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("arch", ["resnet", "ecapa"])
def test_train_then_score(tmp_path, arch):
    out = tmp_path / "models" / ("lfcc_%s_ocs" % arch)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main_train.py"), "-o", str(out), "-m", arch, "--add_loss", "ang_iso",
                        "--gpu", "0", "--synthetic", "32", "--dev_synthetic", "8", "--batch_size", "8", "--num_epochs", "2",
                        "--log_every", "2"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = open(out / "train_loss.log").read().strip().splitlines()
    assert lines[0].startswith("Start recording") and len(lines) == 1 + 2 * 4
    e, s, l = lines[1].split("\t")
    assert (e, s) == ("0", "0") and float(l) > 0
    dev = open(out / "dev_loss.log").read().strip().splitlines()
    assert len(dev) == 3
    for ln in dev[1:]:                                   # epoch \t dev loss \t EER (sorted / reduced on the GPU)
        ep, dl, eer = ln.split("\t")
        assert float(dl) == float(dl) and (0.0 <= float(eer) <= 1.0 or float(eer) != float(eer)), ln
    for name in ("args.json", "anti-spoofing_feat_model.pt", "anti-spoofing_loss_model.pt",
                 "checkpoint/anti-spoofing_feat_model_2.pt", "checkpoint/anti-spoofing_loss_model_2.pt"):
        assert os.path.exists(out / name), name
    losses = [float(x.split("\t")[2]) for x in lines[1:]]
    assert all(v == v for v in losses)
    # --- scoring
    r = subprocess.run([sys.executable, os.path.join(ROOT, "generate_score.py"), "--model_folder", str(tmp_path / "models"),
                        "-n", "lfcc_%s_ocs" % arch, "-s", str(tmp_path / "scores"), "-t", "LA", "-l", "ocsoftmax", "--gpu", "0",
                        "--synthetic", "12", "--batch_size", "8"], capture_output=True, text=True, timeout=600, env=env,
                       cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    path = tmp_path / "scores" / ("lfcc_%s_ocs_LA" % arch) / "score.txt"
    rows = [ln.split() for ln in open(path).read().strip().splitlines()]
    assert len(rows) == 12 and all(len(x) == 2 for x in rows) and rows[0][0] == "SYN_0000000"
    # the same scores through the unpickled drop-in modules (generate_score.py:84-119 semantics)
    sys.path.insert(0, ROOT)
    from asvspoof2021_air_b200 import data
    from asvspoof2021_air_b200.feature_extraction import LFCC
    model = torch.load(out / "anti-spoofing_feat_model.pt", weights_only=False).cuda().eval()
    loss_model = torch.load(out / "anti-spoofing_loss_model.pt", weights_only=False).cuda()
    waves = data.SyntheticWaves(12).batch(list(range(12)))[0].cuda()
    lf = LFCC(320, 160, 512, 16000, 20).cuda()
    feats = lf.extract(waves, feat_len=750, padding="repeat", layout="btd", dtype=torch.float32)       # (B,750,60)
    x = feats.unsqueeze(1).transpose(2, 3)                                                         # generate_score.py:91-94
    if arch == "ecapa":
        x = x.squeeze(1)
    with torch.no_grad():
        f, _ = model(x)
        _, score = loss_model(f, torch.zeros(12, device="cuda"))
    want = (-score).cpu()
    got = torch.tensor([float(x[1]) for x in rows])
    assert torch.allclose(got, want, atol=5e-3), (got - want).abs().max()
