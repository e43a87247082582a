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


def test_train_and_score_from_a_flac_folder(tmp_path):
    """(f) row 1 end to end: ASVspoof-style FLAC folder -> native batch decoder -> prefetch thread + copy stream ->
    fused LFCC / ResNet step; ragged lengths (shorter and longer than feat_len frames) in one batch."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import flac_writer as fw
    rng = np.random.RandomState(0)
    wav = tmp_path / "flac"
    wav.mkdir()
    proto = []
    for i in range(12):
        n = 125000 if i == 5 else int(rng.randint(9000, 26000))          # one utterance longer than 750 frames: cropped
        x = np.round(np.cumsum(rng.randn(n)) * 30).astype(np.int64).clip(-30000, 30000)
        blocks = [4096] * (n // 4096) + ([n % 4096] if n % 4096 else [])
        fw.write_flac(str(wav / ("LA_T_%07d.flac" % i)), x, 16, 16000, [fw.FrameSpec(b, [fw.Sub("fixed", 1, rice=9)]) for b in blocks])
        proto.append("LA_0079 LA_T_%07d - %s %s" % (i, "-" if i % 2 == 0 else "A01", "bonafide" if i % 2 == 0 else "spoof"))
    (tmp_path / "proto.txt").write_text("\n".join(proto) + "\n")
    out = tmp_path / "models" / "lfcc_resnet_flac"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main_train.py"), "-o", str(out), "-m", "resnet", "--add_loss", "ang_iso",
                        "--gpu", "0", "--wave_dir", str(wav), "--protocol", str(tmp_path / "proto.txt"), "--dev_wave_dir", str(wav),
                        "--dev_protocol", str(tmp_path / "proto.txt"), "--batch_size", "4", "--num_epochs", "1", "--log_every", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = open(out / "train_loss.log").read().strip().splitlines()
    assert len(lines) == 1 + 3 and all(float(x.split("\t")[2]) == float(x.split("\t")[2]) for x in lines[1:])
    ep, dl, eer = open(out / "dev_loss.log").read().strip().splitlines()[1].split("\t")
    assert 0.0 <= float(eer) <= 1.0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "generate_score.py"), "--model_folder", str(tmp_path / "models"),
                        "-n", "lfcc_resnet_flac", "-s", str(tmp_path / "scores"), "-t", "19eval", "-l", "ocsoftmax", "--gpu", "0",
                        "--wave_dir", str(wav), "--protocol", str(tmp_path / "proto.txt"), "--batch_size", "5"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    # '19' tasks write under ./scores relative to the working directory (generate_score.py:29-30 of the reference)
    path = os.path.join(str(tmp_path), r.stdout.strip().splitlines()[-1])
    rows = [ln.split() for ln in open(path).read().strip().splitlines()]
    assert [x[0] for x in rows] == ["LA_T_%07d" % i for i in range(12)]
    assert [x[2] for x in rows] == ["bonafide" if i % 2 == 0 else "spoof" for i in range(12)]
    assert all(-1.0001 <= float(x[1]) <= 1.0001 for x in rows)


def test_adversarial_training_cli(tmp_path):
    """(f) row 4 end to end: `main_train.py --ADV_AUG --LA_aug` on a folder of originals + channel-augmented copies.
    Epoch 0: three-column log (the classifier trains on detached features, main_train.py:420-453); epoch 1: the
    gradient-reversed term joins and the log carries adv loss and both accuracies (main_train.py:377,471-477)."""
    import wave
    import numpy as np
    sys.path.insert(0, ROOT)
    from asvspoof2021_air_b200 import data
    ch, _ = data.channel_tables("LA")
    ori, aug = tmp_path / "ori", tmp_path / "aug"
    ori.mkdir(); aug.mkdir()
    rng = np.random.RandomState(1)
    proto = []
    for i in range(8):
        x = (rng.randn(20000 + 500 * i) * 3000).astype(np.int16)
        for folder, name in [(ori, "LA_T_%d" % i)] + [(aug, "LA_T_%d_%s" % (i, ch[1 + (2 * i + k) % (len(ch) - 1)])) for k in range(2)]:
            with wave.open(str(folder / (name + ".wav")), "wb") as f:
                f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(x.tobytes())
        proto.append("LA_0001 LA_T_%d - - %s" % (i, "bonafide" if i % 2 == 0 else "spoof"))
    (tmp_path / "p.txt").write_text("\n".join(proto) + "\n")
    out = tmp_path / "m"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main_train.py"), "-o", str(out), "-m", "resnet", "--add_loss", "ang_iso",
                        "--gpu", "0", "--ADV_AUG", "--LA_aug", "--wave_dir", str(ori), "--aug_wave_dir", str(aug), "--protocol",
                        str(tmp_path / "p.txt"), "--batch_size", "4", "--num_epochs", "2", "--log_every", "1", "--lr_d", "0.001"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = [ln.split("\t") for ln in open(out / "train_loss.log").read().strip().splitlines()[1:]]
    assert [len(x) for x in lines] == [3] * 4 + [6] * 4, lines
    for e, s, adv_loss, acc_m, acc_c, loss in lines[4:]:
        assert e == "1" and 0.0 < float(adv_loss) < 20.0 and 0.0 <= float(acc_m) <= 100.0 and 0.0 <= float(acc_c) <= 100.0
        assert float(loss) == float(loss)
    assert os.path.exists(out / "checkpoint" / "anti-spoofing_feat_model_2.pt")
