"""CPU tests of the host-side logic: drop-in module surfaces (constructor signatures, state_dict keys,
whole-module pickles), the CLI argparse surfaces against the reference's, data sources, score-file
formatting, LR schedule, and the data-parallel helpers under gloo with world_size 2."""
import io
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from oracle import ref_shim, state_spec as ss

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_resnet_module_surface_and_pickle():
    from asvspoof2021_air_b200.resnet import ResNet
    m = ResNet(3, 256, resnet_type='18', nclasses=2, device="cpu")
    spec = ss.resnet_spec()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s, _ in spec]
    sd = ss.seeded_state(spec, 3)
    m.load_state_dict(sd)
    buf = io.BytesIO()
    torch.save(m, buf)                                        # main_train.py:674 saves whole modules
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert type(m2).__name__ == "ResNet"
    for k, v in m2.state_dict().items():
        assert torch.equal(v.cpu().float(), sd[k].float()), k
    with pytest.raises(Exception):
        m(torch.zeros(1, 1, 60, 750))                         # no CPU path: fails loudly


def test_ecapa_module_surface_and_pickle():
    from asvspoof2021_air_b200.ecapa_tdnn import Res2Net2, Bottle2neck
    m = Res2Net2(Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60, device="cpu")
    spec = ss.ecapa_spec()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s, _ in spec]
    assert sum(p.numel() for p in m.parameters()) == 6337734
    sd = ss.seeded_state(spec, 4)
    m.load_state_dict(sd)
    buf = io.BytesIO()
    torch.save(m, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    for k, v in m2.state_dict().items():
        assert torch.equal(v.cpu().float(), sd[k].float()), k
    with pytest.raises(Exception):
        m(torch.zeros(1, 60, 750))


def test_loss_and_lfcc_module_surfaces():
    from asvspoof2021_air_b200.loss import OCSoftmax, AngularIsoLoss
    from asvspoof2021_air_b200.feature_extraction import LFCC
    l = AngularIsoLoss(256, r_real=0.9, r_fake=0.2, alpha=20.0)
    assert list(l.state_dict().keys()) == ["center"] and tuple(l.center.shape) == (1, 256)
    assert OCSoftmax().r_fake == 0.5                               # class default, loss.py:177
    f = LFCC(320, 160, 512, 16000, 20)
    assert sorted(f.state_dict().keys()) == ["l_dct.weight", "lfcc_fb"]
    assert tuple(f.lfcc_fb.shape) == (257, 20) and tuple(f.l_dct.weight.shape) == (20, 20)
    with pytest.raises(Exception):
        f(torch.zeros(1, 1000))


def _flag_table(parser):
    out = {}
    for a in parser._actions:
        for s in a.option_strings:
            if s not in ("-h", "--help"):
                out[s] = (a.default, tuple(a.choices) if a.choices else None)
    return out


def test_cli_surfaces_cover_the_reference_flags():
    sys.path.insert(0, ROOT)
    import main_train
    import generate_score
    ours = _flag_table(main_train.build_parser())
    want = {"--seed": 688, "--feat_len": 750, "--padding": "repeat", "--enc_dim": 256, "--model": "lcnn", "--batch_size": 64,
            "--lr": 0.0005, "--lr_decay": 0.5, "--interval": 30, "--beta_1": 0.9, "--beta_2": 0.999, "--eps": 1e-8,
            "--add_loss": None, "--weight_loss": 1, "--r_real": 0.9, "--r_fake": 0.2, "--alpha": 20, "--num_epochs": 200,
            "--ratio": 0.5, "--gpu": "1", "--lambda_": 0.05, "--lr_d": 0.0001, "--ADV_AUG": False}
    for k, v in want.items():
        assert ours[k][0] == v, k
    gs = _flag_table(generate_score.build_parser())
    assert gs["--task"][1] == tuple(generate_score.TASKS) and gs["--loss"][1] == (None, "ocsoftmax", "amsoftmax", "p2sgrad")
    if ref_shim.reference_available():
        for fname, table in (("main_train.py", ours), ("generate_score.py", gs)):
            src = open(os.path.join(ref_shim.REFERENCE_ROOT, fname)).read()
            flags = set(re.findall(r"add_argument\((?:['\"](-\w)['\"],\s*)?['\"](--\w+)['\"]", src))
            for short, long in flags:
                assert long in table, (fname, long)
                if short:
                    assert short in table, (fname, short)
    args = main_train.build_parser().parse_args(["-o", "/tmp/x", "--add_loss", "ang_iso", "-m", "resnet", "--pad_chop", "False"])
    assert args.pad_chop is False and args.model == "resnet"
    assert main_train.adjust_learning_rate(args, 5e-4, 61) == 5e-4 * 0.25          # main_train.py:144-147


def test_score_line_format_and_paths(tmp_path):
    sys.path.insert(0, ROOT)
    import generate_score as gs
    assert gs.format_line("LA", "LA_E_1", 0.5, 0) == "LA_E_1 0.5\n"
    assert gs.format_line("19eval", "LA_E_1", -0.25, 1) == "LA_E_1 -0.25 spoof\n"
    assert gs.score_file_path(str(tmp_path), "m", "19dev").endswith("m_19dev_score.txt")
    assert gs.score_file_path(str(tmp_path), "m", "DF").endswith(os.path.join("m_DF", "score.txt"))


def test_data_sources(tmp_path):
    from asvspoof2021_air_b200 import data
    s = data.SyntheticWaves(10, length=3200, seed=1)
    w, lens, lab, names, start = s.batch([0, 1, 5])
    w2 = s.batch([5])[0]
    assert w.shape == (3, 3200) and torch.equal(w[2], w2[0]) and lab.tolist() == [0, 1, 1] and start is None
    import wave
    a = (np.sin(np.arange(200000) * 0.01) * 1000).astype(np.int16)
    with wave.open(str(tmp_path / "u1.wav"), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(a.tobytes())
    np.save(tmp_path / "u2.npy", np.zeros(1000, dtype=np.float32))
    (tmp_path / "proto.txt").write_text("LA_0001 u1 - A07 spoof\nu2 bonafide\n")
    d = data.WaveFolder(str(tmp_path), str(tmp_path / "proto.txt"), feat_len=750, seed=0)
    w, lens, lab, names, start = d.batch([0, 1])
    assert w.shape == (2, 200000) and lens.tolist() == [200000, 1000] and lab.tolist() == [1, 0] and names == ["u1", "u2"]
    assert abs(float(w[0, 157]) - a[157] / 32768.0) < 1e-7 and (w[1] == 0).all()
    assert 0 <= int(start[0]) < (1 + 200000 // 160) - 750 and int(start[1]) == 0      # dataset.py:66-69


def test_flac_folder_and_prefetcher_on_cpu(tmp_path):
    """ASVspoof-style folder of FLAC files through the native batch decoder, iterated by the prefetch thread."""
    import flac_writer as fw
    from asvspoof2021_air_b200 import data
    rng = np.random.RandomState(4)
    proto, waves = [], {}
    for i in range(9):
        n = int(rng.randint(1000, 5000))
        x = np.round(np.cumsum(rng.randn(n)) * 20).astype(np.int64).clip(-30000, 30000)
        blocks = [1152] * (n // 1152) + ([n % 1152] if n % 1152 else [])
        fw.write_flac(str(tmp_path / ("LA_T_%07d.flac" % i)), x, 16, 16000,
                      [fw.FrameSpec(b, [fw.Sub("fixed", 1, porder=0)]) for b in blocks])
        waves["LA_T_%07d" % i] = (x / 32768.0).astype(np.float32)
        proto.append("LA_0079 LA_T_%07d - %s %s" % (i, "-" if i % 2 == 0 else "A01", "bonafide" if i % 2 == 0 else "spoof"))
    (tmp_path / "proto.txt").write_text("\n".join(proto) + "\n")
    src = data.WaveFolder(str(tmp_path), str(tmp_path / "proto.txt"), feat_len=750, seed=0, threads=3, verify=True)
    order = [[0, 1, 2, 3], [4, 5, 6, 7], [8]]
    seen = []
    for w, lens, lab, names, start in data.Prefetcher(src, order, depth=2, device=None):
        lens = lens if lens is not None else torch.full((len(names),), w.shape[1], dtype=torch.int32)
        assert w.shape == (len(names), int(lens.max())) and lab.tolist() == [int(n[-1]) % 2 for n in names]
        for j, n in enumerate(names):
            assert int(lens[j]) == len(waves[n]) and np.array_equal(w[j, :lens[j]].numpy(), waves[n]) and not w[j, lens[j]:].any()
        seen += names
    assert seen == ["LA_T_%07d" % i for i in range(9)]
    # a missing file is an error at construction, a damaged one surfaces in the consumer
    (tmp_path / "p2.txt").write_text("x LA_T_9999999 - - bonafide\n")
    with pytest.raises(FileNotFoundError):
        data.WaveFolder(str(tmp_path), str(tmp_path / "p2.txt"))
    raw = bytearray((tmp_path / "LA_T_0000003.flac").read_bytes())
    raw[-9] ^= 4
    (tmp_path / "LA_T_0000003.flac").write_bytes(bytes(raw))
    from asvspoof2021_air_b200.audio_io import AudioError
    with pytest.raises(AudioError, match="LA_T_0000003"):
        list(data.Prefetcher(src, order, device=None))


def test_parallel_helpers_single_process():
    from asvspoof2021_air_b200 import parallel
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    # cuts: the offsets at which the backward pass reports progress are bucket boundaries (no bucket straddles one)
    bc = parallel.bucket_bounds(20, 8, cuts=[3, 15])
    assert bc == [(15, 20), (7, 15), (3, 7), (0, 3)] and all(hi - lo <= 8 for lo, hi in bc)
    b = parallel.bucket_bounds(10, 4)
    assert b == [(6, 10), (2, 6), (0, 2)]
    r = parallel.GradReducer(torch.zeros(10), 10, bucket_elems=4)
    r.begin(); r.ready(3)
    assert r.finish() == 1.0


_WORKER = textwrap.dedent('''
    import os, sys, torch
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from asvspoof2021_air_b200 import parallel
    rank, world, _ = parallel.init_from_env("gloo")
    assert world == 2 and dist.get_backend() == "gloo"
    n = 1000
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)
    extra = torch.full((4,), float(rank + 1))
    red = parallel.GradReducer(flat, n - 100, bucket_elems=256, overlap=False)     # the tail is not reduced (frozen params)
    red.begin()
    launched = []
    for off in (900, 700, 300):          # backward order: the layers owning the highest offsets finish first
        red.ready(off)
        launched.append(red.next)
    scale = red.finish(extra)
    assert launched == [0, 0, 2] and red.next == 4, (launched, red.next)
    want = torch.arange(n, dtype=torch.float32) * 3
    assert torch.equal(flat[:900], want[:900]) and torch.equal(flat[900:], torch.arange(900, n, dtype=torch.float32) * (rank + 1))
    assert torch.equal(extra, torch.full((4,), 3.0)) and scale == 0.5
    p = torch.full((5,), float(rank))
    parallel.broadcast_state([p], 0)
    assert torch.equal(p, torch.zeros(5))
    lo, hi = parallel.shard_range(7, rank, world)
    t = torch.tensor([hi - lo], dtype=torch.float32); dist.all_reduce(t); assert int(t) == 7
    # validation scores / labels of unequal shards (main_train.dev_eer): rank 0 holds 3 trials, rank 1 holds 5
    k = 3 + 2 * rank
    sc = torch.arange(k, dtype=torch.float32) + 10 * rank
    lab = (torch.arange(k) %% 2).long()
    g_sc, g_lab = parallel.gather_ragged([sc, lab])
    assert g_sc.tolist() == [0.0, 1.0, 2.0, 10.0, 11.0, 12.0, 13.0, 14.0] and g_sc.dtype == torch.float32
    assert g_lab.tolist() == [0, 1, 0, 0, 1, 0, 1, 0] and g_lab.dtype == torch.int64
    e_sc, e_lab = parallel.gather_ragged([sc[:0] if rank == 0 else sc, lab[:0] if rank == 0 else lab])
    assert e_sc.tolist() == [10.0, 11.0, 12.0, 13.0, 14.0] and e_lab.numel() == 5
    dist.destroy_process_group()
    sys.stdout.write("rank %%d ok\\n" %% rank); sys.stdout.flush()
''')


def test_gradient_reducer_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    port = 29500 + os.getpid() % 1000
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok") == 2 and "rank 0" in r.stdout and "rank 1" in r.stdout, r.stdout


# ---------------------------------------------------------------------------------------------
# checkpoint compatibility with the reference's whole-module pickles (compat.py)
# ---------------------------------------------------------------------------------------------
_STANDIN = '''
import sys, torch, torch.nn as nn
sys.path.insert(0, %(root)r)
class PreActBlock(nn.Module): pass
class SelfAttention(nn.Module): pass
class ResNet(nn.Module): pass
class SEModule(nn.Module): pass
class Bottle2neck(nn.Module): pass
class Res2Net2(nn.Module): pass
class AngularIsoLoss(nn.Module): pass
'''

_STANDIN_MAIN = '''
import sys, torch, torch.nn as nn
sys.path.insert(0, %(tmp)r)
import model, ECAPA_TDNN, loss
LEAF = {"resnet": {"attention": model.SelfAttention}, "ecapa": {"se": ECAPA_TDNN.SEModule}}
def tree(root, sd, kinds, block_cls, depth_cls):
    """Hang the tensors of a flat state_dict on a module tree whose classes carry the reference's names."""
    for key, t in sd.items():
        parts, node = key.split("."), root
        for i, p in enumerate(parts[:-1]):
            if p not in node._modules:
                if p in depth_cls: cls = depth_cls[p]
                elif p.isdigit() and i == 1 and parts[0].startswith("layer"): cls = block_cls
                elif parts[0].startswith("layer") and i == 0 and block_cls is ECAPA_TDNN.Bottle2neck: cls = block_cls
                else: cls = nn.Sequential if not (parts[i + 1:] and parts[i + 1] in ("weight", "bias", "running_mean")) else nn.Module
                node.add_module(p, cls())
            node = node._modules[p]
        if kinds[key] == "param": node.register_parameter(parts[-1], nn.Parameter(t.clone()))
        else: node.register_buffer(parts[-1], t.clone())
    return root
blob = torch.load(%(blob)r)
m = tree(model.ResNet(), blob["resnet"], blob["resnet_kinds"], model.PreActBlock, LEAF["resnet"]); m.eval()
torch.save(m, %(tmp)r + "/ref_resnet.pt")
e = tree(ECAPA_TDNN.Res2Net2(), blob["ecapa"], blob["ecapa_kinds"], ECAPA_TDNN.Bottle2neck, LEAF["ecapa"]); e.train()
torch.save(e, %(tmp)r + "/ref_ecapa.pt")
l = loss.AngularIsoLoss(); l.center = nn.Parameter(blob["center"]); l.r_real, l.r_fake, l.alpha, l.feat_dim = 0.9, 0.2, 20.0, 256
l.softplus = nn.Softplus()
torch.save(l, %(tmp)r + "/ref_loss.pt")
'''


def test_reference_style_pickles_load_without_the_reference(tmp_path):
    """Pickles whose classes live in modules named `model`, `ECAPA_TDNN`, `loss` (as the reference's do) are written
    by a subprocess that has such modules, then loaded here, where they do not exist."""
    import subprocess
    from asvspoof2021_air_b200 import compat
    from asvspoof2021_air_b200.ecapa_tdnn import Bottle2neck, Res2Net2
    from asvspoof2021_air_b200.resnet import ResNet
    torch.manual_seed(5)
    r = ResNet(3, 256, '18', nclasses=2, device="cpu")
    e = Res2Net2(Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60, device="cpu")
    for m in (r, e):
        for k, v in m.state_dict().items():
            if v.dtype.is_floating_point:
                v.copy_(torch.randn_like(v) * 0.1 + (1.0 if k.endswith("running_var") else 0.0))
    kinds = lambda m: {k: ("param" if k in dict(m.named_parameters()) else "buffer") for k in m.state_dict()}
    center = torch.randn(1, 256)
    torch.save({"resnet": r.state_dict(), "resnet_kinds": kinds(r), "ecapa": e.state_dict(), "ecapa_kinds": kinds(e),
                "center": center}, tmp_path / "blob.pt")
    for name in ("model", "ECAPA_TDNN", "loss"):
        (tmp_path / (name + ".py")).write_text(_STANDIN % {"root": ROOT})
    (tmp_path / "make.py").write_text(_STANDIN_MAIN % {"tmp": str(tmp_path), "blob": str(tmp_path / "blob.pt")})
    p = subprocess.run([sys.executable, str(tmp_path / "make.py")], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    for name in ("model", "ECAPA_TDNN", "loss"):
        assert name not in sys.modules
    with pytest.raises(Exception):
        torch.load(tmp_path / "ref_resnet.pt", weights_only=False)          # plain torch.load cannot resolve `model.ResNet`
    r2 = compat.load_module(str(tmp_path / "ref_resnet.pt"), device="cpu")
    e2 = compat.load_module(str(tmp_path / "ref_ecapa.pt"), device="cpu")
    l2 = compat.load_module(str(tmp_path / "ref_loss.pt"))
    assert type(r2) is ResNet and not r2.training and type(e2) is Res2Net2 and e2.training
    for a, b in ((r, r2), (e, e2)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    assert type(l2).__name__ == "AngularIsoLoss" and torch.equal(l2.center.detach(), center)
    assert (l2.r_real, l2.r_fake, l2.alpha) == (0.9, 0.2, 20.0)
    for name in ("model", "ECAPA_TDNN", "loss"):
        assert name not in sys.modules                                       # the shims do not outlive the load
    # this package's own pickles pass straight through
    buf = tmp_path / "own.pt"
    torch.save(r, buf)
    r3 = compat.load_module(str(buf), device="cpu")
    assert type(r3) is ResNet and all(torch.equal(v, r3.state_dict()[k]) for k, v in r.state_dict().items())
    with pytest.raises(NotImplementedError):
        compat.adopt(torch.nn.Linear(2, 2))


def test_real_reference_pickles_load_through_compat(tmp_path):
    """The same with the reference's own classes (authoring container only: needs /root/reference)."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not mounted")
    import subprocess
    code = '''
import sys, torch
sys.path.insert(0, %r)
from oracle import ref_shim
rn, ec, ls = ref_shim.load("resnet"), ref_shim.load("ecapa_tdnn"), ref_shim.load("loss")
torch.manual_seed(688)
torch.save(rn.ResNet(3, 256, "18", nclasses=2), %r)
torch.save(ec.Res2Net2(ec.Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60), %r)
torch.save(ls.AngularIsoLoss(256, r_real=0.9, r_fake=0.2, alpha=20.0), %r)
torch.save({"resnet": rn.ResNet(3, 256, "18", nclasses=2).state_dict()}, %r)
''' % (ROOT, str(tmp_path / "r.pt"), str(tmp_path / "e.pt"), str(tmp_path / "l.pt"), str(tmp_path / "sd.pt"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    from asvspoof2021_air_b200 import compat
    with compat.reference_modules():
        raw = torch.load(tmp_path / "r.pt", map_location="cpu", weights_only=False)
    want = compat.named_state(raw)
    m = compat.load_module(str(tmp_path / "r.pt"), device="cpu")
    sd = m.state_dict()
    assert len(sd) == 117 and set(sd) == set(want) and all(torch.equal(sd[k], want[k]) for k in sd)
    e = compat.load_module(str(tmp_path / "e.pt"), device="cpu")
    assert len(e.state_dict()) == 248 and e.C == 512 and e.scale == 8
    l = compat.load_module(str(tmp_path / "l.pt"))
    assert (l.r_real, l.r_fake, l.alpha) == (0.9, 0.2, 20.0) and l.center.shape == (1, 256)


def test_adv_head_orchestration_with_emulated_kernels(monkeypatch, golden_dir):
    """The call sequence of adv.ChannelClassifier (argument order, shapes, accumulate semantics of the C entry points it
    strings together) checked on the CPU: every `ops.*` it uses is replaced by a torch emulation of that entry point's
    documented contract (include/air_b200.h), the result compared with the reference golden.  The CUDA kernels
    themselves are covered by tests/test_adv_gpu.py."""
    from asvspoof2021_air_b200 import adv

    class FakeOps:
        @staticmethod
        def linear_fwd(x, W, bias, y, M, N, K):
            y.copy_(x[:M, :K] @ W.t() + bias)

        @staticmethod
        def linear_bwd(x, W, dy, dx, dW, db, M, N, K):
            if dx is not None:
                dx.copy_(dy @ W)
            if dW is not None:
                dW.add_(dy.t() @ x)
                if db is not None:
                    db.add_(dy.sum(0))

        @staticmethod
        def dropout_relu_fwd(x, keep, y, p, generate, seed=0):
            assert not generate
            y.copy_(torch.relu(x * keep.float() / (1.0 - p)))

        @staticmethod
        def dropout_relu_bwd(dy, x, keep, dx, p):
            dx.copy_(dy * ((keep > 0) & (x > 0)).float() / (1.0 - p))

        @staticmethod
        def relu_ce_fwd_bwd(z, labels, B, C, grad_scale, loss_sum, correct, dz):
            logits = torch.relu(z.double())
            sm = torch.softmax(logits, 1)
            loss_sum.add_(-torch.log(sm[torch.arange(B), labels]).mean())
            correct.add_((logits.argmax(1) == labels).sum().int())
            if dz is not None:
                d = sm.clone()
                d[torch.arange(B), labels] -= 1
                dz.copy_((grad_scale * d / B * (z > 0)).float())

        @staticmethod
        def sgd_step(p, g, n, lr, grad_scale=1.0):
            p.sub_(lr * grad_scale * g)

        @staticmethod
        def adam_l2_step(p, g, m, v, n, lr, b1, b2, eps, wd, step, grad_scale=1.0):
            gg = g * grad_scale + wd * p
            m.mul_(b1).add_((1 - b1) * gg)
            v.mul_(b2).add_((1 - b2) * gg * gg)
            p.sub_(lr * (m / (1 - b1 ** step)) / ((v / (1 - b2 ** step)).sqrt() + eps))

    monkeypatch.setattr(adv, "ops", FakeOps)
    monkeypatch.setattr(adv, "_require_cuda", lambda t: None)
    g = np.load(os.path.join(golden_dir, "adv_golden.npz"))
    x = torch.from_numpy(g["x"])
    for C in (60, 13):
        p = "c%d_" % C
        clf = adv.ChannelClassifier(256, C, float(g["lambda"]), device="cpu")
        clf.load_state_dict({"classifier.0.weight": torch.from_numpy(g[p + "w1"]), "classifier.0.bias": torch.from_numpy(g[p + "b1"]),
                             "classifier.3.weight": torch.from_numpy(g[p + "w2"]), "classifier.3.bias": torch.from_numpy(g[p + "b2"])})
        labels, keep = torch.from_numpy(g[p + "labels"]), torch.from_numpy(g[p + "keep"])
        dfeat = torch.zeros(16, 256)
        loss, correct = clf.head_loss_and_feat_grad(x, labels, dfeat, keep_mask=keep)
        assert abs(float(loss) - float(g[p + "loss"])) < 1e-5 and int(correct) == int((g[p + "pred"] == g[p + "labels"]).sum())
        assert np.abs(dfeat.numpy() - g[p + "dfeat"]).max() <= 1e-7 + 1e-4 * np.abs(g[p + "dfeat"]).max()
        before = clf.flat.clone()
        clf.classifier_step(x, labels, lr=1e-4, keep_mask=keep)
        for k, name in (("classifier.0.weight", "dw1"), ("classifier.0.bias", "db1"), ("classifier.3.weight", "dw2"), ("classifier.3.bias", "db2")):
            ref = g[p + name]
            assert np.abs(clf._gviews[k].numpy() - ref).max() <= 1e-7 + 1e-4 * np.abs(ref).max(), k
        step = (clf.flat - before).abs()
        assert 0.5e-4 < float(step.max()) <= 1.0001e-4                  # first Adam step: |delta| ~ lr
        assert clf.step_count == 1 and clf(x).shape == (16, C) and float(clf(x).min()) >= 0.0


def test_augmented_folder_channel_labels_and_half_batches(tmp_path):
    """--ADV_AUG data side: originals + `<utt>_<channel>[_<device>]` copies, class indices in the reference's order."""
    import wave
    from asvspoof2021_air_b200 import data
    ch, dev = data.channel_tables("LAPA")
    assert len(ch) == 60 and ch[0] == "no_channel" and len(dev) == 13 and dev[-1] == "" and data.channel_tables("DF")[1] is None
    assert data.channel_tables("DF")[0] == ['no_channel', 'aac[16k]', 'aac[32k]', 'aac[8k]', 'mp3[16k]', 'mp3[32k]', 'mp3[8k]']
    ori, aug = tmp_path / "ori", tmp_path / "aug"
    ori.mkdir(); aug.mkdir()

    def wav(path, n, v):
        with wave.open(str(path), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
            f.writeframes((np.full(n, v, np.int16)).tobytes())
    proto = []
    for i in range(4):
        wav(ori / ("LA_T_%d.wav" % i), 1000 + i, i)
        proto.append("LA_0001 LA_T_%d - %s %s" % (i, "-" if i % 2 == 0 else "A03", "bonafide" if i % 2 == 0 else "spoof"))
    (tmp_path / "p.txt").write_text("\n".join(proto) + "\n")
    wav(aug / ("LA_T_1_%s_%s.wav" % (ch[7], dev[2])), 500, 100)
    wav(aug / ("LA_T_2_%s_%s.wav" % (ch[59], dev[0])), 700, 101)
    wav(aug / ("LA_T_0_%s_%s.wav" % (ch[1], dev[11])), 900, 102)
    src = data.AugWaveFolder(str(ori), str(aug), str(tmp_path / "p.txt"), "LAPA")
    assert len(src) == 7 and src.n_ori == 4
    w, lens, lab, names, start, chan = src.batch([0, 3, 4, 5, 6])
    assert lens.tolist() == [1000, 1003, 900, 500, 700] and lab.tolist() == [0, 1, 0, 1, 0]
    assert chan.tolist() == [[0, 12], [0, 12], [1, 11], [7, 2], [59, 0]]
    assert names[2].startswith("LA_T_0_") and abs(float(w[2, 0]) - 102 / 32768) < 1e-7
    b = next(iter(data.Prefetcher(src, [[4, 0]], device=None)))
    assert b.channels.tolist() == [[1, 11], [0, 12]] and b.labels_host.tolist() == [0, 0]
    with pytest.raises(KeyError):
        wav(aug / "LA_T_3_notacodec_x.wav", 10, 0)
        data.AugWaveFolder(str(ori), str(aug), str(tmp_path / "p.txt"), "LAPA")
    # single-label kinds
    aug2 = tmp_path / "aug2"
    aug2.mkdir()
    wav(aug2 / "LA_T_3_mp3[8k].wav", 300, 5)
    s2 = data.AugWaveFolder(str(ori), str(aug2), str(tmp_path / "p.txt"), "DF")
    assert s2.batch([4, 1])[5].tolist() == [6, 0]
    # main_train.py:226-233: int(B * ratio) originals + the rest augmented per step, pools reshuffled when exhausted
    steps = data.half_batches(4, 7, 4, 0.5, 5, np.random.RandomState(0))
    assert all(len(s) == 4 and all(i < 4 for i in s[:2]) and all(i >= 4 for i in s[2:]) for s in steps)
    assert sorted(i for s in steps[:2] for i in s[:2]) == [0, 1, 2, 3]          # the first pass over the originals is a permutation


class _FakeTrainer:
    """Stands in for trainer.Trainer in the CLI dry run: same call surface, no kernels."""
    instances = []

    def __init__(self, **kw):
        self.kw, self.calls, self.adv, self.lr_d, self.adv_stats, self.adv_stats_c = kw, [], [], None, [], []
        _FakeTrainer.instances.append(self)

    def attach_adversaries(self, class_counts, lambda_=0.05, lr_d=1e-4, seed=None):
        self.adv, self.lr_d, self.lambda_ = list(class_counts), lr_d, lambda_

    def train_step(self, waves, labels, lengths=None, start=None, lr=None, channels=None, step_seed=0, grl=True):
        self.calls.append(dict(B=waves.shape[0], lr=lr, lr_d=self.lr_d, channels=None if channels is None else channels.clone(),
                               ragged=lengths is not None, seed=step_seed, grl=grl))
        if channels is not None:
            n = len(self.adv)
            self.adv_stats = [(torch.tensor([0.5 + i], dtype=torch.float64), torch.tensor([1], dtype=torch.int32))
                              for i in range(n)] if grl else []
            self.adv_stats_c = [(torch.tensor([0.25]), torch.tensor([2], dtype=torch.int32)) for _ in range(n)]
        return torch.tensor([float(len(self.calls))])

    def eval_loss(self, waves, labels, lengths=None, start=None):
        return torch.tensor([2.0]), torch.arange(waves.shape[0], dtype=torch.float32)

    def modules(self):
        return torch.nn.Linear(1, 1), torch.nn.Linear(1, 1)


def _cli_dry_run(monkeypatch, argv):
    sys.path.insert(0, ROOT)
    import main_train
    from asvspoof2021_air_b200 import trainer
    _FakeTrainer.instances = []
    from asvspoof2021_air_b200 import parallel
    monkeypatch.setattr(trainer, "Trainer", _FakeTrainer)
    monkeypatch.setattr(parallel, "init_from_env", lambda backend=None: (0, 1, 0))
    from asvspoof2021_air_b200 import data
    monkeypatch.setattr(data.WaveFolder, "PIN", False)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(main_train, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(main_train, "dev_eer", lambda s, l, w: 0.125 if torch.cat(s).numel() == torch.cat(l).numel() else -1)
    args = main_train.init_params(argv)
    args.cuda = False
    main_train.train(args)
    return _FakeTrainer.instances[-1]


def test_training_cli_loop_dry_run_default_path(monkeypatch, tmp_path):
    """main_train.train() with the Trainer replaced by a recorder: batch order, LR schedule, log formats, checkpoints."""
    out = tmp_path / "m"
    tr = _cli_dry_run(monkeypatch, ["-o", str(out), "-m", "resnet", "--add_loss", "ang_iso", "--synthetic", "24", "--dev_synthetic", "10",
                                    "--batch_size", "8", "--num_epochs", "3", "--interval", "2", "--log_every", "2", "--lr", "0.001"])
    assert [c["B"] for c in tr.calls] == [8] * 9 and all(c["channels"] is None and not c["ragged"] for c in tr.calls)
    assert [c["lr"] for c in tr.calls] == [0.001] * 6 + [0.0005] * 3                       # lr * 0.5 ** (epoch // 2)
    lines = open(out / "train_loss.log").read().strip().splitlines()
    assert lines[1:] == ["%d\t%d\t%s" % (e, s, float(3 * e + s + 1)) for e in range(3) for s in range(3)]
    dev = open(out / "dev_loss.log").read().strip().splitlines()
    assert dev[1:] == ["%d\t2.0\t0.125" % e for e in range(3)]
    assert os.path.exists(out / "checkpoint" / "anti-spoofing_feat_model_3.pt") and os.path.exists(out / "anti-spoofing_loss_model.pt")


def test_training_cli_loop_dry_run_adversarial_path(monkeypatch, tmp_path):
    """--ADV_AUG bookkeeping (main_train.py:211-233,300-325,377,471-477): gate, heads, half-batches, adversaries from
    epoch 1, lr_d schedule, six-column log lines."""
    import wave
    from asvspoof2021_air_b200 import data
    ori, aug = tmp_path / "ori", tmp_path / "aug"
    ori.mkdir(); aug.mkdir()
    ch, dev = data.channel_tables("LAPA")

    def wav(path):
        with wave.open(str(path), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(np.zeros(400, np.int16).tobytes())
    proto = []
    for i in range(8):
        wav(ori / ("LA_T_%d.wav" % i))
        proto.append("LA_0001 LA_T_%d - - %s" % (i, "bonafide" if i % 2 == 0 else "spoof"))
        for k in range(2):
            wav(aug / ("LA_T_%d_%s_%s.wav" % (i, ch[1 + (3 * i + k) % 59], dev[(i + k) % 12])))
    (tmp_path / "p.txt").write_text("\n".join(proto) + "\n")
    argv = ["-o", str(tmp_path / "m"), "-m", "ecapa", "--add_loss", "ang_iso", "--ADV_AUG", "--LAPA_aug", "--wave_dir", str(ori),
            "--aug_wave_dir", str(aug), "--protocol", str(tmp_path / "p.txt"), "--batch_size", "4", "--num_epochs", "2",
            "--interval", "1", "--lr_d", "0.01", "--lambda_", "0.3", "--log_every", "1"]
    with pytest.raises(SystemExit, match="exactly one of"):
        _cli_dry_run(monkeypatch, argv + ["--DF_aug"])
    tr = _cli_dry_run(monkeypatch, argv)
    assert tr.adv == [60, 13] and tr.lambda_ == 0.3
    assert len(tr.calls) == 8 and all(c["B"] == 4 for c in tr.calls)                       # 8 originals / int(4 * 0.5) per epoch
    # epoch 0: the classifiers already train on the detached features (main_train.py:420-453), only the gradient-reversed
    # term waits for epoch 1 (main_train.py:377)
    assert [c["grl"] for c in tr.calls] == [False] * 4 + [True] * 4
    for c in tr.calls:
        assert c["channels"].shape == (4, 2) and c["channels"][:2].tolist() == [[0, 12], [0, 12]]      # originals first
        assert (c["channels"][2:, 0] > 0).all() and (c["channels"][2:, 1] < 12).all()
    assert [c["lr_d"] for c in tr.calls] == [0.01] * 4 + [0.005] * 4
    lines = open(tmp_path / "m" / "train_loss.log").read().strip().splitlines()
    assert [len(ln.split("\t")) for ln in lines[1:]] == [3] * 4 + [6] * 4
    e, s, adv_loss, acc_m, acc_c, loss = lines[5].split("\t")
    assert (e, s, float(adv_loss), float(acc_m), float(acc_c), float(loss)) == ("1", "0", 2.0, 25.0, 50.0, 5.0)


def test_adversarial_statistics_do_not_alias_the_work_buffers():
    """head_loss_and_feat_grad / classifier_step hand out copies: the classifier pass of the same train step re-uses the
    per-batch work buffers, so an alias would make the logged adversarial loss the classifier's."""
    import inspect
    from asvspoof2021_air_b200 import adv
    for fn in (adv.ChannelClassifier.head_loss_and_feat_grad, adv.ChannelClassifier.classifier_step):
        src = inspect.getsource(fn)
        assert 'return w["loss"].clone(), w["correct"].clone()' in src


def test_scoring_cli_loop_dry_run(monkeypatch, tmp_path):
    """generate_score.test_on_ASVspoof2021 with the Trainer replaced by a recorder: checkpoint loading through compat,
    architecture detection, batch order, `utt score [label]` lines for a '19' task and a 2021 task."""
    import types
    import wave
    sys.path.insert(0, ROOT)
    import generate_score as gs
    from asvspoof2021_air_b200 import compat, data, trainer

    class FakeScorer(_FakeTrainer):
        def load_modules(self, model, loss_model=None):
            self.loaded = (type(model).__name__, None if loss_model is None else tuple(loss_model.center.shape))

        def score_step(self, waves, lengths=None, start=None):
            self.calls.append(dict(B=waves.shape[0], ragged=lengths is not None))
            return torch.arange(waves.shape[0], dtype=torch.float32) * 0.25 - 0.5

    Res2Net2 = type("Res2Net2", (), {})
    loss_model = types.SimpleNamespace(center=torch.zeros(1, 256), r_real=0.9, r_fake=0.2, alpha=20.0)
    monkeypatch.setattr(trainer, "Trainer", FakeScorer)
    monkeypatch.setattr(compat, "load_module", lambda path, device=None: Res2Net2() if "feat" in path else loss_model)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(gs, "_device", lambda: torch.device("cpu"))
    monkeypatch.setattr(data.WaveFolder, "PIN", False)
    proto = []
    for i in range(5):
        with wave.open(str(tmp_path / ("LA_E_%d.wav" % i)), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(np.zeros(300 + 10 * i, np.int16).tobytes())
        proto.append("LA_0001 LA_E_%d - %s %s" % (i, "-" if i < 2 else "A09", "bonafide" if i < 2 else "spoof"))
    (tmp_path / "p.txt").write_text("\n".join(proto) + "\n")
    for task, want_cols in (("19eval", 3), ("LA", 2)):
        args = gs.build_parser().parse_args(["--model_folder", str(tmp_path), "-n", "m", "-s", str(tmp_path / "scores"), "-t", task,
                                             "-l", "ocsoftmax", "--wave_dir", str(tmp_path), "--protocol", str(tmp_path / "p.txt"),
                                             "--batch_size", "2"])
        _FakeTrainer.instances = []
        path = gs.test_on_ASVspoof2021(task, str(tmp_path / "feat.pt"), str(tmp_path / "loss.pt"), str(tmp_path / "scores"), "m",
                                       "ocsoftmax", args)
        tr = _FakeTrainer.instances[-1]
        assert tr.kw["arch"] == "ecapa" and tr.loaded == ("Res2Net2", (1, 256)) and [c["B"] for c in tr.calls] == [2, 2, 1]
        rows = [ln.split() for ln in open(path).read().strip().splitlines()]
        assert [r[0] for r in rows] == ["LA_E_%d" % i for i in range(5)] and all(len(r) == want_cols for r in rows)
        assert [float(r[1]) for r in rows] == [-0.5, -0.25, -0.5, -0.25, -0.5]
        if want_cols == 3:
            assert [r[2] for r in rows] == ["bonafide"] * 2 + ["spoof"] * 3


def _conv_route(layer, H, W):
    """Which kernel family ConvLayer.fprop / dgrad / wgrad pick for an (H, W) input -- the branch order of engine.py."""
    f = "patch" if layer.use_patch(H, W) else "patch1x1" if layer.p1x1_ok else "patch1d" if layer.d1_ok else "gemm"
    if not layer.need_dgrad:
        d = "-"
    else:
        d = ("patch" if layer.use_patch(H, W) else "patch1x1" if layer.p1x1_ok else "patch1d" if layer.d1_ok
             else "patch_s2" if layer.s2_dgrad_ok else "gemm")
    w = "patch1d" if layer.d1_ok else "patch" if (layer.wpatch_ok and layer._patch_efficiency(H, W) >= 0.5) else "gemm"
    return "%s/%s/%s" % (f, d, w)


def test_kernel_dispatch_table_at_the_benchmark_geometry():
    """Performance guard that needs no GPU: which tcgen05 kernel serves fprop / dgrad / wgrad of every conv layer at the
    BASELINE geometry ((60, 750) features).  The table is the one the measured numbers in profiles/ were taken with; a
    change that silently drops a layer back to the generic gather kernels shows up here, not only in the next bench."""
    from asvspoof2021_air_b200.engine import ResNetEngine
    from asvspoof2021_air_b200.engine_ecapa import EcapaEngine
    eng = ResNetEngine(enc_dim=256, nclasses=2, device="cpu")
    H, W = 18, 750                                            # after the 9x3 stem (stride 3 x 1): resnet.py:131
    got = {}
    for blk in eng.blocks:
        Ho, Wo = blk.conv1.out_hw(H, W)
        got[blk.name] = (H, W, _conv_route(blk.conv1, H, W), _conv_route(blk.conv2, Ho, Wo),
                         _conv_route(blk.sc, H, W) if blk.sc is not None else None)
        H, W = Ho, Wo
    want = {
        "layer1.0": (18, 750, "patch/patch/patch", "patch/patch/patch", "patch1x1/patch1x1/patch"),
        "layer1.1": (18, 750, "patch/patch/patch", "patch/patch/patch", None),
        "layer2.0": (18, 750, "gemm/patch_s2/gemm", "patch/patch/patch", "gemm/patch_s2/gemm"),
        "layer2.1": (9, 375, "patch/patch/patch", "patch/patch/patch", None),
        "layer3.0": (9, 375, "gemm/patch_s2/gemm", "patch/patch/patch", "gemm/patch_s2/gemm"),
        "layer3.1": (5, 188, "patch/patch/patch", "patch/patch/patch", None),
        "layer4.0": (5, 188, "gemm/patch_s2/gemm", "gemm/gemm/patch", "gemm/patch_s2/gemm"),
        "layer4.1": (3, 94, "gemm/gemm/patch", "gemm/gemm/patch", None),
    }
    assert got == want, got
    # ECAPA-TDNN-512 at T = 750: 1x1 layers on the TMA GEMM path, the 21 dilated Res2 branches on the 1-D patch kernels
    ec = EcapaEngine(device="cpu")
    routes = {}
    for layer in ec.convs():
        routes[layer.name] = _conv_route(layer, 1, 750)
    branch = [v for k, v in routes.items() if ".convs." in k]
    assert len(branch) == 21 and set(branch) == {"patch1d/patch1d/patch1d"}, routes
    # the biased 1x1 layers go through the generic entry points, which take their own TMA path for 1x1 / stride 1 in C
    rest = {k: v for k, v in routes.items() if ".convs." not in k}
    assert rest.pop("conv1") == "gemm/-/gemm" and set(rest.values()) == {"gemm/gemm/gemm"}, routes
    assert sorted(rest) == ["attention.3", "layer1.conv1", "layer1.conv3", "layer2.conv1", "layer2.conv3", "layer3.conv1",
                            "layer3.conv3", "layer4"]


def test_algorithmic_flops_per_utterance_match_the_survey():
    """The roofline numerator: conv / linear MACs of one forward pass computed from the engines' own layer geometry
    equal the figures SURVEY.md section 8(d) probed by hooking the reference modules (15.709 and 7.698 GFLOP/utt)."""
    from asvspoof2021_air_b200.engine import ResNetEngine
    from asvspoof2021_air_b200.engine_ecapa import EcapaEngine
    eng = ResNetEngine(enc_dim=256, nclasses=2, device="cpu")
    H, W = 18, 750
    total = 2 * H * W * 16 * 27                                    # 9x3 stem, resnet.py:131
    for b in eng.blocks:
        Ho, Wo = b.conv1.out_hw(H, W)
        total += 2 * Ho * Wo * (b.conv1.cout * b.conv1.K + b.conv2.cout * b.conv2.K + (b.sc.cout * b.sc.K if b.sc is not None else 0))
        H, W = Ho, Wo
    Ho, Wo = eng.conv5.out_hw(H, W)
    total += 2 * Ho * Wo * eng.conv5.cout * eng.conv5.K + 2 * (512 * 256 + 256 * 2)
    assert (Ho, Wo) == (1, 94) and abs(total / 1e9 - 15.709) < 1e-3, total
    ec = EcapaEngine(device="cpu")
    T = 750
    t = sum(2 * T * l.cout * l.taps * l.cin for l in ec.convs())
    t += 2 * T * 128 * 4608                                        # attention.0 on [x | mean | std] (ecapa_tdnn.py:139-145,174)
    t += 3 * 2 * (512 * 128 + 128 * 512) + 2 * (3072 * 256 + 256 * 2)    # SE bottlenecks, fc6, fc7
    assert abs(t / 1e9 - 7.698) < 2e-3, t


def test_binding_smoke_every_product_path_on_cpu():
    """tests/binding_smoke.py: the real orchestration of train step / validation / scoring (both nets, ragged batches,
    the adversarial branch), fp32 + bf16 LFCC with every pad mode and the detection metrics, run on CPU tensors with
    kernel launches failing softly -- every ctypes call must be accepted by the header-derived argtypes and no
    Python-level error may occur anywhere on those paths."""
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU tests when a device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "binding_smoke.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0 and "binding smoke ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


def test_packed_corpus_round_trip_and_threaded_gather(tmp_path):
    """pack_folder() -> PackedWaves: the same batches as the FLAC / WAV folder it was made from, bit for bit."""
    import flac_writer as fw
    import wave
    from asvspoof2021_air_b200 import data
    rng = np.random.RandomState(8)
    proto = []
    for i in range(7):
        n = int(rng.randint(300, 4000))
        x = rng.randint(-32768, 32768, n).astype(np.int64)
        x[:3] = [-32768, 32767, 0]                                         # both ends of the 16-bit range survive
        if i % 2:
            with wave.open(str(tmp_path / ("u%d.wav" % i)), "wb") as f:
                f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(x.astype(np.int16).tobytes())
        else:
            fw.write_flac(str(tmp_path / ("u%d.flac" % i)), x, 16, 16000, [fw.FrameSpec(n, [fw.Sub("verbatim")])])
        proto.append("spk u%d - - %s" % (i, "spoof" if i % 3 == 0 else "bonafide"))
    (tmp_path / "p.txt").write_text("\n".join(proto) + "\n")
    src = data.WaveFolder(str(tmp_path), str(tmp_path / "p.txt"), seed=3)
    meta = data.pack_folder(src, str(tmp_path / "corpus"), batch=3)
    assert meta["samples"] == sum(meta["lengths"]) and os.path.getsize(tmp_path / "corpus.i16") == 2 * meta["samples"]
    for threads in (1, 3):
        pk = data.PackedWaves(str(tmp_path / "corpus"), seed=3, threads=threads)
        ref = data.WaveFolder(str(tmp_path), str(tmp_path / "p.txt"), seed=3)
        assert len(pk) == 7
        for idx in ([0, 1, 2, 3], [6, 2], [5]):
            a, b = pk.batch(idx), ref.batch(idx)
            assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and a[3] == b[3]
            assert torch.equal(a[4], b[4]) and len(a) == 5
    batches = list(data.Prefetcher(data.PackedWaves(str(tmp_path / "corpus")), [[0, 1], [2, 3, 4]], device=None))
    assert [b[0].shape[0] for b in batches] == [2, 3]
    # a float source that is not 16-bit PCM is refused rather than silently quantised
    np.save(tmp_path / "f.npy", (rng.randn(100) * 0.1).astype(np.float32))
    (tmp_path / "p2.txt").write_text("f bonafide\n")
    with pytest.raises(ValueError, match="not 16-bit"):
        data.pack_folder(data.WaveFolder(str(tmp_path), str(tmp_path / "p2.txt")), str(tmp_path / "c2"))
    # a stale index (an item reaching past the blob) is refused at load time and by the native gather itself
    import json, ctypes
    from asvspoof2021_air_b200 import _lib
    good = json.load(open(tmp_path / "corpus.json"))
    stale = dict(good, lengths=good["lengths"][:-1] + [good["lengths"][-1] + 5])
    json.dump(stale, open(tmp_path / "corpus.json", "w"))
    with pytest.raises(ValueError, match="lies outside"):
        data.PackedWaves(str(tmp_path / "corpus"))
    json.dump(good, open(tmp_path / "corpus.json", "w"))
    pk = data.PackedWaves(str(tmp_path / "corpus"))
    offs, lens = np.array([good["samples"] - 3], dtype=np.int64), np.array([4], dtype=np.int32)
    row = np.zeros(8, dtype=np.float32)
    assert _lib.lib().air_audio_gather_i16_f32(ctypes.c_void_p(pk.blob.ctypes.data), _lib.LL(pk.blob.shape[0]),
                                               offs.ctypes.data_as(ctypes.c_void_p), lens.ctypes.data_as(ctypes.c_void_p), 1,
                                               row.ctypes.data_as(ctypes.c_void_p), _lib.LL(8), 1) == -1
    with open(tmp_path / "corpus.i16", "ab") as f:
        f.write(b"\x00\x00")
    with pytest.raises(ValueError, match="announces"):
        data.PackedWaves(str(tmp_path / "corpus"))


def test_training_cli_dry_run_from_packed_corpora(monkeypatch, tmp_path):
    """--packed_waves / --dev_packed_waves, plain and with the adversarial branch (classes travel inside the corpus)."""
    import wave
    from asvspoof2021_air_b200 import data
    ori, aug = tmp_path / "ori", tmp_path / "aug"
    ori.mkdir(); aug.mkdir()
    ch, _ = data.channel_tables("DF")
    proto = []
    for i in range(6):
        for folder, name in ((ori, "u%d" % i), (aug, "u%d_%s" % (i, ch[1 + i % 6]))):
            with wave.open(str(folder / (name + ".wav")), "wb") as f:
                f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(np.full(200 + i, i, np.int16).tobytes())
        proto.append("s u%d - - %s" % (i, "bonafide" if i % 2 else "spoof"))
    (tmp_path / "p.txt").write_text("\n".join(proto) + "\n")
    data._main(["pack", "--wave_dir", str(ori), "--protocol", str(tmp_path / "p.txt"), "--out", str(tmp_path / "plain")])
    data._main(["pack", "--wave_dir", str(ori), "--protocol", str(tmp_path / "p.txt"), "--out", str(tmp_path / "adv"),
                "--aug_wave_dir", str(aug), "--kind", "DF"])
    monkeypatch.setattr(data.PackedWaves, "PIN", False)
    tr = _cli_dry_run(monkeypatch, ["-o", str(tmp_path / "m1"), "-m", "resnet", "--add_loss", "ang_iso", "--packed_waves",
                                    str(tmp_path / "plain"), "--dev_packed_waves", str(tmp_path / "plain"), "--batch_size", "3",
                                    "--num_epochs", "1"])
    assert [c["B"] for c in tr.calls] == [3, 3] and all(c["ragged"] for c in tr.calls)
    assert open(tmp_path / "m1" / "dev_loss.log").read().strip().splitlines()[1] == "0\t2.0\t0.125"
    tr = _cli_dry_run(monkeypatch, ["-o", str(tmp_path / "m2"), "-m", "resnet", "--add_loss", "ang_iso", "--ADV_AUG", "--DF_aug",
                                    "--packed_waves", str(tmp_path / "adv"), "--batch_size", "4", "--num_epochs", "2"])
    assert tr.adv == [7] and len(tr.calls) == 6
    assert [c["grl"] for c in tr.calls] == [False] * 3 + [True] * 3
    for c in tr.calls:
        assert c["channels"].shape == (4,) and c["channels"][:2].tolist() == [0, 0] and (c["channels"][2:] > 0).all()


def test_sm_reservation_for_the_gradient_exchange(monkeypatch):
    """ops.reserve_sms(k) takes k SMs out of every persistent grid launched afterwards (the NCCL kernels of the gradient
    exchange occupy them during the backward pass, parallel.collective_ctas() = the NCCL_MAX_CTAS cap); 0 restores them."""
    from asvspoof2021_air_b200 import ops, parallel
    monkeypatch.setitem(ops._SMS, 0, 148)
    try:
        assert ops.num_sms(0) == 148
        ops.reserve_sms(2)
        assert ops.num_sms(0) == 146
        ops.reserve_sms(1000)
        assert ops.num_sms(0) == 1                      # never an empty grid
    finally:
        ops.reserve_sms(0)
    assert ops.num_sms(0) == 148
    monkeypatch.setenv("NCCL_MAX_CTAS", "2")
    assert parallel.collective_ctas() == 2
    monkeypatch.setenv("NCCL_MAX_CTAS", "x")
    assert parallel.collective_ctas() == 0
    monkeypatch.delenv("NCCL_MAX_CTAS")
    assert parallel.collective_ctas() == 0
