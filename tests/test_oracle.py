"""CPU tests: the oracle restatements against the golden vectors generated from the reference,
and (when /root/reference is mounted) against the reference modules themselves."""
import os

import numpy as np
import pytest
import torch

from oracle import lfcc_oracle as lo
from oracle import nets_oracle as no
from oracle import ref_shim, state_spec as ss
from tolerances import lfcc_close, lfcc_worst


@pytest.fixture(scope="module")
def gl(golden_dir):
    return np.load(os.path.join(golden_dir, "lfcc_golden.npz"))


def test_constants_bit_identical(gl):
    assert np.array_equal(lo.linear_filterbank(), gl["lfcc_fb"])
    assert np.abs(lo.dct_matrix() - gl["dct_weight"]).max() < 1e-7
    # bins 0 and 256 carry no weight; every bin feeds at most two filters
    fb = gl["lfcc_fb"]
    assert fb[0].sum() == 0 and fb[256].sum() == 0
    assert ((fb > 0).sum(axis=1) <= 2).all()


def test_lfcc_oracle_vs_reference_golden(gl):
    w = ss.seeded_waves(32, 64000, seed=0, edge_rows=True).numpy()
    rows = gl["full_rows"]
    y = lo.lfcc(w[rows])
    assert y.shape == (len(rows), 401, 60)
    assert lfcc_close(y, gl["full"]).all(), lfcc_worst(y, gl["full"])
    yall = lo.lfcc(w)
    assert lfcc_close(yall[:, gl["sparse_frames"]], gl["sparse"]).all()
    # checksum over frames: 401 terms, each within the per-element tolerance
    assert (np.abs(yall.sum(axis=1) - gl["colsum"]) <= 1e-4 * (np.abs(gl["colsum"]) + 401.0)).all()


@pytest.mark.parametrize("L", [3200, 12345, 5920, 6079, 321, 800])
def test_lfcc_oracle_ragged_lengths(gl, L):
    w = ss.seeded_waves(2, L, seed=L).numpy()
    y = lo.lfcc(w)
    ref = gl["ragged_%d" % L]
    assert y.shape == ref.shape == (2, 1 + L // 160, 60)      # frame count is exact
    assert lfcc_close(y, ref).all(), lfcc_worst(y, ref)


def test_silence_vector(gl):
    s = lo.silence_vector()
    assert lfcc_close(s, gl["silence"][0]).all()
    assert abs(s[0] + 30.96) < 0.01          # c0 of log10(eps) frames


def test_pad_crop_index_maps_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "padcrop_golden.npz"))
    sil_c0 = float(g["silence_pad_value"][0, 0, 0])
    for T in (1, 21, 401, 749):
        assert np.array_equal(lo.frame_index_map(T, 750, "repeat"), g["repeat_%d" % T])
        assert np.array_equal(lo.frame_index_map(T, 750, "zero"), g["zero_%d" % T])
        m = lo.frame_index_map(T, 750, "silence")
        ref = g["silence_%d" % T]
        ours = np.where(m == lo.SRC_SILENCE, sil_c0, m + 1000.0)
        assert np.array_equal(ours.astype(np.float32), ref)
    m = lo.frame_index_map(1000, 750, "repeat", startp=17)
    assert m[0] == 17 and m[-1] == 17 + 749
    assert np.array_equal(lo.frame_index_map(750, 750, "zero"), np.arange(750))


def test_kernel_algorithm_emulation_matches_oracle():
    """The lane/register algorithm of csrc/lfcc.cu, emulated in float64, against the oracle."""
    from asvspoof2021_air_b200 import lfcc_tables as lt
    import lfcc_emulation as em
    tbl = lt.pack_table(lt.linear_filterbank(512, 16000, 20), lt.dct_ortho_matrix(20)).numpy()
    for L in (800, 2000):
        w = ss.seeded_waves(1, L, seed=L).numpy()
        assert np.abs(em.lfcc_emulated(w[0], tbl) - lo.lfcc(w)[0]).max() < 2e-6


def _nets_inputs(batch, seed):
    feats = lo.apply_frame_map(lo.lfcc(ss.seeded_waves(batch, 64000, seed=seed).numpy()),
                               lo.frame_index_map(401, 750, "repeat"))
    return torch.from_numpy(feats).float(), ss.seeded_labels(batch, seed)


@pytest.mark.parametrize("arch", ["resnet", "ecapa"])
def test_nets_oracle_vs_reference_golden(golden_dir, arch):
    g = np.load(os.path.join(golden_dir, "nets_golden.npz"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    feats, labels = _nets_inputs(int(g["batch"]), int(g["seed"]))
    assert np.array_equal(labels.numpy(), g["labels"])
    spec = ss.resnet_spec() if arch == "resnet" else ss.ecapa_spec()
    sd = ss.seeded_state(spec, 11)
    for k in ss.trainable_keys(spec):
        sd[k].requires_grad_(True)
    center = ss.seeded_center(256, 11).requires_grad_(True)
    if arch == "resnet":
        feat, logits = no.resnet_forward(sd, feats.unsqueeze(1).transpose(2, 3).contiguous(), True,
                                         update_running=True)
    else:
        feat, logits = no.ecapa_forward(sd, feats.transpose(1, 2).contiguous(), True, update_running=True)
    loss, score = no.ocsoftmax(center, feat, labels, 0.9, 0.2, 20.0)
    loss.backward()
    assert abs(float(loss) - float(g[arch + "_loss"])) <= 1e-3 * abs(float(g[arch + "_loss"]))
    assert abs(float(no.cross_entropy(logits, labels)) - float(g[arch + "_ce"])) <= 1e-3
    assert np.allclose(feat.detach().numpy(), g[arch + "_feat"], rtol=1e-3, atol=1e-3)
    assert np.allclose(logits.detach().numpy(), g[arch + "_logits"], rtol=1e-3, atol=1e-3)
    assert np.allclose(score.detach().numpy(), g[arch + "_score"], rtol=1e-3, atol=1e-4)
    assert np.allclose(center.grad.numpy(), g[arch + "_center_grad"], rtol=1e-3, atol=1e-5)
    gmax = float(g[arch + "_grad_norm"].max())
    for k, n, h in zip(g[arch + "_grad_keys"], g[arch + "_grad_norm"], g[arch + "_grad_head"]):
        gr = sd[str(k)].grad
        assert gr is not None, k
        # deep gradients are ill-conditioned (1e-6 input noise -> ~1% on single elements), so a
        # different host CPU's conv kernels may move them: norms to 2%, leading elements loosely
        assert abs(float(gr.double().norm()) - n) <= 2e-2 * n + 1e-5 * gmax, k
        hh = gr.reshape(-1)[:8].double().numpy()
        assert np.allclose(hh, h[:len(hh)], rtol=5e-2, atol=2e-2 * n + 1e-5 * gmax), k
    # parameters the reference leaves without gradient under --add_loss ang_iso
    nograd = [k for k in ss.trainable_keys(spec) if sd[k].grad is None]
    expect = {"fc_mu.weight", "fc_mu.bias"} if arch == "resnet" else {"fc7.weight", "fc7.bias", "bn7.weight", "bn7.bias"}
    assert set(nograd) == expect
    rs = {str(k): v for k, v in zip(g[arch + "_running_keys"], g[arch + "_running_sum"])}
    for k, v in rs.items():
        assert abs(float(sd[k].double().sum()) - v) <= 1e-3 * abs(v) + 1e-3, k
    # eval-mode (scoring) path with the updated running statistics
    with torch.no_grad():
        if arch == "resnet":
            fe, le = no.resnet_forward(sd, feats.unsqueeze(1).transpose(2, 3).contiguous(), False)
        else:
            fe, le = no.ecapa_forward(sd, feats.transpose(1, 2).contiguous(), False)
        _, sc = no.ocsoftmax(center, fe, torch.zeros(len(labels)), 0.9, 0.2, 20.0)
    assert np.allclose(fe.numpy(), g[arch + "_eval_feat"], rtol=1e-3, atol=1e-3)
    assert np.allclose((-sc).numpy(), g[arch + "_eval_score"], rtol=1e-3, atol=1e-4)


def test_optimizer_restatement_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = torch.randn(1000)
        ref.grad = g.clone()
        opt.step()
        no.adam_l2_step(p, g, m, v, step, 5e-4)
        assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-7)
    assert no.lr_at_epoch(5e-4, 61) == 5e-4 * 0.25


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not mounted")
def test_state_specs_match_reference_modules():
    rn, ec = ref_shim.load("resnet"), ref_shim.load("ecapa_tdnn")
    m = rn.ResNet(3, 256, "18", nclasses=2)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, s) for k, s, _ in ss.resnet_spec()]
    e = ec.Res2Net2(ec.Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60)
    assert [(k, tuple(v.shape)) for k, v in e.state_dict().items()] == [(k, s) for k, s, _ in ss.ecapa_spec()]
    assert len(ss.resnet_spec()) == 117 and len(ss.ecapa_spec()) == 248


def test_tensor_core_lfcc_algorithm_and_tables_on_cpu():
    """The folded real DFT with the 3-term bf16 split of csrc/lfcc_tc.cu, emulated from the PACKED device tables
    (pre-swizzled DFT chunks, per-bin filter pairs), meets the fp32 LFCC tolerance against the float64 oracle."""
    import lfcc_emulation as em
    from asvspoof2021_air_b200 import lfcc_tables as lt
    from tolerances import lfcc_close, lfcc_worst
    fb, dct = lt.linear_filterbank(512, 16000, 20), lt.dct_ortho_matrix(20)
    assert lt.tc_filter_structure_ok(fb)
    tbl, wmat = lt.pack_tc_table(fb, dct), lt.pack_tc_dft()
    w = ss.seeded_waves(2, 4000, seed=3, edge_rows=True).numpy()
    got = em.lfcc_tc_emulate(w, tbl.numpy(), wmat)
    want = lo.lfcc(w)[:, :, :20]
    assert lfcc_close(got, want).all(), lfcc_worst(got, want)
    assert lfcc_worst(got, want) < 2e-5


def test_tensor_core_lfcc_accuracy_envelope_and_kernel_choice_on_cpu():
    """Accuracy envelope of the tensor-core arithmetic (fp16 hi parts + bf16 residuals, emulated from the PACKED tables): on
    a frame whose weak bands lie ~60 dB under its strongest harmonics -- where a bf16 / bf16 split leaves the 1e-4 bar
    (1.0e-4) -- and on the same signal 60 dB quieter -- where an fp16 / fp16 split underflows (3e-4) -- it stays within a
    few 1e-5 of the float64 oracle, like the reference's own fp32 arithmetic.  Hence LFCC.impl_for() serves every output
    dtype from the tensor-core kernel."""
    import lfcc_emulation as em
    from asvspoof2021_air_b200 import lfcc_tables as lt
    from asvspoof2021_air_b200.feature_extraction import LFCC
    from oracle import lfcc_torch
    from tolerances import lfcc_worst
    n = np.arange(8000)
    rng = np.random.RandomState(0)
    speech = sum(0.3 / h * np.sin(2 * np.pi * 140 * h * n / 16000) for h in range(1, 9)) + 3e-4 * rng.randn(8000)
    fb, dct = lt.linear_filterbank(512, 16000, 20), lt.dct_ortho_matrix(20)
    tbl, wmat = lt.pack_tc_table(fb, dct).numpy(), lt.pack_tc_dft()
    for scale, bar in ((1.0, 3e-5), (1e-3, 1e-5)):
        w = (scale * speech)[None].astype(np.float32)
        want = lo.lfcc(w)[:, :, :20]
        tc = em.lfcc_tc_emulate(w, tbl, wmat)
        assert lfcc_worst(tc, want) < bar, (scale, lfcc_worst(tc, want))      # measured 1.7e-5 / 1.4e-6
        if scale == 1.0:
            fp32 = lfcc_torch.TorchLFCC()(torch.from_numpy(w)).numpy()[:, :, :20]
            assert lfcc_worst(fp32, want) < 3e-5                               # measured 8.6e-6
    m = LFCC(320, 160, 512, 16000, 20)
    if m.impl == "auto":
        assert m.impl_for(torch.float32) == "tc" and m.impl_for(torch.bfloat16) == "tc"
    m.impl = "fft"
    assert m.impl_for(torch.bfloat16) == "fft" and m.impl_for(torch.float32) == "fft"


# ---------------------------------------------------------------------------------------------
# detection metrics (eval_metrics.py): the numpy restatement against the reference's own outputs
# ---------------------------------------------------------------------------------------------
def test_det_oracle_bit_identical_to_reference_golden(golden_dir):
    import det_cases as dc
    from oracle import metrics_oracle as mo
    g = dc.load(golden_dir)
    names = dc.case_names(g)
    assert len(names) >= 12
    for name in names:
        tar, non = g[name + "__target"], g[name + "__nontarget"]
        for tag, neg in (("", False), ("_neg", True)):
            frr, far, thr = mo.det_curve(tar, non, negate=neg)
            e, eth, idx = mo.eer(tar, non, negate=neg)
            want = g[name + tag + "__eer"]
            assert dc.same(e, want[0]) and idx == int(want[2]), (name, tag)
            if idx > 0 or tar.dtype == np.float64:       # numpy 2 subtracts the 0.001 of point 0 in the input dtype
                assert dc.same(eth, want[1]), (name, tag)
            assert dc.same([frr.sum(), far.sum()], g[name + tag + "__sums"]), (name, tag)
            if name + tag + "__frr" in g.files:
                assert dc.same(frr, g[name + tag + "__frr"]) and dc.same(far, g[name + tag + "__far"]), (name, tag)
                assert dc.same(thr[1:], g[name + tag + "__thr"]), (name, tag)
        if name + "__tdcf" in g.files:
            c1, c2 = mo.tdcf_weights(*dc.DET_ASV, dc.DET_COST)
            curve, cthr = mo.tdcf_curve(tar, non, c1, c2)
            i = int(np.argmin(curve))
            want = g[name + "__tdcf"]
            assert dc.same(curve[i], want[0]) and i == int(want[2]) and dc.same(curve.sum(), want[3]), name
            if name + "__tdcf_curve" in g.files:
                assert dc.same(curve, g[name + "__tdcf_curve"]), name
    e, thr, _ = mo.eer(g["asv__tar"], g["asv__non"])
    rates = mo.asv_error_rates(g["asv__tar"], g["asv__non"], g["asv__spoof"], thr)
    assert dc.same([e, thr] + list(rates), g["asv__expected"])


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")
def test_det_oracle_vs_reference_on_fresh_random_scores():
    import sys
    import det_cases as dc
    from oracle import metrics_oracle as mo
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    import eval_metrics as rem
    rng = np.random.RandomState(99)
    for trial in range(20):
        nt, nn = int(rng.randint(1, 400)), int(rng.randint(1, 400))
        q = [None, 1, 2][trial % 3]
        tar, non = rng.randn(nt) + 0.5, rng.randn(nn) - 0.5
        if q:
            tar, non = np.round(tar, q), np.round(non, q)
        frr, far, thr = rem.compute_det_curve(tar, non)
        a, b, c = mo.det_curve(tar, non)
        assert dc.same(a, frr) and dc.same(b, far) and dc.same(c, thr)
        assert dc.same(mo.eer(tar, non)[:2], rem.compute_eer(tar, non))


def test_adv_classifier_oracle_vs_reference_golden(golden_dir):
    """ChannelClassifier + gradient reversal + CE (model.py:976-1023, main_train.py:377-403) for a fixed dropout mask."""
    from oracle import adv_oracle as ao
    g = np.load(os.path.join(golden_dir, "adv_golden.npz"))
    for C in (60, 13):
        p = "c%d_" % C
        r = ao.forward_backward(g["x"], g[p + "labels"], g[p + "w1"], g[p + "b1"], g[p + "w2"], g[p + "b2"], g[p + "keep"],
                                float(g["lambda"]))
        assert abs(r["loss"] - float(g[p + "loss"])) < 1e-5 and np.array_equal(r["pred"], g[p + "pred"])
        assert np.abs(r["logits"] - g[p + "logits"]).max() < 1e-5
        for k in ("dfeat", "dw1", "db1", "dw2", "db2"):
            ref = g[p + k]
            assert np.abs(r[k] - ref).max() <= 1e-6 + 1e-4 * np.abs(ref).max(), (C, k)
        assert np.abs(g[p + "dfeat"]).max() > 0        # the reversed gradient really reaches the features


def test_reference_copy_for_the_cpu_arm_is_byte_identical():
    """oracle/_ref (built by oracle/build_ref.sh, git-ignored, travels to the GPU box) holds the reference's own hot-path
    modules unmodified; bench.py --impl reference imports them under the shim (kind "reference")."""
    import hashlib
    import subprocess
    from oracle import ref_shim
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isfile("/root/reference/feature_extraction.py"):
        pytest.skip("reference tree not mounted")
    subprocess.run(["sh", os.path.join(root, "oracle", "build_ref.sh")], check=True, capture_output=True)
    assert ref_shim.copy_available()
    for line in open(os.path.join(ref_shim.REF_COPY, "MANIFEST")).read().strip().splitlines():
        digest, name = line.split()
        for base in ("/root/reference", ref_shim.REF_COPY):
            assert hashlib.sha256(open(os.path.join(base, name), "rb").read()).hexdigest() == digest, (base, name)
    ignored = subprocess.run(["git", "check-ignore", "oracle/_ref/resnet.py"], cwd=root, capture_output=True, text=True)
    assert ignored.returncode == 0, "oracle/_ref must stay out of the history"


def test_gradient_noise_floor_of_the_reference_fp32_arithmetic():
    """Why end-to-end gradient comparisons cannot be held to the forward tolerance: the oracle (= the reference's torch
    ops) in fp32 against the SAME code in fp64.  The forward agrees to ~1e-6, the parameter gradients only to ~1e-3: an
    activation within the forward deviation of zero flips its ReLU mask, so gradients agree to ~sqrt(forward deviation).
    tests/test_parity_fp32_gpu.py quotes these numbers for the bar of its golden gradient-norm check."""
    B = 2
    spec = ss.resnet_spec()
    x = torch.from_numpy(lo.apply_frame_map(lo.lfcc(ss.seeded_waves(B, 64000, seed=3).numpy()),
                                            lo.frame_index_map(401, 750, "repeat"))).float().unsqueeze(1).transpose(2, 3)
    labels = ss.seeded_labels(B, 3)
    out = {}
    for dt in (torch.float32, torch.float64):
        sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in ss.seeded_state(spec, 11).items()}
        keys = ss.trainable_keys(spec)
        for k in keys:
            sd[k].requires_grad_(True)
        feat, _ = no.resnet_forward(sd, x.to(dt), True)
        loss, _ = no.ocsoftmax(ss.seeded_center(256, 11).to(dt), feat, labels, 0.9, 0.2, 20.0)
        loss.backward()
        out[dt] = (feat.detach().double(), {k: sd[k].grad.double() for k in keys if sd[k].grad is not None})
    rel = lambda a, b: float((a - b).norm() / b.norm())                   # noqa: E731
    fwd = rel(out[torch.float32][0], out[torch.float64][0])
    grads = sorted(rel(out[torch.float32][1][k], g) for k, g in out[torch.float64][1].items())
    median = grads[len(grads) // 2]
    print("fp32 vs fp64 reference arithmetic: feat %.1e, gradients median %.1e max %.1e" % (fwd, median, grads[-1]))
    assert fwd < 1e-5 and 10 * fwd < median < 1e-2
