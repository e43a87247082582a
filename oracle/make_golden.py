"""Generate tests/golden/*.npz by running the UNMODIFIED reference (under oracle/ref_shim.py)
on seeded synthetic inputs.  Run in the authoring container only:

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The golden files are what the GPU box (no /root/reference) checks
the oracle restatement and the CUDA path against.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim, state_spec  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SPARSE_FRAMES = [0, 1, 2, 3, 100, 200, 397, 398, 399, 400]


def ref_lfcc(wave):
    fe = ref_shim.load("feature_extraction")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = fe.LFCC(320, 160, 512, 16000, 20, with_energy=False)
        return mod(wave.clone()), mod      # the reference mutates its input in place


def golden_lfcc():
    out = {}
    # config[0]: 32 seeded 4 s waves + 3 edge rows
    w = state_spec.seeded_waves(32, 64000, seed=0, edge_rows=True)
    y, mod = ref_lfcc(w)
    y = y.numpy()
    out["full_rows"] = np.array([0, 1, 2, 31, 32, 33, 34])
    out["full"] = y[out["full_rows"]]
    out["sparse_frames"] = np.array(SPARSE_FRAMES)
    out["sparse"] = y[:, SPARSE_FRAMES, :]
    out["colsum"] = y.astype(np.float64).sum(axis=1)           # (35,60) checksum over frames
    out["lfcc_fb"] = mod.lfcc_fb.detach().numpy()
    out["dct_weight"] = mod.l_dct.weight.detach().numpy()
    # ragged lengths: T = 1 + L//160
    for L in (3200, 12345, 160 * 37, 160 * 37 + 159, 321, 800):
        wl = state_spec.seeded_waves(2, L, seed=L)
        out["ragged_%d" % L] = ref_lfcc(wl)[0].numpy()
    # the silence vector of dataset.py:13-16
    out["silence"] = ref_lfcc(torch.zeros(1, 3200))[0][:, 0, :].numpy()
    np.savez_compressed(os.path.join(GOLD, "lfcc_golden.npz"), **out)
    print("lfcc_golden.npz", {k: v.shape for k, v in out.items()})


def golden_padcrop():
    """Run the reference's own pad helpers (dataset.py:513-528) on index-valued tensors."""
    import importlib.util
    # dataset.py imports librosa + builds an LFCC at import: fine under the shim
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ds = ref_shim.load("dataset")
    out = {}
    for T in (1, 21, 401, 749):
        spec = torch.arange(T, dtype=torch.float32).reshape(1, T, 1).repeat(1, 1, 60)
        out["repeat_%d" % T] = ds.repeat_padding_Tensor(spec, 750)[0, :, 0].numpy().astype(np.int64)
        z = ds.padding_Tensor(spec + 1, 750)[0, :, 0].numpy().astype(np.int64) - 1   # -1 where zero
        out["zero_%d" % T] = z
        s = ds.silence_padding_Tensor(spec + 1000, 750)[0, :, 0].numpy()
        out["silence_%d" % T] = s       # silence rows carry c0 of the silence vector (< 0)
    out["silence_pad_value"] = ds.silence_pad_value.numpy()
    np.savez_compressed(os.path.join(GOLD, "padcrop_golden.npz"), **out)
    print("padcrop_golden.npz", list(out))


def _features(batch, seed):
    """(B,750,60) float32 model input: float64 oracle LFCC (deterministic numpy arithmetic) of the
    seeded waves, repeat-padded (dataset.py:519-522) and rounded to float32.  The reference
    nets then run on exactly the tensor the tests can regenerate anywhere: the nets amplify
    1e-6 input perturbations to ~1% on some deep gradients, so the input must be identical."""
    from oracle import lfcc_oracle as lo
    y = lo.lfcc(state_spec.seeded_waves(batch, 64000, seed=seed).numpy())
    y = lo.apply_frame_map(y, lo.frame_index_map(y.shape[1], 750, "repeat"))
    return torch.from_numpy(y).float()


def _grad_summary(named_params):
    keys, norms, sums, heads = [], [], [], []
    for k, p in named_params:
        if p.grad is None:
            continue
        g = p.grad.detach().double().reshape(-1)
        keys.append(k)
        norms.append(float(g.norm()))
        sums.append(float(g.sum()))
        h = np.zeros(8)
        h[:min(8, g.numel())] = g[:8].numpy()
        heads.append(h)
    return np.array(keys), np.array(norms), np.array(sums), np.stack(heads)


def golden_nets(batch=4, seed=3):
    rn = ref_shim.load("resnet")
    ec = ref_shim.load("ecapa_tdnn")
    ls = ref_shim.load("loss")
    feats = _features(batch, seed)
    labels = state_spec.seeded_labels(batch, seed)
    out = {"labels": labels.numpy(), "batch": batch, "seed": seed}

    real_randn = torch.randn
    for arch in ("resnet", "ecapa"):
        if arch == "resnet":
            model = rn.ResNet(3, 256, "18", nclasses=2)
            spec = state_spec.resnet_spec()
            x = feats.unsqueeze(1).transpose(2, 3).contiguous()          # main_train.py:338
        else:
            model = ec.Res2Net2(ec.Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60)
            spec = state_spec.ecapa_spec()
            x = feats.transpose(1, 2).contiguous()                       # main_train.py:338,347-348
        ref_keys = list(model.state_dict().keys())
        assert ref_keys == [k for k, _, _ in spec], "state spec mismatch for " + arch
        model.load_state_dict(state_spec.seeded_state(spec, seed=11))
        loss_mod = ls.AngularIsoLoss(256, r_real=0.9, r_fake=0.2, alpha=20.0)
        with torch.no_grad():
            loss_mod.center.copy_(state_spec.seeded_center(256, seed=11))
        model.train()
        torch.randn = lambda *a, **k: torch.zeros(*a, **k)                # SelfAttention noise -> 0
        try:
            feat, logits = model(x)
        finally:
            torch.randn = real_randn
        loss, score = loss_mod(feat, labels)
        ce = torch.nn.functional.cross_entropy(logits, labels)
        loss.backward()
        keys, norms, sums, heads = _grad_summary(list(model.named_parameters()))
        out[arch + "_feat"] = feat.detach().numpy()
        out[arch + "_logits"] = logits.detach().numpy()
        out[arch + "_loss"] = float(loss)
        out[arch + "_ce"] = float(ce)
        out[arch + "_score"] = score.detach().numpy()
        out[arch + "_grad_keys"] = keys
        out[arch + "_grad_norm"] = norms
        out[arch + "_grad_sum"] = sums
        out[arch + "_grad_head"] = heads
        out[arch + "_center_grad"] = loss_mod.center.grad.numpy()
        # running stats after the one train-mode forward
        sd = model.state_dict()
        rk = [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]
        out[arch + "_running_keys"] = np.array(rk)
        out[arch + "_running_sum"] = np.array([float(sd[k].double().sum()) for k in rk])
        # eval-mode forward with the updated running stats (the scoring path)
        model.eval()
        torch.randn = lambda *a, **k: torch.zeros(*a, **k)
        try:
            with torch.no_grad():
                feat_e, logits_e = model(x)
                _, score_e = loss_mod(feat_e, torch.zeros(batch))
        finally:
            torch.randn = real_randn
        out[arch + "_eval_feat"] = feat_e.numpy()
        out[arch + "_eval_logits"] = logits_e.numpy()
        out[arch + "_eval_score"] = (-score_e).numpy()       # generate_score.py:117 writes -score
        print(arch, "loss", float(loss), "ce", float(ce), "ngrads", len(keys))
    np.savez_compressed(os.path.join(GOLD, "nets_golden.npz"), **out)


DET_COST = {"Pspoof": 0.05, "Ptar": 0.95 * 0.99, "Pnon": 0.95 * 0.01, "Cmiss_asv": 1, "Cfa_asv": 10,
            "Cmiss_cm": 1, "Cfa_cm": 10}                      # evaluate_tDCF_asvspoof19.py:10-19
DET_ASV = (0.02372, 0.02478, 0.61542)                         # Pfa_asv, Pmiss_asv, Pmiss_spoof_asv of a typical ASV


def det_cases():
    """Seeded score sets for the detection metrics: name -> (target, nontarget) in the dtype the case is about."""
    rng = np.random.RandomState(2021)
    cases = {}
    ref_file = os.path.join(ref_shim.REFERENCE_ROOT, "scores", "lfcc_ecapa512ctst_ocs_19dev_score.txt")
    rows = np.genfromtxt(ref_file, dtype=str)[::7]            # every 7th trial of the reference's own dev scores
    sc, key = rows[:, 1].astype(np.float64), rows[:, 2]
    cases["devfile_f64"] = (sc[key == "bonafide"], sc[key == "spoof"])
    cases["devfile_f32"] = (sc[key == "bonafide"].astype(np.float32), sc[key == "spoof"].astype(np.float32))
    t = np.round(rng.randn(700) * 0.5 + 0.6, 2)
    n = np.round(rng.randn(3300) * 0.5 - 0.2, 2)
    cases["ties_f32"] = (t.astype(np.float32), n.astype(np.float32))      # ~150 distinct values, many cross-class ties
    cases["ties_f64"] = (t, n)
    cases["normal_f32"] = ((rng.randn(2548) + 1.5).astype(np.float32), (rng.randn(22296) - 1.0).astype(np.float32))
    cases["f64_dense"] = (rng.randn(1500) * 1e-9 + 1.0, rng.randn(2600) * 1e-9 + 1.0)   # differ only in low mantissa bits
    cases["one_each"] = (np.array([0.25], np.float32), np.array([-0.5], np.float32))
    cases["inverted"] = (np.array([-1.0, -2.0, -3.0], np.float32), np.array([0.5], np.float32))
    cases["all_equal"] = (np.full(5, 0.125, np.float32), np.full(9, 0.125, np.float32))
    cases["signed_zero"] = (np.array([0.0, -0.0, 1.0, -1.0], np.float32), np.array([-0.0, 0.0, -1.0, 2.0], np.float32))
    cases["tile_edge"] = ((rng.rand(1000) * 2 - 1).astype(np.float32), (rng.rand(1049) * 2 - 1.2).astype(np.float32))
    cases["subnormal"] = ((rng.randn(40) * 1e-41).astype(np.float32), (rng.randn(60) * 1e-41).astype(np.float32))
    return cases


def golden_det():
    """Outputs of the reference's own eval_metrics functions (imported unmodified) on det_cases()."""
    sys.path.insert(0, ref_shim.REFERENCE_ROOT)
    import eval_metrics as rem
    out = {}
    for name, (tar, non) in det_cases().items():
        out[name + "__target"], out[name + "__nontarget"] = tar, non
        for tag, sign in (("", 1.0), ("_neg", -1.0)):
            t, n = (sign * tar).astype(tar.dtype), (sign * non).astype(non.dtype)
            frr, far, thr = rem.compute_det_curve(t, n)
            e, eth = rem.compute_eer(t, n)
            out[name + tag + "__eer"] = np.array([e, float(eth), float(np.argmin(np.abs(frr - far)))])
            out[name + tag + "__sums"] = np.array([frr.sum(), far.sum()])
            if frr.size <= 4200:
                out[name + tag + "__frr"], out[name + tag + "__far"] = frr, far
                out[name + tag + "__thr"] = thr[1:].astype(np.float64)
            if tag == "" and np.unique(np.concatenate((t, n))).size >= 3:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    curve, cthr = rem.compute_tDCF(t, n, DET_ASV[0], DET_ASV[1], DET_ASV[2], DET_COST, False)
                i = int(np.argmin(curve))
                out[name + "__tdcf"] = np.array([curve[i], float(cthr[i]), float(i), curve.sum()])
                if curve.size <= 4200:
                    out[name + "__tdcf_curve"] = curve
    rng = np.random.RandomState(7)
    tar_asv, non_asv, spoof_asv = rng.randn(900) * 2 + 3, rng.randn(1100) * 2 - 3, rng.randn(1300) * 2 + 1
    non_asv[:40] = tar_asv[:40]                                # exact cross-array ties around the threshold
    e, thr = rem.compute_eer(tar_asv, non_asv)
    rates = rem.obtain_asv_error_rates(tar_asv, non_asv, spoof_asv, thr)
    out["asv__tar"], out["asv__non"], out["asv__spoof"] = tar_asv, non_asv, spoof_asv
    out["asv__expected"] = np.array([e, thr] + [float(r) for r in rates])
    np.savez_compressed(os.path.join(GOLD, "det_golden.npz"), **out)
    print("det_golden.npz", len(out), "arrays,", os.path.getsize(os.path.join(GOLD, "det_golden.npz")), "bytes")


def golden_adv(batch=16, enc=256, classes=(60, 13), lambda_=0.05, seed=11):
    """The reference ChannelClassifier (model.py:1002-1023) + CrossEntropyLoss with a FIXED dropout keep-mask injected
    into torch.nn.functional.dropout, for the two head sizes of the LAPA / DFPA branch (60 codecs, 13 devices)."""
    md = ref_shim.load("model")
    import torch.nn.functional as F
    out = {"lambda": np.array(lambda_)}
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, enc, generator=g)
    out["x"] = x.numpy()
    real_dropout = F.dropout
    for C in classes:
        torch.manual_seed(seed + C)
        clf = md.ChannelClassifier(enc, C, lambda_)
        clf.train()
        labels = torch.randint(0, C, (batch,), generator=g)
        keep = (torch.rand(batch, enc // 2, generator=g) >= 0.3).float()
        F.dropout = lambda inp, p=0.5, training=True, inplace=False: inp * keep / (1.0 - p)
        try:
            xr = x.clone().requires_grad_(True)
            logits = clf(xr)
            loss = torch.nn.CrossEntropyLoss()(logits, labels)
            loss.backward()
        finally:
            F.dropout = real_dropout
        sd = clf.state_dict()
        pre = "c%d_" % C
        out[pre + "labels"], out[pre + "keep"] = labels.numpy(), keep.numpy()
        for k, name in (("classifier.0.weight", "w1"), ("classifier.0.bias", "b1"), ("classifier.3.weight", "w2"), ("classifier.3.bias", "b2")):
            out[pre + name] = sd[k].numpy()
        out[pre + "logits"], out[pre + "loss"] = logits.detach().numpy(), np.array(float(loss))
        out[pre + "pred"] = torch.max(logits.data, 1)[1].numpy()
        out[pre + "dfeat"] = xr.grad.numpy()
        ps = dict(clf.named_parameters())
        for k, name in (("classifier.0.weight", "dw1"), ("classifier.0.bias", "db1"), ("classifier.3.weight", "dw2"), ("classifier.3.bias", "db2")):
            out[pre + name] = ps[k].grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "adv_golden.npz"), **out)
    print("adv_golden.npz", os.path.getsize(os.path.join(GOLD, "adv_golden.npz")), "bytes")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    only = [a[2:-5] for a in sys.argv[1:] if a.startswith("--") and a.endswith("-only")]
    if not only:
        golden_lfcc()
        golden_padcrop()
        golden_nets()
    if not only or "det" in only:
        golden_det()
    if not only or "adv" in only:
        golden_adv()
