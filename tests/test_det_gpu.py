"""GPU parity tests of the detection-metric kernels (csrc/det.cu) through the C ABI: bit-for-bit against the
outputs of the reference's eval_metrics functions (tests/golden/det_golden.npz) and against the numpy oracle."""
import numpy as np
import pytest
import torch

import det_cases as dc
from asvspoof2021_air_b200 import _lib, ops
from asvspoof2021_air_b200 import eval_metrics as em
from asvspoof2021_air_b200 import evaluate_tDCF_asvspoof19 as ev
from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return dc.load(golden_dir)


def _weights():
    return mo.tdcf_weights(*dc.DET_ASV, dc.DET_COST)


def test_curves_eer_tdcf_bit_identical_to_reference_golden(g):
    c1, c2 = _weights()
    for name in dc.case_names(g):
        tar, non = g[name + "__target"], g[name + "__nontarget"]
        for tag, neg in (("", False), ("_neg", True)):
            r = em.det(tar, non, negate=neg, c1=c1, c2=c2, curves=True)
            h = r.host()
            frr, far, thr = r.frr.cpu().numpy(), r.far.cpu().numpy(), r.thresholds.cpu().numpy()
            want = g[name + tag + "__eer"]
            assert dc.same(h["eer"], want[0]) and h["eer_index"] == int(want[2]), (name, tag, h, want)
            if h["eer_index"] > 0 or tar.dtype == np.float64:
                assert dc.same_value(h["eer_threshold"], want[1]), (name, tag)
            assert h["n_target"] == tar.size and h["n_nontarget"] == non.size
            assert dc.same([frr.sum(), far.sum()], g[name + tag + "__sums"]), (name, tag)
            if name + tag + "__frr" in g.files:
                assert dc.same(frr, g[name + tag + "__frr"]) and dc.same(far, g[name + tag + "__far"]), (name, tag)
                assert dc.same_value(thr[1:], g[name + tag + "__thr"]), (name, tag)
            assert dc.same(thr[0], np.float64(thr[1]) - 0.001)
            if tag == "" and name + "__tdcf" in g.files:
                want = g[name + "__tdcf"]
                curve = r.tdcf.cpu().numpy()
                assert dc.same(h["min_tdcf"], want[0]) and h["tdcf_index"] == int(want[2]), (name, h, want)
                assert dc.same(curve.sum(), want[3]), name
                if h["tdcf_index"] > 0:
                    assert dc.same_value(h["tdcf_threshold"], want[1]), name
                if name + "__tdcf_curve" in g.files:
                    assert dc.same(curve, g[name + "__tdcf_curve"]), name


@pytest.mark.parametrize("n_tar,n_non,decimals,dtype", [
    (7355, 63882, None, np.float32),          # ASVspoof 2019 LA eval trial counts
    (7355, 63882, 2, np.float32),             # heavy cross-class ties
    (2049, 1, None, np.float64),
    (1, 4097, 1, np.float32),
    (300001, 700002, 3, np.float32),          # > 488 tiles, ties everywhere
    (40000, 90000, None, np.float64),
])
def test_full_size_against_oracle_and_curve_properties(n_tar, n_non, decimals, dtype):
    rng = np.random.RandomState(n_tar % 1000 + n_non % 77)
    tar, non = rng.randn(n_tar) + 1.0, rng.randn(n_non) - 1.0
    if decimals is not None:
        tar, non = np.round(tar, decimals), np.round(non, decimals)
    tar, non = tar.astype(dtype), non.astype(dtype)
    c1, c2 = _weights()
    for neg in (False, True):
        r = em.det(torch.from_numpy(tar).cuda(), torch.from_numpy(non).cuda(), negate=neg, c1=c1, c2=c2, curves=True)
        h = r.host()
        frr, far, thr = r.frr.cpu().numpy(), r.far.cpu().numpy(), r.thresholds.cpu().numpy()
        ofrr, ofar, othr = mo.det_curve(tar, non, negate=neg)
        assert dc.same(frr, ofrr) and dc.same(far, ofar) and dc.same_value(thr, othr)
        e, eth, idx = mo.eer(tar, non, negate=neg)
        assert dc.same(h["eer"], e) and dc.same_value(h["eer_threshold"], eth) and h["eer_index"] == idx
        curve, _ = mo.tdcf_curve(tar, non, c1, c2, negate=neg)
        assert dc.same(r.tdcf.cpu().numpy(), curve)
        assert h["tdcf_index"] == int(np.argmin(curve)) and dc.same(h["min_tdcf"], curve.min())
        # size-independent properties of a detection-error curve
        assert (np.diff(thr) >= 0).all() and (np.diff(frr) >= 0).all() and (np.diff(far) <= 0).all()
        assert frr[0] == 0.0 and far[0] == 1.0 and frr[-1] == 1.0 and far[-1] == 0.0
        assert np.array_equal(np.sort(thr[1:]), np.sort((-1.0 if neg else 1.0) * np.concatenate((tar, non)).astype(np.float64)))


def test_reference_named_functions_and_input_kinds(g):
    tar, non = g["devfile_f32__target"], g["devfile_f32__nontarget"]
    want = g["devfile_f32__eer"]
    for a, b in ((tar, non), (torch.from_numpy(tar), torch.from_numpy(non)),
                 (torch.from_numpy(tar).cuda(), torch.from_numpy(non).cuda())):
        e, thr = em.compute_eer(a, b)
        assert dc.same(e, want[0]) and dc.same_value(thr, want[1])
    frr, far, thr = em.compute_det_curve(tar, non)
    assert dc.same(frr, g["devfile_f32__frr"]) and dc.same(far, g["devfile_f32__far"])
    curve, cthr = em.compute_tDCF(tar, non, *dc.DET_ASV, dc.DET_COST, False)
    assert dc.same(curve, g["devfile_f32__tdcf_curve"]) and dc.same_value(cthr[1:], g["devfile_f32__thr"])
    # main_train.py:662-664: eer = min(eer, other_eer)
    other = em.compute_eer(-tar, -non)[0]
    assert dc.same(other, g["devfile_f32_neg__eer"][0])
    assert dc.same(em.det(tar, non, negate=True).host()["eer"], other)


def test_error_behaviour_matches_reference(g):
    tar, non = g["devfile_f32__target"], g["devfile_f32__nontarget"]
    with pytest.raises(SystemExit, match="soft CM scores"):
        em.compute_tDCF(np.ones(4, np.float32), np.zeros(6, np.float32), *dc.DET_ASV, dc.DET_COST, False)
    with pytest.raises(SystemExit, match="nan or inf"):
        em.compute_tDCF(np.array([1.0, np.nan, 0.5]), non, *dc.DET_ASV, dc.DET_COST, False)
    with pytest.raises(SystemExit, match="miss rate of spoof"):
        em.compute_tDCF(tar, non, 0.02, 0.02, None, dc.DET_COST, False)
    bad = dict(dc.DET_COST, Ptar=0.5)
    with pytest.raises(SystemExit, match="prior probabilities"):
        em.compute_tDCF(tar, non, *dc.DET_ASV, bad, False)
    # C-ABI argument checks: workspace too small, no scores
    t = torch.from_numpy(tar).cuda()
    n = torch.from_numpy(non).cuda()
    out = torch.empty(10, dtype=torch.float64, device="cuda")
    small = torch.empty(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.AirError, match="argument error"):
        ops.det_curve(t, n, False, 0.0, 0.0, False, small, None, None, None, None, out)
    with pytest.raises(_lib.AirError, match="argument error"):
        ops.det_curve(t[:0], n[:0], False, 0.0, 0.0, False, small, None, None, None, None, out)


def test_asv_error_rates_and_tandem_evaluation(g):
    tar, non, spoof = g["asv__tar"], g["asv__non"], g["asv__spoof"]
    e, thr = em.compute_eer(tar, non)
    rates = em.obtain_asv_error_rates(tar, non, spoof, thr)
    assert dc.same([e] + list(rates), g["asv__expected"][[0, 2, 3, 4]]) and dc.same_value(thr, g["asv__expected"][1])
    assert em.obtain_asv_error_rates(tar, non, spoof[:0], thr)[2] is None
    # evaluate_tDCF_asvspoof19.py:45-62 on arrays: the orientation with the lower EER supplies the t-DCF
    bona, sp = g["normal_f32__target"], g["normal_f32__nontarget"]
    for flip in (1.0, -1.0):
        b, s = (flip * bona).astype(np.float32), (flip * sp).astype(np.float32)
        eer_cm, min_tdcf = ev.eer_and_tdcf(b, s, tar, non, spoof, verbose=False)
        c1, c2 = mo.tdcf_weights(*rates, ev.cost_model_asvspoof19())
        e_fwd, e_neg = mo.eer(b, s)[0], mo.eer(b, s, negate=True)[0]
        curve, _ = mo.tdcf_curve(b, s, c1, c2, negate=not (e_fwd < e_neg))
        assert dc.same(eer_cm, min(e_fwd, e_neg)) and dc.same(min_tdcf, curve.min())
