"""CPU restatement of the reference's detection metrics -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
file; the product path (asvspoof2021_air_b200/eval_metrics.py -> csrc/det.cu) never does.

Follows eval_metrics.py of the reference:
  det_curve            eval_metrics.py:19-37  (compute_det_curve)
  eer                  eval_metrics.py:40-46  (compute_eer)
  tdcf_curve           eval_metrics.py:160-172 (C1, C2, tDCF, tDCF_norm inside compute_tDCF)
  asv_error_rates      eval_metrics.py:4-16   (obtain_asv_error_rates)
Parity is PINNED: tests/golden/det_golden.npz holds the outputs of the reference's own functions (imported
unmodified from /root/reference by oracle/make_golden.py) and tests/test_oracle.py checks this file against
them bit for bit.

Formulated independently of the reference: the stable merge sort of concat(target, nontarget) is replaced by
a lexicographic sort on (score, class) -- inside a group of equal scores a stable sort keeps the concat order,
i.e. targets before nontargets, which is what the secondary key states explicitly -- and the two running sums
are derived from one running count of nontargets, the formulation the CUDA kernels use.
"""
import numpy as np


def det_curve(target, nontarget, negate=False):
    target = np.asarray(target, dtype=np.float64).reshape(-1)
    nontarget = np.asarray(nontarget, dtype=np.float64).reshape(-1)
    if negate:
        target, nontarget = -target, -nontarget
    scores = np.concatenate((target, nontarget))
    is_non = np.concatenate((np.zeros(target.size, np.int64), np.ones(nontarget.size, np.int64)))
    order = np.lexsort((is_non, scores))                     # primary key: score, secondary: class
    non_seen = np.cumsum(is_non[order])                      # nontargets among the k lowest scores
    k = np.arange(1, scores.size + 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        frr = np.concatenate(([0.0], (k - non_seen) / np.float64(target.size)))
        far = np.concatenate(([1.0], (nontarget.size - non_seen) / np.float64(nontarget.size)))
    thresholds = np.concatenate(([scores[order[0]] - 0.001], scores[order]))
    return frr, far, thresholds


def eer(target, nontarget, negate=False):
    frr, far, thr = det_curve(target, nontarget, negate)
    i = int(np.argmin(np.abs(frr - far)))
    return (frr[i] + far[i]) / 2.0, thr[i], i


def tdcf_weights(Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost):
    c1 = cost["Ptar"] * (cost["Cmiss_cm"] - cost["Cmiss_asv"] * Pmiss_asv) - cost["Pnon"] * cost["Cfa_asv"] * Pfa_asv
    c2 = cost["Cfa_cm"] * cost["Pspoof"] * (1 - Pmiss_spoof_asv)
    return c1, c2


def tdcf_curve(bonafide, spoof, c1, c2, negate=False):
    frr, far, thr = det_curve(bonafide, spoof, negate)
    return (c1 * frr + c2 * far) / min(c1, c2), thr


def asv_error_rates(tar, non, spoof, threshold):
    tar, non, spoof = (np.asarray(a, dtype=np.float64) for a in (tar, non, spoof))
    pfa = np.count_nonzero(non >= threshold) / non.size
    pmiss = np.count_nonzero(tar < threshold) / tar.size
    pmiss_spoof = None if spoof.size == 0 else np.count_nonzero(spoof < threshold) / spoof.size
    return pfa, pmiss, pmiss_spoof
