"""Drop-in for the reference's eval_metrics.py (compute_det_curve / compute_eer / compute_tDCF /
obtain_asv_error_rates, eval_metrics.py:4-46,49-193) with the sort, the running counts, the error-rate
curves and both argmins on the GPU (csrc/det.cu) -- SURVEY.md section 8(f) row 3.

Same names, argument meaning, return values and error behaviour (sys.exit messages) as the reference.
Inputs may be numpy arrays (what the reference passes) or torch tensors; CUDA tensors are consumed where
they are, without a host round trip.  float32 scores (what the model produces) and float64 scores (what
np.genfromtxt reads from a score file) are both sorted as exact fp64 values, so every returned number is
bit-identical to the reference's numpy result for the same input dtype, with two exceptions in the thresholds:
numpy >= 2 computes `thresholds[0] = min_score - 0.001` in float32 for float32 input (here always float64), and
a score of -0.0 is reported as +0.0 (the two tie in the sort).

There is no CPU path: without a CUDA device or libair_b200.so every function raises."""
import sys

import numpy as np
import torch

from . import _lib, ops


def _device_scores(a, b):
    """Both score vectors as contiguous CUDA tensors of one dtype (float32 unless either side is float64)."""
    if not torch.cuda.is_available():
        raise _lib.AirError("eval_metrics runs on CUDA only (csrc/det.cu); there is no CPU path")
    ts = []
    for x in (a, b):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
        if not t.dtype.is_floating_point:
            t = t.double()
        ts.append(t.detach().reshape(-1))
    dt = torch.float64 if any(t.dtype == torch.float64 for t in ts) else torch.float32
    dev = next((t.device for t in ts if t.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    return [t.to(device=dev, dtype=dt).contiguous() for t in ts]


class DetResult:
    """Device-side result of one air_det_curve call; `.host()` reads the ten summary doubles (one D2H copy)."""

    def __init__(self, out, frr, far, thresholds, tdcf):
        self.out, self.frr, self.far, self.thresholds, self.tdcf = out, frr, far, thresholds, tdcf

    def host(self):
        o = self.out.cpu().numpy()
        return {"eer": float(o[0]), "eer_threshold": float(o[1]), "frr": float(o[2]), "far": float(o[3]),
                "min_tdcf": float(o[4]), "tdcf_threshold": float(o[5]), "n_target": int(o[6]),
                "n_nontarget": int(o[7]), "eer_index": int(o[8]), "tdcf_index": int(o[9])}


def det(target_scores, nontarget_scores, negate=False, c1=None, c2=None, curves=False, workspace=None):
    """One pass over the scores: sort, curve, EER (and min normalised t-DCF when c1, c2 are given)."""
    tar, non = _device_scores(target_scores, nontarget_scores)
    n = tar.numel() + non.numel()
    if n < 1:
        raise ValueError("compute_det_curve needs at least one score")
    need = ops.det_workspace_bytes(n)
    if workspace is None or workspace.numel() * workspace.element_size() < need or workspace.device != tar.device:
        workspace = torch.empty(need, dtype=torch.uint8, device=tar.device)
    want = c1 is not None and c2 is not None
    mk = (lambda: torch.empty(n + 1, dtype=torch.float64, device=tar.device)) if curves else (lambda: None)
    frr, far, thr = mk(), mk(), mk()
    tdcf = mk() if want else None
    out = torch.empty(10, dtype=torch.float64, device=tar.device)
    with torch.cuda.device(tar.device):
        ops.det_curve(tar, non, negate, float(c1) if want else 0.0, float(c2) if want else 0.0, want, workspace,
                      frr, far, thr, tdcf, out)
    return DetResult(out, frr, far, thr, tdcf)


def obtain_asv_error_rates(tar_asv, non_asv, spoof_asv, asv_threshold):
    """eval_metrics.py:4-16."""
    res = []
    for scores, want_ge in ((non_asv, True), (tar_asv, False), (spoof_asv, False)):
        size = scores.numel() if isinstance(scores, torch.Tensor) else np.asarray(scores).size
        if size == 0:
            if scores is spoof_asv:
                res.append(None)
                continue
            res.append(float("nan"))                                     # numpy: 0 / 0
            continue
        s, _ = _device_scores(scores, scores[:0])
        counts = torch.empty(2, dtype=torch.int64, device=s.device)
        with torch.cuda.device(s.device):
            ops.det_threshold_counts(s, float(asv_threshold), counts)
        c = counts.cpu().numpy()
        res.append(float(c[0] if want_ge else c[1]) / size)
    Pfa_asv, Pmiss_asv, Pmiss_spoof_asv = res
    return Pfa_asv, Pmiss_asv, Pmiss_spoof_asv


def compute_det_curve(target_scores, nontarget_scores):
    """eval_metrics.py:19-37 -> (frr, far, thresholds) as float64 numpy arrays of n + 1 points."""
    r = det(target_scores, nontarget_scores, curves=True)
    return r.frr.cpu().numpy(), r.far.cpu().numpy(), r.thresholds.cpu().numpy()


def compute_eer(target_scores, nontarget_scores):
    """eval_metrics.py:40-46 -> (eer, threshold).  Only ten doubles cross to the host."""
    h = det(target_scores, nontarget_scores).host()
    return h["eer"], h["eer_threshold"]


def tdcf_constants(Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model):
    """C1, C2 of eval_metrics.py:160-162 (host scalars)."""
    C1 = cost_model['Ptar'] * (cost_model['Cmiss_cm'] - cost_model['Cmiss_asv'] * Pmiss_asv) - \
        cost_model['Pnon'] * cost_model['Cfa_asv'] * Pfa_asv
    C2 = cost_model['Cfa_cm'] * cost_model['Pspoof'] * (1 - Pmiss_spoof_asv)
    return C1, C2


def compute_tDCF(bonafide_score_cm, spoof_score_cm, Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model, print_cost):
    """eval_metrics.py:49-193 -> (tDCF_norm curve, CM_thresholds); same sanity checks and exits."""
    if cost_model['Cfa_asv'] < 0 or cost_model['Cmiss_asv'] < 0 or \
            cost_model['Cfa_cm'] < 0 or cost_model['Cmiss_cm'] < 0:
        print('WARNING: Usually the cost values should be positive!')
    if cost_model['Ptar'] < 0 or cost_model['Pnon'] < 0 or cost_model['Pspoof'] < 0 or \
            np.abs(cost_model['Ptar'] + cost_model['Pnon'] + cost_model['Pspoof'] - 1) > 1e-10:
        sys.exit('ERROR: Your prior probabilities should be positive and sum up to one.')
    if Pmiss_spoof_asv is None:
        sys.exit('ERROR: you should provide miss rate of spoof tests against your ASV system.')
    bona, spoof = _device_scores(bonafide_score_cm, spoof_score_cm)
    combined = torch.cat((bona, spoof))
    if not bool(torch.isfinite(combined).all()):
        sys.exit('ERROR: Your scores contain nan or inf.')
    C1, C2 = tdcf_constants(Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model)
    if C1 < 0 or C2 < 0:
        sys.exit('You should never see this error but I cannot evalute tDCF with negative weights - please check '
                 'whether your ASV error rates are correctly computed?')
    r = det(bona, spoof, c1=C1, c2=C2, curves=True)
    thr = r.thresholds.cpu().numpy()
    if np.unique(thr[1:]).size < 3:
        sys.exit('ERROR: You should provide soft CM scores - not binary decisions')
    if print_cost:
        print('t-DCF evaluation from [Nbona={}, Nspoof={}] trials\n'.format(bona.numel(), spoof.numel()))
        if C2 == np.minimum(C1, C2):
            print('   tDCF_norm(s) = {:8.5f} x Pmiss_cm(s) + Pfa_cm(s)\n'.format(C1 / C2))
        else:
            print('   tDCF_norm(s) = Pmiss_cm(s) + {:8.5f} x Pfa_cm(s)\n'.format(C2 / C1))
    return r.tdcf.cpu().numpy(), thr
