"""Drop-in for evaluate_tDCF_asvspoof19.compute_eer_and_tdcf (evaluate_tDCF_asvspoof19.py:6-68): EER and
min t-DCF of a countermeasure score file against the organisers' ASV scores, with every sort / curve /
argmin on the GPU (csrc/det.cu).  The matplotlib figures of the reference (:78-117) are not produced."""
import os

import numpy as np
import torch

from . import eval_metrics as em


def cost_model_asvspoof19():
    Pspoof = 0.05                                                       # evaluate_tDCF_asvspoof19.py:10-19
    return {'Pspoof': Pspoof, 'Ptar': (1 - Pspoof) * 0.99, 'Pnon': (1 - Pspoof) * 0.01,
            'Cmiss_asv': 1, 'Cfa_asv': 10, 'Cmiss_cm': 1, 'Cfa_cm': 10}


def eer_and_tdcf(bona_cm, spoof_cm, tar_asv, non_asv, spoof_asv, cost_model=None, verbose=True):
    """The numeric part of compute_eer_and_tdcf on score arrays (numpy or torch, any device)."""
    cost_model = cost_model or cost_model_asvspoof19()
    eer_asv, asv_threshold = em.compute_eer(tar_asv, non_asv)
    Pfa_asv, Pmiss_asv, Pmiss_spoof_asv = em.obtain_asv_error_rates(tar_asv, non_asv, spoof_asv, asv_threshold)
    C1, C2 = em.tdcf_constants(Pfa_asv, Pmiss_asv, Pmiss_spoof_asv, cost_model)
    if C1 < 0 or C2 < 0:
        raise SystemExit('cannot evaluate the t-DCF with negative weights - check the ASV error rates')
    # both orientations in one go; the better EER decides which t-DCF counts (evaluate_tDCF_asvspoof19.py:45-62)
    fwd = em.det(bona_cm, spoof_cm, c1=C1, c2=C2).host()
    neg = em.det(bona_cm, spoof_cm, negate=True, c1=C1, c2=C2).host()
    pick = fwd if fwd["eer"] < neg["eer"] else neg
    eer_cm, min_tDCF = min(fwd["eer"], neg["eer"]), pick["min_tdcf"]
    if verbose:
        print('\nCM SYSTEM')
        print('   EER            = {:8.5f} % (Equal error rate for countermeasure)'.format(eer_cm * 100))
        print('\nTANDEM')
        print('   min-tDCF       = {:8.5f}'.format(min_tDCF))
    return eer_cm, min_tDCF


def compute_eer_and_tdcf(cm_score_file, path_to_database):
    asv_score_file = os.path.join(path_to_database,
                                  'LA/ASVspoof2019_LA_asv_scores/ASVspoof2019.LA.asv.eval.gi.trl.scores.txt')
    asv_data = np.genfromtxt(asv_score_file, dtype=str)
    asv_keys = asv_data[:, 1]
    asv_scores = asv_data[:, 2].astype(np.float64)
    cm_data = np.genfromtxt(cm_score_file, dtype=str)
    cm_keys = cm_data[:, 2]
    cm_scores = cm_data[:, 3].astype(np.float64)
    dev = torch.device("cuda")
    asv = torch.from_numpy(asv_scores).to(dev)
    cm = torch.from_numpy(cm_scores).to(dev)
    sel = lambda t, keys, k: t[torch.from_numpy(keys == k).to(dev)]
    return eer_and_tdcf(sel(cm, cm_keys, 'bonafide'), sel(cm, cm_keys, 'spoof'), sel(asv, asv_keys, 'target'),
                        sel(asv, asv_keys, 'nontarget'), sel(asv, asv_keys, 'spoof'))
