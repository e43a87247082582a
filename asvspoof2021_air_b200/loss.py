"""Drop-ins for the reference's OCSoftmax and AngularIsoLoss (loss.py:176-206 and :62-97, identical
bodies; AngularIsoLoss is what `--add_loss ang_iso` instantiates, main_train.py:269-272).

forward(x (B,D), labels (B,)) -> (loss, -cos (B,)); parameter `center` (1,D).  One fused sm_100a
kernel computes the loss, the scores and both analytic gradients (csrc/head.cu)."""
import torch
import torch.nn as nn

from . import _lib, ops


class _OCSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, center, labels, r_real, r_fake, alpha):
        B, D = x.shape
        x = x.contiguous().float()
        loss = torch.empty(1, device=x.device)
        score = torch.empty(B, device=x.device)
        dfeat = torch.empty(B, D, device=x.device)
        dcenter = torch.zeros(1, D, device=x.device)
        lab = labels.to(device=x.device, dtype=torch.long).contiguous()
        ops.ocsoftmax(x, lab, center.contiguous(), B, D, r_real, r_fake, alpha, 1.0, loss, score, dfeat, dcenter)
        ctx.save_for_backward(dfeat, dcenter)
        ctx.mark_non_differentiable(score)
        return loss[0], score

    @staticmethod
    def backward(ctx, gloss, gscore):
        dfeat, dcenter = ctx.saved_tensors
        return dfeat * gloss, dcenter * gloss, None, None, None, None


class OCSoftmax(nn.Module):
    def __init__(self, feat_dim=2, r_real=0.9, r_fake=0.5, alpha=20.0):
        super().__init__()
        self.feat_dim, self.r_real, self.r_fake, self.alpha = feat_dim, r_real, r_fake, alpha
        self.center = nn.Parameter(torch.randn(1, self.feat_dim))
        nn.init.kaiming_uniform_(self.center, 0.25)           # loss.py:183-184
        self.softplus = nn.Softplus()                         # kept for state/pickle parity; unused

    def forward(self, x, labels):
        if not x.is_cuda:
            raise _lib.AirError("OCSoftmax runs on CUDA only (no CPU path); got a %s tensor" % x.device)
        return _OCSoftmaxFn.apply(x, self.center, labels, float(self.r_real), float(self.r_fake), float(self.alpha))


class AngularIsoLoss(OCSoftmax):
    """loss.py:62-97 -- same computation under the name main_train.py uses."""
