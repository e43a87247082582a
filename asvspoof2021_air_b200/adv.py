"""Adversarial channel-classifier head of the `--ADV_AUG` branch.

Drop-in for `ChannelClassifier` / `GradientReversal` (model.py:976-1023) as main_train.py:211-224,377-403,420-453 uses
them: a gradient-reversal layer and a two-layer MLP on the 256-d embedding, trained with CrossEntropyLoss on the channel
(codec / device) label.  Per training step the reference

  1. adds CE(classifier(feats), channel) to the feature loss, so that -lambda * dCE/dfeats flows into the encoder
     (`head_loss_and_feat_grad`; the classifier's own gradients of this pass are discarded by the later zero_grad()),
  2. runs the encoder a second time, detaches the features and takes one Adam step of the classifier alone
     (`classifier_step`).

The two nn.Linear run on air_linear_fwd / air_linear_bwd, the rest on csrc/adv.cu; parameters live in one flat fp32
buffer with Adam moments beside it, exposed under the reference's state_dict keys (classifier.0.*, classifier.3.*).

Pinned on the CPU by oracle/adv_oracle.py against the reference module and on a B200 by tests/test_adv_gpu.py
(golden from the reference, oracle with an injected dropout mask, mask statistics, Adam step, inside a train step).
"""
import math

import torch
import torch.nn as nn

from . import _lib, ops

DROPOUT_P = 0.3                      # nn.Dropout(0.3), model.py:1008


def _require_cuda(t):
    if not t.is_cuda:
        raise _lib.AirError("ChannelClassifier runs on CUDA only (no CPU path); got a %s tensor" % t.device)


class ChannelClassifier(nn.Module):
    def __init__(self, enc_dim, nclasses, lambda_, device=None):
        super().__init__()
        self.enc_dim, self.hidden, self.nclasses, self.lambda_ = int(enc_dim), int(enc_dim) // 2, int(nclasses), float(lambda_)
        device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        E, H, C = self.enc_dim, self.hidden, self.nclasses
        self.sizes = [("classifier.0.weight", (H, E)), ("classifier.0.bias", (H,)),
                      ("classifier.3.weight", (C, H)), ("classifier.3.bias", (C,))]
        n = sum(math.prod(s) for _, s in self.sizes)
        self.flat = torch.zeros(n, device=device)
        self.grad = torch.zeros(n, device=device)
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device)
        self.step_count = 0
        self._views, self._gviews, o = {}, {}, 0
        for name, shape in self.sizes:
            k = math.prod(shape)
            self._views[name] = self.flat[o:o + k].view(shape)
            self._gviews[name] = self.grad[o:o + k].view(shape)
            o += k
        with torch.no_grad():                                            # nn.Linear default init (model.py:1006-1010)
            for w, b in (("classifier.0.weight", "classifier.0.bias"), ("classifier.3.weight", "classifier.3.bias")):
                nn.init.kaiming_uniform_(self._views[w], a=math.sqrt(5))
                bound = 1.0 / math.sqrt(self._views[w].shape[1])
                nn.init.uniform_(self._views[b], -bound, bound)
        self._work = {}

    # ---- reference surface -----------------------------------------------------------------------
    def state_dict(self, *args, **kwargs):
        return {k: v.detach().clone() for k, v in self._views.items()}

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._views if k not in sd]
        extra = [k for k in sd if k not in self._views]
        if strict and (missing or extra):
            raise KeyError("ChannelClassifier state_dict mismatch: missing %s, unexpected %s" % (missing, extra))
        with torch.no_grad():
            for k, v in self._views.items():
                if k in sd:
                    v.copy_(sd[k])
        return self

    def initialize_params(self):
        """model.py:1014-1017 (never called by main_train.py)."""
        with torch.no_grad():
            nn.init.kaiming_uniform_(self._views["classifier.0.weight"])
            nn.init.kaiming_uniform_(self._views["classifier.3.weight"])

    # ---- compute -----------------------------------------------------------------------------------
    def _buffers_for(self, B, dev):
        w = self._work.get(B)
        if w is None or w["h"].device != dev:
            H, C = self.hidden, self.nclasses
            w = {"h": torch.empty(B, H, device=dev), "a": torch.empty(B, H, device=dev),
                 "keep": torch.empty(B, H, dtype=torch.uint8, device=dev), "z": torch.empty(B, C, device=dev),
                 "dz": torch.empty(B, C, device=dev), "da": torch.empty(B, H, device=dev), "dh": torch.empty(B, H, device=dev),
                 "dx": torch.empty(B, self.enc_dim, device=dev), "logits": torch.empty(B, C, device=dev),
                 "keepz": torch.ones(B, C, dtype=torch.uint8, device=dev),
                 "loss": torch.zeros(1, dtype=torch.float64, device=dev),
                 "correct": torch.zeros(1, dtype=torch.int32, device=dev)}
            self._work[B] = w
        return w

    def _forward_backward(self, feats, labels, keep_mask, seed, want_dx, want_param_grads, training=True):
        _require_cuda(feats)
        feats = feats.detach().contiguous().float()
        labels = labels.to(device=feats.device, dtype=torch.long).contiguous()
        B, E, H, C = feats.shape[0], self.enc_dim, self.hidden, self.nclasses
        w, P = self._buffers_for(B, feats.device), self._views
        w["loss"].zero_()
        w["correct"].zero_()
        ops.linear_fwd(feats, P["classifier.0.weight"], P["classifier.0.bias"], w["h"], B, H, E)
        p = DROPOUT_P if training else 0.0
        if keep_mask is not None:
            w["keep"].copy_(keep_mask.to(torch.uint8))
        elif not training:
            w["keep"].fill_(1)
        ops.dropout_relu_fwd(w["h"], w["keep"], w["a"], p, generate=training and keep_mask is None, seed=seed)
        ops.linear_fwd(w["a"], P["classifier.3.weight"], P["classifier.3.bias"], w["z"], B, C, H)
        need_bwd = want_dx or want_param_grads
        ops.relu_ce_fwd_bwd(w["z"], labels, B, C, 1.0, w["loss"], w["correct"], w["dz"] if need_bwd else None)
        if need_bwd:
            G = self._gviews
            if want_param_grads:
                self.grad.zero_()
            ops.linear_bwd(w["a"], P["classifier.3.weight"], w["dz"], w["da"], G["classifier.3.weight"] if want_param_grads else None,
                           G["classifier.3.bias"] if want_param_grads else None, B, C, H)
            ops.dropout_relu_bwd(w["da"], w["h"], w["keep"], w["dh"], p)
            ops.linear_bwd(feats, P["classifier.0.weight"], w["dh"], w["dx"] if want_dx else None,
                           G["classifier.0.weight"] if want_param_grads else None,
                           G["classifier.0.bias"] if want_param_grads else None, B, H, E)
        return w

    def head_loss_and_feat_grad(self, feats, labels, dfeat, keep_mask=None, seed=0):
        """Step 1 of the reference (main_train.py:377-403): returns (CE loss, #correct) as device tensors and ADDS the
        gradient-reversed -lambda * dCE/dfeats into `dfeat` (the encoder's feature gradient, (B, enc_dim) fp32)."""
        w = self._forward_backward(feats, labels, keep_mask, seed, want_dx=True, want_param_grads=False)
        ops.sgd_step(dfeat, w["dx"], dfeat.numel(), self.lambda_)         # dfeat -= lambda * dx  (GRL, model.py:990-994)
        # copies: the work buffers are overwritten by the classifier pass of the same train step
        return w["loss"].clone(), w["correct"].clone()

    def classifier_step(self, feats, labels, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0005, keep_mask=None, seed=0,
                        group=None):
        """Step 2 (main_train.py:420-453): CE on DETACHED features, one Adam(L2) step of the classifier alone.  Under
        data parallelism (`group`) the 40 k gradient values are sum-all-reduced and averaged first."""
        w = self._forward_backward(feats, labels, keep_mask, seed, want_dx=False, want_param_grads=True)
        scale = 1.0
        if group is not None:
            import torch.distributed as dist
            if dist.get_world_size(group) > 1:
                dist.all_reduce(self.grad, group=group)
                scale = 1.0 / dist.get_world_size(group)
        self.step_count += 1
        ops.adam_l2_step(self.flat, self.grad, self.m, self.v, self.flat.numel(), lr, beta1, beta2, eps, weight_decay,
                         self.step_count, scale)
        return w["loss"].clone(), w["correct"].clone()

    def forward(self, x):
        """Eval-style logits (validation, main_train.py:560-566): relu(W2 relu(W1 x + b1) + b2), no dropout, no grad."""
        _require_cuda(x)
        B = x.shape[0]
        w = self._buffers_for(B, x.device)
        P = self._views
        xf = x.detach().contiguous().float()
        ops.linear_fwd(xf, P["classifier.0.weight"], P["classifier.0.bias"], w["h"], B, self.hidden, self.enc_dim)
        w["keep"].fill_(1)
        ops.dropout_relu_fwd(w["h"], w["keep"], w["a"], 0.0, generate=False)
        ops.linear_fwd(w["a"], P["classifier.3.weight"], P["classifier.3.bias"], w["z"], B, self.nclasses, self.hidden)
        ops.dropout_relu_fwd(w["z"], w["keepz"], w["logits"], 0.0, generate=False)      # the trailing ReLU, model.py:1012
        return w["logits"]
