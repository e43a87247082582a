"""Drop-in for the reference's `ResNet` (resnet.py:122-191 == model.py:177-253; the copy that
main_train.py:8,162-163 instantiates as ResNet(3, enc_dim, resnet_type='18', nclasses=2)).

Same constructor signature, same 117 state_dict keys / shapes, same forward contract
`(B,1,60,T) -> (feat (B,enc_dim), mu (B,nclasses))`, train()/eval() BatchNorm semantics and an
autograd backward -- but every layer runs in the hand-written sm_100a kernels of engine.py
(tcgen05 implicit-GEMM convs, fused BN/ReLU, attentive pooling).  CUDA only; no CPU path.
"""
import torch
import torch.nn as nn

from . import _lib
from .engine import ResNetEngine
from .module_utils import bind_state


class _ResNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.engine
        feat, mu = eng.forward(x, training=module.training)
        ctx.module = module
        return feat.clone(), mu.clone()

    @staticmethod
    def backward(ctx, dfeat, dmu):
        module = ctx.module
        eng = module.engine
        eng.zero_grad()
        has_mu = dmu is not None and bool((dmu != 0).any())
        if dfeat is None:
            dfeat = torch.zeros_like(eng.feat)
        eng.backward(dfeat.float(), dmu.float().contiguous() if has_mu else None)
        grads = []
        for name in module._param_names:
            if name.startswith("fc_mu.") and not has_mu:
                grads.append(None)          # the reference leaves these without grad under OC-Softmax
            else:
                grads.append(eng.store.pt_view(name, eng.store.grads).clone())
        return (None, None, *grads)


def _rebuild_resnet(args, sd, training):
    m = ResNet(*args)
    m.load_state_dict(sd)
    m.train(training)
    return m


class ResNet(nn.Module):
    def __init__(self, num_nodes, enc_dim, resnet_type='18', nclasses=2, device=None, engine=None):
        super().__init__()
        if str(resnet_type) != '18':
            raise NotImplementedError("only the ResNet-18 configuration used by main_train.py:162-163 is implemented")
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.num_nodes, self.enc_dim, self.nclasses = num_nodes, enc_dim, nclasses
        self.engine = engine if engine is not None else ResNetEngine(
            enc_dim=enc_dim, nclasses=nclasses if nclasses >= 2 else 1, device=device, train_head_mu=True,
            num_nodes=num_nodes)
        self._bind()

    def __reduce__(self):
        """Whole-module pickles (torch.save(feat_model, ...), main_train.py:674-706) carry the constructor
        arguments and a CPU state_dict; unpickling rebuilds the engine on the current default device."""
        sd = {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}
        return (_rebuild_resnet, ((self.num_nodes, self.enc_dim, '18', self.nclasses), sd, self.training))

    # ---- state binding ------------------------------------------------------------------
    def _ordered_keys(self):
        keys = [("conv1.weight", "param")]

        def bn(p):
            return [(p + ".weight", "param"), (p + ".bias", "param"), (p + ".running_mean", "buffer"),
                    (p + ".running_var", "buffer"), (p + ".num_batches_tracked", "buffer")]
        keys += bn("bn1")
        for blk in self.engine.blocks:
            p = blk.name
            keys += bn(p + ".bn1") + [(p + ".conv1.weight", "param")] + bn(p + ".bn2") + [(p + ".conv2.weight", "param")]
            if blk.sc is not None:
                keys.append((p + ".shortcut.0.weight", "param"))
        keys += [("conv5.weight", "param")] + bn("bn5")
        keys += [("fc.weight", "param"), ("fc.bias", "param"), ("fc_mu.weight", "param"), ("fc_mu.bias", "param"),
                 ("attention.att_weights", "param")]
        return keys

    def _bind(self):
        eng = self.engine
        pviews = {n: eng.store.pt_view(n) for n in eng.store.names()}
        bviews = {}
        for name, bn in eng.bns():
            bviews[name + ".running_mean"] = bn.running_mean
            bviews[name + ".running_var"] = bn.running_var
            bviews[name + ".num_batches_tracked"] = bn.num_batches_tracked
        keys = self._ordered_keys()
        bind_state(self, keys, pviews, bviews)
        self._param_names = [k for k, kind in keys if kind == "param"]

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.engine.mark_dirty()
        return out

    def _apply(self, fn, recurse=True):
        probe = fn(torch.zeros(1, device=self.engine.device))
        if probe.device != self.engine.device or probe.dtype != torch.float32:
            if probe.dtype != torch.float32:
                raise NotImplementedError("parameters are kept in fp32 master copies; bf16 is used inside the kernels")
            sd = {k: v.detach().clone() for k, v in self.state_dict().items()}
            self.engine = ResNetEngine(enc_dim=self.enc_dim, nclasses=self.nclasses if self.nclasses >= 2 else 1,
                                       device=probe.device, train_head_mu=True, num_nodes=self.num_nodes)
            self._bind()
            super().load_state_dict({k: v.to(probe.device) for k, v in sd.items()})
            self.engine.mark_dirty()
        return self

    # ---- forward --------------------------------------------------------------------------
    def forward(self, x):
        if not x.is_cuda:
            raise _lib.AirError("ResNet runs on CUDA only (no CPU path); got a %s tensor" % x.device)
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError("expected (B, 1, 60, T) features")
        xb = x[:, 0].to(torch.bfloat16).contiguous()
        if getattr(self, "_params_touched", True):
            self.engine.mark_dirty()       # parameters may have been updated by an external optimiser
        params = [self.get_parameter(n) for n in self._param_names]
        return _ResNetFn.apply(self, xb, *params)
