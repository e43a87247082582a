"""Drop-in for the reference's ECAPA-TDNN classes (ecapa_tdnn.py:15-198): `Res2Net2`, `Bottle2neck`,
`SEModule`.  main_train.py:167-168 builds `Res2Net2(Bottle2neck, C=512, model_scale=8, nOut=2, n_mels=60)`.

Same constructor signature, the same 248 state_dict keys / shapes in the same order, the same forward
contract `(B, n_mels, T) -> (feat (B,256), logits (B,nOut))`, train()/eval() BatchNorm semantics and an
autograd backward -- executed by the hand-written sm_100a kernels of engine_ecapa.py.  CUDA only.
`Bottle2neck` / `SEModule` are kept as constructor-compatible markers: the block structure is fixed by
`Res2Net2` (the only way the reference uses them) and run fused inside the engine.
"""
import torch
import torch.nn as nn

from . import _lib
from .engine_ecapa import EcapaEngine
from .module_utils import bind_state


class SEModule(nn.Module):
    def __init__(self, channels, bottleneck=128):
        super().__init__()
        self.channels, self.bottleneck = channels, bottleneck

    def forward(self, input):
        raise _lib.AirError("SEModule runs fused inside Res2Net2 (engine_ecapa.py); it has no standalone kernel path")


class Bottle2neck(nn.Module):
    def __init__(self, inplanes, planes, kernel_size=None, dilation=None, scale=4):
        super().__init__()
        self.inplanes, self.planes, self.kernel_size, self.dilation, self.scale = inplanes, planes, kernel_size, dilation, scale

    def forward(self, x):
        raise _lib.AirError("Bottle2neck runs fused inside Res2Net2 (engine_ecapa.py); it has no standalone kernel path")


class _EcapaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        feat, logits = module.engine.forward(x, training=module.training)
        ctx.module = module
        return feat.clone(), logits.clone()

    @staticmethod
    def backward(ctx, dfeat, dlogits):
        module = ctx.module
        eng = module.engine
        eng.zero_grad()
        has_head = dlogits is not None and bool((dlogits != 0).any())
        if dfeat is None:
            dfeat = torch.zeros_like(eng.feat)
        eng.backward(dfeat.float(), dlogits.float().contiguous() if has_head else None)
        grads = []
        for name in module._param_names:
            if name.startswith(("fc7.", "bn7.")) and not has_head:
                grads.append(None)          # no gradient under OC-Softmax
            else:
                grads.append(eng.store.pt_view(name, eng.store.grads).clone())
        return (None, None, *grads)


def _rebuild_ecapa(args, sd, training):
    C, scale, n_out, n_mels = args
    m = Res2Net2(Bottle2neck, C=C, model_scale=scale, nOut=n_out, n_mels=n_mels)
    m.load_state_dict(sd)
    m.train(training)
    return m


class Res2Net2(nn.Module):
    def __init__(self, block, C, model_scale, nOut, n_mels, encoder_type='ECA', context=True, summed=False, out_bn=True,
                 device=None, engine=None, **kwargs):
        super().__init__()
        if encoder_type != 'ECA' or not context or summed or not out_bn:
            raise NotImplementedError("only encoder_type='ECA', context=True, summed=False, out_bn=True "
                                      "(the main_train.py:167-168 configuration) is implemented")
        if block is not Bottle2neck and getattr(block, "__name__", "") != "Bottle2neck":
            raise NotImplementedError("block must be Bottle2neck")
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.context, self.summed, self.n_mfcc, self.encoder_type, self.out_bn = context, summed, n_mels, encoder_type, out_bn
        self.scale, self.C, self.nOut = model_scale, C, nOut
        self.engine = engine if engine is not None else EcapaEngine(C=C, scale=model_scale, n_out=nOut, n_mels=n_mels,
                                                                    device=device, train_head=True)
        self._bind()

    def __reduce__(self):
        """Whole-module pickles (main_train.py:674-706): constructor arguments + CPU state_dict."""
        sd = {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}
        return (_rebuild_ecapa, ((self.C, self.scale, self.nOut, self.n_mfcc), sd, self.training))

    def _ordered_keys(self):
        def conv(p):
            return [(p + ".weight", "param"), (p + ".bias", "param")]

        def bn(p):
            return [(p + ".weight", "param"), (p + ".bias", "param"), (p + ".running_mean", "buffer"),
                    (p + ".running_var", "buffer"), (p + ".num_batches_tracked", "buffer")]
        keys = conv("conv1") + bn("bn1")
        for li in (1, 2, 3):
            p = "layer%d" % li
            keys += conv(p + ".conv1") + bn(p + ".bn1")
            for i in range(self.scale - 1):
                keys += conv(p + ".convs.%d" % i)
            for i in range(self.scale - 1):
                keys += bn(p + ".bns.%d" % i)
            keys += conv(p + ".conv3") + bn(p + ".bn3")
            keys += conv(p + ".se.se.1") + bn(p + ".se.se.3") + conv(p + ".se.se.4")
        keys += conv("layer4") + conv("attention.0") + bn("attention.2") + conv("attention.3")
        keys += bn("bn5") + conv("fc6") + conv("fc7") + bn("bn7")
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

    def load_state_dict(self, state_dict, strict=True, assign=False):
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.engine.mark_dirty()
        return out

    def _apply(self, fn, recurse=True):
        probe = fn(torch.zeros(1, device=self.engine.device))
        if probe.dtype != torch.float32:
            raise NotImplementedError("parameters are kept in fp32 master copies; bf16 is used inside the kernels")
        if probe.device != self.engine.device:
            sd = {k: v.detach().clone() for k, v in self.state_dict().items()}
            self.engine = EcapaEngine(C=self.C, scale=self.scale, n_out=self.nOut, n_mels=self.n_mfcc, device=probe.device,
                                      train_head=True)
            self._bind()
            super().load_state_dict({k: v.to(probe.device) for k, v in sd.items()})
            self.engine.mark_dirty()
        return self

    def forward(self, x):
        """x: (B, n_mels, T) float features (main_train.py:347-348 squeezes the channel axis)."""
        if not x.is_cuda:
            raise _lib.AirError("Res2Net2 runs on CUDA only (no CPU path); got a %s tensor" % x.device)
        if x.dim() != 3 or x.shape[1] != self.n_mfcc:
            raise ValueError("expected (B, %d, T) features" % self.n_mfcc)
        eng = self.engine
        xb = torch.zeros(x.shape[0], x.shape[2], eng.mels_g, device=x.device, dtype=torch.bfloat16)
        xb[:, :, :self.n_mfcc] = x.transpose(1, 2)
        eng.mark_dirty()                   # parameters may have been updated by an external optimiser
        params = [self.get_parameter(n) for n in self._param_names]
        return _EcapaFn.apply(self, xb, *params)
