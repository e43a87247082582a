"""Checkpoint compatibility with the reference.

The reference saves WHOLE MODULES with `torch.save(feat_model, ...)` (main_train.py:674-706) and loads them with
`torch.load(path)` (main_train.py:172-173, generate_score.py:46-48).  Such a pickle names its classes by module path:
`model.ResNet` / `resnet.ResNet` (+ `PreActBlock`, `SelfAttention`), `ECAPA_TDNN.Res2Net2` / `ecapa_tdnn.Res2Net2`
(+ `Bottle2neck`, `SEModule`), `loss.AngularIsoLoss` / `loss.OCSoftmax` -- modules that do not exist here.

`load_module(path)` unpickles such a file without the reference's source tree: for the duration of the load the
reference's module names resolve to empty nn.Module shells (the tensors, buffers and hyper-parameters live in the
pickled instance state, not in the class), then the shell is ADOPTED: its parameters / buffers are walked by name and
loaded into the corresponding drop-in of this package (same state_dict keys, tests/test_*_gpu.py
::test_state_dict_keys_match_reference).  A pickle written by this package loads unchanged.

The other direction needs nothing special: this package's pickles name `asvspoof2021_air_b200.resnet.ResNet` etc.
and rebuild themselves from (constructor arguments, state_dict) wherever the package is importable."""
import contextlib
import sys
import types

import torch
import torch.nn as nn

# reference module name -> class names its pickles may mention
REFERENCE_CLASSES = {
    "model": ["ResNet", "PreActBlock", "PreActBottleneck", "SelfAttention"],
    "resnet": ["ResNet", "PreActBlock", "PreActBottleneck", "SelfAttention"],
    "ecapa_tdnn": ["Res2Net2", "Bottle2neck", "SEModule"],
    "ECAPA_TDNN": ["Res2Net2", "Bottle2neck", "SEModule"],
    "loss": ["OCSoftmax", "AngularIsoLoss"],
}


class ReferenceShell(nn.Module):
    """Unpickle target for a reference class: holds whatever instance state the pickle carries; never called."""

    def forward(self, *a, **k):
        raise RuntimeError("%s is an unpickled reference module shell; pass it through compat.adopt()" % type(self).__name__)


@contextlib.contextmanager
def reference_modules():
    """Make the reference's module names importable (as shells) while a pickle is being read."""
    installed = []
    for mod_name, classes in REFERENCE_CLASSES.items():
        if mod_name in sys.modules:
            continue
        m = types.ModuleType(mod_name)
        m.__doc__ = "pickle shim installed by asvspoof2021_air_b200.compat"
        for c in classes:
            setattr(m, c, type(c, (ReferenceShell,), {"__module__": mod_name}))
        sys.modules[mod_name] = m
        installed.append(mod_name)
    try:
        yield
    finally:
        for mod_name in installed:
            sys.modules.pop(mod_name, None)


def named_state(module, prefix=""):
    """(key, tensor) pairs like nn.Module.state_dict(), read straight from _parameters / _buffers / _modules so that
    pickles of any torch vintage work (no hooks, no attributes newer than the pickle)."""
    out = {}
    d = module.__dict__
    for name, p in (d.get("_parameters") or {}).items():
        if p is not None:
            out[prefix + name] = p.detach()
    skip = d.get("_non_persistent_buffers_set") or ()
    for name, b in (d.get("_buffers") or {}).items():
        if b is not None and name not in skip:
            out[prefix + name] = b.detach()
    for name, child in (d.get("_modules") or {}).items():
        if child is not None:
            out.update(named_state(child, prefix + name + "."))
    return out


def adopt(obj, device=None):
    """Turn an unpickled reference module (shell) into this package's drop-in; pass drop-ins through."""
    from . import ecapa_tdnn, loss, resnet
    if isinstance(obj, (resnet.ResNet, ecapa_tdnn.Res2Net2, loss.OCSoftmax)):
        return obj
    if not isinstance(obj, nn.Module):
        raise TypeError("not a module checkpoint: %r" % type(obj))
    kind = type(obj).__name__
    sd = {k: v.cpu() for k, v in named_state(obj).items()}
    training = bool(obj.__dict__.get("training", False))
    if kind == "ResNet":
        if "layer1.0.conv1.weight" not in sd or "layer1.0.conv3.weight" in sd:
            raise NotImplementedError("only ResNet-18 checkpoints (PreActBlock) are supported")
        enc_dim, nclasses = sd["fc.weight"].shape[0], sd["fc_mu.weight"].shape[0]
        num_nodes = sd["conv5.weight"].shape[2]                        # conv5 kernel height = num_nodes (resnet.py:140)
        m = resnet.ResNet(num_nodes, enc_dim, '18', nclasses=nclasses, device=device)
    elif kind == "Res2Net2":
        C, n_mels = sd["conv1.weight"].shape[0], sd["conv1.weight"].shape[1]
        scale = 1 + sum(1 for k in sd if k.startswith("layer1.convs.") and k.endswith(".weight"))
        m = ecapa_tdnn.Res2Net2(ecapa_tdnn.Bottle2neck, C=C, model_scale=scale, nOut=sd["fc7.weight"].shape[0],
                                n_mels=n_mels, device=device)
    elif kind in ("OCSoftmax", "AngularIsoLoss"):
        d = obj.__dict__
        cls = loss.AngularIsoLoss if kind == "AngularIsoLoss" else loss.OCSoftmax
        m = cls(sd["center"].shape[1], r_real=d.get("r_real", 0.9), r_fake=d.get("r_fake", 0.5), alpha=d.get("alpha", 20.0))
        m.center.data.copy_(sd["center"])
        return m
    else:
        raise NotImplementedError("checkpoints of %s are outside the B200 path (ResNet-18 / Res2Net2 / OC-Softmax only)" % kind)
    missing = [k for k in m.state_dict() if k not in sd]
    extra = [k for k in sd if k not in m.state_dict()]
    if missing or extra:
        raise KeyError("checkpoint does not match the %s drop-in: missing %s, unexpected %s" % (kind, missing[:5], extra[:5]))
    m.load_state_dict(sd)
    m.train(training)
    return m


def load_module(path, device=None):
    """torch.load of a whole-module checkpoint written by the reference OR by this package -> a drop-in module."""
    with reference_modules():
        obj = torch.load(path, map_location="cpu", weights_only=False)
    return adopt(obj, device=device)
