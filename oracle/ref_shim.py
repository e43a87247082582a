"""Import the UNMODIFIED reference modules from /root/reference under a compatibility shim.

TEST INFRASTRUCTURE ONLY.  This file is used in the authoring container to (a) pin the
oracle restatements in this directory against the reference's own code and (b) generate the
golden vectors under tests/golden/ (see oracle/make_golden.py).  /root/reference does not
exist on the GPU box, so nothing at run time (tests -m gpu, smoke, bench) imports this.

The reference targets torch 1.1 / python 3.6 (README.md:16-19) and does not import on
torch 2.11:
  * feature_extraction.py:6 imports librosa (absent)            -> stub module
  * ecapa_tdnn.py:12 imports pytorch_model_summary (absent)     -> stub module
  * utils_dsp.py:162 calls torch.rfft(v, 1, onesided=False)     -> torch.fft.fft + view_as_real
  * feature_extraction.py:109 calls torch.stft without return_complex -> wrapper
No reference source is copied; the modules are imported from where they lie.
"""
import importlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("AIR_REFERENCE_ROOT", "/root/reference")
# byte-for-byte copies of the hot-path modules made by oracle/build_ref.sh (git-ignored; they travel to the GPU box so
# that the reference arm of bench.py can time the reference itself there)
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "feature_extraction.py"))


def copy_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_COPY, f)) for f in
               ("feature_extraction.py", "utils_dsp.py", "resnet.py", "ecapa_tdnn.py", "loss.py"))


def use_copy_if_needed() -> bool:
    """Point the shim at oracle/_ref when the full tree is absent (GPU box).  True when some reference is importable."""
    global REFERENCE_ROOT
    if reference_available():
        return True
    if copy_available():
        REFERENCE_ROOT = REF_COPY
        return True
    return False


_installed = False


def _install():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in ("librosa", "pytorch_model_summary"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.summary = lambda *a, **k: ""
            sys.modules[name] = m
    if not hasattr(torch, "rfft"):
        def rfft(x, signal_ndim, onesided=True):
            assert signal_ndim == 1 and not onesided
            return torch.view_as_real(torch.fft.fft(x, dim=-1))
        torch.rfft = rfft
    if not getattr(torch.stft, "_air_shim", False):
        _stft = torch.stft

        def stft(x, n_fft, hop_length=None, win_length=None, window=None, center=True,
                 pad_mode="reflect", normalized=False, onesided=None, return_complex=None):
            out = _stft(x, n_fft, hop_length, win_length, window=window, center=center,
                        pad_mode=pad_mode, normalized=normalized, onesided=onesided,
                        return_complex=True)
            return out if return_complex else torch.view_as_real(out)
        stft._air_shim = True
        torch.stft = stft
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load(name: str):
    """Import reference module `name` (e.g. 'feature_extraction', 'resnet', 'ecapa_tdnn', 'loss')."""
    _install()
    return importlib.import_module(name)
