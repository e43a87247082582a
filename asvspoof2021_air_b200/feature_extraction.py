"""Drop-in for the reference's feature_extraction.LFCC (feature_extraction.py:61-138), backed by
the fused sm_100a kernel (csrc/lfcc.cu) through the C ABI `air_lfcc_fwd`.

Same constructor signature, same registered state (`lfcc_fb` (257,20), `l_dct.weight` (20,20))
and the same forward contract `(B, L) float32 -> (B, 1 + L//fs, 3*filter_num)`.  Additionally
`extract()` produces the features already cropped/padded (dataset.py:66-79,513-528) and laid out
for the first conv of ResNet / ECAPA (main_train.py:338,347-348) so raw waves go end to end on
device.  There is no CPU implementation: inputs must live on a CUDA device.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib, lfcc_tables

PAD_MODES = {"none": 0, "zero": 1, "repeat": 2, "silence": 3}


class _FrozenDCT(nn.Module):
    """Holds `weight` under the reference's key `l_dct.weight` (utils_dsp.LinearDCT, :220-244)."""

    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(lfcc_tables.dct_ortho_matrix(n), requires_grad=False)


class LFCC(nn.Module):
    def __init__(self, fl, fs, fn, sr, filter_num, with_energy=False, with_emphasis=True, with_delta=True):
        super().__init__()
        if (fl, fs, fn, filter_num) != (320, 160, 512, 20):
            raise NotImplementedError(
                "the fused sm_100a LFCC kernel is specialised for LFCC(320, 160, 512, sr, 20), the only "
                "configuration the reference instantiates (dataset.py:13, preprocess.py:237)")
        if with_energy:
            raise NotImplementedError("with_energy=True is unused by the reference and not implemented")
        self.fl, self.fs, self.fn, self.sr, self.filter_num = fl, fs, fn, sr, filter_num
        self.with_energy, self.with_emphasis, self.with_delta = with_energy, with_emphasis, with_delta
        self.lfcc_fb = nn.Parameter(lfcc_tables.linear_filterbank(fn, sr, filter_num), requires_grad=False)
        self.l_dct = _FrozenDCT(filter_num)
        self._table = None
        self._silence = None
        self._tc = None
        # "tc": tensor-core DFT (csrc/lfcc_tc.cu, 3-term fp16 / bf16 split); "fft": radix FFT in fp32 on the CUDA cores
        # (csrc/lfcc.cu); "auto" (default) = "tc", see impl_for().  AIR_LFCC_IMPL / `.impl` force one.
        self.impl = os.environ.get("AIR_LFCC_IMPL", "auto")

    # -- constants -------------------------------------------------------------------------
    def _consts(self, device):
        if self._table is None or self._table.device != device:
            self._table = lfcc_tables.pack_table(self.lfcc_fb, self.l_dct.weight).to(device)
            self._silence = None
        return self._table

    def _tc_consts(self, device):
        """Tables of the tensor-core path, or None when the filterbank buffer does not have the canonical band structure."""
        if self._tc is None or self._tc[0].device != device:
            if not lfcc_tables.tc_filter_structure_ok(self.lfcc_fb):
                return None
            self._tc = (lfcc_tables.pack_tc_table(self.lfcc_fb, self.l_dct.weight).to(device),
                        lfcc_tables.pack_tc_dft().to(device))
        return self._tc

    def impl_for(self, dtype):
        """Which kernel serves an output dtype: the tensor-core kernel for every dtype.  Its 3-term split keeps the hi parts
        in fp16 and the residuals in bf16 (csrc/lfcc_tc.cu), which holds the fp32 contract on any input: worst deviation from
        the float64 oracle 1.4e-6 (|ref|+1) on the golden waves, 1.7e-5 on frames whose weak bands lie 60 dB under the
        strongest, 1.4e-6 on the same frames 60 dB quieter (tolerance 1e-4; the reference's own fp32 arithmetic: 9e-6).
        "fft" (AIR_LFCC_IMPL / `.impl`): the radix FFT in fp32 on the CUDA cores (csrc/lfcc.cu, 2.4e-6, 147 us instead of
        112 us per 256 utterances); the fp32 parity mode of the Trainer pins it."""
        if self.impl in ("tc", "fft"):
            return self.impl
        return "tc"

    def num_frames(self, length):
        return 1 + length // self.fs

    def silence_vector(self, device):
        """LFCC frame 0 of 3200 zero samples (dataset.py:13-16), computed by the same kernel."""
        if self._silence is None or self._silence.device != device:
            z = torch.zeros(1, 3200, device=device)
            self._silence = self._run(z, None, 0, 0, None, "btd", torch.float32)[0, 0].contiguous()
        return self._silence

    # -- kernel launch ---------------------------------------------------------------------
    def _run(self, x, lengths, feat_len, pad_mode, start, layout, dtype, out=None, fseg=0):
        if not x.is_cuda:
            raise _lib.AirError("LFCC runs on CUDA only (no CPU path); got a %s tensor" % x.device)
        if x.dim() != 2 or x.dtype != torch.float32:
            raise ValueError("LFCC expects a (batch, length) float32 tensor")
        if x.stride(1) != 1:
            x = x.contiguous()
        B, L = x.shape
        dev = x.device
        Tout = feat_len if feat_len > 0 else self.num_frames(L)
        D = 3 * self.filter_num
        if layout == "btd":            # (B, T, 60)  -- the reference LFCC.forward layout
            shape, sb, sj, sd = (B, Tout, D), Tout * D, D, 1
        elif layout == "resnet":       # (B, 1, 60, T): H = coefficient, W = time (NHWC with C = 1)
            shape, sb, sj, sd = (B, 1, D, Tout), D * Tout, 1, Tout
        elif layout == "ecapa":        # (B, T, 64) channels-last, channels 60..63 stay zero
            shape, sb, sj, sd = (B, Tout, 64), Tout * 64, 64, 1
        else:
            raise ValueError(layout)
        if out is None:
            need_zero = layout == "ecapa" or lengths is not None or (feat_len > 0 and pad_mode == 0)
            out = (torch.zeros if need_zero else torch.empty)(shape, device=dev, dtype=dtype)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
            raise ValueError("bad `out` tensor")
        if lengths is not None:
            lengths = lengths.to(device=dev, dtype=torch.int32).contiguous()
        if start is not None:
            start = start.to(device=dev, dtype=torch.int32).contiguous()
        sil = self.silence_vector(dev) if (feat_len > 0 and pad_mode == 3) else None
        tc = self._tc_consts(dev) if self.impl_for(dtype) == "tc" else None
        if tc is not None:
            st = _lib.lib().air_lfcc_tc_fwd(
                _lib.ptr(x), _lib.LL(x.stride(0)), _lib.ptr(lengths), B, L, _lib.ptr(tc[0]), _lib.ptr(tc[1]),
                _lib.ptr(out), _lib.LL(sb), _lib.LL(sj), _lib.LL(sd), int(dtype == torch.bfloat16),
                Tout, feat_len, pad_mode, _lib.ptr(start), _lib.ptr(sil),
                ctypes.c_float(0.97 if self.with_emphasis else 0.0),
                torch.cuda.get_device_properties(dev).multi_processor_count, _lib.stream_ptr())
            _lib.check(st, "air_lfcc_tc_fwd")
            return out
        table = self._consts(dev)
        st = _lib.lib().air_lfcc_fwd(
            _lib.ptr(x), _lib.LL(x.stride(0)), _lib.ptr(lengths), B, L, _lib.ptr(table),
            _lib.ptr(out), _lib.LL(sb), _lib.LL(sj), _lib.LL(sd), int(dtype == torch.bfloat16),
            Tout, feat_len, pad_mode, _lib.ptr(start), _lib.ptr(sil),
            ctypes.c_float(0.97 if self.with_emphasis else 0.0), fseg, _lib.stream_ptr())
        _lib.check(st, "air_lfcc_fwd")
        return out

    # -- public API ------------------------------------------------------------------------
    def forward(self, x):
        """x (batch, length) -> (batch, frame_num, dim_num).  Unlike the reference
        (feature_extraction.py:106) the caller's tensor is NOT modified in place."""
        y = self._run(x, None, 0, 0, None, "btd", torch.float32)
        return y if self.with_delta else y[:, :, :self.filter_num].contiguous()

    def extract(self, x, lengths=None, feat_len=750, padding="repeat", start=None,
                layout="resnet", dtype=torch.bfloat16, out=None, fseg=0):
        """Fused wave -> LFCC -> crop/pad -> model layout.

        start: int tensor (B,) of crop offsets for utterances longer than feat_len (the
        reference draws np.random.randint(T - feat_len), dataset.py:68); None crops at 0."""
        return self._run(x, lengths, feat_len, PAD_MODES[padding], start, layout, dtype, out, fseg)
