"""Thin Python wrappers over the C ABI (ctypes) for the tensor-core conv kernels and the
pointwise / reduction kernels.  All tensors are CUDA tensors owned by the caller."""
import ctypes

import torch

from . import _lib

_SMS = {}


def num_sms(device=None):
    d = torch.cuda.current_device() if device is None else device
    if d not in _SMS:
        _SMS[d] = torch.cuda.get_device_properties(d).multi_processor_count
    return _SMS[d]


def conv_block_n(n):
    return _lib.lib().air_conv_block_n(int(n))


def packed_elems(n, k):
    f = _lib.lib().air_conv_packed_elems
    f.restype = ctypes.c_longlong
    return int(f(int(n), int(k)))


def pack_weights(w, mode, cin, cout, taps, out=None):
    """w: fp32 [Cout][taps][Cin] (GEMM layout). mode 0 = fprop operand, 1 = dgrad operand."""
    assert w.dtype == torch.float32 and w.is_contiguous() and w.numel() == cout * taps * cin
    n, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
    if out is None:
        out = torch.empty(packed_elems(n, k), device=w.device, dtype=torch.bfloat16)
    st = _lib.lib().air_conv_pack_weights(_lib.ptr(w), _lib.ptr(out), n, k, mode, cin, cout, taps, _lib.stream_ptr())
    _lib.check(st, "air_conv_pack_weights")
    return out


def conv_gemm(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K,
              out, out_ld, bias=None, res=None, res_ld=0, relu=False, flags=0):
    st = _lib.lib().air_conv_gemm_bf16(
        _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode,
        _lib.ptr(wpk), N, K, _lib.ptr(out), _lib.LL(out_ld), _lib.ptr(bias), _lib.ptr(res), _lib.LL(res_ld),
        int(relu), num_sms(), flags, _lib.stream_ptr())
    _lib.check(st, "air_conv_gemm_bf16")
    return out


def conv_wgrad(x, x_ld, B, H, W, C, dy, dy_ld, Ho, Wo, N, kh, kw, sh, sw, ph, pw, dh, dw, dw_out, flags=0):
    """dw_out: fp32 [N][kh*kw*C] accumulated in place (caller zeroes)."""
    st = _lib.lib().air_conv_wgrad_bf16(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), Ho, Wo, N,
        kh, kw, sh, sw, ph, pw, dh, dw, _lib.ptr(dw_out), num_sms(), flags, _lib.stream_ptr())
    _lib.check(st, "air_conv_wgrad_bf16")
    return dw_out


def conv_out_size(n, k, s, p, d):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1
