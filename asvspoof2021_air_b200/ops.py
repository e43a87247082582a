"""Thin Python wrappers over the C ABI (ctypes) for the tensor-core conv kernels and the
pointwise / reduction kernels.  All tensors are CUDA tensors owned by the caller."""
import ctypes

import torch

from . import _lib

_SMS = {}


_RESERVED = [0]


def reserve_sms(k):
    """Leave k SMs out of the persistent grids launched from now on (0 = use them all).  The conv kernels are one CTA per SM
    with the whole register file: a collective kernel resident on a few SMs would push that many of their CTAs into a second
    wave (a ~2x longer launch); sized to the collective's CTA cap they run side by side (parallel.GradReducer)."""
    _RESERVED[0] = max(0, int(k))


def num_sms(device=None):
    d = torch.cuda.current_device() if device is None else device
    if d not in _SMS:
        _SMS[d] = torch.cuda.get_device_properties(d).multi_processor_count
    return max(1, _SMS[d] - _RESERVED[0])


F32 = torch.float32
F32_OUT = 1                     # AIR_CONV_F32_OUT (include/air_b200.h): out / res / out2 of a conv are float tensors


def _is_f32(t):
    return t is not None and t.dtype == F32


def _conv_flags(out, *others):
    """flags of a conv launch from the dtype of its output; residual / second output must share it (fp32 parity mode)."""
    f32 = _is_f32(out)
    for t in others:
        if t is not None and _is_f32(t) != f32:
            raise TypeError("conv output, residual and second output must share one dtype")
    return F32_OUT if f32 else 0


def split_terms(x, x_ld, M, C, out, out_ld, nterms, lo_mask):
    """fp32 rows -> bf16 (or bf16-exact fp32) split terms concatenated along the channels (csrc/split.cu)."""
    assert x.dtype == F32
    _lib.check(_lib.lib().air_split_terms(_lib.ptr(x), _lib.LL(x_ld), _lib.LL(M), int(C), _lib.ptr(out), _lib.LL(out_ld),
                                          int(out.dtype == F32), int(nterms), int(lo_mask), _lib.stream_ptr()), "air_split_terms")
    return out


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
        int(relu), num_sms(), flags | _conv_flags(out, res), _lib.stream_ptr())
    _lib.check(st, "air_conv_gemm_bf16")
    return out


def conv_gemm_stats(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K,
                    out, out_ld, bias, res, res_ld, relu, stats):
    """conv_gemm whose epilogue also accumulates sum / sum of squares of the stored output into stats (fp64 [2N])."""
    st = _lib.lib().air_conv_gemm_bf16_stats(
        _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode,
        _lib.ptr(wpk), N, K, _lib.ptr(out), _lib.LL(out_ld), _lib.ptr(bias), _lib.ptr(res), _lib.LL(res_ld),
        int(relu), _lib.ptr(stats), num_sms(), _conv_flags(out, res), _lib.stream_ptr())
    _lib.check(st, "air_conv_gemm_bf16_stats")
    return out


def conv_wgrad(x, x_ld, B, H, W, C, dy, dy_ld, Ho, Wo, N, kh, kw, sh, sw, ph, pw, dh, dw, dw_out, flags=0):
    """dw_out: fp32 [N][kh*kw*C] accumulated in place (caller zeroes)."""
    st = _lib.lib().air_conv_wgrad_bf16(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), Ho, Wo, N,
        kh, kw, sh, sw, ph, pw, dh, dw, _lib.ptr(dw_out), num_sms(), flags, _lib.stream_ptr())
    _lib.check(st, "air_conv_wgrad_bf16")
    return dw_out


def patch_supported(C, N, H, W):
    return bool(_lib.lib().air_conv3x3_patch_supported(int(C), int(N), int(H), int(W)))


def pack3x3(w, C, N, mode, out):
    """w: fp32 [Cout][3][3][Cin] (GEMM layout); mode 0: (C, N) = (Cin, Cout); mode 1 (dgrad): (C, N) = (Cout, Cin)."""
    _lib.check(_lib.lib().air_conv3x3_pack_weights(_lib.ptr(w), _lib.ptr(out), C, N, mode, _lib.stream_ptr()),
               "air_conv3x3_pack_weights")
    return out


def conv3x3_patch(a, a_ld, B, H, W, C, wpk, N, out, out_ld, res=None, res_ld=0, relu=False, mode=0):
    """mode is only a label for the profiler (0 fprop, 1 dgrad): the arithmetic is identical."""
    if _is_f32(out):
        return _patch_taps_ex2(a, a_ld, B, H, W, C, wpk, 9, N, out, out_ld, H, W, res, res_ld, relu, None, None, 0, None,
                               H, W, -1, -1, 1, 1, 0, 0, 9, _DR3, _DC3, _SL3)
    _lib.check(_lib.lib().air_conv3x3_patch_bf16(_lib.ptr(a), _lib.LL(a_ld), B, H, W, C, _lib.ptr(wpk), N, _lib.ptr(out),
                                                 _lib.LL(out_ld), _lib.ptr(res), _lib.LL(res_ld), int(relu), num_sms(),
                                                 _lib.stream_ptr()), "air_conv3x3_patch_bf16")
    return out


def pack_patch(w, C, N, taps, mode, out):
    """w: fp32 [Cout][taps][Cin]; mode 0: (C, N) = (Cin, Cout); mode 1: (C, N) = (Cout, Cin), taps flipped."""
    _lib.check(_lib.lib().air_conv_patch_pack_weights(_lib.ptr(w), _lib.ptr(out), C, N, taps, mode, _lib.stream_ptr()),
               "air_conv_patch_pack_weights")
    return out


class PackPlan:
    """Batched weight packing: collect the (source view, destination buffer, layout) jobs of a model once, then
    re-pack all of them with ONE launch per optimiser step (csrc/pack.cu)."""

    def __init__(self, device):
        self.device, self.records, self.keep, self.table, self.max_total = device, [], [], None, 0

    def _rec(self):
        return (ctypes.c_longlong * 16)()

    def add_gemm(self, w, w_ld, mode, cin, cout, taps, out):
        n, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
        r = self._rec()
        _lib.check(_lib.lib().air_pack_job_gemm(r, _lib.ptr(w), _lib.LL(w_ld), _lib.ptr(out), n, k, mode, cin, cout, taps),
                   "air_pack_job_gemm", 0)
        self._push(r, w, out)

    def add_patch(self, w, C, N, taps, mode, out):
        r = self._rec()
        _lib.check(_lib.lib().air_pack_job_patch(r, _lib.ptr(w), _lib.ptr(out), C, N, taps, mode), "air_pack_job_patch", 0)
        self._push(r, w, out)

    def _push(self, r, w, out):
        self.records.append(list(r))
        self.keep.append((w, out))          # the table holds raw pointers: keep the tensors alive
        self.max_total = max(self.max_total, int(r[3]))
        self.table = None

    def run(self):
        if not self.records:
            return
        if self.table is None:
            self.table = torch.tensor(self.records, dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib().air_pack_jobs(_lib.ptr(self.table), len(self.records), _lib.LL(self.max_total),
                                            _lib.stream_ptr()), "air_pack_jobs")


_ONE_TAP = (ctypes.c_int * 1)(0)
_DR3 = (ctypes.c_int * 9)(*[t // 3 for t in range(9)])
_DC3 = (ctypes.c_int * 9)(*[t % 3 for t in range(9)])
_SL3 = (ctypes.c_int * 9)(*range(9))


def _patch_taps_ex2(a, a_ld, B, Hin, Win, C, wpk, wtaps, N, out, out_ld, OH, OW, res, res_ld, relu, bias, out2, out2_ld, stats,
                    GH, GW, org_h, org_w, osh, osw, oph, opw, ntaps, dr, dc, sl):
    """The general patch-kernel entry point; the storage type of out / res / out2 (bf16 or float) travels in `flags`."""
    _lib.check(_lib.lib().air_conv_patch_taps_ex2_bf16(
        _lib.ptr(a), _lib.LL(a_ld), B, Hin, Win, C, _lib.ptr(wpk), wtaps, N, _lib.ptr(out), _lib.LL(out_ld), OH, OW,
        _lib.ptr(res), _lib.LL(res_ld), int(relu), _lib.ptr(bias), _lib.ptr(out2), _lib.LL(out2_ld), _lib.ptr(stats),
        GH, GW, org_h, org_w, osh, osw, oph, opw, ntaps, dr, dc, sl, _conv_flags(out, res, out2), num_sms(),
        _lib.stream_ptr()), "air_conv_patch_taps_ex2_bf16")
    return out


def conv1x1_patch(a, a_ld, B, H, W, C, wpk, N, out, out_ld, res=None, res_ld=0, relu=False, mode=0):
    """1x1 / stride-1 convolution (a plain GEMM over pixels) through the TMA patch kernel; mode labels the profile."""
    if _is_f32(out):
        return _patch_taps_ex2(a, a_ld, B, H, W, C, wpk, 1, N, out, out_ld, H, W, res, res_ld, relu, None, None, 0, None,
                               H, W, 0, 0, 1, 1, 0, 0, 1, _ONE_TAP, _ONE_TAP, _ONE_TAP)
    _lib.check(_lib.lib().air_conv_patch_taps_bf16(
        _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, _lib.ptr(wpk), 1, N, _lib.ptr(out), _lib.LL(out_ld), H, W,
        _lib.ptr(res), _lib.LL(res_ld), int(relu), H, W, 0, 0, 1, 1, 0, 0, 1, _ONE_TAP, _ONE_TAP, _ONE_TAP,
        num_sms(), _lib.stream_ptr()), "air_conv_patch_taps_bf16")
    return out


def conv1d_patch(a, a_ld, B, H, W, C, wpk, k, d, N, out, out_ld, bias=None, res=None, res_ld=0, relu=False,
                 out2=None, out2_ld=0, mode=0, post_scale=None, post_shift=None, stats=None):
    """1-D convolution over W (H independent rows), kernel k, dilation d, 'same' padding, through the TMA patch kernel.
    mode 0: forward (mode-0 packed weights); mode 1: data gradient (mode-1 packed weights).
    post_scale / post_shift (fp32 [N], optional): t = relu(acc + bias) * scale + shift; out2 <- t; out <- round(t) + res
    (eval-mode BatchNorm folded into the epilogue).  stats (fp64 [2N], optional): += per-channel sum / sum of squares of the
    stored output (the batch statistics of the BatchNorm that follows)."""
    zero = (ctypes.c_int * k)(*([0] * k))
    dc = (ctypes.c_int * k)(*[t * d for t in range(k)])
    sl = (ctypes.c_int * k)(*range(k))
    if H == 1:
        # no tap crosses rows, so the B sequences are the rows of ONE image: a work item (two rows x 128 columns) then
        # carries two sequences instead of one sequence and an empty row (same addresses: (B, 1, W, C) == (1, B, W, C))
        B, H = 1, B
    if post_scale is not None:
        _lib.check(_lib.lib().air_conv_patch_taps_ex3_bf16(
            _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, _lib.ptr(wpk), k, N, _lib.ptr(out), _lib.LL(out_ld), H, W,
            _lib.ptr(res), _lib.LL(res_ld), int(relu), _lib.ptr(bias), _lib.ptr(post_scale), _lib.ptr(post_shift),
            _lib.ptr(out2), _lib.LL(out2_ld), None, H, W, 0, -d * (k - 1) // 2, 1, 1, 0, 0, k, zero, dc, sl,
            _conv_flags(out, res, out2), num_sms(), _lib.stream_ptr()), "air_conv_patch_taps_ex3_bf16")
        return out
    return _patch_taps_ex2(a, a_ld, B, H, W, C, wpk, k, N, out, out_ld, H, W, res, res_ld, relu, bias, out2, out2_ld, stats,
                           H, W, 0, -d * (k - 1) // 2, 1, 1, 0, 0, k, zero, dc, sl)


def conv1d_wgrad_patch(x, x_ld, B, H, W, C, dy, dy_ld, N, k, d, dw_out, dw_ld=None):
    if H == 1:
        B, H = 1, B                          # as conv1d_patch: independent rows, two sequences per work item
    _lib.check(_lib.lib().air_conv1d_wgrad_patch_bf16(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), N, k, d, _lib.ptr(dw_out),
        _lib.LL(k * C if dw_ld is None else dw_ld), num_sms(), _lib.stream_ptr()), "air_conv1d_wgrad_patch_bf16")
    return dw_out


def conv_s2_dgrad_patch(dy, dy_ld, B, Ho, Wo, Cout, wpk, k, Cin, dx, dx_ld, H, W, res=None, res_ld=0):
    """Data gradient of a stride-2 k x k (k = 3 pad 1 / k = 1 pad 0) convolution by output parity classes."""
    _lib.check(_lib.lib().air_conv_s2_dgrad_patch_ex_bf16(
        _lib.ptr(dy), _lib.LL(dy_ld), B, Ho, Wo, Cout, _lib.ptr(wpk), k, Cin, _lib.ptr(dx), _lib.LL(dx_ld), H, W,
        _lib.ptr(res), _lib.LL(res_ld), _conv_flags(dx, res), num_sms(), _lib.stream_ptr()),
        "air_conv_s2_dgrad_patch_ex_bf16", 4 if k == 3 else 1)
    return dx


def wgrad_patch_supported(C, N):
    return bool(_lib.lib().air_conv3x3_wgrad_patch_supported(int(C), int(N)))


def conv3x3_wgrad_patch(x, x_ld, B, H, W, C, dy, dy_ld, N, dw_out, dw_ld=None):
    """dw_out: fp32 [N][9*C] (GEMM layout [Cout][tap][Cin]) accumulated in place (caller zeroes)."""
    _lib.check(_lib.lib().air_conv3x3_wgrad_patch_bf16(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), N, _lib.ptr(dw_out),
        _lib.LL(9 * C if dw_ld is None else dw_ld), num_sms(), _lib.stream_ptr()), "air_conv3x3_wgrad_patch_bf16")
    return dw_out


def conv_wgrad_patch(x, x_ld, B, H, W, C, dy, dy_ld, N, k, dw_out, dw_ld=None):
    """k = 3 (3x3 / s1 / p1) or 1 (1x1 / s1 / p0); dw_out fp32 [N][k*k*C] accumulated in place (caller zeroes)."""
    _lib.check(_lib.lib().air_conv_wgrad_patch_bf16(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), N, k, _lib.ptr(dw_out),
        _lib.LL(k * k * C if dw_ld is None else dw_ld), num_sms(), _lib.stream_ptr()), "air_conv_wgrad_patch_bf16")
    return dw_out


def conv_s2_wgrad_patch(x, x_ld, B, H, W, C, dy, dy_ld, Ho, Wo, N, k, dw_out, dw_ld=None):
    """Weight gradient of a stride-2 k x k (k = 3 pad 1 / k = 1 pad 0) layer: one stride-1 patch problem per input parity
    class (csrc/conv_wgrad_patch.cu); dw_out fp32 [N][k*k*C] accumulated in place."""
    _lib.check(_lib.lib().air_conv_s2_wgrad_patch_bf16(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), Ho, Wo, N, k, _lib.ptr(dw_out),
        _lib.LL(k * k * C if dw_ld is None else dw_ld), num_sms(), _lib.stream_ptr()), "air_conv_s2_wgrad_patch_bf16",
        4 if k == 3 else 1)
    return dw_out


def conv3x3_patch_stats(a, a_ld, B, H, W, C, wpk, N, out, out_ld, res, res_ld, relu, stats):
    """3x3 / s1 / p1 forward that also adds the per-channel sum / sum of squares of its output to `stats` (fp64 [2N])."""
    if _is_f32(out):
        return _patch_taps_ex2(a, a_ld, B, H, W, C, wpk, 9, N, out, out_ld, H, W, res, res_ld, relu, None, None, 0, stats,
                               H, W, -1, -1, 1, 1, 0, 0, 9, _DR3, _DC3, _SL3)
    _lib.check(_lib.lib().air_conv3x3_patch_stats_bf16(
        _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, _lib.ptr(wpk), N, _lib.ptr(out), _lib.LL(out_ld), _lib.ptr(res),
        _lib.LL(res_ld), int(relu), _lib.ptr(stats), num_sms(), _lib.stream_ptr()), "air_conv3x3_patch_stats_bf16")
    return out


def conv_out_size(n, k, s, p, d):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


# ------------------------------------------------------------------------------------------
# BatchNorm / stem / head / loss / optimiser wrappers
# ------------------------------------------------------------------------------------------
def bn_stats(x, x_ld, M, C, sums):
    fn = _lib.lib().air_bn_stats_f32 if _is_f32(x) else _lib.lib().air_bn_stats
    _lib.check(fn(_lib.ptr(x), _lib.LL(x_ld), _lib.LL(M), C, _lib.ptr(sums), num_sms(), _lib.stream_ptr()), "air_bn_stats")


def bn_apply(x, x_ld, y, y_ld, M, C, sums, gamma, beta, relu, training, save_mean, save_invstd,
             running_mean, running_var, eps=1e-5, momentum=0.1):
    if _is_f32(x):
        return bn_apply_add.__wrapped__(x, x_ld, y, y_ld, M, C, sums, gamma, beta, relu, training, save_mean, save_invstd,
                                        running_mean, running_var, None, 0, None, 0, eps, momentum)
    _lib.check(_lib.lib().air_bn_apply(
        _lib.ptr(x), _lib.LL(x_ld), _lib.ptr(y), _lib.LL(y_ld), _lib.LL(M), C, _lib.ptr(sums), _lib.ptr(gamma),
        _lib.ptr(beta), _lib.F(eps), int(relu), int(training), _lib.ptr(save_mean), _lib.ptr(save_invstd),
        _lib.ptr(running_mean), _lib.ptr(running_var), _lib.F(momentum), num_sms(), _lib.stream_ptr()), "air_bn_apply")


def bn_bwd(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum, dgamma, dbeta):
    if _is_f32(x):
        return bn_bwd_bias.__wrapped__(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum,
                                       dgamma, dbeta, None)
    _lib.check(_lib.lib().air_bn_bwd(
        _lib.ptr(dy), _lib.LL(dy_ld), _lib.ptr(x), _lib.LL(x_ld), _lib.ptr(add), _lib.LL(add_ld), _lib.ptr(dx),
        _lib.LL(dx_ld), _lib.LL(M), C, order, _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(gamma), _lib.ptr(beta),
        _lib.ptr(rsum), _lib.ptr(dgamma), _lib.ptr(dbeta), num_sms(), _lib.stream_ptr()), "air_bn_bwd", 2)


def stem_fwd(x, B, H, W, kh, kw, sh, sw, ph, pw, w, cout, y):
    fn = _lib.lib().air_stem_conv_fwd_f32 if _is_f32(x) else _lib.lib().air_stem_conv_fwd
    _lib.check(fn(_lib.ptr(x), B, H, W, kh, kw, sh, sw, ph, pw, _lib.ptr(w), cout, _lib.ptr(y), _lib.stream_ptr()),
               "air_stem_conv_fwd")


def stem_wgrad(x, B, H, W, kh, kw, sh, sw, ph, pw, dy, cout, dw):
    fn = _lib.lib().air_stem_conv_wgrad_f32 if _is_f32(x) else _lib.lib().air_stem_conv_wgrad
    _lib.check(fn(_lib.ptr(x), B, H, W, kh, kw, sh, sw, ph, pw, _lib.ptr(dy), cout, _lib.ptr(dw), _lib.stream_ptr()),
               "air_stem_conv_wgrad")


def selfattn_pool_fwd(x, att, stats, p, th, B, T, C, seed=-1):
    fn = _lib.lib().air_selfattn_pool_fwd_f32 if _is_f32(x) else _lib.lib().air_selfattn_pool_fwd
    _lib.check(fn(_lib.ptr(x), _lib.ptr(att), _lib.ptr(stats), _lib.ptr(p), _lib.ptr(th), B, T, C, _lib.LL(seed),
                  _lib.stream_ptr()), "air_selfattn_pool_fwd")


def selfattn_pool_bwd(x, att, p, th, stats, dstats, dx, datt, B, T, C, seed=-1):
    fn = _lib.lib().air_selfattn_pool_bwd_f32 if _is_f32(x) else _lib.lib().air_selfattn_pool_bwd
    _lib.check(fn(_lib.ptr(x), _lib.ptr(att), _lib.ptr(p), _lib.ptr(th), _lib.ptr(stats), _lib.ptr(dstats), _lib.ptr(dx),
                  _lib.ptr(datt), B, T, C, _lib.LL(seed), _lib.stream_ptr()), "air_selfattn_pool_bwd")


def linear_fwd(x, W, bias, y, M, N, K):
    _lib.check(_lib.lib().air_linear_fwd(_lib.ptr(x), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), M, N, K,
                                         _lib.stream_ptr()), "air_linear_fwd")


def linear_fwd_splitk(x, W, bias, y, M, N, K, scratch, splits):
    """linear_fwd with K dealt to `splits` CTAs per output tile; `scratch`: fp64 device tensor of >= splits*M*N elements."""
    assert scratch.dtype == torch.float64 and scratch.numel() >= splits * M * N, "split-K scratch too small"
    _lib.check(_lib.lib().air_linear_fwd_splitk(_lib.ptr(x), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(y), M, N, K,
                                                _lib.ptr(scratch), splits, _lib.stream_ptr()), "air_linear_fwd_splitk", 2)


def linear_bwd(x, W, dy, dx, dW, db, M, N, K):
    _lib.check(_lib.lib().air_linear_bwd(_lib.ptr(x), _lib.ptr(W), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(dW), _lib.ptr(db),
                                         M, N, K, _lib.stream_ptr()), "air_linear_bwd", 2)


def ocsoftmax(x, labels, center, B, D, r_real, r_fake, alpha, grad_scale, loss, score, dfeat, dcenter,
              logits=None, ncls=0, ce=None):
    _lib.check(_lib.lib().air_ocsoftmax_fwd_bwd(
        _lib.ptr(x), _lib.ptr(labels), _lib.ptr(center), B, D, _lib.F(r_real), _lib.F(r_fake), _lib.F(alpha),
        _lib.F(grad_scale), _lib.ptr(loss), _lib.ptr(score), _lib.ptr(dfeat), _lib.ptr(dcenter), _lib.ptr(logits),
        ncls, _lib.ptr(ce), _lib.stream_ptr()), "air_ocsoftmax_fwd_bwd")


def adam_l2_step(p, g, m, v, n, lr, beta1, beta2, eps, wd, step, grad_scale=1.0):
    _lib.check(_lib.lib().air_adam_l2_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), _lib.LL(n), _lib.F(lr),
                                           _lib.F(beta1), _lib.F(beta2), _lib.F(eps), _lib.F(wd), int(step),
                                           _lib.F(grad_scale), _lib.stream_ptr()), "air_adam_l2_step")


def sgd_step(p, g, n, lr, grad_scale=1.0):
    _lib.check(_lib.lib().air_sgd_step(_lib.ptr(p), _lib.ptr(g), _lib.LL(n), _lib.F(lr), _lib.F(grad_scale),
                                       _lib.stream_ptr()), "air_sgd_step")


# ------------------------------------------------------------------------------------------
# ECAPA-TDNN wrappers (csrc/ecapa.cu and the *_ex / *_ld entry points)
# ------------------------------------------------------------------------------------------
def pack_weights_ld(w, w_ld, mode, cin, cout, taps, out):
    """w: fp32 view whose output-channel rows are w_ld elements apart (column slice of a wider weight)."""
    n, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
    st = _lib.lib().air_conv_pack_weights_ld(_lib.ptr(w), _lib.LL(w_ld), _lib.ptr(out), n, k, mode, cin, cout, taps,
                                             _lib.stream_ptr())
    _lib.check(st, "air_conv_pack_weights_ld")
    return out


def conv_gemm_ex(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K,
                 out, out_ld, bias=None, res=None, res_ld=0, relu=False, flags=0, bias_rows=0, out2=None, out2_ld=0):
    st = _lib.lib().air_conv_gemm_bf16_ex(
        _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode,
        _lib.ptr(wpk), N, K, _lib.ptr(out), _lib.LL(out_ld), _lib.ptr(bias), int(bias_rows), _lib.ptr(res), _lib.LL(res_ld),
        int(relu), _lib.ptr(out2), _lib.LL(out2_ld), num_sms(), flags | _conv_flags(out, res, out2), _lib.stream_ptr())
    _lib.check(st, "air_conv_gemm_bf16_ex")
    return out


def conv_gemm_affine(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode, wpk, N, K,
                     out, out_ld, bias, relu, post_scale, post_shift):
    """fprop with a per-channel affine after the ReLU (eval-mode BatchNorm folded into the conv epilogue)."""
    st = _lib.lib().air_conv_gemm_bf16_affine(
        _lib.ptr(a), _lib.LL(a_ld), B, H, W, C, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, mode,
        _lib.ptr(wpk), N, K, _lib.ptr(out), _lib.LL(out_ld), _lib.ptr(bias), 0, None, _lib.LL(0), int(relu),
        None, _lib.LL(0), _lib.ptr(post_scale), _lib.ptr(post_shift), num_sms(), _conv_flags(out), _lib.stream_ptr())
    _lib.check(st, "air_conv_gemm_bf16_affine")
    return out


def conv_wgrad_ld(x, x_ld, B, H, W, C, dy, dy_ld, Ho, Wo, N, kh, kw, sh, sw, ph, pw, dh, dw, dw_out, dw_ld, flags=0):
    st = _lib.lib().air_conv_wgrad_bf16_ld(
        _lib.ptr(x), _lib.LL(x_ld), B, H, W, C, _lib.ptr(dy), _lib.LL(dy_ld), Ho, Wo, N,
        kh, kw, sh, sw, ph, pw, dh, dw, _lib.ptr(dw_out), _lib.LL(dw_ld), num_sms(), flags, _lib.stream_ptr())
    _lib.check(st, "air_conv_wgrad_bf16_ld")
    return dw_out


def bn_apply_add(x, x_ld, y, y_ld, M, C, sums, gamma, beta, relu, training, save_mean, save_invstd,
                 running_mean, running_var, add, add_ld, y2, y2_ld, eps=1e-5, momentum=0.1):
    fn = _lib.lib().air_bn_apply_add_f32 if _is_f32(x) else _lib.lib().air_bn_apply_add
    _lib.check(fn(
        _lib.ptr(x), _lib.LL(x_ld), _lib.ptr(y), _lib.LL(y_ld), _lib.LL(M), C, _lib.ptr(sums), _lib.ptr(gamma),
        _lib.ptr(beta), _lib.F(eps), int(relu), int(training), _lib.ptr(save_mean), _lib.ptr(save_invstd),
        _lib.ptr(running_mean), _lib.ptr(running_var), _lib.F(momentum), _lib.ptr(add), _lib.LL(add_ld), _lib.ptr(y2),
        _lib.LL(y2_ld), num_sms(), _lib.stream_ptr()), "air_bn_apply_add")


def bn_bwd_bias(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, C, order, mean, invstd, gamma, beta, rsum, dgamma, dbeta,
                dbias):
    fn = _lib.lib().air_bn_bwd_bias_f32 if _is_f32(x) else _lib.lib().air_bn_bwd_bias
    _lib.check(fn(
        _lib.ptr(dy), _lib.LL(dy_ld), _lib.ptr(x), _lib.LL(x_ld), _lib.ptr(add), _lib.LL(add_ld), _lib.ptr(dx),
        _lib.LL(dx_ld), _lib.LL(M), C, order, _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(gamma), _lib.ptr(beta),
        _lib.ptr(rsum), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dbias), num_sms(), _lib.stream_ptr()), "air_bn_bwd_bias", 2)


def linear_fwd_ld(x, W, ldw, bias, y, M, N, K, accumulate=False):
    _lib.check(_lib.lib().air_linear_fwd_ld(_lib.ptr(x), _lib.ptr(W), _lib.LL(ldw), _lib.ptr(bias), _lib.ptr(y), M, N, K,
                                            int(accumulate), _lib.stream_ptr()), "air_linear_fwd_ld")


def linear_bwd_ld(x, W, ldw, dy, dx, dW, db, M, N, K):
    _lib.check(_lib.lib().air_linear_bwd_ld(_lib.ptr(x), _lib.ptr(W), _lib.LL(ldw), _lib.ptr(dy), _lib.ptr(dx), _lib.ptr(dW),
                                            _lib.ptr(db), M, N, K, _lib.stream_ptr()), "air_linear_bwd_ld", 2)


def time_stats(x, x_ld, B, T, C, mean_out, std_out=None, clampv=0.0):
    _lib.check((_lib.lib().air_time_stats_fwd_f32 if _is_f32(x) else _lib.lib().air_time_stats_fwd)(_lib.ptr(x), _lib.LL(x_ld), B, T, C, _lib.ptr(mean_out), _lib.ptr(std_out),
                                             _lib.F(clampv), _lib.stream_ptr()), "air_time_stats_fwd")


def asp_fwd(e, e_ld, x, x_ld, B, T, C, out, smax, ssum, sq):
    _lib.check((_lib.lib().air_ecapa_asp_fwd_f32 if _is_f32(x) else _lib.lib().air_ecapa_asp_fwd)(_lib.ptr(e), _lib.LL(e_ld), _lib.ptr(x), _lib.LL(x_ld), B, T, C, _lib.ptr(out),
                                            _lib.ptr(smax), _lib.ptr(ssum), _lib.ptr(sq), _lib.stream_ptr()), "air_ecapa_asp_fwd")


def asp_bwd(e, e_ld, x, x_ld, B, T, C, out, dout, smax, ssum, sq, cmean, cstd, dcmean, dcstd, clampv, de, de_ld, dx, dx_ld):
    _lib.check((_lib.lib().air_ecapa_asp_bwd_f32 if _is_f32(x) else _lib.lib().air_ecapa_asp_bwd)(
        _lib.ptr(e), _lib.LL(e_ld), _lib.ptr(x), _lib.LL(x_ld), B, T, C, _lib.ptr(out), _lib.ptr(dout), _lib.ptr(smax),
        _lib.ptr(ssum), _lib.ptr(sq), _lib.ptr(cmean), _lib.ptr(cstd), _lib.ptr(dcmean), _lib.ptr(dcstd), _lib.F(clampv),
        _lib.ptr(de), _lib.LL(de_ld), _lib.ptr(dx), _lib.LL(dx_ld), _lib.stream_ptr()), "air_ecapa_asp_bwd")


def ctx_bwd_mask(x, x_ld, B, T, C, cmean, cstd, dcmean, dcstd, clampv, dx, dx_ld):
    _lib.check((_lib.lib().air_ctx_stats_bwd_mask_f32 if _is_f32(x) else _lib.lib().air_ctx_stats_bwd_mask)(_lib.ptr(x), _lib.LL(x_ld), B, T, C, _lib.ptr(cmean), _lib.ptr(cstd),
                                                 _lib.ptr(dcmean), _lib.ptr(dcstd), _lib.F(clampv), _lib.ptr(dx), _lib.LL(dx_ld),
                                                 _lib.stream_ptr()), "air_ctx_stats_bwd_mask")


def scale_residual(x, x_ld, gate, res, res_ld, out, out_ld, B, T, C):
    _lib.check((_lib.lib().air_scale_residual_fwd_f32 if _is_f32(x) else _lib.lib().air_scale_residual_fwd)(_lib.ptr(x), _lib.LL(x_ld), _lib.ptr(gate), _lib.ptr(res), _lib.LL(res_ld),
                                                 _lib.ptr(out), _lib.LL(out_ld), B, T, C, _lib.stream_ptr()),
               "air_scale_residual_fwd")


def se_dgate(dout, d_ld, x, x_ld, B, T, C, dgate):
    _lib.check((_lib.lib().air_se_dgate_f32 if _is_f32(x) else _lib.lib().air_se_dgate)(_lib.ptr(dout), _lib.LL(d_ld), _lib.ptr(x), _lib.LL(x_ld), B, T, C, _lib.ptr(dgate),
                                       _lib.stream_ptr()), "air_se_dgate")


def se_apply_bwd(dout, d_ld, gate, dmean, dx, dx_ld, B, T, C):
    _lib.check((_lib.lib().air_se_apply_bwd_f32 if _is_f32(dout) else _lib.lib().air_se_apply_bwd)(_lib.ptr(dout), _lib.LL(d_ld), _lib.ptr(gate), _lib.ptr(dmean), _lib.ptr(dx),
                                           _lib.LL(dx_ld), B, T, C, _lib.stream_ptr()), "air_se_apply_bwd")


def bn1d_fwd(x, y, M, C, relu_in, gamma, beta, training, save_mean, save_invstd, running_mean, running_var,
             eps=1e-5, momentum=0.1):
    _lib.check(_lib.lib().air_bn1d_f32_fwd(_lib.ptr(x), _lib.ptr(y), M, C, int(relu_in), _lib.ptr(gamma), _lib.ptr(beta),
                                           _lib.F(eps), int(training), _lib.ptr(save_mean), _lib.ptr(save_invstd),
                                           _lib.ptr(running_mean), _lib.ptr(running_var), _lib.F(momentum),
                                           _lib.stream_ptr()), "air_bn1d_f32_fwd")


def bn1d_bwd(dy, x, dx, M, C, relu_in, gamma, mean, invstd, dgamma, dbeta):
    _lib.check(_lib.lib().air_bn1d_f32_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(dx), M, C, int(relu_in), _lib.ptr(gamma),
                                           _lib.ptr(mean), _lib.ptr(invstd), _lib.ptr(dgamma), _lib.ptr(dbeta),
                                           _lib.stream_ptr()), "air_bn1d_f32_bwd")


def sigmoid_fwd(x, y, n):
    _lib.check(_lib.lib().air_sigmoid_fwd(_lib.ptr(x), _lib.ptr(y), _lib.LL(n), _lib.stream_ptr()), "air_sigmoid_fwd")


def sigmoid_bwd(dy, y, dx, n):
    _lib.check(_lib.lib().air_sigmoid_bwd(_lib.ptr(dy), _lib.ptr(y), _lib.ptr(dx), _lib.LL(n), _lib.stream_ptr()), "air_sigmoid_bwd")


def copy_channels(src, s_ld, dst, d_ld, M, C, mask=None, m_ld=0):
    _lib.check((_lib.lib().air_copy_channels_f32 if _is_f32(src) else _lib.lib().air_copy_channels)(_lib.ptr(src), _lib.LL(s_ld), _lib.ptr(mask), _lib.LL(m_ld), _lib.ptr(dst),
                                            _lib.LL(d_ld), _lib.LL(M), C, _lib.stream_ptr()), "air_copy_channels")


def colsum(x, ld, M, C, out):
    _lib.check((_lib.lib().air_colsum_f32 if _is_f32(x) else _lib.lib().air_colsum_bf16)(_lib.ptr(x), _lib.LL(ld), _lib.LL(M), C, _lib.ptr(out), _lib.stream_ptr()),
               "air_colsum_bf16")


# ------------------------------------------------------------------------------------------
# Optional per-kernel-family device timing (bench.py roofline pass): CUDA events recorded on the
# launching stream around every C-ABI call while a Profile is active.  Off by default (no cost).
# ------------------------------------------------------------------------------------------
class Profile:
    def __init__(self):
        self.records = []          # (family, start_event, end_event, flops, bytes)
        self.details = []          # scalar arguments of each recorded call

    def per_call(self):
        torch.cuda.synchronize()
        return [(r[0], r[1].elapsed_time(r[2]), r[3], d) for r, d in zip(self.records, self.details)]

    def __enter__(self):
        global _ACTIVE
        _ACTIVE = self
        return self

    def __exit__(self, *exc):
        global _ACTIVE
        _ACTIVE = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for fam, e0, e1, fl, by in self.records:
            d = out.setdefault(fam, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
            d["launches"] += 1
        return out


_ACTIVE = None


class region:
    """`with ops.region("family"):` -- CUDA events around a block of launches when a Profile is active."""

    def __init__(self, family, flops=0.0, nbytes=0.0):
        self.family, self.flops, self.nbytes = family, flops, nbytes

    def __enter__(self):
        self.prof = _ACTIVE
        if self.prof is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.prof is not None:
            self.e1.record()
            self.prof.records.append((self.family, self.e0, self.e1, self.flops, self.nbytes))
            self.prof.details.append(())


def _conv_work(args):
    # conv_gemm(a, a_ld, B, H, W, C, Ho, Wo, kh, kw, ..., mode, wpk, N, K, ...)
    # algorithmic FLOPs: a strided dgrad enumerates the input grid, of which 1/(sh*sw) of the
    # gathered taps are structurally non-zero
    B, Ho, Wo, N, K = args[2], args[6], args[7], args[18], args[19]
    f = 2.0 * B * Ho * Wo * N * K
    return f / (args[10] * args[11]) if args[16] == 1 else f


def _wgrad_work(args):
    # conv_wgrad(x, x_ld, B, H, W, C, dy, dy_ld, Ho, Wo, N, kh, kw, ...)
    B, C, Ho, Wo, N, kh, kw = args[2], args[5], args[8], args[9], args[10], args[11], args[12]
    return 2.0 * B * Ho * Wo * N * C * kh * kw


def _timed(fn, family, work=None):
    def wrapper(*args, **kw):
        prof = _ACTIVE
        if prof is None:
            return fn(*args, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*args, **kw)
        e1.record()
        fam = family(args) if callable(family) else family
        prof.records.append((fam, e0, e1, work(args) if work else 0.0, 0.0))
        prof.details.append(tuple(a for a in args if isinstance(a, (int, float, bool))))
        return r
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    wrapper.__wrapped__ = fn
    return wrapper


split_terms = _timed(split_terms, "split")
conv_gemm = _timed(conv_gemm, lambda a: "conv_dgrad" if a[16] == 1 else "conv_fprop", _conv_work)
conv_wgrad = _timed(conv_wgrad, "conv_wgrad", _wgrad_work)
bn_stats = _timed(bn_stats, "bn_stats")
bn_apply = _timed(bn_apply, "bn_apply")
bn_bwd = _timed(bn_bwd, "bn_bwd")
stem_fwd = _timed(stem_fwd, "stem_fwd")
stem_wgrad = _timed(stem_wgrad, "stem_wgrad")
selfattn_pool_fwd = _timed(selfattn_pool_fwd, "pool_fwd")
selfattn_pool_bwd = _timed(selfattn_pool_bwd, "pool_bwd")
linear_fwd = _timed(linear_fwd, "linear")
linear_fwd_splitk = _timed(linear_fwd_splitk, "linear")
linear_bwd = _timed(linear_bwd, "linear")
ocsoftmax = _timed(ocsoftmax, "ocsoftmax")
adam_l2_step = _timed(adam_l2_step, "optim")
sgd_step = _timed(sgd_step, "optim")
pack_weights = _timed(pack_weights, "pack_weights")
conv_gemm_ex = _timed(conv_gemm_ex, lambda a: "conv_dgrad" if a[16] == 1 else "conv_fprop", _conv_work)
conv_wgrad_ld = _timed(conv_wgrad_ld, "conv_wgrad", _wgrad_work)
pack_weights_ld = _timed(pack_weights_ld, "pack_weights")
bn_apply_add = _timed(bn_apply_add, "bn_apply")
bn_bwd_bias = _timed(bn_bwd_bias, "bn_bwd")
linear_fwd_ld = _timed(linear_fwd_ld, "linear")
linear_bwd_ld = _timed(linear_bwd_ld, "linear")
time_stats = _timed(time_stats, "time_stats")
asp_fwd = _timed(asp_fwd, "asp_pool")
asp_bwd = _timed(asp_bwd, "asp_pool")
scale_residual = _timed(scale_residual, "se")
se_dgate = _timed(se_dgate, "se")
se_apply_bwd = _timed(se_apply_bwd, "se")
bn1d_fwd = _timed(bn1d_fwd, "bn1d")
bn1d_bwd = _timed(bn1d_bwd, "bn1d")
sigmoid_fwd = _timed(sigmoid_fwd, "se")
sigmoid_bwd = _timed(sigmoid_bwd, "se")
copy_channels = _timed(copy_channels, "copy")
colsum = _timed(colsum, "colsum")
ctx_bwd_mask = _timed(ctx_bwd_mask, "asp_pool")
pack3x3 = _timed(pack3x3, "pack_weights")
conv3x3_patch = _timed(conv3x3_patch, lambda a: "conv_dgrad" if (len(a) > 13 and a[13] == 1) else "conv_fprop",
                       lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[7] * 9)
conv3x3_wgrad_patch = _timed(conv3x3_wgrad_patch, "conv_wgrad", lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[8] * 9)
pack_patch = _timed(pack_patch, "pack_weights")
conv1x1_patch = _timed(conv1x1_patch, lambda a: "conv_dgrad" if (len(a) > 13 and a[13] == 1) else "conv_fprop",
                       lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[7])
conv_s2_dgrad_patch = _timed(conv_s2_dgrad_patch, "conv_dgrad", lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[8] * a[7] * a[7])
PackPlan.run = _timed(PackPlan.run, "pack_weights")
conv_wgrad_patch = _timed(conv_wgrad_patch, "conv_wgrad", lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[8] * a[9] * a[9])
conv_gemm_affine = _timed(conv_gemm_affine, "conv_fprop", _conv_work)
conv_gemm_stats = _timed(conv_gemm_stats, "conv_fprop", _conv_work)
conv1d_patch = _timed(conv1d_patch, lambda a: "conv_dgrad" if (len(a) > 18 and a[18] == 1) else "conv_fprop",
                      lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[9] * a[7])
conv1d_wgrad_patch = _timed(conv1d_wgrad_patch, "conv_wgrad", lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[8] * a[9])
conv3x3_patch_stats = _timed(conv3x3_patch_stats, "conv_fprop", lambda a: 2.0 * a[2] * a[3] * a[4] * a[5] * a[7] * 9)
conv_s2_wgrad_patch = _timed(conv_s2_wgrad_patch, "conv_wgrad", lambda a: 2.0 * a[2] * a[8] * a[9] * a[5] * a[10] * a[11] * a[11])


# ------------------------------------------------------------------------------------------
# detection-error curve / EER / t-DCF (csrc/det.cu)
# ------------------------------------------------------------------------------------------
def det_workspace_bytes(n):
    out = ctypes.c_longlong(0)
    st = _lib.lib().air_det_workspace_bytes(_lib.LL(n), ctypes.byref(out))
    if st != 0:
        raise _lib.AirError("air_det_workspace_bytes failed: argument error %d" % st)
    return int(out.value)


def det_curve(target, nontarget, negate, c1, c2, want_tdcf, workspace, frr, far, thresholds, tdcf, out):
    """target / nontarget: contiguous CUDA score vectors of one dtype (float32 or float64)."""
    f64 = target.dtype == torch.float64
    fn = _lib.lib().air_det_curve_f64 if f64 else _lib.lib().air_det_curve_f32
    n = target.numel() + nontarget.numel()
    _lib.check(fn(_lib.ptr(target), _lib.LL(target.numel()), _lib.ptr(nontarget), _lib.LL(nontarget.numel()),
                  int(bool(negate)), _lib.D(c1), _lib.D(c2), int(bool(want_tdcf)), _lib.ptr(workspace),
                  _lib.LL(workspace.numel() * workspace.element_size()), _lib.ptr(frr), _lib.ptr(far),
                  _lib.ptr(thresholds), _lib.ptr(tdcf), _lib.ptr(out), _lib.stream_ptr()),
               "air_det_curve", n=_lib.lib().air_det_launches(_lib.LL(n), int(f64)))


def det_threshold_counts(scores, threshold, counts):
    fn = _lib.lib().air_det_threshold_counts_f64 if scores.dtype == torch.float64 else _lib.lib().air_det_threshold_counts_f32
    _lib.check(fn(_lib.ptr(scores), _lib.LL(scores.numel()), _lib.D(threshold), _lib.ptr(counts), _lib.stream_ptr()),
               "air_det_threshold_counts")


det_curve = _timed(det_curve, "det_curve")


# ------------------------------------------------------------------------------------------
# adversarial channel-classifier head (csrc/adv.cu)
# ------------------------------------------------------------------------------------------
def dropout_relu_fwd(x, keep, y, p, generate, seed=0):
    _lib.check(_lib.lib().air_dropout_relu_fwd(_lib.ptr(x), _lib.ptr(keep), _lib.ptr(y), _lib.LL(x.numel()), _lib.F(p),
                                               int(bool(generate)), ctypes.c_ulonglong(int(seed)), _lib.stream_ptr()),
               "air_dropout_relu_fwd")


def dropout_relu_bwd(dy, x, keep, dx, p):
    _lib.check(_lib.lib().air_dropout_relu_bwd(_lib.ptr(dy), _lib.ptr(x), _lib.ptr(keep), _lib.ptr(dx), _lib.LL(x.numel()),
                                               _lib.F(p), _lib.stream_ptr()), "air_dropout_relu_bwd")


def relu_ce_fwd_bwd(z, labels, B, C, grad_scale, loss_sum, correct, dz):
    _lib.check(_lib.lib().air_relu_ce_fwd_bwd(_lib.ptr(z), _lib.ptr(labels), int(B), int(C), _lib.F(grad_scale),
                                              _lib.ptr(loss_sum), _lib.ptr(correct), _lib.ptr(dz), _lib.stream_ptr()),
               "air_relu_ce_fwd_bwd")
