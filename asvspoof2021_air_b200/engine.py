"""Hand-scheduled forward/backward engines over flat device buffers.

PyTorch is used here for device memory, streams and (in parallel.py) torch.distributed only: every
arithmetic step is one of the hand-written sm_100a kernels behind the C ABI (ops.py).  Parameters
live in ONE flat fp32 buffer in "engine layout" (conv weights as [Cout][kh][kw][Cin], the GEMM
layout the tensor-core kernels consume); the nn.Module drop-ins expose them under the reference's
state_dict names as (permuted) views of that buffer, so the optimiser, the NCCL all-reduce and
checkpoints all work on the same storage.
"""
import math
import os

import torch

from . import ops

BF16 = torch.bfloat16


class ParamStore:
    """Flat fp32 parameter / gradient / Adam-state buffers with named views.

    entries: (name, engine_shape, to_pt) where to_pt maps the engine-layout view to the PyTorch
    layout view (e.g. a permute for conv weights).  Entries listed in `frozen` are placed after
    the trainable prefix [0, n_train)."""

    def __init__(self, entries, device, frozen=()):
        entries = [e for e in entries if e[0] not in frozen] + [e for e in entries if e[0] in frozen]
        self.offsets, off = {}, 0
        for name, shape, _ in entries:
            n = int(math.prod(shape))
            self.offsets[name] = (off, n, tuple(shape))
            off = (off + n + 3) // 4 * 4            # keep every tensor 16-byte aligned
            if name not in frozen:
                self.n_train = off
        self.total = off
        self.entries = entries
        self.device = device
        self.params = torch.zeros(self.total, device=device)
        self.grads = torch.zeros(self.total, device=device)
        self.exp_avg = None
        self.exp_avg_sq = None
        self.step = 0
        self._to_pt = {name: fn for name, _, fn in entries}

    def view(self, name, buf=None):
        off, n, shape = self.offsets[name]
        return (self.params if buf is None else buf)[off:off + n].view(shape)

    def grad(self, name):
        return self.view(name, self.grads)

    def pt_view(self, name, buf=None):
        return self._to_pt[name](self.view(name, buf))

    def names(self):
        return [e[0] for e in self.entries]

    def adam_step(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4, grad_scale=1.0):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros(self.n_train, device=self.device)
            self.exp_avg_sq = torch.zeros(self.n_train, device=self.device)
        self.step += 1
        ops.adam_l2_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.n_train, lr, beta1, beta2,
                         eps, weight_decay, self.step, grad_scale)


def _ident(t):
    return t


def _conv2d_pt(t):          # [Cout][kh][kw][Cin] -> (Cout, Cin, kh, kw)
    return t.permute(0, 3, 1, 2)


def _conv1d_pt(t):          # [Cout][k][Cin] -> (Cout, Cin, k)
    return t.permute(0, 2, 1)


class ConvLayer:
    """One convolution: geometry, weight views, packed bf16 operands, fprop / dgrad / wgrad launches."""
    # BatchNorm statistics in the epilogue of the generic kernel as well (AIR_GEMM_STATS=0: separate bn_stats launches)
    gemm_stats = os.environ.get("AIR_GEMM_STATS", "1") != "0"

    def __init__(self, store, name, cin, cout, kh, kw, sh=1, sw=1, ph=0, pw=0, dh=1, dw=1,
                 bias=False, need_dgrad=True, cin_pad=None, need_fprop=True):
        self.store, self.name = store, name
        self.need_fprop = need_fprop
        self.cin, self.cout = cin, cout
        self.cin_g = cin_pad or cin                 # channels the gather sees (ECAPA conv1: 60 -> 64)
        self.kh, self.kw, self.sh, self.sw, self.ph, self.pw, self.dh, self.dw = kh, kw, sh, sw, ph, pw, dh, dw
        self.taps = kh * kw
        self.bias = bias
        self.need_dgrad = need_dgrad
        dev = store.device
        self.K = self.taps * self.cin_g
        self.wpk = torch.empty(ops.packed_elems(cout, self.K), device=dev, dtype=BF16)
        self.wpk_d = torch.empty(ops.packed_elems(self.cin_g, self.taps * cout), device=dev, dtype=BF16) if need_dgrad else None
        self._wtmp = torch.zeros(cout, self.taps, self.cin_g, device=dev) if self.cin_g != cin else None
        # 3x3 / stride 1 / pad 1 layers may run on the patch-resident kernel (chosen per call from the image size)
        geom = (kh, kw, sh, sw, ph, pw, dh, dw)
        plain = not bias and self._wtmp is None
        self.patch_ok = (geom == (3, 3, 1, 1, 1, 1, 1, 1) and plain
                         and ops.patch_supported(cin, cout, 1, 1) and ops.patch_supported(cout, cin, 1, 1)
                         and cin <= 256 and cout <= 256)
        self.wpatch_ok = (geom in ((3, 3, 1, 1, 1, 1, 1, 1), (1, 1, 1, 1, 0, 0, 1, 1)) and self._wtmp is None
                          and ops.wgrad_patch_supported(cin, cout))
        # stride-2 layers: data gradient by output parity classes on the patch kernel (3x3 pad 1, or the 1x1 shortcut)
        self.s2_dgrad_ok = (need_dgrad and plain and geom in ((3, 3, 2, 2, 1, 1, 1, 1), (1, 1, 2, 2, 0, 0, 1, 1))
                            and ops.patch_supported(cout, cin, 1, 1))
        # stride-2 layers: weight gradient as one stride-1 patch problem per input parity class (strided TMA sub-images)
        self.s2_wgrad_ok = (geom in ((3, 3, 2, 2, 1, 1, 1, 1), (1, 1, 2, 2, 0, 0, 1, 1)) and self._wtmp is None
                            and cin % 64 == 0 and ops.wgrad_patch_supported(cin, cout))
        # 1x1 / stride-1 layers: a GEMM over pixels through the same TMA kernel (single tap)
        self.p1x1_ok = (geom == (1, 1, 1, 1, 0, 0, 1, 1) and plain and ops.patch_supported(cin, cout, 1, 1)
                        and (not need_dgrad or ops.patch_supported(cout, cin, 1, 1)))
        # dilated 1-D layers (k = 3, "same" padding, reach d*(k-1) <= 8): the Res2 branches of ECAPA
        self.d1_ok = (kh == 1 and kw in (3, 5) and (sh, sw, ph, dh) == (1, 1, 0, 1) and pw == dw * (kw - 1) // 2
                      and dw * (kw - 1) <= 8 and self._wtmp is None and cin % 64 == 0 and cout % 64 == 0 and cout <= 256
                      and ops.patch_supported(cin, cout, 1, 1) and ops.patch_supported(cout, cin, 1, 1)
                      and ops.wgrad_patch_supported(cin, cout))
        n = self.taps * cin * cout
        self.wpk3 = torch.empty(n, device=dev, dtype=BF16) if (self.patch_ok or self.p1x1_ok or self.d1_ok) else None
        self.wpk3_d = torch.empty(n, device=dev, dtype=BF16) if (need_dgrad and (self.patch_ok or self.s2_dgrad_ok
                                                                                 or self.p1x1_ok or self.d1_ok)) else None

    @staticmethod
    def _patch_efficiency(H, W):
        return (W / (-(-W // 128) * 128)) * (H / (-(-H // 2) * 2))

    def use_patch(self, H, W):
        return self.patch_ok and self._patch_efficiency(H, W) >= 0.6

    def out_hw(self, H, W):
        return (ops.conv_out_size(H, self.kh, self.sh, self.ph, self.dh),
                ops.conv_out_size(W, self.kw, self.sw, self.pw, self.dw))

    def weight(self):
        return self.store.view(self.name + ".weight")

    def _pack_source(self):
        w = self.weight().reshape(self.cout, self.taps, self.cin)
        if self._wtmp is not None:                  # zero-padded input channels
            self._wtmp[:, :, :self.cin].copy_(w)
            w = self._wtmp
        return w.reshape(-1)

    def pack(self):
        """Per-tensor launches (used when the layer is not part of a PackPlan)."""
        w = self._pack_source()
        if self.need_fprop:
            ops.pack_weights(w, 0, self.cin_g, self.cout, self.taps, self.wpk)
        if self.need_dgrad:
            ops.pack_weights(w, 1, self.cin_g, self.cout, self.taps, self.wpk_d)
        if self.wpk3 is not None and self.need_fprop:
            ops.pack_patch(w, self.cin, self.cout, self.taps, 0, self.wpk3)
        if self.wpk3_d is not None:
            ops.pack_patch(w, self.cout, self.cin, self.taps, 1, self.wpk3_d)

    def add_pack_jobs(self, plan):
        """Register this layer's packings in a batched PackPlan; returns False when the layer needs its own path."""
        if self._wtmp is not None:
            return False
        w = self.weight().reshape(-1)
        plan.add_gemm(w, self.taps * self.cin_g, 0, self.cin_g, self.cout, self.taps, self.wpk)
        if self.need_dgrad:
            plan.add_gemm(w, self.taps * self.cin_g, 1, self.cin_g, self.cout, self.taps, self.wpk_d)
        if self.wpk3 is not None:
            plan.add_patch(w, self.cin, self.cout, self.taps, 0, self.wpk3)
        if self.wpk3_d is not None:
            plan.add_patch(w, self.cout, self.cin, self.taps, 1, self.wpk3_d)
        return True

    def fprop(self, x, x_ld, B, H, W, out, out_ld, res=None, res_ld=0, relu=False, stats=None):
        """stats (optional, fp64 [2*cout]): when the layer runs on the patch kernel the per-channel sum / sum of squares of
        the output are accumulated there by the conv epilogue and `self.stats_fused` is set (bn_stats can be skipped)."""
        Ho, Wo = self.out_hw(H, W)
        self.stats_fused = False
        if self.use_patch(H, W):
            if stats is not None and self.cout % 32 == 0:
                ops.conv3x3_patch_stats(x, x_ld, B, H, W, self.cin, self.wpk3, self.cout, out, out_ld, res, res_ld, relu, stats)
                self.stats_fused = True
            else:
                ops.conv3x3_patch(x, x_ld, B, H, W, self.cin, self.wpk3, self.cout, out, out_ld, res, res_ld, relu, 0)
            return Ho, Wo
        if self.p1x1_ok:
            ops.conv1x1_patch(x, x_ld, B, H, W, self.cin, self.wpk3, self.cout, out, out_ld, res, res_ld, relu, 0)
            return Ho, Wo
        bias = self.store.view(self.name + ".bias") if self.bias else None
        if self.d1_ok:
            fuse = stats is not None and self.cout % 32 == 0 and out.dtype == BF16
            ops.conv1d_patch(x, x_ld, B, H, W, self.cin, self.wpk3, self.kw, self.dw, self.cout, out, out_ld, bias,
                             res, res_ld, relu, None, 0, 0, None, None, stats if fuse else None)
            self.stats_fused = fuse
            return Ho, Wo
        if (stats is not None and self.gemm_stats and out.dtype == BF16 and self.cout % 32 == 0 and (self.cout <= 256 or self.cout % 256 == 0)
                and not (self.kh == 1 and self.kw == 1 and self.sh == 1 and self.sw == 1)):
            # generic kernel (gather mode): the statistics of the BatchNorm that follows ride its epilogue too
            ops.conv_gemm_stats(x, x_ld, B, H, W, self.cin_g, Ho, Wo, self.kh, self.kw, self.sh, self.sw, self.ph, self.pw,
                                self.dh, self.dw, 0, self.wpk, self.cout, self.K, out, out_ld, bias, res, res_ld, relu, stats)
            self.stats_fused = True
            return Ho, Wo
        ops.conv_gemm(x, x_ld, B, H, W, self.cin_g, Ho, Wo, self.kh, self.kw, self.sh, self.sw, self.ph, self.pw,
                      self.dh, self.dw, 0, self.wpk, self.cout, self.K, out, out_ld, bias, res, res_ld, relu)
        return Ho, Wo

    def fprop_affine(self, x, x_ld, B, H, W, out, out_ld, relu, scale, shift, add=None, add_ld=0, out_sum=None, out_sum_ld=0):
        """conv (+bias) -> [ReLU] -> per-channel affine, in one launch (eval-mode conv -> ReLU -> BN).
        add / out_sum (dilated 1-D layers on the patch kernel only): out_sum = round(out) + add, the next Res2 branch's
        input (ecapa_tdnn.py:77-80)."""
        Ho, Wo = self.out_hw(H, W)
        bias = self.store.view(self.name + ".bias") if self.bias else None
        if self.d1_ok and out.dtype == BF16 and self.cout % 32 == 0:
            if out_sum is not None:          # kernel roles: out2 = the affine output, out = rounded output + residual
                ops.conv1d_patch(x, x_ld, B, H, W, self.cin, self.wpk3, self.kw, self.dw, self.cout, out_sum, out_sum_ld, bias,
                                 add, add_ld, relu, out, out_ld, 0, scale, shift)
            else:
                ops.conv1d_patch(x, x_ld, B, H, W, self.cin, self.wpk3, self.kw, self.dw, self.cout, out, out_ld, bias,
                                 None, 0, relu, None, 0, 0, scale, shift)
            return Ho, Wo
        assert out_sum is None, "fused next-branch sum needs the patch kernel"
        ops.conv_gemm_affine(x, x_ld, B, H, W, self.cin_g, Ho, Wo, self.kh, self.kw, self.sh, self.sw, self.ph, self.pw,
                             self.dh, self.dw, 0, self.wpk, self.cout, self.K, out, out_ld, bias, relu, scale, shift)
        return Ho, Wo

    def dgrad(self, dy, dy_ld, B, H, W, dx, dx_ld, accumulate=False, res=None, res_ld=0, out2=None, out2_ld=0):
        """dy on the (Ho,Wo) grid -> dx on the (H,W) input grid.  accumulate adds into dx; `res` adds
        another tensor instead; out2 (optional) receives the gradient without the residual."""
        Ho, Wo = self.out_hw(H, W)
        if accumulate:
            res, res_ld = dx, dx_ld
        if out2 is None and self.use_patch(H, W):
            ops.conv3x3_patch(dy, dy_ld, B, H, W, self.cout, self.wpk3_d, self.cin, dx, dx_ld, res, res_ld, False, 1)
            return
        if out2 is None and self.p1x1_ok:
            ops.conv1x1_patch(dy, dy_ld, B, H, W, self.cout, self.wpk3_d, self.cin, dx, dx_ld, res, res_ld, False, 1)
            return
        if self.d1_ok:
            ops.conv1d_patch(dy, dy_ld, B, H, W, self.cout, self.wpk3_d, self.kw, self.dw, self.cin, dx, dx_ld, None,
                             res, res_ld, False, out2, out2_ld, 1)
            return
        if out2 is None and self.s2_dgrad_ok and (self.kh == 3 or accumulate):
            ops.conv_s2_dgrad_patch(dy, dy_ld, B, Ho, Wo, self.cout, self.wpk3_d, self.kh, self.cin, dx, dx_ld, H, W,
                                    res, res_ld)
            return
        ops.conv_gemm_ex(dy, dy_ld, B, Ho, Wo, self.cout, H, W, self.kh, self.kw, self.sh, self.sw, self.ph, self.pw,
                         self.dh, self.dw, 1, self.wpk_d, self.cin_g, self.taps * self.cout, dx, dx_ld, None,
                         res, res_ld, False, 0, 0, out2, out2_ld)

    def wgrad(self, x, x_ld, B, H, W, dy, dy_ld):
        Ho, Wo = self.out_hw(H, W)
        if self.d1_ok:
            ops.conv1d_wgrad_patch(x, x_ld, B, H, W, self.cin, dy, dy_ld, self.cout, self.kw, self.dw,
                                   self.store.grad(self.name + ".weight"))
            return
        if self.s2_wgrad_ok and self._patch_efficiency(Ho, Wo) >= 0.5:
            ops.conv_s2_wgrad_patch(x, x_ld, B, H, W, self.cin, dy, dy_ld, Ho, Wo, self.cout, self.kh,
                                    self.store.grad(self.name + ".weight"))
            return
        if self.wpatch_ok and self._patch_efficiency(H, W) >= 0.5:
            ops.conv_wgrad_patch(x, x_ld, B, H, W, self.cin, dy, dy_ld, self.cout, self.kh,
                                 self.store.grad(self.name + ".weight"))
            return
        if self._wtmp is not None:
            g = self._gtmp if hasattr(self, "_gtmp") else None
            if g is None:
                g = self._gtmp = torch.zeros(self.cout, self.taps * self.cin_g, device=self.store.device)
            g.zero_()
            ops.conv_wgrad(x, x_ld, B, H, W, self.cin_g, dy, dy_ld, Ho, Wo, self.cout, self.kh, self.kw, self.sh,
                           self.sw, self.ph, self.pw, self.dh, self.dw, g)
            self.store.grad(self.name + ".weight").view(self.cout, self.taps, self.cin).add_(
                g.view(self.cout, self.taps, self.cin_g)[:, :, :self.cin])
        else:
            ops.conv_wgrad(x, x_ld, B, H, W, self.cin_g, dy, dy_ld, Ho, Wo, self.cout, self.kh, self.kw, self.sh,
                           self.sw, self.ph, self.pw, self.dh, self.dw, self.store.grad(self.name + ".weight"))


class Scratch:
    """Grow-only device scratch buffers of an engine (the split operands of the fp32 parity mode)."""

    def __init__(self, device):
        self.device, self.bufs = device, {}

    def get(self, key, numel, dtype=BF16):
        b = self.bufs.get(key)
        if b is None or b.numel() < numel or b.dtype != dtype:
            b = self.bufs[key] = torch.empty(int(numel), device=self.device, dtype=dtype)
        return b[:numel]


class _DerivedStore:
    """Stands in for the ParamStore inside the helper layers of a SplitConvLayer: `<name>.weight` resolves to a derived
    fp32 tensor (split terms of the master weight), every other name to the real store."""

    def __init__(self, real, name, weight):
        self.real, self.key, self.weight, self.device = real, name + ".weight", weight, real.device

    def view(self, name):
        return self.weight if name == self.key else self.real.view(name)


class SplitConvLayer(ConvLayer):
    """fp32 parity mode of one convolution: float activations / gradients in and out, the contraction
    on the SAME bf16 tcgen05 kernels with 3-term split operands concatenated along the contraction axis
    (csrc/split.cu):

        fprop : [hi_x | lo_x | hi_x] (3 Cin)  *  [hi_w | hi_w | lo_w]           one launch, fp32 epilogue
        dgrad : [hi_dy | lo_dy | hi_dy] (3 Cout) * [hi_w ; hi_w ; lo_w]         one launch, fp32 epilogue
        wgrad : (hi_x, hi_dy) + (lo_x, hi_dy) + (hi_x, lo_dy)                   three launches into the fp32 gradient

    Each product is exact in the fp32 accumulator; the dropped lo*lo term and the 16 significant bits of a hi + lo pair
    leave a relative error of ~2^-16 per layer instead of the 2^-9 of a bf16 operand."""
    X_MASK, W_MASK = 0b010, 0b100          # lo positions in [hi, lo, hi] and [hi, hi, lo]

    def __init__(self, store, name, cin, cout, kh, kw, sh=1, sw=1, ph=0, pw=0, dh=1, dw=1,
                 bias=False, need_dgrad=True, cin_pad=None, scratch=None):
        super().__init__(store, name, cin, cout, kh, kw, sh, sw, ph, pw, dh, dw, bias, need_dgrad, cin_pad, need_fprop=False)
        self.scratch = scratch
        dev, cg = store.device, self.cin_g
        self.wf = torch.zeros(cout, kh, kw, 3 * cg, device=dev)
        self.f = ConvLayer(_DerivedStore(store, name, self.wf), name, 3 * cg, cout, kh, kw, sh, sw, ph, pw, dh, dw,
                           bias=bias, need_dgrad=False)
        self.d = None
        if need_dgrad:
            self.wd = torch.zeros(3 * cout, kh, kw, cg, device=dev)
            self.d = ConvLayer(_DerivedStore(store, name, self.wd), name, cg, 3 * cout, kh, kw, sh, sw, ph, pw, dh, dw,
                               bias=False, need_dgrad=True, need_fprop=False)
        self.wpk = self.wpk_d = self.wpk3 = self.wpk3_d = None        # the plain operands are never used in this mode

    def add_pack_jobs(self, plan):
        return False                                  # packed per layer by pack(): derived weights first

    def pack(self):
        w = self._pack_source()                       # fp32 [cout][taps][cin_g]
        rows, cg = self.cout * self.taps, self.cin_g
        ops.split_terms(w, cg, rows, cg, self.wf, 3 * cg, 3, self.W_MASK)
        self.f.pack()
        if self.d is not None:
            n = w.numel()
            ops.split_terms(w, n, 1, n, self.wd, 3 * n, 3, self.W_MASK)       # three whole-tensor copies: [hi ; hi ; lo]
            self.d.pack()

    def _split(self, key, t, ld, M, C, nterms, mask):
        buf = self.scratch.get(key, M * nterms * C).view(M, nterms * C)
        return ops.split_terms(t, ld, M, C, buf, nterms * C, nterms, mask)

    def fprop(self, x, x_ld, B, H, W, out, out_ld, res=None, res_ld=0, relu=False, stats=None):
        xs = self._split("a", x, x_ld, B * H * W, self.cin_g, 3, self.X_MASK)
        r = self.f.fprop(xs, 3 * self.cin_g, B, H, W, out, out_ld, res, res_ld, relu, stats)
        self.stats_fused = self.f.stats_fused
        return r

    def fprop_affine(self, x, x_ld, B, H, W, out, out_ld, relu, scale, shift):
        xs = self._split("a", x, x_ld, B * H * W, self.cin_g, 3, self.X_MASK)
        return self.f.fprop_affine(xs, 3 * self.cin_g, B, H, W, out, out_ld, relu, scale, shift)

    def dgrad(self, dy, dy_ld, B, H, W, dx, dx_ld, accumulate=False, res=None, res_ld=0, out2=None, out2_ld=0):
        Ho, Wo = self.out_hw(H, W)
        dys = self._split("a", dy, dy_ld, B * Ho * Wo, self.cout, 3, self.X_MASK)
        self.d.dgrad(dys, 3 * self.cout, B, H, W, dx, dx_ld, accumulate, res, res_ld, out2, out2_ld)

    def wgrad(self, x, x_ld, B, H, W, dy, dy_ld):
        Ho, Wo = self.out_hw(H, W)
        cg, co = self.cin_g, self.cout
        xs = self._split("a", x, x_ld, B * H * W, cg, 2, 0b10)                  # [hi | lo]
        dys = self._split("b", dy, dy_ld, B * Ho * Wo, co, 2, 0b10)
        for xi, di in ((0, 0), (1, 0), (0, 1)):
            super().wgrad(xs[:, xi * cg:], 2 * cg, B, H, W, dys[:, di * co:], 2 * co)


class BNLayer:
    """BatchNorm over channels-last rows; owns its statistics workspaces (slices of shared buffers)."""

    def __init__(self, store, buffers, name, C):
        self.store, self.name, self.C = store, name, C
        self.running_mean = buffers.add(name + ".running_mean", C, 0.0)
        self.running_var = buffers.add(name + ".running_var", C, 1.0)
        self.num_batches_tracked = torch.zeros((), dtype=torch.long, device=store.device)
        self.sums = buffers.add_f64(2 * C)
        self.rsum = buffers.add_f64(2 * C)
        self.save_mean = torch.empty(C, device=store.device)
        self.save_invstd = torch.empty(C, device=store.device)

    def forward(self, x, x_ld, y, y_ld, M, relu, training, have_stats=False):
        """have_stats: the producing conv already accumulated sum / sum of squares of x into self.sums."""
        g, b = self.store.view(self.name + ".weight"), self.store.view(self.name + ".bias")
        if training:
            if not have_stats:
                ops.bn_stats(x, x_ld, M, self.C, self.sums)
            self.num_batches_tracked += 1
        ops.bn_apply(x, x_ld, y, y_ld, M, self.C, self.sums, g, b, relu, training, self.save_mean, self.save_invstd,
                     self.running_mean, self.running_var)

    def backward(self, dy, dy_ld, x, x_ld, dx, dx_ld, M, order, add=None, add_ld=0):
        g, b = self.store.view(self.name + ".weight"), self.store.view(self.name + ".bias")
        ops.bn_bwd(dy, dy_ld, x, x_ld, add, add_ld, dx, dx_ld, M, self.C, order, self.save_mean, self.save_invstd,
                   g, b, self.rsum, self.store.grad(self.name + ".weight"), self.store.grad(self.name + ".bias"))


class BufferStore:
    """Running statistics (flat fp32) and per-step fp64 reduction workspaces (flat, zeroed per step)."""

    def __init__(self, device):
        self.device = device
        self._f32, self._f32_items, self._n32 = None, [], 0
        self._f64_items, self._n64 = [], 0
        self.f64 = None

    def add(self, name, n, init):
        self._f32_items.append((name, self._n32, n, init))
        self._n32 += n
        return _Lazy(self, "f32", len(self._f32_items) - 1)

    def add_f64(self, n):
        self._f64_items.append((self._n64, n))
        self._n64 += n
        return _Lazy(self, "f64", len(self._f64_items) - 1)

    def finalize(self):
        self._f32 = torch.zeros(max(self._n32, 1), device=self.device)
        for name, off, n, init in self._f32_items:
            self._f32[off:off + n] = init
        self.f64 = torch.zeros(max(self._n64, 1), device=self.device, dtype=torch.float64)

    def resolve(self, kind, idx):
        if kind == "f32":
            _, off, n, _ = self._f32_items[idx]
            return self._f32[off:off + n]
        off, n = self._f64_items[idx]
        return self.f64[off:off + n]

    def named_f32(self):
        return {name: self._f32[off:off + n] for name, off, n, _ in self._f32_items}


class _Lazy:
    """Placeholder for a slice of a buffer that is allocated in BufferStore.finalize()."""

    def __init__(self, owner, kind, idx):
        self.owner, self.kind, self.idx = owner, kind, idx


def _resolve_lazies(obj):
    for k, v in list(vars(obj).items()):
        if isinstance(v, _Lazy):
            setattr(obj, k, v.owner.resolve(v.kind, v.idx))


class AsyncWgrad:
    """Mixin of the engines: weight-gradient launches on a side stream."""
    # Weight gradients are leaves of the backward graph (they only feed the optimiser), so they run on a side stream:
    # the tensor-core bound wgrad kernels then overlap the HBM-bound BatchNorm-backward kernels of the main chain.
    overlap_wgrad = os.environ.get("AIR_OVERLAP_WGRAD", "1") != "0"
    _side = None
    _side_pending = False

    def _wgrad_async(self, fn, *args):
        if not self.overlap_wgrad:
            fn(*args)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record()                                   # the output gradient (and zero_grad) precede this point on the main stream
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            fn(*args)
        self._side_pending = True

    # Weight packing (fp32 master weights -> pre-swizzled bf16 operand tiles, once per optimiser step) also runs on the side
    # stream: launched by the Trainer BEFORE the LFCC kernel, it overlaps the front end and the stem, which need no packed
    # weights; the compute stream waits for it in front of the first tensor-core layer.
    _pack_event = None

    def prepack(self):
        """Start packing the current weights on the side stream (no-op when they are already packed)."""
        if self._packed_version == self.store.step:
            return
        if not self.overlap_wgrad:
            self.pack_weights()
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        ev = torch.cuda.Event()
        ev.record()                                   # the optimiser step and the last users of the old tiles precede this point
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            self.pack_weights()
            self._pack_event = torch.cuda.Event()
            self._pack_event.record()

    def _await_pack(self):
        if self._pack_event is not None:
            torch.cuda.current_stream().wait_event(self._pack_event)
            self._pack_event = None

    def _side_events(self):
        if not self._side_pending:
            return ()
        ev = torch.cuda.Event()
        ev.record(self._side)
        return (ev,)

    def _join_side(self):
        if self._side_pending:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_pending = False



# ==========================================================================================
# ResNet-18 (pre-activation) + SelfAttention pooling      resnet.py:49-69,122-191
# ==========================================================================================
class _Block:
    pass


class ResNetEngine(AsyncWgrad):
    """Forward / backward of ResNet(num_nodes=3, enc_dim, '18', nclasses) for a fixed batch size on
    channels-last bf16 activations.  Input: (B, 60, T) bf16 from the fused LFCC kernel."""

    def __init__(self, enc_dim=256, nclasses=2, device="cuda", train_head_mu=False, F=60, num_nodes=3, precision="bf16"):
        """precision: "bf16" (the product path: bf16 activations, bf16 tensor-core operands) or "fp32" (parity mode:
        float activations, 3-term split operands on the same kernels; several times slower, used to show the residual
        against the reference's fp32 arithmetic is rounding)."""
        assert precision in ("bf16", "fp32")
        self.precision = precision
        self.act_dtype = BF16 if precision == "bf16" else torch.float32
        self.B, self.T, self.F = None, None, F
        self.enc_dim, self.nclasses, self.num_nodes = enc_dim, nclasses, num_nodes
        dev = torch.device(device)
        self.device = dev
        ent = [("conv1.weight", (16, 9, 3, 1), _conv2d_pt), ("bn1.weight", (16,), _ident), ("bn1.bias", (16,), _ident)]
        cfg, inp = [], 16
        for li, (planes, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            for bi in range(2):
                cin = inp if bi == 0 else planes
                p = "layer%d.%d" % (li, bi)
                cfg.append((p, cin, planes, stride if bi == 0 else 1, bi == 0))
                ent += [(p + ".bn1.weight", (cin,), _ident), (p + ".bn1.bias", (cin,), _ident),
                        (p + ".conv1.weight", (planes, 3, 3, cin), _conv2d_pt),
                        (p + ".bn2.weight", (planes,), _ident), (p + ".bn2.bias", (planes,), _ident),
                        (p + ".conv2.weight", (planes, 3, 3, planes), _conv2d_pt)]
                if bi == 0:
                    ent.append((p + ".shortcut.0.weight", (planes, 1, 1, cin), _conv2d_pt))
            inp = planes
        ent += [("conv5.weight", (256, 3, 3, 512), _conv2d_pt), ("bn5.weight", (256,), _ident), ("bn5.bias", (256,), _ident),
                ("fc.weight", (enc_dim, 512), _ident), ("fc.bias", (enc_dim,), _ident),
                ("fc_mu.weight", (nclasses, enc_dim), _ident), ("fc_mu.bias", (nclasses,), _ident),
                ("attention.att_weights", (1, 256), _ident)]
        frozen = () if train_head_mu else ("fc_mu.weight", "fc_mu.bias")
        self.store = ParamStore(ent, dev, frozen=frozen)
        self.buffers = BufferStore(dev)
        st, bufs = self.store, self.buffers

        # ---- layers ----
        if precision == "fp32":
            self.scratch = Scratch(dev)
            self.overlap_wgrad = False              # the split operands of fprop / dgrad / wgrad share scratch buffers

            def ConvLayer(*a, **k):                 # noqa: N802 (shadows the module-level class inside this constructor)
                return SplitConvLayer(*a, scratch=self.scratch, **k)
        else:
            ConvLayer = globals()["ConvLayer"]
        self.bn1 = BNLayer(st, bufs, "bn1", 16)
        self.blocks = []
        for p, cin, planes, stride, has_sc in cfg:
            blk = _Block()
            blk.name, blk.cin, blk.planes, blk.stride = p, cin, planes, stride
            blk.bn1 = BNLayer(st, bufs, p + ".bn1", cin)
            blk.conv1 = ConvLayer(st, p + ".conv1", cin, planes, 3, 3, stride, stride, 1, 1)
            blk.bn2 = BNLayer(st, bufs, p + ".bn2", planes)
            blk.conv2 = ConvLayer(st, p + ".conv2", planes, planes, 3, 3, 1, 1, 1, 1)
            blk.sc = ConvLayer(st, p + ".shortcut.0", cin, planes, 1, 1, stride, stride, 0, 0) if has_sc else None
            self.blocks.append(blk)
        self.conv5 = ConvLayer(st, "conv5", 512, 256, num_nodes, 3, 1, 1, 0, 1)
        self.bn5 = BNLayer(st, bufs, "bn5", 256)
        bufs.finalize()
        for obj in [self.bn1, self.bn5] + [b.bn1 for b in self.blocks] + [b.bn2 for b in self.blocks]:
            _resolve_lazies(obj)
        self.noise_seed = -1
        self._packed_version = -1
        self.init_parameters()

    # batch statistics of a BatchNorm accumulated by the epilogue of the patch conv that produces its input
    fuse_bn_stats = os.environ.get("AIR_FUSE_BN_STATS", "1") != "0"

    def bind(self, batch, T):
        """(Re)allocate activation / gradient buffers for a (batch, T) input."""
        if self.B == batch and self.T == T:
            return
        self.B, self.T = batch, T
        B, dev, F, enc_dim, nclasses = batch, self.device, self.F, self.enc_dim, self.nclasses
        self.H0, self.W0 = F, T
        self.H1, self.W1 = ops.conv_out_size(F, 9, 3, 1, 1), ops.conv_out_size(T, 3, 1, 1, 1)
        H, W = self.H1, self.W1
        for blk in self.blocks:
            blk.H, blk.W = H, W
            blk.Ho, blk.Wo = blk.conv1.out_hw(H, W)
            H, W = blk.Ho, blk.Wo
        self.H5, self.W5 = self.conv5.out_hw(H, W)
        assert self.H5 == 1, "conv5 must collapse the frequency axis (num_nodes == 3 for 60-dim LFCC)"

        def act(*shape):
            return torch.empty(shape, device=dev, dtype=self.act_dtype)
        self.c1 = act(B, self.H1, self.W1, 16)
        self.z1 = act(B, self.H1, self.W1, 16)
        self.g_z1 = act(B, self.H1, self.W1, 16)
        self.g_c1 = act(B, self.H1, self.W1, 16)
        for blk in self.blocks:
            blk.a1 = act(B, blk.H, blk.W, blk.cin)
            blk.h = act(B, blk.Ho, blk.Wo, blk.planes)
            blk.a2 = act(B, blk.Ho, blk.Wo, blk.planes)
            blk.y = act(B, blk.Ho, blk.Wo, blk.planes)
            blk.g_y = act(B, blk.Ho, blk.Wo, blk.planes)
            blk.g_a2 = act(B, blk.Ho, blk.Wo, blk.planes)
            blk.g_h = act(B, blk.Ho, blk.Wo, blk.planes)
            blk.g_a1 = act(B, blk.H, blk.W, blk.cin)
        self.c5 = act(B, self.W5, 256)
        self.z5 = act(B, self.W5, 256)
        self.g_z5 = act(B, self.W5, 256)
        self.g_c5 = act(B, self.W5, 256)
        f32 = dict(device=dev, dtype=torch.float32)
        self.stats = torch.empty(B, 512, **f32)
        self.g_stats = torch.empty(B, 512, **f32)
        self.pool_p = torch.empty(B, self.W5, **f32)
        self.pool_th = torch.empty(B, self.W5, **f32)
        self.feat = torch.empty(B, enc_dim, **f32)
        self.mu = torch.empty(B, nclasses, **f32)

    # ---- parameters ---------------------------------------------------------------------
    def init_parameters(self, seed=None):
        """kaiming-normal(fan_out) convs, kaiming-uniform linears, BN weight 1 / bias 0
        (resnet.py:149-157); attention weights kaiming-uniform (resnet.py:21)."""
        g = torch.Generator(device="cpu")
        if seed is not None:
            g.manual_seed(seed)
        else:
            g.manual_seed(torch.initial_seed() % (2 ** 31))
        for name, (off, n, shape) in self.store.offsets.items():
            v = self.store.view(name)
            if name.endswith("conv1.weight") or name.endswith("conv2.weight") or name.endswith("shortcut.0.weight") \
                    or name.endswith("conv5.weight"):
                fan_out = shape[0] * shape[1] * shape[2]
                v.copy_(torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out))
            elif name in ("fc.weight", "fc_mu.weight", "attention.att_weights"):
                bound = math.sqrt(6.0 / shape[1])
                v.copy_((torch.rand(shape, generator=g) * 2 - 1) * bound)
            elif name in ("fc.bias", "fc_mu.bias"):
                fan_in = 512 if name == "fc.bias" else self.enc_dim
                bound = 1.0 / math.sqrt(fan_in)
                v.copy_((torch.rand(shape, generator=g) * 2 - 1) * bound)
            elif name.endswith(".weight"):
                v.fill_(1.0)
            else:
                v.zero_()
        self.mark_dirty()

    def mark_dirty(self):
        self._packed_version = -1

    def load_state(self, sd):
        """Copy a reference-layout state dict (resnet.py key names) into the flat engine buffers."""
        bufs = self.buffers.named_f32()
        with torch.no_grad():
            for k, t in sd.items():
                t = t.detach()
                if k in self.store.offsets:
                    self.store.pt_view(k).copy_(t.to(self.device))
                elif k in bufs:
                    bufs[k].copy_(t.to(self.device))
                elif k.endswith("num_batches_tracked"):
                    dict(self.bns())[k[:-len(".num_batches_tracked")]].num_batches_tracked.fill_(int(t))
                else:
                    raise KeyError(k)
        self.mark_dirty()

    def state(self):
        """Reference-layout state dict (detached copies on the engine's device)."""
        out = {k: self.store.pt_view(k).detach().clone().contiguous() for k in self.store.names()}
        out.update({k: v.clone() for k, v in self.buffers.named_f32().items()})
        for name, bn in self.bns():
            out[name + ".num_batches_tracked"] = bn.num_batches_tracked.clone()
        return out

    def convs(self):
        out = []
        for blk in self.blocks:
            out += [blk.conv1, blk.conv2] + ([blk.sc] if blk.sc else [])
        return out + [self.conv5]

    def bns(self):
        out = [("bn1", self.bn1)]
        for blk in self.blocks:
            out += [(blk.name + ".bn1", blk.bn1), (blk.name + ".bn2", blk.bn2)]
        return out + [("bn5", self.bn5)]

    def pack_weights(self):
        if getattr(self, "_pack_plan", None) is None:
            self._pack_plan = ops.PackPlan(self.device)
            self._pack_rest = [c for c in self.convs() if not c.add_pack_jobs(self._pack_plan)]
        self._pack_plan.run()
        for c in self._pack_rest:
            c.pack()
        self._packed_version = self.store.step

    # ---- forward --------------------------------------------------------------------------
    def forward(self, x0, training=True):
        """x0: (B, 60, T) bf16 (the fused LFCC output, layout 'resnet').  Returns (feat, mu) fp32."""
        assert x0.dtype == self.act_dtype and x0.is_contiguous() and x0.dim() == 3 and x0.shape[1] == self.F
        self.bind(x0.shape[0], x0.shape[2])
        B = self.B
        self.prepack()                          # (normally started by the Trainer before the LFCC kernel)
        self.x0 = x0
        if training:
            self.buffers.f64.zero_()
        st = self.store
        ops.stem_fwd(x0, B, self.H0, self.W0, 9, 3, 3, 1, 1, 1, st.view("conv1.weight"), 16, self.c1)
        M1 = B * self.H1 * self.W1
        self.bn1.forward(self.c1, 16, self.z1, 16, M1, True, training)
        self._await_pack()                      # the first layer that reads packed weights follows
        x = self.z1
        x_stats = False                         # were the statistics of x accumulated by the conv that produced it?
        for bi, blk in enumerate(self.blocks):
            Min, Mout = B * blk.H * blk.W, B * blk.Ho * blk.Wo
            blk.x = x
            blk.bn1.forward(x, blk.cin, blk.a1, blk.cin, Min, True, training, have_stats=x_stats)
            if blk.sc is not None:
                blk.sc.fprop(blk.a1, blk.cin, B, blk.H, blk.W, blk.y, blk.planes)
            fuse = training and self.fuse_bn_stats
            blk.conv1.fprop(blk.a1, blk.cin, B, blk.H, blk.W, blk.h, blk.planes, stats=blk.bn2.sums if fuse else None)
            blk.bn2.forward(blk.h, blk.planes, blk.a2, blk.planes, Mout, True, training, have_stats=blk.conv1.stats_fused)
            res = blk.y if blk.sc is not None else x
            nxt = self.blocks[bi + 1].bn1.sums if (fuse and bi + 1 < len(self.blocks)) else None
            blk.conv2.fprop(blk.a2, blk.planes, B, blk.Ho, blk.Wo, blk.y, blk.planes, res=res, res_ld=blk.planes, stats=nxt)
            x_stats = blk.conv2.stats_fused
            x = blk.y
        last = self.blocks[-1]
        self.conv5.fprop(x, 512, B, last.Ho, last.Wo, self.c5, 256, stats=self.bn5.sums if (training and self.fuse_bn_stats) else None)
        M5 = B * self.W5
        self.bn5.forward(self.c5, 256, self.z5, 256, M5, True, training, have_stats=self.conv5.stats_fused)
        ops.selfattn_pool_fwd(self.z5, st.view("attention.att_weights"), self.stats, self.pool_p, self.pool_th,
                              B, self.W5, 256, self.noise_seed)
        ops.linear_fwd(self.stats, st.view("fc.weight"), st.view("fc.bias"), self.feat, B, self.enc_dim, 512)
        ops.linear_fwd(self.feat, st.view("fc_mu.weight"), st.view("fc_mu.bias"), self.mu, B, self.nclasses, self.enc_dim)
        return self.feat, self.mu

    # ---- backward -------------------------------------------------------------------------
    def zero_grad(self):
        self.store.grads.zero_()

    grad_hook = None        # optional callable(offset, events): every gradient at flat index >= offset is final

    def ready_names(self):
        """Parameters at which backward() reports progress to the gradient reducer (bucket boundaries)."""
        return ["conv5.weight"] + [blk.name + ".bn1.weight" for blk in self.blocks]

    def _ready(self, name):
        """Every gradient at or above `name` in the flat buffer has been LAUNCHED: hand the reducer the offset and an
        event of the weight-gradient side stream -- the exchange stream waits for it, the compute stream does not."""
        if self.grad_hook is not None:
            self.grad_hook(self.store.offsets[name][0], self._side_events())

    def backward(self, dfeat, dmu=None):
        """Accumulates parameter gradients into store.grads (call zero_grad() first).
        dfeat (B, enc_dim) fp32 [, dmu (B, nclasses) fp32]."""
        B, st = self.B, self.store
        if dmu is not None:
            dfe = torch.empty_like(self.feat)
            ops.linear_bwd(self.feat, st.view("fc_mu.weight"), dmu, dfe, st.grad("fc_mu.weight"), st.grad("fc_mu.bias"),
                           B, self.nclasses, self.enc_dim)
            dfeat = dfeat + dfe if dfeat is not None else dfe
        dfeat = dfeat.contiguous()
        ops.linear_bwd(self.stats, st.view("fc.weight"), dfeat, self.g_stats, st.grad("fc.weight"), st.grad("fc.bias"),
                       B, self.enc_dim, 512)
        ops.selfattn_pool_bwd(self.z5, st.view("attention.att_weights"), self.pool_p, self.pool_th, self.stats,
                              self.g_stats, self.g_z5, st.grad("attention.att_weights"), B, self.W5, 256, self.noise_seed)
        M5 = B * self.W5
        self.bn5.backward(self.g_z5, 256, self.c5, 256, self.g_c5, 256, M5, 0)
        last = self.blocks[-1]
        self._wgrad_async(self.conv5.wgrad, last.y, 512, B, last.Ho, last.Wo, self.g_c5, 256)
        self._ready("conv5.weight")
        self.conv5.dgrad(self.g_c5, 256, B, last.Ho, last.Wo, last.g_y, 512)
        for i in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[i]
            Min, Mout = B * blk.H * blk.W, B * blk.Ho * blk.Wo
            g_x = self.blocks[i - 1].g_y if i > 0 else self.g_z1
            self._wgrad_async(blk.conv2.wgrad, blk.a2, blk.planes, B, blk.Ho, blk.Wo, blk.g_y, blk.planes)
            blk.conv2.dgrad(blk.g_y, blk.planes, B, blk.Ho, blk.Wo, blk.g_a2, blk.planes)
            blk.bn2.backward(blk.g_a2, blk.planes, blk.h, blk.planes, blk.g_h, blk.planes, Mout, 0)
            self._wgrad_async(blk.conv1.wgrad, blk.a1, blk.cin, B, blk.H, blk.W, blk.g_h, blk.planes)
            blk.conv1.dgrad(blk.g_h, blk.planes, B, blk.H, blk.W, blk.g_a1, blk.cin)
            if blk.sc is not None:
                self._wgrad_async(blk.sc.wgrad, blk.a1, blk.cin, B, blk.H, blk.W, blk.g_y, blk.planes)
                blk.sc.dgrad(blk.g_y, blk.planes, B, blk.H, blk.W, blk.g_a1, blk.cin, accumulate=True)
                blk.bn1.backward(blk.g_a1, blk.cin, blk.x, blk.cin, g_x, blk.cin, Min, 0)
            else:
                blk.bn1.backward(blk.g_a1, blk.cin, blk.x, blk.cin, g_x, blk.cin, Min, 0, add=blk.g_y, add_ld=blk.planes)
            self._ready(blk.name + ".bn1.weight")
        M1 = B * self.H1 * self.W1
        self.bn1.backward(self.g_z1, 16, self.c1, 16, self.g_c1, 16, M1, 0)
        self._wgrad_async(ops.stem_wgrad, self.x0, B, self.H0, self.W0, 9, 3, 3, 1, 1, 1, self.g_c1, 16, st.grad("conv1.weight"))
        self._join_side()
