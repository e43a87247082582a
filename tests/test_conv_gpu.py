"""GPU parity of the tcgen05 implicit-GEMM convolution kernels (fprop / dgrad / wgrad, through the
C ABI) against a plain PyTorch fp32 convolution of the same bf16-rounded operands.

Shapes are the ResNet-18 / ECAPA layer geometries of SURVEY.md section 8(a) at small batch.
Tolerance: operands are exactly representable in bf16, products accumulate in fp32 in TMEM, the
output is rounded once to bf16 -> |err| <= 2^-8 * |ref| + small fp32 accumulation slack."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = {
    # name: B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw
    "gemm_1x1_64_64": (1, 1, 256, 64, 64, 1, 1, 1, 1, 0, 0, 1, 1),
    "gemm_1x1_128_256": (2, 1, 300, 128, 256, 1, 1, 1, 1, 0, 0, 1, 1),
    "res_l1_3x3_64_64": (2, 18, 75, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1),
    "res_l1_3x3_16_64": (2, 18, 75, 16, 64, 3, 3, 1, 1, 1, 1, 1, 1),
    "res_l1_sc_16_64": (2, 18, 75, 16, 64, 1, 1, 1, 1, 0, 0, 1, 1),
    "res_l2_3x3_s2_64_128": (2, 18, 75, 64, 128, 3, 3, 2, 2, 1, 1, 1, 1),
    "res_l2_sc_s2_64_128": (2, 18, 75, 64, 128, 1, 1, 2, 2, 0, 0, 1, 1),
    "res_l3_3x3_s2_128_256": (2, 9, 38, 128, 256, 3, 3, 2, 2, 1, 1, 1, 1),
    "res_l4_3x3_512_512": (2, 3, 94, 512, 512, 3, 3, 1, 1, 1, 1, 1, 1),
    "res_conv5_512_256": (2, 3, 94, 512, 256, 3, 3, 1, 1, 0, 1, 1, 1),
    "ecapa_k3_dil2_64_64": (3, 1, 750, 64, 64, 1, 3, 1, 1, 0, 2, 1, 2),
    "ecapa_k3_dil4_64_64": (3, 1, 750, 64, 64, 1, 3, 1, 1, 0, 4, 1, 4),
    "ecapa_k5_64_512": (2, 1, 750, 64, 512, 1, 5, 1, 1, 0, 2, 1, 1),
    "ecapa_k1_1536_1536": (1, 1, 750, 1536, 1536, 1, 1, 1, 1, 0, 0, 1, 1),
    "ecapa_k1_512_128": (2, 1, 750, 512, 128, 1, 1, 1, 1, 0, 0, 1, 1),
}


def _setup(name):
    from asvspoof2021_air_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES[name]
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).cuda()
    Ho, Wo = ops.conv_out_size(H, kh, sh, ph, dh), ops.conv_out_size(W, kw, sw, pw, dw)
    dy = torch.randn(B, Cout, Ho, Wo, generator=g).cuda().to(torch.bfloat16)
    return ops, x, w, dy, Ho, Wo


def _check(got, ref, what):
    assert not torch.isnan(got).any(), what
    err = (got - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-3 * ref.abs().max()
    bad = err > tol
    assert not bad.any(), "%s: %d bad, max err %g (ref max %g)" % (what, int(bad.sum()), float(err.max()), float(ref.abs().max()))


@pytest.mark.parametrize("name", list(CASES))
def test_fprop(name):
    ops, x, w, dy, Ho, Wo = _setup(name)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES[name]
    wq = w.to(torch.bfloat16).float()
    ref = F.conv2d(x.float(), wq, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))
    xn = x.permute(0, 2, 3, 1).contiguous()
    wpk = ops.pack_weights(w.permute(0, 2, 3, 1).contiguous(), 0, Cin, Cout, kh * kw)
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(xn, Cin, B, H, W, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, 0, wpk, Cout, kh * kw * Cin, out, Cout)
    torch.cuda.synchronize()
    _check(out.float().permute(0, 3, 1, 2), ref, name + " fprop")


@pytest.mark.parametrize("name", list(CASES))
def test_dgrad(name):
    ops, x, w, dy, Ho, Wo = _setup(name)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES[name]
    wq = w.to(torch.bfloat16).float()
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), wq, dy.float(), stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    wpk = ops.pack_weights(w.permute(0, 2, 3, 1).contiguous(), 1, Cin, Cout, kh * kw)
    out = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(dyn, Cout, B, Ho, Wo, Cout, H, W, kh, kw, sh, sw, ph, pw, dh, dw, 1, wpk, Cin, kh * kw * Cout, out, Cin)
    torch.cuda.synchronize()
    _check(out.float().permute(0, 3, 1, 2), ref, name + " dgrad")


@pytest.mark.parametrize("name", list(CASES))
def test_wgrad(name):
    ops, x, w, dy, Ho, Wo = _setup(name)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES[name]
    ref = torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, kh, kw), dy.float(), stride=(sh, sw), padding=(ph, pw),
                                      dilation=(dh, dw))
    xn = x.permute(0, 2, 3, 1).contiguous()
    dyn = dy.permute(0, 2, 3, 1).contiguous()
    dwb = torch.zeros(Cout, kh * kw * Cin, device="cuda")
    ops.conv_wgrad(xn, Cin, B, H, W, Cin, dyn, Cout, Ho, Wo, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dwb)
    torch.cuda.synchronize()
    got = dwb.view(Cout, kh, kw, Cin).permute(0, 3, 1, 2)
    err = (got - ref).abs()
    assert float(err.max()) <= 1e-3 * float(ref.abs().max()) + 1e-3, (name, float(err.max()), float(ref.abs().max()))


def test_fused_epilogue_bias_residual_relu():
    ops, x, w, dy, Ho, Wo = _setup("res_l1_3x3_64_64")
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES["res_l1_3x3_64_64"]
    g = torch.Generator(device="cpu").manual_seed(2)
    bias = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(B, Ho, Wo, Cout, generator=g).cuda().to(torch.bfloat16)
    wq = w.to(torch.bfloat16).float()
    ref = F.relu(F.conv2d(x.float(), wq, bias, stride=1, padding=1) + res.float().permute(0, 3, 1, 2))
    xn = x.permute(0, 2, 3, 1).contiguous()
    wpk = ops.pack_weights(w.permute(0, 2, 3, 1).contiguous(), 0, Cin, Cout, 9)
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(xn, Cin, B, H, W, Cin, Ho, Wo, 3, 3, 1, 1, 1, 1, 1, 1, 0, wpk, Cout, 9 * Cin, out, Cout,
                  bias=bias, res=res, res_ld=Cout, relu=True)
    torch.cuda.synchronize()
    _check(out.float().permute(0, 3, 1, 2), ref, "fused epilogue")


def test_channel_slices_with_leading_dimension():
    """ECAPA Res2 branches read / write 64-channel slices of 512-channel rows (ld = 512)."""
    from asvspoof2021_air_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(3)
    B, T, C, width = 2, 750, 512, 64
    big = torch.randn(B, T, C, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(width, width, 1, 3, generator=g) / (3 * width) ** 0.5).cuda()
    out = torch.zeros(B, T, C, device="cuda", dtype=torch.bfloat16)
    wpk = ops.pack_weights(w.permute(0, 2, 3, 1).contiguous(), 0, width, width, 3)
    src = big[:, :, 128:192]
    ops.conv_gemm(src, C, B, 1, T, width, 1, T, 1, 3, 1, 1, 0, 3, 1, 3, 0, wpk, width, 3 * width, out[:, :, 64:128], C)
    torch.cuda.synchronize()
    ref = F.conv1d(src.float().permute(0, 2, 1), w.to(torch.bfloat16).float()[:, :, 0], padding=3, dilation=3)
    _check(out[:, :, 64:128].float().permute(0, 2, 1), ref, "slice conv")
    assert (out[:, :, :64] == 0).all() and (out[:, :, 128:] == 0).all()


PATCH_CASES = {
    # name: B, H, W, Cin, Cout
    "l1_64_64": (2, 18, 750, 64, 64),
    "l1_16_64": (2, 18, 750, 16, 64),
    "l2_128_128": (2, 9, 375, 128, 128),
    "odd_rows_ragged_width": (3, 5, 131, 64, 128),
    "single_row_narrow": (2, 1, 40, 32, 16),
    "wide_n256": (1, 4, 260, 64, 256),
}


@pytest.mark.parametrize("name", list(PATCH_CASES))
def test_patch_conv3x3_fprop_and_dgrad(name):
    """csrc/conv_patch.cu (shared-memory resident patch, shifted descriptor windows) vs torch fp32 conv."""
    from asvspoof2021_air_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, Cin, Cout = PATCH_CASES[name]
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).cuda()
    dy = torch.randn(B, Cout, H, W, generator=g).cuda().to(torch.bfloat16)
    res = torch.randn(B, H, W, Cout, generator=g).cuda().to(torch.bfloat16)
    wq = w.to(torch.bfloat16).float()
    wg = w.permute(0, 2, 3, 1).contiguous()
    assert ops.patch_supported(Cin, Cout, H, W) and ops.patch_supported(Cout, Cin, H, W)
    wpk = torch.empty(9 * Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack3x3(wg, Cin, Cout, 0, wpk)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv3x3_patch(x.permute(0, 2, 3, 1).contiguous(), Cin, B, H, W, Cin, wpk, Cout, out, Cout, res, Cout, True)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float(), wq, padding=1) + res.float().permute(0, 3, 1, 2))
    _check(out.float().permute(0, 3, 1, 2), ref, name + " patch fprop")
    wpk_d = torch.empty(9 * Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack3x3(wg, Cout, Cin, 1, wpk_d)
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv3x3_patch(dy.permute(0, 2, 3, 1).contiguous(), Cout, B, H, W, Cout, wpk_d, Cin, dx, Cin, None, 0, False, 1)
    torch.cuda.synchronize()
    refd = torch.nn.grad.conv2d_input((B, Cin, H, W), wq, dy.float(), padding=1)
    _check(dx.float().permute(0, 3, 1, 2), refd, name + " patch dgrad")


WPATCH_CASES = {
    # name: B, H, W, Cin, Cout[, k]
    "l1_16_64_narrow": (2, 18, 750, 16, 64),
    "narrow_ragged": (3, 5, 131, 16, 128),
    "l1_shortcut_1x1_narrow": (2, 18, 750, 16, 64, 1),
    "wide_1x1": (2, 7, 200, 64, 64, 1),
    "l1_64_64": (2, 18, 750, 64, 64),
    "l2_128_128": (2, 9, 375, 128, 128),
    "odd_rows_ragged_width": (3, 5, 131, 64, 128),
    "single_row_narrow": (2, 1, 40, 128, 64),
    "l3_256_256": (1, 5, 188, 256, 256),
}


@pytest.mark.parametrize("name", list(WPATCH_CASES))
def test_patch_conv3x3_wgrad(name):
    """csrc/conv_wgrad_patch.cu (TMA patches, MN-major shifted windows, tap pairs through LBO) vs torch fp32 wgrad."""
    from asvspoof2021_air_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, Cin, Cout = WPATCH_CASES[name][:5]
    k = WPATCH_CASES[name][5] if len(WPATCH_CASES[name]) > 5 else 3
    g = torch.Generator(device="cpu").manual_seed(13)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().to(torch.bfloat16)
    dy = torch.randn(B, Cout, H, W, generator=g).cuda().to(torch.bfloat16)
    assert ops.wgrad_patch_supported(Cin, Cout)
    dw = torch.zeros(Cout, k * k * Cin, device="cuda")
    xc, dyc = x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous()
    ops.conv_wgrad_patch(xc, Cin, B, H, W, Cin, dyc, Cout, Cout, k, dw)
    ops.conv_wgrad_patch(xc, Cin, B, H, W, Cin, dyc, Cout, Cout, k, dw)          # accumulates
    torch.cuda.synchronize()
    ref = 2 * torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, k, k), dy.float(), padding=k // 2)
    got = dw.view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    err = (got - ref).norm() / ref.norm()
    worst = (got - ref).abs().max() / ref.abs().max()
    assert err < 1e-3 and worst < 2e-3, "%s: normwise %.3g, worst %.3g" % (name, float(err), float(worst))


S2_CASES = {
    # name: B, H, W, Cin, Cout, k
    "l2_3x3": (2, 18, 750, 64, 128, 3),
    "l3_3x3_odd": (2, 9, 375, 128, 256, 3),
    "l4_3x3_odd_h": (2, 5, 188, 256, 256, 3),
    "tiny_3x3": (1, 3, 5, 64, 64, 3),
    "l2_1x1": (2, 18, 750, 64, 128, 1),
    "l3_1x1_odd": (2, 9, 375, 128, 256, 1),
}


@pytest.mark.parametrize("name", list(S2_CASES))
def test_patch_stride2_dgrad_by_parity(name):
    """air_conv_s2_dgrad_patch_bf16: output-parity decomposition of the stride-2 data gradient vs torch fp32."""
    from asvspoof2021_air_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    B, H, W, Cin, Cout, k = S2_CASES[name]
    pad = 1 if k == 3 else 0
    g = torch.Generator(device="cpu").manual_seed(17)
    Ho, Wo = ops.conv_out_size(H, k, 2, pad, 1), ops.conv_out_size(W, k, 2, pad, 1)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda()
    dy = torch.randn(B, Cout, Ho, Wo, generator=g).cuda().to(torch.bfloat16)
    prev = torch.randn(B, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
    wq = w.to(torch.bfloat16).float()
    wpk_d = torch.empty(k * k * Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack_patch(w.permute(0, 2, 3, 1).contiguous(), Cout, Cin, k * k, 1, wpk_d)
    dx = prev.clone()
    ops.conv_s2_dgrad_patch(dy.permute(0, 2, 3, 1).contiguous(), Cout, B, Ho, Wo, Cout, wpk_d, k, Cin, dx, Cin, H, W, dx, Cin)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_input((B, Cin, H, W), wq, dy.float(), stride=2, padding=pad) + prev.float().permute(0, 3, 1, 2)
    _check(dx.float().permute(0, 3, 1, 2), ref, name + " s2 dgrad (accumulate)")
    if k == 3:
        dx2 = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv_s2_dgrad_patch(dy.permute(0, 2, 3, 1).contiguous(), Cout, B, Ho, Wo, Cout, wpk_d, k, Cin, dx2, Cin, H, W)
        torch.cuda.synchronize()
        _check(dx2.float().permute(0, 3, 1, 2), ref - prev.float().permute(0, 3, 1, 2), name + " s2 dgrad")


@pytest.mark.parametrize("shape", [(2, 18, 750, 16, 64), (2, 7, 130, 64, 16), (1, 3, 94, 128, 256)])
def test_patch_conv1x1(shape):
    """single-tap use of the patch kernel (1x1 / stride 1 = GEMM over pixels) vs torch fp32."""
    from asvspoof2021_air_b200 import ops
    B, H, W, Cin, Cout = shape
    g = torch.Generator(device="cpu").manual_seed(19)
    x = torch.randn(B, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).cuda().to(torch.bfloat16)
    wpk = torch.empty(Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack_patch(w.contiguous(), Cin, Cout, 1, 0, wpk)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv1x1_patch(x, Cin, B, H, W, Cin, wpk, Cout, out, Cout, res, Cout, False)
    torch.cuda.synchronize()
    ref = x.float() @ w.to(torch.bfloat16).float().t() + res.float()
    _check(out.float(), ref, "1x1 patch fprop")
    wpk_d = torch.empty(Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack_patch(w.contiguous(), Cout, Cin, 1, 1, wpk_d)
    dy = torch.randn(B, H, W, Cout, generator=g).cuda().to(torch.bfloat16)
    dx = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv1x1_patch(dy, Cout, B, H, W, Cout, wpk_d, Cin, dx, Cin, None, 0, False, 1)
    torch.cuda.synchronize()
    _check(dx.float(), dy.float() @ w.to(torch.bfloat16).float(), "1x1 patch dgrad")


def test_batched_pack_plan_matches_per_tensor_packing():
    """csrc/pack.cu: one launch over a job table == the per-tensor packing kernels, bit for bit."""
    from asvspoof2021_air_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(23)
    plan = ops.PackPlan(torch.device("cuda"))
    checks = []
    for cin, cout, taps in ((64, 64, 9), (16, 64, 9), (128, 256, 9), (64, 128, 1), (512, 256, 9)):
        w = torch.randn(cout, taps, cin, generator=g).cuda().reshape(-1)
        for mode in (0, 1):
            n, k = (cout, taps * cin) if mode == 0 else (cin, taps * cout)
            ref = ops.pack_weights(w, mode, cin, cout, taps)
            out = torch.zeros_like(ref)
            plan.add_gemm(w, taps * cin, mode, cin, cout, taps, out)
            checks.append((ref, out))
            C, N = (cin, cout) if mode == 0 else (cout, cin)
            if ops.patch_supported(C, N, 1, 1):
                ref3 = torch.zeros(taps * cin * cout, device="cuda", dtype=torch.bfloat16)
                ops.pack_patch(w, C, N, taps, mode, ref3)
                out3 = torch.zeros_like(ref3)
                plan.add_patch(w, C, N, taps, mode, out3)
                checks.append((ref3, out3))
    plan.run()
    torch.cuda.synchronize()
    for ref, out in checks:
        assert torch.equal(ref, out)


@pytest.mark.parametrize("d", [2, 3, 4])
def test_patch_dilated_conv1d_fprop_dgrad_wgrad(d):
    """The Res2-branch Conv1d(64, 64, k=3, dilation=d, padding=d) of ecapa_tdnn.py:50 on the TMA patch kernels
    (136-pixel patch, taps at 0 / d / 2d), on 64-channel slices of 512-wide tensors, vs torch fp32."""
    from asvspoof2021_air_b200 import ops
    B, T, C, Wd = 3, 750, 512, 64
    g = torch.Generator(device="cpu").manual_seed(29 + d)
    big = torch.randn(B, T, C, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Wd, Wd, 3, generator=g) / (3 * Wd) ** 0.5).cuda()            # (Cout, Cin, k)
    bias = torch.randn(Wd, generator=g).cuda()
    wq = w.to(torch.bfloat16).float()
    wg = w.permute(0, 2, 1).contiguous().reshape(-1)                              # GEMM layout [Cout][k][Cin]
    wpk = torch.empty(3 * Wd * Wd, device="cuda", dtype=torch.bfloat16)
    wpk_d = torch.empty_like(wpk)
    ops.pack_patch(wg, Wd, Wd, 3, 0, wpk)
    ops.pack_patch(wg, Wd, Wd, 3, 1, wpk_d)
    src = big[:, :, 128:192]
    out = torch.full((B, T, Wd), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv1d_patch(src, C, B, 1, T, Wd, wpk, 3, d, Wd, out, Wd, bias, None, 0, True)
    torch.cuda.synchronize()
    xs = src.float().permute(0, 2, 1)
    ref = F.relu(F.conv1d(xs, wq, bias, padding=d, dilation=d))
    _check(out.float().permute(0, 2, 1), ref, "dilated conv1d fprop d=%d" % d)
    # training form: the epilogue also accumulates the batch statistics of the BatchNorm that follows (sum / sum of squares of
    # the stored output)
    st = torch.zeros(2 * Wd, device="cuda", dtype=torch.float64)
    out_s = torch.full((B, T, Wd), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv1d_patch(src, C, B, 1, T, Wd, wpk, 3, d, Wd, out_s, Wd, bias, None, 0, True, None, 0, 0, None, None, st)
    torch.cuda.synchronize()
    assert torch.equal(out_s, out)
    o64 = out.double().reshape(-1, Wd)
    assert torch.allclose(st[:Wd], o64.sum(0), rtol=1e-5, atol=1e-6 * float(o64.abs().sum(0).max())), (st[:Wd] - o64.sum(0)).abs().max()
    assert torch.allclose(st[Wd:], (o64 * o64).sum(0), rtol=1e-5)
    # scoring form: conv -> ReLU -> eval-mode BatchNorm affine in the epilogue; second launch also emits the next branch's
    # input = round_bf16(affine output) + next split (ecapa_tdnn.py:73-83)
    sc, sh = (torch.rand(Wd, generator=g) + 0.5).cuda(), torch.randn(Wd, generator=g).cuda()
    refa = ref * sc[None, :, None] + sh[None, :, None]
    oa = torch.full((B, T, Wd), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv1d_patch(src, C, B, 1, T, Wd, wpk, 3, d, Wd, oa, Wd, bias, None, 0, True, None, 0, 0, sc, sh)
    torch.cuda.synchronize()
    _check(oa.float().permute(0, 2, 1), refa, "dilated conv1d fprop + affine d=%d" % d)
    cat = torch.zeros(B, T, C, device="cuda", dtype=torch.bfloat16)
    nxt = torch.full((B, T, Wd), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.conv1d_patch(src, C, B, 1, T, Wd, wpk, 3, d, Wd, nxt, Wd, bias, big[:, :, 192:256], C, True, cat[:, :, 64:128], C, 0, sc, sh)
    torch.cuda.synchronize()
    assert torch.equal(cat[:, :, 64:128], oa) and (cat[:, :, :64] == 0).all() and (cat[:, :, 128:] == 0).all()
    want = (oa.float() + big[:, :, 192:256].float()).to(torch.bfloat16)
    assert torch.equal(nxt, want)
    # data gradient with a residual and the un-accumulated second output
    dy = torch.randn(B, T, Wd, generator=g).cuda().to(torch.bfloat16)
    res = torch.randn(B, T, C, generator=g).cuda().to(torch.bfloat16)
    dx = torch.full((B, T, Wd), float("nan"), device="cuda", dtype=torch.bfloat16)
    dx2 = torch.zeros(B, T, C, device="cuda", dtype=torch.bfloat16)
    ops.conv1d_patch(dy, Wd, B, 1, T, Wd, wpk_d, 3, d, Wd, dx, Wd, None, res[:, :, 64:128], C, False, dx2[:, :, 192:256], C, 1)
    torch.cuda.synchronize()
    refd = torch.nn.grad.conv1d_input((B, Wd, T), wq, dy.float().permute(0, 2, 1), padding=d, dilation=d)
    _check(dx2[:, :, 192:256].float().permute(0, 2, 1), refd, "dilated conv1d dgrad (out2) d=%d" % d)
    _check(dx.float().permute(0, 2, 1), refd + res[:, :, 64:128].float().permute(0, 2, 1), "dilated conv1d dgrad (+res) d=%d" % d)
    assert (dx2[:, :, :192] == 0).all() and (dx2[:, :, 256:] == 0).all()
    # weight gradient
    dw = torch.zeros(Wd, 3 * Wd, device="cuda")
    ops.conv1d_wgrad_patch(src, C, B, 1, T, Wd, dy, Wd, Wd, 3, d, dw)
    torch.cuda.synchronize()
    refw = torch.nn.grad.conv1d_weight(xs, (Wd, Wd, 3), dy.float().permute(0, 2, 1), padding=d, dilation=d)
    got = dw.view(Wd, 3, Wd).permute(0, 2, 1)
    err = (got - refw).norm() / refw.norm()
    assert err < 1e-3, "dilated conv1d wgrad d=%d: normwise %.3g" % (d, float(err))


@pytest.mark.parametrize("shape", [(2, 18, 750, 64, 64), (3, 5, 131, 64, 128), (2, 9, 375, 128, 128)])
def test_patch_conv_fused_batchnorm_statistics(shape):
    """air_conv3x3_patch_stats_bf16: per-channel sum / sum of squares of the stored output, accumulated in the conv
    epilogue (warp transpose-reduce + shared-memory partials + fp64 atomics) == sums of the output tensor."""
    from asvspoof2021_air_b200 import ops
    B, H, W, Cin, Cout = shape
    g = torch.Generator(device="cpu").manual_seed(31)
    x = torch.randn(B, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / (Cin * 9) ** 0.5).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).cuda().to(torch.bfloat16)
    wpk = torch.empty(9 * Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack3x3(w.contiguous(), Cin, Cout, 0, wpk)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ref = torch.empty_like(out)
    stats = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    ops.conv3x3_patch_stats(x, Cin, B, H, W, Cin, wpk, Cout, out, Cout, res, Cout, False, stats)
    ops.conv3x3_patch(x, Cin, B, H, W, Cin, wpk, Cout, ref, Cout, res, Cout, False)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)                                   # the fused path stores the same tensor
    o = out.double().reshape(-1, Cout)
    s, q = o.sum(0), (o * o).sum(0)
    assert torch.allclose(stats[:Cout], s, rtol=1e-5, atol=1e-3 * float(o.abs().sum(0).max()) * 1e-3)
    assert torch.allclose(stats[Cout:], q, rtol=1e-5)


S2_WGRAD_CASES = {
    # name: B, H, W, Cin, Cout, k        (stride 2; k = 3 pad 1 or k = 1 pad 0)
    "l2_3x3_even": (2, 18, 150, 64, 128, 3),
    "l3_3x3_odd": (2, 9, 375, 128, 256, 3),             # odd H and W: the last row / column has no partner in one parity class
    "l4_3x3_odd_h": (2, 5, 188, 256, 512, 3),
    "l2_sc_1x1": (2, 18, 150, 64, 128, 1),
    "l3_sc_1x1_odd": (3, 9, 375, 128, 256, 1),
}


@pytest.mark.parametrize("name", list(S2_WGRAD_CASES))
def test_stride2_wgrad_by_parity_classes(name):
    """air_conv_s2_wgrad_patch_bf16 (strided TMA sub-images, one stride-1 patch problem per input parity class) against
    torch's conv2d_weight on the same bf16 operands; x is a channel slice of a wider tensor (pixel stride != C)."""
    from asvspoof2021_air_b200 import ops
    B, H, W, Cin, Cout, k = S2_WGRAD_CASES[name]
    pad = 1 if k == 3 else 0
    g = torch.Generator(device="cpu").manual_seed(3)
    Ho, Wo = ops.conv_out_size(H, k, 2, pad, 1), ops.conv_out_size(W, k, 2, pad, 1)
    xw = torch.randn(B, H, W, Cin + 64, generator=g).cuda().to(torch.bfloat16)
    xn = xw[..., 64:]                                                     # view: pixel stride Cin + 64
    dyn = torch.randn(B, Ho, Wo, Cout, generator=g).cuda().to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_weight(xn.float().permute(0, 3, 1, 2), (Cout, Cin, k, k), dyn.float().permute(0, 3, 1, 2),
                                      stride=2, padding=pad)
    dwb = torch.zeros(Cout, k * k * Cin, device="cuda")
    ops.conv_s2_wgrad_patch(xn, Cin + 64, B, H, W, Cin, dyn, Cout, Ho, Wo, Cout, k, dwb)
    torch.cuda.synchronize()
    got = dwb.view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    err = (got - ref).abs()
    assert float(err.max()) <= 1e-3 * float(ref.abs().max()) + 1e-3, (name, float(err.max()), float(ref.abs().max()))


@pytest.mark.parametrize("shape", [(3, 5, 188, 64, 128), (4, 9, 131, 128, 64), (5, 3, 94, 64, 64)])
def test_patch_conv_flat_rows_pair_rows_of_neighbouring_images_bit_identically(shape):
    """Odd image heights: the patch kernel treats the batch as one image of B * H rows, so that the last row of an image
    shares a work item with the first row of the next one, and skips the taps that would cross the boundary (they would
    have multiplied the zero padding).  The result is bit-identical to one launch per image, with and without a residual."""
    from asvspoof2021_air_b200 import ops
    B, H, W, Cin, Cout = shape
    g = torch.Generator(device="cpu").manual_seed(41)
    x = torch.randn(B, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / (Cin * 9) ** 0.5).cuda()
    res = torch.randn(B, H, W, Cout, generator=g).cuda().to(torch.bfloat16)
    wpk = torch.empty(9 * Cin * Cout, device="cuda", dtype=torch.bfloat16)
    ops.pack3x3(w.contiguous(), Cin, Cout, 0, wpk)
    for r in (None, res):
        out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        ref = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv3x3_patch(x, Cin, B, H, W, Cin, wpk, Cout, out, Cout, r, Cout if r is not None else 0, False)
        for b in range(B):
            ops.conv3x3_patch(x[b:b + 1], Cin, 1, H, W, Cin, wpk, Cout, ref[b:b + 1], Cout,
                              None if r is None else r[b:b + 1], Cout if r is not None else 0, False)
        torch.cuda.synchronize()
        assert not torch.isnan(out.float()).any()
        assert torch.equal(out, ref)
    st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    out_s = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    ops.conv3x3_patch_stats(x, Cin, B, H, W, Cin, wpk, Cout, out_s, Cout, res, Cout, False, st)
    torch.cuda.synchronize()
    assert torch.equal(out_s, out)
    o64 = out.double().reshape(-1, Cout)
    assert torch.allclose(st[:Cout], o64.sum(0), rtol=1e-5, atol=1e-6 * float(o64.abs().sum(0).max()))


@pytest.mark.parametrize("name", ["res_l2_3x3_s2_64_128", "res_l4_3x3_512_512", "res_conv5_512_256"])
def test_generic_conv_fused_batchnorm_statistics(name):
    """air_conv_gemm_bf16_stats: the generic (gather) kernel stores the same tensor as air_conv_gemm_bf16 and its epilogue
    accumulates the per-channel sum / sum of squares of the stored output (rows beyond M excluded)."""
    ops, x, w, dy, Ho, Wo = _setup(name)
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES[name]
    xn = x.permute(0, 2, 3, 1).contiguous()
    wg = w.permute(0, 2, 3, 1).contiguous().reshape(-1)
    K = kh * kw * Cin
    wpk = ops.pack_weights(wg, 0, Cin, Cout, kh * kw)
    res = torch.randn(B, Ho, Wo, Cout, generator=torch.Generator(device="cpu").manual_seed(5)).cuda().to(torch.bfloat16)
    out = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    ref = torch.empty_like(out)
    st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
    ops.conv_gemm_stats(xn, Cin, B, H, W, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, 0, wpk, Cout, K, out, Cout, None, res, Cout,
                        False, st)
    ops.conv_gemm(xn, Cin, B, H, W, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, 0, wpk, Cout, K, ref, Cout, None, res, Cout, False)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    o64 = out.double().reshape(-1, Cout)
    assert torch.allclose(st[:Cout], o64.sum(0), rtol=1e-5, atol=1e-6 * float(o64.abs().sum(0).max()))
    assert torch.allclose(st[Cout:], (o64 * o64).sum(0), rtol=1e-5)
