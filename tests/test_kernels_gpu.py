"""GPU parity of the HBM-bound kernels (BatchNorm fwd/bwd, stem conv, SelfAttention pooling,
linear, OC-Softmax, Adam/SGD) through the C ABI against the fp32 oracle restatement
(oracle/nets_oracle.py) / plain PyTorch fp32 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import nets_oracle as no

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("M,C,relu", [(5000, 16, True), (3001, 64, True), (777, 256, False), (300, 512, True),
                                       (64, 1536, False)])
def test_batchnorm_train_forward_backward(M, C, relu):
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    gamma = 1 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    dy = torch.randn(M, C, generator=g).to(torch.bfloat16)
    rm, rv = torch.zeros(C), torch.ones(C)
    # oracle (fp32, autograd)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_o, rv_o = rm.clone(), rv.clone()
    y = F.batch_norm(xr, rm_o, rv_o, gr, br, training=True, momentum=0.1, eps=1e-5)
    if relu:
        y = F.relu(y)
    y.backward(dy.float())
    # ours
    xd, dyd = x.cuda(), dy.cuda()
    yd, dxd = torch.empty_like(xd), torch.empty_like(xd)
    sums = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    rsum = torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    sm, si = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    rmd, rvd = rm.cuda(), rv.cuda()
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn_stats(xd, C, M, C, sums)
    ops.bn_apply(xd, C, yd, C, M, C, sums, gamma.cuda(), beta.cuda(), relu, True, sm, si, rmd, rvd)
    if relu:
        ops.bn_bwd(dyd, C, xd, C, None, 0, dxd, C, M, C, 0, sm, si, gamma.cuda(), beta.cuda(), rsum, dg, db)
    else:
        # order 1 (ECAPA): y = bn(x), x = relu(.) upstream -> mask by x > 0; emulate by comparing on x > 0 only
        ops.bn_bwd(dyd, C, xd, C, None, 0, dxd, C, M, C, 1, sm, si, gamma.cuda(), beta.cuda(), rsum, dg, db)
    torch.cuda.synchronize()
    assert torch.allclose(yd.float().cpu(), y.detach(), atol=2e-2, rtol=1e-2)
    assert _rel(rmd, rm_o) < 1e-5 and _rel(rvd, rv_o) < 1e-5
    ref_dx = xr.grad
    got_dx = dxd.float().cpu()
    if not relu:
        ref_dx = ref_dx * (x.float() > 0)
    assert _rel(got_dx, ref_dx) < 1e-2
    assert _rel(dg, gr.grad) < 2e-3 and _rel(db, br.grad) < 2e-3


def test_batchnorm_eval_uses_running_stats():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(5)
    M, C = 1000, 128
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    rm, rv = 0.2 * torch.randn(C, generator=g), 1 + 0.2 * torch.rand(C, generator=g)
    y = F.relu(F.batch_norm(x.float(), rm, rv, gamma, beta, training=False, eps=1e-5))
    yd = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    ops.bn_apply(x.cuda(), C, yd, C, M, C, None, gamma.cuda(), beta.cuda(), True, False, None, None, rm.cuda(), rv.cuda())
    torch.cuda.synchronize()
    assert torch.allclose(yd.float().cpu(), y, atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("W", [750, 1000, 97])          # 1000: wider than one staged segment of the weight-gradient kernel
def test_stem_conv_forward_and_wgrad(W):
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(6)
    B, H = 3, 60
    x = torch.randn(B, H, W, generator=g).to(torch.bfloat16)
    w = torch.randn(16, 1, 9, 3, generator=g) * 0.2
    ref = F.conv2d(x.float().unsqueeze(1), w, stride=(3, 1), padding=(1, 1))            # resnet.py:131
    Ho, Wo = ref.shape[2], ref.shape[3]
    y = torch.empty(B, Ho, Wo, 16, device="cuda", dtype=torch.bfloat16)
    wg = w.permute(0, 2, 3, 1).contiguous().view(16, 27).cuda()
    ops.stem_fwd(x.cuda(), B, H, W, 9, 3, 3, 1, 1, 1, wg, 16, y)
    dy = torch.randn(B, 16, Ho, Wo, generator=g).to(torch.bfloat16)
    refw = torch.nn.grad.conv2d_weight(x.float().unsqueeze(1), (16, 1, 9, 3), dy.float(), stride=(3, 1), padding=(1, 1))
    dw = torch.zeros(16, 27, device="cuda")
    ops.stem_wgrad(x.cuda(), B, H, W, 9, 3, 3, 1, 1, 1, dy.permute(0, 2, 3, 1).contiguous().cuda(), 16, dw)
    torch.cuda.synchronize()
    assert torch.allclose(y.float().cpu().permute(0, 3, 1, 2), ref, atol=3e-2, rtol=1e-2)
    assert _rel(dw.view(16, 9, 3, 1).permute(0, 3, 1, 2), refw) < 1e-4


def test_selfattention_pooling_forward_backward():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(7)
    B, T, C = 5, 94, 256
    x = torch.relu(torch.randn(B, T, C, generator=g)).to(torch.bfloat16)
    x[:, :, 3] = 0                                       # a dead channel: std == 0, gradient must stay finite
    att = 0.1 * torch.randn(1, C, generator=g)
    dstats = torch.randn(B, 2 * C, generator=g)
    xr = x.float().requires_grad_(True)
    ar = att.clone().requires_grad_(True)
    # reference semantics (resnet.py:38-42) keep d std/dx finite through the 1e-5 noise; without the
    # noise the oracle's autograd gives nan for the dead channel, which the kernel defines as 0
    out = no.self_attention_pool(xr, ar)
    out.backward(dstats)
    stats = torch.empty(B, 2 * C, device="cuda")
    p, th = torch.empty(B, T, device="cuda"), torch.empty(B, T, device="cuda")
    dx = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
    datt = torch.zeros(C, device="cuda")
    ops.selfattn_pool_fwd(x.cuda(), att.cuda(), stats, p, th, B, T, C, -1)
    ops.selfattn_pool_bwd(x.cuda(), att.cuda(), p, th, stats, dstats.cuda(), dx, datt, B, T, C, -1)
    torch.cuda.synchronize()
    assert _rel(stats, out.detach()) < 1e-5
    live = [c for c in range(C) if c != 3]
    ref_dx = xr.grad[:, :, live]
    assert torch.isfinite(dx.float()).all()
    assert _rel(dx.float().cpu()[:, :, live], ref_dx) < 1e-2          # bf16 output rounding
    ref_da = torch.nan_to_num(ar.grad.reshape(-1), nan=0.0)
    got_da = datt.cpu()
    keep = torch.tensor(live)
    # att gradient: contributions of the dead channel are zero in both (x == 0 there)
    assert _rel(got_da[keep], ref_da[keep]) < 1e-3 or torch.isnan(ar.grad).any()


def test_selfattention_noise_matches_reference_scale():
    """The counter-based noise has the reference's distribution: std(noise) ~ 1e-5, effect on stats < 1e-4 rel."""
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(8)
    B, T, C = 4, 94, 256
    x = torch.relu(torch.randn(B, T, C, generator=g)).to(torch.bfloat16).cuda()
    att = (0.1 * torch.randn(1, C, generator=g)).cuda()
    s0, s1 = torch.empty(B, 2 * C, device="cuda"), torch.empty(B, 2 * C, device="cuda")
    p, th = torch.empty(B, T, device="cuda"), torch.empty(B, T, device="cuda")
    ops.selfattn_pool_fwd(x, att, s0, p, th, B, T, C, -1)
    ops.selfattn_pool_fwd(x, att, s1, p, th, B, T, C, 1234)
    torch.cuda.synchronize()
    assert torch.equal(s0[:, :C], s1[:, :C])
    assert not torch.equal(s0[:, C:], s1[:, C:])
    assert _rel(s1, s0) < 1e-4


@pytest.mark.parametrize("M,N,K", [(17, 256, 512), (1024, 256, 3072), (37, 250, 2085), (2100, 40, 33)])
def test_linear_forward_backward(M, N, K):
    """Ragged tiles in all three dimensions; K >= 2048 takes the fp64-accumulator form (fc6), M >= 2048 takes it in the
    weight-gradient product."""
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(9)
    x, W, b, dy = (torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5,
                   torch.randn(N, generator=g), torch.randn(M, N, generator=g))
    xr, Wr, br = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    F.linear(xr, Wr, br).backward(dy)
    y = torch.empty(M, N, device="cuda")
    dx, dW, db = torch.empty(M, K, device="cuda"), torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    ops.linear_fwd(x.cuda(), W.cuda(), b.cuda(), y, M, N, K)
    ops.linear_bwd(x.cuda(), W.cuda(), dy.cuda(), dx, dW, db, M, N, K)
    torch.cuda.synchronize()
    assert _rel(y, F.linear(x, W, b)) < 1e-5
    assert _rel(dx, xr.grad) < 1e-5 and _rel(dW, Wr.grad) < 1e-5 and _rel(db, br.grad) < 1e-5
    # split-K form (fc6): fp64 partial sums added in split order -- within an fp32 ulp of the fp64 product for any number of
    # splits (more splits than K tiles included), and the same bits on a second call
    ref64 = (x.double() @ W.double().t() + b.double())
    for splits in (1, 3, 8, 64):
        part = torch.full((splits * M * N,), float("nan"), device="cuda", dtype=torch.float64)
        ys, ys2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
        ops.linear_fwd_splitk(x.cuda(), W.cuda(), b.cuda(), ys, M, N, K, part, splits)
        ops.linear_fwd_splitk(x.cuda(), W.cuda(), b.cuda(), ys2, M, N, K, part, splits)
        torch.cuda.synchronize()
        assert torch.equal(ys, ys2)
        assert float((ys.cpu().double() - ref64).abs().max()) <= 1.5e-7 * float(ref64.abs().max())
        if K >= 2048:
            assert float((ys - y).abs().max()) <= 1.5e-7 * float(ref64.abs().max())


@pytest.mark.parametrize("B", [4, 256, 1024])
def test_ocsoftmax_forward_backward_and_ce(B):
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(B)
    D = 256
    x = torch.randn(B, D, generator=g)
    labels = torch.randint(0, 2, (B,), generator=g)
    center = torch.randn(1, D, generator=g) * 0.3
    logits = torch.randn(B, 2, generator=g)
    xr, cr = x.clone().requires_grad_(True), center.clone().requires_grad_(True)
    loss, score = no.ocsoftmax(cr, xr, labels, 0.9, 0.2, 20.0)            # loss.py:187-206
    loss.backward()
    ce = no.cross_entropy(logits, labels)
    l, s = torch.empty(1, device="cuda"), torch.empty(B, device="cuda")
    dfeat, dc, cev = torch.empty(B, D, device="cuda"), torch.zeros(1, D, device="cuda"), torch.empty(1, device="cuda")
    ops.ocsoftmax(x.cuda(), labels.cuda(), center.cuda(), B, D, 0.9, 0.2, 20.0, 1.0, l, s, dfeat, dc, logits.cuda(), 2, cev)
    torch.cuda.synchronize()
    assert abs(float(l) - float(loss)) <= 1e-5 * abs(float(loss)) + 1e-7
    assert torch.allclose(s.cpu(), score.detach(), atol=1e-6)
    assert _rel(dfeat, xr.grad) < 1e-4 and _rel(dc, cr.grad) < 1e-4
    assert abs(float(cev) - float(ce)) <= 1e-5 * abs(float(ce))


def test_adam_l2_and_sgd_steps_match_torch_optim():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(10)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4)   # main_train.py:175
    p = p0.clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        gr = torch.randn(n, generator=g)
        ref.grad = gr.clone()
        opt.step()
        ops.adam_l2_step(p, gr.cuda(), m, v, n, 5e-4, 0.9, 0.999, 1e-8, 5e-4, step)
    torch.cuda.synchronize()
    assert torch.allclose(p.cpu(), ref.detach(), atol=1e-6, rtol=1e-5)
    c = torch.randn(256, generator=g)
    gc = torch.randn(256, generator=g)
    cd = c.clone().cuda()
    ops.sgd_step(cd, gc.cuda(), 256, 5e-4)
    torch.cuda.synchronize()
    assert torch.allclose(cd.cpu(), c - 5e-4 * gc, atol=1e-7)
