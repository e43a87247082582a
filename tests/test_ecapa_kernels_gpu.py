"""GPU parity of the ECAPA non-GEMM kernels (csrc/ecapa.cu) and the extended conv / BN / linear entry
points against plain PyTorch fp32 (autograd for the backward) on the same bf16-rounded inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _rel(a, b):
    a, b = a.double().reshape(-1).cpu(), b.double().reshape(-1).cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_time_stats_mean_std_sum():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, T, C, ld = 3, 750, 520, 1040
    big = torch.relu(torch.randn(B, T, ld, generator=g)).to(BF)
    big[:, :, 8] = 0.0                                   # dead channel -> clamp
    x = big[:, :, :C]
    mean, std, tot = (torch.empty(B, C, device="cuda") for _ in range(3))
    xd = big.cuda()[:, :, :C]
    ops.time_stats(xd, ld, B, T, C, mean, std, 1e-4)
    ops.time_stats(xd, ld, B, T, C, tot, None, -1.0)
    torch.cuda.synchronize()
    xf = x.float()
    assert _rel(mean, xf.mean(1)) < 1e-5
    assert _rel(std, torch.sqrt(xf.var(1).clamp(min=1e-4))) < 1e-4         # ecapa_tdnn.py:171
    assert _rel(tot, xf.sum(1)) < 1e-5
    assert abs(float(std[0, 8]) - 1e-2) < 1e-6


def test_attentive_stats_pooling_forward_backward():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, T, C = 3, 200, 264
    e = (2 * torch.randn(B, T, C, generator=g)).to(BF)
    x = torch.relu(torch.randn(B, T, C, generator=g)).to(BF)
    x[:, :, 5] = 0.25                                    # constant channel -> sg clamps at 1e-2, ctx std clamps
    dout = torch.randn(B, 2 * C, generator=g)
    dcm, dcs = torch.randn(B, C, generator=g), torch.randn(B, C, generator=g)
    xr, er = x.float().requires_grad_(True), e.float().requires_grad_(True)
    w = F.softmax(er, dim=1)
    mu = (xr * w).sum(1)
    sg = torch.sqrt(((xr ** 2) * w).sum(1) - mu ** 2).clamp(min=1e-4) if False else torch.sqrt((((xr ** 2) * w).sum(1) - mu ** 2).clamp(min=1e-4))
    cmean = xr.mean(1)
    cstd = torch.sqrt(xr.var(1).clamp(min=1e-4))
    ((torch.cat((mu, sg), 1) * dout).sum() + (cmean * dcm).sum() + (cstd * dcs).sum()).backward()
    out, smax, ssum, sq = torch.empty(B, 2 * C, device="cuda"), *(torch.empty(B, C, device="cuda") for _ in range(3))
    cm, cs = torch.empty(B, C, device="cuda"), torch.empty(B, C, device="cuda")
    ed, xd = e.cuda(), x.cuda()
    ops.asp_fwd(ed, C, xd, C, B, T, C, out, smax, ssum, sq)
    ops.time_stats(xd, C, B, T, C, cm, cs, 1e-4)
    de, dx = torch.empty(B, T, C, device="cuda", dtype=BF), torch.empty(B, T, C, device="cuda", dtype=BF)
    ops.asp_bwd(ed, C, xd, C, B, T, C, out, dout.cuda(), smax, ssum, sq, cm, cs, dcm.cuda(), dcs.cuda(), 1e-4, de, C, dx, C)
    torch.cuda.synchronize()
    assert _rel(out, torch.cat((mu, sg), 1).detach()) < 1e-4
    assert _rel(de.float(), er.grad) < 1e-2 and _rel(dx.float(), xr.grad) < 1e-2
    # split form used by the engine: direct part first, context part + ReLU mask later
    z = torch.zeros(B, C, device="cuda")
    ops.asp_bwd(ed, C, xd, C, B, T, C, out, dout.cuda(), smax, ssum, sq, cm, cs, z, z, 1e-4, de, C, dx, C)
    ops.ctx_bwd_mask(xd, C, B, T, C, cm, cs, dcm.cuda(), dcs.cuda(), 1e-4, dx, C)
    torch.cuda.synchronize()
    assert _rel(dx.float(), xr.grad * (x.float() > 0)) < 1.5e-2


def test_se_gate_residual_forward_backward():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, T, C = 2, 300, 512
    x, res = torch.randn(B, T, C, generator=g).to(BF), torch.randn(B, T, C, generator=g).to(BF)
    gate = torch.rand(B, C, generator=g)
    dout = torch.randn(B, T, C, generator=g).to(BF)
    ds = torch.randn(B, C, generator=g)
    out, dx = torch.empty(B, T, C, device="cuda", dtype=BF), torch.empty(B, T, C, device="cuda", dtype=BF)
    dg = torch.empty(B, C, device="cuda")
    ops.scale_residual(x.cuda(), C, gate.cuda(), res.cuda(), C, out, C, B, T, C)
    ops.se_dgate(dout.cuda(), C, x.cuda(), C, B, T, C, dg)
    ops.se_apply_bwd(dout.cuda(), C, gate.cuda(), ds.cuda(), dx, C, B, T, C)
    torch.cuda.synchronize()
    assert _rel(out.float(), x.float() * gate[:, None] + res.float()) < 4e-3
    assert _rel(dg, (dout.float() * x.float()).sum(1)) < 1e-5
    assert _rel(dx.float(), dout.float() * gate[:, None] + ds[:, None] / T) < 4e-3


@pytest.mark.parametrize("M,C,relu_in", [(4, 128, True), (256, 3072, False), (16, 2, False)])
def test_bn1d_f32_forward_backward(M, C, relu_in):
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    x, dy = torch.randn(M, C, generator=g), torch.randn(M, C, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(C), torch.ones(C)
    y = F.batch_norm(F.relu(xr) if relu_in else xr, rm, rv, gr, br, training=True, momentum=0.1, eps=1e-5)
    y.backward(dy)
    yd, dxd = torch.empty(M, C, device="cuda"), torch.empty(M, C, device="cuda")
    sm, si, rmd, rvd = torch.empty(C, device="cuda"), torch.empty(C, device="cuda"), torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn1d_fwd(x.cuda(), yd, M, C, relu_in, gamma.cuda(), beta.cuda(), True, sm, si, rmd, rvd)
    ops.bn1d_bwd(dy.cuda(), x.cuda(), dxd, M, C, relu_in, gamma.cuda(), sm, si, dg, db)
    torch.cuda.synchronize()
    assert _rel(yd, y.detach()) < 1e-4 and _rel(rmd, rm) < 1e-5 and _rel(rvd, rv) < 1e-4
    assert _rel(dxd, xr.grad) < 2e-3 and _rel(dg, gr.grad) < 1e-3 and _rel(db, br.grad) < 1e-4
    ye = torch.empty(M, C, device="cuda")
    ops.bn1d_fwd(x.cuda(), ye, M, C, relu_in, gamma.cuda(), beta.cuda(), False, None, None, rmd, rvd)
    ref = F.batch_norm(F.relu(x) if relu_in else x, rm, rv, gamma, beta, training=False, eps=1e-5)
    assert _rel(ye, ref) < 1e-5


def test_sigmoid_copy_colsum():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1000, generator=g)
    y, dx = torch.empty(1000, device="cuda"), torch.empty(1000, device="cuda")
    ops.sigmoid_fwd(x.cuda(), y, 1000)
    ops.sigmoid_bwd(torch.ones(1000, device="cuda"), y, dx, 1000)
    s = torch.sigmoid(x)
    assert _rel(y, s) < 1e-6 and _rel(dx, s * (1 - s)) < 1e-5
    M, C = 999, 1536
    a = torch.randn(M, C, generator=g).to(BF).cuda()
    dst = torch.zeros(M, 512, device="cuda", dtype=BF)
    ops.copy_channels(a[:, 448:512], C, dst[:, 64:128], 512, M, 64)
    assert torch.equal(dst[:, 64:128], a[:, 448:512]) and (dst[:, :64] == 0).all() and (dst[:, 128:] == 0).all()
    cs = torch.zeros(C, device="cuda")
    ops.colsum(a, C, M, C, cs)
    torch.cuda.synchronize()
    assert _rel(cs, a.float().sum(0)) < 1e-5


def test_bn_apply_add_and_bwd_bias():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(6)
    M, C = 3000, 64
    x = torch.relu(torch.randn(M, C, generator=g)).to(BF)
    add = torch.randn(M, C, generator=g).to(BF)
    dy = torch.randn(M, C, generator=g).to(BF)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    xr = x.float().requires_grad_(True)
    y = F.batch_norm(xr, torch.zeros(C), torch.ones(C), gamma, beta, training=True, eps=1e-5)
    y.backward(dy.float())
    yd, y2, dxd = (torch.empty(M, C, device="cuda", dtype=BF) for _ in range(3))
    sums, rsum = torch.zeros(2 * C, dtype=torch.float64, device="cuda"), torch.zeros(2 * C, dtype=torch.float64, device="cuda")
    sm, si = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dg, db, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
    ops.bn_stats(x.cuda(), C, M, C, sums)
    ops.bn_apply_add(x.cuda(), C, yd, C, M, C, sums, gamma.cuda(), beta.cuda(), False, True, sm, si, None, None,
                     add.cuda(), C, y2, C)
    ops.bn_bwd_bias(dy.cuda(), C, x.cuda(), C, None, 0, dxd, C, M, C, 1, sm, si, gamma.cuda(), beta.cuda(), rsum, dg, db, dbias)
    torch.cuda.synchronize()
    assert torch.allclose(yd.float().cpu(), y.detach(), atol=2e-2, rtol=1e-2)
    assert torch.equal(y2.float().cpu(), (yd.float().cpu() + add.float()).to(BF).float())
    ref_dx = xr.grad * (x.float() > 0)
    assert _rel(dxd.float(), ref_dx) < 1e-2
    assert _rel(dbias, ref_dx.sum(0)) < 5e-3


def test_conv_gemm_ex_per_utterance_bias_and_second_output():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(7)
    B, T, K, N = 3, 750, 256, 128
    x = torch.randn(B, T, K, generator=g).to(BF)
    wfull = torch.randn(N, 3 * K, generator=g) / K ** 0.5                 # only the first K columns go through the GEMM
    u = torch.randn(B, N, generator=g)
    res = torch.randn(B, T, N, generator=g).to(BF)
    wpk = torch.empty(ops.packed_elems(N, K), device="cuda", dtype=BF)
    wd = wfull.cuda()
    ops.pack_weights_ld(wd, 3 * K, 0, K, N, 1, wpk)
    out, out2 = torch.empty(B, T, N, device="cuda", dtype=BF), torch.empty(B, T, N, device="cuda", dtype=BF)
    ops.conv_gemm_ex(x.cuda(), K, B, 1, T, K, 1, T, 1, 1, 1, 1, 0, 0, 1, 1, 0, wpk, N, K, out, N, u.cuda(), res.cuda(), N,
                     False, 0, T, out2, N)
    torch.cuda.synchronize()
    acc = x.float() @ wfull[:, :K].to(BF).float().t() + u[:, None, :]
    assert _rel(out2.float(), acc) < 4e-3 and _rel(out.float(), acc + res.float()) < 4e-3
    # wgrad into a column slice of a wider gradient
    dy = torch.randn(B, T, N, generator=g).to(BF)
    gw = torch.zeros(N, 3 * K, device="cuda")
    ops.conv_wgrad_ld(x.cuda(), K, B, 1, T, K, dy.cuda(), N, 1, T, N, 1, 1, 1, 1, 0, 0, 1, 1, gw, 3 * K)
    torch.cuda.synchronize()
    ref = dy.float().reshape(-1, N).t() @ x.float().reshape(-1, K)
    assert _rel(gw[:, :K], ref) < 1e-3 and float(gw[:, K:].abs().max()) == 0.0


def test_linear_ld_forward_backward():
    from asvspoof2021_air_b200 import ops
    g = torch.Generator().manual_seed(8)
    M, N, K = 5, 128, 1536
    W = torch.randn(N, 3 * K, generator=g) / K ** 0.5
    x1, x2, b, dy = torch.randn(M, K, generator=g), torch.randn(M, K, generator=g), torch.randn(N, generator=g), torch.randn(M, N, generator=g)
    Wd = W.cuda()
    y = torch.empty(M, N, device="cuda")
    ops.linear_fwd_ld(x1.cuda(), Wd[:, K:], 3 * K, b.cuda(), y, M, N, K)
    ops.linear_fwd_ld(x2.cuda(), Wd[:, 2 * K:], 3 * K, None, y, M, N, K, accumulate=True)
    dx, gW = torch.empty(M, K, device="cuda"), torch.zeros(N, 3 * K, device="cuda")
    ops.linear_bwd_ld(x1.cuda(), Wd[:, K:], 3 * K, dy.cuda(), dx, gW[:, K:], None, M, N, K)
    torch.cuda.synchronize()
    assert _rel(y, x1 @ W[:, K:2 * K].t() + b + x2 @ W[:, 2 * K:].t()) < 1e-5
    assert _rel(dx, dy @ W[:, K:2 * K]) < 1e-5 and _rel(gW[:, K:2 * K], dy.t() @ x1) < 1e-5
    assert float(gW[:, :K].abs().max()) == 0.0 and float(gW[:, 2 * K:].abs().max()) == 0.0
