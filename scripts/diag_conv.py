"""GPU diagnostic for the tcgen05 implicit-GEMM conv: every (shape, descriptor-variant) in its own
process so a trapped launch cannot poison the others.  Prints max errors vs torch fp32 conv."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw
    "gemm_1x1_64_64": (1, 1, 256, 64, 64, 1, 1, 1, 1, 0, 0, 1, 1),
    "gemm_1x1_128_256": (2, 1, 300, 128, 256, 1, 1, 1, 1, 0, 0, 1, 1),
    "c3x3_64_64": (2, 18, 75, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1),
    "c3x3_s2_64_128": (2, 18, 75, 64, 128, 3, 3, 2, 2, 1, 1, 1, 1),
    "c3x3_16_64": (2, 18, 75, 16, 64, 3, 3, 1, 1, 1, 1, 1, 1),
    "c1x1_s2_64_128": (2, 18, 75, 64, 128, 1, 1, 2, 2, 0, 0, 1, 1),
    "c3x3_512_512": (2, 3, 94, 512, 512, 3, 3, 1, 1, 1, 1, 1, 1),
    "conv5_512_256": (2, 3, 94, 512, 256, 3, 3, 1, 1, 0, 1, 1, 1),
    "k3_dil3_64_64": (3, 1, 750, 64, 64, 1, 3, 1, 1, 0, 3, 1, 3),
    "k5_64_512": (2, 1, 750, 64, 512, 1, 5, 1, 1, 0, 2, 1, 1),
    "k1_1536_1536": (1, 1, 750, 1536, 1536, 1, 1, 1, 1, 0, 0, 1, 1),
}


def run_case(name, flags, what):
    import torch
    import torch.nn.functional as F
    from asvspoof2021_air_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W, Cin, Cout, kh, kw, sh, sw, ph, pw, dh, dw = CASES[name]
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5).cuda()
    wq = w.to(torch.bfloat16).float()
    Ho, Wo = ops.conv_out_size(H, kh, sh, ph, dh), ops.conv_out_size(W, kw, sw, pw, dw)
    w_gemm = w.permute(0, 2, 3, 1).contiguous()          # [Cout][kh][kw][Cin]
    res = {"case": name, "flags": flags, "what": what}
    if what == "fprop":
        ref = F.conv2d(x.float(), wq, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))
        xn = x.permute(0, 2, 3, 1).contiguous()
        wpk = ops.pack_weights(w_gemm, 0, Cin, Cout, kh * kw)
        out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv_gemm(xn, Cin, B, H, W, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw, dh, dw, 0, wpk, Cout, kh * kw * Cin,
                      out, Cout, flags=flags)
        torch.cuda.synchronize()
        got = out.float().permute(0, 3, 1, 2)
    elif what == "wgrad":
        dy = torch.randn(B, Cout, Ho, Wo, generator=g).cuda().to(torch.bfloat16)
        ref = torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, kh, kw), dy.float(), stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))
        xn = x.permute(0, 2, 3, 1).contiguous()
        dyn = dy.permute(0, 2, 3, 1).contiguous()
        dwb = torch.zeros(Cout, kh * kw * Cin, device="cuda")
        ops.conv_wgrad(xn, Cin, B, H, W, Cin, dyn, Cout, Ho, Wo, Cout, kh, kw, sh, sw, ph, pw, dh, dw, dwb, flags=flags)
        torch.cuda.synchronize()
        got = dwb.view(Cout, kh, kw, Cin).permute(0, 3, 1, 2)
    else:   # dgrad
        dy = torch.randn(B, Cout, Ho, Wo, generator=g).cuda().to(torch.bfloat16)
        ref = torch.nn.grad.conv2d_input((B, Cin, H, W), wq, dy.float(), stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw))
        dyn = dy.permute(0, 2, 3, 1).contiguous()
        wpk = ops.pack_weights(w_gemm, 1, Cin, Cout, kh * kw)
        out = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv_gemm(dyn, Cout, B, Ho, Wo, Cout, H, W, kh, kw, sh, sw, ph, pw, dh, dw, 1, wpk, Cin, kh * kw * Cout,
                      out, Cin, flags=flags)
        torch.cuda.synchronize()
        got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    res.update(max_err=float(err.max()), ref_max=float(ref.abs().max()), nan=int(torch.isnan(got).sum()),
               mean_err=float(err[~torch.isnan(err)].mean()) if (~torch.isnan(err)).any() else None)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 4:
        run_case(sys.argv[1], int(sys.argv[2]), sys.argv[3])
    else:
        names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CASES)
        for what in (os.environ.get("DIAG_WHAT", "fprop,dgrad,wgrad").split(",")):
            for flags in (0, 1):
                for n in names:
                    r = subprocess.run([sys.executable, __file__, n, str(flags), what], capture_output=True, text=True, timeout=300)
                    out = r.stdout.strip().splitlines()
                    print(out[-1] if out else json.dumps({"case": n, "flags": flags, "what": what, "rc": r.returncode,
                                                           "err": r.stderr.strip().splitlines()[-1:]}), flush=True)
                if flags == 0 and what == "fprop":
                    pass
