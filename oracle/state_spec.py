"""Parameter/buffer inventories (reference state_dict key names and shapes) and a seeded,
reference-independent weight generator used by the parity tests and golden vectors.

TEST INFRASTRUCTURE ONLY -- see oracle/lfcc_oracle.py header.

Key names follow the reference modules (resnet.py:122-147, ecapa_tdnn.py:31-150,
loss.py:176-185); the lists are checked against the real modules' state_dict() in
tests/test_oracle_pinned.py when /root/reference is present (117 ResNet keys, 248 ECAPA keys).
"""
import math

import torch


def _bn_keys(prefix, c):
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
            (prefix + ".running_mean", (c,), "zeros"), (prefix + ".running_var", (c,), "ones"),
            (prefix + ".num_batches_tracked", (), "count")]


def resnet_spec(enc_dim=256, nclasses=2, num_nodes=3):
    """ResNet(num_nodes, enc_dim, '18', nclasses) -- resnet.py:122-147."""
    spec = [("conv1.weight", (16, 1, 9, 3), "conv")]
    spec += _bn_keys("bn1", 16)
    inp = 16
    for li, planes in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in range(2):
            p = "layer%d.%d" % (li, bi)
            cin = inp if bi == 0 else planes
            spec += _bn_keys(p + ".bn1", cin)
            spec.append((p + ".conv1.weight", (planes, cin, 3, 3), "conv"))
            spec += _bn_keys(p + ".bn2", planes)
            spec.append((p + ".conv2.weight", (planes, planes, 3, 3), "conv"))
            if bi == 0:   # stride != 1 or in_planes != planes holds for every first block of '18'
                spec.append((p + ".shortcut.0.weight", (planes, cin, 1, 1), "conv"))
        inp = planes
    spec.append(("conv5.weight", (256, 512, num_nodes, 3), "conv"))
    spec += _bn_keys("bn5", 256)
    spec += [("fc.weight", (enc_dim, 512), "linear_w"), ("fc.bias", (enc_dim,), "linear_b"),
             ("fc_mu.weight", (nclasses, enc_dim), "linear_w"), ("fc_mu.bias", (nclasses,), "linear_b"),
             ("attention.att_weights", (1, 256), "att")]
    return spec


def ecapa_spec(C=512, scale=8, n_out=2, n_mels=60, bottleneck=128):
    """Res2Net2(Bottle2neck, C, model_scale, nOut, n_mels) -- ecapa_tdnn.py:97-150."""
    def conv(prefix, co, ci, k):
        return [(prefix + ".weight", (co, ci, k), "conv"), (prefix + ".bias", (co,), "conv_b")]
    spec = conv("conv1", C, n_mels, 5) + _bn_keys("bn1", C)
    width = C // scale
    for li in (1, 2, 3):
        p = "layer%d" % li
        spec += conv(p + ".conv1", C, C, 1) + _bn_keys(p + ".bn1", C)
        for i in range(scale - 1):
            spec += conv(p + ".convs.%d" % i, width, width, 3)
        for i in range(scale - 1):
            spec += _bn_keys(p + ".bns.%d" % i, width)
        spec += conv(p + ".conv3", C, C, 1) + _bn_keys(p + ".bn3", C)
        spec += conv(p + ".se.se.1", bottleneck, C, 1) + _bn_keys(p + ".se.se.3", bottleneck)
        spec += conv(p + ".se.se.4", C, bottleneck, 1)
    spec += conv("layer4", 1536, 3 * C, 1)
    spec += conv("attention.0", 128, 1536 * 3, 1) + _bn_keys("attention.2", 128)
    spec += conv("attention.3", 1536, 128, 1)
    spec += _bn_keys("bn5", 3072)
    spec += [("fc6.weight", (256, 3072), "linear_w"), ("fc6.bias", (256,), "linear_b"),
             ("fc7.weight", (n_out, 256), "linear_w"), ("fc7.bias", (n_out,), "linear_b")]
    spec += _bn_keys("bn7", n_out)
    return spec


def seeded_state(spec, seed):
    """Deterministic fp32 state dict: every tensor drawn from its own torch CPU generator."""
    sd = {}
    for i, (key, shape, kind) in enumerate(spec):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if kind == "conv":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind in ("conv_b", "linear_b"):
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "linear_w":
            t = torch.randn(shape, generator=g) / math.sqrt(shape[1])
        elif kind == "bn_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_b":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "att":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "ones":
            t = torch.ones(shape)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise KeyError(kind)
        sd[key] = t
    return sd


def seeded_center(feat_dim, seed):
    g = torch.Generator().manual_seed(seed * 100003 + 99991)
    return torch.randn(1, feat_dim, generator=g) * 0.3


def trainable_keys(spec):
    return [k for k, _, kind in spec if kind not in ("zeros", "ones", "count")]


def seeded_waves(batch, length=64000, seed=0, edge_rows=False):
    """SURVEY.md section 8(d) synthetic input: 0.1*randn; optional 3 edge rows appended
    (all-zeros, full-scale +-1 square wave, single impulse at n=0)."""
    g = torch.Generator().manual_seed(seed)
    w = 0.1 * torch.randn(batch, length, generator=g)
    if edge_rows:
        e = torch.zeros(3, length)
        n = torch.arange(length)
        e[1] = torch.where((n // 40) % 2 == 0, 1.0, -1.0)
        e[2, 0] = 1.0
        w = torch.cat([w, e], 0)
    return w


def seeded_labels(batch, seed=0):
    g = torch.Generator().manual_seed(seed + 7)
    lab = torch.randint(0, 2, (batch,), generator=g)
    if batch >= 2:
        lab[0], lab[1] = 0, 1       # both classes present
    return lab
