"""CPU oracle (torch fp32, functional) for ResNet-18-OC, ECAPA-TDNN (Res2Net2), OC-Softmax,
the logged CE loss, and the Adam(L2)/SGD optimiser step.

TEST INFRASTRUCTURE ONLY -- see oracle/lfcc_oracle.py header.  These are floating-point
kernels, so the oracle is a plain torch fp32 restatement (autograd provides the backward the
reference gets from PyTorch).  State is a plain dict with the REFERENCE state_dict key names
(SURVEY.md section 8b), so weights copied from the reference modules drive both.

Parity status: pinned against the reference's own modules run under oracle/ref_shim.py
(tests/test_oracle_pinned.py in the authoring container; tests/golden/*.npz elsewhere).
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------
# BatchNorm helper: train mode uses batch stats (biased var) and updates running stats with
# unbiased var, momentum 0.1, eps 1e-5 (nn.BatchNorm defaults used at resnet.py:54-56,132).
# ------------------------------------------------------------------------------------
def _bn(sd, prefix, x, training, update_running=False):
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if training and not update_running:
        rm, rv = rm.clone(), rv.clone()
    return F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"],
                        training=training, momentum=0.1, eps=1e-5)


# ------------------------------------------------------------------------------------
# ResNet-18 (pre-activation) + SelfAttention pooling            resnet.py:23-46,49-69,122-191
# ------------------------------------------------------------------------------------
def self_attention_pool(x, att_weights, noise=None):
    """x (B,T,C) -> (B,2C).  resnet.py:23-46.

    weights = x . a ; attentions = softmax_T(tanh(weights)); weighted = x * attentions
    avg = sum_T weighted ; std = unbiased std_T(weighted + noise); noise = 1e-5*randn in the
    reference (drawn every forward, train and eval) -- passed in explicitly here."""
    w = torch.matmul(x, att_weights.reshape(-1))            # (B,T)   :26
    a = F.softmax(torch.tanh(w), dim=1)                     # :29-33
    weighted = x * a.unsqueeze(2)
    avg = weighted.sum(1)                                   # :41
    if noise is None:
        noise = torch.zeros_like(weighted)
    std = (weighted + noise).std(1)                         # unbiased (N-1)  :41
    return torch.cat((avg, std), 1)


def _q_none(t):
    return t


def _q_bf16(t):
    """Round to bf16 and back (differentiable: gradient passes straight through)."""
    return t.to(torch.bfloat16).float()


def _preact_block(sd, p, x, stride, training, upd, q=_q_none):
    """resnet.py:63-69.  q marks the points where the bf16 tensor-core path stores bf16."""
    out = q(F.relu(_bn(sd, p + ".bn1", x, training, upd)))
    if (p + ".shortcut.0.weight") in sd:
        sc = q(F.conv2d(out, q(sd[p + ".shortcut.0.weight"]), stride=stride))
    else:
        sc = x
    out = q(F.conv2d(out, q(sd[p + ".conv1.weight"]), stride=stride, padding=1))
    out = F.conv2d(q(F.relu(_bn(sd, p + ".bn2", out, training, upd))), q(sd[p + ".conv2.weight"]),
                   stride=1, padding=1)
    return q(out + sc)


def resnet_forward(sd, x, training=True, noise=None, update_running=False, bf16_points=False):
    """x (B,1,60,T) -> (feat (B,enc_dim), mu (B,nclasses)).  resnet.py:174-191.

    bf16_points=True rounds activations / conv weights to bf16 at exactly the points where the
    sm_100a path stores bf16 (everything else stays fp32): the oracle for kernel-level parity of
    the bf16 tensor-core path; bf16_points=False is the reference's fp32 arithmetic."""
    upd = update_running
    q = _q_bf16 if bf16_points else _q_none
    x = q(F.conv2d(q(x), sd["conv1.weight"], stride=(3, 1), padding=(1, 1)))  # :131,176
    x = q(F.relu(_bn(sd, "bn1", x, training, upd)))
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):                      # :135-138
        x = _preact_block(sd, "layer%d.0" % li, x, stride, training, upd, q)
        x = _preact_block(sd, "layer%d.1" % li, x, 1, training, upd, q)
    x = q(F.conv2d(x, q(sd["conv5.weight"]), stride=1, padding=(0, 1)))       # :140
    x = q(F.relu(_bn(sd, "bn5", x, training, upd))).squeeze(2)               # (B,256,T')
    stats = self_attention_pool(x.permute(0, 2, 1), sd["attention.att_weights"], noise)
    feat = F.linear(stats, sd["fc.weight"], sd["fc.bias"])
    mu = F.linear(feat, sd["fc_mu.weight"], sd["fc_mu.bias"])
    return feat, mu


# ------------------------------------------------------------------------------------
# ECAPA-TDNN (Res2Net2 / Bottle2neck / SEModule)                 ecapa_tdnn.py:15-198
# NB the order in this file is conv -> ReLU -> BN everywhere.
# ------------------------------------------------------------------------------------
def _conv1d(sd, p, x, q=_q_none, **kw):
    return F.conv1d(x, q(sd[p + ".weight"]), sd[p + ".bias"], **kw)


def _se(sd, p, x, training, upd):
    """ecapa_tdnn.py:15-29: mean_T -> conv 512->128 -> ReLU -> BN -> conv 128->512 -> sigmoid (fp32 throughout)."""
    s = x.mean(dim=2, keepdim=True)
    s = F.relu(_conv1d(sd, p + ".se.1", s))
    s = _bn(sd, p + ".se.3", s, training, upd)
    s = torch.sigmoid(_conv1d(sd, p + ".se.4", s))
    return x * s


def _bottle2neck(sd, p, x, dilation, scale, training, upd, q=_q_none):
    """ecapa_tdnn.py:64-95.  q marks the points where the bf16 tensor-core path stores bf16."""
    out = q(_bn(sd, p + ".bn1", q(F.relu(_conv1d(sd, p + ".conv1", x, q))), training, upd))
    width = out.shape[1] // scale
    spx = torch.split(out, width, 1)
    outs = []
    sp = None
    for i in range(scale - 1):
        sp = spx[i] if i == 0 else q(sp + spx[i])
        sp = q(F.relu(_conv1d(sd, p + ".convs.%d" % i, sp, q, dilation=dilation, padding=dilation)))
        sp = q(_bn(sd, p + ".bns.%d" % i, sp, training, upd))
        outs.append(sp)
    outs.append(spx[scale - 1])
    out = torch.cat(outs, 1)
    out = q(_bn(sd, p + ".bn3", q(F.relu(_conv1d(sd, p + ".conv3", out, q))), training, upd))
    out = _se(sd, p + ".se", out, training, upd)
    return q(out + x)


def ecapa_forward(sd, x, training=True, update_running=False, scale=8, bf16_points=False):
    """x (B,n_mels,T) -> (feat (B,256), logits (B,nOut)).  ecapa_tdnn.py:152-198
    (encoder_type='ECA', context=True, summed=False, out_bn=True: the main_train.py:167 config).
    bf16_points: as in resnet_forward (rounding where the sm_100a path stores bf16)."""
    upd = update_running
    q = _q_bf16 if bf16_points else _q_none
    x = q(_bn(sd, "bn1", q(F.relu(_conv1d(sd, "conv1", q(x), q, padding=2))), training, upd))
    x1 = _bottle2neck(sd, "layer1", x, 2, scale, training, upd, q)
    x2 = _bottle2neck(sd, "layer2", x1, 3, scale, training, upd, q)
    x3 = _bottle2neck(sd, "layer3", x2, 4, scale, training, upd, q)
    x = q(F.relu(_conv1d(sd, "layer4", torch.cat((x1, x2, x3), dim=1), q)))
    t = x.shape[-1]
    if bf16_points:
        # the x-block of attention.0 runs on the tensor cores (bf16 weights); the time-constant mean / std
        # blocks are folded into an fp32 per-utterance bias
        c3 = x.shape[1]
        w0, b0 = sd["attention.0.weight"], sd["attention.0.bias"]
        mean = x.mean(dim=2)
        std = torch.sqrt(x.var(dim=2).clamp(min=1e-4))
        u = b0 + mean @ w0[:, c3:2 * c3, 0].t() + std @ w0[:, 2 * c3:, 0].t()
        w = q(F.relu(F.conv1d(x, q(w0[:, :c3])) + u.unsqueeze(2)))
    else:
        gx = torch.cat((x, x.mean(dim=2, keepdim=True).repeat(1, 1, t),
                        torch.sqrt(x.var(dim=2, keepdim=True).clamp(min=1e-4)).repeat(1, 1, t)), dim=1)
        w = F.relu(_conv1d(sd, "attention.0", gx))
    w = q(_bn(sd, "attention.2", w, training, upd))
    w = F.softmax(q(_conv1d(sd, "attention.3", w, q)), dim=2)
    mu = torch.sum(x * w, dim=2)
    sg = torch.sqrt((torch.sum((x ** 2) * w, dim=2) - mu ** 2).clamp(min=1e-4))
    x = torch.cat((mu, sg), 1)
    x = _bn(sd, "bn5", x, training, upd)
    feat = F.linear(x, sd["fc6.weight"], sd["fc6.bias"])
    x = F.linear(feat, sd["fc7.weight"], sd["fc7.bias"])
    x = _bn(sd, "bn7", x, training, upd)
    return feat, x


# ------------------------------------------------------------------------------------
# OC-Softmax == AngularIsoLoss                                   loss.py:187-206 == :73-97
# ------------------------------------------------------------------------------------
def ocsoftmax(center, x, labels, r_real=0.9, r_fake=0.5, alpha=20.0):
    """-> (loss, -cos).  softplus has beta=1, threshold=20 (nn.Softplus defaults, loss.py:185)."""
    w = F.normalize(center, p=2, dim=1)
    xn = F.normalize(x, p=2, dim=1)
    s = xn @ w.t()                                           # (B,1)
    out_scores = s.clone()
    m = torch.where((labels == 0).unsqueeze(1), r_real - s, s)
    m = torch.where((labels == 1).unsqueeze(1), s - r_fake, m)
    loss = F.softplus(alpha * m).mean()
    return loss, -out_scores.squeeze(1)


def cross_entropy(logits, labels):
    """nn.CrossEntropyLoss (logged only when --add_loss is set; main_train.py:250-251,355-357)."""
    return F.cross_entropy(logits, labels)


# ------------------------------------------------------------------------------------
# optimiser steps                                                main_train.py:175,272,404-409
# ------------------------------------------------------------------------------------
def adam_l2_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=5e-4):
    """torch.optim.Adam semantics with coupled L2 weight decay; in place; `step` is 1-based."""
    g = g + weight_decay * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def sgd_step(p, g, lr):
    p.add_(g, alpha=-lr)


def lr_at_epoch(lr, epoch, lr_decay=0.5, interval=30):
    """main_train.py:144-147."""
    return lr * (lr_decay ** (epoch // interval))
