"""Isolate the stages at the top of the ResNet backward in fp32 mode on REAL data: each stage's output from the engine is
compared with an fp64 torch evaluation of the same stage on the engine's own inputs."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import state_spec as ss
from asvspoof2021_air_b200 import ops
from asvspoof2021_air_b200.trainer import Trainer

def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-300))

B = 4
spec = ss.resnet_spec()
waves, labels = ss.seeded_waves(B, 64000, seed=3), ss.seeded_labels(B, 3)
tr = Trainer(arch="resnet", seed=5, precision=sys.argv[1] if len(sys.argv) > 1 else "fp32")
tr.load_state(ss.seeded_state(spec, 11), ss.seeded_center(256, 11))
eng = tr.engine
x0 = tr.features(waves.cuda())
feat, logits = eng.forward(x0, training=True)
dfeat, score = torch.empty_like(feat), torch.empty(B, device="cuda")
eng.zero_grad(); tr.center_grad.zero_()
ops.ocsoftmax(feat, labels.cuda(), tr.center, B, 256, 0.9, 0.2, 20.0, 1.0, tr.loss, score, dfeat, tr.center_grad, logits, 2, tr.ce)
eng.backward(dfeat)
torch.cuda.synchronize()
st = eng.store
last = eng.blocks[-1]
# --- bn5 backward: inputs g_z5 (dL/dz5), c5; z5 = relu(bn(c5))
c5 = eng.c5.double().reshape(-1, 256); gz = eng.g_z5.double().reshape(-1, 256)
g, b = st.view("bn5.weight").double(), st.view("bn5.bias").double()
c5r = c5.clone().requires_grad_(True)
mu, var = c5r.mean(0), c5r.var(0, unbiased=False)
z = F.relu((c5r - mu) / torch.sqrt(var + 1e-5) * g + b)
(z * gz).sum().backward()
print("bn5 bwd dx           rel %.2e   (|dx| %.3e)" % (rel(eng.g_c5.reshape(-1, 256), c5r.grad), float(c5r.grad.norm())))
# --- conv5 wgrad on the engine's own (y, g_c5)
y = last.y.double().permute(0, 3, 1, 2); gc5 = eng.g_c5.double().reshape(B, 1, -1, 256).permute(0, 3, 1, 2)
wref = torch.nn.grad.conv2d_weight(y, (256, 512, 3, 3), gc5, padding=(0, 1))
print("conv5 wgrad          rel %.2e" % rel(st.pt_view("conv5.weight", st.grads), wref))
w5 = st.pt_view("conv5.weight").double()
dref = torch.nn.grad.conv2d_input(y.shape, w5, gc5, padding=(0, 1))
print("conv5 dgrad          rel %.2e" % rel(last.g_y.double().permute(0, 3, 1, 2), dref))
# --- layer4.1: conv2 wgrad / dgrad, bn2 bwd, conv1 wgrad / dgrad
gy = last.g_y.double().permute(0, 3, 1, 2); a2 = last.a2.double().permute(0, 3, 1, 2)
print("l4.1 conv2 wgrad     rel %.2e" % rel(st.pt_view("layer4.1.conv2.weight", st.grads),
      torch.nn.grad.conv2d_weight(a2, (512, 512, 3, 3), gy, padding=1)))
print("l4.1 conv2 dgrad     rel %.2e" % rel(last.g_a2.double().permute(0, 3, 1, 2),
      torch.nn.grad.conv2d_input(a2.shape, st.pt_view("layer4.1.conv2.weight").double(), gy, padding=1)))
h = last.h.double().reshape(-1, 512); ga2 = last.g_a2.double().reshape(-1, 512)
g, b = st.view("layer4.1.bn2.weight").double(), st.view("layer4.1.bn2.bias").double()
hr = h.clone().requires_grad_(True)
mu, var = hr.mean(0), hr.var(0, unbiased=False)
(F.relu((hr - mu) / torch.sqrt(var + 1e-5) * g + b) * ga2).sum().backward()
print("l4.1 bn2 bwd dx      rel %.2e" % rel(last.g_h.reshape(-1, 512), hr.grad))
# forward sanity: a2 == relu(bn(h))
print("l4.1 bn2 fwd         rel %.2e" % rel(last.a2.reshape(-1, 512), F.relu((h - h.mean(0)) / torch.sqrt(h.var(0, unbiased=False) + 1e-5) * g + b)))
# --- head: OC-Softmax dfeat, fc backward, pooling backward on the engine's own tensors
from oracle import nets_oracle as no
featr = eng.feat.double().clone().requires_grad_(True)
loss, _ = no.ocsoftmax(tr.center.double(), featr, labels.cuda(), 0.9, 0.2, 20.0)
loss.backward()
print("ocsoftmax dfeat      rel %.2e" % rel(dfeat, featr.grad))
W, bfc = st.view("fc.weight").double(), st.view("fc.bias").double()
print("fc bwd g_stats       rel %.2e" % rel(eng.g_stats, dfeat.double() @ W))
z5 = eng.z5.double().reshape(B, -1, 256).clone().requires_grad_(True)
pooled = no.self_attention_pool(z5, st.view("attention.att_weights").double())
(pooled * eng.g_stats.double()).sum().backward()
print("pool fwd stats       rel %.2e" % rel(eng.stats, pooled))
print("pool bwd g_z5        rel %.2e" % rel(eng.g_z5.reshape(B, -1, 256), z5.grad))
half = eng.g_stats.double().clone(); half[:, 256:] = 0
z5b = eng.z5.double().reshape(B, -1, 256).clone().requires_grad_(True)
(no.self_attention_pool(z5b, st.view("attention.att_weights").double()) * half).sum().backward()
print("   |g_z5| %.3e, mean-part only |g| %.3e" % (float(z5.grad.norm()), float(z5b.grad.norm())))
sd_ = pooled[:, 256:]
print("   std min %.3e  median %.3e ; dead channels (std == 0): %d" % (float(sd_.min()), float(sd_.median()), int((sd_ == 0).sum())))
