"""CPU restatement of the reference's adversarial channel-classifier head -- TEST INFRASTRUCTURE ONLY.

Follows model.py:976-1023 (GradientReversalFunction, GradientReversal, ChannelClassifier) and the way
main_train.py:377-403,420-453 uses it with `criterion = nn.CrossEntropyLoss()` (main_train.py:251):

    x -> GRL(lambda) -> Linear(enc, enc/2) -> Dropout(0.3) -> ReLU -> Linear(enc/2, C) -> ReLU -> CE(mean)

forward_backward() returns the loss, the predicted classes (torch.max: first maximum), the gradient that reaches the
features THROUGH the gradient-reversal layer (-lambda * dL/dx) and the classifier's parameter gradients, for an explicit
dropout keep-mask (the reference draws it from torch's RNG; parity is defined for a given mask).
Pinned by tests/golden/adv_golden.npz (reference module run with the same mask injected into F.dropout).
"""
import numpy as np


def forward_backward(x, labels, w1, b1, w2, b2, keep_mask, lambda_, p=0.3):
    x = np.asarray(x, dtype=np.float64)
    keep = np.asarray(keep_mask, dtype=np.float64)
    B = x.shape[0]
    h = x @ w1.T.astype(np.float64) + b1                       # Linear 0
    hd = h * keep / (1.0 - p)                                  # Dropout (training)
    a = np.maximum(hd, 0.0)                                    # ReLU
    z = a @ w2.T.astype(np.float64) + b2                       # Linear 3
    logits = np.maximum(z, 0.0)                                # trailing ReLU (model.py:1012)
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    sm = e / e.sum(axis=1, keepdims=True)
    loss = float(np.mean(-np.log(sm[np.arange(B), labels])))
    pred = logits.argmax(axis=1)
    dlogits = sm.copy()
    dlogits[np.arange(B), labels] -= 1.0
    dlogits /= B
    dz = dlogits * (z > 0)
    dw2, db2 = dz.T @ a, dz.sum(axis=0)
    da = dz @ w2.astype(np.float64)
    dh = da * (hd > 0) * keep / (1.0 - p)
    dw1, db1 = dh.T @ x, dh.sum(axis=0)
    dx = dh @ w1.astype(np.float64)
    return {"loss": loss, "pred": pred, "logits": logits, "dfeat": -lambda_ * dx, "dw1": dw1, "db1": db1, "dw2": dw2, "db2": db2}
