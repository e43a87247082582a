"""Tolerances used by the parity tests, stated once.

LFCC (fp32): north_star asks for "within 1e-4 rel".  Cepstra are log-energies that cross zero,
so the criterion is written with an absolute floor: |a - b| <= 1e-4 * (|b| + 1).  (The reference's
own fp32 output differs from a float64 restatement by up to 7e-6 absolute)
Training loss / logits / scores: 1e-3 rel for the fp32 oracle restatement; the bf16 tensor-core
path is checked against an oracle that applies the same bf16 rounding points, and against the
fp32 golden values with the bf16 tolerance stated in the individual tests.
"""
import numpy as np

LFCC_RTOL = 1e-4


def lfcc_close(a, b, rtol=LFCC_RTOL):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= rtol * (np.abs(b) + 1.0)


def lfcc_worst(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1.0)))
