"""GPU parity of the --ADV_AUG channel-classifier head.

The head was written after the round's GPU budget was spent: its arithmetic is pinned on the CPU (oracle vs the
reference module, tests/test_oracle.py::test_adv_classifier_oracle_vs_reference_golden) but these checks have not run
on hardware.  They therefore run in a subprocess (a kernel fault cannot poison the CUDA context of the other GPU tests)
and are xfail(strict=False) until a B200 run confirms them -- then the mark goes and main_train.py stops rejecting
--ADV_AUG."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="adv head not yet validated on hardware (written without GPU time)")
def test_adv_head_against_reference_golden_and_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "adv_gpu_checks.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "adv checks ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
