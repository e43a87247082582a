"""GPU parity of the --ADV_AUG channel-classifier head: golden vectors of the
reference module, the numpy oracle with the mask the kernel drew, mask statistics, the Adam step, and the head inside a
train step.  The checks live in tests/adv_gpu_checks.py and run in a subprocess so that their full output is kept
(scripts/gpu_full.sh runs the same file directly)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_adv_head_against_reference_golden_and_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "adv_gpu_checks.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "adv checks ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
