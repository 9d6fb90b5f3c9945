"""CTA-pair LOCAL likelihood kernel (`lik_tc_pair.cuh`, opt-in with SCDEF_TC_PAIR): the bring-up cases of
tests/tc_bringup.py against the float64 oracle, in a subprocess because the switch is read once per process.
Bar: 1e-4 on the loss and on every gradient tensor (same as tests/test_parity_gpu.py).  The persistent variant (=2) is the one tested here; these four cases ran green on a
B200 in round 1 (profiles/r01_pair_flat_bringup.txt)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_persistent_pair_kernel_matches_the_oracle():
    env = dict(os.environ, SCDEF_TC_PAIR="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tc_bringup.py"), "pair"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "pair: worst relative error" in r.stdout
