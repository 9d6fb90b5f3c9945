"""(Named test_zz_* so that it runs last: the driver runs the GPU suite with -x.)
CTA-pair LOCAL likelihood kernel (`lik_tc_pair.cuh`, opt-in with SCDEF_TC_PAIR): the bring-up cases of
tests/tc_bringup.py against the float64 oracle, in a subprocess because the switch is read once per process.
Bar: 1e-4 on the loss and on every gradient tensor (same as tests/test_parity_gpu.py).  The persistent variant (=2) is the one tested here; these four cases ran green on a
B200 in round 1 (profiles/r01_pair_flat_bringup.txt)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(mode):
    env = dict(os.environ, SCDEF_TC_PAIR=mode)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tc_bringup.py"), "pair"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "pair: worst relative error" in r.stdout


@pytest.mark.gpu
def test_persistent_pair_kernel_matches_the_oracle():
    _run("2")


@pytest.mark.gpu_pending
@pytest.mark.parametrize("mode", ["3", "4", "5"])
def test_pending_epilogue_variants_match_the_oracle(mode):
    """SCDEF_TC_PAIR=3 (whole-pair-tile epilogue) and 4 (+ per-k-step Q hand-off): written after the round's GPU minutes
    were spent, never run; their barrier protocol is modelled in tests/test_pair_protocol.py."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    _run(mode)
