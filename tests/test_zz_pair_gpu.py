"""(Named test_zz_* so that it runs last: the driver runs the GPU suite with -x.)
CTA-pair LOCAL likelihood kernel (`lik_tc_pair.cuh`, the default LOCAL path; SCDEF_TC_PAIR selects a variant): the bring-up cases of
tests/tc_bringup.py against the float64 oracle, in a subprocess because the switch is read once per process.
Bar: 1e-4 on the loss and on every gradient tensor (same as tests/test_parity_gpu.py)."""
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
@pytest.mark.parametrize("mode", ["3", "2", "0"])
def test_likelihood_kernel_variants_match_the_oracle(mode):
    """3 = the default (persistent CTA pairs, whole-pair-tile epilogue), 2 = its narrow-epilogue sibling, 0 = the
    single-CTA pipeline that odd batch sizes still take.  All green on a B200 (profiles/r02_first_call_summary.txt)."""
    _run(mode)
