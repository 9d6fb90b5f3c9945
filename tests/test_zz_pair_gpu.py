"""(Named test_zz_* so that it runs last: the driver runs the GPU suite with -x.)
The fused likelihood kernels against the float64 oracle on the bring-up cases of tests/tc_bringup.py (one pair-tile, C5-like
layers with a phantom tile, batches with two cell blocks, 100 samples whose chunks cross sample boundaries), in a
subprocess because the kernel switch is read once per process: the default CTA-pair kernel (`csrc/lik_flat.cuh`, both
passes) and the monolithic single-CTA kernels (`SCDEF_TC_MONO=1`, which the likelihood-off phases also use).
Bar: 1e-4 on the loss and on every gradient tensor (same as tests/test_parity_gpu.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mono", [False, True])
def test_likelihood_kernels_match_the_oracle(mono):
    env = dict(os.environ)
    env.pop("SCDEF_TC_MONO", None)
    if mono:
        env["SCDEF_TC_MONO"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tc_bringup.py"), "pair"], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-3000:]
    assert "pair: worst relative error" in r.stdout
