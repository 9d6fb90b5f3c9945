import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line(
        "markers",
        "gpu_pending: GPU test of code that has not had its first run on a device yet (none at the moment: the 16 of "
        "round 1 ran green in round 2's first GPU call and are ordinary `gpu` tests now). Not selected by `-m gpu`.")


def pytest_sessionstart(session):
    """A fresh checkout has no libscdef_b200.so (it is git-ignored): build it once if nvcc is here, so that the
    ABI tests (`-m "not gpu"`) and the GPU tests do not depend on the order in which the driver runs things."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "scdef_b200", "libscdef_b200.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.call(["bash", os.path.join(ROOT, "scdef_b200", "csrc", "build.sh")])
