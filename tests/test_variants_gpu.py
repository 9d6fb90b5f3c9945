"""Kernel variants must reproduce each other.  `k_sample_v2` draws the same Philox normals and does the same arithmetic as
`k_sample`, so every sample — and with it the loss and every gradient — must agree to the last bits (1e-6 leaves room
for the float atomics of the kernels downstream)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle.elbo_ref import Noise
from scdef_b200 import _lib
from scdef_b200.engine import NoiseSpec
from tests.helpers import make_engine, noise_to_spec, rel_err, small_problem

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def _set_sample_impl(v):
    lib = _lib.load()
    lib.scdef_debug_set_sample_impl.argtypes = [ctypes.c_int]
    lib.scdef_debug_set_sample_impl.restype = None
    lib.scdef_debug_set_sample_impl(v)


def _eval(eng, idx_d, spec, B, S, sg, model):
    out = []
    for which in ("local", "global"):
        eng.elbo_grad(which, idx_d, spec, num_samples=S, annealing=1.3, stop_gradients=sg, alpha=model.alpha)
        torch.cuda.synchronize()
        grads = eng.local_grads(B) if which == "local" else eng.glob_grad.views
        out.append((eng.loss_value(), [g.cpu().numpy().copy() for g in grads]))
    return out


@pytest.mark.parametrize("N,G,Ks,nb,B,S,explicit", [
    (96, 70, (9, 5, 2), 1, 40, 3, True),
    (300, 200, (100, 60, 30, 10), 1, 256, 7, False),
    (160, 72, (24, 6, 1), 3, 130, 100, False),
])
def test_sample_v2_reproduces_the_validated_kernel(N, G, Ks, nb, B, S, explicit):
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=0)
    eng = make_engine(X, batch, model, lp, gp)
    idx = np.random.RandomState(0).permutation(N)[:B]
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    spec = noise_to_spec(Noise.draw(S, B, nb, G, Ks, seed=3, dtype=torch.float32)) if explicit else NoiseSpec.philox(5, 11)
    sg = [0] * len(Ks)
    try:
        _set_sample_impl(1)
        ref = _eval(eng, idx_d, spec, B, S, sg, model)
        _set_sample_impl(2)
        got = _eval(eng, idx_d, spec, B, S, sg, model)
    finally:
        _set_sample_impl(1)
    for (l0, g0), (l1, g1) in zip(ref, got):
        assert l1 == pytest.approx(l0, rel=1e-9)
        for a, b in zip(g0, g1):
            assert (rel_err(b, a) < 1e-6) if np.abs(a).max() > 0 else (np.abs(b).max() == 0)
