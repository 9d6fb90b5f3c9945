"""Generates tests/golden/elbo_small.npz: inputs + float64 oracle outputs for one small case.

The reference (JAX) cannot run in this image, so these vectors come from the oracle's
line-by-line autograd restatement (`oracle/elbo_ref.py`), cross-checked against the closed-form
one.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import elbo_closed, elbo_ref  # noqa: E402
from oracle.elbo_ref import Noise  # noqa: E402
from tests.helpers import small_problem  # noqa: E402


def main():
    N, G, Ks, nb, B, S = 64, 48, (8, 4, 2), 2, 24, 3
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=11)
    idx = np.random.RandomState(0).permutation(N)[:B]
    noise = Noise.draw(S, B, nb, G, Ks, seed=13, dtype=torch.float32)
    sg = [0.0, 0.0, 0.0]
    args = (1.25, sg, 0.0, 0.0, model.alpha)
    loss, gl = elbo_ref.local_loss_grad(model, X[idx], idx, lp, gp, noise, *args)
    _, gg = elbo_ref.global_loss_grad(model, X[idx], idx, lp, gp, noise, *args)
    loss2, cl, cg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, *args)
    assert abs(float(loss) - loss2) < 1e-9 * abs(loss2)
    out = dict(X=X, batch=batch, idx=idx, layer_sizes=np.array(Ks), annealing=1.25, alpha=model.alpha,
               loss=float(loss))
    for i, p in enumerate(lp):
        out[f"lp{i}"] = p
        out[f"gl{i}"] = gl[i].numpy()
    for i, p in enumerate(gp):
        out[f"gp{i}"] = p
        out[f"gg{i}"] = gg[i].numpy()
    for k in ("c", "s", "tau", "omega", "a", "z"):
        out[f"eps_{k}"] = getattr(noise, k).numpy()
    for l, w in enumerate(noise.W):
        out[f"eps_W{l}"] = w.numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "elbo_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss))


if __name__ == "__main__":
    main()
