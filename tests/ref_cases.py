"""Cases shared by `tests/golden/make_golden_ref.py` (runs the UNMODIFIED reference source through
`oracle/ref_exec.py` and stores what it returned) and `tests/test_ref_pin.py` (requires the restated
oracle to reproduce that execution).  Test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch

from oracle import jaxshim
from oracle.elbo_ref import Noise, OracleModel
from scdef_b200.synth import synth_counts_numpy

# name -> problem definition.  Shapes are chosen so that the reference's same-key coupling shows up
# (tau/omega always; c and a width-1 root z; W_l / z_l coincidences), nb = 1 and > 1, BRD on/off,
# q(alpha) on/off, every stop-gradient phase `fit` uses, annealing != 1, a short last minibatch.
ELBO_CASES = {
    "base":        dict(N=96, G=70, Ks=[9, 5, 2], nb=1, B=32, S=3, sg=[0, 0, 0]),
    "two_batches": dict(N=96, G=70, Ks=[9, 5, 2], nb=2, B=32, S=2, sg=[0, 0, 0], anneal=2.5),
    "no_brd":      dict(N=80, G=50, Ks=[8, 4], nb=1, B=80, S=2, sg=[0, 0], use_brd=False),
    "marg_alpha":  dict(N=96, G=70, Ks=[9, 5, 1], nb=3, B=40, S=2, sg=[0, 0, 0], marginalize_alpha=True),
    "root_frozen": dict(N=96, G=70, Ks=[9, 5, 2], nb=1, B=32, S=2, sg=[0, 0, 1]),
    "root_only":   dict(N=96, G=70, Ks=[9, 5, 1], nb=2, B=32, S=2, sg=[1, 1, 0]),
    "budgets_off": dict(N=96, G=70, Ks=[9, 5, 2], nb=2, B=32, S=2, sg=[0, 0, 0], stop_cell=1.0, stop_gene=1.0),
    "coupled":     dict(N=64, G=6, Ks=[6, 6, 1], nb=1, B=6, S=2, sg=[0, 0, 0]),  # W_1 (6,6) = z_0/z_1 (B,6) = W_0 (6,G)
    "five_layers": dict(N=120, G=90, Ks=[12, 8, 5, 3, 1], nb=1, B=50, S=2, sg=[0, 0, 0, 0, 0], top_alpha=3.0),
    "hyper":       dict(N=96, G=70, Ks=[9, 5, 2], nb=2, B=32, S=2, sg=[0, 1, 0], brd_strength=2.0, brd_mean=0.5,
                        shrinkage_shape=2.0, shrinkage_rate=0.5, cell_scale_shape=2.0, gene_scale_shape=0.7,
                        factor_shape=0.3, anneal=0.5),
}


def make_counts(case, seed=0):
    X, batch = synth_counts_numpy(case["N"], case["G"], n_batches=case["nb"], seed=seed, lib=800.0)
    return np.asarray(X, dtype=np.float64), np.asarray(batch)


def ctor_kwargs(case):
    kw = dict(layer_sizes=list(case["Ks"]), seed=1)
    for k in ("use_brd", "marginalize_alpha", "top_alpha", "brd_strength", "brd_mean", "shrinkage_shape",
              "shrinkage_rate", "cell_scale_shape", "gene_scale_shape", "factor_shape"):
        if k in case:
            kw[k] = case[k]
    return kw


def build_reference_model(ref, case, seed=0, cls=None):
    X, batch = make_counts(case, seed)
    ad = ref.adata(X, batch if case["nb"] > 1 else None)
    kw = ctor_kwargs(case)
    if case["nb"] > 1:
        kw["batch_key"] = "batch"
    m = (cls or ref.scDEF)(ad, **kw)
    m.elbos, m.step_sizes = [], []  # `fit` does this before `_learn` (`_scdef.py:2638-2639`)
    return m, X, batch


def perturb(params, seed, scale=0.15):
    """Moves every variational parameter off its initial value (so that no term is accidentally zero)."""
    rng = np.random.default_rng(seed)
    return [np.asarray(p, dtype=np.float64) + scale * rng.normal(size=tuple(p.shape)) for p in params]


def oracle_model_from_reference(m, alpha_floor=0.1, supervised_layer=None, fixed_top_z=None) -> OracleModel:
    """The attributes `elbo` reads, taken from the reference object itself."""
    as_np = lambda v: np.asarray(v, dtype=np.float64)
    w_priors = [[as_np(a) if np.ndim(a) else float(a), as_np(b) if np.ndim(b) else float(b)] for a, b in m.w_priors]
    gr = m.gene_ratio
    om = OracleModel(
        n_cells=int(m.X.shape[0]), layer_sizes=[int(k) for k in m.layer_sizes],
        batch_indices_onehot=as_np(m.batch_indices_onehot), batch_lib_ratio=as_np(m.batch_lib_ratio),
        gene_ratio=as_np(gr) if np.ndim(gr) else float(gr), cell_alpha_factor=as_np(m.cell_alpha_factor),
        w_priors=w_priors, use_brd=bool(m.use_brd), marginalize_alpha=bool(m.marginalize_alpha),
        top_alpha=float(m.top_alpha), brd=float(m.brd), brd_mean=float(m.brd_mean),
        shrinkage_shape=float(m.shrinkage_shape), shrinkage_rate=float(m.shrinkage_rate),
        cell_scale_shape=float(m.cell_scale_shape), gene_scale_shape=float(m.gene_scale_shape),
        alpha_floor=alpha_floor, supervised_layer=supervised_layer, fixed_top_z=fixed_top_z)
    om.alpha = float(m.alpha)
    return om


def noise_from_key(ref_random, key, S, B, nb, G, Ks, dtype=torch.float64) -> Noise:
    """The N(0,1) draws the reference's `elbo` consumes under the stand-in PRNG, block by block:
    `rngs = random.split(key, S)` (`_scdef.py:2163`), then every `lognormal_sample(rngs[i], m, s)` draws
    `normal(rngs[i], broadcast shape)` (`jax_utils.py:6-8`) — identical shapes give identical draws."""
    Ks = [int(k) for k in Ks]
    rngs = ref_random.split(key, S)
    outs = [G] + Ks[:-1]
    draw = lambda i, shape: jaxshim.normal_like_reference(rngs[i], shape, dtype)
    stack = lambda shape: torch.stack([draw(i, shape) for i in range(S)])
    z = torch.cat([stack((B, k)) for k in Ks], dim=2)
    return Noise(c=stack((B, 1)), s=stack((nb, G)), tau=stack((Ks[0], 1)), omega=stack((Ks[0], 1)),
                 a=stack((1, 1)), W=[stack((Ks[l], outs[l])) for l in range(len(Ks))], z=z)


def noise_arrays(noise: Noise):
    d = dict(eps_c=noise.c, eps_s=noise.s, eps_tau=noise.tau, eps_omega=noise.omega, eps_a=noise.a, eps_z=noise.z)
    for l, w in enumerate(noise.W):
        d[f"eps_W{l}"] = w
    return {k: np.asarray(v, dtype=np.float64) for k, v in d.items()}


def noise_from_arrays(d, n_layers, prefix="", dtype=torch.float64) -> Noise:
    t = lambda k: torch.as_tensor(np.asarray(d[prefix + k]), dtype=dtype)
    return Noise(c=t("eps_c"), s=t("eps_s"), tau=t("eps_tau"), omega=t("eps_omega"), a=t("eps_a"),
                 W=[t(f"eps_W{l}") for l in range(n_layers)], z=t("eps_z"))
