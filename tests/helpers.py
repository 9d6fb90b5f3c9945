"""Shared builders for the tests: a small problem, its oracle model and a matching Engine."""
import numpy as np
import torch

from oracle import host_ref
from oracle.elbo_ref import Noise
from scdef_b200.synth import synth_counts_numpy


def small_problem(N=96, G=70, Ks=(9, 5, 2), nb=1, seed=0, use_brd=True, marginalize_alpha=False,
                  perturb=0.15, **hyper):
    X, batch = synth_counts_numpy(N, G, n_batches=nb, seed=seed, lib=800.0)
    model = host_ref.build_model(X, list(Ks), batch_ids=batch if nb > 1 else None, use_brd=use_brd,
                                 marginalize_alpha=marginalize_alpha, **hyper)
    lp, gp = host_ref.init_var_params(model, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    lp = [p + perturb * rng.normal(size=p.shape).astype(np.float32) for p in lp]
    gp = [p + perturb * rng.normal(size=p.shape).astype(np.float32) for p in gp]
    return X, batch, model, lp, gp


def make_engine(X, batch, model, lp, gp, resident=True, device="cuda:0", **kw):
    from scdef_b200.engine import Engine

    eng = Engine(
        layer_sizes=model.layer_sizes, n_genes=X.shape[1], n_batches=model.batch_indices_onehot.shape[1],
        n_cells=X.shape[0], batch_id=batch, batch_lib_ratio=model.batch_lib_ratio,
        cell_alpha_factor=model.cell_alpha_factor, gene_ratio=model.gene_ratio, w_priors=model.w_priors,
        use_brd=model.use_brd, marginalize_alpha=model.marginalize_alpha, top_alpha=model.top_alpha,
        brd=model.brd, brd_mean=model.brd_mean, shrinkage_shape=model.shrinkage_shape,
        shrinkage_rate=model.shrinkage_rate, cell_scale_shape=model.cell_scale_shape,
        gene_scale_shape=model.gene_scale_shape, alpha_floor=model.alpha_floor,
        supervised_layer=model.supervised_layer, fixed_top_z=model.fixed_top_z,
        X=X if resident else None, device=device, **kw)
    eng.set_params(lp, gp)
    return eng


def noise_to_spec(noise: Noise):
    from scdef_b200.engine import NoiseSpec

    return NoiseSpec.explicit(c=noise.c, z=noise.z, s=noise.s, tau=noise.tau, omega=noise.omega,
                              a=noise.a, W=noise.W)


def rel_err(a, b):
    """max|a-b| / max|b|  (SURVEY.md §8c parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / den)
