"""GPU parity against the (reference-pinned) float64 oracle AT THE BASELINE PER-STEP SHAPES, and in the clip regimes.

Shapes (BASELINE.json configs; the kernels' work per step does not depend on the number of cells beyond the
minibatch, so N is kept small): C5 atlas step (5000 genes, layers [100,60,30,10], 256 cells), C4 (the same with 8
batches), C2 (2000 genes; also with the library's default ladder [100,40,16,6,3,1]), C3 (iscDEF-shaped: 3000 genes,
layers [48,24,1], matrix-valued W_0 / W_1 priors).  Ten MC samples (the paper-benchmark setting,
`/root/reference/reproducibility/config/methods.yaml:10-11`): the oracle evaluates those in a few seconds.

Clip regimes (`_scdef.py:1835-1838, 1855-1911`, `jax_utils.py:6-8`): mu outside [ln 1e-10, ln 1e6], rho outside
[ln 1e-10, ln 1e10], samples that hit the 1e-15 / 1e15 clip: value AND the zeroed gradients must match the oracle.

Bar: 1e-4 on the loss and on max|g - g_ref| / max|g_ref| per gradient tensor (BASELINE.json north_star).
"""
import numpy as np
import pytest
import torch

from oracle import elbo_closed
from oracle.elbo_ref import Noise
from tests.helpers import make_engine, noise_to_spec, rel_err, small_problem

pytestmark = pytest.mark.gpu
TOL = 1e-4
NAMES_L = ("cell_scale", "z")


def _check(model, X, batch, lp, gp, idx, noise, sg, anneal=1.2, lik_impl=0, tol=TOL):
    Ks = model.layer_sizes
    S = noise.num_samples
    ref_loss, ref_gl, ref_gg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, anneal, sg, 0.0, 0.0, model.alpha)
    assert np.isfinite(ref_loss)
    eng = make_engine(X, batch, model, lp, gp)
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    kw = dict(num_samples=S, annealing=anneal, stop_gradients=sg, alpha=model.alpha, lik_impl=lik_impl)
    spec = noise_to_spec(noise)
    errs = {}
    eng.elbo_grad("local", idx_d, spec, **kw)
    errs["loss_local"] = abs(eng.loss_value() - ref_loss) / abs(ref_loss)
    gl = [g.cpu().numpy() for g in eng.local_grads(len(idx))]
    eng.elbo_grad("global", idx_d, spec, **kw)
    errs["loss_global"] = abs(eng.loss_value() - ref_loss) / abs(ref_loss)
    gg = [g.cpu().numpy() for g in eng.glob_grad.views]
    names_g = ["gene_scale", "brd"] + [f"W{l}" for l in range(len(Ks))] + ["ard", "alpha"]
    for name, g, r in list(zip(NAMES_L, gl, [r[:, idx] for r in ref_gl])) + list(zip(names_g, gg, ref_gg)):
        assert np.all(np.isfinite(g)), name
        if np.abs(r).max() == 0:
            assert np.abs(g).max() == 0, name
        else:
            errs["g_" + name] = rel_err(g, r)
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"parity failures {bad} (all: {errs})"
    return errs, (ref_gl, ref_gg, gl, gg)


@pytest.mark.parametrize("name,G,Ks,nb,B", [
    ("C5 step", 5000, (100, 60, 30, 10), 1, 256),
    ("C4 step (8 batches)", 5000, (100, 60, 30, 10), 8, 256),
    ("C2 step", 2000, (100, 60, 30, 10), 1, 256),
    ("C2 step, default ladder", 2000, (100, 40, 16, 6, 3, 1), 1, 256),
    ("C5 step, short last minibatch", 5000, (100, 60, 30, 10), 1, 64),
])
def test_baseline_step_shapes_match_the_oracle(name, G, Ks, nb, B):
    N, S = 640, 10
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=21)
    idx = np.random.RandomState(3).permutation(N)[:B]
    noise = Noise.draw(S, B, nb, G, Ks, seed=5, dtype=torch.float32, couple_like_reference=True)
    _check(model, X, batch, lp, gp, idx, noise, [0] * len(Ks))


def test_c3_iscdef_step_shape_matches_the_oracle():
    """100k x 3000 iscDEF config: layers [48,24,1] (24 marker sets x 2 sub-factors, root), W_0 prior per (factor, gene)
    from marker lists, block connectivity prior on W_1, root frozen as in the main phase of `iscDEF.fit`."""
    from oracle import host_ref

    N, G, B, S = 640, 3000, 256, 10
    Ks = (48, 24, 1)
    X, batch, model, lp, gp = small_problem(N, G, Ks, 1, seed=23)
    rng = np.random.default_rng(1)
    genes = [f"g{i}" for i in range(G)]
    markers = {f"t{i}": [genes[j] for j in rng.choice(G, 10, replace=False)] for i in range(20)}
    names = list(markers) + [f"other{i}" for i in range(4)]
    P0, R0 = host_ref.iscdef_geneset_prior(genes, markers, names, list(Ks), 1, 2, gs_small_scale=1.0, gs_big_scale=10.0,
                                           marker_strength=1.0, nonmarker_strength=1.0, other_strength=0.1)
    model.w_priors[0] = [P0, R0]
    conn = host_ref.iscdef_connectivity_prior(list(Ks), 24, 1, True, 2, 1.0, 10.0, 0.1, 1.0)
    for l, (sh, rt) in conn.items():
        model.w_priors[l] = [np.asarray(model.w_priors[l][0]) * sh, np.asarray(model.w_priors[l][1]) * rt]
    idx = np.random.RandomState(3).permutation(N)[:B]
    noise = Noise.draw(S, B, 1, G, Ks, seed=5, dtype=torch.float32, couple_like_reference=True)
    _check(model, X, batch, lp, gp, idx, noise, [0, 0, 1])


def _clip_problem(seed=31, G=136, Ks=(24, 6, 1), nb=2, N=300, B=200, S=3):
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=seed)
    idx = np.random.RandomState(seed).permutation(N)[:B]
    noise = Noise.draw(S, B, nb, G, Ks, seed=seed + 1, dtype=torch.float32)
    return X, batch, model, [np.array(p) for p in lp], [np.array(p) for p in gp], idx, noise


@pytest.mark.parametrize("lik_impl", [1, 2])
def test_parameters_outside_their_clips(lik_impl):
    """mu below ln 1e-10 (and above ln 1e6 where the model stays finite: ard, W_1) and rho below ln 1e-10 for entries
    of every block: the clipped values enter the objective, the clipped parameter gets exactly zero gradient
    (`jnp.clip`'s sub-gradient), its partner keeps its own.  (mu of the gene scale is only bounded below, `_scdef.py:1872`.)"""
    X, batch, model, lp, gp, idx, noise = _clip_problem()
    L = model.n_layers
    rng = np.random.default_rng(0)

    def poke(a, n, hi_mu=None, lo_mu=-30.0):
        mu, rho = a[0].reshape(-1), a[1].reshape(-1)     # views: a[0] / a[1] are contiguous
        pick = rng.choice(mu.size, size=min(n, mu.size), replace=False)
        third = max(1, len(pick) // 3)
        if lo_mu is not None:
            mu[pick[:third]] = lo_mu
        if hi_mu is not None:
            mu[pick[third:2 * third]] = hi_mu
        rho[pick[2 * third:]] = -30.0

    rows = idx[:40]
    K0 = model.layer_sizes[0]   # z: the bottom layer only (a vanishing upper-layer z makes the prior rate alpha / m of the
    zsub, csub = np.ascontiguousarray(lp[1][:, rows, :K0]), np.ascontiguousarray(lp[0][:, rows])   # layer below astronomic)
    poke(zsub, 60)
    poke(csub, 12)
    lp[1][:, rows, :K0], lp[0][:, rows] = zsub, csub
    poke(gp[0], 20)                    # gene scale
    poke(gp[2], 200)                   # W_0
    poke(gp[3], 20, hi_mu=30.0)        # W_1: both sides
    poke(gp[2 + L], 6, hi_mu=30.0, lo_mu=None)   # ard: upper side only (a vanishing ard makes the W_0 prior rate astronomic)
    errs, (ref_gl, ref_gg, gl, gg) = _check(model, X, batch, lp, gp, idx, noise, [0] * L, lik_impl=lik_impl)
    for blk in (2, 3, 2 + L):
        mu_out = (gp[blk][0] < np.log(1e-10)) | (gp[blk][0] > np.log(1e6))
        rho_out = gp[blk][1] < np.log(1e-10)
        assert mu_out.any() and rho_out.any()
        assert np.all(gg[blk][0][mu_out] == 0.0) and np.all(ref_gg[blk][0][mu_out] == 0.0)
        assert np.all(gg[blk][1][rho_out] == 0.0) and np.all(ref_gg[blk][1][rho_out] == 0.0)


@pytest.mark.parametrize("lik_impl", [1, 2])
def test_samples_that_hit_the_sample_clip(lik_impl):
    """Large sigma on part of W_0, z (means near the lower bound) and ard (means near the upper bound):
    exp(mu + sigma eps) leaves [1e-15, 1e15] for many draws, `lognormal_sample` clips them (`jax_utils.py:6-8`) and
    those draws carry no gradient through the sample."""
    X, batch, model, lp, gp, idx, noise = _clip_problem(seed=37)
    L = model.n_layers
    rng = np.random.default_rng(1)
    w0 = gp[2]
    pick = rng.random(w0[0].shape) < 0.05
    w0[0][pick] = -22.0                       # mean near the lower parameter clip ...
    w0[1][pick] = 2.3                         # ... sigma ~ 10: one draw in ten falls below 1e-15
    rows = idx[:30]
    K0 = model.layer_sizes[0]
    zsub = np.ascontiguousarray(lp[1][:, rows, :K0])
    zpick = rng.random(zsub[0].shape) < 0.1
    zsub[0][zpick] = -22.0
    zsub[1][zpick] = 2.3
    lp[1][:, rows, :K0] = zsub
    om = gp[2 + L]
    om[0][:4] = 13.0                          # ard near ln 1e6 with sigma ~ 12 and draws pushed upwards: above 1e15
    om[1][:4] = 2.5                           # (a vanishing ard would make the W_0 prior rate astronomic instead)
    noise.omega[:, :4] = noise.omega[:, :4].abs() + 1.0
    errs, (ref_gl, ref_gg, gl, gg) = _check(model, X, batch, lp, gp, idx, noise, [0] * L, lik_impl=lik_impl)
    raw = w0[0][None] + np.exp(w0[1])[None] * noise.W[0].numpy()
    assert (raw[:, pick] < np.log(1e-15)).mean() > 0.05           # the lower clip really was hit ...
    raw_om = om[0][None] + np.exp(om[1])[None] * noise.omega.numpy()
    assert (raw_om[:, :4] > np.log(1e15)).any()                    # ... and the upper one
