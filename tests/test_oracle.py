"""CPU: the oracle pins itself (two independent restatements + finite differences + golden file)."""
import os

import numpy as np
import pytest
import torch

from oracle import elbo_closed, elbo_ref, host_ref
from oracle.elbo_ref import Noise
from tests.helpers import rel_err, small_problem

GOLD = os.path.join(os.path.dirname(__file__), "golden", "elbo_small.npz")


def _both(model, X, idx, lp, gp, noise, args):
    l1, gl = elbo_ref.local_loss_grad(model, X[idx], idx, lp, gp, noise, *args)
    l2, gg = elbo_ref.global_loss_grad(model, X[idx], idx, lp, gp, noise, *args)
    l3, cl, cg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, *args)
    assert float(l1) == pytest.approx(float(l2), rel=1e-14)
    assert float(l1) == pytest.approx(l3, rel=1e-11)
    for a, b in zip(gl, cl):
        assert rel_err(a.numpy(), b) < 1e-8 if np.abs(b).max() > 0 else np.abs(a.numpy()).max() == 0
    for a, b in zip(gg, cg):
        assert rel_err(a.numpy(), b) < 1e-8 if np.abs(b).max() > 0 else np.abs(a.numpy()).max() == 0


@pytest.mark.parametrize("nb", [1, 3])
@pytest.mark.parametrize("use_brd", [True, False])
@pytest.mark.parametrize("marg", [False, True])
@pytest.mark.parametrize("sg", [[0, 0, 0], [1, 0, 0], [0, 0, 1], [1, 1, 0]])
def test_closed_form_matches_autograd(nb, use_brd, marg, sg):
    X, batch, model, lp, gp = small_problem(60, 40, (7, 4, 2), nb, use_brd=use_brd, marginalize_alpha=marg)
    idx = np.random.RandomState(0).permutation(60)[:16]
    noise = Noise.draw(3, 16, nb, 40, (7, 4, 2), seed=5)
    for scb, sgb in ((0, 0), (1, 0), (0, 1)):
        _both(model, X, idx, lp, gp, noise, (1.3, sg, scb, sgb, model.alpha))


def test_reference_coupling_shares_draws():
    n = Noise.draw(2, 10, 1, 12, (5, 3, 1), seed=1, couple_like_reference=True)
    assert torch.equal(n.tau, n.omega)                 # (K0,1) and (K0,1)
    assert torch.equal(n.c, n.z[:, :, 8:9])            # (B,1) and the width-1 root z
    X, batch, model, lp, gp = small_problem(40, 12, (5, 3, 1), 1)
    idx = np.arange(10)
    _both(model, X, idx, lp, gp, n, (1.0, [0, 0, 0], 0, 0, model.alpha))


def test_finite_differences():
    X, batch, model, lp, gp = small_problem(30, 20, (5, 3, 2), 2, marginalize_alpha=True)
    idx = np.random.RandomState(1).permutation(30)[:9]
    noise = Noise.draw(2, 9, 2, 20, (5, 3, 2), seed=3)
    args = (1.2, [0, 0, 0], 0, 0, model.alpha)
    lp = [p.astype(np.float64) for p in lp]
    gp = [p.astype(np.float64) for p in gp]
    _, cl, cg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, *args)
    rng = np.random.default_rng(0)
    h = 1e-6

    def f(lp_, gp_):
        return elbo_closed.loss_and_grads(model, X[idx], idx, lp_, gp_, noise, *args)[0]

    for which, params, grads in (("l", lp, cl), ("g", gp, cg)):
        for t in range(len(params)):
            for _ in range(4):
                pos = tuple(rng.integers(0, s) for s in params[t].shape)
                if which == "l" and pos[1] not in idx:
                    pos = (pos[0], int(idx[0])) + pos[2:]
                pp = [p.copy() for p in params]
                pm = [p.copy() for p in params]
                pp[t][pos] += h
                pm[t][pos] -= h
                fd = ((f(pp, gp) - f(pm, gp)) if which == "l" else (f(lp, pp) - f(lp, pm))) / (2 * h)
                an = grads[t][pos]
                assert fd == pytest.approx(an, rel=2e-4, abs=1e-3 * max(1.0, np.abs(grads[t]).max()) * 1e-2), (which, t, pos)


def test_invariants():
    X, batch, model, lp, gp = small_problem(40, 25, (6, 3, 1), 1)
    idx = np.arange(12)
    noise = Noise.draw(2, 12, 1, 25, (6, 3, 1), seed=2)
    S = 2
    # layer 0 stopped: likelihood contributes nothing, c / s / tau / W0 / z0 get exactly zero gradient
    loss, cl, cg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, 1.0, [1, 0, 0], 0, 0, model.alpha)
    assert np.all(cl[0] == 0) and np.all(cg[0] == 0) and np.all(cg[1] == 0) and np.all(cg[2] == 0)
    assert np.all(cl[1][:, :, :6] == 0) and np.any(cl[1][:, :, 6:] != 0)
    # rows outside the minibatch are exactly zero
    assert np.all(cl[1][:, 12:] == 0)
    # ard is never frozen
    assert np.any(cg[5] != 0)


def test_golden_file_matches_oracle():
    g = np.load(GOLD)
    Ks = tuple(int(k) for k in g["layer_sizes"])
    X, batch, idx = g["X"], g["batch"], g["idx"]
    model = host_ref.build_model(X, list(Ks), batch_ids=batch)
    lp = [g["lp0"], g["lp1"]]
    gp = [g[f"gp{i}"] for i in range(4 + len(Ks))]
    noise = Noise(c=torch.tensor(g["eps_c"]), s=torch.tensor(g["eps_s"]), tau=torch.tensor(g["eps_tau"]),
                  omega=torch.tensor(g["eps_omega"]), a=torch.tensor(g["eps_a"]),
                  W=[torch.tensor(g[f"eps_W{l}"]) for l in range(len(Ks))], z=torch.tensor(g["eps_z"]))
    loss, cl, cg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, float(g["annealing"]),
                                              [0, 0, 0], 0, 0, float(g["alpha"]))
    assert loss == pytest.approx(float(g["loss"]), rel=1e-11)
    for i in range(2):
        assert rel_err(cl[i], g[f"gl{i}"]) < 1e-8
    for i in range(len(gp)):
        if np.abs(g[f"gg{i}"]).max() > 0:
            assert rel_err(cg[i], g[f"gg{i}"]) < 1e-8


def test_float32_oracle_calibrates_tolerance():
    """What '1e-4' means against fp32 rounding: the same closed form in float32 stays well inside it."""
    X, batch, model, lp, gp = small_problem(80, 60, (9, 5, 2), 1)
    idx = np.random.RandomState(0).permutation(80)[:32]
    noise = Noise.draw(4, 32, 1, 60, (9, 5, 2), seed=4, dtype=torch.float32)
    args = (1.0, [0, 0, 0], 0, 0, model.alpha)
    l64, cl64, cg64 = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, *args)
    l32, cl32, cg32 = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, *args, dtype=np.float32)
    assert abs(l32 - l64) / abs(l64) < 1e-5
    for a, b in zip(cl32 + cg32, cl64 + cg64):
        if np.abs(b).max() > 0:
            assert rel_err(a, b) < 1e-4


# ---- the third-party arithmetic the oracle restates (tensorflow-probability, jax.scipy, optax are not
# ---- in the reference tree) against INDEPENDENT implementations of the same published definitions -----------------
def test_distribution_primitives_match_scipy_and_torch_distributions():
    """`jax_utils.py:6-46` are thin wrappers over tfd.LogNormal / tfd.Gamma and `_scdef.py:2119` over
    jax.scipy.stats.poisson: the oracle's closed forms must agree with scipy.stats and torch.distributions, which
    implement the same distributions independently."""
    import scipy.stats as st

    rng = np.random.default_rng(0)
    x = rng.gamma(2.0, 1.0, size=(6, 5)) + 1e-3
    shape = rng.uniform(0.05, 4.0, size=(6, 5))
    rate = rng.uniform(0.1, 9.0, size=(6, 5))
    got = float(elbo_ref.gamma_logpdf(torch.as_tensor(x), torch.as_tensor(shape), torch.as_tensor(rate)))
    assert got == pytest.approx(float(st.gamma.logpdf(x, shape, scale=1.0 / rate).sum()), rel=1e-12)
    td_g = torch.distributions.Gamma(torch.as_tensor(shape), torch.as_tensor(rate)).log_prob(torch.as_tensor(x)).sum()
    assert got == pytest.approx(float(td_g), rel=1e-12)

    m, s = rng.normal(size=(4, 7)), rng.uniform(0.05, 2.0, size=(4, 7))
    got = float(elbo_ref.lognormal_entropy(torch.as_tensor(m), torch.as_tensor(s)))
    assert got == pytest.approx(float(st.lognorm.entropy(s, scale=np.exp(m)).sum()), rel=1e-12)
    assert got == pytest.approx(float(torch.distributions.LogNormal(torch.as_tensor(m), torch.as_tensor(s)).entropy().sum()), rel=1e-12)

    k = rng.poisson(3.0, size=(5, 8)).astype(np.float64)
    mu = rng.uniform(0.01, 10.0, size=(5, 8))
    got = elbo_ref.poisson_logpmf(torch.as_tensor(k), torch.as_tensor(mu)).numpy()
    np.testing.assert_allclose(got, st.poisson.logpmf(k, mu), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got, torch.distributions.Poisson(torch.as_tensor(mu)).log_prob(torch.as_tensor(k)).numpy(), rtol=1e-12, atol=1e-12)

    eps = rng.normal(size=(3, 4))
    got = elbo_ref.lognormal_sample(torch.as_tensor(eps), torch.as_tensor(m[:3, :4]), torch.as_tensor(s[:3, :4])).numpy()
    base = torch.distributions.Normal(torch.as_tensor(m[:3, :4]), torch.as_tensor(s[:3, :4]))
    want = torch.distributions.LogNormal(base.loc, base.scale).icdf(torch.distributions.Normal(0.0, 1.0).cdf(torch.as_tensor(eps))).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9)


def test_adam_restatement_matches_torch_adam():
    """optax.adam(lr) (`_scdef.py:3079-3082`) = bias-corrected moments, update -lr m_hat / (sqrt(v_hat) + eps): the same
    published algorithm as torch.optim.Adam (amsgrad off, no weight decay).  20 steps, float64."""
    rng = np.random.default_rng(1)
    p0 = [rng.normal(size=(2, 5, 3)), rng.normal(size=(2, 4, 1))]
    grads = [[rng.normal(size=p.shape) * (1.0 + it) for p in p0] for it in range(20)]
    opt = host_ref.Adam(p0, 1e-2, dtype=np.float64)
    params = [p.copy() for p in p0]
    tp = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in p0]
    topt = torch.optim.Adam(tp, lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    for g in grads:
        params = opt.update(params, g)
        for t, gi in zip(tp, g):
            t.grad = torch.tensor(gi, dtype=torch.float64)
        topt.step()
    for a, t in zip(params, tp):
        np.testing.assert_allclose(a, t.detach().numpy(), rtol=1e-10, atol=1e-12)
