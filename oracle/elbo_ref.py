"""Oracle #1 — line-by-line restatement of the reference ELBO in torch (CPU autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned to the reference's own source executed under
`oracle/jaxshim.py` (`tests/test_ref_pin.py`, 1e-11); the third-party formulas are restated, see there.

Follows `/root/reference/src/scdef/models/_scdef.py`:
  * `elbo`            1823-2146   -> `elbo_sample`
  * `batch_elbo`      2148-2181   -> `batch_elbo`
  * `objective`       2822-2846   -> `objective`  (loss = -batch_elbo)
  * value_and_grad    2848-2849   -> `local_loss_grad`, `global_loss_grad`
and `/root/reference/src/scdef/utils/jax_utils.py`:
  * `lognormal_sample` 6-8, `lognormal_entropy` 11-18, `gamma_logpdf` 35-46.
sscDEF variant (`_sscdef.py:484-487, 505-507, 530-531`) through `OracleModel.supervised_layer`.

The only deliberate difference: the reference draws its N(0,1) noise from a JAX PRNG key
inside `lognormal_sample`; here the noise is INJECTED (`Noise`), one tensor per variational
block, so that the oracle and the CUDA path can be fed identical draws.
The reference's "same key for every block" quirk (identical shapes -> identical noise) is
reproduced by the caller passing the same tensor for those blocks (`Noise.couple_like_reference`).

Gradients come from torch autograd on the restated graph; `jax.lax.stop_gradient` is
`.detach()`, `jnp.clip` is `torch.clamp` (both pass the gradient strictly inside the
interval and block it outside; they differ only AT the bounds, a measure-zero set).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

LOG_1E_10 = math.log(1e-10)  # min_shape / min_rate   (_scdef.py:1835,1837)
LOG_1E6 = math.log(1e6)  # max_shape              (_scdef.py:1836)
LOG_1E10 = math.log(1e10)  # max_rate               (_scdef.py:1838)


@dataclass
class OracleModel:
    """The attributes of the reference `scDEF` object that `elbo` reads."""

    n_cells: int
    layer_sizes: List[int]
    batch_indices_onehot: np.ndarray  # (N, nb)            _scdef.py:498,546-560
    batch_lib_ratio: np.ndarray  # (N, 1)             _scdef.py:500-504,562-564
    gene_ratio: object  # scalar or (nb, G)  _scdef.py:521-524,568-573
    cell_alpha_factor: np.ndarray  # (N,)               _scdef.py:1279-1281
    w_priors: list  # [[shape, rate]]*L  _scdef.py:1258-1277
    use_brd: bool = True
    marginalize_alpha: bool = False
    top_alpha: float = 1.0
    brd: float = 1.0
    brd_mean: float = 1.0
    shrinkage_shape: float = 1.0
    shrinkage_rate: float = 1.0
    cell_scale_shape: float = 1.0
    gene_scale_shape: float = 1.0
    alpha_floor: float = 0.1  # 1.0 in sscDEF (_sscdef.py:505-507)
    supervised_layer: Optional[int] = None  # sscDEF pinned layer
    fixed_top_z: Optional[np.ndarray] = None  # (N, K_t)  _sscdef.py:484-487

    @property
    def n_layers(self) -> int:
        return len(self.layer_sizes)


@dataclass
class Noise:
    """Injected standard-normal draws, leading axis = MC sample."""

    c: torch.Tensor  # (S, B, 1)
    s: torch.Tensor  # (S, nb, G)
    tau: torch.Tensor  # (S, K0, 1)
    omega: torch.Tensor  # (S, K0, 1)
    a: torch.Tensor  # (S, 1, 1)
    W: List[torch.Tensor]  # (S, K_l, K_{l-1}), K_{-1}=G
    z: torch.Tensor  # (S, B, sum K)

    @property
    def num_samples(self) -> int:
        return int(self.c.shape[0])

    @staticmethod
    def draw(S, B, nb, G, layer_sizes, seed=0, dtype=torch.float64, couple_like_reference=False):
        g = torch.Generator().manual_seed(int(seed))
        rn = lambda *shape: torch.randn(*shape, generator=g, dtype=torch.float64).to(dtype)
        Ks = list(layer_sizes)
        outs = [G] + Ks[:-1]
        n = Noise(
            c=rn(S, B, 1),
            s=rn(S, nb, G),
            tau=rn(S, Ks[0], 1),
            omega=rn(S, Ks[0], 1),
            a=rn(S, 1, 1),
            W=[rn(S, Ks[l], outs[l]) for l in range(len(Ks))],
            z=rn(S, B, sum(Ks)),
        )
        if couple_like_reference:
            n = n.couple_like_reference(Ks)
        return n

    def couple_like_reference(self, layer_sizes):
        """Blocks of identical shape share their draws (one PRNG key per `elbo` call,
        `_scdef.py:1924-2061`).  Order of first use in `elbo`:
        c, s, tau, omega, [a], then per layer top->bottom W_l, z_l."""
        Ks = list(layer_sizes)
        seen = {}

        def share(t):
            key = tuple(t.shape[1:])
            if key in seen:
                return seen[key]
            seen[key] = t
            return t

        c = share(self.c)
        s = share(self.s)
        tau = share(self.tau)
        omega = share(self.omega)
        a = share(self.a)
        W = list(self.W)
        z_parts = [None] * len(Ks)
        start = np.concatenate([[0], np.cumsum(Ks)]).astype(int)
        for l in range(len(Ks) - 1, -1, -1):
            W[l] = share(W[l])
            z_parts[l] = share(self.z[:, :, start[l] : start[l + 1]])
        return Noise(c=c, s=s, tau=tau, omega=omega, a=a, W=W, z=torch.cat(z_parts, dim=2))

    def sample(self, i):
        return dict(
            c=self.c[i], s=self.s[i], tau=self.tau[i], omega=self.omega[i], a=self.a[i],
            W=[w[i] for w in self.W], z=self.z[i],
        )


# ----- jax_utils.py:6-46 -------------------------------------------------------------------

def lognormal_sample(eps, m, s):
    # tfd.LogNormal(m, s).sample() = exp(m + s * N(0,1)); clip  (jax_utils.py:6-8)
    return torch.clamp(torch.exp(m + s * eps), 1e-15, 1e15)


def lognormal_entropy(m, s):
    # tfd.LogNormal.entropy = loc + 0.5 + 0.5 log(2 pi) + log(scale)  (jax_utils.py:11-18)
    return torch.sum(m + 0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(s))


def gamma_logpdf(x, shape, rate):
    # tfd.Gamma(concentration, rate).log_prob(x)  (jax_utils.py:35-46)
    shape = shape * torch.ones_like(x)
    rate = rate * torch.ones_like(x)
    return torch.sum(
        torch.xlogy(shape - 1.0, x) - rate * x - torch.lgamma(shape) + shape * torch.log(rate)
    )


def poisson_logpmf(k, mu):
    # jax.scipy.stats.poisson.logpmf: xlogy(k, mu) - gammaln(k+1) - mu  (_scdef.py:2119)
    return torch.xlogy(k, mu) - torch.lgamma(k + 1.0) - mu


def _stop(flag, t):
    # jax.lax.cond(flag, stop_gradient(t), t): any non-zero flag stops
    return t.detach() if bool(flag) else t


def _t(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ----- _scdef.py:1823-2146 ------------------------------------------------------------------

def elbo_sample(model: OracleModel, eps: dict, batch, indices, local_params, global_params,
                annealing_parameter, stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha):
    dtype = local_params[0].dtype
    L = model.n_layers
    annealing_z = annealing_parameter
    annealing_w = annealing_parameter
    annealing_brd = annealing_ard = annealing_scales = 1.0  # 1844-1846
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    sg = [float(v) for v in np.asarray(stop_gradients, dtype=float).reshape(-1)]

    onehot = _t(model.batch_indices_onehot, dtype)[idx]  # 1849

    cell_budget_params = local_params[0]
    z_params = local_params[1]
    gene_budget_params = global_params[0]
    fscale_params = global_params[1]

    cell_budget_shape = torch.clamp(cell_budget_params[0][idx], LOG_1E_10, LOG_1E6)  # 1856
    cell_budget_rate = torch.exp(torch.clamp(cell_budget_params[1][idx], LOG_1E_10, LOG_1E10))
    cell_budget_shape = _stop(sg[0], cell_budget_shape)  # 1861-1870
    cell_budget_rate = _stop(sg[0], cell_budget_rate)

    gene_budget_shape = torch.clamp(gene_budget_params[0], min=LOG_1E_10)  # 1872 (max only)
    gene_budget_rate = torch.exp(torch.clamp(gene_budget_params[1], LOG_1E_10, LOG_1E10))
    gene_budget_shape = _stop(sg[0], gene_budget_shape)
    gene_budget_rate = _stop(sg[0], gene_budget_rate)

    fscale_shapes = torch.clamp(fscale_params[0], LOG_1E_10, LOG_1E6)  # 1885
    fscale_rates = torch.exp(torch.clamp(fscale_params[1], LOG_1E_10, LOG_1E10))
    fscale_shapes = _stop(sg[0], fscale_shapes)
    fscale_rates = _stop(sg[0], fscale_rates)

    wm_shape = torch.clamp(global_params[2 + L][0], LOG_1E_10, LOG_1E6)  # 1898
    wm_rate = torch.exp(torch.clamp(global_params[2 + L][1], LOG_1E_10, LOG_1E10))

    s_shape = torch.clamp(global_params[2 + L + 1][0], LOG_1E_10, LOG_1E6)  # 1903
    s_rate = torch.exp(torch.clamp(global_params[2 + L + 1][1], LOG_1E_10, LOG_1E10))

    z_shapes = torch.clamp(z_params[0][idx], LOG_1E_10, LOG_1E6)  # 1910
    z_rates = torch.exp(torch.clamp(z_params[1][idx], LOG_1E_10, LOG_1E10))

    cell_budget_shape = _stop(stop_cell_budgets, cell_budget_shape)  # 1914-1923
    cell_budget_rate = _stop(stop_cell_budgets, cell_budget_rate)
    cell_budget_sample = lognormal_sample(eps["c"], cell_budget_shape, cell_budget_rate)
    gene_budget_shape = _stop(stop_gene_budgets, gene_budget_shape)  # 1925-1934
    gene_budget_rate = _stop(stop_gene_budgets, gene_budget_rate)
    gene_budget_sample = lognormal_sample(eps["s"], gene_budget_shape, gene_budget_rate)
    if model.use_brd:  # 1936-1939
        fscale_samples = lognormal_sample(eps["tau"], fscale_shapes, fscale_rates)
        _wm_sample = lognormal_sample(eps["omega"], wm_shape, wm_rate)
        brd_samples = fscale_samples

    gene_ratio = _t(model.gene_ratio, dtype)
    global_pl = gamma_logpdf(  # 1946-1950
        gene_budget_sample, model.gene_scale_shape, model.gene_scale_shape * gene_ratio
    )
    global_en = annealing_scales * lognormal_entropy(gene_budget_shape, gene_budget_rate)
    local_pl = gamma_logpdf(  # 1961-1965
        cell_budget_sample,
        model.cell_scale_shape,
        model.cell_scale_shape * _t(model.batch_lib_ratio, dtype)[idx],
    )
    local_en = annealing_scales * lognormal_entropy(cell_budget_shape, cell_budget_rate)

    alpha_value = torch.as_tensor(float(alpha), dtype=dtype)  # 1971
    if model.marginalize_alpha:  # 1972-1981
        s_sample = lognormal_sample(eps["a"], s_shape, s_rate)
        alpha_value = torch.squeeze(s_sample)
        global_pl = global_pl + gamma_logpdf(s_sample, 2.0, 2.0 / max(float(alpha), 1e-8))
        global_en = global_en + lognormal_entropy(s_shape, s_rate)
    if model.use_brd:  # 1982-1990
        global_pl = global_pl + gamma_logpdf(fscale_samples, model.brd, model.brd / model.brd_mean)
        global_en = global_en + annealing_brd * lognormal_entropy(fscale_shapes, fscale_rates)
        global_pl = global_pl + gamma_logpdf(_wm_sample, model.shrinkage_shape, model.shrinkage_rate)
        global_en = global_en + annealing_ard * lognormal_entropy(wm_shape, wm_rate)

    z_mean = 1.0
    caf = _t(model.cell_alpha_factor, dtype)[idx][:, None]
    for l in range(L - 1, -1, -1):  # 1993
        start = int(np.sum(model.layer_sizes[:l]))
        end = start + int(model.layer_sizes[l])
        K_l = float(model.layer_sizes[l])

        _w_shape = torch.clamp(global_params[2 + l][0], LOG_1E_10, LOG_1E6)  # 1998
        _w_rate = torch.exp(torch.clamp(global_params[2 + l][1], LOG_1E_10, LOG_1E10))
        _w_shape = _stop(sg[l], _w_shape)
        _w_rate = _stop(sg[l], _w_rate)
        _w_sample = lognormal_sample(eps["W"][l], _w_shape, _w_rate)  # 2012

        P = _t(model.w_priors[l][0], dtype)
        R = _t(model.w_priors[l][1], dtype)
        if l == 0 and model.use_brd:  # 2014-2022
            global_pl = global_pl + gamma_logpdf(
                _w_sample,
                P * (1.0 / brd_samples),
                R * (1.0 / brd_samples) * (1.0 / _wm_sample) * K_l,
            )
        elif l == 0:  # 2023-2028
            global_pl = global_pl + gamma_logpdf(_w_sample, P, R * K_l)
        else:
            if l == 1 and model.use_brd:  # 2030-2038
                global_pl = global_pl + gamma_logpdf(_w_sample, P, R * 1.0 / brd_samples.T * K_l)
            else:  # 2039-2044
                global_pl = global_pl + gamma_logpdf(_w_sample, P, R * K_l)
        global_en = global_en + annealing_w * lognormal_entropy(_w_shape, _w_rate)  # 2045

        _z_shape = z_shapes[:, start:end]
        _z_rate = z_rates[:, start:end]
        supervised = model.supervised_layer is not None and l == model.supervised_layer
        if supervised:  # _sscdef.py:484-487
            _z_shape = _z_shape.detach()
            _z_rate = _z_rate.detach()
            _z_sample = _t(model.fixed_top_z, dtype)[idx]
        else:
            _z_shape = _stop(sg[l], _z_shape)  # 2051-2060
            _z_rate = _stop(sg[l], _z_rate)
            _z_sample = lognormal_sample(eps["z"][:, start:end], _z_shape, _z_rate)  # 2061

        alpha_layer = torch.clamp(alpha_value * caf, min=model.alpha_floor)  # 2066-2069
        root_frozen = sg[L - 1] > 0.5  # 2075
        use_top_prior = (l == L - 1 and not root_frozen) or (l == L - 2 and root_frozen)
        if use_top_prior:  # 2082-2090
            local_pl = local_pl + gamma_logpdf(_z_sample, model.top_alpha, model.top_alpha)
        else:
            local_pl = local_pl + gamma_logpdf(_z_sample, alpha_layer, alpha_layer / z_mean)
        if not supervised:
            local_en = local_en + annealing_z * lognormal_entropy(_z_shape, _z_rate)  # 2092
        z_mean = _z_sample @ _w_sample  # 2094

    mean_bottom = (_z_sample @ _w_sample) * cell_budget_sample  # 2100-2102
    mean_bottom = mean_bottom * (onehot @ gene_budget_sample)  # 2103
    ll = torch.sum(poisson_logpmf(_t(batch, dtype), mean_bottom))  # 2119
    if sg[0]:  # 2132-2136
        ll = ll * 0.0

    return (ll + local_pl + local_en) * (model.n_cells / len(idx)) + global_pl + global_en  # 2142


def batch_elbo(model, noise: Noise, X, indices, local_params, global_params,
               annealing_parameter, stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha):
    # _scdef.py:2148-2181 — mean over the S per-sample ELBOs
    vals = [
        elbo_sample(model, noise.sample(i), X, indices, local_params, global_params,
                    annealing_parameter, stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha)
        for i in range(noise.num_samples)
    ]
    return torch.mean(torch.stack(vals))


def objective(model, X, indices, local_params, global_params, noise, annealing_parameter,
              stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha):
    # _scdef.py:2822-2846
    return -batch_elbo(model, noise, X, indices, local_params, global_params, annealing_parameter,
                       stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha)


def _as_leaves(params, dtype, requires_grad):
    out = []
    for p in params:
        t = _t(p, dtype).clone().detach()
        t.requires_grad_(requires_grad)
        out.append(t)
    return out


def _loss_grad(model, which, X, indices, local_params, global_params, noise, annealing_parameter,
               stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, dtype):
    lp = _as_leaves(local_params, dtype, which == "local")
    gp = _as_leaves(global_params, dtype, which == "global")
    noise = Noise(c=noise.c.to(dtype), s=noise.s.to(dtype), tau=noise.tau.to(dtype),
                  omega=noise.omega.to(dtype), a=noise.a.to(dtype),
                  W=[w.to(dtype) for w in noise.W], z=noise.z.to(dtype))
    loss = objective(model, X, indices, lp, gp, noise, annealing_parameter, stop_gradients,
                     stop_cell_budgets, stop_gene_budgets, alpha)
    leaves = lp if which == "local" else gp
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = [torch.zeros_like(p) if g is None else g for p, g in zip(leaves, grads)]
    return loss.detach(), grads


def local_loss_grad(model, X, indices, local_params, global_params, noise, annealing_parameter,
                    stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, dtype=torch.float64):
    """jit(value_and_grad(objective, argnums=2)) — `_scdef.py:2848`.  Dense (2,N,.) grads."""
    return _loss_grad(model, "local", X, indices, local_params, global_params, noise,
                      annealing_parameter, stop_gradients, stop_cell_budgets, stop_gene_budgets,
                      alpha, dtype)


def global_loss_grad(model, X, indices, local_params, global_params, noise, annealing_parameter,
                     stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, dtype=torch.float64):
    """jit(value_and_grad(objective, argnums=3)) — `_scdef.py:2849`."""
    return _loss_grad(model, "global", X, indices, local_params, global_params, noise,
                      annealing_parameter, stop_gradients, stop_cell_budgets, stop_gene_budgets,
                      alpha, dtype)
