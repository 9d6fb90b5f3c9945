"""Oracle #2 — closed-form ELBO value and gradients in NumPy (no autodiff).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned to the reference's own source executed under
`oracle/jaxshim.py` (`tests/test_ref_pin.py`, 1e-9); the third-party formulas are restated, see there.

Same function as `oracle.elbo_ref` (reference `_scdef.py:1823-2181`, `jax_utils.py:6-46`)
but with the backward pass derived by hand.  This is the
arithmetic the CUDA kernels implement; `tests/test_oracle.py` requires it to agree with the
autograd restatement to ~1e-10 in float64 and with central finite differences.

All MC samples are evaluated at once (leading axis S).  Returns the reference's outputs:
loss = -mean_s ELBO_s and gradients shaped like `local_params` (dense (2,N,.), zero outside
the minibatch rows) or `global_params`.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import digamma, gammaln

from .elbo_ref import LOG_1E_10, LOG_1E6, LOG_1E10, OracleModel

HALF_LOG_2PI_PLUS_HALF = 0.5 + 0.5 * math.log(2.0 * math.pi)


# Hook for operand-precision experiments ONLY (scripts/numerics_probe.py): callable(A, B, which) -> A @ B with the operands
# rounded to a tensor-core format; None = exact float64 products (every test).
LIK_GEMM = None

def _np(x, dtype):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=dtype)


class _Block:
    """One LogNormal variational block: A.1 of SURVEY.md (`jax_utils.py:6-18`)."""

    def __init__(self, mu, rho, eps, stopped, w_H, scale_H, mu_hi=LOG_1E6):
        self.mu_raw, self.rho_raw, self.eps = mu, rho, eps
        self.mu = np.clip(mu, LOG_1E_10, mu_hi) if mu_hi is not None else np.maximum(mu, LOG_1E_10)
        self.in_mu = (mu >= LOG_1E_10) & ((mu <= mu_hi) if mu_hi is not None else True)
        rc = np.clip(rho, LOG_1E_10, LOG_1E10)
        self.in_rho = (rho >= LOG_1E_10) & (rho <= LOG_1E10)
        self.sigma = np.exp(rc)
        self.log_sigma = rc
        u = self.mu[None] + self.sigma[None] * eps  # (S, ...)
        raw = np.exp(u)
        self.v = np.clip(raw, 1e-15, 1e15)
        self.in_v = (raw >= 1e-15) & (raw <= 1e15)
        self.stopped, self.w_H, self.scale_H = bool(stopped), w_H, scale_H
        self.G = np.zeros_like(self.v)  # dELBO_s / dv, accumulated by the caller

    def entropy(self):
        return np.sum(self.mu + self.log_sigma + HALF_LOG_2PI_PLUS_HALF)

    def param_grads(self, S):
        """d(loss)/d(mu, rho) with loss = -(1/S) sum_s ELBO_s."""
        if self.stopped:
            return np.zeros_like(self.mu_raw), np.zeros_like(self.rho_raw)
        gv = self.in_v * self.G * self.v
        dmu = self.in_mu * (gv.sum(0) / S + self.w_H * self.scale_H)
        drho = self.in_rho * ((gv * self.eps * self.sigma[None]).sum(0) / S + self.w_H * self.scale_H)
        return -dmu, -drho


def _gamma_lp(x, a, b):
    # jax_utils.py:35-46; x>0 always here (clipped samples), so xlogy == (a-1) log x
    return (a - 1.0) * np.log(x) - b * x - gammaln(a) + a * np.log(b)


def loss_and_grads(model: OracleModel, X, indices, local_params, global_params, noise,
                   annealing_parameter, stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha,
                   dtype=np.float64):
    """Returns (loss, local_grads [2 arrays like local_params], global_grads [list])."""
    L = model.n_layers
    Ks = [int(k) for k in model.layer_sizes]
    K0 = Ks[0]
    off = np.concatenate([[0], np.cumsum(Ks)]).astype(int)
    idx = np.asarray(indices, dtype=np.int64)
    B = len(idx)
    N = model.n_cells
    scale = N / B
    sg = np.asarray(stop_gradients, dtype=float).reshape(-1)
    anneal = float(annealing_parameter)
    lp = [_np(p, dtype) for p in local_params]
    gp = [_np(p, dtype) for p in global_params]
    x = _np(X, dtype)
    S = noise.num_samples
    e = lambda t: _np(t, dtype)

    onehot = _np(model.batch_indices_onehot, dtype)[idx]  # (B, nb)
    nb = onehot.shape[1]
    G = x.shape[1]
    blr = _np(model.batch_lib_ratio, dtype)[idx]  # (B,1)
    caf = _np(model.cell_alpha_factor, dtype)[idx][:, None]  # (B,1)
    gene_ratio = np.broadcast_to(_np(model.gene_ratio, dtype), (nb, G))
    lik_on = 0.0 if sg[0] else 1.0

    # ---- blocks (A.1) -------------------------------------------------------------------
    c = _Block(lp[0][0][idx], lp[0][1][idx], e(noise.c), sg[0] or stop_cell_budgets, 1.0, scale)
    s = _Block(gp[0][0], gp[0][1], e(noise.s), sg[0] or stop_gene_budgets, 1.0, 1.0, mu_hi=None)
    tau = _Block(gp[1][0], gp[1][1], e(noise.tau), sg[0], 1.0, 1.0)
    om = _Block(gp[2 + L][0], gp[2 + L][1], e(noise.omega), False, 1.0, 1.0)
    aq = _Block(gp[3 + L][0], gp[3 + L][1], e(noise.a), False, 1.0, 1.0)
    W = [_Block(gp[2 + l][0], gp[2 + l][1], e(noise.W[l]), sg[l], anneal, 1.0) for l in range(L)]
    zeps = e(noise.z)
    z = []
    for l in range(L):
        sl = slice(off[l], off[l + 1])
        blk = _Block(lp[1][0][idx][:, sl], lp[1][1][idx][:, sl], zeps[:, :, sl], sg[l], anneal, scale)
        if model.supervised_layer is not None and l == model.supervised_layer:
            fz = _np(model.fixed_top_z, dtype)[idx]
            blk.v = np.broadcast_to(fz[None], blk.v.shape).copy()
            blk.stopped = True
            blk.w_H = 0.0
            blk.fixed = True
        z.append(blk)

    elbo = np.zeros(S, dtype=dtype)  # per-sample ELBO

    # ---- scale priors and entropies (A.4 last lines; _scdef.py:1946-1990) -----------------
    a_g, a_c = model.gene_scale_shape, model.cell_scale_shape
    elbo += _gamma_lp(s.v, a_g, a_g * gene_ratio[None]).sum((1, 2)) + s.entropy()
    s.G += (a_g - 1.0) / s.v - a_g * gene_ratio[None]
    elbo += scale * (_gamma_lp(c.v, a_c, a_c * blr[None]).sum((1, 2)) + c.entropy())
    c.G += scale * ((a_c - 1.0) / c.v - a_c * blr[None])

    if model.marginalize_alpha:
        b_a = 2.0 / max(float(alpha), 1e-8)
        elbo += _gamma_lp(aq.v, 2.0, b_a).sum((1, 2)) + aq.entropy()
        aq.G += 1.0 / aq.v - b_a
        alpha_value = aq.v.reshape(S, 1, 1)
    else:
        alpha_value = np.full((S, 1, 1), float(alpha), dtype=dtype)
    if model.use_brd:
        b_t = model.brd / model.brd_mean
        elbo += _gamma_lp(tau.v, model.brd, b_t).sum((1, 2)) + tau.entropy()
        tau.G += (model.brd - 1.0) / tau.v - b_t
        elbo += _gamma_lp(om.v, model.shrinkage_shape, model.shrinkage_rate).sum((1, 2)) + om.entropy()
        om.G += (model.shrinkage_shape - 1.0) / om.v - model.shrinkage_rate

    # ---- W priors (A.4; _scdef.py:2014-2045) ---------------------------------------------
    for l in range(L):
        P = np.broadcast_to(_np(model.w_priors[l][0], dtype), W[l].v.shape[1:])
        R = np.broadcast_to(_np(model.w_priors[l][1], dtype), W[l].v.shape[1:])
        w = W[l].v
        if l == 0 and model.use_brd:
            A = P[None] / tau.v  # (S,K0,G)
            Bm = R[None] * Ks[0] / (tau.v * om.v)
            dA = np.log(w) - digamma(A) + np.log(Bm)
            dB = -w + A / Bm
            tau.G += (dA * (-A / tau.v) + dB * (-Bm / tau.v)).sum(2, keepdims=True)
            om.G += (dB * (-Bm / om.v)).sum(2, keepdims=True)
        elif l == 1 and model.use_brd:
            A = np.broadcast_to(P[None], w.shape)
            tauT = np.swapaxes(tau.v, 1, 2)  # (S,1,K0)
            Bm = R[None] * Ks[1] / tauT
            dB = -w + A / Bm
            tau.G += np.swapaxes((dB * (-Bm / tauT)).sum(1, keepdims=True), 1, 2)
        else:
            A = np.broadcast_to(P[None], w.shape)
            Bm = np.broadcast_to(R[None] * Ks[l], w.shape)
        elbo += _gamma_lp(w, A, Bm).sum((1, 2)) + anneal * W[l].entropy()
        W[l].G += (A - 1.0) / w - Bm

    # ---- hierarchy z priors (A.4; _scdef.py:2047-2094) ------------------------------------
    alpha_n = np.maximum(alpha_value * caf[None], model.alpha_floor)  # (S,B,1)
    alpha_active = (alpha_value * caf[None]) > model.alpha_floor
    root_frozen = sg[L - 1] > 0.5
    for l in range(L - 1, -1, -1):
        zl = z[l].v
        use_top = (l == L - 1 and not root_frozen) or (l == L - 2 and root_frozen)
        if not getattr(z[l], "fixed", False):
            elbo += scale * anneal * z[l].entropy()
        if use_top:
            ta = model.top_alpha
            elbo += scale * _gamma_lp(zl, ta, ta).sum((1, 2))
            z[l].G += scale * ((ta - 1.0) / zl - ta)
        else:
            m = 1.0 if l == L - 1 else z[l + 1].v @ W[l + 1].v  # (S,B,K_l)
            m = np.broadcast_to(m, zl.shape)
            elbo += scale * _gamma_lp(zl, alpha_n, alpha_n / m).sum((1, 2))
            z[l].G += scale * ((alpha_n - 1.0) / zl - alpha_n / m)
            if l < L - 1:
                Pm = (alpha_n / m) * (zl / m - 1.0)  # dELBO/dm
                z[l + 1].G += scale * (Pm @ np.swapaxes(W[l + 1].v, 1, 2))
                W[l + 1].G += scale * (np.swapaxes(z[l + 1].v, 1, 2) @ Pm)
            if model.marginalize_alpha:
                dal = np.log(zl) - digamma(alpha_n) + np.log(alpha_n / m) + 1.0 - zl / m
                aq.G += scale * (dal * alpha_active * caf[None]).sum((1, 2)).reshape(S, 1, 1)

    # ---- Poisson likelihood (A.3; _scdef.py:2096-2136) -------------------------------------
    if lik_on:
        Rm = z[0].v @ W[0].v if LIK_GEMM is None else LIK_GEMM(z[0].v, W[0].v, "rate")  # (S,B,G)
        s_cell = np.einsum("nb,sbg->sng", onehot, s.v)
        lam = Rm * c.v * s_cell
        with np.errstate(divide="ignore", invalid="ignore"):
            xlogy = np.where(x[None] > 0, x[None] * np.log(lam), 0.0)
        elbo += scale * (xlogy - lam - gammaln(x + 1.0)[None]).sum((1, 2))
        E = x[None] - lam
        if LIK_GEMM is None:
            Q = E / Rm
            z[0].G += scale * (Q @ np.swapaxes(W[0].v, 1, 2))
            W[0].G += scale * (np.swapaxes(z[0].v, 1, 2) @ Q)
        else:
            # numerics experiments (scripts/numerics_probe.py): the kernels' split Q = x/R - c s^T, the x/R part through
            # the operand-rounding GEMM under test, the rank-1 part exact
            with np.errstate(divide="ignore", invalid="ignore"):
                Qx = np.where(x[None] > 0, x[None] / Rm, 0.0)
            cs = c.v * s_cell
            z[0].G += scale * (LIK_GEMM(Qx, np.swapaxes(W[0].v, 1, 2), "second") - cs @ np.swapaxes(W[0].v, 1, 2))
            W[0].G += scale * (LIK_GEMM(np.swapaxes(z[0].v, 1, 2), Qx, "second") - np.swapaxes(z[0].v, 1, 2) @ cs)
        c.G += scale * E.sum(2, keepdims=True) / c.v
        s.G += scale * np.einsum("nb,sng->sbg", onehot, E) / s.v

    loss = -float(np.mean(elbo))

    # ---- assemble (A.5) ------------------------------------------------------------------
    g_local = [np.zeros_like(lp[0]), np.zeros_like(lp[1])]
    dmu, drho = c.param_grads(S)
    np.add.at(g_local[0][0], idx, dmu)
    np.add.at(g_local[0][1], idx, drho)
    for l in range(L):
        dmu, drho = z[l].param_grads(S)
        sl = slice(off[l], off[l + 1])
        tmp0 = np.zeros((B, off[-1]), dtype=dtype)
        tmp1 = np.zeros((B, off[-1]), dtype=dtype)
        tmp0[:, sl], tmp1[:, sl] = dmu, drho
        np.add.at(g_local[1][0], idx, tmp0)
        np.add.at(g_local[1][1], idx, tmp1)

    g_global = [np.zeros_like(p) for p in gp]
    g_global[0][0], g_global[0][1] = s.param_grads(S)
    if model.use_brd:
        g_global[1][0], g_global[1][1] = tau.param_grads(S)
        g_global[2 + L][0], g_global[2 + L][1] = om.param_grads(S)
    for l in range(L):
        g_global[2 + l][0], g_global[2 + l][1] = W[l].param_grads(S)
    if model.marginalize_alpha:
        g_global[3 + L][0], g_global[3 + L][1] = aq.param_grads(S)
    return loss, g_local, g_global
