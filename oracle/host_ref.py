"""Oracle — host-side pieces around the ELBO: constants, layer ladder, parameter
initialisation, minibatch stream, Adam + clip, one `_learn` step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Constants, Adam / clip and the step order are pinned to
the reference's own `__init__` / `_learn` executed under `oracle/jaxshim.py` (`tests/test_ref_pin.py`);
the integer parts (ladder, permutation) also by the reference's own tests
(`/root/reference/tests/test_sscdef.py:55-83`) and by NumPy itself.

Restates `/root/reference/src/scdef/models/_scdef.py`:
  `load_adata` 482-578, `_geometric_layer_sizes` 580-628, `update_model_priors` 1246-1287,
  `init_var_params` 1463-1732 (cold start), `clip_params` 3034-3042, optimisers 3079-3082,
  `local_update` 3087-3117, `global_update` 3124-3163, `data_stream` 3167-3173,
  hot loop 3230-3271, `set_posterior_means/variances` 3394-3476.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .elbo_ref import OracleModel


# ---- _scdef.py:580-628 -----------------------------------------------------------------------
def geometric_layer_sizes(n_factors: int, top_factors: int, n_layers: int) -> List[int]:
    n_layers = max(2, int(n_layers))
    min_k, top_k = float(n_factors), float(top_factors)
    ratio = top_k / max(min_k, 1.0)
    add_root = int(top_factors) > 1
    if n_layers == 2:
        out = [max(1, int(round(min_k))), int(top_factors)]
        if out[0] < out[1]:
            out[0] = out[1]
    else:
        out, prev = [], None
        for l in range(n_layers):
            if l == n_layers - 1:
                k_i = int(top_factors)
            else:
                k_i = max(1, int(round(min_k * ratio ** (l / float(n_layers - 1)))))
                if prev is not None and k_i >= prev:
                    k_i = max(1, prev - 1)
            out.append(k_i)
            prev = k_i
        out[-1] = int(top_factors)
        for i in range(len(out) - 2, -1, -1):
            out[i] = max(out[i], out[i + 1])
    if add_root:
        out.append(1)
    return out


# ---- _scdef.py:482-578 + 1246-1287 ----------------------------------------------------------
def build_model(X, layer_sizes, batch_ids=None, use_brd=True, factor_shape=0.1, alpha=1.0,
                set_alpha_from_cov=True, hierarchy_weight=1.0, **hyper) -> OracleModel:
    """Constants `elbo` reads, computed from the count matrix as the reference does."""
    X = np.asarray(X, dtype=np.float64)
    N, G = X.shape
    eps = 1e-12
    lib = X.sum(1)
    blr = np.ones((N, 1)) * lib.mean() / lib.var()
    gene_size = X.sum(0)
    gene_ratio = gene_size.mean() / gene_size.var()
    onehot = np.ones((N, 1))
    mu = X.mean(0) + eps
    gene_ratio_init = (mu.mean() / mu)[None, :]
    if batch_ids is not None:
        batches = np.unique(batch_ids)
        nb = len(batches)
        onehot = np.zeros((N, nb))
        if nb > 1:
            gene_ratio = np.ones((nb, G))
            gene_ratio_init = np.ones((nb, G))
            for i, b in enumerate(batches):
                cells = np.where(np.asarray(batch_ids) == b)[0]
                onehot[cells, i] = 1
                blr[cells] = lib[cells].mean() / (lib[cells].var() + eps)
                mu_b = X[cells].mean(0) + eps
                gene_ratio_init[i] = mu_b.mean() / mu_b
                gs = X[cells].sum(0)
                gene_ratio[i] = gs.mean() / gs.var()
        # nb == 1 with a batch_key: the reference leaves the one-hot all-zero (546-548 never
        # filled, loop at 552 is skipped) -> mean_bottom == 0.  Not reproduced: treated as no key.
        else:
            onehot = np.ones((N, 1))
    L = len(layer_sizes)
    fshapes = ([1.0] + [factor_shape] * (L - 1)) if use_brd else [factor_shape] * L
    w_priors = []
    for l in range(L):
        if l == 0:
            w_priors.append([fshapes[0], fshapes[0]])
        else:
            m = np.clip(np.ones((layer_sizes[l], layer_sizes[l - 1])) * fshapes[l], 1e-12, 1e12)
            w_priors.append([m.copy(), m.copy()])
    caf = lib / float(np.median(lib))
    if set_alpha_from_cov:
        alpha = float(np.median(lib)) / float(layer_sizes[0]) * hierarchy_weight
    model = OracleModel(
        n_cells=N, layer_sizes=[int(k) for k in layer_sizes], batch_indices_onehot=onehot,
        batch_lib_ratio=blr, gene_ratio=gene_ratio, cell_alpha_factor=caf, w_priors=w_priors,
        use_brd=use_brd, **hyper,
    )
    model.alpha = alpha
    model.batch_lib_sizes = lib
    model.gene_ratio_init = gene_ratio_init
    return model


def _moment_match(m, v):
    # mean/variance -> LogNormal (mu, log sigma); used throughout init_var_params 1496-1732
    return np.log(m**2 / np.sqrt(m**2 + v)), np.log(np.sqrt(np.log(1.0 + v / m**2)))


def init_var_params(model: OracleModel, seed=42, z_init_concentration=0.05, brd_mean=1.0):
    """Cold-start initialisation (distribution-level restatement: the reference draws its
    Gamma variates from JAX keys, here from NumPy's Generator)."""
    rng = np.random.default_rng(seed)
    N, Ks = model.n_cells, model.layer_sizes
    L = len(Ks)
    G = np.asarray(model.gene_ratio_init).shape[1]
    lib = model.batch_lib_sizes
    m = np.clip((lib / lib.mean())[:, None], 1e-3, 1e2)
    local = [np.stack(_moment_match(m, m / 10.0)).astype(np.float32)]
    m = np.clip(1.0 / np.asarray(model.gene_ratio_init), 1e-6, 1e6)
    glob = [np.stack(_moment_match(m, m / 10.0)).astype(np.float32)]
    m_brd = rng.gamma(100.0, brd_mean / 100.0, size=(Ks[0], 1))
    glob.append(np.stack(_moment_match(m_brd, m_brd**2 / 100.0)).astype(np.float32))
    zs = []
    for l in range(L):
        a = z_init_concentration if l == 0 else 1.0
        if l == L - 1:
            a = 100.0
        m = np.clip(rng.gamma(a, 1.0 / a, size=(N, Ks[l])), 1e-3, 4.0)
        zs.append(_moment_match(m, m / 100.0 if l == 0 else m / 10.0))
    local.append(np.stack([np.hstack([z[0] for z in zs]), np.hstack([z[1] for z in zs])]).astype(np.float32))
    for l in range(L):
        out = G if l == 0 else Ks[l - 1]
        m = np.ones((Ks[l], out)) / Ks[l] * np.asarray(model.w_priors[l][0]) / np.asarray(model.w_priors[l][1])
        if l < L - 1:
            m = rng.gamma(10.0, m / 10.0)
        glob.append(np.stack(_moment_match(m, m / 100.0 if l == 0 else m / 10.0)).astype(np.float32))
    m = model.shrinkage_shape / model.shrinkage_rate * np.ones((Ks[0], 1))
    glob.append(np.stack(_moment_match(m, m**2 / 10.0)).astype(np.float32))
    m = float(model.alpha) * np.ones((1, 1))
    glob.append(np.stack(_moment_match(m, m**2 / 10.0)).astype(np.float32))
    return local, glob


# ---- _scdef.py:3167-3173 --------------------------------------------------------------------
def data_stream_indices(n_cells: int, batch_size: Optional[int]):
    """Yields the minibatch index arrays exactly as the reference: RandomState(0), a fresh
    permutation per epoch, contiguous slices, short last slice."""
    if batch_size is None:
        batch_size = n_cells
    num_complete, leftover = divmod(n_cells, batch_size)
    num_batches = num_complete + bool(leftover)
    rng = np.random.RandomState(0)
    while True:
        perm = rng.permutation(n_cells)
        for i in range(num_batches):
            yield perm[i * batch_size : (i + 1) * batch_size]


# ---- optax.adam (3079-3082) + clip_params (3034-3042) ---------------------------------------
class Adam:
    """optax.adam(lr): scale_by_adam(b1=.9, b2=.999, eps=1e-8, eps_root=0) then scale(-lr)."""

    def __init__(self, params, lr, b1=0.9, b2=0.999, eps=1e-8, dtype=np.float32):
        self.lr, self.b1, self.b2, self.eps, self.dtype = lr, b1, b2, eps, dtype
        self.count = 0
        self.m = [np.zeros_like(np.asarray(p, dtype=dtype)) for p in params]
        self.v = [np.zeros_like(np.asarray(p, dtype=dtype)) for p in params]

    def update(self, params, grads):
        dt = self.dtype
        self.count += 1
        bc1 = dt(1.0 - np.power(dt(self.b1), self.count, dtype=dt))
        bc2 = dt(1.0 - np.power(dt(self.b2), self.count, dtype=dt))
        out = []
        for i, (p, g) in enumerate(zip(params, grads)):
            g = np.asarray(g, dtype=dt)
            self.m[i] = dt(self.b1) * self.m[i] + dt(1.0 - self.b1) * g
            self.v[i] = dt(self.b2) * self.v[i] + dt(1.0 - self.b2) * g * g
            upd = (self.m[i] / bc1) / (np.sqrt(self.v[i] / bc2) + dt(self.eps))
            out.append(np.asarray(p, dtype=dt) + dt(-self.lr) * upd)
        return out


def clip_params(params, min_mu=-1e10, max_mu=1e4, min_logstd=-1e10, max_logstd=1e1):
    # `for i in range(len(params))[:-2]` — the last two tensors are never clipped, so a
    # 2-element list (the locals) is never clipped at all.
    params = list(params)
    for i in range(len(params))[:-2]:
        p = np.array(params[i])
        p[0] = np.clip(p[0], min_mu, max_mu)
        p[1] = np.clip(p[1], min_logstd, max_logstd)
        params[i] = p
    return params


def learn_step(model, X_batch, indices, local_params, global_params, noise, local_opt: Adam,
               global_opt: Adam, annealing, stop_gradients, stop_cell_budgets, stop_gene_budgets,
               alpha, update_locals=True, update_globals=True, freeze_w=False,
               freeze_z_slices=(), z_orig=None, dtype=np.float64, post_local=None):
    """One iteration of the hot loop (`_scdef.py:3230-3271`): local pass + Adam, optional
    z snap-back, global pass (same noise, updated locals) + Adam + freeze_w restore + clip."""
    from . import elbo_closed

    L = model.n_layers
    loss = None
    if update_locals:
        loss, g_loc, _ = elbo_closed.loss_and_grads(
            model, X_batch, indices, local_params, global_params, noise, annealing, stop_gradients,
            stop_cell_budgets, stop_gene_budgets, alpha, dtype=dtype)
        local_params = clip_params(local_opt.update(local_params, g_loc))
        for (s, e) in freeze_z_slices:
            local_params[1] = np.array(local_params[1])
            local_params[1][:, :, s:e] = z_orig[:, :, s:e]
        if post_local is not None:  # sscDEF: `_sscdef.py:892-893` runs between the two updates
            local_params = post_local(local_params)
    if update_globals:
        loss, _, g_glob = elbo_closed.loss_and_grads(
            model, X_batch, indices, local_params, global_params, noise, annealing, stop_gradients,
            stop_cell_budgets, stop_gene_budgets, alpha, dtype=dtype)
        new_glob = global_opt.update(global_params, g_glob)
        if freeze_w:
            for i in range(2, 2 + L):
                new_glob[i] = np.asarray(global_params[i])
        global_params = clip_params(new_glob)
    return loss, local_params, global_params


# ---- _scdef.py:3394-3476 --------------------------------------------------------------------
def lognormal_mean(mu, log_sigma):
    return np.exp(mu + 0.5 * np.exp(log_sigma) ** 2)


def lognormal_var(mu, log_sigma):
    s2 = np.exp(log_sigma) ** 2
    return (np.exp(s2) - 1.0) * np.exp(2.0 * mu + s2)


def filter_factors_upper_only(pmeans_z, pmeans_W, layer_sizes, n_cells, min_cells_upper=0.001, filter_up=True):
    """`filter_factors(upper_only=True)` as `_learn` calls it (`_scdef.py:3386-3387, 3518-3568`): layer 0 is kept
    whole; an upper layer keeps the factors that at least `min_cells_upper` cells are hard-assigned to
    (argmax of the posterior mean z, `3540-3549`), intersected — `filter_up` — with the factors that the kept
    factors of the layer below attach to most strongly (argmax over the kept rows of W, `3550-3558`); an empty
    result keeps the whole layer (`3560-3566`).  pmeans_z[l]: (N, K_l); pmeans_W[l]: (K_l, K_{l-1})."""
    if min_cells_upper != 0 and min_cells_upper < 1.0:
        min_cells_upper = max(min_cells_upper * n_cells, 10)
    out = []
    for i, K in enumerate(layer_sizes):
        if i == 0:
            keep = np.arange(K)
        else:
            assignments = np.argmax(pmeans_z[i], axis=1)
            counts = np.array([np.count_nonzero(assignments == a) for a in range(K)])
            keep = np.arange(K)[np.where(counts >= min_cells_upper)[0]]
            if filter_up and len(keep) > 0 and len(out[i - 1]) > 0:
                mat = pmeans_W[i][keep]
                att = [keep[np.argmax(mat[:, f])] for f in out[i - 1]]
                keep = np.unique(list(set(np.unique(att)).intersection(keep)))
        if len(keep) == 0:
            keep = np.arange(K)
        out.append(np.asarray(keep, dtype=int))
    return out


# ---------------------------------------------------------------------------------------------
# iscDEF prior construction (`/root/reference/src/scdef/models/_iscdef.py`), restated loop by loop
# ---------------------------------------------------------------------------------------------
def iscdef_layer_sizes(n_markers, markers_layer, add_root, decay_factor=2, n_layers_schedule=6):
    """`set_layer_sizes` (`_iscdef.py:167-212`) + the decay schedule of `update_model_size` (`_scdef.py:1206-1234`).
    Returns (layer_sizes, layer_names, n_factors_per_marker or None)."""
    if markers_layer == 0:
        n = int(n_markers)
        sizes = [n]
        if int(n_layers_schedule) > 1:
            while len(sizes) < int(n_layers_schedule):
                n = int(np.floor(n / float(decay_factor)))
                sizes.append(n)
                if n == 1:
                    break
            if sizes[-1] > 1:
                sizes[-1] = 1
        names = [f"L{i}" for i in range(len(sizes))]
        names[0] = "marker"
        return sizes, names, None
    n_layers = markers_layer + 1
    if n_layers_schedule == 1:
        n_layers = 1
    sizes, names = [], []
    if n_layers == 1:
        per = int(decay_factor)
        sizes.append(n_markers * per)
        names.append("marker")
    else:
        for layer in range(n_layers):
            rev = (n_layers - 1) - layer
            sizes.append(n_markers * int(decay_factor) ** rev)
            names.append("marker" if layer == n_layers - 1 else f"L{layer}")
        per = int(decay_factor) ** (n_layers - 1)
    if add_root:
        sizes.append(1)
        names.append("root")
    return sizes, names, per


def iscdef_geneset_prior(var_names, markers_dict, marker_names, layer_sizes, markers_layer, n_factors_per_marker,
                         gs_small_scale, gs_big_scale, marker_strength, nonmarker_strength, other_strength,
                         penalize_other=True):
    """`set_geneset_prior` (`_iscdef.py:368-450`) with all layer-0 factors kept: returns (strengths, strengths / gene_sets),
    the Gamma shape and rate matrices of W_0."""
    var_names = np.asarray(var_names)
    K0, G = layer_sizes[0], len(var_names)
    gene_sets = np.ones((K0, G)) * gs_small_scale
    strengths = np.ones((K0, G)) * nonmarker_strength
    kept = np.arange(K0)
    for i, cellgroup in enumerate(marker_names):
        if markers_layer == 0:
            rows = np.where(kept == i)[0]
        else:
            rows = np.where((kept >= i * n_factors_per_marker) & (kept < (i + 1) * n_factors_per_marker))[0]
        if len(rows) == 0:
            continue
        if "other" not in cellgroup:
            for gene in markers_dict[cellgroup]:
                loc = np.where(var_names == gene)[0]
                if len(loc) == 0:
                    continue
                gene_sets[np.ix_(rows, loc)] = gs_big_scale
                strengths[np.ix_(rows, loc)] = marker_strength
        if penalize_other:
            for group in markers_dict:
                if group != cellgroup:
                    for gene in markers_dict[group]:
                        if "other" not in cellgroup:
                            if gene not in markers_dict[cellgroup]:
                                loc = np.where(var_names == gene)[0]
                                if len(loc) == 0:
                                    continue
                                gene_sets[np.ix_(rows, loc)] = 1e-6
                                strengths[np.ix_(rows, loc)] = other_strength
                        else:
                            loc = np.where(var_names == gene)[0]
                            if len(loc) == 0:
                                continue
                            gene_sets[np.ix_(rows, loc)] = 1e-6
                            strengths[np.ix_(rows, loc)] = other_strength
    return strengths, strengths / gene_sets


def iscdef_connectivity_prior(layer_sizes, n_markers, markers_layer, add_root, decay_factor, cn_small_mean, cn_big_mean,
                              cn_small_strength, cn_big_strength):
    """`set_connectivity_prior` (`_iscdef.py:272-366`) with all factors kept: per layer l >= 1 the factors
    (strength_matrix, strength_matrix / max(connectivity_matrix, 1e-12)) that multiply the scDEF W_l shape / rate."""
    n_layers = len(layer_sizes)
    content = n_layers - 1 if (add_root and markers_layer > 0) else n_layers
    connectivity_layers = markers_layer + 1 if (add_root and markers_layer > 0) else n_layers
    out = {}
    for layer_idx in range(1, connectivity_layers):
        n_upper, n_lower = layer_sizes[layer_idx], layer_sizes[layer_idx - 1]
        conn = cn_small_mean * np.ones((n_upper, n_lower))
        stren = cn_small_strength * np.ones((n_upper, n_lower))
        rev = content - 1 - layer_idx
        upper_per, lower_per = int(decay_factor) ** rev, int(decay_factor) ** (rev + 1)
        for i in range(n_markers):
            upper_start, lower_start = i * upper_per, i * lower_per
            for upper_factor in range(upper_start, (i + 1) * upper_per):
                if upper_factor >= n_upper:
                    continue
                block = lower_per // upper_per
                child_start = lower_start + (upper_factor - upper_start) * block
                children = [c for c in range(child_start, child_start + block) if c < n_lower]
                if not children:
                    continue
                conn[upper_factor, children] = cn_big_mean
                stren[upper_factor, children] = cn_big_strength
        out[layer_idx] = (stren, stren / np.maximum(conn, 1e-12))
    return out
