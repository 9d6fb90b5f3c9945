"""TEST INFRASTRUCTURE — CPU restatement (NumPy) of `scdef.tl.assign_confident`, the post-fit Monte-Carlo
assignment of cells to layers (`/root/reference/src/scdef/tools/factor.py:1107-1477`).  PARITY PIN: the reference's own
`assign_confident` is executed unmodified under `oracle/jaxshim.py` and must agree with this module on the draws it
consumed (`tests/test_ref_pin.py::test_restated_assign_confident_equals_the_reference_execution`); the statements below
follow the reference line by line with the N(0,1) draws injected (`eps[s, n, k]` replaces
`random.normal(sample_keys[s], mu_k.shape)`, factor.py:1288, 1318).

Only `tests/` may import this module; the product path is `scdef_b200/tools.py` -> `scdef_assign_confident`.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def layer_stats(mu_k: np.ndarray, log_std_k: np.ndarray, eps: np.ndarray, credible_level: float = 0.9):
    """Per-cell statistics of ONE layer (factor.py:1264-1379), float32 like the reference's JAX arrays.

    mu_k, log_std_k: (N, K) posterior parameters of the kept factors; eps: (S, N, K).
    Returns dict of (N,) arrays: effect_size, posterior_sd, confidence, winner_mass, winner_probability,
    entropy_confidence, argmax (int)."""
    mu_k = np.asarray(mu_k, dtype=F32)
    sigma_k = np.exp(np.asarray(log_std_k, dtype=F32))            # 1265
    S, N, K = eps.shape
    if K == 1:                                                     # 1270-1281: the only factor has all the mass
        one = np.ones(N)
        return dict(effect_size=one.copy(), posterior_sd=np.zeros(N), confidence=one.copy(), winner_mass=one.copy(),
                    winner_probability=one.copy(), entropy_confidence=one.copy(), argmax=np.zeros(N, dtype=int))
    counts = np.zeros((N, K), dtype=F32)
    sum_zhat = np.zeros((N, K), dtype=F32)
    zhats = []
    for s in range(S):                                             # pass 1, lax.scan 1286-1302
        z = np.exp(mu_k + sigma_k * eps[s].astype(F32))            # 1289
        zhat = (z / np.clip(np.sum(z, axis=-1, keepdims=True, dtype=F32), F32(1e-30), None)).astype(F32)  # 1290
        a = np.argmax(zhat, axis=-1)                               # 1291 (first maximum on ties)
        counts[np.arange(N), a] += F32(1.0)                        # 1292-1293
        sum_zhat = (sum_zhat + zhat).astype(F32)
        zhats.append(zhat)
    p = counts / F32(S)                                            # 1303
    mean_zhat = sum_zhat / F32(S)                                  # 1304
    winner = np.argmax(mean_zhat, axis=-1)                         # 1305
    rows = np.arange(N)
    gaps = np.zeros((S, N), dtype=F32)
    wms = np.zeros((S, N), dtype=F32)
    for s in range(S):                                             # pass 2, same draws, 1317-1329
        zhat = zhats[s]
        wm = zhat[rows, winner]
        others = zhat.copy()
        others[rows, winner] = -np.inf                             # 1322
        gaps[s] = wm - others.max(axis=-1)                         # 1323-1324
        wms[s] = wm
    q_level = 1.0 - float(credible_level)                          # 1333
    conf = np.quantile(gaps.astype(np.float64), q_level, axis=0)   # 1334 (linear interpolation, jnp default)
    mean_gap = gaps.mean(axis=0, dtype=np.float64)                 # 1335
    sd_gap = gaps.std(axis=0, dtype=np.float64)                    # 1336 (population SD)
    wm_mean = wms.mean(axis=0, dtype=np.float64)                   # 1337
    p64 = p.astype(np.float64)
    safe_p = np.clip(p64, 1e-30, 1.0)                              # 1365
    H = -np.sum(p64 * np.log(safe_p), axis=-1)                     # 1366
    return dict(
        effect_size=np.clip(mean_gap, -1.0, 1.0),                  # 1343
        posterior_sd=np.clip(sd_gap, 0.0, 1.0),                    # 1351
        confidence=np.clip(conf, 0.0, 1.0),                        # 1354
        winner_mass=np.clip(wm_mean, 0.0, 1.0),                    # 1358
        winner_probability=np.clip(p64.max(axis=-1), 0.0, 1.0),    # 1361
        entropy_confidence=np.clip(1.0 - H / float(np.log(K)), 0.0, 1.0),  # 1367
        argmax=winner.astype(int),
    )


def select_layers(confidence_mat: np.ndarray, effect_size_mat, posterior_sd_mat, argmax_mat, kept_sizes, tau: float):
    """Cross-layer selection, "finest layer that clears tau" (factor.py:1381-1451), restated cell by cell.

    confidence_mat etc.: (N, L); kept_sizes[l] = len(model.factor_lists[l]).
    Returns dict: best_layer_idx, best_confidence, best_effect_size, best_posterior_sd, best_factor_slot, depth_score."""
    N, L = confidence_mat.shape
    fallback = -1                                                   # 1408-1412: top-most single-factor layer
    for i in range(L - 1, -1, -1):
        if kept_sizes[i] == 1:
            fallback = i
            break
    out = dict(best_layer_idx=np.full(N, -1, dtype=int), best_confidence=np.full(N, np.nan),
               best_effect_size=np.full(N, np.nan), best_posterior_sd=np.full(N, np.nan),
               best_factor_slot=np.full(N, -1, dtype=int), depth_score=np.full(N, np.nan))
    for n in range(N):
        chosen = fallback
        for l in range(L):                                          # 1401-1406: finest multi-factor layer clearing tau
            if kept_sizes[l] >= 2 and confidence_mat[n, l] >= float(tau):
                chosen = l
                break
        if chosen < 0:                                              # 1419: no candidate at all
            continue
        out["best_layer_idx"][n] = chosen                           # 1432-1441
        out["best_confidence"][n] = confidence_mat[n, chosen]
        out["best_effect_size"][n] = effect_size_mat[n, chosen]
        out["best_posterior_sd"][n] = posterior_sd_mat[n, chosen]
        out["best_factor_slot"][n] = argmax_mat[n, chosen]
        out["depth_score"][n] = 0.0 if L <= 1 else chosen / float(L - 1)   # 1444-1451
    return out


# =================================================================================================
# `scdef.tl.set_confident_signatures` (factor.py:29-131, 182-363, 441-552, 1652-1768) restated.
# PARITY PIN: `tests/test_ref_pin.py::test_restated_confident_signatures_equal_the_reference_execution` runs the
# reference's own function under `oracle/jaxshim.py` and replays the draws it consumed into these functions.
# Genes are indices into `adata.var_names`; factors are positions in `model.factor_lists[l]`.
# =================================================================================================
def lognormal_draw(mu, rho, eps):
    """`jax_utils.lognormal_sample` (jax_utils.py:6-8) with the N(0,1) draw injected."""
    return np.clip(np.exp(np.asarray(mu, float) + np.exp(np.asarray(rho, float)) * np.asarray(eps, float)), 1e-15, 1e15)


def layer_term_means(pmeans_W, factor_lists, l0_keep, layer_idx):
    """`_get_layer_term_means` (factor.py:182-217): mean loadings of the kept factors of a layer on the genes, upper
    layers propagated through the posterior means of the W's; `l0_keep` = positions (in factor_lists[0]) of the L0
    factors that are not dropped as technical."""
    fl = [np.asarray(f, dtype=int) for f in factor_lists]
    if layer_idx == 0:
        return np.asarray(pmeans_W[0], float)[fl[0]]
    t = np.asarray(pmeans_W[layer_idx], float)[fl[layer_idx]][:, fl[layer_idx - 1]]
    for l in range(layer_idx - 1, 0, -1):
        t = t @ np.asarray(pmeans_W[l], float)[fl[l]][:, fl[l - 1]]
    w0 = np.asarray(pmeans_W[0], float)[fl[0]][np.asarray(l0_keep, int)]
    return t[:, np.asarray(l0_keep, int)] @ w0


def hierarchy_scores_draw(W_mu, W_rho, factor_lists, l0_keep, eps):
    """`_hierarchy_gene_scores_draw` (factor.py:220-268): ONE joint draw of all W_l, propagated down to the genes.
    eps[l]: N(0,1) array of the kept sub-matrix of layer l.  Returns {layer: (n_factors_l, G)} for l >= 1."""
    fl = [np.asarray(f, dtype=int) for f in factor_lists]
    keep = np.asarray(l0_keep, int)
    w0 = lognormal_draw(np.asarray(W_mu[0])[fl[0]], np.asarray(W_rho[0])[fl[0]], eps[0])[keep]
    scores, path = {}, None
    for l in range(1, len(fl)):
        w = lognormal_draw(np.asarray(W_mu[l])[fl[l]][:, fl[l - 1]], np.asarray(W_rho[l])[fl[l]][:, fl[l - 1]], eps[l])
        path = w if path is None else w @ path
        scores[l] = path[:, keep] @ w0
    return scores


def confidence_mean_score(conf, means, eps=1e-12):
    """`_confidence_mean_score` (factor.py:116-131)."""
    return np.asarray(means, float) * -np.log10(np.clip(1.0 - np.asarray(conf, float), eps, 1.0))


def rank_signature(mu, confidences, confidence_threshold, min_effect=None):
    """Keep mask + DE-style ordering shared by both branches (factor.py:311-322, 1725-1736)."""
    keep = confidences >= confidence_threshold
    if min_effect is not None:
        keep = keep & (mu >= min_effect)
    idx = np.where(keep)[0]
    if len(idx) > 0:
        idx = idx[np.argsort(confidence_mean_score(confidences[idx], mu[idx]))[::-1]]
    return idx


def confident_from_scores(term_means, mc_scores, confidence_threshold, tau_quantile, min_effect=None):
    """`_confident_signatures_from_mc_scores` (factor.py:294-335): per factor the ordered gene indices and their
    exceedance confidences `mean_s [score_s > tau]`, tau = quantile of the factor's mean loadings."""
    out = []
    for f in range(term_means.shape[0]):
        mu = term_means[f]
        tau = float(np.quantile(mu, tau_quantile))
        conf = np.mean(np.asarray(mc_scores[:, f, :], float) > tau, axis=0)
        idx = rank_signature(mu, conf, confidence_threshold, min_effect)
        out.append((idx, conf[idx]))
    return out


def layer0_confident(term_means, term_vars, confidence_threshold, tau_quantile, min_effect=None):
    """Layer 0 of `get_confident_signatures` (factor.py:1712-1741): normal approximation of P(W_kg > tau)."""
    from scipy.stats import norm

    sd = np.sqrt(np.maximum(term_vars, 0.0) + 1e-12)
    out = []
    for f in range(term_means.shape[0]):
        mu = term_means[f]
        tau = float(np.quantile(mu, tau_quantile))
        conf = norm.cdf((mu - tau) / sd[f])
        idx = rank_signature(mu, conf, confidence_threshold, min_effect)
        out.append((idx, conf[idx]))
    return out


def weighted_jaccard(ref_idx, weights, draw_idx):
    """`_weighted_signature_jaccard` (factor.py:41-70) on gene indices."""
    k = len(ref_idx)
    if k == 0:
        return 1.0
    w = np.asarray(weights[:k], float)
    if w.size < k:
        w = np.pad(w, (0, k - w.size))
    if np.all(w <= 0):
        w = np.ones(k)
    w = w / w.sum()
    draw = set(int(i) for i in draw_idx)
    inter = float(sum(w[i] for i, g in enumerate(ref_idx) if int(g) in draw))
    n_only = len(draw - set(int(g) for g in ref_idx))
    return inter / (1.0 + float(n_only) / float(k))


def jaccard_from_scores(sig_idx, combined, scores_of):
    """`_signature_jaccard_from_mc_scores` / `_compute_signature_jaccard_confidences` (factor.py:338-362, 73-113):
    mean weighted Jaccard of every draw's top-k list against the signature.  `scores_of(f)` -> (S, G) draws of factor f."""
    out = []
    for f, ref in enumerate(sig_idx):
        k = len(ref)
        if k == 0:
            out.append(float("nan"))
            continue
        sc = np.asarray(scores_of(f), float)
        out.append(float(np.mean([weighted_jaccard(ref, combined[f], np.argsort(sc[s])[::-1][:k]) for s in range(sc.shape[0])])))
    return out
