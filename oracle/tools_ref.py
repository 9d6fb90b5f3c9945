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
