"""Post-fit tools that the reference runs through JAX (`scdef.tl`) — device versions.

* `assign_confident`: the Monte-Carlo assignment of every cell to the finest layer at which its winning factor leads
  with posterior confidence (`/root/reference/src/scdef/tools/factor.py:1107-1477`).  The per-layer statistics (two
  passes over `n_samples` draws of `q(z)` per cell) run in `scdef_assign_confident` (`scdef_b200/csrc/tools.cuh`, one
  warp per cell); the cross-layer selection rule and the AnnData fields are host code with the reference's names.
* `set_confident_signatures` / `get_stored_confident_signatures` (`factor.py:441-552, 382-438`): gene signatures with
  posterior confidences for every factor of every layer; the Monte-Carlo gene scores of the upper layers (and the
  per-draw top-k lists of all layers) are formed and reduced on the device by `scdef_signature_scores`.

No CPU fallback: without CUDA or the shared library these raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

STAT_NAMES = ("effect_size", "posterior_sd", "confidence", "winner_mass", "winner_probability", "entropy_confidence")


class ScdefToolsError(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def layer_confidence_stats(mu: torch.Tensor, rho: torch.Tensor, cols, n_samples: int, quantile_level: float,
                           seed: int = 0, stream_id: int = 0, row0: int = 0, eps: Optional[torch.Tensor] = None):
    """Statistics of one layer for every row of the z block halves `mu`, `rho` ((N, sumK) float32 CUDA tensors).

    cols: columns of the kept factors (len >= 2).  Returns (stats (N, 6) float32, argmax (N,) int32) on the device;
    the columns of `stats` are `STAT_NAMES`.  `eps` (n_samples, N, len(cols)) injects the draws (parity tests)."""
    if not (mu.is_cuda and rho.is_cuda):
        raise ScdefToolsError("assign_confident needs CUDA tensors: there is no CPU fallback")
    assert mu.dtype == torch.float32 and rho.dtype == torch.float32 and mu.dim() == 2 and mu.shape == rho.shape
    assert mu.stride(1) == 1 and rho.stride(1) == 1 and mu.stride(0) == rho.stride(0)
    lib = _lib.load()
    N, ld = int(mu.shape[0]), int(mu.stride(0))
    cols_d = torch.as_tensor(np.asarray(cols, dtype=np.int32), device=mu.device)
    stats = torch.empty((N, 6), dtype=torch.float32, device=mu.device)
    argmax = torch.empty((N,), dtype=torch.int32, device=mu.device)
    if eps is not None:
        assert eps.is_cuda and eps.dtype == torch.float32 and eps.is_contiguous()
        assert tuple(eps.shape) == (int(n_samples), N, int(cols_d.numel()))
    st = C.c_void_p(torch.cuda.current_stream(mu.device).cuda_stream)
    with torch.cuda.device(mu.device):
        rc = lib.scdef_assign_confident(_ptr(mu), _ptr(rho), N, ld, _ptr(cols_d), int(cols_d.numel()), int(n_samples),
                                        float(quantile_level), int(seed) & (2**64 - 1), int(stream_id), int(row0),
                                        _ptr(eps), _ptr(stats), _ptr(argmax), st)
    if rc != 0:
        raise ScdefToolsError(f"scdef_assign_confident failed ({rc}): n_cols={int(cols_d.numel())}, n_samples={n_samples}")
    return stats, argmax


def select_layers(confidence_mat, effect_size_mat, posterior_sd_mat, argmax_mat, kept_sizes, tau: float):
    """"Finest layer that clears tau" (factor.py:1381-1451): multi-factor layers are scanned from the finest; cells
    that clear `tau` nowhere fall back to the top-most single-factor layer."""
    N, L = confidence_mat.shape
    is_primary = np.asarray([k >= 2 for k in kept_sizes], dtype=bool)
    has_factors = np.asarray([k > 0 for k in kept_sizes], dtype=bool)
    primary_mask = (confidence_mat >= float(tau)) & is_primary[None, :]
    has_primary = np.any(primary_mask, axis=1)
    finest_primary = np.argmax(primary_mask, axis=1)
    fallback_layer = -1
    for i in range(L - 1, -1, -1):
        if has_factors[i] and not is_primary[i]:
            fallback_layer = i
            break
    raw = np.where(has_primary, finest_primary, fallback_layer)
    any_cand = raw >= 0
    out = {
        "best_layer_idx": np.full(N, -1, dtype=int),
        "best_confidence": np.full(N, np.nan, dtype=float),
        "best_effect_size": np.full(N, np.nan, dtype=float),
        "best_posterior_sd": np.full(N, np.nan, dtype=float),
        "best_factor_slot": np.full(N, -1, dtype=int),
        "depth_score": np.full(N, np.nan, dtype=float),
    }
    if np.any(any_cand):
        idxs = np.where(any_cand)[0]
        sel = raw[any_cand]
        out["best_layer_idx"][idxs] = sel
        out["best_confidence"][idxs] = confidence_mat[idxs, sel]
        out["best_effect_size"][idxs] = effect_size_mat[idxs, sel]
        out["best_posterior_sd"][idxs] = posterior_sd_mat[idxs, sel]
        out["best_factor_slot"][idxs] = argmax_mat[idxs, sel]
    if L <= 1:
        out["depth_score"][out["best_layer_idx"] >= 0] = 0.0
    else:
        v = out["best_layer_idx"] >= 0
        out["depth_score"][v] = out["best_layer_idx"][v].astype(float) / float(L - 1)
    return out


def _categorical(values):
    try:
        import pandas as pd

        return pd.Categorical(np.asarray(values).astype(str))
    except Exception:  # noqa: BLE001 - pandas is optional for the stand-in AnnData
        return np.asarray(values).astype(str)


def assign_confident(model, n_samples: int = 500, tau: float = 0.3, credible_level: float = 0.9,
                     key_added: str = "confident", rng_key=None, device=None) -> None:
    """Drop-in for `scdef.tl.assign_confident` (same arguments, same fields written to `model.adata`).

    `rng_key`: an integer seed (the reference takes a `jax.random` key; `None` -> `model.seed`, factor.py:1237-1238).
    The draws come from the package's Philox stream, keyed by (seed, layer, cell, factor, sample) — not from JAX's
    threefry, so individual draws differ from the reference's while their distribution is the same."""
    if int(n_samples) <= 0:
        raise ValueError("n_samples must be > 0.")
    if not (0.0 <= float(tau) <= 1.0):
        raise ValueError("tau must be in [0, 1].")
    if not (0.0 < float(credible_level) < 1.0):
        raise ValueError("credible_level must be in (0, 1).")
    if not torch.cuda.is_available():
        raise ScdefToolsError("assign_confident needs a CUDA device: there is no CPU fallback")
    seed = int(getattr(model, "seed", 0) or 0) if rng_key is None else int(rng_key)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

    z_params = model.local_params[1]
    mu_full = torch.as_tensor(np.ascontiguousarray(z_params[0], dtype=np.float32), device=dev)
    rho_full = torch.as_tensor(np.ascontiguousarray(z_params[1], dtype=np.float32), device=dev)
    n_cells = int(mu_full.shape[0])
    n_layers = int(model.n_layers)
    layer_sizes = [int(s) for s in model.layer_sizes]
    layer_names = [str(model.layer_names[i]) for i in range(n_layers)]
    n_samples = int(n_samples)

    # data parallelism: this rank holds a block of the cells; the draws of a cell are keyed by its global index
    row0 = 0
    shard = getattr(model, "_shard", None)
    if shard is not None and getattr(shard, "world", 1) > 1:
        from . import dist as _dist

        row0 = int(_dist.shard_bounds(n_cells, shard)[0])

    mats = {name: np.zeros((n_cells, n_layers), dtype=float) for name in STAT_NAMES}
    argmax_mat = np.full((n_cells, n_layers), -1, dtype=int)
    label_mat = np.empty((n_cells, n_layers), dtype=object)
    label_mat[:] = ""

    for layer_idx in range(n_layers):
        start = int(sum(layer_sizes[:layer_idx]))
        kept = np.asarray(model.factor_lists[layer_idx], dtype=int)
        K_k = int(len(kept))
        if K_k == 0:
            continue
        names_arr = np.asarray(model.factor_names[layer_idx], dtype=object)
        if K_k == 1:   # the only factor has all the normalised mass: confidence 1 (factor.py:1270-1281)
            for name in STAT_NAMES:
                mats[name][:, layer_idx] = 0.0 if name == "posterior_sd" else 1.0
            argmax_mat[:, layer_idx] = 0
            label_mat[:, layer_idx] = names_arr[np.zeros(n_cells, dtype=int)]
            continue
        stats, amax = layer_confidence_stats(mu_full, rho_full, start + kept, n_samples, 1.0 - float(credible_level),
                                             seed=seed, stream_id=layer_idx, row0=row0)
        stats = stats.cpu().numpy().astype(float)
        amax = amax.cpu().numpy().astype(int)
        for c, name in enumerate(STAT_NAMES):
            mats[name][:, layer_idx] = stats[:, c]
        argmax_mat[:, layer_idx] = amax
        label_mat[:, layer_idx] = names_arr[amax]

    kept_sizes = [int(len(model.factor_lists[i])) for i in range(n_layers)]
    best = select_layers(mats["confidence"], mats["effect_size"], mats["posterior_sd"], argmax_mat, kept_sizes, tau)
    best_layer_name = np.empty(n_cells, dtype=object)
    best_layer_name[:] = ""
    best_label = np.empty(n_cells, dtype=object)
    best_label[:] = ""
    ok = best["best_layer_idx"] >= 0
    if np.any(ok):
        idxs = np.where(ok)[0]
        sel = best["best_layer_idx"][idxs]
        best_layer_name[idxs] = np.asarray([layer_names[int(k)] for k in sel], dtype=object)
        best_label[idxs] = label_mat[idxs, sel]

    adata = model.adata
    for name in STAT_NAMES:
        adata.obsm[f"{key_added}_{name}"] = mats[name]
    adata.obsm[f"{key_added}_argmax_factor"] = argmax_mat
    for layer_idx in range(n_layers):
        layer_name = layer_names[layer_idx]
        adata.obs[f"{key_added}_confidence_{layer_name}"] = mats["confidence"][:, layer_idx]
        adata.obs[f"{key_added}_argmax_{layer_name}"] = _categorical(label_mat[:, layer_idx])
    adata.obs[f"{key_added}_best_layer"] = _categorical(best_layer_name)
    adata.obs[f"{key_added}_best_layer_idx"] = best["best_layer_idx"]
    adata.obs[f"{key_added}_best_factor_idx"] = best["best_factor_slot"]
    adata.obs[f"{key_added}_factor"] = _categorical(best_label)
    adata.obs[f"{key_added}_best_effect_size"] = best["best_effect_size"]
    adata.obs[f"{key_added}_best_posterior_sd"] = best["best_posterior_sd"]
    adata.obs[f"{key_added}_best_confidence"] = best["best_confidence"]
    adata.obs[f"{key_added}_depth_score"] = best["depth_score"]
    adata.uns[key_added] = {
        "layer_names": list(layer_names),
        "layer_sizes_filtered": kept_sizes,
        "tau": float(tau),
        "n_samples": int(n_samples),
        "credible_level": float(credible_level),
        "quantile_level": float(1.0 - float(credible_level)),
        "metric": "empirical_lower_quantile_winner_runner_up_gap",
        "selection_rule": "finest_layer_clearing_tau",
        "depth_score": "best_layer_index / max(n_layers - 1, 1); single-layer models use 0",
        "fit_revision": int(getattr(model, "_fit_revision", 0)),
    }


# =====================================================================================================================
# set_confident_signatures (`/root/reference/src/scdef/tools/factor.py:441-552`): gene signatures of every factor of every
# layer with per-gene posterior confidences, the DE-style combined score and the weighted-Jaccard stability of each list.
# Layer 0 is closed form (normal approximation, factor.py:1712-1741).  The upper layers need Monte-Carlo gene scores
# (`_collect_hierarchy_mc_scores`, 271-291: `mc_samples` joint draws of every W_l propagated down to the genes, an
# (S, F, G) tensor per layer in the reference); here `scdef_signature_scores` forms them tile by tile on the device and
# reduces them on the spot to exceedance counts and to the ranks of the signature genes (csrc/tools.cuh).
# =====================================================================================================================
def _technical_drop_factors(model):
    factor_obs = model.adata.uns.get("factor_obs") if hasattr(model.adata, "uns") else None      # factor.py:134-139
    if factor_obs is None or "technical" not in getattr(factor_obs, "columns", ()):
        return []
    return factor_obs.index[factor_obs["technical"].astype(bool)].tolist()


def _l0_keep_positions(model, drop_factors):
    drop = set(drop_factors)                                                                      # factor.py:151-159
    return np.array([i for i, n in enumerate(model.factor_names[0]) if n not in drop], dtype=int)


def _layer_term_means(model, layer_idx, l0_keep):
    """`_get_layer_term_means` (factor.py:182-217) from `model.pmeans`."""
    names, fl = model.layer_names, [np.asarray(f, dtype=int) for f in model.factor_lists]
    if layer_idx == 0:
        return np.asarray(model.pmeans[f"{names[0]}W"], dtype=float)[fl[0]]
    t = np.asarray(model.pmeans[f"{names[layer_idx]}W"], dtype=float)[fl[layer_idx]][:, fl[layer_idx - 1]]
    for l in range(layer_idx - 1, 0, -1):
        t = t.dot(np.asarray(model.pmeans[f"{names[l]}W"], dtype=float)[fl[l]][:, fl[l - 1]])
    w0 = np.asarray(model.pmeans[f"{names[0]}W"], dtype=float)[fl[0]][l0_keep]
    return t[:, l0_keep].dot(w0)


def _confidence_mean_score(confidences, means, eps: float = 1e-12):
    confidences, means = np.asarray(confidences, dtype=float), np.asarray(means, dtype=float)   # factor.py:116-131
    return means * -np.log10(np.clip(1.0 - confidences, eps, 1.0))


def _rank_signature(mu, confidences, confidence_threshold, min_effect):
    keep = confidences >= confidence_threshold                                                    # factor.py:311-322
    if min_effect is not None:
        keep = keep & (mu >= min_effect)
    idx = np.where(keep)[0]
    if len(idx) > 0:
        idx = idx[np.argsort(_confidence_mean_score(confidences[idx], mu[idx]))[::-1]]
    return idx


def _lognormal(mu, rho, eps):
    return torch.clamp(torch.exp(mu + torch.exp(rho) * eps), 1e-15, 1e15)                         # jax_utils.py:6-8


def _sig_call(lib, w_mu, w_rho, rows, eps, path, mode, tau=None, counts=None, ref_idx=None, ref_score=None, rank_count=None):
    S, R, G = (int(x) for x in eps.shape)
    F = R if path is None else int(path.shape[1])
    st = C.c_void_p(torch.cuda.current_stream(w_mu.device).cuda_stream)
    with torch.cuda.device(w_mu.device):
        rc = lib.scdef_signature_scores(_ptr(w_mu), _ptr(w_rho), G, _ptr(rows), R, _ptr(eps), _ptr(path), S, F, int(mode),
                                        _ptr(tau), _ptr(counts), _ptr(ref_idx), 0 if ref_idx is None else int(ref_idx.shape[1]),
                                        _ptr(ref_score), _ptr(rank_count), st)
    if rc != 0:
        raise ScdefToolsError(f"scdef_signature_scores failed ({rc}): rows={R}, factors={F}, genes={G}, draws={S}, mode={mode}")


def _jaccard_from_ranks(sig_idx, combined, rank_count):
    """Mean weighted Jaccard (factor.py:41-70, 338-362) from the ranks of the signature genes in every draw: gene i is in
    the draw's top-k list iff fewer than k genes score above it; the list has exactly k genes, so the genes it holds
    that are not in the signature number k - |intersection|."""
    out = []
    rc = rank_count.cpu().numpy()
    for f, ref in enumerate(sig_idx):
        k = len(ref)
        if k == 0:
            out.append(float("nan"))
            continue
        w = np.asarray(combined[f][:k], dtype=float)
        if w.size < k:
            w = np.pad(w, (0, k - w.size), constant_values=0.0)
        if np.all(w <= 0):
            w = np.ones(k, dtype=float)
        w = w / np.sum(w)
        in_top = rc[:, f, :k] < k                                     # (S, k)
        inter = in_top.astype(float) @ w
        n_only = k - in_top.sum(axis=1)
        out.append(float(np.mean(inter / (1.0 + n_only / float(k)))))
    return out


def _ranks_of_signatures(lib, w_mu, w_rho, rows, eps, path, sig_idx):
    F = len(sig_idx)
    kmax = max(1, max((len(s) for s in sig_idx), default=1))
    ref = np.full((F, kmax), -1, dtype=np.int32)
    for f, s in enumerate(sig_idx):
        ref[f, :len(s)] = s
    ref_d = torch.as_tensor(ref, device=w_mu.device)
    S = int(eps.shape[0])
    ref_score = torch.zeros((S, F, kmax), dtype=torch.float32, device=w_mu.device)
    rank = torch.zeros((S, F, kmax), dtype=torch.int32, device=w_mu.device)
    _sig_call(lib, w_mu, w_rho, rows, eps, path, 1, ref_idx=ref_d, ref_score=ref_score)
    _sig_call(lib, w_mu, w_rho, rows, eps, path, 2, ref_idx=ref_d, ref_score=ref_score, rank_count=rank)
    return rank


def set_confident_signatures(model, confidence_threshold: float = 0.9, tau_quantile: float = 0.99,
                             min_effect: Optional[float] = None, mc_samples: int = 100, random_seed: int = 0,
                             device=None, _eps=None):
    """Drop-in for `scdef.tl.set_confident_signatures` (same arguments; same structure written to
    `model.adata.uns['confident_signatures']` and `['factor_signatures']`).

    The N(0,1) draws come from `torch.randn` on the device (generator seeded with `random_seed`), not from JAX's threefry:
    individual draws differ from the reference's, their distribution is the same.  `_eps` injects them (parity tests):
    {"w0": (S, R, G), "upper": {l: (S, F_l, F_{l-1})}, "l0": (S, F_0, G)} float32 CUDA tensors."""
    if not (0.0 < confidence_threshold < 1.0):
        raise ValueError("confidence_threshold must be in (0, 1).")
    if not (0.0 < tau_quantile < 1.0):
        raise ValueError("tau_quantile must be in (0, 1).")
    if mc_samples <= 0:
        raise ValueError("mc_samples must be > 0.")
    if not torch.cuda.is_available():
        raise ScdefToolsError("set_confident_signatures needs a CUDA device: there is no CPU fallback")
    from scipy.stats import norm

    lib = _lib.load()
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    S, L = int(mc_samples), int(model.n_layers)
    names = [str(model.layer_names[i]) for i in range(L)]
    fl = [np.asarray(model.factor_lists[i], dtype=int) for i in range(L)]
    term_names = np.asarray(model.adata.var_names)
    G = int(len(term_names))
    l0_keep = _l0_keep_positions(model, _technical_drop_factors(model))
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(random_seed))

    def randn(key, shape):
        if _eps is not None:
            e = _eps[key[0]] if len(key) == 1 else _eps[key[0]][key[1]]
            assert tuple(e.shape) == tuple(shape) and e.is_cuda and e.dtype == torch.float32
            return e.contiguous()
        return torch.randn(shape, generator=gen, device=dev, dtype=torch.float32)

    W = [(torch.as_tensor(np.ascontiguousarray(model.global_params[2 + l][0], dtype=np.float32), device=dev),
          torch.as_tensor(np.ascontiguousarray(model.global_params[2 + l][1], dtype=np.float32), device=dev)) for l in range(L)]
    w0_mu, w0_rho = W[0]

    # ---- upper layers: path of every draw down to the kept L0 factors (tiny batched products), then the device kernel ----
    upper = {}
    if L > 1 and len(l0_keep) > 0 and all(len(f) > 0 for f in fl):
        rows = torch.as_tensor((fl[0][l0_keep]).astype(np.int32), device=dev)
        eps_w0 = randn(("w0",), (S, len(l0_keep), G))
        paths, path = [], None
        for l in range(1, L):
            mu_l = W[l][0][torch.as_tensor(fl[l], device=dev)][:, torch.as_tensor(fl[l - 1], device=dev)]
            rho_l = W[l][1][torch.as_tensor(fl[l], device=dev)][:, torch.as_tensor(fl[l - 1], device=dev)]
            w_l = _lognormal(mu_l[None], rho_l[None], randn(("upper", l), (S, len(fl[l]), len(fl[l - 1]))))
            path = w_l if path is None else torch.bmm(w_l, path)                                  # factor.py:256-259
            paths.append(path[:, :, torch.as_tensor(l0_keep, device=dev)])
        path_all = torch.cat(paths, dim=1).contiguous()                                           # (S, sum F_l, R)
        offs = np.concatenate([[0], np.cumsum([len(fl[l]) for l in range(1, L)])])
        means_u = [_layer_term_means(model, l, l0_keep) for l in range(1, L)]
        tau_u = np.concatenate([np.quantile(m, tau_quantile, axis=1) for m in means_u]).astype(np.float32)
        counts = torch.zeros((path_all.shape[1], G), dtype=torch.int32, device=dev)
        _sig_call(lib, w0_mu, w0_rho, rows, eps_w0, path_all, 0, tau=torch.as_tensor(tau_u, device=dev), counts=counts)
        conf_u = counts.cpu().numpy().astype(float) / float(S)                                    # factor.py:308
        sig_all = []
        for li, l in enumerate(range(1, L)):
            for f in range(len(fl[l])):
                mu = means_u[li][f]
                sig_all.append(_rank_signature(mu, conf_u[offs[li] + f], confidence_threshold, min_effect))
        comb_all = []
        for li, l in enumerate(range(1, L)):
            for f in range(len(fl[l])):
                idx = sig_all[offs[li] + f]
                comb_all.append(_confidence_mean_score(conf_u[offs[li] + f][idx], means_u[li][f][idx]))
        rank = _ranks_of_signatures(lib, w0_mu, w0_rho, rows, eps_w0, path_all, sig_all)
        jac_all = _jaccard_from_ranks(sig_all, comb_all, rank)
        for li, l in enumerate(range(1, L)):
            sl = slice(offs[li], offs[li + 1])
            upper[l] = (sig_all[sl], [conf_u[offs[li] + f][sig_all[offs[li] + f]] for f in range(len(fl[l]))],
                        comb_all[sl], jac_all[sl])

    cache = {
        "fit_revision": int(getattr(model, "_fit_revision", 0)),
        "params": {"confidence_threshold": float(confidence_threshold), "tau_quantile": float(tau_quantile),
                   "min_effect": None if min_effect is None else float(min_effect), "mc_samples": int(mc_samples),
                   "random_seed": int(random_seed)},
        "by_layer": {},
    }
    signatures_flat = {}
    for l in range(L):
        fnames = list(model.factor_names[l])
        if l == 0:
            means = np.asarray(model.pmeans[f"{names[0]}W"], dtype=float)[fl[0]]                 # factor.py:1712-1741
            sd = np.sqrt(np.maximum(np.asarray(model.pvars[f"{names[0]}W"], dtype=float)[fl[0]], 0.0) + 1e-12)
            sig, confs, comb = [], [], []
            for f in range(len(fl[0])):
                mu = means[f]
                conf = norm.cdf((mu - float(np.quantile(mu, tau_quantile))) / sd[f])
                idx = _rank_signature(mu, conf, confidence_threshold, min_effect)
                sig.append(idx)
                confs.append(conf[idx])
                comb.append(_confidence_mean_score(conf[idx], mu[idx]))
            if len(fl[0]) > 0:
                rows0 = torch.as_tensor(fl[0].astype(np.int32), device=dev)
                eps_l0 = randn(("l0",), (S, len(fl[0]), G))                                       # every factor's own draws
                rank = _ranks_of_signatures(lib, w0_mu, w0_rho, rows0, eps_l0, None, sig)
                jac = _jaccard_from_ranks(sig, comb, rank)
            else:
                jac = []
        elif l in upper:
            sig, confs, comb, jac = upper[l]
        else:
            sig, confs, comb, jac = [], [], [], []
        layer = {"layer_name": names[l], "signatures": {}, "confidences": {}, "combined_scores": {}, "signature_confidences": {}}
        for f, fname in enumerate(fnames):
            genes = term_names[sig[f]].tolist() if f < len(sig) else []
            layer["signatures"][fname] = genes
            layer["confidences"][fname] = np.asarray(confs[f], dtype=float).tolist() if f < len(sig) else []
            layer["combined_scores"][fname] = np.asarray(comb[f], dtype=float).tolist() if f < len(sig) else []
            layer["signature_confidences"][fname] = float(jac[f]) if f < len(jac) else float("nan")
            signatures_flat[fname] = genes
        cache["by_layer"][str(int(l))] = layer
    model.adata.uns["confident_signatures"] = cache
    model.adata.uns["factor_signatures"] = signatures_flat
    return None


def get_stored_confident_signatures(model, layer_idx: int = 0, max_genes: Optional[int] = None, return_confidences: bool = False,
                                    return_combined_scores: bool = False, return_signature_confidences: bool = False):
    """`scdef.tl.get_stored_confident_signatures` (factor.py:382-438): read the cache written above."""
    if layer_idx < 0 or layer_idx >= model.n_layers:
        raise ValueError(f"layer_idx must be in [0, {model.n_layers - 1}].")
    cache = model.adata.uns.get("confident_signatures", None)
    if cache is None:
        raise KeyError("Confident signatures were not precomputed. Run `tl.set_confident_signatures(model)` first.")
    if int(cache.get("fit_revision", -1)) != int(getattr(model, "_fit_revision", 0)):
        raise KeyError("Stored confident signatures are stale for this fitted model. Run `tl.set_confident_signatures(model)` again.")
    layer = cache["by_layer"][str(int(layer_idx))]
    cut = (lambda v: v) if max_genes is None else (lambda v: v[: int(max_genes)])
    out = [{k: cut(list(v)) for k, v in layer["signatures"].items()}]
    if return_confidences:
        out.append({k: cut(np.asarray(v, dtype=float)) for k, v in layer["confidences"].items()})
    if return_combined_scores:
        out.append({k: cut(np.asarray(v, dtype=float)) for k, v in layer["combined_scores"].items()})
    if return_signature_confidences:
        out.append({k: float(v) for k, v in layer.get("signature_confidences", {}).items()})
    return out[0] if len(out) == 1 else tuple(out)
