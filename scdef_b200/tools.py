"""Post-fit tools that the reference runs through JAX (`scdef.tl`) — device versions.

Only `assign_confident` is built: the Monte-Carlo assignment of every cell to the finest layer at which its winning
factor leads with posterior confidence (`/root/reference/src/scdef/tools/factor.py:1107-1477`).  The per-layer
statistics (two passes over `n_samples` draws of `q(z)` per cell) run in `scdef_assign_confident`
(`scdef_b200/csrc/tools.cuh`, one warp per cell); the cross-layer selection rule and the AnnData fields are host code
with the reference's names and shapes.  No CPU fallback: without CUDA or the shared library this raises.
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
