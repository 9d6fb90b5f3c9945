"""`iscDEF` — marker-informed scDEF on the B200 training path.

Mirrors the model-definition part of `/root/reference/src/scdef/models/_iscdef.py`: the layer ladder built from
the marker sets (`167-212`), the gene-set prior on W_0 (`368-450`: a `(K0, G)` matrix of Gamma shapes/rates — both
likelihood paths take matrix-valued W_0 priors), the block-diagonal connectivity prior
on the upper W_l (`272-366`), the marker-aware factor names (`452-526`) and `fit` with the frozen-root second
phase (`936-1134`).  Not built: the Scanpy `score_genes` initialisation of z (`588-934`; Scanpy is not part of
this image — `fit` then uses the ordinary initialisation and says so), the refit warm start and the
diagnostics-based `filter_factors` override (`528-586`), which are host-side library code around the hot path.
"""
from __future__ import annotations

import logging
from typing import Any, List, Mapping, Optional, Sequence

import numpy as np

from .model import scDEF


class iscDEF(scDEF):
    def __init__(self, adata, markers_dict: Mapping[str, Sequence[str]], add_other: Optional[int] = 0,
                 markers_layer: Optional[int] = 0, add_root: Optional[bool] = None,
                 cn_small_mean: float = 1.0, cn_big_mean: float = 10.0, cn_small_strength: float = 0.1,
                 cn_big_strength: float = 1.0, gs_small_scale: float = 1.0, gs_big_scale: float = 10.0,
                 marker_strength: float = 1.0, nonmarker_strength: float = 0.1, other_strength: float = 0.1,
                 penalize_other: bool = True, **kwargs: Any):
        self.markers_dict = markers_dict
        self.penalize_other = bool(penalize_other)
        self.add_other = int(add_other) if add_other is not None else 0
        if self.add_other < 0:
            raise ValueError("add_other must be >= 0.")
        self.markers_layer = int(markers_layer)
        if self.markers_layer == 0:
            if add_root is True:
                raise ValueError("add_root=True requires markers_layer > 0.")
            self.add_root = False
        else:
            self.add_root = True if add_root is None else bool(add_root)
        self.cn_small_mean, self.cn_big_mean = cn_small_mean, cn_big_mean
        self.cn_small_strength, self.cn_big_strength = cn_small_strength, cn_big_strength
        self.gs_small_scale, self.gs_big_scale = gs_small_scale, gs_big_scale
        self.marker_strength, self.nonmarker_strength, self.other_strength = marker_strength, nonmarker_strength, other_strength
        self.other_names = [f"other{i}" for i in range(self.add_other)]
        self.marker_names = list(self.markers_dict.keys()) + self.other_names
        self.n_markers = len(self.marker_names)
        self.decay_factor = kwargs.pop("decay_factor", 2)
        self.n_layers_schedule = kwargs.pop("n_layers", 6)
        if "use_brd" not in kwargs and self.markers_layer != 0:
            kwargs["use_brd"] = True
        self.use_brd = kwargs.get("use_brd", True)
        if self.use_brd:
            self.nonmarker_strength = 1.0           # `_iscdef.py:140-141`
        self.set_layer_sizes()
        kwargs.setdefault("hierarchy_weight", 0.25)
        schedule = self.n_layers_schedule
        super().__init__(adata, layer_sizes=self.layer_sizes, layer_names=self.layer_names,
                         n_layers=len(self.layer_sizes), **kwargs)
        self.n_layers_schedule = schedule
        level = self.logger.level
        self.logger = logging.getLogger("iscDEF")
        self.logger.setLevel(level)
        self._preserve_factor_names_on_annotate = True
        self.set_geneset_prior(gs_big_scale=self.gs_big_scale, gs_small_scale=self.gs_small_scale,
                               marker_strength=self.marker_strength, nonmarker_strength=self.nonmarker_strength,
                               other_strength=self.other_strength)
        self.init_var_params()
        self.set_posterior_means()
        self.set_posterior_variances()
        self.set_factor_names()

    # ---- ladder ---------------------------------------------------------------------------------------
    def set_layer_sizes(self) -> None:
        """`_iscdef.py:167-212`: markers at layer 0 -> halve (`decay_factor`) up to a single top factor; markers at layer
        t > 0 -> `n_markers * decay^(t-l)` factors at layer l (a block per marker), optional width-1 root on top."""
        df = int(self.decay_factor)
        if self.markers_layer == 0:
            sizes = [self.n_markers]
            n = self.n_markers
            while len(sizes) < int(self.n_layers_schedule):
                n = int(np.floor(n / float(self.decay_factor)))
                sizes.append(n)
                if n == 1:
                    break
            if len(sizes) > 1 and sizes[-1] > 1:
                sizes[-1] = 1
            names = [f"L{i}" for i in range(len(sizes))]
            names[0] = "marker"
        else:
            depth = 1 if self.n_layers_schedule == 1 else self.markers_layer + 1
            if depth == 1:
                self.n_factors_per_marker = df
                sizes, names = [self.n_markers * df], ["marker"]
            else:
                sizes = [self.n_markers * df ** (depth - 1 - l) for l in range(depth)]
                names = [f"L{l}" for l in range(depth - 1)] + ["marker"]
                self.n_factors_per_marker = df ** (depth - 1)
            if self.add_root:
                sizes.append(1)
                names.append("root")
        self.layer_sizes, self.layer_names, self.n_layers = sizes, names, len(sizes)

    def _hierarchy_content_layers(self) -> int:
        """Layers that carry marker blocks (the optional root excluded), `_iscdef.py:730-735`."""
        return self.n_layers - 1 if (self.add_root and self.markers_layer > 0) else self.n_layers

    # ---- priors ----------------------------------------------------------------------------------------
    def update_model_priors(self, update_alpha_from_cov: bool = True) -> None:
        super().update_model_priors(update_alpha_from_cov=update_alpha_from_cov)
        if self.markers_layer != 0 and self.n_layers > 1:
            self.set_connectivity_prior(cn_small_strength=self.cn_small_strength, cn_big_strength=self.cn_big_strength,
                                        cn_small_mean=self.cn_small_mean, cn_big_mean=self.cn_big_mean)
        self.set_geneset_prior(gs_big_scale=self.gs_big_scale, gs_small_scale=self.gs_small_scale,
                               marker_strength=self.marker_strength, nonmarker_strength=self.nonmarker_strength,
                               other_strength=self.other_strength)

    def set_connectivity_prior(self, cn_small_mean=1e-3, cn_big_mean=1.0, cn_small_strength=1.0, cn_big_strength=0.1) -> None:
        """`_iscdef.py:272-366`: an upper factor of marker i is strongly connected (mean `cn_big_mean`) to its own
        `decay_factor` children in the layer below and weakly (`cn_small_mean`) to everything else."""
        self.cn_small_mean, self.cn_big_mean = cn_small_mean, cn_big_mean
        self.cn_small_strength, self.cn_big_strength = cn_small_strength, cn_big_strength
        df = int(self.decay_factor)
        top = self.markers_layer + 1 if (self.add_root and self.markers_layer > 0) else self.n_layers
        content = self._hierarchy_content_layers()
        for l in range(1, top):
            upper = np.asarray(self.factor_lists[l], dtype=int)
            lower = np.asarray(self.factor_lists[l - 1], dtype=int)
            mean = np.full((len(upper), len(lower)), float(cn_small_mean))
            strength = np.full((len(upper), len(lower)), float(cn_small_strength))
            per_up = df ** (content - 1 - l)          # factors per marker in the upper layer
            children = (df ** (content - l)) // per_up
            # original upper factor u (marker i, sub-index j) owns the original lower factors [u*children, (u+1)*children)
            own = (lower[None, :] // children) == upper[:, None]
            own &= (upper[:, None] < self.n_markers * per_up)
            mean[own] = float(cn_big_mean)
            strength[own] = float(cn_big_strength)
            self.w_priors[l][0] = (np.asarray(self.w_priors[l][0]) * strength).astype(np.float32)
            self.w_priors[l][1] = (np.asarray(self.w_priors[l][1]) * strength / np.maximum(mean, 1e-12)).astype(np.float32)

    def set_geneset_prior(self, gs_small_scale=1.0, gs_big_scale=100.0, marker_strength=10.0, nonmarker_strength=0.1,
                          other_strength=0.1) -> None:
        """`_iscdef.py:368-450`: W_0 ~ Gamma(strength, strength / scale) per (factor, gene): scale `gs_big_scale` on a
        typed factor's own markers, `gs_small_scale` elsewhere, and — `penalize_other` — 1e-6 on other groups' markers."""
        self.gs_big_scale, self.gs_small_scale = gs_big_scale, gs_small_scale
        self.marker_strength, self.nonmarker_strength, self.other_strength = marker_strength, nonmarker_strength, other_strength
        K0, G = int(self.layer_sizes[0]), int(self.n_genes)
        scale = np.full((K0, G), float(gs_small_scale))
        strength = np.full((K0, G), float(nonmarker_strength))
        var_index = np.asarray(self.adata.var.index if getattr(self.adata.var, "index", None) is not None else self.adata.var_names)
        col = {g: i for i, g in enumerate(var_index.tolist())}
        self.marker_gene_locs = []
        kept = np.asarray(self.factor_lists[0], dtype=int)
        for i, group in enumerate(self.marker_names):
            if self.markers_layer == 0:
                rows = np.where(kept == i)[0]
            else:
                rows = np.where((kept >= i * self.n_factors_per_marker) & (kept < (i + 1) * self.n_factors_per_marker))[0]
            if len(rows) == 0:
                continue
            typed = "other" not in group
            own = set(self.markers_dict[group]) if typed else set()
            if typed:
                for gene in self.markers_dict[group]:
                    if gene not in col:
                        self.logger.warning(f"Did not find gene {gene} for set {group} in AnnData object.")
                        continue
                    self.marker_gene_locs.append(np.array([col[gene]]))
                    scale[rows, col[gene]] = float(gs_big_scale)
                    strength[rows, col[gene]] = float(marker_strength)
            if self.penalize_other:
                for other_group, genes in self.markers_dict.items():
                    if other_group == group:
                        continue
                    for gene in genes:
                        if gene in col and gene not in own:
                            scale[rows, col[gene]] = 1e-6
                            strength[rows, col[gene]] = float(other_strength)
        self.gene_sets, self.strengths = scale, strength
        prior = [strength.astype(np.float32), (strength / scale).astype(np.float32)]
        if self.n_layers == 1:
            self.w_priors = [prior]
        else:
            self.w_priors[0] = prior
        self._engine_stale = True

    # ---- names -----------------------------------------------------------------------------------------
    def _refresh_top_factor_names(self) -> None:
        if self.markers_layer != 0 and len(self.factor_names) > self.markers_layer:
            self._top_factor_names = list(self.factor_names[self.markers_layer])
        elif len(self.factor_names) > 0:
            self._top_factor_names = list(self.factor_names[-1])

    def _block_names(self, idx: int, per_marker: int) -> List[str]:
        kept = set(int(f) for f in self.factor_lists[idx])
        out = []
        for m, name in enumerate(self.marker_names):
            n_kept = sum(1 for f in range(m * per_marker, (m + 1) * per_marker) if f in kept)
            out += [f"{name}_{self.layer_names[idx]}_{j}" for j in range(n_kept)]
        return out

    def set_factor_names(self) -> None:
        """`_iscdef.py:452-526`."""
        if not hasattr(self, "marker_names") or not hasattr(self, "factor_lists"):
            return super().set_factor_names()
        df = int(self.decay_factor)
        names = []
        for idx in range(self.n_layers):
            generic = [f"{self.layer_names[idx]}_{i}" for i in range(len(self.factor_lists[idx]))]
            if self.markers_layer == 0:
                names.append([str(self.marker_names[i]) for i in self.factor_lists[idx]] if idx == 0 else generic)
            elif self.n_layers == 1:
                names.append(self._block_names(idx, self.n_factors_per_marker))
            elif idx == self.markers_layer:
                top = getattr(self, "_top_factor_names", None)
                if top is not None and len(top) == len(self.factor_lists[idx]):
                    names.append(list(top))
                else:
                    names.append([m for i, m in enumerate(self.marker_names) if i in self.factor_lists[idx]])
            elif self.add_root and idx == self.n_layers - 1:
                names.append(generic)
            else:
                names.append(self._block_names(idx, df ** (self._hierarchy_content_layers() - 1 - idx)))
        self.factor_names = names
        self._refresh_top_factor_names()

    # ---- training --------------------------------------------------------------------------------------
    def fit(self, nmf_init: bool = False, max_cells_init: int = 1024, z_init_concentration: float = 10.0,
            z_init_from_score_genes: bool = True, root_epochs: int = 0, **kwargs: Any) -> None:
        """First fit (`_iscdef.py:936-1134`): all layers but the root, then `root_epochs` on the root alone (10 by default
        when there is a root).  The Scanpy marker-score initialisation of z is not available here."""
        for k in ("z_init_score_temperature", "z_init_other_mass", "z_init_other_mode", "score_genes_layer", "score_genes_kwargs"):
            kwargs.pop(k, None)
        if nmf_init:
            self.logger.warning("iscDEF does not use nmf_init.")
        if z_init_from_score_genes:
            self.logger.info("scanpy.tl.score_genes initialisation of z is outside this package; using the default initialisation.")
        if self.add_root and int(root_epochs) == 0:
            root_epochs = 10
        super().fit(nmf_init=False, max_cells_init=max_cells_init, z_init_concentration=z_init_concentration,
                    root_epochs=int(root_epochs), **kwargs)

    def filter_factors(self, *args, **kwargs):
        if kwargs.get("upper_only", False):
            return super().filter_factors(*args, **kwargs)
        raise NotImplementedError("iscDEF.filter_factors relies on scdef.tools.factor_diagnostics (out of scope here)")
