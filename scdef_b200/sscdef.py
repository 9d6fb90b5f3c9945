"""`sscDEF` — the supervised variant on the B200 training path.

What the reference's `sscDEF` (`/root/reference/src/scdef/models/_sscdef.py`) changes about the hot path is three
facts about ONE layer `t` (the layer that carries the cell-type labels of `obs[top_key]`):
  * its z is a constant, one-hot(label) with 1.0 on the annotated population and 1e-3 elsewhere (`484-487`);
  * it has no entropy term and gets no gradient (`530-531`);
  * the per-cell prior strength is floored at 1.0 instead of 0.1 (`505-507`).
The kernels take those as `desc.supervised_layer`, `desc.fixed_top_z`, `desc.alpha_floor`; this class only derives
them from the labels, keeps the variational parameters of layer `t` at the moment-matched one-hot values, and maps the
reference's constructor / `fit` arguments (`79-184`, `1032-1050`) onto `scDEF`.

Hot loop, literal vs intended.  In the reference's loop the pinning is followed by `local_params = self.local_params`
(`_sscdef.py:892-893`), and `self.local_params` is never assigned the updated values: every local update is discarded,
the global pass sees the per-cell blocks of the START of `_learn`, and only the global blocks train
(`tests/test_ref_pin.py::test_reference_sscdef_loop_discards_every_local_update` runs the reference and shows it).
Default here = that literal behaviour (`local_updates="reference"`: the local pass is skipped, which gives the same
parameters and the same recorded losses); `local_updates="keep"` trains the per-cell blocks and pins only layer `t`.
"""
from __future__ import annotations

import logging
from typing import Any, List, Optional, Sequence

import numpy as np

from .model import geometric_ladder, scDEF

_OFF, _ON = 1e-3, 1.0   # `_sscdef.py:41-42`


def _labels_of(adata, key: str) -> np.ndarray:
    obs = adata.obs
    names = list(obs.columns) if hasattr(obs, "columns") else list(obs.keys())
    if key not in names:
        raise KeyError(f"top_key `{key}` not found in adata.obs.")
    col = obs[key]
    vals = np.asarray(col.values if hasattr(col, "values") and not isinstance(col, np.ndarray) else col, dtype=object)
    if any(v is None or (isinstance(v, float) and np.isnan(v)) for v in vals):
        raise ValueError(f"obs[`{key}`] contains missing values; fill or drop them first.")
    return vals.astype(str)


def _lognormal_of_mean(mean):
    """(mu, log sigma) of a LogNormal with this mean and variance mean * 1e-3 (`_sscdef.py:206-214`)."""
    m = np.clip(np.asarray(mean, dtype=np.float64), 1e-3, 10.0)
    v = np.maximum(m * 1e-3, 1e-12)
    return (np.log(m * m / np.sqrt(m * m + v)).astype(np.float32),
            np.log(np.sqrt(np.log1p(v / (m * m)))).astype(np.float32))


class sscDEF(scDEF):
    _Z_OFF_TYPE, _Z_ON = _OFF, _ON
    alpha_floor = 1.0   # `_sscdef.py:505-507`

    @staticmethod
    def _geometric_layer_sizes_no_root(n_factors: int, top_factors: int, n_layers: int) -> List[int]:
        return geometric_ladder(n_factors, top_factors, n_layers, root=False)

    def __init__(self, adata, top_key: str, add_root: bool = True, *, local_updates: str = "reference", **kwargs: Any):
        if local_updates not in ("reference", "keep"):
            raise ValueError("local_updates must be 'reference' (the literal loop of the reference) or 'keep'")
        self.local_updates = local_updates
        self.top_key = str(top_key)
        labels = _labels_of(adata, self.top_key)
        self.population_names = sorted(set(labels.tolist()))
        self.n_populations = len(self.population_names)
        if self.n_populations < 1:
            raise ValueError(f"obs[`{self.top_key}`] must have at least one category.")
        self.population_to_idx = {name: i for i, name in enumerate(self.population_names)}
        self.add_root = bool(add_root)
        if int(kwargs.get("n_layers", 6)) < 2:
            raise ValueError("sscDEF requires n_layers >= 2 (one supervised top layer plus at least one learned "
                             "layer below).")

        sizes = kwargs.get("layer_sizes")
        if sizes is None:
            sizes = geometric_ladder(int(kwargs.get("n_factors", 100)), self.n_populations,
                                     int(kwargs.get("n_layers", 6)), root=False) + ([1] if self.add_root else [])
        else:
            sizes = [int(k) for k in sizes]
            t = len(sizes) - 1 - int(self.add_root)
            if self.add_root and sizes[-1] != 1:
                raise ValueError("With add_root=True, the last entry of layer_sizes must be 1.")
            if t < 0 or sizes[t] != self.n_populations:
                where = "penultimate layer_sizes entry" if self.add_root else "last entry of layer_sizes"
                raise ValueError(f"The {where} must equal the number of categories in obs[`{self.top_key}`] "
                                 f"({self.n_populations}).")
        t = len(sizes) - 1 - int(self.add_root)
        names = kwargs.get("layer_names")
        if names is None:
            names = [f"L{i}" for i in range(t)] + [self.top_key] + (["root"] if self.add_root else [])
        elif len(names) != len(sizes):
            raise ValueError("layer_names must have the same length as layer_sizes.")
        self.supervised_top_layer_idx = t

        onehot = np.full((len(labels), self.n_populations), _OFF, dtype=np.float32)
        onehot[np.arange(len(labels)), [self.population_to_idx[v] for v in labels]] = _ON
        kwargs = dict(kwargs, layer_sizes=sizes, layer_names=names, n_factors=sizes[0],
                      top_factors=self.n_populations, n_layers=len(sizes))
        super().__init__(adata, **kwargs)
        self.logger = logging.getLogger("sscDEF")
        self.logger.setLevel(kwargs.get("logginglevel", logging.INFO))
        self._supervised_top_z = onehot
        self._fixed_top_z = onehot              # what the engine hands to the kernels (desc.fixed_top_z)
        self.set_factor_names()
        self.init_var_params(init_budgets=True, init_alpha=True, nmf_init=False)
        self.set_posterior_means()
        self.set_posterior_variances()

    # ---- the three facts, as the base class and the engine read them -------------------------------------------
    @property
    def supervised_layer(self) -> Optional[int]:
        return self.supervised_top_layer_idx if hasattr(self, "_supervised_top_z") else None

    def _layer_column_range(self, layer_idx: int):
        lo = int(sum(self.layer_sizes[:layer_idx]))
        return lo, lo + int(self.layer_sizes[layer_idx])

    def _pin_supervised_top_z_local_params(self) -> None:
        """Layer t's (mu, log sigma) <- the LogNormal matched to the one-hot labels (`_sscdef.py:221-233`)."""
        if not hasattr(self, "_supervised_top_z") or len(self._local_params) < 2:
            return
        lo, hi = self._layer_column_range(self.supervised_top_layer_idx)
        z = np.array(self._local_params[1])
        z[0][:, lo:hi], z[1][:, lo:hi] = _lognormal_of_mean(self._supervised_top_z)
        self._local_params[1] = z
        self._engine_stale = True

    def init_var_params(self, *args, **kwargs):
        if hasattr(self, "_supervised_top_z") and kwargs.get("init_z") is None:
            init_z: List[Optional[np.ndarray]] = [None] * self.n_layers
            init_z[self.supervised_top_layer_idx] = self._supervised_top_z
            kwargs["init_z"] = init_z
        super().init_var_params(*args, **kwargs)
        self._pin_supervised_top_z_local_params()

    def set_factor_names(self) -> None:
        super().set_factor_names()
        t = getattr(self, "supervised_top_layer_idx", None)
        if t is not None and hasattr(self, "population_names") and len(self.factor_lists[t]) == self.n_populations:
            self.factor_names[t] = [self.population_names[int(i)] for i in self.factor_lists[t]]

    def _summary_key(self) -> str:
        return f"{self.layer_names[self.supervised_top_layer_idx]}z"

    def set_posterior_means(self) -> None:
        super().set_posterior_means()
        if hasattr(self, "_supervised_top_z"):   # `_sscdef.py:1052-1057`
            self.pmeans[self._summary_key()] = np.asarray(self._supervised_top_z, dtype=float)

    def set_posterior_variances(self) -> None:
        super().set_posterior_variances()
        if hasattr(self, "_supervised_top_z"):   # `_sscdef.py:1059-1066`
            self.pvars[self._summary_key()] = np.zeros_like(self._supervised_top_z, dtype=float)

    # ---- training ---------------------------------------------------------------------------------------------
    def _learn(self, *args, **kwargs):
        """`sscDEF._learn` (`_sscdef.py:588-1030`) = `scDEF._learn` with layer t out of `optimize_layers`
        (`240-253`), its columns never moved, and — literal mode — the local updates discarded (`892-893`)."""
        rounds = int(kwargs.get("n_rounds", args[0] if args else 1))
        if rounds <= 1:
            t = int(self.supervised_top_layer_idx)
            wanted = kwargs.get("optimize_layers")
            if wanted is None and kwargs.get("opt_layer") is not None:
                wanted = [int(kwargs["opt_layer"])]
            if wanted is None:
                layers = [l for l in range(self.n_layers - 1)]
            else:
                layers = sorted({int(l) for l in wanted})
                if t in layers:
                    self.logger.warning("Top supervised layer %s is fixed; excluding it from optimize_layers.", t)
            kwargs["optimize_layers"] = [l for l in layers if l != t]
            kwargs["freeze_z_layers"] = sorted({int(l) for l in (kwargs.get("freeze_z_layers") or [])} | {t})
            if self.local_updates == "reference":
                kwargs["update_locals"] = False
        return super()._learn(*args, **kwargs)

    learn = _learn

    def fit(self, root_epochs: int = 0, **kwargs: Any) -> None:
        """`_sscdef.py:1032-1050`: the two-phase schedule of `scDEF.fit`, `root_epochs` defaulting to 10 with a root."""
        kwargs = {k: v for k, v in kwargs.items() if k != "optimize_layers"}
        kwargs["root_epochs"] = 10 if (self.add_root and int(root_epochs) == 0) else int(root_epochs)
        super().fit(**kwargs)
        self._pin_supervised_top_z_local_params()
        self.set_posterior_means()
        self.set_posterior_variances()
