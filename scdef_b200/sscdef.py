"""`sscDEF` — supervised scDEF on the B200 training path.

Mirrors `/root/reference/src/scdef/models/_sscdef.py`: the cell-type labels in `obs[top_key]` fix
the top layer's z to one-hot(label) (1.0 on the annotated population, 1e-3 elsewhere, `484-487`),
that layer has no entropy term (`530-531`), the cell-specific prior strength is floored at 1.0
instead of 0.1 (`505-507`), and its variational parameters stay pinned (`221-233, 892-893`).
The kernels implement all of this through `desc.supervised_layer / fixed_top_z / alpha_floor`.

Reference quirk (documented, not reproduced by default): in the reference's hot loop the pinning is
followed by `local_params = self.local_params` (`_sscdef.py:892-893`), which resets ALL local blocks
to their values at the start of `_learn` after every step, i.e. the local updates are discarded and
only the global blocks learn.  Here only the supervised columns are pinned (the evident intent);
`reference_local_reset=True` reproduces the literal behaviour (equivalent to `update_locals=False`).
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from .model import scDEF


class sscDEF(scDEF):
    _Z_OFF_TYPE = 1e-3
    _Z_ON = 1.0
    alpha_floor = 1.0  # `_sscdef.py:505-507`

    @staticmethod
    def _geometric_layer_sizes_no_root(n_factors: int, top_factors: int, n_layers: int) -> List[int]:
        """The scDEF geometric ladder without the appended root (`_sscdef.py:44-77`)."""
        n_layers = max(2, int(n_layers))
        min_k, top_k = float(n_factors), float(top_factors)
        ratio = top_k / max(min_k, 1.0)
        if n_layers == 2:
            out = [max(1, int(round(min_k))), int(top_factors)]
            out[0] = max(out[0], out[1])
            return out
        out: List[int] = []
        prev = None
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
        return out

    def __init__(self, adata, top_key: str, add_root: bool = True, *, reference_local_reset: bool = False,
                 **kwargs: Any):
        obs = adata.obs
        cols = obs.columns if hasattr(obs, "columns") else list(obs.keys())
        if top_key not in list(cols):
            raise KeyError(f"top_key `{top_key}` not found in adata.obs.")
        self.top_key = str(top_key)
        self.reference_local_reset = bool(reference_local_reset)
        col = obs[self.top_key]
        raw = np.asarray(col.values if hasattr(col, "values") and not isinstance(col, np.ndarray) else col)
        if any(v is None or (isinstance(v, float) and np.isnan(v)) for v in raw):
            raise ValueError(f"obs[`{self.top_key}`] contains missing values; fill or drop them first.")
        labels = raw.astype(str)
        self.population_names: List[str] = sorted(set(labels.tolist()))
        self.n_populations = len(self.population_names)
        if self.n_populations < 1:
            raise ValueError(f"obs[`{self.top_key}`] must have at least one category.")
        self.population_to_idx: Dict[str, int] = {n: i for i, n in enumerate(self.population_names)}

        n_factors = int(kwargs.get("n_factors", 100))
        n_layers_schedule = int(kwargs.get("n_layers", 6))
        if n_layers_schedule < 2:
            raise ValueError("sscDEF requires n_layers >= 2 (one supervised top layer plus "
                             "at least one learned layer below).")
        self.add_root = bool(add_root)
        layer_sizes = kwargs.get("layer_sizes")
        if layer_sizes is not None:
            layer_sizes = [int(x) for x in layer_sizes]
            if self.add_root:
                if int(layer_sizes[-1]) != 1:
                    raise ValueError("With add_root=True, the last entry of layer_sizes must be 1.")
                if int(layer_sizes[-2]) != self.n_populations:
                    raise ValueError("With add_root=True, the penultimate layer_sizes entry must "
                                     f"equal the number of categories in obs[`{self.top_key}`] ({self.n_populations}).")
            elif int(layer_sizes[-1]) != self.n_populations:
                raise ValueError("The last entry of layer_sizes must equal the number of "
                                 f"categories in obs[`{self.top_key}`] ({self.n_populations}).")
        else:
            layer_sizes = self._geometric_layer_sizes_no_root(n_factors, self.n_populations, n_layers_schedule)
            if self.add_root:
                layer_sizes.append(1)
        layer_names = kwargs.get("layer_names")
        if layer_names is None:
            n_below = len(layer_sizes) - 1 - int(self.add_root)
            layer_names = [f"L{i}" for i in range(n_below)] + [self.top_key] + (["root"] if self.add_root else [])
        elif len(layer_names) != len(layer_sizes):
            raise ValueError("layer_names must have the same length as layer_sizes.")
        kwargs = dict(kwargs)
        kwargs.update(layer_sizes=layer_sizes, layer_names=layer_names, n_factors=int(layer_sizes[0]),
                      top_factors=self.n_populations, n_layers=len(layer_sizes))
        self.supervised_top_layer_idx = len(layer_sizes) - 2 if self.add_root else len(layer_sizes) - 1
        self._labels = labels
        super().__init__(adata, **kwargs)
        level = self.logger.level
        self.logger = logging.getLogger("sscDEF")
        self.logger.setLevel(level)
        if int(self.layer_sizes[self.supervised_top_layer_idx]) != self.n_populations:
            raise ValueError("Top layer size must match the number of supervised populations.")
        self._supervised_top_z = self._build_supervised_top_z()
        self._fixed_top_z = self._supervised_top_z          # read by the engine (desc.fixed_top_z)
        self.set_factor_names()
        self.init_var_params(init_budgets=True, init_alpha=True, init_z=self._build_init_z_list(), nmf_init=False)
        self._pin_supervised_top_z_local_params()
        self.set_posterior_means()
        self.set_posterior_variances()

    # the engine reads these two through the base class
    @property
    def supervised_layer(self):
        return getattr(self, "supervised_top_layer_idx", None) if hasattr(self, "_supervised_top_z") else None

    def set_factor_names(self) -> None:
        super().set_factor_names()
        if hasattr(self, "population_names") and hasattr(self, "supervised_top_layer_idx"):
            top = int(self.supervised_top_layer_idx)
            if len(self.factor_lists[top]) == self.n_populations:
                self.factor_names[top] = [self.population_names[int(i)] for i in self.factor_lists[top]]

    def _build_supervised_top_z(self) -> np.ndarray:
        k_top = int(self.layer_sizes[self.supervised_top_layer_idx])
        z = np.full((self.n_cells, k_top), self._Z_OFF_TYPE, dtype=np.float32)
        for i, lab in enumerate(self._labels):
            z[i, self.population_to_idx[lab]] = self._Z_ON
        return z

    @staticmethod
    def _gamma_params_from_mean(mean):
        mean = np.clip(np.asarray(mean, dtype=np.float64), 1e-3, 10.0)
        v = np.maximum(mean * 1e-3, 1e-12)
        return (np.log(mean**2 / np.sqrt(mean**2 + v)).astype(np.float32),
                np.log(np.sqrt(np.log(1.0 + v / mean**2))).astype(np.float32))

    def _layer_column_range(self, layer_idx: int):
        start = int(sum(self.layer_sizes[:layer_idx]))
        return start, start + int(self.layer_sizes[layer_idx])

    def _pin_supervised_top_z_local_params(self) -> None:
        if len(self._local_params) < 2:
            return
        start, end = self._layer_column_range(self.supervised_top_layer_idx)
        mu, rho = self._gamma_params_from_mean(self._supervised_top_z)
        z = np.array(self._local_params[1])
        z[0][:, start:end] = mu
        z[1][:, start:end] = rho
        self._local_params[1] = z
        self._engine_stale = True

    def _build_init_z_list(self) -> List[Optional[np.ndarray]]:
        init_z: List[Optional[np.ndarray]] = [None] * self.n_layers
        init_z[self.supervised_top_layer_idx] = self._supervised_top_z
        return init_z

    def _supervised_optimize_layers(self, optimize_layers: Optional[Sequence[int]]) -> List[int]:
        if optimize_layers is None:
            return list(range(self.n_layers - 1))
        layers = sorted({int(i) for i in optimize_layers})
        top = self.supervised_top_layer_idx
        if top in layers:
            self.logger.warning("Top supervised layer %s is fixed; excluding it from optimize_layers.", top)
            layers = [i for i in layers if i != top]
        return layers

    def _learn(self, *args, **kwargs):
        if int(kwargs.get("n_rounds", args[0] if args else 1)) <= 1:
            opt = kwargs.get("optimize_layers")
            if opt is None and kwargs.get("opt_layer") is not None:
                opt = [int(kwargs["opt_layer"])]
            kwargs["optimize_layers"] = self._supervised_optimize_layers(opt)
            fz = sorted({int(i) for i in (kwargs.get("freeze_z_layers") or [])} | {int(self.supervised_top_layer_idx)})
            kwargs["freeze_z_layers"] = fz      # the pinned columns never move (`_sscdef.py:892`)
            if self.reference_local_reset:
                kwargs["update_locals"] = False  # literal reference behaviour, see the module docstring
        return super()._learn(*args, **kwargs)

    learn = _learn

    def fit(self, root_epochs: int = 0, **kwargs: Any) -> None:
        kwargs = dict(kwargs)
        kwargs.pop("optimize_layers", None)
        if self.add_root and int(root_epochs) == 0:
            root_epochs = 10
        kwargs["root_epochs"] = int(root_epochs)
        self._pin_supervised_top_z_local_params()
        # scDEF.fit re-initialises the variational parameters; the init_var_params override below
        # warm-starts and re-pins the supervised layer like the constructor does
        super().fit(**kwargs)
        self._pin_supervised_top_z_local_params()
        self.set_posterior_means()
        self.set_posterior_variances()

    def init_var_params(self, *args, **kwargs):
        if kwargs.get("init_z") is None and hasattr(self, "_supervised_top_z"):
            kwargs["init_z"] = self._build_init_z_list()
        super().init_var_params(*args, **kwargs)
        if hasattr(self, "_supervised_top_z"):
            self._pin_supervised_top_z_local_params()

    def set_posterior_means(self) -> None:
        super().set_posterior_means()
        if hasattr(self, "_supervised_top_z"):
            self.pmeans[f"{self.layer_names[self.supervised_top_layer_idx]}z"] = np.asarray(self._supervised_top_z, dtype=float)

    def set_posterior_variances(self) -> None:
        super().set_posterior_variances()
        if hasattr(self, "_supervised_top_z"):
            self.pvars[f"{self.layer_names[self.supervised_top_layer_idx]}z"] = np.zeros_like(self._supervised_top_z, dtype=float)
