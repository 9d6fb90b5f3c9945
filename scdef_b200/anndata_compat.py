"""Minimal AnnData stand-in.

`anndata` is not installed in this image.  The reference's constructor takes an
`AnnData` (`_scdef.py:483-485`); `scdef_b200.scDEF` accepts a real one when importable and this
duck-typed container otherwise (`.X`, `.layers`, `.obs`, `.var`, `.uns`, `.obsm`, `.shape`,
`.obs_names`, `.var_names`, `.copy()`).  `X` may be a NumPy array, a SciPy CSR matrix or a
torch tensor (host or CUDA — a CUDA tensor is used in place, never copied to the host).
"""
from __future__ import annotations

import copy

import numpy as np

try:  # pragma: no cover - not available in this image
    from anndata import AnnData as _RealAnnData
except Exception:  # noqa: BLE001
    _RealAnnData = None


class _Frame(dict):
    """Tiny column store standing in for a pandas DataFrame (`obs` / `var`)."""

    def __init__(self, data=None, index=None):
        super().__init__(data or {})
        self.index = np.asarray(index) if index is not None else None

    @property
    def columns(self):
        return list(self.keys())

    def copy(self):
        return _Frame({k: np.array(v, copy=True) for k, v in self.items()}, self.index)


class MiniAnnData:
    def __init__(self, X, obs=None, var=None, uns=None, obsm=None, layers=None,
                 obs_names=None, var_names=None):
        self.X = X
        n, g = X.shape
        self.obs_names = np.asarray(obs_names if obs_names is not None else [f"cell{i}" for i in range(n)]) \
            if n <= 2_000_000 else None
        self.var_names = np.asarray(var_names if var_names is not None else [f"g{i}" for i in range(g)])
        self.obs = obs if isinstance(obs, _Frame) else _Frame(obs or {}, self.obs_names)
        self.var = var if isinstance(var, _Frame) else _Frame(var or {}, self.var_names)
        self.uns = dict(uns or {})
        self.obsm = dict(obsm or {})
        self.layers = dict(layers or {})
        self.raw = None

    @property
    def shape(self):
        return tuple(self.X.shape)

    @property
    def n_obs(self):
        return self.shape[0]

    @property
    def n_vars(self):
        return self.shape[1]

    def copy(self):
        # the count matrix is shared, not duplicated (it can be tens of GB on the device)
        out = MiniAnnData(self.X, self.obs.copy(), self.var.copy(), copy.deepcopy(self.uns),
                          dict(self.obsm), dict(self.layers), self.obs_names, self.var_names)
        return out

    def __str__(self):
        return f"MiniAnnData object with n_obs x n_vars = {self.shape[0]} x {self.shape[1]}"


def is_anndata(obj) -> bool:
    if _RealAnnData is not None and isinstance(obj, _RealAnnData):
        return True
    return isinstance(obj, MiniAnnData)
