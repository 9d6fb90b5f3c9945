"""Run the reference's OWN, UNMODIFIED Python source on the torch-backed jax stand-in.

TEST INFRASTRUCTURE (see oracle/__init__.py): nothing under `scdef_b200/` may import this.

`reference_runtime()` imports, straight from `/root/reference/src/scdef` (read-only, never copied):
  * `utils/jax_utils.py`     (lognormal_sample / lognormal_entropy / gamma_logpdf, 6-46)
  * `models/_scdef.py`       (scDEF: load_adata, update_model_priors, init_var_params, elbo 1823-2146,
                              batch_elbo 2148-2181, _get_or_build_learn_grad_fns 2808-2856, _learn 2923-3392)
  * `models/_sscdef.py`, `models/_iscdef.py`
with `jax`, `optax`, `tensorflow_probability`, `anndata`, `seaborn`, `matplotlib` resolved to
`oracle/jaxshim.py`.  The package `__init__` files are NOT executed (they pull in the plotting and
tools trees, which need scanpy / graphviz): the parents `scdef`, `scdef.utils`, `scdef.models` are
bare namespace modules whose `__path__` points into the reference tree.

This is what pins `oracle/elbo_ref.py` / `oracle/host_ref.py` (tests/test_ref_pin.py) and what
`tests/golden/make_golden_ref.py` runs to produce `tests/golden/ref_*.npz`.  It only exists where
`/root/reference` does (this container, not the GPU box): `available()` says which.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import jaxshim

REFERENCE_SRC = os.environ.get("SCDEF_REFERENCE_SRC", "/root/reference/src")
_PKG = "scdef"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, _PKG, "models", "_scdef.py"))


def _namespace(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    return m


class ReferenceRuntime:
    """Handles to the reference's classes and functions, valid inside `reference_runtime()`."""

    def __init__(self, mods, dtype):
        self.dtype = dtype
        self.shim = mods
        self.jnp = mods["jax.numpy"]
        self.random = mods["jax.random"]
        self.jax_utils = importlib.import_module(f"{_PKG}.utils.jax_utils")
        self._scdef_mod = importlib.import_module(f"{_PKG}.models._scdef")
        self.scDEF = self._scdef_mod.scDEF
        self.AnnData = jaxshim.AnnData

    @property
    def sscDEF(self):
        return importlib.import_module(f"{_PKG}.models._sscdef").sscDEF

    @property
    def iscDEF(self):
        return importlib.import_module(f"{_PKG}.models._iscdef").iscDEF

    @property
    def tools_factor(self):
        """`scdef.tools.factor` (assign_confident, factor.py:1107-1477): needs inert `scanpy` / `matplotlib.pyplot`."""
        name = f"{_PKG}.tools.factor"
        if name not in sys.modules:
            sys.modules.setdefault(f"{_PKG}.tools", _namespace(f"{_PKG}.tools", os.path.join(REFERENCE_SRC, _PKG, "tools")))
            self._extra = getattr(self, "_extra", [])
            if "scanpy" not in sys.modules:
                sys.modules["scanpy"] = types.ModuleType("scanpy")
                self._extra.append("scanpy")
            mpl = sys.modules["matplotlib"]
            if not hasattr(mpl, "__path__"):
                mpl.__path__ = []
            if "matplotlib.pyplot" not in sys.modules:
                sys.modules["matplotlib.pyplot"] = mpl.pyplot = types.ModuleType("matplotlib.pyplot")
                self._extra.append("matplotlib.pyplot")
        return importlib.import_module(name)

    def source_lines(self, obj):
        import inspect

        lines, start = inspect.getsourcelines(obj)
        return inspect.getsourcefile(obj), start, start + len(lines) - 1

    # -- helpers the pin tests share ---------------------------------------------------------
    def adata(self, X, batch=None, batch_key="batch"):
        import pandas as pd

        X = np.asarray(X, dtype=np.float64)
        obs = pd.DataFrame(index=[f"cell{i}" for i in range(X.shape[0])])
        if batch is not None:
            obs[batch_key] = pd.Categorical([f"b{int(b)}" for b in np.asarray(batch)])
        return jaxshim.AnnData(X, obs=obs)

    def to_numpy(self, tree):
        if isinstance(tree, (list, tuple)):
            return [self.to_numpy(t) for t in tree]
        return np.asarray(tree.detach() if isinstance(tree, torch.Tensor) else tree)


@contextlib.contextmanager
def reference_runtime(dtype=torch.float64):
    if not available():
        raise RuntimeError(f"reference source not found under {REFERENCE_SRC}")
    root = os.path.join(REFERENCE_SRC, _PKG)
    parents = {
        _PKG: _namespace(_PKG, root),
        f"{_PKG}.utils": _namespace(f"{_PKG}.utils", os.path.join(root, "utils")),
        f"{_PKG}.models": _namespace(f"{_PKG}.models", os.path.join(root, "models")),
    }
    before = set(sys.modules)
    dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only: no __pycache__ there
    with jaxshim.installed(dtype=dtype, extra=parents):
        rt = None
        try:
            rt = ReferenceRuntime(sys.modules, dtype)
            yield rt
        finally:
            sys.dont_write_bytecode = dont_write
            for name in getattr(rt, "_extra", []):
                sys.modules.pop(name, None)
            for name in set(sys.modules) - before:
                if name == _PKG or name.startswith(_PKG + "."):
                    sys.modules.pop(name, None)
