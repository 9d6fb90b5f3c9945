"""`scDEF` — host-side mirror of the reference model class for the training hot path.

Same constructor / `fit()` / `_learn()` (`learn()` alias) surface, kwarg names, defaults and
error behaviour as `/root/reference/src/scdef/models/_scdef.py` (constructor 106-212, `fit`
2346-2737 first-fit path, `pretrain` 2739-2780, `_learn` 2923-3392), but every device-side
piece — ELBO, gradients, Adam, clipping, posterior summaries — runs in the sm_100a kernels
behind `include/scdef_b200.h`.  No JAX, no CPU fallback for the training step.

State handed back to the rest of the library keeps the reference layout:
`local_params`, `global_params` (lists of float32 NumPy arrays, stacked [mu, log sigma]),
`pmeans`, `pvars`, `elbos`, `step_sizes`, `adata.uns["entropy_annealing_trace"]`, ...

Out of scope here: refit warm-start, factor filtering, annotation, plotting.
"""
from __future__ import annotations

import logging
from typing import Any, List, Optional, Sequence

import numpy as np

from . import anndata_compat, dist as _dist

try:  # progress bar is optional
    from tqdm import tqdm as _tqdm
except Exception:  # noqa: BLE001
    _tqdm = None


def _moment_match(m, v):
    """LogNormal (mu, log sigma) with mean m and variance v (`_scdef.py:1496-1505` etc.)."""
    m = np.asarray(m, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    return (np.log(m**2 / np.sqrt(m**2 + v)).astype(np.float32),
            np.log(np.sqrt(np.log1p(v / m**2))).astype(np.float32))


def geometric_ladder(n_factors: int, top_factors: int, n_layers: int, root: bool = True) -> List[int]:
    """Layer widths from `n_factors` down to `top_factors` in `n_layers` geometric steps, strictly decreasing where
    rounding allows, plus a width-1 root when `root` and `top_factors > 1`.
    MIRRORED ON PURPOSE: the integers must equal `scDEF._geometric_layer_sizes` (`_scdef.py:580-628`; without the root
    `sscDEF._geometric_layer_sizes_no_root`, `_sscdef.py:44-77`) for every input — pinned by the reference's own values
    (`/root/reference/tests/test_sscdef.py:55-83`) in tests/test_host.py."""
    n_layers = max(2, int(n_layers))
    k_lo, k_hi = float(n_factors), int(top_factors)
    if n_layers == 2:
        widths = [max(1, int(round(k_lo)), k_hi), k_hi]
    else:
        step = k_hi / max(k_lo, 1.0)
        widths: List[int] = []
        for l in range(n_layers - 1):
            k = max(1, int(round(k_lo * step ** (l / float(n_layers - 1)))))
            if widths and k >= widths[-1]:
                k = max(1, widths[-1] - 1)
            widths.append(k)
        widths.append(k_hi)
        for i in range(n_layers - 2, -1, -1):   # never narrower than the layer above
            widths[i] = max(widths[i], widths[i + 1])
    if root and k_hi > 1:
        widths.append(1)
    return widths


class scDEF:
    """Single-cell Deep Exponential Families — B200-native training path."""

    alpha_floor = 0.1  # `_scdef.py:2068` (sscDEF: 1.0)
    supervised_layer = None

    def __init__(
        self,
        adata,
        counts_layer: Optional[str] = None,
        batch_key: Optional[str] = None,
        seed: Optional[int] = 42,
        n_factors: Optional[int] = 100,
        top_factors: int = 1,
        n_layers: Optional[float] = 6,
        layer_sizes: Optional[list] = None,
        layer_names: Optional[list] = None,
        logginglevel: Optional[int] = logging.INFO,
        alpha: Optional[float] = 1.0,
        shrinkage_shape: Optional[float] = 1.0,
        shrinkage_rate: Optional[float] = 1.0,
        shrinkage_mean: Optional[float] = 1.0,
        top_alpha: Optional[float] = 1.0,
        factor_shape: Optional[float] = 0.1,
        brd_strength: Optional[float] = 1.0,
        brd_mean: Optional[float] = 1.0,
        use_brd: Optional[bool] = True,
        cell_scale_shape: Optional[float] = 1.0,
        gene_scale_shape: Optional[float] = 1.0,
        batch_cpal: Optional[str] = "Dark2",
        layer_cpal: Optional[str] = "tab10",
        lightness_mult: Optional[float] = 0.15,
        set_alpha_from_cov: Optional[bool] = True,
        hierarchy_weight: Optional[float] = 1.0,
        marginalize_alpha: Optional[bool] = False,
        *,
        device: Optional[str] = None,
        x_placement: str = "device",
        noise_coupling: bool = True,
        lik_impl: int = 0,
        process_group=None,
        lazy_adam: bool = True,
        epoch_reshard: bool = True,
    ):
        # --- extensions of this package (keyword-only, all optional) ---
        self.device = device
        # "device": dense counts resident in HBM; "csr": CSR counts resident in HBM (a minibatch is densified per
        # step); "host": dense counts in host memory, streamed per step
        if x_placement not in ("device", "device_u16", "csr", "host"):
            raise ValueError("x_placement must be 'device', 'device_u16', 'csr' or 'host'")
        self.x_placement = x_placement
        self.noise_coupling = bool(noise_coupling)
        self.lik_impl = int(lik_impl)
        self.lazy_adam = bool(lazy_adam)  # exact lazy form of the dense local Adam (bit-identical)
        self.epoch_reshard = bool(epoch_reshard)  # data parallel: position-based ownership (dist.epoch_layout)
        self._shard = _dist.ShardInfo.from_group(process_group)
        self.max_steps_per_learn = None  # benchmark hook: stop the hot loop after this many steps
        self._step_hook = None           # benchmark hook: called as hook(step_index, engine) around each step
        self._engine = None

        self.n_cells, self.n_genes = adata.shape
        self.n_batches = 1
        self.batches = [""]
        self.batch_key = batch_key
        self.seed = seed
        self.batch_cpal, self.layer_cpal, self.lightness_mult = batch_cpal, layer_cpal, lightness_mult
        self.logger = logging.getLogger("scDEF")
        self.logger.setLevel(logginglevel)
        self.counts_layer = counts_layer
        self.load_adata(adata, layer=counts_layer, batch_key=batch_key)

        self.top_alpha = top_alpha
        self.alpha = alpha
        self.factor_shape = factor_shape
        self.brd = brd_strength
        self.brd_mean = brd_mean
        self.shrinkage_shape = shrinkage_shape
        self.shrinkage_rate = shrinkage_rate
        self.shrinkage_mean = shrinkage_mean
        self.cell_scale_shape = cell_scale_shape
        self.gene_scale_shape = gene_scale_shape
        self.use_brd = use_brd
        self.set_alpha_from_cov = set_alpha_from_cov
        self.hierarchy_weight = float(hierarchy_weight)
        self.marginalize_alpha = bool(marginalize_alpha)
        self._marginalize_alpha_init = bool(marginalize_alpha)

        if n_layers is None:
            n_layers = 6.0
        self.n_layers_schedule = n_layers
        k0 = int(n_factors if n_factors is not None else 100)
        if k0 < 1:
            raise ValueError("n_factors must be >= 1.")
        self.top_factors = int(top_factors)
        if self.top_factors < 1:
            raise ValueError("top_factors must be >= 1.")
        if self.top_factors > k0:
            raise ValueError(
                "top_factors must be <= n_factors (layer sizes are non-increasing "
                "from L0 toward the top layer)."
            )
        self.n_factors = k0

        if layer_sizes is not None:
            self.layer_sizes = [int(x) for x in layer_sizes]
            self.n_factors = int(layer_sizes[0])
            self.n_layers = len(self.layer_sizes)
        else:
            self.update_model_size(n_layers=int(self.n_layers_schedule))

        self.layer_names = layer_names if layer_names is not None else [f"L{i}" for i in range(self.n_layers)]
        self.factor_lists = [np.arange(size) for size in self.layer_sizes]
        self.set_factor_names()
        self.update_model_priors()
        self.init_var_params(nmf_init=False)
        self.set_posterior_means()
        self.set_posterior_variances()
        self.elbos: List[list] = []
        self.step_sizes: List[float] = []
        self._has_fit = False
        self._fit_revision = 0

    def __repr__(self):
        out = f"scDEF object with {self.n_layers} layers"
        out += "\n\tLayer names: " + ", ".join(str(n) for n in self.layer_names)
        out += "\n\tLayer sizes: " + ", ".join(str(len(f)) for f in self.factor_lists)
        out += "\n\talpha parameter: " + str(self.alpha)
        if self.use_brd:
            out += "\n\tUsing BRD"
        out += "\n\tNumber of batches: " + str(self.n_batches)
        out += "\nContains " + str(self.adata)
        return out

    # ------------------------------------------------------------------------------------------
    # data ingestion — `load_adata`, `_scdef.py:482-578`
    # ------------------------------------------------------------------------------------------
    def load_adata(self, adata, layer=None, batch_key=None):
        if not anndata_compat.is_anndata(adata):
            raise TypeError("adata must be an instance of AnnData.")
        self.adata = adata.copy()
        X = self.adata.X if layer is None else self.adata.layers[layer]
        import scipy.sparse as sp

        self._X_is_torch = hasattr(X, "is_cuda")
        if self.x_placement == "csr":   # keep the counts sparse (the reference densifies, `_scdef.py:491-496`)
            if self._X_is_torch:
                raise TypeError("x_placement='csr' takes a scipy sparse or NumPy count matrix")
            X = sp.csr_matrix(X, dtype=np.float32)
            X.sort_indices()
        elif self.x_placement == "device_u16":   # raw counts as uint16 in HBM (a quarter of the reference's float64 copy)
            import torch

            if sp.issparse(X):
                X = X.toarray()
            if not self._X_is_torch:
                Xn = np.asarray(X)
                if Xn.size and (Xn.min() < 0 or Xn.max() > 65535 or np.any(Xn != np.rint(Xn))):
                    raise ValueError("x_placement='device_u16' needs integer counts in [0, 65535]")
                X = torch.from_numpy(np.ascontiguousarray(Xn.astype(np.uint16)))
            elif X.dtype not in (torch.uint16, torch.int16):
                X = X.to(torch.uint16)
            if not torch.cuda.is_available():
                raise RuntimeError("x_placement='device_u16' needs a CUDA device; there is no CPU fallback")
            X = X.to(self.device or f"cuda:{self._shard.local_rank}")   # the statistics below are computed on the device
            self._X_is_torch = True
        else:
            if sp.issparse(X):
                X = X.toarray()
            if not self._X_is_torch:
                X = np.ascontiguousarray(np.asarray(X), dtype=np.float32)
        self.X = X  # float32 (the reference keeps float64 on the host; counts are exact in f32)
        stats = _dist.count_statistics(X, self._batch_labels(batch_key), self._shard)
        self.n_cells_total = stats["n_cells_total"]
        self.batch_lib_sizes = stats["lib"]
        self.batch_lib_ratio = stats["batch_lib_ratio"]
        self.gene_ratio = stats["gene_ratio"]
        self.gene_ratio_init = stats["gene_ratio_init"]
        self.gene_means, self.gene_vars = stats["gene_means"], stats["gene_vars"]
        self.batch_id = stats["batch_id"]
        self.batches = stats["batches"]
        self.n_batches = len(self.batches)
        self._median_lib = stats["median_lib"]
        self._mean_lib = stats["mean_lib"]
        if batch_key is not None and self.n_batches > 0 and stats["has_key"]:
            self.logger.info(f"Found {self.n_batches} values for `{batch_key}` in data: {self.batches}")

    def _batch_labels(self, batch_key):
        if batch_key is None:
            return None
        obs = self.adata.obs
        cols = obs.columns if hasattr(obs, "columns") else list(obs.keys())
        if batch_key not in list(cols):
            return None
        col = obs[batch_key]
        return np.asarray(col.values if hasattr(col, "values") and not isinstance(col, np.ndarray) else col)

    @property
    def batch_indices_onehot(self):
        oh = np.zeros((self.n_cells, self.n_batches), dtype=np.float32)
        oh[np.arange(self.n_cells), np.asarray(self.batch_id)] = 1.0
        return oh

    # ------------------------------------------------------------------------------------------
    # model size / priors — `_scdef.py:580-628, 1175-1287`
    # ------------------------------------------------------------------------------------------
    def _geometric_layer_sizes(self, n_layers: int) -> List[int]:
        return geometric_ladder(self.n_factors, self.top_factors, n_layers, root=True)

    def update_model_size(self, max_n_factors=None, n_layers=None, layer_sizes=None,
                          use_decay_factor_schedule: bool = False):
        if layer_sizes is not None:
            sizes = [int(x) for x in layer_sizes]
            for i in range(1, len(sizes)):  # non-increasing, consecutive duplicates collapsed
                sizes[i] = min(sizes[i], sizes[i - 1])
            sizes = [s for i, s in enumerate(sizes) if i == 0 or s != sizes[i - 1]]
            self.layer_sizes = sizes
        elif use_decay_factor_schedule:
            raise NotImplementedError("decay-factor schedule belongs to iscDEF (out of scope here)")
        else:
            if n_layers is None:
                n_layers = self.n_layers_schedule
            self.layer_sizes = self._geometric_layer_sizes(int(n_layers))
            self.n_factors = int(self.layer_sizes[0])
        self.n_layers = len(self.layer_sizes)
        self.layer_names = [f"L{i}" for i in range(self.n_layers)]
        self.factor_lists = [np.arange(size) for size in self.layer_sizes]
        self.set_factor_names()

    def set_factor_names(self):
        self.factor_names = [
            [f"{self.layer_names[idx]}_{i}" for i in range(len(self.factor_lists[idx]))]
            for idx in range(self.n_layers)
        ]

    def update_model_priors(self, update_alpha_from_cov: bool = True):
        fs = self.factor_shape
        self.factor_shapes = ([1.0] + [fs] * (self.n_layers - 1)) if self.use_brd else [fs] * self.n_layers
        self.w_priors = []
        for idx in range(self.n_layers):
            if idx == 0:
                self.w_priors.append([self.factor_shapes[0], self.factor_shapes[0]])
            else:
                m = np.clip(np.ones((self.layer_sizes[idx], self.layer_sizes[idx - 1])) * self.factor_shapes[idx],
                            1e-12, 1e12).astype(np.float32)
                self.w_priors.append([m.copy(), m.copy()])
        self.cell_alpha_factor = self.batch_lib_sizes / float(self._median_lib)
        if self.set_alpha_from_cov and update_alpha_from_cov:
            self.alpha = float(self._median_lib) / float(self.layer_sizes[0]) * self.hierarchy_weight

    # ------------------------------------------------------------------------------------------
    # variational parameters — `init_var_params`, `_scdef.py:1463-1732`
    # ------------------------------------------------------------------------------------------
    def init_var_params(self, init_budgets=True, init_alpha=True, init_z=None, init_w=None,
                        init_brd=None, init_ard=None, init_gene_scale=None, nmf_init=False,
                        z_init_concentration=0.05, device_init=False, **kwargs):
        """Cold / warm start of every block.  The reference draws its Gamma variates with JAX
        keys; here NumPy's Generator(seed) — distribution-level equivalence.

        device_init=True (opt-in): the per-cell z layers without an `init_z` — N x sum K Gamma draws, the
        only part whose cost grows with the number of cells — are drawn on the GPU by `scdef_init_gamma_block` (Philox
        stream; a cell's draws do not depend on the sharding) and copied into the host arrays."""
        if nmf_init:
            raise NotImplementedError("nmf_init is host-side preprocessing outside the hot path")
        rng = np.random.default_rng(None if self.seed is None else int(self.seed) + 1000 * self._shard.rank)
        grng = np.random.default_rng(None if self.seed is None else int(self.seed))  # replicated draws
        N, Ks, L = self.n_cells, self.layer_sizes, self.n_layers

        def layer_init(val, l):
            if val is None:
                return None
            if isinstance(val, (list, tuple)):
                return None if l >= len(val) or val[l] is None else np.asarray(val[l], dtype=np.float32)
            return np.asarray(val, dtype=np.float32) if l == 0 else None

        if init_budgets:
            m = np.clip((np.asarray(self.batch_lib_sizes, dtype=np.float64) / self._mean_lib)[:, None], 1e-3, 1e2)
            local = [np.stack(_moment_match(m, m / 10.0))]
            if init_gene_scale is not None:
                m = np.asarray(init_gene_scale, dtype=np.float32)
                if m.ndim == 1:
                    m = np.tile(m[None, :], (max(int(self.n_batches), 1), 1))
                elif m.shape[0] == 1 and int(self.n_batches) > 1:
                    m = np.tile(m, (int(self.n_batches), 1))
            else:
                m = 1.0 / np.asarray(self.gene_ratio_init)
            m = np.clip(m, 1e-6, 1e6)
            glob = [np.stack(_moment_match(m, m / 10.0))]
        else:
            local = [np.asarray(self._local_params[0])]
            glob = [np.asarray(self._global_params[0])]

        m_brd = grng.gamma(100.0, self.brd_mean / 100.0, size=(Ks[0], 1))
        if init_brd is not None:
            m_brd = np.asarray(init_brd, dtype=np.float64).reshape(Ks[0], 1)
        glob.append(np.stack(_moment_match(m_brd, m_brd**2 / 100.0)))

        z_mu, z_rho = [], []
        for l in range(L):
            a, clip = (z_init_concentration, 1e-3) if l == 0 else (1.0, 1e-3)
            if l == L - 1:
                a = 100.0
            zi = layer_init(init_z, l)
            if zi is None and device_init:
                mu, rho = self._device_gamma_block(N, Ks[l], a, clip, 4.0, 100.0 if l == 0 else 10.0, stream_id=l)
                z_mu.append(mu)
                z_rho.append(rho)
                continue
            if zi is not None:
                m = np.clip(zi, clip, 1e6).astype(np.float64)
                m = np.clip(rng.gamma(z_init_concentration, m / z_init_concentration), clip, 1e6)
                v = m**2 / 100.0
            else:
                m = np.clip(rng.standard_gamma(a, size=(N, Ks[l])) / a, clip, 4.0)
                v = m / 100.0 if l == 0 else m / 10.0
            mu, rho = _moment_match(m, v)
            z_mu.append(mu)
            z_rho.append(rho)
        local.append(np.stack([np.hstack(z_mu), np.hstack(z_rho)]))

        for l in range(L):
            out = self.n_genes if l == 0 else Ks[l - 1]
            wi = layer_init(init_w, l)
            if wi is not None:
                m = np.clip(wi.astype(np.float64), 1e-3, 1e8)
                v = m**2 / 100.0
            else:
                m = np.ones((Ks[l], out)) / Ks[l] * np.asarray(self.w_priors[l][0]) / np.asarray(self.w_priors[l][1])
                if l < L - 1:
                    m = grng.gamma(10.0, m / 10.0)
                v = m / 100.0 if l == 0 else m / 10.0
            glob.append(np.stack(_moment_match(m, v)))

        m = np.full((Ks[0], 1), self.shrinkage_shape / self.shrinkage_rate, dtype=np.float64)
        v = m**2 / 10.0
        if init_ard is not None:
            m = np.asarray(init_ard, dtype=np.float64).reshape(Ks[0], 1)
            v = m**2 / 100.0
        glob.append(np.stack(_moment_match(m, v)))
        m = np.full((1, 1), float(self.alpha if init_alpha else np.asarray(self.pmeans["alpha"]).reshape(-1)[0]))
        glob.append(np.stack(_moment_match(m, m**2 / 10.0)))

        self._local_params = [np.ascontiguousarray(p, dtype=np.float32) for p in local]
        self._global_params = [np.ascontiguousarray(p, dtype=np.float32) for p in glob]
        self._engine_stale = True

    def _device_gamma_block(self, n_rows, n_cols, shape, clip_lo, clip_hi, var_div, stream_id):
        """(mu, rho) of an (n_rows, n_cols) block drawn by `scdef_init_gamma_block`; NumPy arrays."""
        import ctypes as C

        import torch

        from . import _lib

        if not torch.cuda.is_available():
            raise RuntimeError("device_init=True needs a CUDA device (there is no CPU fallback for it; use device_init=False)")
        lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        out = torch.empty((2, int(n_rows), int(n_cols)), dtype=torch.float32, device=dev)
        seed = 0 if self.seed is None else int(self.seed)
        row0 = int(_dist.shard_bounds(self.n_cells, self._shard)[0]) if self._shard.world > 1 else 0   # global index of this rank's first cell
        with torch.cuda.device(dev):
            rc = lib.scdef_init_gamma_block(C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()), int(n_rows), int(n_cols),
                                            int(n_cols), float(shape), float(clip_lo), float(clip_hi), float(var_div),
                                            seed & (2**64 - 1), int(stream_id), row0,
                                            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"scdef_init_gamma_block failed ({rc})")
        host = out.cpu().numpy()
        return host[0], host[1]

    @property
    def local_params(self):
        return self._local_params

    @local_params.setter
    def local_params(self, value):
        self._local_params = [np.ascontiguousarray(np.asarray(p), dtype=np.float32) for p in value]
        self._engine_stale = True

    @property
    def global_params(self):
        return self._global_params

    @global_params.setter
    def global_params(self, value):
        self._global_params = [np.ascontiguousarray(np.asarray(p), dtype=np.float32) for p in value]
        self._engine_stale = True

    # ------------------------------------------------------------------------------------------
    # posterior summaries — `_scdef.py:3394-3476`
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _ln_mean(p):
        return np.exp(p[0] + 0.5 * np.exp(p[1]) ** 2)

    @staticmethod
    def _ln_var(p):
        s2 = np.exp(p[1]) ** 2
        return (np.exp(s2) - 1.0) * np.exp(2.0 * p[0] + s2)

    def _summaries(self, fn, device_index):
        eng = self._engine
        if eng is not None and not self._engine_stale and getattr(self, "_device_summaries", None) is not None:
            loc, glo = self._device_summaries[device_index]
            loc = [t.cpu().numpy() for t in loc]
            glo = [t.cpu().numpy() for t in glo]
        else:
            loc = [fn(p) for p in self._local_params]
            glo = [fn(p) for p in self._global_params]
        L = self.n_layers
        out = {
            "cell_scale": loc[0],
            "gene_scale": glo[0],
            "factor_concentrations": glo[1],
            "factor_means": glo[2 + L],
            "alpha": glo[3 + L],
        }
        for idx in range(L):
            start = int(sum(self.layer_sizes[:idx]))
            end = start + self.layer_sizes[idx]
            out[f"{self.layer_names[idx]}z"] = loc[1][:, start:end]
            out[f"{self.layer_names[idx]}W"] = glo[2 + idx]
        return out

    def set_posterior_means(self):
        self.pmeans = self._summaries(self._ln_mean, 0)
        self.pmeans["brd"] = self.pmeans["factor_concentrations"]
        self.pmeans["wm"] = self.pmeans["factor_means"]
        self.pmeans["ard"] = self.pmeans["wm"]

    def set_posterior_variances(self):
        self.pvars = self._summaries(self._ln_var, 1)

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------------------------------
    def _get_engine(self):
        from .engine import Engine

        if self._engine is None:
            import torch

            dev = self.device or (f"cuda:{self._shard.local_rank}" if torch.cuda.is_available() else "cuda:0")
            self._engine = Engine(
                layer_sizes=self.layer_sizes, n_genes=self.n_genes, n_batches=self.n_batches,
                n_cells=self.n_cells, batch_id=self.batch_id, batch_lib_ratio=self.batch_lib_ratio,
                cell_alpha_factor=self.cell_alpha_factor, gene_ratio=self.gene_ratio, w_priors=self.w_priors,
                use_brd=self.use_brd, marginalize_alpha=self.marginalize_alpha, top_alpha=self.top_alpha,
                brd=self.brd, brd_mean=self.brd_mean, shrinkage_shape=self.shrinkage_shape,
                shrinkage_rate=self.shrinkage_rate, cell_scale_shape=self.cell_scale_shape,
                gene_scale_shape=self.gene_scale_shape, alpha_floor=self.alpha_floor,
                supervised_layer=self.supervised_layer, fixed_top_z=getattr(self, "_fixed_top_z", None),
                X=self.X if self.x_placement == "device" else None,
                X_u16=self.X if self.x_placement == "device_u16" else None,
                X_csr=(self.X.indptr.astype(np.int64), self.X.indices.astype(np.int32), self.X.data.astype(np.float32))
                if self.x_placement == "csr" else None, device=dev, lazy_adam=self.lazy_adam)
            self._engine_key = self._engine_signature()
            self._engine_stale = True
        elif self._engine_key != self._engine_signature():
            self._engine = None
            return self._get_engine()
        if self._engine_stale:
            self._engine.set_params(self._local_params, self._global_params)
            self._engine_stale = False
        return self._engine

    def _engine_signature(self):
        """Everything the engine copies into its device-side model description ONCE: a change of any of it after the
        first `_get_engine()` (`set_geneset_prior`, `set_connectivity_prior`, `update_model_priors`, `model.brd = ...`)
        must rebuild the engine, or training would silently keep the old priors."""
        import hashlib

        h = hashlib.blake2b(digest_size=16)

        def feed(v):
            if v is None:
                h.update(b"none")
            elif hasattr(v, "is_cuda"):   # torch tensor (possibly on the device): identity + shape is enough
                h.update(repr((int(v.data_ptr()), tuple(v.shape), str(v.dtype))).encode())
            else:
                a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
                h.update(repr(a.shape).encode())
                h.update(a.tobytes())

        for shape, rate in self.w_priors:
            feed(shape)
            feed(rate)
        for v in (self.cell_alpha_factor, self.gene_ratio, self.batch_lib_ratio, getattr(self, "_fixed_top_z", None),
                  [self.brd, self.brd_mean, self.shrinkage_shape, self.shrinkage_rate, self.cell_scale_shape,
                   self.gene_scale_shape, self.top_alpha, self.alpha_floor]):
            feed(v)
        return (tuple(self.layer_sizes), bool(self.marginalize_alpha), bool(self.use_brd), self.x_placement,
                self.supervised_layer, h.hexdigest())

    def _pull_params(self):
        lp, gp = self._engine.get_params()
        self._local_params, self._global_params = lp, gp
        self._engine_stale = False

    # ------------------------------------------------------------------------------------------
    # the seam: `local_loss_grad` / `global_loss_grad` (`_scdef.py:2808-2856`)
    # ------------------------------------------------------------------------------------------
    def _loss_grad(self, which, X, indices, local_params, global_params, key, annealing_parameter,
                   stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, num_samples):
        import torch

        from .engine import NoiseSpec

        eng = self._get_engine()
        eng.set_params(local_params, global_params)
        self._engine_stale = True  # engine now holds the caller's arrays, not self's
        idx = torch.as_tensor(np.asarray(indices), dtype=torch.int64, device=eng.device)
        Xb = None
        if X is not None:
            Xb = torch.as_tensor(np.asarray(X), dtype=torch.float32).to(eng.device).contiguous()
        noise = key if isinstance(key, NoiseSpec) else NoiseSpec.philox(self.seed or 0, int(key), self.noise_coupling)
        eng.elbo_grad(which, idx, noise, num_samples=int(num_samples), annealing=float(annealing_parameter),
                      stop_gradients=list(np.asarray(stop_gradients, dtype=float)),
                      stop_cell_budgets=float(stop_cell_budgets), stop_gene_budgets=float(stop_gene_budgets),
                      alpha=float(alpha), scale=self.n_cells / max(len(idx), 1), X_batch=Xb, lik_impl=self.lik_impl)
        loss = eng.loss_value()
        if which == "local":
            B = len(idx)
            dense = [np.zeros_like(np.asarray(p), dtype=np.float32) for p in local_params]
            for d, g in zip(dense, eng.local_grads(B)):
                d[:, np.asarray(indices)] = g.cpu().numpy()
            return loss, dense
        return loss, [g.cpu().numpy().copy() for g in eng.glob_grad.views]

    def local_loss_grad(self, X, indices, local_params, global_params, key, annealing_parameter,
                        stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, num_samples=100):
        """(loss, grads shaped like local_params) — dense (2,N,.), zero outside `indices`."""
        return self._loss_grad("local", X, indices, local_params, global_params, key, annealing_parameter,
                               stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, num_samples)

    def global_loss_grad(self, X, indices, local_params, global_params, key, annealing_parameter,
                         stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, num_samples=100):
        return self._loss_grad("global", X, indices, local_params, global_params, key, annealing_parameter,
                               stop_gradients, stop_cell_budgets, stop_gene_budgets, alpha, num_samples)

    # ------------------------------------------------------------------------------------------
    # fit / pretrain — `_scdef.py:2346-2737` (first-fit path), `2739-2780`
    # ------------------------------------------------------------------------------------------
    def fit(self, nmf_init: bool = False, max_cells_init: int = 5000, z_init_concentration: float = 0.05,
            n_rounds: int = 1, pretraining: bool = False, force_decay_factor: bool = False,
            root_epochs: int = 0, collapse_l1_fraction: Optional[float] = 0.8,
            collapse_l0_min_factors: int = 10, collapse_upper_min_factors: int = 5,
            learn_budgets_on_refit: bool = False, refit_top_layer=None, refit_top_factors=None,
            refit_rescale_relevance: bool = True, refit_relevance_max_ratio: float = 50.0,
            refit_rescale_w_by_layer_sizes: bool = True, freeze_w: bool = False, **kwargs: Any) -> None:
        marginalize_alpha_for_main_fit = bool(self.marginalize_alpha)
        self.root_epochs = int(root_epochs)
        if bool(getattr(self, "_has_fit", False)):
            raise NotImplementedError(
                "refit warm-start (fit -> filter_factors -> fit) is host-side orchestration outside the "
                "accelerated hot path (SURVEY.md §2 row 2); build a new model or call _learn() to continue.")
        self.init_var_params(init_budgets=True, init_alpha=True, nmf_init=nmf_init,
                             z_init_concentration=z_init_concentration)
        self.elbos = []
        self.step_sizes = []
        if pretraining:
            pre_n = int(kwargs.pop("pretraining_n_epoch", max(10, int(kwargs.get("n_epoch", 1000)) // 5)))
            prune = kwargs.pop("pretraining_prune_alpha", None)
            pre_kw = dict(kwargs)
            pre_kw.pop("n_epoch", None)
            pre_kw.pop("n_epochs", None)
            self.pretrain(n_epoch=pre_n, prune_alpha=None if prune is None else float(prune), **pre_kw)
        self.marginalize_alpha = marginalize_alpha_for_main_fit
        optimize_layers = list(range(self.n_layers - 1)) if root_epochs > 0 else list(range(self.n_layers))
        main_kw = dict(kwargs)
        if int(root_epochs) > 0:
            main_kw["filter"] = False
            main_kw["annotate"] = False
        self._learn(n_rounds=n_rounds, optimize_layers=optimize_layers, stop_gene_budgets=False,
                    stop_cell_budgets=False, freeze_w=freeze_w, **main_kw)
        self.qc_elbos = [np.asarray(x).copy() for x in self.elbos]
        trace = np.asarray(getattr(self, "entropy_annealing_trace", np.array([]))).copy()
        trace_ep = np.asarray(getattr(self, "entropy_annealing_trace_epochs", np.array([]))).copy()
        if root_epochs > 0:
            root_kw = dict(kwargs)
            root_kw["n_epoch"] = root_epochs
            root_kw["annealing"] = 1.0
            root_kw["entropy_anneal"] = False
            self._learn(n_rounds=1, optimize_layers=[self.n_layers - 1], freeze_w=freeze_w, **root_kw)
            self.entropy_annealing_trace, self.entropy_annealing_trace_epochs = trace, trace_ep
            if len(trace) > 0:
                self.adata.uns["entropy_annealing_trace"] = trace.copy()
                self.adata.uns["entropy_annealing_trace_epochs"] = trace_ep.copy()
            else:
                self.adata.uns.pop("entropy_annealing_trace", None)
                self.adata.uns.pop("entropy_annealing_trace_epochs", None)
        self._has_fit = True
        self._fit_revision += 1

    def pretrain(self, n_epoch: int = 200, prune_alpha: Optional[float] = None, **kwargs: Any) -> None:
        base_alpha = float(self.alpha)
        if prune_alpha is None:
            prune_alpha = max(10.0, base_alpha * 5.0)
        kw = dict(kwargs)
        kw.update(n_rounds=1, n_epoch=int(n_epoch), filter=False, annotate=False)
        self.logger.info("Starting two-pass pretraining (n_epoch=%s).", int(n_epoch))
        prev = bool(self.marginalize_alpha)
        self.marginalize_alpha = False
        try:
            self.alpha = 1.0
            self._learn(**kw)
            self.alpha = float(prune_alpha)
            self._learn(**kw)
        finally:
            self.alpha = base_alpha
            self.marginalize_alpha = prev

    # ------------------------------------------------------------------------------------------
    # the hot loop — `_learn`, `_scdef.py:2923-3392`
    # ------------------------------------------------------------------------------------------
    def _learn(self, n_rounds=1, n_epoch=1000, lr=1e-1, local_lr=1e-2, annealing=1.0, num_samples=100,
               batch_size=256, layerwise=False, min_epochs=50, tolerance=1e-5, patience=50,
               update_locals=True, update_globals=True, stop_cell_budgets=0, stop_gene_budgets=0,
               stop_gene_budgets_after_burnin=False, opt_layer=None, optimize_layers=None, filter=True,
               annotate=True, entropy_anneal=False, entropy_window=50, entropy_check_every=10,
               entropy_rel_change_low=1e-3, entropy_rel_change_high=1e-2, entropy_increase_factor=1.2,
               entropy_decrease_factor=0.9, entropy_min_annealing=1.0, entropy_max_annealing=5.0,
               entropy_optimizer_reset_threshold=0.25, freeze_w=False, freeze_z_layers=None, **kwargs):
        """Fit the model (same contract as the reference's `_learn`)."""
        if "n_epochs" in kwargs:
            n_epoch = kwargs.pop("n_epochs")
        if len(kwargs) > 0:
            unknown = ", ".join(sorted(kwargs.keys()))
            raise TypeError(f"Unexpected keyword arguments for _learn: {unknown}")
        freeze_z_layers = sorted({int(i) for i in (freeze_z_layers or [])})

        if int(n_rounds) > 1:
            total_rounds = int(n_rounds)
            args = dict(locals())
            for r in range(total_rounds):
                last = r == total_rounds - 1
                self._learn(
                    n_rounds=1, n_epoch=n_epoch, lr=lr if layerwise else lr * 0.5**r,
                    local_lr=local_lr if layerwise else local_lr * 0.5**r, annealing=annealing,
                    num_samples=num_samples, batch_size=batch_size, layerwise=layerwise, min_epochs=min_epochs,
                    tolerance=tolerance, patience=patience, update_locals=update_locals,
                    update_globals=update_globals, stop_cell_budgets=stop_cell_budgets,
                    stop_gene_budgets=stop_gene_budgets,
                    stop_gene_budgets_after_burnin=stop_gene_budgets_after_burnin, opt_layer=opt_layer,
                    optimize_layers=optimize_layers, filter=bool(filter) and last, annotate=bool(annotate) and last,
                    entropy_anneal=entropy_anneal, entropy_window=entropy_window,
                    entropy_check_every=entropy_check_every, entropy_rel_change_low=entropy_rel_change_low,
                    entropy_rel_change_high=entropy_rel_change_high,
                    entropy_increase_factor=entropy_increase_factor,
                    entropy_decrease_factor=entropy_decrease_factor, entropy_min_annealing=entropy_min_annealing,
                    entropy_max_annealing=entropy_max_annealing,
                    entropy_optimizer_reset_threshold=entropy_optimizer_reset_threshold, freeze_w=freeze_w,
                    freeze_z_layers=freeze_z_layers)
            return

        shard = self._shard
        n_total = int(self.n_cells_total)
        if batch_size is None:
            batch_size = n_total
        batch_size = int(batch_size)
        # under data parallelism `batch_size` is per rank: the global minibatch is world * batch_size
        global_batch = batch_size * shard.world
        num_complete, leftover = divmod(n_total, global_batch)
        num_batches = num_complete + bool(leftover)
        self.logger.info(f"Each epoch contains {num_batches} batches of size {int(min(global_batch, n_total))}")

        if optimize_layers is None:
            optimize_layers = [int(opt_layer)] if opt_layer is not None else list(range(self.n_layers))
        layers_to_optimize = sorted({int(i) for i in optimize_layers})
        if len(layers_to_optimize) == 0:
            raise ValueError("optimize_layers must contain at least one layer index.")
        if min(layers_to_optimize) < 0 or max(layers_to_optimize) >= self.n_layers:
            raise ValueError(f"optimize_layers must be within [0, {self.n_layers - 1}].")
        stop_gradients = np.ones(self.n_layers, dtype=np.float32)
        stop_gradients[layers_to_optimize] = 0.0
        stop_cell_budgets = float(stop_cell_budgets)
        stop_gene_budgets = float(stop_gene_budgets)

        if int(entropy_window) <= 0:
            raise ValueError("entropy_window must be > 0.")
        if int(entropy_check_every) <= 0:
            raise ValueError("entropy_check_every must be > 0.")
        if float(entropy_rel_change_low) < 0.0 or float(entropy_rel_change_high) < 0.0:
            raise ValueError("entropy relative-change thresholds must be >= 0.")
        if float(entropy_rel_change_low) > float(entropy_rel_change_high):
            raise ValueError("entropy_rel_change_low must be <= entropy_rel_change_high.")
        if float(entropy_increase_factor) <= 1.0:
            raise ValueError("entropy_increase_factor must be > 1.0.")
        if not (0.0 < float(entropy_decrease_factor) < 1.0):
            raise ValueError("entropy_decrease_factor must be in (0, 1).")
        if float(entropy_min_annealing) <= 0.0:
            raise ValueError("entropy_min_annealing must be > 0.")
        if float(entropy_max_annealing) < float(entropy_min_annealing):
            raise ValueError("entropy_max_annealing must be >= entropy_min_annealing.")
        if float(entropy_optimizer_reset_threshold) < 0.0:
            raise ValueError("entropy_optimizer_reset_threshold must be >= 0.")

        import torch

        from .engine import NoiseSpec
        from .stream import MinibatchStream

        eng = self._get_engine()
        eng.reset_optimizer()  # fresh optax state per _learn call (3079-3082)
        start, _, counts = _dist.shard_bounds(self.n_cells, shard)
        # data parallel: position-based ownership, the cells' rows move once per epoch so that every rank has
        # exactly batch_size rows per step (dist.epoch_layout); in host placement the counts stay with their home
        # rank and travel per step (stream.py); static ownership when the shard sizes cannot be kept
        migrator = None
        if (shard.world > 1 and self.epoch_reshard and num_batches > 1 and self.x_placement != "csr"  # CSR rows: static
                and _dist.epoch_layout(np.arange(n_total), global_batch, counts) is not None):
            migrator = _dist.RowMigrator(shard, counts)
        # Host placement: a GPU step shorter than the host needs to gather and launch it is fed by two staging threads
        # (measured at C5 on a B200: S=10, 0.45 ms per step: 0.90 -> 0.62 ms end to end; S=100, 1.35 ms per step: a second
        # thread only takes the GIL from the launching thread, 1.37 -> 1.43).  Likelihood work per step ~ S B G K0.
        light_step = float(num_samples) * batch_size * self.n_genes * self.layer_sizes[0] < 4e9
        stream = MinibatchStream(n_total, global_batch, shard, eng.device,
                                 host_X=self.X if self.x_placement == "host" else None,
                                 n_local=self.n_cells, start=start, migrator=migrator,
                                 tensors_fn=eng.per_cell_tensors, stage=(2, 5, 3) if light_step else (1, 3, 2))
        self._last_stream = stream
        annealing_parameter = float(annealing)
        alpha_f = float(self.alpha)
        scale_den = None  # set per step: global minibatch size
        gw = 1.0 / shard.world
        step_counter = getattr(self, "_noise_step", 0)
        all_losses, total_epochs = [], 0
        min_loss, early_stop_counter = np.inf, 0
        stop_message, interrupted = None, False
        trace, trace_epochs = [], []
        gene_budgets_frozen = bool(stop_gene_budgets >= 1.0)
        loss_dev = torch.zeros(num_batches + 1, dtype=torch.float64, device=eng.device)   # last element: stop flag
        # data parallel: a Ctrl-C on one rank must not leave that rank alone in (or out of) the collectives.  SIGINT
        # only sets a flag; the flag rides with the per-epoch loss all-reduce, and every rank leaves at the same epoch.
        import signal
        import threading

        sigint = {"hit": False, "old": None}
        if shard.world > 1 and threading.current_thread() is threading.main_thread():
            def _on_sigint(signum, frame):
                sigint["hit"] = True
            sigint["old"] = signal.signal(signal.SIGINT, _on_sigint)
        steps_done = 0
        limit = self.max_steps_per_learn
        reducer = _dist.GradReducer(eng, shard) if shard.world > 1 else None

        epochs = range(n_epoch)
        pbar = _tqdm(epochs, disable=shard.rank != 0 or self.logger.level >= 30) if _tqdm is not None else None
        try:
            for epoch in (pbar if pbar is not None else epochs):
                if stop_gene_budgets_after_burnin and not gene_budgets_frozen and total_epochs >= min_epochs:
                    stop_gene_budgets = 1.0
                    gene_budgets_frozen = True
                    self.logger.info(f"Freezing gene budgets at epoch {total_epochs} (post burn-in).")
                stream.new_epoch()
                n_it = 0
                for it in range(num_batches):
                    if self._step_hook is not None:
                        self._step_hook(steps_done, eng, stream)
                    step_counter += 1
                    noise = NoiseSpec.philox(self.seed or 0, step_counter, self.noise_coupling, *stream.noise_window(it))
                    idx, Xb, n_glob = stream.batch(it)
                    if Xb is None and (eng.X_csr is not None or eng.X_u16 is not None) and int(idx.numel()) > 0:
                        Xb = eng.gather_rows(idx)   # CSR placement: densify the minibatch ONCE for both passes
                    kw = dict(num_samples=int(num_samples), annealing=annealing_parameter,
                              stop_gradients=stop_gradients, stop_cell_budgets=stop_cell_budgets,
                              stop_gene_budgets=stop_gene_budgets, alpha=alpha_f, scale=n_total / n_glob,
                              global_weight=gw, X_batch=Xb, lik_impl=self.lik_impl)
                    if update_locals:
                        eng.local_catchup(idx)  # lazy Adam: the minibatch rows become current before they are read
                        # the local-pass loss is overwritten by the global pass (3235-3271): skip it when unused
                        eng.elbo_grad("local", idx, noise, want_loss=not update_globals, **kw)
                        eng.adam_local(idx, float(local_lr), freeze_z_layers)
                    if update_globals:
                        # same rng_input, globals untouched since the local pass: its W_0 sample tiles are reusable
                        eng.elbo_grad("global", idx, noise, reuse_w0=bool(update_locals),
                                      lik_first=reducer is not None and reducer.lik_first, **kw)
                        if reducer is not None:
                            reducer.reduce()   # W_0's share overlaps the tail of the pass (dist.GradReducer)
                        eng.adam_global(float(lr), freeze_w=bool(freeze_w))
                    stream.release()
                    stream.prefetch(it + 1)  # host placement: stage the next minibatch while this step runs
                    loss_dev[it] = eng.loss_buf.sum()
                    n_it += 1
                    steps_done += 1
                    if limit is not None and steps_done >= int(limit):
                        break
                if shard.world > 1:
                    loss_dev[num_batches] = 1.0 if sigint["hit"] else 0.0
                    _dist.all_reduce_sum(loss_dev, shard)
                current_loss = float(loss_dev[:n_it].mean().item())  # one host sync per epoch (3273)
                if shard.world > 1 and float(loss_dev[num_batches].item()) > 0.0:
                    raise KeyboardInterrupt   # agreed on by all ranks: everybody takes the "exit safely" path
                all_losses.append(current_loss)
                total_epochs += 1
                trace.append(float(annealing_parameter))
                trace_epochs.append(int(total_epochs))
                if limit is not None and steps_done >= int(limit):
                    stop_message = f"Stopping learning: reached max_steps_per_learn={int(limit)}."
                    break

                recent = None
                if bool(entropy_anneal) and int(total_epochs) % int(entropy_check_every) == 0:
                    if len(all_losses) >= int(entropy_window):
                        recent = np.asarray(all_losses[-int(entropy_window):], dtype=float)
                if recent is not None:
                    denom = max(abs(float(np.mean(recent))), 1e-12)
                    relative_change = float((np.max(recent) - np.min(recent)) / denom)
                    old, new = annealing_parameter, annealing_parameter
                    if relative_change < float(entropy_rel_change_low):
                        new = min(old * float(entropy_increase_factor), float(entropy_max_annealing))
                    elif relative_change > float(entropy_rel_change_high):
                        new = max(old * float(entropy_decrease_factor), float(entropy_min_annealing))
                    if abs(new - old) > 0.0:
                        annealing_parameter = float(np.float32(new))
                        if abs(new - old) / max(abs(old), 1e-12) >= float(entropy_optimizer_reset_threshold):
                            eng.reset_optimizer()
                        self.logger.info("Entropy annealing update at epoch %s: %.4f -> %.4f (relative_change=%.6g).",
                                         int(total_epochs), float(old), float(new), relative_change)

                relative_improvement = np.nan
                if total_epochs >= min_epochs:
                    if min_loss == np.inf:
                        min_loss = current_loss
                    relative_improvement = (min_loss - current_loss) / np.abs(min_loss)
                    min_loss = min(min_loss, current_loss)
                    early_stop_counter = early_stop_counter + 1 if relative_improvement < tolerance else 0
                if pbar is not None:
                    postfix = {"Loss": current_loss}
                    if bool(entropy_anneal):
                        postfix["anneal"] = float(annealing_parameter)
                    if total_epochs >= min_epochs:
                        postfix["Rel. impr."] = relative_improvement
                    pbar.set_postfix(postfix)
                if total_epochs >= min_epochs and early_stop_counter >= patience:
                    stop_message = f"Converged at epoch {total_epochs}."
                    break
        except KeyboardInterrupt:
            interrupted = True
            self.logger.info("Interrupted. Exiting safely...")
        finally:
            if sigint["old"] is not None:
                signal.signal(signal.SIGINT, sigint["old"])
            if pbar is not None:
                pbar.close()
            import sys as _sys

            failing = _sys.exc_info()[0] is not None   # a rank-local error is propagating (KeyboardInterrupt is caught above)
            if self._step_hook is not None and not failing:
                self._step_hook(steps_done, eng, stream)
            # position-based ownership: the rows go back to block order before anything reads them.  Not on a
            # rank-local error: the way home is a collective the other ranks will never enter — fail fast instead
            # (torchrun tears the job down) and leave the rows where they are.
            stream.finish(collective=not (failing and shard.world > 1))

        if stop_message is None and not interrupted:
            stop_message = f"Stopping learning: reached max epochs (n_epoch={int(n_epoch)})."
        if stop_message is not None:
            self.logger.info(stop_message)
        self._noise_step = step_counter

        self._device_summaries = eng.posterior()
        self._assignment_counts = self._device_assignment_counts(eng)
        self._pull_params()
        self.elbos.append(all_losses)
        self.step_sizes.append(lr)
        for k in ("alpha_trace", "alpha_trace_epochs", "n_eff_parents_trace", "n_eff_parents_trace_epochs",
                  "active_l0_factor_counts_trace", "alpha_schedule_alphas", "alpha_schedule_losses",
                  "alpha_schedule_epochs"):
            self.adata.uns.pop(k, None)
        self.entropy_annealing_trace = np.asarray(trace, dtype=float)
        self.entropy_annealing_trace_epochs = np.asarray(trace_epochs, dtype=int)
        self.adata.uns["entropy_annealing_trace"] = self.entropy_annealing_trace.copy()
        self.adata.uns["entropy_annealing_trace_epochs"] = self.entropy_annealing_trace_epochs.copy()
        self.set_posterior_means()
        self.set_posterior_variances()
        self._device_summaries = None
        if self.use_brd and filter:
            self.filter_factors(upper_only=True, annotate=annotate)   # `_scdef.py:3386-3387`
        self._assignment_counts = None
        # annotation (`annotate_adata`, `make_layercolors`, `3388-3392`) is unchanged library code that
        # consumes `pmeans` / `factor_lists`; it is outside this package's scope and is skipped here.

    # ------------------------------------------------------------------------------------------
    # factor filtering — `filter_factors`, `_scdef.py:3478-3582`
    # ------------------------------------------------------------------------------------------
    def _device_assignment_counts(self, eng=None):
        """Per layer, how many cells (over all ranks) have their largest posterior-mean z at each factor:
        computed on the device from the local blocks (`scdef_assignment_counts`), summed over the ranks."""
        import torch

        eng = eng if eng is not None else self._get_engine()
        counts = eng.assignment_counts()
        if self._shard.world > 1:
            flat = torch.cat(counts)
            _dist.all_reduce_sum(flat, self._shard)
            offs = np.concatenate([[0], np.cumsum(self.layer_sizes)])
            counts = [flat[offs[l]: offs[l + 1]] for l in range(self.n_layers)]
        return [c.cpu().numpy().astype(np.int64) for c in counts]

    def filter_factors(self, brd_min=1.0, ard_min=0.001, clarity_min=0.5, n_eff_parents_max=1.5,
                       local_l0_scores=False, batch_purity_max=None, batch_purity_soft_max=None,
                       min_cells_upper=0.001, min_cells_lower=0.0, filter_up=True, annotate=True,
                       upper_only=False):
        """The reference's `filter_factors` for `upper_only=True` — the call `_learn` makes (`_scdef.py:3386-3387`).
        Layer 0 is left whole; upper layers keep the factors with enough hard-assigned cells (counts from the
        device, all ranks) and, with `filter_up`, only those the kept factors below attach to (`3540-3568`).
        Filtering layer 0 (`upper_only=False`) needs `tools.factor_diagnostics` (BRD/ARD/clarity table), which is
        host-side library code outside this package."""
        if not upper_only:
            raise NotImplementedError("filter_factors(upper_only=False) needs scdef.tools.factor_diagnostics "
                                      "(out of scope here); `_learn` itself only uses upper_only=True")
        n_all = int(self.n_cells_total)
        if min_cells_upper != 0 and min_cells_upper < 1.0:
            min_cells_upper = max(min_cells_upper * n_all, 10)
        counts_all = getattr(self, "_assignment_counts", None)
        if counts_all is None:
            counts_all = self._device_assignment_counts()
        new_factor_lists = []
        for i, layer_name in enumerate(self.layer_names):
            if i == 0:
                keep = np.arange(self.layer_sizes[i])
            else:
                counts = np.asarray(counts_all[i])
                keep = np.arange(self.layer_sizes[i])[np.where(counts >= min_cells_upper)[0]]
                if filter_up and len(keep) > 0 and len(new_factor_lists[i - 1]) > 0:
                    mat = np.asarray(self.pmeans[f"{layer_name}W"])[keep]
                    att = [keep[np.argmax(mat[:, f])] for f in new_factor_lists[i - 1]]
                    keep = np.unique(list(set(np.unique(att)).intersection(keep)))
            if len(keep) == 0:
                self.logger.info(f"No factors in layer {i} satisfy the filtering criterion. Please adjust the filtering "
                                 f"parameters.Keeping all factors for layer {i} for now.")
                keep = np.arange(self.layer_sizes[i])
            new_factor_lists.append(np.asarray(keep, dtype=int))
        self.factor_lists = new_factor_lists
        self.set_factor_names()

    learn = _learn  # legacy name used by the docs notebook / benchmarks
