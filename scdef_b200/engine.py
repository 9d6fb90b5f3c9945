"""Device engine: owns the PyTorch buffers of one rank and drives the C ABI.

PyTorch is plumbing here (device memory, streams, NCCL); all arithmetic of the hot path runs
in the hand-written kernels behind `include/scdef_b200.h`.

Layouts (reference `_scdef.py:1463-1732`): every variational block is a stacked
pair [mu, rho] along axis 0, float32.  All local blocks live in one flat buffer and all global
blocks in another (so the global gradient is ONE contiguous buffer for the data-parallel
all-reduce and the Adam moments are two more); `local_params` / `global_params` are views with
the reference's shapes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Blocks, Call, ModelDesc, Noise, MAX_LAYERS


class ScdefError(RuntimeError):
    pass


@dataclass
class NoiseSpec:
    """Noise of one objective evaluation (one `rng_input`, `_scdef.py:3231`)."""

    mode: int = _lib.NOISE_PHILOX
    seed: int = 0
    step: int = 0
    coupling: bool = True  # reference behaviour: same-shape blocks share draws
    eps: Optional[Dict[str, object]] = None  # explicit: c, z, s, tau, omega, a, W (list)
    row_offset: int = 0  # data parallelism: this rank's rows are [row_offset, row_offset+B) of rows_total
    rows_total: int = 0  # (0: no window) — per-cell philox draws then equal the single-rank run's

    @staticmethod
    def philox(seed: int, step: int, coupling: bool = True, row_offset: int = 0, rows_total: int = 0) -> "NoiseSpec":
        return NoiseSpec(_lib.NOISE_PHILOX, int(seed), int(step), bool(coupling), None, int(row_offset), int(rows_total))

    @staticmethod
    def explicit(c, z, s, tau, omega, a, W) -> "NoiseSpec":
        return NoiseSpec(_lib.NOISE_EXPLICIT, 0, 0, False, dict(c=c, z=z, s=s, tau=tau, omega=omega, a=a, W=list(W)))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _block_shapes(n_rows, layer_sizes, n_batches, n_genes):
    Ks = list(layer_sizes)
    outs = [n_genes] + Ks[:-1]
    local = [("cell_scale", (2, n_rows, 1)), ("z", (2, n_rows, sum(Ks)))]
    glob = [("gene_scale", (2, n_batches, n_genes)), ("brd", (2, Ks[0], 1))]
    glob += [(f"W{l}", (2, Ks[l], outs[l])) for l in range(len(Ks))]
    glob += [("ard", (2, Ks[0], 1)), ("alpha", (2, 1, 1))]
    return local, glob


class _FlatGroup:
    """A flat float32 buffer carved into named views; every view starts 16-byte aligned."""

    def __init__(self, shapes, device, last=None):
        """`last`: name of the block placed at the END of the flat buffer (the views keep the order of `shapes`): the
        global groups put W0 there, so that "W_0" and "everything else" are two contiguous ranges for the data-parallel
        all-reduce (dist.GradReducer)."""
        self.names = [n for n, _ in shapes]
        order = [i for i, (n, _) in enumerate(shapes) if n != last] + [i for i, (n, _) in enumerate(shapes) if n == last]
        offs, o = [0] * len(shapes), 0
        for i in order:
            offs[i] = o
            o += (int(np.prod(shapes[i][1])) + 3) // 4 * 4
        self.flat = torch.zeros(max(o, 4), dtype=torch.float32, device=device)
        self.views = [self.flat[off : off + int(np.prod(shp))].view(*shp) for off, (_, shp) in zip(offs, shapes)]
        self.split = offs[self.names.index(last)] if last in self.names else None   # flat[:split] | flat[split:] = the `last` block

    def blocks(self, n_layers, local: bool) -> Blocks:
        b = Blocks()
        if local:
            b.cell_scale, b.z = _ptr(self.views[0]), _ptr(self.views[1])
        else:
            b.gene_scale, b.brd = _ptr(self.views[0]), _ptr(self.views[1])
            for l in range(n_layers):
                b.W[l] = self.views[2 + l].data_ptr()
            b.ard, b.alpha = _ptr(self.views[2 + n_layers]), _ptr(self.views[3 + n_layers])
        return b


def _merge(local: Blocks, glob: Blocks, n_layers) -> Blocks:
    b = Blocks()
    b.cell_scale, b.z = local.cell_scale, local.z
    b.gene_scale, b.brd, b.ard, b.alpha = glob.gene_scale, glob.brd, glob.ard, glob.alpha
    for l in range(n_layers):
        b.W[l] = glob.W[l]
    return b


class Engine:
    """One rank's device state + the calls of the hot path."""

    def __init__(self, *, layer_sizes: Sequence[int], n_genes: int, n_batches: int, n_cells: int,
                 batch_id, batch_lib_ratio, cell_alpha_factor, gene_ratio, w_priors,
                 use_brd=True, marginalize_alpha=False, top_alpha=1.0, brd=1.0, brd_mean=1.0,
                 shrinkage_shape=1.0, shrinkage_rate=1.0, cell_scale_shape=1.0, gene_scale_shape=1.0,
                 alpha_floor=0.1, supervised_layer=None, fixed_top_z=None, X=None, X_csr=None, X_u16=None,
                 device="cuda:0", batch_max=256, lazy_adam=True):
        if not torch.cuda.is_available():
            raise ScdefError("scdef_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
        self.lib = _lib.load()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.Ks = [int(k) for k in layer_sizes]
        self.L = len(self.Ks)
        if self.L > MAX_LAYERS:
            raise ScdefError(f"at most {MAX_LAYERS} layers are supported, got {self.L}")
        self.Kt = sum(self.Ks)
        self.G, self.nb, self.N = int(n_genes), int(n_batches), int(n_cells)
        dev = self.device
        f32 = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.array(a, copy=True))).to(dev, torch.float32).contiguous()
        self.X = None if X is None else (X if isinstance(X, torch.Tensor) else torch.as_tensor(np.asarray(X)))
        if self.X is not None:
            self.X = self.X.to(dev, torch.float32).contiguous()
        self.batch_id = torch.as_tensor(np.asarray(batch_id) if not isinstance(batch_id, torch.Tensor) else batch_id).to(dev, torch.int32).contiguous()
        self.batch_lib_ratio = f32(batch_lib_ratio).reshape(-1)
        self.cell_alpha_factor = f32(cell_alpha_factor).reshape(-1)
        gr = np.broadcast_to(np.asarray(gene_ratio, dtype=np.float64), (self.nb, self.G))
        self.gene_ratio = f32(np.ascontiguousarray(gr))
        self.row_slot = torch.full((max(self.N, 1),), -1, dtype=torch.int32, device=dev)
        self.x_lgamma = None
        # counts kept as CSR on the device: (indptr int64, indices int32, data float32); a minibatch is densified
        # per call (`scdef_csr_gather_rows`) and handed to the kernels as X_batch
        self.X_csr = None
        if X_csr is not None:
            if self.X is not None:
                raise ScdefError("give either X or X_csr")
            ip, ix, dv = X_csr
            as_t = lambda a, dt: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev, dt).contiguous()
            self.X_csr = (as_t(ip, torch.int64), as_t(ix, torch.int32), as_t(dv, torch.float32))
            if int(self.X_csr[0].numel()) != self.N + 1:
                raise ScdefError("X_csr row pointers must have n_cells + 1 entries")
            self.x_lgamma = torch.empty(max(self.N, 1), dtype=torch.float32, device=dev)
            rc = self.lib.scdef_csr_row_stats(_ptr(self.X_csr[0]), _ptr(self.X_csr[2]), self.N, None, _ptr(self.x_lgamma), self._stream())
            if rc != 0:
                raise ScdefError(f"scdef_csr_row_stats failed ({rc})")
            self._xb_scratch = None
        # counts kept as uint16 on the device (half of float32's HBM); a minibatch is converted per call
        # (`scdef_u16_gather_rows`) into the dense float32 X_batch, like the CSR rows
        self.X_u16 = None
        if X_u16 is not None:
            if self.X is not None or self.X_csr is not None:
                raise ScdefError("give one of X, X_csr, X_u16")
            t = X_u16 if isinstance(X_u16, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X_u16))
            if t.dtype not in (torch.uint16, torch.int16):
                t = t.to(torch.uint16)
            self.X_u16 = t.to(dev).contiguous()
            self.x_lgamma = torch.empty(max(self.N, 1), dtype=torch.float32, device=dev)
            rc = self.lib.scdef_u16_row_stats(_ptr(self.X_u16), self.N, self.G, None, _ptr(self.x_lgamma), self._stream())
            if rc != 0:
                raise ScdefError(f"scdef_u16_row_stats failed ({rc})")
            self._xb_scratch = None
        if self.X is not None and self.N > 0:
            self.x_lgamma = torch.empty(self.N, dtype=torch.float32, device=dev)
            rc = self.lib.scdef_row_stats(_ptr(self.X), self.N, self.G, None, _ptr(self.x_lgamma), self._stream())
            if rc != 0:
                raise ScdefError(f"scdef_row_stats failed ({rc})")
        self._prior_tensors = []
        d = ModelDesc()
        d.abi_version = _lib.ABI_VERSION
        d.n_layers = self.L
        for l, k in enumerate(self.Ks):
            d.layer_sizes[l] = k
        d.n_cells, d.n_genes, d.n_batches = self.N, self.G, self.nb
        d.use_brd, d.marginalize_alpha = int(bool(use_brd)), int(bool(marginalize_alpha))
        d.supervised_layer = -1 if supervised_layer is None else int(supervised_layer)
        d.alpha_floor, d.top_alpha, d.brd, d.brd_mean = alpha_floor, top_alpha, brd, brd_mean
        d.shrinkage_shape, d.shrinkage_rate = shrinkage_shape, shrinkage_rate
        d.cell_scale_shape, d.gene_scale_shape = cell_scale_shape, gene_scale_shape
        d.X = _ptr(self.X)
        d.batch_id = _ptr(self.batch_id)
        d.batch_lib_ratio = _ptr(self.batch_lib_ratio)
        d.cell_alpha_factor = _ptr(self.cell_alpha_factor)
        d.gene_ratio = _ptr(self.gene_ratio)
        d.x_lgamma_rowsum = _ptr(self.x_lgamma)
        d.row_slot = _ptr(self.row_slot)
        outs = [self.G] + self.Ks[:-1]
        for l in range(self.L):
            for which, arr_name, sc_name in ((0, "w_prior_shape", "w_prior_shape_scalar"), (1, "w_prior_rate", "w_prior_rate_scalar")):
                val = w_priors[l][which]
                a = np.asarray(val.detach().cpu() if isinstance(val, torch.Tensor) else val, dtype=np.float64)
                if a.ndim == 0:
                    getattr(d, sc_name)[l] = float(a)
                    getattr(d, arr_name)[l] = None
                else:
                    t = f32(np.ascontiguousarray(np.broadcast_to(a, (self.Ks[l], outs[l]))))
                    self._prior_tensors.append(t)
                    getattr(d, arr_name)[l] = t.data_ptr()
        self.fixed_top_z = None
        if fixed_top_z is not None:
            self.fixed_top_z = f32(fixed_top_z)
            d.fixed_top_z = _ptr(self.fixed_top_z)
        self.desc = d
        h = C.c_void_p()
        rc = self.lib.scdef_ctx_create(C.byref(d), C.byref(h))
        if rc != 0:
            raise ScdefError(f"scdef_ctx_create failed ({rc})")
        self.ctx = h

        lshapes, gshapes = _block_shapes(self.N, self.Ks, self.nb, self.G)
        self.local = _FlatGroup(lshapes, dev)
        self.glob = _FlatGroup(gshapes, dev)
        self.local_m, self.local_v = _FlatGroup(lshapes, dev), _FlatGroup(lshapes, dev)
        self.glob_m, self.glob_v = _FlatGroup(gshapes, dev), _FlatGroup(gshapes, dev)
        self.glob_grad = _FlatGroup(gshapes, dev, last="W0")   # [small blocks | W_0]: two contiguous all-reduce ranges
        self._w0_event = None
        self.batch_max = 0
        self.local_grad = None
        self._ensure_batch(int(batch_max))
        self.loss_buf = torch.zeros(2, dtype=torch.float64, device=dev)
        self.workspace = torch.empty(0, dtype=torch.uint8, device=dev)
        self._ws_key = None
        self.local_step = 0
        self.global_step = 0
        # exact lazy local Adam (bit-identical to the dense kernel): per-row "last applied step"
        self.lazy_adam = bool(lazy_adam)
        self.row_step = torch.zeros(max(self.N, 1), dtype=torch.int32, device=dev)
        self._lazy_pending = False   # some rows are behind local_step
        self._lazy_hyper = None      # (lr, b1, b2, eps, freeze tuple) of the pending steps
        self._bc_key, self._bc_tab = None, None
        self.launches = 0  # kernels enqueued through the ABI (for bench.py's gpu_launches)

    # ------------------------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.scdef_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:  # noqa: BLE001
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc, what):
        if rc != 0:
            msg = self.lib.scdef_last_error(self.ctx)
            raise ScdefError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _ensure_batch(self, B):
        if B <= self.batch_max and self.local_grad is not None:
            return
        self.batch_max = max(B, 1)
        lshapes, _ = _block_shapes(self.batch_max, self.Ks, self.nb, self.G)
        self.local_grad = _FlatGroup(lshapes, self.device)

    def _ensure_workspace(self, B, S, lik_impl):
        key = (max(B, 1), S, lik_impl)
        if self._ws_key is not None and self._ws_key[1:] == key[1:] and self._ws_key[0] >= key[0]:
            return
        need = int(self.lib.scdef_workspace_bytes(self.ctx, key[0], S, lik_impl))
        if need == 0:
            raise ScdefError("scdef_workspace_bytes returned 0")
        if self.workspace.numel() < need:
            self.workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        self._ws_key = key

    # ---- parameter I/O -------------------------------------------------------------------
    @property
    def local_params(self) -> List[torch.Tensor]:
        return self.local.views

    @property
    def global_params(self) -> List[torch.Tensor]:
        return self.glob.views

    def set_params(self, local_params=None, global_params=None):
        if local_params is not None:
            self.local_sync()
            for v, p in zip(self.local.views, local_params):
                v.copy_(torch.as_tensor(np.asarray(p) if not isinstance(p, torch.Tensor) else p).to(self.device, torch.float32).reshape(v.shape))
        if global_params is not None:
            for v, p in zip(self.glob.views, global_params):
                v.copy_(torch.as_tensor(np.asarray(p) if not isinstance(p, torch.Tensor) else p).to(self.device, torch.float32).reshape(v.shape))

    def get_params(self):
        self.local_sync()
        return ([v.detach().cpu().numpy().copy() for v in self.local.views],
                [v.detach().cpu().numpy().copy() for v in self.glob.views])

    def per_cell_tensors(self) -> List[torch.Tensor]:
        """Every per-cell array of this rank as an in-place 2-D (rows, width) view: the counts, the per-cell
        constants, the local blocks and their Adam moments (mu and rho planes), the lazy-Adam row clock.
        This is what `dist.RowMigrator` moves when the cells change ranks at an epoch boundary."""
        out = []
        x16 = None if self.X_u16 is None else self.X_u16.view(torch.int16)   # NCCL moves it as int16 bits
        for t in (self.X, x16, self.batch_id, self.batch_lib_ratio, self.cell_alpha_factor, self.x_lgamma,
                  self.row_step, self.fixed_top_z):
            if t is not None and t.numel() > 0:
                out.append(t.view(self.N, -1))
        for grp in (self.local, self.local_m, self.local_v):
            for v in grp.views:          # (2, N, width): the two planes are separate row-major matrices
                out.extend([v[0], v[1]])
        return out

    def reset_optimizer(self):
        """Fresh optax state (`_scdef.py:3081-3082, 3317-3319`)."""
        self.local_sync()  # pending zero-gradient steps belong to the old state
        self.row_step.zero_()
        for g in (self.local_m, self.local_v, self.glob_m, self.glob_v):
            g.flat.zero_()
        self.local_step = 0
        self.global_step = 0

    # ---- the objective ---------------------------------------------------------------------
    def _noise_struct(self, noise: NoiseSpec, keep):
        n = Noise()
        n.mode, n.coupling, n.seed, n.step = noise.mode, int(noise.coupling), noise.seed, noise.step
        n.local_row_offset, n.local_rows_total = int(noise.row_offset), int(noise.rows_total)
        if noise.mode == _lib.NOISE_EXPLICIT:
            dev = self.device
            conv = lambda t: torch.as_tensor(np.asarray(t) if not isinstance(t, torch.Tensor) else t).to(dev, torch.float32).contiguous()
            e = noise.eps
            ts = dict(c=conv(e["c"]), z=conv(e["z"]), s=conv(e["s"]), tau=conv(e["tau"]), omega=conv(e["omega"]), a=conv(e["a"]))
            Ws = [conv(w) for w in e["W"]]
            keep.extend(list(ts.values()) + Ws)
            n.eps_cell_scale, n.eps_z, n.eps_gene_scale = _ptr(ts["c"]), _ptr(ts["z"]), _ptr(ts["s"])
            n.eps_brd, n.eps_ard, n.eps_alpha = _ptr(ts["tau"]), _ptr(ts["omega"]), _ptr(ts["a"])
            for l, w in enumerate(Ws):
                n.eps_W[l] = w.data_ptr()
        return n

    def w0_grad_event(self) -> torch.cuda.Event:
        """Event recorded by every GLOBAL pass when the W_0 gradient is final (`scdef_set_w0_grad_event`): lets the
        data-parallel loop start reducing it while the rest of the pass still runs."""
        if self._w0_event is None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))   # creates the underlying cudaEvent_t
            rc = self.lib.scdef_set_w0_grad_event(self.ctx, C.c_void_p(ev.cuda_event))
            if rc != 0:
                raise ScdefError(f"scdef_set_w0_grad_event failed ({rc})")
            self._w0_event = ev
        return self._w0_event

    def elbo_grad(self, which: str, indices: torch.Tensor, noise: NoiseSpec, *, num_samples: int,
                  annealing=1.0, stop_gradients=None, stop_cell_budgets=0.0, stop_gene_budgets=0.0,
                  alpha=1.0, scale=None, global_weight=1.0, X_batch: Optional[torch.Tensor] = None,
                  lik_impl=_lib.LIK_AUTO, zero_loss=True, reuse_w0=False, want_loss=True, lik_first=False):
        """Enqueue loss + gradient (`local_loss_grad` / `global_loss_grad`, `_scdef.py:2848-2849`).
        `indices`: int64 device tensor of this rank's rows.  Results: `self.loss_buf` (device
        double[2]: cell part, global part) and `self.local_grad` / `self.glob_grad`."""
        B = int(indices.numel())
        self._ensure_batch(B)
        self._ensure_workspace(B, int(num_samples), lik_impl)
        if X_batch is None and (self.X_csr is not None or self.X_u16 is not None) and B > 0:
            X_batch = self.gather_rows(indices)
        call = Call()
        call.pass_ = _lib.PASS_LOCAL if which == "local" else _lib.PASS_GLOBAL
        call.num_samples, call.batch, call.lik_impl = int(num_samples), B, int(lik_impl)
        call.annealing = float(annealing)
        sg = [0.0] * self.L if stop_gradients is None else [float(x) for x in stop_gradients]
        for l in range(self.L):
            call.stop_gradients[l] = sg[l]
        call.stop_cell_budgets, call.stop_gene_budgets = float(stop_cell_budgets), float(stop_gene_budgets)
        call.alpha = float(alpha)
        call.scale = float(scale) if scale is not None else (self.N / max(B, 1))
        call.global_weight = float(global_weight)
        call.flags = ((_lib.FLAG_REUSE_W0_SAMPLES if reuse_w0 else 0) | (0 if want_loss else _lib.FLAG_SKIP_LOSS)
                      | (_lib.FLAG_LIK_FIRST if lik_first else 0))
        keep = []
        if B > 0:
            assert indices.dtype == torch.int64 and indices.is_cuda and indices.is_contiguous()
            call.indices = indices.data_ptr()
        if X_batch is not None:
            assert X_batch.dtype == torch.float32 and X_batch.is_cuda and X_batch.is_contiguous()
            assert tuple(X_batch.shape) == (B, self.G)
            call.X_batch = X_batch.data_ptr()
        nz = self._noise_struct(noise, keep)
        if zero_loss:
            self.loss_buf.zero_()
        params, grads = self._cached_blocks()
        rc = self.lib.scdef_elbo_grad(self.ctx, C.byref(call), C.byref(params), C.byref(nz),
                                      _ptr(self.loss_buf), C.byref(grads), _ptr(self.workspace),
                                      self.workspace.numel(), self._stream())
        self._check(rc, "scdef_elbo_grad")
        self._keep = keep  # explicit-noise tensors must outlive the enqueued kernels
        return self.loss_buf

    def gather_rows(self, indices: torch.Tensor) -> torch.Tensor:
        """CSR- or uint16-resident counts: the dense float32 (B, G) rows of this rank's cells `indices` (a view of a reused scratch)."""
        B = int(indices.numel())
        if self._xb_scratch is None or self._xb_scratch.shape[0] < B:
            self._xb_scratch = torch.empty((max(B, self.batch_max), self.G), dtype=torch.float32, device=self.device)
        out = self._xb_scratch[:B]
        if self.X_u16 is not None:
            rc = self.lib.scdef_u16_gather_rows(_ptr(self.X_u16), _ptr(indices), B, self.G, _ptr(out), self._stream())
            if rc != 0:
                raise ScdefError(f"scdef_u16_gather_rows failed ({rc})")
            return out
        ip, ix, dv = self.X_csr
        rc = self.lib.scdef_csr_gather_rows(_ptr(ip), _ptr(ix), _ptr(dv), _ptr(indices), B, self.G, _ptr(out), self._stream())
        if rc != 0:
            raise ScdefError(f"scdef_csr_gather_rows failed ({rc})")
        return out

    def _cached_blocks(self):
        key = self.local_grad.flat.data_ptr()
        if getattr(self, "_blk_key", None) != key:
            self._blk_params = _merge(self.local.blocks(self.L, True), self.glob.blocks(self.L, False), self.L)
            self._blk_grads = _merge(self.local_grad.blocks(self.L, True), self.glob_grad.blocks(self.L, False), self.L)
            self._blk_lm = self.local_m.blocks(self.L, True)
            self._blk_lv = self.local_v.blocks(self.L, True)
            self._blk_gm = self.glob_m.blocks(self.L, False)
            self._blk_gv = self.glob_v.blocks(self.L, False)
            self._blk_key = key
        return self._blk_params, self._blk_grads

    def local_grads(self, B: int) -> List[torch.Tensor]:
        """Views of the compact local gradient written by the last LOCAL pass: (2,B,1), (2,B,sumK)."""
        c, z = self.local_grad.views
        return [c.reshape(-1)[: 2 * B].view(2, B, 1), z.reshape(-1)[: 2 * B * self.Kt].view(2, B, self.Kt)]

    def loss_value(self) -> float:
        return float(self.loss_buf.sum().item())

    # ---- optimiser ----------------------------------------------------------------------------
    def _bc_table(self, b1, b2):
        """(1/(1-b1^k), 1/(1-b2^k)) for k = 1.., computed exactly as the C side does for the dense kernel
        (double pow of the float32 betas, rounded to float32); stops once both are exactly 1."""
        key = (float(b1), float(b2))
        if self._bc_key != key:
            cap = 1 << 20
            host = np.empty((cap, 2), dtype=np.float32)
            n = self.lib.scdef_adam_bias_table(float(b1), float(b2), host.ctypes.data_as(C.c_void_p), cap)
            if n < 1:
                raise ScdefError("scdef_adam_bias_table failed")
            self._bc_tab = torch.from_numpy(host[:n].copy()).to(self.device).contiguous()
            self._bc_key = key
        return self._bc_tab

    def _lazy_call(self, phase, indices, step, lr, b1, b2, eps, freeze):
        tab = self._bc_table(b1, b2)
        fz = (C.c_int32 * MAX_LAYERS)()
        for l in freeze:
            fz[int(l)] = 1
        params, grads = self._cached_blocks()
        B = 0 if indices is None else int(indices.numel())
        rc = self.lib.scdef_adam_local_lazy(
            self.ctx, phase, C.byref(params), C.byref(grads), C.byref(self._blk_lm), C.byref(self._blk_lv),
            C.c_void_p(indices.data_ptr()) if B > 0 else None, B, int(step), lr, b1, b2, eps, fz,
            _ptr(self.row_step), _ptr(tab), int(tab.shape[0]), self._stream())
        self._check(rc, "scdef_adam_local_lazy")

    def local_catchup(self, indices: torch.Tensor):
        """Lazy mode: bring the minibatch rows up to date before the objective reads them."""
        if self.lazy_adam and self._lazy_pending:
            lr, b1, b2, eps, freeze = self._lazy_hyper
            self._lazy_call(_lib.LAZY_CATCHUP, indices, self.local_step + 1, lr, b1, b2, eps, freeze)

    def local_sync(self):
        """Lazy mode: replay the pending zero-gradient steps of every row (before anything reads all rows)."""
        if getattr(self, "lazy_adam", False) and self._lazy_pending:
            lr, b1, b2, eps, freeze = self._lazy_hyper
            self._lazy_call(_lib.LAZY_CATCHUP_ALL, None, self.local_step, lr, b1, b2, eps, freeze)
            self._lazy_pending = False

    def adam_local(self, indices: torch.Tensor, lr: float, freeze_z_layers: Sequence[int] = (),
                   b1=0.9, b2=0.999, eps=1e-8):
        """optax.adam over the local blocks (`_scdef.py:3112-3116`).  Dense: one pass over all N rows.
        Lazy (default): only the minibatch rows are written now; the other rows' zero-gradient steps are
        replayed on demand — the same fp32 operations, bit-identical parameters."""
        B = int(indices.numel())
        if self.lazy_adam:
            hyper = (float(lr), float(b1), float(b2), float(eps), tuple(int(l) for l in freeze_z_layers))
            if self._lazy_pending and self._lazy_hyper != hyper:
                self.local_sync()
            self._lazy_hyper = hyper
            if self._lazy_pending:  # rows of this minibatch may still be behind (if local_catchup was not called)
                self._lazy_call(_lib.LAZY_CATCHUP, indices, self.local_step + 1, *hyper)
            self.local_step += 1
            self._lazy_call(_lib.LAZY_UPDATE, indices, self.local_step, *hyper)
            self._lazy_pending = True
            return
        self.local_step += 1
        fz = (C.c_int32 * MAX_LAYERS)()
        for l in freeze_z_layers:
            fz[int(l)] = 1
        params, grads = self._cached_blocks()
        rc = self.lib.scdef_adam_local(
            self.ctx, C.byref(params), C.byref(grads), C.byref(self._blk_lm), C.byref(self._blk_lv),
            C.c_void_p(indices.data_ptr()) if B > 0 else None, B, self.local_step, lr, b1, b2, eps, fz, self._stream())
        self._check(rc, "scdef_adam_local")

    def adam_global(self, lr: float, freeze_w=False, b1=0.9, b2=0.999, eps=1e-8):
        self.global_step += 1
        params, grads = self._cached_blocks()
        rc = self.lib.scdef_adam_global(
            self.ctx, C.byref(params), C.byref(grads), C.byref(self._blk_gm), C.byref(self._blk_gv),
            self.global_step, lr, b1, b2, eps, int(bool(freeze_w)), self._stream())
        self._check(rc, "scdef_adam_global")

    # ---- summaries -----------------------------------------------------------------------------
    def posterior(self):
        """(means, variances) of every block as lists shaped like one half of the params."""
        self.local_sync()
        lshapes, gshapes = _block_shapes(self.N, self.Ks, self.nb, self.G)
        half = lambda shapes: [(n, s[1:]) for n, s in shapes]
        lm, lv = _FlatGroup(half(lshapes), self.device), _FlatGroup(half(lshapes), self.device)
        gm, gv = _FlatGroup(half(gshapes), self.device), _FlatGroup(half(gshapes), self.device)
        params = _merge(self.local.blocks(self.L, True), self.glob.blocks(self.L, False), self.L)
        means = _merge(lm.blocks(self.L, True), gm.blocks(self.L, False), self.L)
        vars_ = _merge(lv.blocks(self.L, True), gv.blocks(self.L, False), self.L)
        rc = self.lib.scdef_posterior(self.ctx, C.byref(params), C.byref(means), C.byref(vars_), self._stream())
        self._check(rc, "scdef_posterior")
        return (lm.views, gm.views), (lv.views, gv.views)

    def assignment_counts(self) -> List[torch.Tensor]:
        """Per layer, the number of this rank's cells whose posterior-mean z is largest at each factor
        (`filter_factors` / `get_effective_factors`, `_scdef.py:3540-3549, 1378-1385`): int64 device tensors (K_l,)."""
        self.local_sync()
        counts = torch.empty(self.Kt, dtype=torch.int64, device=self.device)
        params, _ = self._cached_blocks()
        rc = self.lib.scdef_assignment_counts(self.ctx, C.byref(params), _ptr(counts), self._stream())
        self._check(rc, "scdef_assignment_counts")
        offs = np.concatenate([[0], np.cumsum(self.Ks)])
        return [counts[offs[l] : offs[l + 1]] for l in range(self.L)]

    def profile(self, on: bool, kinds: Optional[Sequence[str]] = None):
        """Time every kernel family (kinds=None) or only the named ones (e.g. ["lik_main"])."""
        flag = int(bool(on))
        if on and kinds is not None:
            flag = sum(1 << (_lib.PROF_KINDS.index(k) + 1) for k in kinds)
        self._check(self.lib.scdef_profile_enable(self.ctx, flag), "scdef_profile_enable")

    def profile_read(self, reset=True):
        """{kind: (ms_sum, n_timed, n_launched)}, total kernel launches through the ABI."""
        n = len(_lib.PROF_KINDS)
        ms = (C.c_double * n)()
        nt = (C.c_int64 * n)()
        nl = (C.c_int64 * n)()
        tot = C.c_int64()
        self._check(self.lib.scdef_profile_read(self.ctx, ms, nt, nl, C.byref(tot), int(reset)), "scdef_profile_read")
        return {k: (ms[i], nt[i], nl[i]) for i, k in enumerate(_lib.PROF_KINDS)}, int(tot.value)

    def noise_fill(self, noise: NoiseSpec, kind: int, layer: int, S: int, B: int, shape):
        out = torch.empty((S,) + tuple(shape), dtype=torch.float32, device=self.device)
        nz = self._noise_struct(noise, [])
        rc = self.lib.scdef_noise_fill(self.ctx, C.byref(nz), kind, layer, S, B, _ptr(out), self._stream())
        self._check(rc, "scdef_noise_fill")
        return out
