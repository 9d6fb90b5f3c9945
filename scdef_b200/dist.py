"""Data-parallel plumbing: cell sharding, replicated count statistics, gradient all-reduce.

The reference has no multi-device code; this is the B200-native addition of
§8(e): cells (rows of X, the local blocks and their Adam state) are block-sharded over the
ranks, the global blocks are replicated, and the only exchange of the hot loop is ONE
all-reduce (sum) of the flat global-gradient buffer per step.  Every rank advances the same
`RandomState(0)` permutation, so a step on R ranks with `batch_size=b` is exactly the
reference's step with `batch_size=R*b` (bit-exact minibatch membership).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class ShardInfo:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    group: object = None

    @staticmethod
    def from_group(group) -> "ShardInfo":
        if group is None:
            return ShardInfo()
        import os

        import torch.distributed as td

        pg = None if group == "world" else group
        return ShardInfo(td.get_rank(pg), td.get_world_size(pg), int(os.environ.get("LOCAL_RANK", 0)), pg)


def all_reduce_sum(t, shard: ShardInfo):
    import torch.distributed as td

    td.all_reduce(t, op=td.ReduceOp.SUM, group=shard.group)
    return t


def _allreduce_np(a: np.ndarray, shard: ShardInfo, device=None) -> np.ndarray:
    if shard.world == 1:
        return a
    import torch

    import torch.distributed as td

    t = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    elif td.get_backend(shard.group) == "nccl":  # NCCL only moves device tensors
        t = t.cuda()
    all_reduce_sum(t, shard)
    return t.cpu().numpy()


def _allgather_np(a: np.ndarray, shard: ShardInfo) -> list:
    if shard.world == 1:
        return [a]
    import torch.distributed as td

    out = [None] * shard.world
    td.all_gather_object(out, a, group=shard.group)
    return out


def shard_bounds(n_local: int, shard: ShardInfo):
    """(start, n_total, all_counts) of this rank's block in the global cell numbering."""
    counts = [int(c) for c in _allgather_np(np.int64(n_local), shard)]
    start = int(sum(counts[: shard.rank]))
    return start, int(sum(counts)), counts


def count_statistics(X, batch_labels, shard: ShardInfo, eps=1e-12):
    """The constants `load_adata` / `update_model_priors` derive from the counts
    (`_scdef.py:498-573, 1279-1287`), computed over ALL ranks' cells.  float64 throughout."""
    is_torch = hasattr(X, "is_cuda")
    dev = X.device if (is_torch and X.is_cuda and shard.world > 1) else None
    n_local, G = int(X.shape[0]), int(X.shape[1])
    if is_torch:
        import torch

        lib = X.sum(dim=1, dtype=torch.float64).cpu().numpy()
    else:
        lib = np.asarray(X, dtype=np.float64).sum(1) if X.dtype != np.float32 else X.sum(1, dtype=np.float64)
    has_key = batch_labels is not None
    if has_key:
        local_vals = np.unique(batch_labels)
        batches = np.unique(np.concatenate([np.atleast_1d(v) for v in _allgather_np(local_vals, shard)]))
        batch_id = np.searchsorted(batches, batch_labels).astype(np.int32)
    else:
        batches = np.array([""])
        batch_id = np.zeros(n_local, dtype=np.int32)
    nb = len(batches)

    def gene_sums(mask=None):
        if is_torch:
            import torch

            rows = X if mask is None else X[torch.as_tensor(mask, device=X.device)]
            return rows.sum(dim=0, dtype=torch.float64).cpu().numpy()
        rows = X if mask is None else X[mask]
        return rows.sum(0, dtype=np.float64)

    # sufficient statistics: per batch [count, sum lib, sum lib^2] and gene sums
    st = np.zeros((nb + 1, 3))
    gs = np.zeros((nb + 1, G))
    st[nb] = [n_local, lib.sum(), (lib**2).sum()]
    gs[nb] = gene_sums()
    if nb > 1:
        for i in range(nb):
            m = batch_id == i
            st[i] = [m.sum(), lib[m].sum(), (lib[m] ** 2).sum()]
            gs[i] = gene_sums(np.where(m)[0])
    st = _allreduce_np(st, shard, dev)
    gs = _allreduce_np(gs, shard, dev)
    n_total = int(round(st[nb, 0]))
    mean_lib = st[nb, 1] / n_total
    var_lib = st[nb, 2] / n_total - mean_lib**2
    blr = np.full((n_local, 1), mean_lib / var_lib)
    mu = gs[nb] / n_total + eps
    gene_ratio_init = (mu.mean() / mu)[None, :]
    gene_means, gene_vars = gs[nb].mean(), gs[nb].var()
    gene_ratio = np.float64(gene_means / gene_vars)
    if nb > 1:
        gene_ratio = np.ones((nb, G))
        gene_ratio_init = np.ones((nb, G))
        for i in range(nb):
            cnt = st[i, 0]
            mb = st[i, 1] / cnt
            vb = st[i, 2] / cnt - mb**2
            blr[batch_id == i] = mb / (vb + eps)
            mu_b = gs[i] / cnt + eps
            gene_ratio_init[i] = mu_b.mean() / mu_b
            gene_ratio[i] = gs[i].mean() / gs[i].var()
    all_lib = np.concatenate(_allgather_np(lib.astype(np.float64), shard))
    return dict(
        n_cells_total=n_total, lib=lib, batch_lib_ratio=blr, gene_ratio=gene_ratio,
        gene_ratio_init=gene_ratio_init, gene_means=gene_means, gene_vars=gene_vars,
        batch_id=batch_id, batches=batches, median_lib=float(np.median(all_lib)), mean_lib=float(mean_lib),
        has_key=has_key,
    )


def local_members(perm_slice: np.ndarray, start: int, n_local: int) -> np.ndarray:
    """Rows of this rank (local numbering) that belong to the global minibatch `perm_slice`,
    in the order they appear in it."""
    m = (perm_slice >= start) & (perm_slice < start + n_local)
    return perm_slice[m] - start
