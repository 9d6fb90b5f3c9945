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

import os

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

    if t.is_cuda and td.get_backend(shard.group) != "nccl":  # gloo (tests): stage through the host
        h = t.cpu()
        td.all_reduce(h, op=td.ReduceOp.SUM, group=shard.group)
        t.copy_(h)
        return t
    td.all_reduce(t, op=td.ReduceOp.SUM, group=shard.group)
    return t


class GradReducer:
    """Sum of the flat global-gradient buffer over the ranks, once per step (the only collective of the data path).

    The buffer is laid out [small blocks | W_0] (`Engine.glob_grad.split`).  W_0's gradient — 97 % of the bytes — is final
    before the hierarchy kernel and `k_finish` of the GLOBAL pass have run (`SCDEF_FLAG_LIK_FIRST`,
    `scdef_set_w0_grad_event`), so its all-reduce is enqueued behind that event and runs on NCCL's stream while the pass
    finishes; only the all-reduce of the small blocks (~0.1 MB) starts after the pass.  gloo (CPU tests) and more than
    two ranks (see the measurements below): one plain all-reduce of the whole buffer."""

    def __init__(self, eng, shard: ShardInfo):
        import torch
        import torch.distributed as td

        self.eng, self.shard = eng, shard
        flat, split = eng.glob_grad.flat, eng.glob_grad.split
        # Measured at C5 on B200s (same box, back to back, ms per step with / without the split): 2 GPUs 1.435 / 1.445,
        # 4 GPUs 1.429 / 1.427, 8 GPUs 1.466 / 1.460 — beyond two ranks the cost of a collective is its latency, not its
        # 4 MB, and a second collective costs what the overlap saves.  So: on by default for two ranks only.
        want = os.environ.get("SCDEF_DP_OVERLAP", "auto")
        want = shard.world == 2 if want == "auto" else want != "0"
        self.overlap = bool(flat.is_cuda and shard.world > 1 and td.get_backend(shard.group) == "nccl" and split is not None and want)
        if self.overlap:
            self.small, self.w0 = flat[:split], flat[split:]
            self.side = torch.cuda.Stream(device=flat.device)
            self.event = eng.w0_grad_event()

    @property
    def lik_first(self) -> bool:
        return self.overlap

    def reduce(self):
        """Call right after the GLOBAL pass has been enqueued on the current stream."""
        import torch
        import torch.distributed as td

        if not self.overlap:
            return all_reduce_sum(self.eng.glob_grad.flat, self.shard)
        self.side.wait_event(self.event)
        with torch.cuda.stream(self.side):      # NCCL's stream waits for `side`, i.e. for the W_0 gradient only
            w_big = td.all_reduce(self.w0, op=td.ReduceOp.SUM, group=self.shard.group, async_op=True)
        w_small = td.all_reduce(self.small, op=td.ReduceOp.SUM, group=self.shard.group, async_op=True)   # after the whole pass
        w_big.wait()                             # the compute stream waits for both (no host synchronisation)
        w_small.wait()


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


def _device_lib(X):
    """Library sizes of device-resident counts (float32 / uint16) through the C ABI (`scdef_row_stats`)."""
    import ctypes as C

    import torch

    from . import _lib

    lib = _lib.load()
    n, G = int(X.shape[0]), int(X.shape[1])
    out = torch.empty(max(n, 1), dtype=torch.float32, device=X.device)
    st = C.c_void_p(torch.cuda.current_stream(X.device).cuda_stream)
    fn = lib.scdef_u16_row_stats if X.dtype in (torch.uint16, torch.int16) else lib.scdef_row_stats
    rc = fn(C.c_void_p(X.data_ptr()), n, G, C.c_void_p(out.data_ptr()), None, st)
    if rc != 0:
        raise RuntimeError(f"row statistics failed ({rc})")
    return out[:n].double().cpu().numpy()


def _device_count_stats(X, batch_id, nb):
    """(per-batch [n, sum lib, sum lib^2], per-batch gene sums), slot nb = all cells: `scdef_count_stats` on the device."""
    import ctypes as C

    import torch

    from . import _lib

    lib = _lib.load()
    n, G = int(X.shape[0]), int(X.shape[1])
    dev = X.device
    st = torch.zeros((nb + 1, 3), dtype=torch.float64, device=dev)
    gs = torch.zeros((nb + 1, G), dtype=torch.float64, device=dev)
    bid = torch.as_tensor(np.ascontiguousarray(batch_id, dtype=np.int32), device=dev)
    is16 = 1 if X.dtype in (torch.uint16, torch.int16) else 0
    if not is16 and X.dtype != torch.float32:
        raise TypeError("device-resident counts must be float32 or uint16")
    rc = lib.scdef_count_stats(C.c_void_p(X.data_ptr()), is16, C.c_void_p(bid.data_ptr()), n, G, nb,
                               C.c_void_p(st.data_ptr()), C.c_void_p(gs.data_ptr()), None,
                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"scdef_count_stats failed ({rc})")
    return st.cpu().numpy(), gs.cpu().numpy()


def count_statistics(X, batch_labels, shard: ShardInfo, eps=1e-12):
    """The constants `load_adata` / `update_model_priors` derive from the counts
    (`_scdef.py:498-573, 1279-1287`), computed over ALL ranks' cells.  float64 throughout."""
    is_torch = hasattr(X, "is_cuda")
    is_sparse = hasattr(X, "tocsr")
    dev = X.device if (is_torch and X.is_cuda and shard.world > 1) else None
    n_local, G = int(X.shape[0]), int(X.shape[1])
    on_device = bool(is_torch and X.is_cuda)
    if on_device:
        lib = _device_lib(X)
    elif is_sparse:
        lib = np.asarray(X.sum(axis=1, dtype=np.float64)).ravel()
    elif is_torch:
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
        if is_sparse:
            rows = X if mask is None else X[mask]
            return np.asarray(rows.sum(axis=0, dtype=np.float64)).ravel()
        if is_torch:
            import torch

            rows = X if mask is None else X[torch.as_tensor(mask, device=X.device)]
            return rows.sum(dim=0, dtype=torch.float64).cpu().numpy()
        rows = X if mask is None else X[mask]
        return rows.sum(0, dtype=np.float64)

    # sufficient statistics: per batch [count, sum lib, sum lib^2] and gene sums
    if on_device:
        # counts resident in HBM (float32 or uint16): one pass of `scdef_count_stats` over them
        st, gs = _device_count_stats(X, batch_id, nb)
    else:
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


# ---------------------------------------------------------------------------------------------
# Position-based ownership ("epoch resharding").
#
# With cells owned statically, the number of members of a global minibatch that fall into one
# rank's block is binomial (b ± sqrt(b)); the slowest rank sets the step time and the likelihood
# kernel works in units of 128 cells, so b = 256 costs 3 units instead of 2 on about half the ranks.
# Instead, at the start of every epoch the cells MOVE: rank r owns positions [r*b, (r+1)*b) of every
# global slice perm[i*R*b : (i+1)*R*b] of that epoch's permutation, stored so that its minibatch i is
# the contiguous row range [i*b, (i+1)*b).  Every rank then has exactly b rows per step, the global
# minibatch is the reference's bit for bit, and the per-cell Philox rows line up with the single-rank
# run (scdef_noise.local_row_offset).  The move is one all-to-all of the per-cell rows per epoch
# (counts, variational parameters, Adam moments, per-cell constants) over NVLink.
# ---------------------------------------------------------------------------------------------
_POS_CACHE = {}


def _position_layout(N: int, gb: int, counts: tuple):
    """The permutation-independent part of `epoch_layout`: owner rank and row of every POSITION of an epoch."""
    key = (N, gb, counts)
    if key in _POS_CACHE:
        return _POS_CACHE[key]
    R = len(counts)
    if gb % R != 0:
        return None
    b = gb // R
    n_full, rem = divmod(N, gb)
    tail = np.asarray(counts, dtype=np.int64) - n_full * b
    if np.any(tail < 0) or int(tail.sum()) != rem:
        return None
    cum = np.concatenate([[0], np.cumsum(tail)])
    pos = np.arange(N, dtype=np.int64)
    i = pos // gb
    t = pos - i * gb
    rank = np.minimum(t // b, R - 1)
    row = i * b + (t - rank * b)
    if rem:
        tl = slice(n_full * gb, N)
        rr = np.searchsorted(cum, t[tl], side="right") - 1
        rank[tl] = rr
        row[tl] = n_full * b + t[tl] - cum[rr]
    offs, row0 = [], []
    for r in range(R):
        o = np.arange(n_full + 1, dtype=np.int64) * b
        z = np.full(n_full, r * b, dtype=np.int64)
        if rem:
            o = np.concatenate([o, [n_full * b + tail[r]]])
            z = np.concatenate([z, [cum[r]]])
        offs.append(o)
        row0.append(z)
    _POS_CACHE.clear()   # one layout at a time is enough (a `_learn` call uses one)
    _POS_CACHE[key] = (rank.astype(np.int32), row, offs, row0)
    return _POS_CACHE[key]


def epoch_layout(perm: np.ndarray, global_batch: int, counts):
    """(rank_of, row_of, offs, row0) for one epoch, or None when the ranks' row counts cannot be kept.

    rank_of[g], row_of[g]: where global cell g lives during this epoch.  Rank r keeps exactly counts[r]
    rows: the full global minibatches give every rank b = global_batch/R rows, the short last one gives
    rank r its remaining counts[r] - n_full*b rows.  offs[r]: row boundaries of rank r's minibatches;
    row0[r][i]: position of rank r's first row inside global minibatch i (the noise row offset)."""
    perm = np.asarray(perm)
    lay = _position_layout(len(perm), int(global_batch), tuple(int(c) for c in counts))
    if lay is None:
        return None
    rank_pos, row_pos, offs, row0 = lay
    rank_of = np.empty(len(perm), dtype=np.int32)
    row_of = np.empty(len(perm), dtype=np.int64)
    rank_of[perm] = rank_pos
    row_of[perm] = row_pos
    return rank_of, row_of, offs, row0


class RowMigrator:
    """Moves per-cell rows between ranks so that cell g ends up at (rank_of[g], row_of[g]).

    Every rank tracks the placement of ALL cells on the host (two integer arrays of n_total), so the
    send and receive orders are computed locally and the only communication is the payload itself:
    one `all_to_all_single` per tensor (column-chunked so the staging buffers stay below `chunk_bytes`)."""

    def __init__(self, shard: ShardInfo, counts, chunk_bytes: int = 1 << 30):
        self.shard = shard
        self.counts = np.asarray(counts, dtype=np.int64)
        starts = np.concatenate([[0], np.cumsum(self.counts)])
        self.n_total = int(starts[-1])
        self.home_rank = np.repeat(np.arange(len(self.counts), dtype=np.int32), self.counts)
        self.home_row = np.arange(self.n_total, dtype=np.int64) - starts[self.home_rank]
        self.cur_rank, self.cur_row = self.home_rank, self.home_row
        self.chunk_bytes = int(chunk_bytes)
        self.moved_bytes = 0
        self.n_moves = 0

    @property
    def at_home(self) -> bool:
        return self.cur_rank is self.home_rank and self.cur_row is self.home_row

    def plan(self, rank_of, row_of):
        """(send_rows, in_splits, place_rows, out_splits) of this rank for the move to (rank_of, row_of)."""
        me, R = self.shard.rank, self.shard.world
        starts = np.concatenate([[0], np.cumsum(self.counts)])
        # destination slot (rank-major, then row) of every cell: a permutation of 0..n_total-1, so "sorted by
        # (destination rank, destination row)" is a scatter, not a sort
        slot = starts[rank_of] + row_of
        src_rank = np.empty(self.n_total, dtype=np.int16)
        src_row = np.empty(self.n_total, dtype=np.int64)
        src_rank[slot] = self.cur_rank
        src_row[slot] = self.cur_row
        from_me = np.nonzero(src_rank == me)[0]                       # ascending destination slot
        send_rows = src_row[from_me]
        in_splits = np.diff(np.searchsorted(from_me, starts)).astype(np.int64)
        mine = src_rank[starts[me] : starts[me + 1]]                  # source rank of each of my future rows
        place_rows = np.argsort(mine, kind="stable").astype(np.int64)  # received order: by source rank, then row
        out_splits = np.bincount(mine, minlength=R).astype(np.int64)
        if len(send_rows) != self.counts[me]:
            raise ValueError("row migration must keep every rank's row count")
        return send_rows, in_splits, place_rows, out_splits

    def _exchange(self, send, in_splits, out_splits):
        import torch
        import torch.distributed as td

        n_out = int(out_splits.sum())
        staged = send.is_cuda and td.get_backend(self.shard.group) != "nccl"  # gloo moves host tensors
        s = send.cpu() if staged else send
        recv = s.new_empty((n_out,) + tuple(s.shape[1:]))
        td.all_to_all_single(recv, s, [int(x) for x in out_splits], [int(x) for x in in_splits], group=self.shard.group)
        self.moved_bytes += int(send.numel() * send.element_size())
        return recv.to(send.device) if staged else recv

    def exchange(self, send, in_splits, out_splits):
        """One all-to-all of rows: `send` holds in_splits[d] rows for every rank d (in rank order); returns the
        out_splits[s] rows received from every rank s, in rank order."""
        if self.shard.world == 1:
            return send
        return self._exchange(send.contiguous(), in_splits, out_splits)

    def move(self, tensors, rank_of, row_of, plan=None):
        """In place: every tensor's row for cell g goes from its current placement to (rank_of[g], row_of[g]).
        `tensors`: 2-D views (rows, width) whose row count is this rank's counts[rank].  `plan`: the result of
        `self.plan(rank_of, row_of)` when it was computed ahead of time."""
        import torch

        send_rows, in_splits, place_rows, out_splits = plan if plan is not None else self.plan(rank_of, row_of)
        if self.shard.world > 1 or not np.array_equal(send_rows, place_rows):
            dev = tensors[0].device if tensors else None
            sr = torch.from_numpy(np.ascontiguousarray(send_rows)).to(dev)
            inv = np.empty_like(place_rows)
            inv[place_rows] = np.arange(len(place_rows))             # row j of the result = received row inv[j]
            pr = torch.from_numpy(inv).to(dev)
            for t in tensors:
                n, w = int(t.shape[0]), int(t.shape[1])
                # the chunk width must be the same on every rank: size it by the largest shard
                step = max(1, min(w, self.chunk_bytes // max(1, int(self.counts.max()) * t.element_size())))
                for c0 in range(0, w, step):
                    blk = t[:, c0 : c0 + step]
                    send = blk.index_select(0, sr).contiguous()
                    recv = self._exchange(send, in_splits, out_splits) if self.shard.world > 1 else send
                    blk.copy_(recv.index_select(0, pr))
        self.cur_rank, self.cur_row = rank_of, row_of
        self.n_moves += 1

    def go_home(self, tensors):
        if not self.at_home:
            self.move(tensors, self.home_rank, self.home_row)
        self.cur_rank, self.cur_row = self.home_rank, self.home_row
