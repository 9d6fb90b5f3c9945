"""Minibatch stream — `data_stream`, `/root/reference/src/scdef/models/_scdef.py:3167-3173`.

Index selection is the reference's, bit for bit: `np.random.RandomState(0)`, a fresh
`permutation(n_cells)` per epoch, contiguous slices of `batch_size`, short last slice.  The
permutation is drawn on the host (it IS NumPy's algorithm) and uploaded once per epoch; batches
are device views of it.  With the counts resident in HBM nothing else moves per step; in
"host" placement the rows `X[batch_idx]` are gathered into pinned memory and copied each step,
as the reference does (`jnp.array(self.X[batch_idx])`).

Data parallel, counts resident (`migrator` given): position-based ownership, see `dist.epoch_layout` — at
every epoch boundary the per-cell rows move so that this rank's minibatch i is its contiguous row range
[i*b, (i+1)*b); every rank gets exactly b rows per step.  In host placement the counts stay in the host
memory of their home rank: each step a rank stages the rows of the global slice that it is home to, and ONE
all-to-all on the device (about 5 MB per rank) hands every row to the rank that processes it.  Otherwise
(row counts that cannot be kept) ownership is static and a rank takes the members of each global slice
that it owns.
"""
from __future__ import annotations

import time

import os
from typing import Optional

import numpy as np
import torch

from . import dist as _dist


class MinibatchStream:
    def __init__(self, n_total: int, global_batch: int, shard: _dist.ShardInfo, device,
                 host_X=None, n_local=None, start=None, pinned_slots: Optional[int] = None, migrator=None, tensors_fn=None,
                 stage=(1, 3, 2)):
        """stage = (staging threads, pinned / device slots, minibatches staged ahead) of the host placement; the
        environment variables SCDEF_STAGE_WORKERS / _SLOTS / _DEPTH override it."""
        self.n_total, self.gb, self.shard, self.device = int(n_total), int(global_batch), shard, torch.device(device)
        self.rng = np.random.RandomState(0)
        num_complete, leftover = divmod(self.n_total, self.gb)
        self.num_batches = num_complete + bool(leftover)
        if shard.world > 1 and (n_local is None or start is None):
            raise ValueError("sharded stream needs n_local and start")
        self.n_local = self.n_total if n_local is None else int(n_local)
        self.start = 0 if start is None else int(start)
        self.host_X = host_X
        self.migrator, self.tensors_fn = migrator, tensors_fn
        self.row0 = None
        self._perm_epoch = None
        self._move_events = []
        self.host_stats = {"plan_s": [], "join_wait_s": [], "move_enqueue_s": []}
        self.h2d_bytes = 0
        self._slots, self._events, self._slot_i = [], [], 0
        self._staged = {}
        self._copy_stream = None
        self._pool = None
        self._gather_pool = None
        # gather threads of the host placement: a few per rank, never more than the host has cores for all ranks
        cores = os.cpu_count() or 8
        self._gather_workers = int(os.environ.get("SCDEF_GATHER_THREADS", 1))   # measured: more threads halve the gather (0.42 -> 0.29 ms) but do not shorten the step
        if pinned_slots is None:
            pinned_slots = int(os.environ.get("SCDEF_STAGE_SLOTS", stage[1]))
        self._stage_workers = int(os.environ.get("SCDEF_STAGE_WORKERS", stage[0]))
        self._stage_depth = int(os.environ.get("SCDEF_STAGE_DEPTH", stage[2]))   # minibatches staged ahead of the running step
        if host_X is not None:
            G = host_X.shape[1]
            cap = min(self.gb, self.n_total)
            for _ in range(pinned_slots):
                self._slots.append(torch.empty((cap, G), dtype=torch.float32).pin_memory())
                self._events.append(None)
            self._dev_slots = [torch.empty((cap, G), dtype=torch.float32, device=device) for _ in range(pinned_slots)]
            self._copy_stream = torch.cuda.Stream(device=device)
            if migrator is not None:   # window order of the rows received from the home ranks
                self._inv_pin = [torch.empty(cap, dtype=torch.int64).pin_memory() for _ in range(pinned_slots)]
                self._inv_dev = [torch.empty(cap, dtype=torch.int64, device=device) for _ in range(pinned_slots)]
                self._starts = np.concatenate([[0], np.cumsum(migrator.counts)])

    # Position-based ownership: the next epoch's permutation, layout and send/receive plan are pure host work on
    # n_total integers (~0.1 s per million cells); a helper thread does it while the current epoch runs.
    def _layout_job(self, box):
        t0 = time.perf_counter()
        perm = self.rng.permutation(self.n_total)
        lay = _dist.epoch_layout(perm, self.gb, self.migrator.counts)
        box.append(lay + (self.migrator.plan(lay[0], lay[1]), perm))
        self.host_stats["plan_s"].append(round(time.perf_counter() - t0, 4))

    def _prepare_ahead(self):
        import threading

        box = []
        th = threading.Thread(target=self._layout_job, args=(box,), daemon=True)
        th.start()
        self._ahead = (th, box)

    def _next_layout(self, perm):
        lay = _dist.epoch_layout(perm, self.gb, self.migrator.counts)
        return lay + (self.migrator.plan(lay[0], lay[1]), perm)

    def new_epoch(self):
        ahead = getattr(self, "_ahead", None)
        if ahead is not None:   # the helper thread drew this epoch's permutation and planned its move
            t0 = time.perf_counter()
            ahead[0].join()
            self.host_stats["join_wait_s"].append(round(time.perf_counter() - t0, 4))
            self._ahead = None
            self._apply_layout(*ahead[1][0])
            return
        perm = self.rng.permutation(self.n_total)
        bounds = np.minimum(np.arange(self.num_batches + 1, dtype=np.int64) * self.gb, self.n_total)
        self.n_glob = np.diff(bounds)
        if self.migrator is not None:
            self._apply_layout(*self._next_layout(perm))
            return
        if self.shard.world == 1:
            local, offs = perm, bounds
        else:
            m = (perm >= self.start) & (perm < self.start + self.n_local)
            local = perm[m] - self.start
            offs = np.concatenate([[0], np.cumsum(np.add.reduceat(m.astype(np.int64), bounds[:-1]))])
            # static ownership: the members of a slice that fall into the lower ranks' blocks come first in the
            # noise numbering, so every rank draws its per-cell normals from a DISJOINT window of the slice's stream
            # (with (0, 0) all ranks would read rows [0, B) of the same stream: correlated noise across ranks)
            self.row0 = np.add.reduceat((perm < self.start).astype(np.int64), bounds[:-1])
        self.perm_host = np.ascontiguousarray(local, dtype=np.int64)
        self.offs = offs
        self.perm_dev = torch.from_numpy(self.perm_host).to(self.device, non_blocking=False)
        self.h2d_bytes += self.perm_host.nbytes
        self._staged = {}

    def _apply_layout(self, rank_of, row_of, offs_all, row0_all, plan, perm):
        self._perm_epoch = perm
        self._rank_pos = _dist._position_layout(self.n_total, self.gb, tuple(int(c) for c in self.migrator.counts))[0]
        bounds = np.minimum(np.arange(self.num_batches + 1, dtype=np.int64) * self.gb, self.n_total)
        self.n_glob = np.diff(bounds)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if self.device.type == "cuda" else None
        if ev:
            ev[0].record()
        t0 = time.perf_counter()
        self.migrator.move(self.tensors_fn(), rank_of, row_of, plan)
        self.host_stats["move_enqueue_s"].append(round(time.perf_counter() - t0, 4))
        self._prepare_ahead()
        if ev:
            ev[1].record()
            self._move_events.append(ev)
        offs, self.row0 = offs_all[self.shard.rank], row0_all[self.shard.rank]
        if getattr(self, "_arange_dev", None) is None:
            self._arange_host = np.arange(self.n_local, dtype=np.int64)
            self._arange_dev = torch.from_numpy(self._arange_host).to(self.device)
            self.h2d_bytes += self._arange_host.nbytes
        self.perm_host, self.perm_dev, self.offs = self._arange_host, self._arange_dev, offs
        self._staged = {}

    def noise_window(self, it: int):
        """(row offset, total rows) of this rank's rows inside global minibatch `it` for the per-cell Philox draws.
        Position-based ownership: the rank's rows ARE the range [offset, offset + B) of the slice, so an R-rank run
        draws the single-rank run's normals.  Static ownership: a disjoint window per rank (members of lower ranks
        first) — independent across ranks, though not the single-rank numbering."""
        if self.row0 is None or self.shard.world == 1:
            return 0, 0
        return int(self.row0[it]), int(self.n_glob[it])

    def finish(self, collective: bool = True):
        """Return every per-cell row to its home rank and row (block order) after the last epoch.
        collective=False: only stop the helper threads (a rank that is failing must not enter the all-to-all)."""
        ahead = getattr(self, "_ahead", None)
        if ahead is not None:
            ahead[0].join()
            self._ahead = None
        if self._pool is not None:          # staging thread of the host placement
            for f in self._staged.values():
                f.cancel()
            self._pool.shutdown(wait=True)
            self._pool, self._staged = None, {}
        if self._gather_pool is not None:
            self._gather_pool.shutdown(wait=True)
            self._gather_pool = None
        if collective and self.migrator is not None and not self.migrator.at_home:
            self.migrator.go_home(self.tensors_fn())

    def reshard_ms(self):
        """Device time of each epoch-boundary move so far (synchronises)."""
        out = []
        for a, b in self._move_events:
            b.synchronize()
            out.append(a.elapsed_time(b))
        return out

    def indices_host(self, it: int) -> np.ndarray:
        return self.perm_host[self.offs[it] : self.offs[it + 1]]

    def _home_rows(self, it: int):
        """Host placement with position-based ownership: (rows of this rank's host block that belong to global
        slice `it`, in slice order; how many of them go to each rank; how many rows of my window come from each
        home rank; the permutation from received order to window order)."""
        me, R = self.shard.rank, self.shard.world
        lo, hi = it * self.gb, min((it + 1) * self.gb, self.n_total)
        sl = self._perm_epoch[lo:hi]
        home = np.searchsorted(self._starts, sl, side="right") - 1
        dest = self._rank_pos[lo:hi]
        mine = home == me
        rows = sl[mine] - self._starts[me]
        in_splits = np.bincount(dest[mine], minlength=R).astype(np.int64)
        hw = home[dest == me]                                   # my window is a contiguous range of the slice
        order = np.argsort(hw, kind="stable")                  # received order: by home rank, then slice position
        inv = np.empty(len(hw), dtype=np.int64)
        inv[order] = np.arange(len(hw))
        out_splits = np.bincount(hw, minlength=R).astype(np.int64)
        return rows, in_splits, out_splits, inv

    def _gather(self, rows, out):
        """X[rows] -> pinned slot.  One thread copies scattered 20 KB rows at ~3.6 GB/s (measured: 1.4 ms for 256 rows of
        5000 genes — longer than the step it feeds, which made the whole host placement wait for it); `np.take` releases
        the GIL, so the rows are split over a few gather threads."""
        n = len(rows)
        w = self._gather_workers
        if w <= 1 or n < 4 * w:
            np.take(self.host_X, rows, axis=0, out=out, mode="clip")
            return
        if self._gather_pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._gather_pool = ThreadPoolExecutor(max_workers=w, thread_name_prefix="scdef-gather")
        cuts = [n * i // w for i in range(w + 1)]
        jobs = [self._gather_pool.submit(np.take, self.host_X, rows[a:b], 0, out[a:b], "clip") for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        for j in jobs:
            j.result()

    def _stage_job(self, rows, k, extra=None):
        """Worker thread: gather X[rows] into pinned slot k and start its H2D copy on the copy stream."""
        torch.cuda.set_device(self.device)
        if self._events[k] is not None:
            self._events[k].synchronize()  # the step that last read this slot's device buffer has finished
        pin = self._slots[k][: len(rows)]
        self._gather(rows, pin.numpy())
        Xb = self._dev_slots[k][: len(rows)]
        inv_dev = None
        if extra is not None:
            inv = extra[2]
            self._inv_pin[k][: len(inv)].copy_(torch.from_numpy(inv))
            inv_dev = self._inv_dev[k][: len(inv)]
        with torch.cuda.stream(self._copy_stream):
            Xb.copy_(pin, non_blocking=True)
            if inv_dev is not None:
                inv_dev.copy_(self._inv_pin[k][: len(inv_dev)], non_blocking=True)
                self.h2d_bytes += inv_dev.numel() * 8
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        self.h2d_bytes += pin.numel() * 4
        return k, Xb, ready, extra, inv_dev

    def _stage(self, it: int):
        """Host placement: hand minibatch `it` to the staging thread (FIFO, one slot per job)."""
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(max_workers=self._stage_workers, thread_name_prefix="scdef-stage")
        k = self._slot_i
        self._slot_i = (k + 1) % len(self._slots)
        if self.migrator is not None:
            rows, in_splits, out_splits, inv = self._home_rows(it)
            self._staged[it] = self._pool.submit(self._stage_job, rows, k, (in_splits, out_splits, inv))
        else:
            self._staged[it] = self._pool.submit(self._stage_job, self.indices_host(it), k)

    def prefetch(self, it: int):
        """Start staging minibatches `it` and `it+1` of the current epoch (host placement only).  Called by the hot loop
        right after step it-1 has been enqueued: the host gather (a helper thread) and the H2D copy (a copy stream)
        overlap that step's kernels and the main thread's launches."""
        if self.host_X is None:
            return
        for j in range(it, it + self._stage_depth):
            if 0 <= j < self.num_batches and j not in self._staged and len(self._staged) < len(self._slots) - 1:
                self._stage(j)

    def batch(self, it: int):
        """(device int64 indices of this rank, device X rows or None, global minibatch size)."""
        idx = self.perm_dev[self.offs[it] : self.offs[it + 1]]
        Xb = None
        if self.host_X is not None:
            if it not in self._staged:
                self._stage(it)
            k, Xb, ready, extra, inv_dev = self._staged.pop(it).result()
            torch.cuda.current_stream(self.device).wait_event(ready)
            if extra is not None:   # hand every staged row to the rank that processes it, then restore window order
                recv = self.migrator.exchange(Xb, extra[0], extra[1])
                Xb = recv.index_select(0, inv_dev)
            done = torch.cuda.Event()   # recorded by release(): the compute stream is finished with this slot
            self._pending_release = (k, done)
        return idx, Xb, int(self.n_glob[it])

    def release(self):
        """Marks the end of the step that consumed the last batch (its device slot may be overwritten afterwards)."""
        pr = getattr(self, "_pending_release", None)
        if pr is not None:
            k, done = pr
            done.record(torch.cuda.current_stream(self.device))
            self._events[k] = done
            self._pending_release = None
