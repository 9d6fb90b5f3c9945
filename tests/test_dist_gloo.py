"""CPU, world_size 2 over gloo: the host side of the data-parallel path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from scdef_b200.synth import synth_counts_numpy


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from scdef_b200 import MiniAnnData, scDEF
    from scdef_b200.dist import all_reduce_sum, shard_bounds
    from scdef_b200.stream import MinibatchStream

    X, b = synth_counts_numpy(121, 40, n_batches=3, seed=0)
    labels = np.array(["u", "v", "w"])[b]
    cut = 70
    sl = slice(0, cut) if rank == 0 else slice(cut, 121)
    m = scDEF(MiniAnnData(X[sl], obs={"batch": labels[sl]}), batch_key="batch", layer_sizes=[6, 3, 1],
              process_group="world")
    start, n_total, counts = shard_bounds(m.n_cells, m._shard)
    st = MinibatchStream(n_total, 16 * world, m._shard, torch.device("cpu"), n_local=m.n_cells, start=start)
    st.new_epoch()
    idx = [st.batch(i)[0].numpy() + start for i in range(st.num_batches)]
    g = torch.full((5,), float(rank + 1))
    all_reduce_sum(g, m._shard)
    res = dict(alpha=m.alpha, blr=m.batch_lib_ratio, gene_ratio=np.asarray(m.gene_ratio), caf=m.cell_alpha_factor,
               n_total=m.n_cells_total, start=start, idx=idx, gsum=g.numpy(), gp0=m.global_params[0],
               gp_w0=m.global_params[2], batch_id=m.batch_id, batches=list(m.batches))
    torch.save(res, out + f".{rank}")
    td.destroy_process_group()


def test_two_rank_statistics_and_stream_equal_single_rank(tmp_path):
    port = _free_port()
    out = str(tmp_path / "res")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0 = torch.load(out + ".0", weights_only=False)
    r1 = torch.load(out + ".1", weights_only=False)
    from scdef_b200 import MiniAnnData, scDEF
    from oracle import host_ref

    X, b = synth_counts_numpy(121, 40, n_batches=3, seed=0)
    labels = np.array(["u", "v", "w"])[b]
    m = scDEF(MiniAnnData(X, obs={"batch": labels}), batch_key="batch", layer_sizes=[6, 3, 1])
    assert r0["n_total"] == r1["n_total"] == 121 and (r0["start"], r1["start"]) == (0, 70)
    assert r0["alpha"] == pytest.approx(m.alpha, rel=1e-12) and r1["alpha"] == pytest.approx(m.alpha, rel=1e-12)
    np.testing.assert_allclose(np.concatenate([r0["blr"], r1["blr"]]), m.batch_lib_ratio, rtol=1e-9)
    np.testing.assert_allclose(np.concatenate([r0["caf"], r1["caf"]]), m.cell_alpha_factor, rtol=1e-12)
    np.testing.assert_allclose(r0["gene_ratio"], np.asarray(m.gene_ratio), rtol=1e-9)
    np.testing.assert_allclose(r1["gene_ratio"], np.asarray(m.gene_ratio), rtol=1e-9)
    np.testing.assert_array_equal(np.concatenate([r0["batch_id"], r1["batch_id"]]), m.batch_id)
    # replicated global blocks are initialised identically on every rank
    np.testing.assert_array_equal(r0["gp_w0"], r1["gp_w0"])
    np.testing.assert_allclose(r0["gp0"], m.global_params[0], rtol=1e-5, atol=1e-6)
    # the union of the two ranks' minibatches is the reference's minibatch of size 2*16
    ref = host_ref.data_stream_indices(121, 32)
    for a, c in zip(r0["idx"], r1["idx"]):
        expect = next(ref)
        assert sorted(np.concatenate([a, c]).tolist()) == sorted(expect.tolist())
    np.testing.assert_array_equal(r0["gsum"], np.full(5, 3.0))


# ---- position-based ownership (dist.epoch_layout / RowMigrator) ----------------------------------
@pytest.mark.parametrize("n,gb,counts", [(121, 32, (70, 51)), (1000, 64, (250, 250, 250, 250)), (96, 32, (48, 48)),
                                         (77, 24, (26, 26, 25))])
def test_epoch_layout_is_balanced_and_keeps_the_reference_minibatches(n, gb, counts):
    from oracle import host_ref
    from scdef_b200.dist import epoch_layout

    R, b = len(counts), gb // len(counts)
    rng = np.random.RandomState(0)
    ref = host_ref.data_stream_indices(n, gb)
    for _ in range(2):
        perm = rng.permutation(n)
        rank_of, row_of, offs, row0 = epoch_layout(perm, gb, counts)
        for r in range(R):  # every rank keeps its row count and every row is used exactly once
            assert sorted(row_of[rank_of == r].tolist()) == list(range(counts[r]))
        inv = {(int(rank_of[g]), int(row_of[g])): g for g in range(n)}
        nb = (n + gb - 1) // gb
        for i in range(nb):
            expect = next(ref)
            got = []
            for r in range(R):
                rows = range(int(offs[r][i]), int(offs[r][i + 1]))
                if i < n // gb:
                    assert len(rows) == b           # exactly b rows per rank on every full minibatch
                got.append([inv[(r, j)] for j in rows])
                # rank r's rows are the contiguous window [row0, row0+len) of the reference's slice, in order
                assert got[-1] == expect[int(row0[r][i]) : int(row0[r][i]) + len(rows)].tolist()
            assert np.concatenate(got).astype(np.int64).tolist() == expect.tolist()   # bit-exact, in order


def test_epoch_layout_refuses_counts_it_cannot_keep():
    from scdef_b200.dist import epoch_layout

    assert epoch_layout(np.arange(100), 32, (90, 10)) is None     # rank 1 owns fewer rows than 3 full minibatches need
    assert epoch_layout(np.arange(100), 33, (50, 50)) is None     # global batch not divisible by the world size


def _migrate_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from scdef_b200.dist import RowMigrator, ShardInfo, epoch_layout

    counts = (70, 51)
    start = sum(counts[:rank])
    gid = np.arange(start, start + counts[rank])
    # three per-cell arrays of different widths/dtypes, each row tagged with its global cell id
    A = torch.from_numpy(np.stack([gid * 10.0 + c for c in range(7)], 1).astype(np.float32))
    Bt = torch.from_numpy(gid.astype(np.int32)).view(-1, 1)
    P = torch.zeros(2, counts[rank], 3)
    P[0] = torch.from_numpy(gid[:, None] + np.zeros((1, 3))).float()
    P[1] = -P[0]
    tensors = [A, Bt, P[0], P[1]]
    mig = RowMigrator(ShardInfo.from_group("world"), counts, chunk_bytes=4 * 70 * 3)   # forces column chunking
    rng = np.random.RandomState(0)
    seen = []
    for _ in range(3):
        perm = rng.permutation(121)
        rank_of, row_of, offs, row0 = epoch_layout(perm, 32, counts)
        mig.move(tensors, rank_of, row_of)
        here = np.nonzero(rank_of == rank)[0]
        want = np.empty(counts[rank], dtype=np.int64)
        want[row_of[here]] = here
        assert np.array_equal(Bt[:, 0].numpy(), want)
        assert np.array_equal(A.numpy(), np.stack([want * 10.0 + c for c in range(7)], 1).astype(np.float32))
        assert np.array_equal(P[0, :, 1].numpy(), want.astype(np.float32)) and np.array_equal(P[1].numpy(), -P[0].numpy())
        seen.append([Bt[int(offs[rank][i]) : int(offs[rank][i + 1]), 0].numpy().copy() for i in range(len(offs[rank]) - 1)])
    mig.go_home(tensors)
    assert mig.at_home and np.array_equal(Bt[:, 0].numpy(), gid) and np.array_equal(P[0, :, 0].numpy(), gid.astype(np.float32))
    torch.save(seen, out + f".{rank}")
    td.destroy_process_group()


def test_row_migration_two_ranks_round_trip(tmp_path):
    port = _free_port()
    out = str(tmp_path / "mig")
    mp.spawn(_migrate_worker, args=(2, port, out), nprocs=2, join=True)
    s0 = torch.load(out + ".0", weights_only=False)
    s1 = torch.load(out + ".1", weights_only=False)
    from oracle import host_ref

    ref = host_ref.data_stream_indices(121, 32)
    for e in range(3):
        for a, c in zip(s0[e], s1[e]):
            np.testing.assert_array_equal(np.concatenate([a, c]), next(ref))   # rank 0's rows then rank 1's = the slice


@pytest.mark.parametrize("n,gb,counts", [(121, 32, (70, 51)), (1000, 96, (334, 333, 333))])
def test_host_placement_row_routing_delivers_each_window_in_order(n, gb, counts):
    """stream._home_rows (host placement + position-based ownership), emulated without devices: every rank stages
    the rows of the global slice it is home to; after the all-to-all and the window permutation each rank holds
    exactly its window of the reference's slice, in order."""
    from scdef_b200.dist import ShardInfo, _position_layout
    from scdef_b200.stream import MinibatchStream

    R = len(counts)
    starts = np.concatenate([[0], np.cumsum(counts)])
    perm = np.random.RandomState(3).permutation(n)
    rank_pos = _position_layout(n, gb, tuple(counts))[0]
    streams = []
    for r in range(R):
        st = object.__new__(MinibatchStream)
        st.shard, st.gb, st.n_total = ShardInfo(rank=r, world=R), gb, n
        st._perm_epoch, st._starts, st._rank_pos = perm, starts, rank_pos
        streams.append(st)
    for it in range((n + gb - 1) // gb):
        plans = [st._home_rows(it) for st in streams]
        sends = [rows + starts[r] for r, (rows, _, _, _) in enumerate(plans)]           # global ids staged by rank r
        lo, hi = it * gb, min((it + 1) * gb, n)
        for d in range(R):
            recv = []
            for s in range(R):                                                             # all-to-all, rank order
                a = int(plans[s][1][:d].sum())
                chunk = sends[s][a : a + int(plans[s][1][d])]
                assert len(chunk) == plans[d][2][s]
                recv.append(chunk)
            got = np.concatenate(recv)[plans[d][3]]
            want = perm[lo:hi][rank_pos[lo:hi] == d]
            np.testing.assert_array_equal(got, want)


def test_epoch_layout_properties_random_shapes():
    """Property test over random shard sizes / batch sizes: every rank keeps its row count, every full global minibatch
    gives each rank exactly b rows, and concatenating the ranks' windows reproduces the reference's slice in order."""
    from hypothesis import given, settings, strategies as st

    from scdef_b200.dist import RowMigrator, ShardInfo, epoch_layout

    @settings(max_examples=60, deadline=None)
    @given(R=st.integers(1, 5), b=st.integers(1, 9), n_full=st.integers(0, 6), data=st.data())
    def check(R, b, n_full, data):
        tail = [data.draw(st.integers(0, b)) for _ in range(R)]
        if sum(tail) == R * b:          # that would be one more full minibatch
            tail[0] -= 1
        counts = [n_full * b + t for t in tail]
        n = sum(counts)
        if n == 0:
            return
        gb = R * b
        perm = np.random.RandomState(data.draw(st.integers(0, 1000))).permutation(n)
        lay = epoch_layout(perm, gb, counts)
        assert lay is not None
        rank_of, row_of, offs, row0 = lay
        for r in range(R):
            assert sorted(row_of[rank_of == r].tolist()) == list(range(counts[r]))
        where = {(int(rank_of[g]), int(row_of[g])): g for g in range(n)}
        for i in range((n + gb - 1) // gb):
            want = perm[i * gb:(i + 1) * gb]
            got = []
            for r in range(R):
                rows = list(range(int(offs[r][i]), int(offs[r][i + 1])))
                if i < n_full:
                    assert len(rows) == b
                assert int(row0[r][i]) == len(got)
                got += [where[(r, j)] for j in rows]
            assert got == want.tolist()
        # the send / receive plans of all ranks are mutually consistent
        plans = []
        for r in range(R):
            m = RowMigrator(ShardInfo(rank=r, world=R), counts)
            plans.append(m.plan(rank_of, row_of))
        for s in range(R):
            for d in range(R):
                assert plans[s][1][d] == plans[d][3][s]          # what s sends to d is what d expects from s
            assert sorted(plans[s][2].tolist()) == list(range(counts[s]))   # every row of a rank is written exactly once

    check()
