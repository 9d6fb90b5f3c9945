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
