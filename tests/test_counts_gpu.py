"""device-resident count storage and the statistics `load_adata` derives from the counts
(`/root/reference/src/scdef/models/_scdef.py:491-573`), on the device.
  * `scdef_count_stats` (library sizes, per-batch moments, per-batch gene sums; float32 and uint16 counts) against the
    NumPy path of `dist.count_statistics`, which tests/test_ref_pin.py pins to the reference's own constructor;
  * `x_placement="device_u16"`: raw counts kept as uint16 in HBM — the minibatch converted per step is the dense float32
    minibatch bit for bit, so training follows the dense path exactly."""
import numpy as np
import pytest
import torch

from scdef_b200 import MiniAnnData, dist, scDEF
from scdef_b200.synth import synth_counts_numpy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nb", [1, 3])
@pytest.mark.parametrize("u16", [False, True])
def test_device_count_statistics_equal_the_numpy_path(nb, u16):
    X, batch = synth_counts_numpy(1000, 203, n_batches=nb, seed=5)   # 203 genes: not a multiple of anything
    labels = np.array([f"b{b}" for b in batch]) if nb > 1 else None
    shard = dist.ShardInfo()
    want = dist.count_statistics(X, labels, shard)
    Xd = torch.as_tensor(X, device="cuda")
    if u16:
        Xd = Xd.to(torch.uint16)
    got = dist.count_statistics(Xd, labels, shard)
    assert got["n_cells_total"] == want["n_cells_total"]
    np.testing.assert_array_equal(got["lib"], want["lib"])                      # integer sums: exact
    np.testing.assert_array_equal(got["batch_id"], want["batch_id"])
    for k in ("batch_lib_ratio", "gene_ratio", "gene_ratio_init", "gene_means", "gene_vars", "median_lib", "mean_lib"):
        np.testing.assert_allclose(np.asarray(got[k], dtype=float), np.asarray(want[k], dtype=float), rtol=1e-12, err_msg=k)


def test_u16_resident_counts_equal_dense():
    X, _ = synth_counts_numpy(300, 120, seed=2)
    md = scDEF(MiniAnnData(X), layer_sizes=[10, 5, 2], seed=1, logginglevel=30)
    mu = scDEF(MiniAnnData(X), layer_sizes=[10, 5, 2], seed=1, logginglevel=30, x_placement="device_u16")
    eng_d, eng_u = md._get_engine(), mu._get_engine()
    assert eng_u.X is None and eng_u.X_u16.dtype == torch.uint16 and eng_u.X_u16.element_size() == 2
    idx = torch.as_tensor(np.random.RandomState(1).permutation(300)[:77], dtype=torch.int64, device="cuda")
    assert torch.equal(eng_u.gather_rows(idx), eng_d.X[idx])
    np.testing.assert_allclose(eng_u.x_lgamma.cpu().numpy(), eng_d.x_lgamma.cpu().numpy(), rtol=1e-6, atol=1e-6)
    for a, b in zip(mu.local_params + mu.global_params, md.local_params + md.global_params):
        np.testing.assert_array_equal(a, b)                                      # same constants -> same initialisation
    kw = dict(n_epoch=3, num_samples=5, batch_size=64)
    md._learn(**kw)
    mu._learn(**kw)
    np.testing.assert_allclose(mu.elbos[-1], md.elbos[-1], rtol=1e-6)
    for a, c in zip(mu.local_params + mu.global_params, md.local_params + md.global_params):
        np.testing.assert_allclose(a, c, rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        scDEF(MiniAnnData(X + 0.5), layer_sizes=[10, 5, 2], x_placement="device_u16")
