"""`scdef_assign_confident` on the device vs the NumPy restatement of the reference (`oracle/tools_ref.py`), same
injected draws.  Tolerances: statistics 2e-5 absolute (they live in [0, 1]; float32 sums in a different order),
winner slots equal."""
import numpy as np
import pytest
import torch

from oracle import tools_ref
from scdef_b200 import MiniAnnData, scDEF, tools

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]
ATOL = 2e-5


def _case(N, Kt, cols, S, level, seed):
    rng = np.random.default_rng(seed)
    mu = rng.normal(-1.0, 1.0, size=(N, Kt)).astype(np.float32)
    rho = (np.log(0.5) + 0.4 * rng.normal(size=(N, Kt))).astype(np.float32)
    eps = rng.normal(size=(S, N, len(cols))).astype(np.float32)
    want = tools_ref.layer_stats(mu[:, cols], rho[:, cols], eps, credible_level=level)
    stats, amax = tools.layer_confidence_stats(torch.as_tensor(mu, device="cuda"), torch.as_tensor(rho, device="cuda"), cols, S,
                                               1.0 - level, eps=torch.as_tensor(eps, device="cuda"))
    return want, stats.cpu().numpy(), amax.cpu().numpy()


@pytest.mark.parametrize("N,Kt,cols,S,level", [
    (37, 9, [0, 1, 2, 3, 4], 33, 0.9),                  # one quad and a ragged one
    (50, 140, list(range(3, 140, 1)), 64, 0.9),         # 137 factors: two quads per lane
    (20, 300, list(range(0, 300, 1)), 20, 0.95),        # 300 factors: the 8-quad instantiation
    (64, 12, [11, 2, 7], 500, 0.9),                     # the reference's default n_samples, gathered columns
    (9, 6, [4, 5], 1, 0.5),                             # a single draw: every statistic is that draw's
])
def test_layer_statistics_match_the_oracle(N, Kt, cols, S, level):
    want, stats, amax = _case(N, Kt, cols, S, level, seed=N + S)
    np.testing.assert_array_equal(amax, want["argmax"])
    for c, name in enumerate(tools.STAT_NAMES):
        np.testing.assert_allclose(stats[:, c], want[name], atol=ATOL, rtol=0, err_msg=name)


def test_philox_draws_do_not_depend_on_the_sharding():
    """Rows [a, b) of a full run equal a run over those rows alone with row0 = a (data-parallel ranks)."""
    rng = np.random.default_rng(5)
    N, K, S = 96, 10, 50
    mu = torch.as_tensor(rng.normal(size=(N, K)).astype(np.float32), device="cuda")
    rho = torch.as_tensor((np.log(0.5) + 0.2 * rng.normal(size=(N, K))).astype(np.float32), device="cuda")
    cols = list(range(K))
    full, amax = tools.layer_confidence_stats(mu, rho, cols, S, 0.1, seed=3, stream_id=2)
    part, amax_p = tools.layer_confidence_stats(mu[40:70].contiguous(), rho[40:70].contiguous(), cols, S, 0.1, seed=3, stream_id=2, row0=40)
    assert torch.equal(full[40:70], part) and torch.equal(amax[40:70], amax_p)
    other, _ = tools.layer_confidence_stats(mu, rho, cols, S, 0.1, seed=4, stream_id=2)
    assert not torch.equal(full, other)


def test_philox_statistics_agree_with_injected_normal_draws():
    """Distribution-level check of the Philox path: with many samples its statistics are within Monte-Carlo error of
    the oracle's run on independent NumPy normals."""
    rng = np.random.default_rng(11)
    N, K, S = 40, 6, 4000
    mu = rng.normal(size=(N, K)).astype(np.float32)
    rho = (np.log(0.4) + 0.1 * rng.normal(size=(N, K))).astype(np.float32)
    want = tools_ref.layer_stats(mu, rho, rng.normal(size=(S, N, K)).astype(np.float32), credible_level=0.9)
    stats, _ = tools.layer_confidence_stats(torch.as_tensor(mu, device="cuda"), torch.as_tensor(rho, device="cuda"), list(range(K)), S, 0.1, seed=1)
    stats = stats.cpu().numpy()
    np.testing.assert_allclose(stats[:, 0], want["effect_size"], atol=0.03)
    np.testing.assert_allclose(stats[:, 3], want["winner_mass"], atol=0.03)
    np.testing.assert_allclose(stats[:, 1], want["posterior_sd"], atol=0.03)


def test_assign_confident_writes_the_reference_fields():
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(150, 60, seed=0, lib=600.0)
    m = scDEF(MiniAnnData(X), layer_sizes=[8, 4, 1], seed=1, logginglevel=30)
    m.fit(n_epoch=3, num_samples=4, batch_size=64, filter=False, annotate=False)
    tools.assign_confident(m, n_samples=40, tau=0.2, credible_level=0.9, key_added="conf")
    ad = m.adata
    for name in tools.STAT_NAMES:
        assert ad.obsm[f"conf_{name}"].shape == (150, 3)
    assert ad.obsm["conf_argmax_factor"].shape == (150, 3)
    assert np.all(ad.obsm["conf_confidence"][:, 2] == 1.0)          # the single-factor root layer
    idx = np.asarray(ad.obs["conf_best_layer_idx"])
    assert set(np.unique(idx)) <= {0, 1, 2}
    conf = ad.obsm["conf_confidence"]
    for n in range(150):                                             # finest multi-factor layer that clears tau, else the root
        want = 0 if conf[n, 0] >= 0.2 else (1 if conf[n, 1] >= 0.2 else 2)
        assert idx[n] == want
    assert ad.uns["conf"]["n_samples"] == 40 and ad.uns["conf"]["selection_rule"] == "finest_layer_clearing_tau"
