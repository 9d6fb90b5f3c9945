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


def _signature_model(G, Ks, kept, technical, seed):
    """A stand-in with exactly the attributes `tl.set_confident_signatures` reads (factor.py:441-552)."""
    import types

    import pandas as pd

    rng = np.random.default_rng(seed)
    L = len(Ks)
    dims = [G] + list(Ks)
    gp = [None, None] + [np.stack([rng.normal(-1.0, 1.0, (Ks[l], dims[l])), rng.normal(-1.5, 0.3, (Ks[l], dims[l]))]).astype(np.float32)
                         for l in range(L)]
    names = [f"L{l}" for l in range(L)]
    m = types.SimpleNamespace(n_layers=L, layer_names=names, global_params=gp, _fit_revision=3,
                              factor_lists=[np.asarray(k) for k in kept], adata=MiniAnnData(np.zeros((4, G), dtype=np.float32)))
    m.factor_names = [[f"L{l}_{k}" for k in kept[l]] for l in range(L)]
    m.pmeans = {f"{names[l]}W": np.exp(gp[2 + l][0].astype(float) + 0.5 * np.exp(2.0 * gp[2 + l][1].astype(float))) for l in range(L)}
    m.pvars = {f"{names[l]}W": (np.exp(np.exp(2.0 * gp[2 + l][1].astype(float))) - 1.0)
               * np.exp(2.0 * gp[2 + l][0].astype(float) + np.exp(2.0 * gp[2 + l][1].astype(float))) for l in range(L)}
    allnames = [n for l in m.factor_names for n in l]
    m.adata.uns["factor_obs"] = pd.DataFrame({"technical": [n in technical for n in allnames]}, index=allnames)
    return m


@pytest.mark.parametrize("G,Ks,kept,technical,S", [
    (200, (6, 4, 2), ([0, 1, 3, 4, 5], [0, 2, 3], [0, 1]), ("L0_3",), 24),
    (333, (40, 12, 3, 1), (list(range(0, 40, 1)), list(range(12)), [0, 1, 2], [0]), (), 40),      # ragged gene tile, two factor chunks
])
def test_confident_signatures_match_the_oracle(G, Ks, kept, technical, S):
    """`tl.set_confident_signatures` (device Monte-Carlo gene scores, `scdef_signature_scores`) against
    `oracle/tools_ref.py` (pinned to the reference's execution, tests/test_ref_pin.py) on the same injected draws."""
    thr, tq = 0.6, 0.9
    m = _signature_model(G, Ks, kept, technical, seed=G)
    L = len(Ks)
    fl = m.factor_lists
    l0_keep = np.array([i for i, n in enumerate(m.factor_names[0]) if n not in technical])
    gcpu = torch.Generator().manual_seed(5)
    eps = {"w0": torch.randn((S, len(l0_keep), G), generator=gcpu), "l0": torch.randn((S, len(fl[0]), G), generator=gcpu),
           "upper": {l: torch.randn((S, len(fl[l]), len(fl[l - 1])), generator=gcpu) for l in range(1, L)}}
    dev_eps = {"w0": eps["w0"].cuda(), "l0": eps["l0"].cuda(), "upper": {l: e.cuda() for l, e in eps["upper"].items()}}
    tools.set_confident_signatures(m, confidence_threshold=thr, tau_quantile=tq, mc_samples=S, random_seed=0, _eps=dev_eps)
    cache = m.adata.uns["confident_signatures"]
    assert cache["fit_revision"] == 3 and cache["params"]["mc_samples"] == S

    tr = tools_ref
    W_mu = [m.global_params[2 + l][0].astype(float) for l in range(L)]
    W_rho = [m.global_params[2 + l][1].astype(float) for l in range(L)]
    pm = [m.pmeans[f"L{l}W"] for l in range(L)]
    upper = {l: [] for l in range(1, L)}
    for s in range(S):
        # the oracle's W_0 draw covers all kept L0 rows and drops the technical ones afterwards (factor.py:243-245)
        e0 = np.zeros((len(fl[0]), G))
        e0[l0_keep] = eps["w0"][s].numpy()
        sc = tr.hierarchy_scores_draw(W_mu, W_rho, fl, l0_keep, [e0] + [eps["upper"][l][s].numpy() for l in range(1, L)])
        for l in sc:
            upper[l].append(sc[l])
    genes = {g: i for i, g in enumerate(np.asarray(m.adata.var_names))}
    n_nonempty = 0
    for l in range(L):
        means = tr.layer_term_means(pm, fl, l0_keep, l)
        if l == 0:
            sig = tr.layer0_confident(means, m.pvars["L0W"][fl[0]], thr, tq)
            draws0 = tr.lognormal_draw(W_mu[0][fl[0]][None], W_rho[0][fl[0]][None], eps["l0"].numpy())   # (S, F0, G)
            scores_of = lambda f: draws0[:, f, :]
        else:
            mc = np.stack(upper[l])
            sig = tr.confident_from_scores(means, mc, thr, tq)
            scores_of = lambda f, mc=mc: mc[:, f, :]
        combined = [tr.confidence_mean_score(c, means[f][idx]) for f, (idx, c) in enumerate(sig)]
        jac = tr.jaccard_from_scores([idx for idx, _ in sig], combined, scores_of)
        got = cache["by_layer"][str(l)]
        for f, name in enumerate(m.factor_names[l]):
            assert [genes[g] for g in got["signatures"][name]] == list(sig[f][0]), (l, name)
            assert np.allclose(got["confidences"][name], sig[f][1], atol=1e-9), (l, name)
            assert np.allclose(got["combined_scores"][name], combined[f], rtol=1e-6, atol=1e-9), (l, name)
            assert np.isclose(got["signature_confidences"][name], jac[f], atol=1e-6, equal_nan=True), (l, name, jac[f])
            n_nonempty += len(sig[f][0]) > 0
    assert n_nonempty >= L
    sigs, confs, jacs = tools.get_stored_confident_signatures(m, layer_idx=1, max_genes=3, return_confidences=True,
                                                              return_signature_confidences=True)
    assert all(len(v) <= 3 for v in sigs.values()) and set(jacs) == set(m.factor_names[1])
    m._fit_revision = 4
    with pytest.raises(KeyError):
        tools.get_stored_confident_signatures(m, layer_idx=0)


def test_confident_signatures_on_a_trained_model():
    """The whole call on a model trained by this package: every factor of every layer gets an entry, lists are ordered by
    the combined score, confidences respect the threshold, and a second call with the same seed reproduces the cache."""
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(150, 72, seed=0, lib=600.0)
    m = scDEF(MiniAnnData(X), layer_sizes=[8, 4, 1], seed=1, logginglevel=30)
    m.fit(n_epoch=3, num_samples=4, batch_size=64, filter=False, annotate=False)
    tools.set_confident_signatures(m, confidence_threshold=0.5, tau_quantile=0.8, mc_samples=30, random_seed=4)
    cache = m.adata.uns["confident_signatures"]
    first = {l: dict(v["signature_confidences"]) for l, v in cache["by_layer"].items()}
    for l in range(m.n_layers):
        layer = cache["by_layer"][str(l)]
        assert list(layer["signatures"]) == list(m.factor_names[l])
        for name in m.factor_names[l]:
            conf, comb = np.asarray(layer["confidences"][name]), np.asarray(layer["combined_scores"][name])
            assert len(conf) == len(comb) == len(layer["signatures"][name])
            assert np.all(conf >= 0.5) and np.all(np.diff(comb) <= 1e-12)
            j = layer["signature_confidences"][name]
            assert np.isnan(j) if len(conf) == 0 else 0.0 <= j <= 1.0
    assert set(m.adata.uns["factor_signatures"]) == {n for names in m.factor_names for n in names}
    tools.set_confident_signatures(m, confidence_threshold=0.5, tau_quantile=0.8, mc_samples=30, random_seed=4)
    second = {l: dict(v["signature_confidences"]) for l, v in m.adata.uns["confident_signatures"]["by_layer"].items()}
    for l in first:
        for name in first[l]:
            assert np.isclose(first[l][name], second[l][name], equal_nan=True)
    assert tools.get_stored_confident_signatures(m, layer_idx=1).keys() == cache["by_layer"]["1"]["signatures"].keys()
