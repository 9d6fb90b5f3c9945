"""`tl.assign_confident`: the oracle restatement, the host-side selection rule and the error
behaviour — CPU.  The device kernel is checked in tests/test_tools_gpu.py."""
import numpy as np
import pytest

from oracle import tools_ref
from scdef_b200 import tools


def _params(N, K, seed=0, sd=0.4):
    rng = np.random.default_rng(seed)
    mu = rng.normal(-1.0, 1.0, size=(N, K)).astype(np.float32)
    rho = np.log(sd) + 0.3 * rng.normal(size=(N, K)).astype(np.float32)
    return mu, rho.astype(np.float32)


def test_oracle_degenerate_posterior_is_the_gap_of_the_means():
    """sigma -> 0: every draw equals exp(mu), so the gap is constant: confidence = effect size = gap, SD = 0,
    the winner is certain (probability 1, entropy confidence 1)."""
    N, K, S = 7, 5, 11
    mu, _ = _params(N, K)
    rho = np.full((N, K), -30.0, dtype=np.float32)
    eps = np.random.default_rng(1).normal(size=(S, N, K)).astype(np.float32)
    out = tools_ref.layer_stats(mu, rho, eps, credible_level=0.9)
    z = np.exp(mu.astype(np.float64))
    zhat = z / z.sum(axis=1, keepdims=True)
    srt = np.sort(zhat, axis=1)
    gap = srt[:, -1] - srt[:, -2]
    np.testing.assert_allclose(out["confidence"], gap, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["effect_size"], gap, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["posterior_sd"], 0.0, atol=1e-6)
    np.testing.assert_allclose(out["winner_mass"], srt[:, -1], rtol=1e-5)
    np.testing.assert_array_equal(out["argmax"], zhat.argmax(axis=1))
    np.testing.assert_allclose(out["winner_probability"], 1.0)
    np.testing.assert_allclose(out["entropy_confidence"], 1.0)


def test_oracle_quantile_is_the_lower_tail_of_the_gap():
    """With two factors the gap at the winner is 2 zhat[w] - 1; the confidence is its (1 - level) quantile with
    linear interpolation, and a stricter credible level can only lower it."""
    N, S = 4, 41
    mu, rho = _params(N, 2, seed=3, sd=0.8)
    eps = np.random.default_rng(4).normal(size=(S, N, 2)).astype(np.float32)
    o90 = tools_ref.layer_stats(mu, rho, eps, credible_level=0.9)
    o99 = tools_ref.layer_stats(mu, rho, eps, credible_level=0.99)
    assert np.all(o99["confidence"] <= o90["confidence"] + 1e-12)
    z = np.exp(mu[None].astype(np.float32) + np.exp(rho)[None] * eps)
    zhat = z / z.sum(axis=-1, keepdims=True)
    w = zhat.mean(axis=0).argmax(axis=-1)
    gap = np.take_along_axis(zhat, w[None, :, None], axis=2)[..., 0] * 2.0 - 1.0          # (S, N)
    srt = np.sort(gap.astype(np.float64), axis=0)
    pos = 0.1 * (S - 1)
    lo = int(np.floor(pos))
    want = srt[lo] * (1 - (pos - lo)) + srt[lo + 1] * (pos - lo)
    np.testing.assert_allclose(o90["confidence"], np.clip(want, 0, 1), rtol=1e-5, atol=1e-6)


def test_oracle_single_factor_layer_is_constant():
    out = tools_ref.layer_stats(np.zeros((3, 1)), np.zeros((3, 1)), np.zeros((5, 3, 1), dtype=np.float32))
    assert np.all(out["confidence"] == 1.0) and np.all(out["posterior_sd"] == 0.0) and np.all(out["argmax"] == 0)


@pytest.mark.parametrize("kept_sizes", [[6, 3, 1], [6, 3, 2], [1, 1], [0, 4, 1], [5], [1], [0, 0]])
def test_selection_rule_matches_the_cell_by_cell_restatement(kept_sizes):
    rng = np.random.default_rng(len(kept_sizes) * 7 + sum(kept_sizes))
    N, L = 60, len(kept_sizes)
    conf = rng.uniform(size=(N, L))
    conf[:5] = 0.0                      # cells that clear tau nowhere
    conf[5:8, 0] = 0.3                  # exactly at the threshold: ">=" clears it
    eff, sd = rng.normal(size=(N, L)), rng.uniform(size=(N, L))
    amax = rng.integers(0, 3, size=(N, L))
    got = tools.select_layers(conf, eff, sd, amax, kept_sizes, tau=0.3)
    want = tools_ref.select_layers(conf, eff, sd, amax, kept_sizes, tau=0.3)
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)


def test_argument_errors_come_before_any_device_work():
    class M:
        pass

    with pytest.raises(ValueError):
        tools.assign_confident(M(), n_samples=0)
    with pytest.raises(ValueError):
        tools.assign_confident(M(), tau=1.5)
    with pytest.raises(ValueError):
        tools.assign_confident(M(), credible_level=1.0)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(tools.ScdefToolsError):
        tools.assign_confident(object())
    with pytest.raises(tools.ScdefToolsError):
        tools.layer_confidence_stats(torch.zeros(2, 3), torch.zeros(2, 3), [0, 1], 4, 0.1)


# ---- the device kernel's ALGORITHM, emulated lane by lane in float32 NumPy (scdef_b200/csrc/tools.cuh) --------------
# Not the CUDA code itself, but the same decomposition: lane l owns the quads l, l+32, ... of 4 consecutive kept
# factors; butterfly (xor) reductions for sums / maxima / arg-maxima with "first maximum" ties; the two order statistics
# of the quantile by rank counting with index tie-break; linear interpolation in float32.  It must agree with the
# line-by-line restatement of the reference, which pins the index arithmetic and the tie rules the kernel relies on.
def _butterfly(vals, op):
    v = list(vals)
    o = 16
    while o > 0:
        v = [op(v[i], v[i ^ o]) for i in range(32)]
        o >>= 1
    return v[0]


def _emulate_cell(mu, sigma, eps, q_level):
    F = np.float32
    S, K = eps.shape
    Q = (K + 3) // 4
    maxq = (Q + 31) // 32
    own = [[4 * (lane + 32 * i) + c for i in range(maxq) for c in range(4) if 4 * (lane + 32 * i) + c < K] for lane in range(32)]
    assert sorted(k for ks in own for k in ks) == list(range(K))          # every factor owned exactly once

    def amax(a, b):   # (value, index): larger value wins, lower index on ties; empty lanes carry (-inf, huge)
        return b if (b[0] > a[0] or (b[0] == a[0] and b[1] < a[1])) else a

    def sample(s):
        z = {k: F(np.exp(F(F(sigma[k]) * F(eps[s, k]) + F(mu[k])))) for k in range(K)}
        part = [F(sum((z[k] for k in own[lane]), F(0))) for lane in range(32)]
        denom = max(_butterfly(part, lambda a, b: F(a + b)), F(1e-30))
        return {k: F(z[k] / denom) for k in range(K)}

    sz = {k: F(0) for k in range(K)}
    cnt = {k: 0 for k in range(K)}
    for s in range(S):
        zh = sample(s)
        loc = []
        for lane in range(32):
            best = (F(-np.inf), 0x7FFFFFFF)
            for k in own[lane]:
                sz[k] = F(sz[k] + zh[k])
                if zh[k] > best[0]:
                    best = (zh[k], k)
            loc.append(best)
        cnt[_butterfly(loc, amax)[1]] += 1
    invS = F(1.0) / F(S)
    loc = []
    for lane in range(32):
        best = (F(-np.inf), 0x7FFFFFFF)
        for k in own[lane]:
            m = F(sz[k] * invS)
            if m > best[0]:
                best = (m, k)
        loc.append(best)
    winner = _butterfly(loc, amax)[1]
    p = {k: F(F(cnt[k]) * invS) for k in range(K)}
    H = F(-sum(float(p[k]) * np.log(min(max(float(p[k]), 1e-30), 1.0)) for k in range(K)))
    gaps = np.zeros(S, dtype=F)
    sum_gap, sum_wm = F(0), F(0)
    for s in range(S):
        zh = sample(s)
        wm = zh[winner]
        ru = max(zh[k] for k in range(K) if k != winner)
        gaps[s] = F(wm - ru)
        sum_gap, sum_wm = F(sum_gap + gaps[s]), F(sum_wm + wm)
    mean_gap = F(sum_gap * invS)
    var = F(np.sum((gaps - mean_gap).astype(F) ** 2, dtype=F) * invS)
    pos = F(F(q_level) * F(S - 1))
    lo, hi = int(np.floor(pos)), min(int(np.ceil(pos)), S - 1)
    wh = F(pos - F(lo))
    rank = [sum(1 for j in range(S) if gaps[j] < gaps[i] or (gaps[j] == gaps[i] and j < i)) for i in range(S)]
    assert sorted(rank) == list(range(S))                                  # a permutation, ties included
    vlo, vhi = gaps[rank.index(lo)], gaps[rank.index(hi)]
    conf = F(vlo * (F(1) - wh) + vhi * wh)
    clip = lambda x, a, b: float(min(max(float(x), a), b))                 # noqa: E731
    return dict(effect_size=clip(mean_gap, -1, 1), posterior_sd=clip(np.sqrt(var), 0, 1), confidence=clip(conf, 0, 1),
                winner_mass=clip(sum_wm * invS, 0, 1), winner_probability=clip(max(p.values()), 0, 1),
                entropy_confidence=clip(1.0 - float(H) / np.log(K), 0, 1), argmax=winner)


@pytest.mark.parametrize("N,K,S,level", [(6, 5, 33, 0.9), (3, 137, 20, 0.9), (4, 3, 64, 0.95), (5, 2, 1, 0.5), (4, 6, 40, 0.9)])
def test_kernel_algorithm_emulated_lane_by_lane_matches_the_oracle(N, K, S, level):
    rng = np.random.default_rng(N * 100 + K + S)
    mu, rho = _params(N, K, seed=K)
    eps = rng.normal(size=(S, N, K)).astype(np.float32)
    if K == 6:                      # exact ties: duplicated factors and a duplicated draw
        mu[:, 3], rho[:, 3], eps[:, :, 3] = mu[:, 1], rho[:, 1], eps[:, :, 1]
        eps[7] = eps[5]
    want = tools_ref.layer_stats(mu, rho, eps, credible_level=level)
    for n in range(N):
        got = _emulate_cell(mu[n], np.exp(rho[n]).astype(np.float32), eps[:, n, :], 1.0 - level)
        assert got["argmax"] == want["argmax"][n]
        for name in tools.STAT_NAMES:
            assert got[name] == pytest.approx(float(want[name][n]), abs=2e-5), (name, n)


# ---- set_confident_signatures: the host-side algebra around `scdef_signature_scores` (no device needed) --------------
def test_rank_counts_reproduce_the_top_k_jaccard_of_the_oracle():
    """The device reports, per draw and signature gene, how many genes score above it; `_jaccard_from_ranks` turns that
    into the weighted Jaccard of the draw's top-k list (factor.py:41-70, 338-362).  Checked here with NumPy rank counts
    against the oracle's argsort-based restatement (itself pinned to the executed reference)."""
    import torch

    rng = np.random.default_rng(3)
    S, F, G = 12, 5, 90
    scores = rng.lognormal(size=(S, F, G))
    sig = [np.sort(rng.choice(G, size=k, replace=False))[::-1].copy() for k in (7, 1, 0, 20, 3)]
    comb = [rng.uniform(0.1, 2.0, size=len(s)) for s in sig]
    comb[3][:] = 0.0                                      # all-zero weights fall back to uniform (factor.py:57-58)
    kmax = max(len(s) for s in sig)
    rank = np.zeros((S, F, kmax), dtype=np.int32)
    for s in range(S):
        for f, ref in enumerate(sig):
            for i, g in enumerate(ref):
                rank[s, f, i] = int(np.sum(scores[s, f] > scores[s, f, g]))
    got = tools._jaccard_from_ranks(sig, comb, torch.as_tensor(rank))
    want = tools_ref.jaccard_from_scores(sig, comb, lambda f: scores[:, f, :])
    assert np.isnan(got[2]) and np.isnan(want[2])
    for f in (0, 1, 3, 4):
        assert abs(got[f] - want[f]) < 1e-12, f


def test_signature_helpers_match_the_oracle_restatement():
    rng = np.random.default_rng(4)
    mu, conf = rng.lognormal(size=60), rng.uniform(0, 1, size=60)
    for min_effect in (None, 1.0):
        assert np.array_equal(tools._rank_signature(mu, conf, 0.4, min_effect), tools_ref.rank_signature(mu, conf, 0.4, min_effect))
    assert np.allclose(tools._confidence_mean_score(conf, mu), tools_ref.confidence_mean_score(conf, mu), rtol=0, atol=0)
    with pytest.raises(ValueError):
        tools.set_confident_signatures(object(), confidence_threshold=1.5)
    with pytest.raises(ValueError):
        tools.set_confident_signatures(object(), mc_samples=0)
