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
