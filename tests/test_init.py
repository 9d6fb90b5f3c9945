"""Device cold start of the z layers (`scdef_init_gamma_block`).  The reference draws its Gamma variates
with JAX keys, so parity is distributional: the CPU tests draw from the SAME sampler code on the host
(`scdef_debug_gamma_host`, no CUDA call) and test it against scipy's Gamma distribution and the NumPy path of
`init_var_params`; the GPU test checks that the kernel writes what that code draws."""
import ctypes

import numpy as np
import pytest
import scipy.stats as st
import torch

from scdef_b200 import MiniAnnData, _lib, scDEF
from scdef_b200.model import _moment_match


def _host_gamma(shape, seed, stream_id, n):
    lib = ctypes.CDLL(_lib.lib_path())
    lib.scdef_debug_gamma_host.argtypes = [ctypes.c_float, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
    lib.scdef_debug_gamma_host.restype = None
    out = np.empty(n, dtype=np.float32)
    lib.scdef_debug_gamma_host(shape, seed, stream_id, n, out.ctypes.data)
    return out


@pytest.mark.parametrize("shape", [0.05, 0.5, 1.0, 2.5, 100.0])   # 0.05 / 1 / 100 are the z layers' concentrations
def test_sampler_is_gamma_distributed(shape):
    x = _host_gamma(shape, 42, 1, 200_000).astype(np.float64)
    assert x.mean() == pytest.approx(shape, rel=0.02) and x.var() == pytest.approx(shape, rel=0.04)
    tiny = 1e-30                                                    # float32 underflow of U^(1/a) for a << 1
    assert (x <= tiny).mean() == pytest.approx(st.gamma.cdf(tiny, shape), abs=2e-3)
    p0 = st.gamma.cdf(tiny, shape)
    ks = st.kstest(x[x > tiny], lambda t: (st.gamma.cdf(t, shape) - p0) / (1.0 - p0))
    assert ks.pvalue > 1e-3, ks
    # streams are independent and reproducible
    assert np.array_equal(_host_gamma(shape, 42, 1, 1000), x[:1000].astype(np.float32))
    assert not np.array_equal(_host_gamma(shape, 42, 2, 1000), x[:1000].astype(np.float32))
    assert not np.array_equal(_host_gamma(shape, 43, 1, 1000), x[:1000].astype(np.float32))


def test_initial_means_have_the_distribution_of_the_numpy_path():
    """What `init_var_params` does with the draws: m = clip(G / a, 1e-3, 4), compared with NumPy's Generator (the
    package's host path, `model.py::init_var_params`) by a two-sample KS test, per z-layer concentration."""
    rng = np.random.default_rng(0)
    for a in (0.05, 1.0, 100.0):
        ours = np.clip(_host_gamma(a, 7, 0, 100_000).astype(np.float64) / a, 1e-3, 4.0)
        ref = np.clip(rng.standard_gamma(a, size=100_000) / a, 1e-3, 4.0)
        assert st.ks_2samp(ours, ref).pvalue > 1e-3
        assert (ours == 1e-3).mean() == pytest.approx((ref == 1e-3).mean(), abs=5e-3)


def test_device_init_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    X = np.random.default_rng(0).poisson(1.0, size=(40, 30)).astype(np.float32)
    m = scDEF(MiniAnnData(X), layer_sizes=[4, 2], seed=1, logginglevel=30)
    with pytest.raises(RuntimeError):
        m.init_var_params(device_init=True)


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")
def test_kernel_writes_the_moment_matched_draws():
    lib = _lib.load()
    N, K, a, lo, hi, div = 333, 7, 0.05, 1e-3, 4.0, 100.0
    out = torch.empty((2, N, K), dtype=torch.float32, device="cuda")
    row0 = 1000
    rc = lib.scdef_init_gamma_block(ctypes.c_void_p(out[0].data_ptr()), ctypes.c_void_p(out[1].data_ptr()), N, K, K, a, lo, hi, div,
                                    5, 3, row0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    got = out.cpu().numpy()
    g = _host_gamma(a, 5, 3, (row0 + N) * K)[row0 * K:].reshape(N, K).astype(np.float64)
    m = np.clip(g / a, lo, hi)
    mu, rho = _moment_match(m, m / div)
    # device and host libm differ in the last bits; a draw that sits on a rejection boundary may differ entirely
    close = np.isclose(got[0], mu, rtol=1e-4, atol=1e-5) & np.isclose(got[1], rho, rtol=1e-4, atol=1e-5)
    assert close.mean() > 0.999


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")
def test_device_init_of_the_model_trains():
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(200, 80, seed=0, lib=600.0)
    m = scDEF(MiniAnnData(X), layer_sizes=[8, 4, 1], seed=1, logginglevel=30)
    m.init_var_params(device_init=True)
    z = m.local_params[1]
    assert z.shape == (2, 200, 13) and np.all(np.isfinite(z))
    mean0 = np.exp(z[0][:, :8] + 0.5 * np.exp(z[1][:, :8]) ** 2)
    assert 1e-3 * 0.99 <= mean0.min() and mean0.max() <= 4.0 * 1.01
    m._learn(n_epoch=2, num_samples=4, batch_size=64, filter=False, annotate=False)
    assert np.isfinite(m.elbos[-1][-1])
