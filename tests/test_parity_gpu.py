"""GPU parity: the CUDA path (through the C ABI) vs the float64 oracle on the same inputs and
the same injected noise.  Tolerance: 1e-4 relative on the loss, 1e-4 on max|g-g_ref|/max|g_ref|
per gradient tensor (BASELINE.json north_star)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import elbo_closed
from oracle.elbo_ref import Noise
from tests.helpers import make_engine, noise_to_spec, rel_err, small_problem

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _run_case(N, G, Ks, nb, B, S, sg, scb=0.0, sgb=0.0, use_brd=True, marg=False, anneal=1.3,
              resident=True, lik_impl=0, seed=0, couple=False, mutate=None, **hyper):
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=seed, use_brd=use_brd,
                                            marginalize_alpha=marg, **hyper)
    if mutate is not None:
        mutate(model, X)
    eng = make_engine(X, batch, model, lp, gp, resident=resident)
    idx = np.random.RandomState(seed).permutation(N)[:B]
    noise = Noise.draw(S, B, nb, G, Ks, seed=seed + 7, dtype=torch.float32, couple_like_reference=couple)
    ref_loss, ref_gl, ref_gg = elbo_closed.loss_and_grads(
        model, X[idx], idx, lp, gp, noise, anneal, sg, scb, sgb, model.alpha)
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    Xb = None if resident else torch.as_tensor(X[idx], device="cuda").contiguous()
    spec = noise_to_spec(noise)
    kw = dict(num_samples=S, annealing=anneal, stop_gradients=sg, stop_cell_budgets=scb,
              stop_gene_budgets=sgb, alpha=model.alpha, X_batch=Xb, lik_impl=lik_impl)
    out = {}
    eng.elbo_grad("local", idx_d, spec, **kw)
    out["loss_local"] = eng.loss_value()
    gl = [g.cpu().numpy() for g in eng.local_grads(B)]
    eng.elbo_grad("global", idx_d, spec, **kw)
    out["loss_global"] = eng.loss_value()
    gg = [g.cpu().numpy() for g in eng.glob_grad.views]
    errs = {"loss_local": abs(out["loss_local"] - ref_loss) / abs(ref_loss),
            "loss_global": abs(out["loss_global"] - ref_loss) / abs(ref_loss)}
    for name, g, r in zip(("cell_scale", "z"), gl, ref_gl):
        r = r[:, idx]
        if np.abs(r).max() == 0:
            assert np.abs(g).max() == 0, name
        else:
            errs["g_" + name] = rel_err(g, r)
    names = ["gene_scale", "brd"] + [f"W{l}" for l in range(len(Ks))] + ["ard", "alpha"]
    for name, g, r in zip(names, gg, ref_gg):
        if np.abs(r).max() == 0:
            assert np.abs(g).max() == 0, name
        else:
            errs["g_" + name] = rel_err(g, r)
    bad = {k: v for k, v in errs.items() if not (v <= TOL)}
    assert not bad, f"parity failures {bad} (all: {errs})"
    return errs


def test_c1_shape_full_batch():
    # BASELINE config C1: 200 x 120, layers [10,5,2], full batch
    _run_case(200, 120, (10, 5, 2), 1, 200, 4, [0, 0, 0])


@pytest.mark.parametrize("sg", [[0, 0, 0], [1, 0, 0], [0, 0, 1], [1, 1, 0], [0, 1, 0]])
@pytest.mark.parametrize("use_brd", [True, False])
def test_stop_gradient_phases(sg, use_brd):
    _run_case(96, 70, (9, 5, 2), 1, 40, 3, sg, use_brd=use_brd)


@pytest.mark.parametrize("scb,sgb", [(1.0, 0.0), (0.0, 1.0), (1.0, 1.0)])
def test_budget_freezes(scb, sgb):
    _run_case(96, 70, (9, 5, 2), 1, 40, 3, [0, 0, 0], scb=scb, sgb=sgb)


@pytest.mark.parametrize("nb", [2, 3, 11])
def test_batches(nb):
    _run_case(120, 75, (8, 4, 1), nb, 50, 3, [0, 0, 0])


def test_marginalize_alpha():
    _run_case(96, 70, (9, 5, 2), 1, 40, 3, [0, 0, 0], marg=True)
    _run_case(96, 70, (9, 5, 2), 2, 40, 3, [0, 0, 1], marg=True, use_brd=False)


def test_streamed_batch_equals_resident():
    _run_case(96, 70, (9, 5, 2), 1, 40, 3, [0, 0, 0], resident=False)


def test_reference_noise_coupling():
    _run_case(96, 70, (9, 5, 1), 1, 40, 3, [0, 0, 0], couple=True)


def test_ragged_shapes_and_single_layer():
    _run_case(70, 131, (13,), 1, 33, 2, [0])            # L = 1
    _run_case(70, 67, (17, 3), 1, 1, 2, [0, 0])          # B = 1
    _run_case(150, 257, (33, 16, 6, 3, 1), 2, 65, 2, [0, 0, 0, 0, 0])
    _run_case(300, 190, (100, 60, 30, 10), 1, 129, 2, [0, 0, 0, 0])  # the C5 ladder


def test_default_ladder_and_s1():
    _run_case(130, 90, (20, 10, 5, 1), 1, 64, 1, [0, 0, 0, 0])


@pytest.mark.parametrize("lik_impl", [1, 2])
def test_both_likelihood_kernels_on_tc_eligible_shapes(lik_impl):
    """SIMT (fp32 CUDA cores) and tcgen05 (split-bf16 tensor cores) agree with the oracle."""
    _run_case(200, 120, (10, 5, 2), 1, 200, 3, [0, 0, 0], lik_impl=lik_impl)
    _run_case(300, 200, (100, 60, 30, 10), 1, 257, 2, [0, 0, 0, 0], lik_impl=lik_impl)       # 2 cell blocks
    _run_case(160, 72, (24, 6, 1), 3, 130, 2, [0, 0, 0], lik_impl=lik_impl)                   # batches
    _run_case(100, 64, (12, 4), 1, 50, 2, [1, 0], lik_impl=lik_impl)                          # likelihood off
    _run_case(100, 64, (12, 4), 1, 50, 2, [0, 1], use_brd=False, lik_impl=lik_impl)
    _run_case(600, 136, (112, 3), 2, 520, 1, [0, 0], lik_impl=lik_impl, resident=False)       # K0 = 112, 3 cell blocks
    _run_case(96, 64, (9, 5, 2), 1, 40, 3, [0, 0, 0], marg=True, couple=True, lik_impl=lik_impl)


def test_tc_request_on_unsupported_shape_fails_loudly():
    from scdef_b200.engine import ScdefError

    with pytest.raises(ScdefError, match="tcgen05"):
        _run_case(96, 70, (9, 5, 2), 1, 40, 3, [0, 0, 0], lik_impl=2)   # G % 8 != 0


def test_sscdef_pinned_layer():
    """a8: supervised top layer (`_sscdef.py:484-531`): constant z = onehot(label) v 1e-3, no entropy for it,
    alpha floor 1.0, zero gradient for its variational parameters."""
    def pin(model, X):
        rng = np.random.default_rng(3)
        lab = rng.integers(0, 3, size=X.shape[0])
        fz = np.full((X.shape[0], 3), 1e-3)
        fz[np.arange(X.shape[0]), lab] = 1.0
        model.supervised_layer = 2
        model.fixed_top_z = fz
        model.alpha_floor = 1.0
    errs = _run_case(110, 64, (9, 5, 3), 1, 48, 3, [0, 0, 1], mutate=pin)
    _run_case(110, 70, (9, 5, 3), 2, 48, 2, [0, 0, 1], mutate=pin)


@pytest.mark.parametrize("lik_impl", [1, 2])
def test_iscdef_matrix_valued_w0_priors(lik_impl):
    """iscDEF feeds (K0,G) matrices as w_priors[0] (`_iscdef.py:442-450`) and block priors for l>=1; both likelihood
    paths take them (the tensor-core path folds the per-element prior gradient in its W epilogue)."""
    def markers(model, X):
        rng = np.random.default_rng(5)
        K0, G = model.layer_sizes[0], X.shape[1]
        strengths = np.where(rng.random((K0, G)) < 0.15, 20.0, 1.0)
        gene_sets = np.where(strengths > 1.0, 1.0, 0.05)
        model.w_priors[0] = [strengths, strengths / gene_sets]
        conn = np.where(rng.random(np.asarray(model.w_priors[1][0]).shape) < 0.5, 1.0, 1e-2)
        model.w_priors[1] = [np.asarray(model.w_priors[1][0]) * 5.0, np.asarray(model.w_priors[1][1]) * 5.0 / conn]
    _run_case(90, 64, (12, 4, 1), 1, 40, 3, [0, 0, 0], mutate=markers, lik_impl=lik_impl)     # BRD: A = P/tau per element
    _run_case(90, 72, (12, 4, 1), 1, 40, 2, [0, 0, 0], use_brd=False, mutate=markers, lik_impl=lik_impl)
    _run_case(300, 136, (24, 6, 1), 2, 200, 2, [0, 0, 0], mutate=markers, lik_impl=lik_impl)  # 3 gene tiles, 2 cell units
    _run_case(90, 64, (12, 4, 1), 1, 40, 2, [1, 0, 0], mutate=markers, lik_impl=lik_impl)     # layer 0 frozen: value + ard only
    if lik_impl == 1:
        _run_case(90, 70, (12, 4, 1), 1, 40, 2, [0, 0, 0], use_brd=False, mutate=markers)     # odd G: auto -> CUDA cores


@pytest.mark.parametrize("nb,S", [(1, 100), (8, 10)])
def test_full_step_shape_tensor_core_path_equals_cuda_core_path(nb, S):
    """BASELINE full per-step shapes (C5: 5000 genes, layers [100,60,30,10], 256 cells, 100 MC samples; C4: 8 batches):
    the oracle cannot run these in seconds, so the size-independent property is that the two independent likelihood
    implementations (tcgen05 split-bf16 vs fp32 CUDA cores), fed the same Philox draws, agree on the loss and on every
    gradient tensor.  (The kernels' work does not depend on the number of cells beyond the minibatch.)"""
    from scdef_b200.engine import NoiseSpec

    N, G, Ks, B = 1536, 5000, (100, 60, 30, 10), 256
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=11)
    eng = make_engine(X, batch, model, lp, gp)
    idx = torch.as_tensor(np.random.RandomState(5).permutation(N)[:B], dtype=torch.int64, device="cuda")
    spec = NoiseSpec.philox(seed=3, step=17, coupling=True)
    kw = dict(num_samples=S, annealing=1.2, stop_gradients=[0, 0, 0, 0], alpha=model.alpha)
    res = {}
    for impl in (1, 2):
        eng.elbo_grad("local", idx, spec, lik_impl=impl, **kw)
        loss_l = eng.loss_value()
        gl = [g.cpu().numpy().copy() for g in eng.local_grads(B)]
        eng.elbo_grad("global", idx, spec, lik_impl=impl, **kw)
        res[impl] = (loss_l, eng.loss_value(), gl, [g.cpu().numpy().copy() for g in eng.glob_grad.views])
    a, b = res[1], res[2]
    assert abs(a[0] - b[0]) / abs(a[0]) < 1e-5 and abs(a[1] - b[1]) / abs(a[1]) < 1e-5
    assert abs(a[0] - a[1]) / abs(a[0]) < 1e-6          # both passes evaluate the same objective
    for x, y in zip(a[2] + a[3], b[2] + b[3]):
        if np.abs(x).max() > 0:
            assert rel_err(y, x) < TOL
        else:
            assert np.abs(y).max() == 0


def test_rank_without_minibatch_rows_scalar_and_matrix_priors():
    """Data parallel, static ownership: a rank may own no member of a global minibatch (B = 0).  Its global pass still
    carries the cell-independent terms (priors / entropies x global_weight); both likelihood paths must agree on them,
    with scalar and with matrix-valued W_0 priors (the latter folds the prior's W gradient outside the main kernel)."""
    from scdef_b200.engine import NoiseSpec

    for matrix in (False, True):
        X, batch, model, lp, gp = small_problem(90, 64, (12, 4, 1), 1, seed=2)
        if matrix:
            rng = np.random.default_rng(5)
            strengths = np.where(rng.random((12, 64)) < 0.15, 20.0, 1.0)
            model.w_priors[0] = [strengths, strengths / np.where(strengths > 1.0, 1.0, 0.05)]
        eng = make_engine(X, batch, model, lp, gp)
        idx = torch.empty(0, dtype=torch.int64, device="cuda")
        spec = NoiseSpec.philox(seed=3, step=5, coupling=True)
        kw = dict(num_samples=3, annealing=1.1, stop_gradients=[0, 0, 0], alpha=model.alpha, scale=90 / 40, global_weight=0.5)
        res = {}
        for impl in (1, 2):
            eng.elbo_grad("global", idx, spec, lik_impl=impl, **kw)
            res[impl] = (eng.loss_value(), [g.cpu().numpy().copy() for g in eng.glob_grad.views])
        assert abs(res[1][0] - res[2][0]) / abs(res[1][0]) < 1e-5
        for x, y in zip(res[1][1], res[2][1]):
            if np.abs(x).max() > 0:
                assert rel_err(y, x) < TOL
            else:
                assert np.abs(y).max() == 0


@pytest.mark.parametrize("lik_impl", [1, 2])
def test_skip_loss_and_reuse_flags_do_not_change_the_gradients(lik_impl):
    """SCDEF_FLAG_SKIP_LOSS (LOCAL pass whose loss the caller discards) and SCDEF_FLAG_REUSE_W0_SAMPLES (GLOBAL pass right
    after a LOCAL pass with the same noise) are pure work-saving hints: gradients and the global-pass loss are unchanged."""
    from scdef_b200.engine import NoiseSpec

    X, batch, model, lp, gp = small_problem(300, 136, (24, 6, 1), 2, seed=6)
    eng = make_engine(X, batch, model, lp, gp)
    idx = torch.as_tensor(np.random.RandomState(2).permutation(300)[:200], dtype=torch.int64, device="cuda")
    spec = NoiseSpec.philox(seed=9, step=4, coupling=True)
    kw = dict(num_samples=3, annealing=1.2, stop_gradients=[0, 0, 0], alpha=model.alpha, lik_impl=lik_impl)
    eng.elbo_grad("local", idx, spec, **kw)
    ref_l = [g.clone() for g in eng.local_grads(200)]
    eng.elbo_grad("global", idx, spec, **kw)
    ref_loss, ref_g = eng.loss_value(), eng.glob_grad.flat.clone()
    eng.elbo_grad("local", idx, spec, want_loss=False, **kw)
    for a, b in zip(ref_l, eng.local_grads(200)):
        assert rel_err(b.cpu().numpy(), a.cpu().numpy()) < 1e-6
    eng.elbo_grad("global", idx, spec, reuse_w0=True, **kw)
    assert abs(eng.loss_value() - ref_loss) / abs(ref_loss) < 1e-6
    assert rel_err(eng.glob_grad.flat.cpu().numpy(), ref_g.cpu().numpy()) < 1e-6


def test_lik_first_order_of_the_global_pass_gives_the_same_gradients():
    """`SCDEF_FLAG_LIK_FIRST` (the data-parallel loop sets it so that the W_0 gradient is final before the hierarchy kernel
    runs) only reorders kernels with disjoint outputs: the gradients are the same (bit-identical where no float atomics
    are involved), the loss agrees to the order of its float64 atomics, and the W_0 event fires."""
    X, batch, model, lp, gp = small_problem(300, 136, (24, 6, 1), 2, seed=3)
    idx = np.random.RandomState(1).permutation(300)[:200]
    noise = Noise.draw(3, 200, 2, 136, (24, 6, 1), seed=4, dtype=torch.float32)
    eng = make_engine(X, batch, model, lp, gp)
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    kw = dict(num_samples=3, annealing=1.0, stop_gradients=[0, 0, 0], alpha=model.alpha)
    ev = eng.w0_grad_event()
    out = []
    for lik_first in (False, True):
        eng.elbo_grad("global", idx_d, noise_to_spec(noise), lik_first=lik_first, **kw)
        ev.synchronize()
        torch.cuda.synchronize()
        out.append((eng.loss_value(), [g.clone() for g in eng.glob_grad.views]))
    assert abs(out[0][0] - out[1][0]) <= 1e-8 * abs(out[0][0])   # block partial sums reach the float64 total in a different order
    for i, (a, b) in enumerate(zip(out[0][1], out[1][1])):
        if i in (0, 2):   # gene scale and W_0: no atomics on their path
            assert torch.equal(a, b), i
        else:             # the hierarchy kernel adds its W_l / alpha contributions with float atomics
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(a.abs().max())), i
    split = eng.glob_grad.split
    assert split is not None and eng.glob_grad.flat[split:].numel() == 2 * 24 * 136       # [small blocks | W_0]
    assert eng.glob_grad.views[2].data_ptr() == eng.glob_grad.flat[split:].data_ptr()
