"""Pins the restated oracle to the reference's OWN source.

The reference cannot be installed here (no jax / optax / tensorflow-probability), so
`oracle/ref_exec.py` imports its unmodified `.py` files from `/root/reference/src/scdef` on top of the
torch-backed stand-ins of `oracle/jaxshim.py` and RUNS them: the constructor (`load_adata`,
`update_model_priors`, `init_var_params`), `_get_or_build_learn_grad_fns` -> `objective` -> `batch_elbo`
-> `elbo` (`_scdef.py:1823-2181, 2808-2856`) with `jax_utils.py:6-46`, and the hot loop `_learn`
(`_scdef.py:2923-3392`: `data_stream`, `local_update`, `global_update`, `optax.adam`, `clip_params`).
`oracle/elbo_ref.py`, `oracle/elbo_closed.py` and `oracle/host_ref.py` must reproduce those executions
(float64, <= 1e-11) on the same counts, the same parameters and the very draws the reference consumed.

Skipped where `/root/reference` does not exist (the GPU box): there `tests/golden/ref_*.npz`, written by
`tests/golden/make_golden_ref.py` from the same executions, stand in (`test_golden_ref_*`).
"""
import os

import numpy as np
import pytest
import torch

from oracle import elbo_closed, elbo_ref, host_ref, ref_exec
from tests import ref_cases
from tests.helpers import rel_err

needs_reference = pytest.mark.skipif(not ref_exec.available(), reason="/root/reference is not on this machine")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-11


@pytest.fixture(scope="module")
def ref():
    with ref_exec.reference_runtime(torch.float64) as r:
        yield r


def _reference_eval(ref, m, case, X, lp, gp, key):
    idx = np.random.RandomState(0).permutation(case["N"])[: case["B"]]
    jnp = ref.jnp
    local_fn, global_fn = m._get_or_build_learn_grad_fns(case["S"])
    args = (jnp.array(X[idx]), jnp.array(idx), [jnp.array(p) for p in lp], [jnp.array(p) for p in gp], key,
            case.get("anneal", 1.0), jnp.array(np.asarray(case["sg"], dtype=np.float64)),
            case.get("stop_cell", 0.0), case.get("stop_gene", 0.0), m.alpha)
    loss_l, g_loc = local_fn(*args)
    loss_g, g_glob = global_fn(*args)
    return idx, float(loss_l), float(loss_g), ref.to_numpy(g_loc), ref.to_numpy(g_glob)


@needs_reference
def test_the_reference_source_is_what_runs(ref):
    path, lo, hi = ref.source_lines(ref.scDEF.elbo)
    assert path == "/root/reference/src/scdef/models/_scdef.py" and (lo, hi) == (1823, 2146)
    path, lo, hi = ref.source_lines(ref.scDEF.batch_elbo)
    assert path.endswith("models/_scdef.py") and (lo, hi) == (2148, 2181)
    path, lo, hi = ref.source_lines(ref.jax_utils.gamma_logpdf)
    assert path == "/root/reference/src/scdef/utils/jax_utils.py" and (lo, hi) == (40, 46)


@needs_reference
@pytest.mark.parametrize("name", sorted(ref_cases.ELBO_CASES))
def test_restated_elbo_and_gradients_equal_the_reference_execution(ref, name):
    case = ref_cases.ELBO_CASES[name]
    m, X, batch = ref_cases.build_reference_model(ref, case)
    lp = ref_cases.perturb(m.local_params, seed=11)
    gp = ref_cases.perturb(m.global_params, seed=12)
    key = ref.random.PRNGKey(7)
    idx, loss_l, loss_g, g_loc, g_glob = _reference_eval(ref, m, case, X, lp, gp, key)
    assert np.isfinite(loss_l) and loss_l == loss_g

    om = ref_cases.oracle_model_from_reference(m)
    noise = ref_cases.noise_from_key(ref.random, key, case["S"], len(idx), case["nb"], case["G"], case["Ks"])
    kw = (case.get("anneal", 1.0), case["sg"], case.get("stop_cell", 0.0), case.get("stop_gene", 0.0), om.alpha)
    # (1) the line-by-line autograd restatement
    l1, gl1 = elbo_ref.local_loss_grad(om, X[idx], idx, lp, gp, noise, *kw)
    l2, gg1 = elbo_ref.global_loss_grad(om, X[idx], idx, lp, gp, noise, *kw)
    assert abs(float(l1) - loss_l) <= TOL * abs(loss_l) and float(l1) == float(l2)
    for a, b in zip(gl1, g_loc):
        assert rel_err(a.numpy(), b) <= TOL
    for a, b in zip(gg1, g_glob):
        assert rel_err(a.numpy(), b) <= TOL
    # (2) the closed-form oracle (the arithmetic of the CUDA kernels)
    l3, gl3, gg3 = elbo_closed.loss_and_grads(om, X[idx], idx, lp, gp, noise, *kw)
    assert abs(l3 - loss_l) <= 1e-10 * abs(loss_l)
    for a, b in zip(list(gl3) + list(gg3), list(g_loc) + list(g_glob)):
        assert rel_err(a, b) <= 1e-9


@needs_reference
@pytest.mark.parametrize("name", ["base", "two_batches", "no_brd", "hyper"])
def test_restated_constants_equal_the_reference_constructor(ref, name):
    """`load_adata` (482-578) + `update_model_priors` (1246-1287) as executed vs `host_ref.build_model`."""
    case = ref_cases.ELBO_CASES[name]
    m, X, batch = ref_cases.build_reference_model(ref, case)
    kw = {k: v for k, v in ref_cases.ctor_kwargs(case).items() if k not in ("layer_sizes", "seed")}
    if "brd_strength" in kw:
        kw["brd"] = kw.pop("brd_strength")
    om = host_ref.build_model(X, case["Ks"], batch_ids=batch if case["nb"] > 1 else None, **kw)
    want = ref_cases.oracle_model_from_reference(m)
    for attr in ("batch_indices_onehot", "batch_lib_ratio", "gene_ratio", "cell_alpha_factor"):
        np.testing.assert_allclose(np.asarray(getattr(om, attr), dtype=float),
                                   np.asarray(getattr(want, attr), dtype=float), rtol=1e-12, err_msg=attr)
    assert om.alpha == pytest.approx(want.alpha, rel=1e-13)
    for (a0, b0), (a1, b1) in zip(om.w_priors, want.w_priors):
        np.testing.assert_allclose(np.asarray(a0, dtype=float), np.asarray(a1, dtype=float), rtol=1e-13)
        np.testing.assert_allclose(np.asarray(b0, dtype=float), np.asarray(b1, dtype=float), rtol=1e-13)
    for attr in ("use_brd", "marginalize_alpha", "top_alpha", "brd", "brd_mean", "shrinkage_shape",
                 "shrinkage_rate", "cell_scale_shape", "gene_scale_shape"):
        assert getattr(om, attr) == getattr(want, attr), attr
    # deterministic blocks of init_var_params (1491-1530, 1699-1732): c, s, omega, alpha
    lp0, gp0 = host_ref.init_var_params(om, seed=1, brd_mean=om.brd_mean)
    L = len(case["Ks"])
    np.testing.assert_allclose(lp0[0], np.asarray(m.local_params[0]), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(gp0[0], np.asarray(m.global_params[0]), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(gp0[2 + L], np.asarray(m.global_params[2 + L]), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(gp0[3 + L], np.asarray(m.global_params[3 + L]), rtol=2e-6, atol=2e-6)
    for a, b in zip(lp0 + gp0, list(m.local_params) + list(m.global_params)):
        assert tuple(a.shape) == tuple(b.shape)


def _run_reference_learn(ref, case, n_epoch, **learn_kw):
    m, X, batch = ref_cases.build_reference_model(ref, case)
    lp0, gp0 = ref.to_numpy(m.local_params), ref.to_numpy(m.global_params)
    m._learn(n_epoch=n_epoch, num_samples=case["S"], batch_size=case["B"], filter=False, annotate=False,
             min_epochs=n_epoch + 1, **learn_kw)
    return m, X, batch, lp0, gp0


def _replay_with_the_oracle(ref, m, case, X, lp0, gp0, n_epoch, lr=0.1, local_lr=0.01, sg=None, anneal=1.0,
                            freeze_w=False, stop_cell=0.0, stop_gene=0.0):
    """`_learn` (`_scdef.py:3181, 3215-3271`) restated with `host_ref`: key chain PRNGKey(seed) -> split per
    step, RandomState(0) minibatches, local pass + Adam, global pass (same key) + Adam + clip."""
    om = ref_cases.oracle_model_from_reference(m)
    N, B, S = case["N"], case["B"], case["S"]
    sg = [0.0] * len(case["Ks"]) if sg is None else sg
    lp, gp = [np.array(p) for p in lp0], [np.array(p) for p in gp0]
    lopt = host_ref.Adam(lp, local_lr, dtype=np.float64)
    gopt = host_ref.Adam(gp, lr, dtype=np.float64)
    stream = host_ref.data_stream_indices(N, B)
    n_batches = -(-N // B)
    rng = ref.random.PRNGKey(m.seed)
    epoch_losses = []
    for _ in range(n_epoch):
        losses = []
        for _ in range(n_batches):
            rng, rng_input = ref.random.split(rng)
            idx = next(stream)
            noise = ref_cases.noise_from_key(ref.random, rng_input, S, len(idx), case["nb"], case["G"], case["Ks"])
            loss, lp, gp = host_ref.learn_step(om, X[idx], idx, lp, gp, noise, lopt, gopt, anneal, sg, stop_cell,
                                               stop_gene, om.alpha, freeze_w=freeze_w)
            losses.append(loss)
        epoch_losses.append(np.mean(losses))
    return epoch_losses, lp, gp


@needs_reference
@pytest.mark.parametrize("name,learn_kw", [
    ("base", {}),                          # 96 cells, batches of 32
    ("marg_alpha", {}),                    # 96 cells in batches of 40: short last minibatch, q(alpha), 3 batches
    ("two_batches", {"freeze_w": True}),   # W restored after every global update (3153-3161)
])
def test_restated_hot_loop_equals_the_reference_learn(ref, name, learn_kw):
    case = dict(ref_cases.ELBO_CASES[name])
    n_epoch = 2
    m, X, batch, lp0, gp0 = _run_reference_learn(ref, case, n_epoch, **learn_kw)
    want_losses = np.asarray(m.elbos[-1], dtype=np.float64)
    assert len(want_losses) == n_epoch
    got_losses, lp, gp = _replay_with_the_oracle(ref, m, case, X, lp0, gp0, n_epoch, **learn_kw)
    np.testing.assert_allclose(got_losses, want_losses, rtol=1e-10)
    for a, b in zip(lp + gp, ref.to_numpy(m.local_params) + ref.to_numpy(m.global_params)):
        assert rel_err(a, b) <= 1e-9
    if learn_kw.get("freeze_w"):
        L = len(case["Ks"])
        for a, b in zip(gp0[2:2 + L], ref.to_numpy(m.global_params)[2:2 + L]):
            np.testing.assert_array_equal(a, b)


@needs_reference
def test_reference_minibatch_stream_is_randomstate0(ref):
    """`data_stream` (3167-3173) observed through the indices the reference hands to its gradient function."""
    case = dict(ref_cases.ELBO_CASES["marg_alpha"])
    m, X, batch = ref_cases.build_reference_model(ref, case)
    seen = []
    local_fn, global_fn = m._get_or_build_learn_grad_fns(case["S"])

    def spy(Xb, indices, *a):
        seen.append((np.asarray(indices).copy(), np.asarray(Xb).copy()))
        return local_fn(Xb, indices, *a)

    m._learn_jit_cache["local_loss_grad"] = spy
    m._learn(n_epoch=2, num_samples=case["S"], batch_size=case["B"], filter=False, annotate=False, min_epochs=3)
    stream = host_ref.data_stream_indices(case["N"], case["B"])
    assert len(seen) == 2 * 3
    for idx, Xb in seen:
        want = next(stream)
        np.testing.assert_array_equal(idx, want)
        np.testing.assert_array_equal(Xb, X[want])
    assert len(seen[2][0]) == 96 - 2 * 40  # short last slice


# ---- sscDEF (`_sscdef.py:256-585` elbo copy, `588-1030` loop) ---------------------------------------------
def _reference_sscdef(ref, add_root):
    case = dict(N=90, G=60, Ks=[8, 3, 1] if add_root else [8, 3], nb=1, B=30, S=2,
                sg=[0, 0, 1] if add_root else [0, 0])
    X, _ = ref_cases.make_counts(case)
    ad = ref.adata(X)
    ad.obs["celltype"] = [f"t{i % 3}" for i in range(case["N"])]
    m = ref.sscDEF(ad, "celltype", add_root=add_root, layer_sizes=list(case["Ks"]), seed=1)
    m.elbos, m.step_sizes = [], []
    return m, X, case


@needs_reference
@pytest.mark.parametrize("add_root", [True, False])
def test_restated_sscdef_elbo_equals_the_reference_execution(ref, add_root):
    m, X, case = _reference_sscdef(ref, add_root)
    t = int(m.supervised_top_layer_idx)
    assert t == (1 if add_root else 1)
    lp = ref_cases.perturb(m.local_params, seed=21)
    gp = ref_cases.perturb(m.global_params, seed=22)
    key = ref.random.PRNGKey(3)
    idx, loss_l, loss_g, g_loc, g_glob = _reference_eval(ref, m, case, X, lp, gp, key)
    om = ref_cases.oracle_model_from_reference(m, alpha_floor=1.0, supervised_layer=t,
                                               fixed_top_z=np.asarray(m._supervised_top_z, dtype=np.float64))
    noise = ref_cases.noise_from_key(ref.random, key, case["S"], len(idx), 1, case["G"], case["Ks"])
    kw = (1.0, case["sg"], 0.0, 0.0, om.alpha)
    l1, gl1 = elbo_ref.local_loss_grad(om, X[idx], idx, lp, gp, noise, *kw)
    _, gg1 = elbo_ref.global_loss_grad(om, X[idx], idx, lp, gp, noise, *kw)
    assert abs(float(l1) - loss_l) <= TOL * abs(loss_l)
    for a, b in zip(list(gl1) + list(gg1), list(g_loc) + list(g_glob)):
        assert rel_err(a.numpy(), b) <= TOL
    l3, gl3, gg3 = elbo_closed.loss_and_grads(om, X[idx], idx, lp, gp, noise, *kw)
    assert abs(l3 - loss_l) <= 1e-10 * abs(loss_l)
    for a, b in zip(list(gl3) + list(gg3), list(g_loc) + list(g_glob)):
        assert rel_err(a, b) <= 1e-9
    # the pinned layer gets no gradient at all
    s, e = sum(case["Ks"][:t]), sum(case["Ks"][:t + 1])
    assert np.abs(g_loc[1][:, :, s:e]).max() == 0.0


@needs_reference
def test_reference_sscdef_loop_discards_every_local_update(ref):
    """`_sscdef.py:892-893`: after `local_update` the loop re-reads `self.local_params` (pinned, but never
    assigned the updated values), so the per-cell blocks stay at their initial values for the whole `_learn`
    call while the globals train — the literal behaviour `scdef_b200.sscDEF` reproduces by default."""
    m, X, case = _reference_sscdef(ref, add_root=True)
    lp0, gp0 = ref.to_numpy(m.local_params), ref.to_numpy(m.global_params)
    m._learn(n_epoch=2, num_samples=2, batch_size=30, filter=False, annotate=False, min_epochs=3,
             optimize_layers=[0, 1])
    lp1, gp1 = ref.to_numpy(m.local_params), ref.to_numpy(m.global_params)
    for a, b in zip(lp0, lp1):
        np.testing.assert_array_equal(a, b)
    assert np.abs(gp1[2] - gp0[2]).max() > 1e-3


# ---- iscDEF (`_iscdef.py:167-450`): matrix-valued W priors into the same `elbo` --------------------------------
def _reference_iscdef(ref, markers_layer, use_brd=True):
    case = dict(N=90, G=60, nb=1, B=30, S=2)
    X, _ = ref_cases.make_counts(case)
    ad = ref.adata(X)
    markers = {"A": ["g1", "g2", "g3"], "B": ["g10", "g11", "g3"], "C": ["g20", "g21", "g22", "g23"]}
    m = ref.iscDEF(ad, markers, add_other=1, markers_layer=markers_layer, use_brd=use_brd, seed=1)
    m.elbos, m.step_sizes = [], []
    case["Ks"] = [int(k) for k in m.layer_sizes]
    case["sg"] = [0] * len(case["Ks"])
    if markers_layer > 0:
        case["sg"][-1] = 1  # main phase of iscDEF.fit: root frozen
    return m, X, case, markers


@needs_reference
@pytest.mark.parametrize("markers_layer,use_brd", [(0, True), (1, True), (2, True), (0, False)])
def test_restated_iscdef_priors_and_elbo_equal_the_reference_execution(ref, markers_layer, use_brd):
    m, X, case, markers = _reference_iscdef(ref, markers_layer, use_brd)
    # prior construction restated in host_ref vs what the reference built
    sizes, names, per = host_ref.iscdef_layer_sizes(len(markers) + 1, markers_layer, add_root=markers_layer > 0,
                                                    n_layers_schedule=6)
    assert sizes == case["Ks"] and list(names) == list(m.layer_names)
    P0, R0 = host_ref.iscdef_geneset_prior(
        [f"g{i}" for i in range(case["G"])], markers, list(markers) + ["other0"], sizes, markers_layer, per,
        gs_small_scale=1.0, gs_big_scale=10.0, marker_strength=1.0, nonmarker_strength=1.0 if use_brd else 0.1,
        other_strength=0.1)
    np.testing.assert_allclose(P0, np.asarray(m.w_priors[0][0], dtype=float), rtol=1e-12)
    np.testing.assert_allclose(R0, np.asarray(m.w_priors[0][1], dtype=float), rtol=1e-12)
    assert len(np.unique(P0)) >= 2
    # the ELBO with those matrix priors
    lp = ref_cases.perturb(m.local_params, seed=31)
    gp = ref_cases.perturb(m.global_params, seed=32)
    key = ref.random.PRNGKey(5)
    idx, loss_l, loss_g, g_loc, g_glob = _reference_eval(ref, m, case, X, lp, gp, key)
    om = ref_cases.oracle_model_from_reference(m)
    noise = ref_cases.noise_from_key(ref.random, key, case["S"], len(idx), 1, case["G"], case["Ks"])
    kw = (1.0, case["sg"], 0.0, 0.0, om.alpha)
    l3, gl3, gg3 = elbo_closed.loss_and_grads(om, X[idx], idx, lp, gp, noise, *kw)
    assert abs(l3 - loss_l) <= 1e-10 * abs(loss_l)
    for a, b in zip(list(gl3) + list(gg3), list(g_loc) + list(g_glob)):
        assert rel_err(a, b) <= 1e-9
    l1, gg1 = elbo_ref.global_loss_grad(om, X[idx], idx, lp, gp, noise, *kw)
    assert abs(float(l1) - loss_l) <= TOL * abs(loss_l)
    for a, b in zip(gg1, g_glob):
        assert rel_err(a.numpy(), b) <= TOL


@needs_reference
def test_layer_ladder_equals_the_reference_method(ref):
    """`scDEF._geometric_layer_sizes` (`_scdef.py:580-628`) and `sscDEF._geometric_layer_sizes_no_root`
    (`_sscdef.py:44-77`) executed from the reference vs the oracle's and the product's ladder, over a grid."""
    import itertools
    import types

    from scdef_b200.model import geometric_ladder

    n = 0
    for nf, tf, nl in itertools.product([1, 2, 5, 8, 20, 33, 64, 100, 128], [1, 2, 3, 5, 10, 24], [1, 2, 3, 4, 5, 6, 8]):
        if tf > nf:
            continue
        me = types.SimpleNamespace(n_factors=nf, top_factors=tf)
        want = [int(k) for k in ref.scDEF._geometric_layer_sizes(me, nl)]
        assert host_ref.geometric_layer_sizes(nf, tf, nl) == want, (nf, tf, nl)
        assert geometric_ladder(nf, tf, nl, root=True) == want, (nf, tf, nl)
        want_nr = [int(k) for k in ref.sscDEF._geometric_layer_sizes_no_root(nf, tf, nl)]
        assert geometric_ladder(nf, tf, nl, root=False) == want_nr, (nf, tf, nl)
        n += 1
    assert n > 200


@needs_reference
def test_restated_assign_confident_equals_the_reference_execution(ref):
    """`tl.assign_confident` (`tools/factor.py:1107-1477`) of the REFERENCE, run on a model whose per-cell blocks were
    set by hand, against `oracle/tools_ref.py` fed the very draws the reference consumed (`random.normal(sample_keys[s],
    shape)` of the two `lax.scan` passes: the shim's keys are replayed here)."""
    from oracle import tools_ref

    rng = np.random.default_rng(5)
    N, G, Ks, S, tau, level = 90, 40, (6, 3, 1), 64, 0.3, 0.9
    X = rng.poisson(1.5, (N, G)).astype(float)
    m = ref.scDEF(ref.adata(X), layer_sizes=list(Ks), seed=3)
    z = np.array(ref.to_numpy(m.local_params[1]), dtype=np.float64)
    z[0] = rng.normal(-0.5, 1.0, z[0].shape)
    z[1] = rng.normal(-1.0, 0.5, z[1].shape)
    m.local_params[1] = ref.jnp.array(z)
    m.factor_lists = [np.array([0, 2, 3, 5]), np.array([0, 1, 2]), np.array([0])]     # a filtered bottom layer
    m.factor_names = [[f"L{l}_{k}" for k in kept] for l, kept in enumerate(m.factor_lists)]
    ref.tools_factor.assign_confident(m, n_samples=S, tau=tau, credible_level=level)
    obs = m.adata.obs

    random = ref.random
    layer_keys = random.split(random.PRNGKey(int(m.seed)), len(Ks))                   # factor.py:1262
    off = np.concatenate([[0], np.cumsum(Ks)])
    conf, eff, sd, am = (np.zeros((N, len(Ks))) for _ in range(4))
    for l, kept in enumerate(m.factor_lists):
        cols = off[l] + kept
        keys = random.split(layer_keys[l], S)                                          # 1274
        eps = np.stack([np.asarray(random.normal(keys[s], shape=(N, len(kept)))) for s in range(S)])
        st = tools_ref.layer_stats(z[0][:, cols], z[1][:, cols], eps, level)
        name = m.layer_names[l]
        for field in ("confidence", "effect_size", "posterior_sd", "winner_mass", "winner_probability", "entropy_confidence"):
            got = np.asarray(m.adata.obsm[f"confident_{field}"], dtype=float)[:, l]     # factor.py:1452-1457
            assert np.allclose(got, st[field], rtol=0, atol=2e-6), (name, field, np.abs(got - st[field]).max())
        assert np.array_equal(np.asarray(m.adata.obsm["confident_argmax_factor"])[:, l], st["argmax"])
        labels = np.asarray(obs[f"confident_argmax_{name}"]).astype(str)
        assert np.array_equal(labels, np.asarray(m.factor_names[l], dtype=object)[st["argmax"]].astype(str))
        conf[:, l], eff[:, l], sd[:, l], am[:, l] = st["confidence"], st["effect_size"], st["posterior_sd"], st["argmax"]
    sel = tools_ref.select_layers(conf, eff, sd, am.astype(int), [len(k) for k in m.factor_lists], tau)
    want_layer = np.array([m.layer_names[i] if i >= 0 else "" for i in sel["best_layer_idx"]], dtype=object).astype(str)
    assert np.array_equal(np.asarray(obs["confident_best_layer"]).astype(str), want_layer)
    assert np.array_equal(np.asarray(obs["confident_best_layer_idx"]), sel["best_layer_idx"])
    assert np.array_equal(np.asarray(obs["confident_best_factor_idx"]), sel["best_factor_slot"])
    assert np.allclose(np.asarray(obs["confident_best_confidence"], dtype=float), sel["best_confidence"], atol=2e-6, equal_nan=True)
    assert np.allclose(np.asarray(obs["confident_best_effect_size"], dtype=float), sel["best_effect_size"], atol=2e-6, equal_nan=True)
    assert np.allclose(np.asarray(obs["confident_depth_score"], dtype=float), sel["depth_score"], atol=1e-12, equal_nan=True)


@needs_reference
def test_restated_confident_signatures_equal_the_reference_execution(ref):
    """`tl.set_confident_signatures` (`tools/factor.py:441-552`; helpers 29-363, 1652-1768; `get_signature_sample`,
    `_scdef.py:3953-3994`) of the REFERENCE on a model with hand-set W blocks, filtered factor lists and one technical
    L0 factor, against `oracle/tools_ref.py` fed the draws the reference consumed (the shim's key chain is replayed)."""
    import pandas as pd
    from oracle import tools_ref as tr

    rng = np.random.default_rng(11)
    N, G, Ks, S = 50, 64, (6, 4, 2), 20
    thr, tq = 0.6, 0.9
    m = ref.scDEF(ref.adata(rng.poisson(1.5, (N, G)).astype(float)), layer_sizes=list(Ks), seed=1)
    gp = [np.array(p, dtype=np.float64) for p in ref.to_numpy(m.global_params)]
    for l in range(len(Ks)):
        gp[2 + l][0] = rng.normal(-1.0, 1.0, gp[2 + l][0].shape)
        gp[2 + l][1] = rng.normal(-1.5, 0.3, gp[2 + l][1].shape)
    m.global_params = [ref.jnp.array(p) for p in gp]
    m.set_posterior_means()
    m.set_posterior_variances()
    m.factor_lists = [np.array([0, 1, 3, 4, 5]), np.array([0, 2, 3]), np.array([0, 1])]
    m.factor_names = [[f"L{l}_{k}" for k in kept] for l, kept in enumerate(m.factor_lists)]
    m.adata.uns["factor_obs"] = pd.DataFrame({"technical": [n == "L0_3" for l in m.factor_names for n in l]},
                                             index=[n for l in m.factor_names for n in l])
    tf = ref.tools_factor
    tf.set_confident_signatures(m, confidence_threshold=thr, tau_quantile=tq, mc_samples=S, random_seed=7)
    cache = m.adata.uns["confident_signatures"]
    genes = {g: i for i, g in enumerate(np.asarray(m.adata.var_names))}

    # ---- the same with the oracle, on the reference's draws --------------------------------------------------------
    random = ref.random
    fl = [np.asarray(f) for f in m.factor_lists]
    l0_keep = np.array([i for i, n in enumerate(m.factor_names[0]) if n != "L0_3"])
    W_mu, W_rho = [gp[2 + l][0] for l in range(len(Ks))], [gp[2 + l][1] for l in range(len(Ks))]
    pm = [np.asarray(m.pmeans[f"{m.layer_names[l]}W"], float) for l in range(len(Ks))]
    pv0 = np.asarray(m.pvars[f"{m.layer_names[0]}W"], float)
    base = random.PRNGKey(7)
    upper = {l: [] for l in range(1, len(Ks))}
    for s in range(S):                                              # factor.py:284-290, 243-266
        key = random.fold_in(base, s)
        eps = []
        key, sk = random.split(key)
        eps.append(np.asarray(random.normal(sk, shape=(len(fl[0]), G))))
        for l in range(1, len(Ks)):
            key, sk = random.split(key)
            eps.append(np.asarray(random.normal(sk, shape=(len(fl[l]), len(fl[l - 1])))))
        for l, sc in tr.hierarchy_scores_draw(W_mu, W_rho, fl, l0_keep, eps).items():
            upper[l].append(sc)
    for l in range(len(Ks)):
        means = tr.layer_term_means(pm, fl, l0_keep, l)
        if l == 0:
            sig = tr.layer0_confident(means, pv0[fl[0]], thr, tq)

            def scores_of(f):                                        # factor.py:95-110, _scdef.py:3975-3989
                rows = []
                for s in range(S):
                    _, sk = random.split(random.fold_in(base, 0 * 1_000_003 + f * S + s))
                    e = np.asarray(random.normal(sk, shape=(len(fl[0]), G)))
                    rows.append(tr.lognormal_draw(W_mu[0][fl[0]], W_rho[0][fl[0]], e)[f])
                return np.stack(rows)
        else:
            mc = np.stack(upper[l])
            sig = tr.confident_from_scores(means, mc, thr, tq)
            scores_of = lambda f, mc=mc: mc[:, f, :]
        combined = [tr.confidence_mean_score(c, means[f][idx]) for f, (idx, c) in enumerate(sig)]
        jac = tr.jaccard_from_scores([idx for idx, _ in sig], combined, scores_of)
        got = cache["by_layer"][str(l)]
        for f, name in enumerate(m.factor_names[l]):
            want_genes = [genes[g] for g in got["signatures"][name]]
            assert want_genes == list(sig[f][0]), (l, name)
            assert np.allclose(got["confidences"][name], sig[f][1], atol=1e-12)
            assert np.allclose(got["combined_scores"][name], combined[f], rtol=1e-10, atol=1e-12)
            assert np.isclose(got["signature_confidences"][name], jac[f], atol=1e-12, equal_nan=True), (l, name)
        assert any(len(idx) > 0 for idx, _ in sig), f"layer {l}: every signature is empty, the case tests nothing"


@needs_reference
@pytest.mark.parametrize("min_cells_upper,filter_up", [(0.001, True), (2, True), (2, False), (0.5, False), (40, True)])
def test_restated_upper_layer_filter_equals_the_reference_execution(ref, min_cells_upper, filter_up):
    """`scDEF.filter_factors(upper_only=True)` (`_scdef.py:3478-3568`, the call `_learn` makes at 3386-3387) of the
    REFERENCE on hand-set posterior means vs `oracle/host_ref.filter_factors_upper_only` (row f1: the device supplies
    its arg-max counts through `scdef_assignment_counts`)."""
    rng = np.random.default_rng(17)
    N, G, Ks = 60, 30, (6, 4, 2)
    m = ref.scDEF(ref.adata(rng.poisson(1.0, (N, G)).astype(float)), layer_sizes=list(Ks), seed=0)
    pz = [rng.gamma(0.3, 1.0, (N, K)) for K in Ks]
    pz[1][:, 3] = 0.0                                      # a factor nobody is assigned to
    pW = [None] + [rng.gamma(0.5, 1.0, (Ks[l], Ks[l - 1])) for l in range(1, len(Ks))]
    for l, name in enumerate(m.layer_names):
        m.pmeans[f"{name}z"] = pz[l]
        if l > 0:
            m.pmeans[f"{name}W"] = pW[l]
    m.filter_factors(upper_only=True, min_cells_upper=min_cells_upper, filter_up=filter_up, annotate=False)
    want = host_ref.filter_factors_upper_only(pz, pW, list(Ks), N, min_cells_upper=min_cells_upper, filter_up=filter_up)
    assert len(m.factor_lists) == len(want)
    for got, w in zip(m.factor_lists, want):
        assert np.array_equal(np.asarray(got), np.asarray(w))


@needs_reference
def test_posterior_summaries_equal_the_reference_execution(ref):
    """`set_posterior_means` / `set_posterior_variances` (`_scdef.py:3394-3476`, row a9) of the REFERENCE vs this package's
    host path on the same parameters: same keys (including the `brd` / `wm` / `ard` aliases), same values."""
    from scdef_b200 import MiniAnnData, scDEF

    rng = np.random.default_rng(23)
    N, G, Ks = 40, 24, (5, 3, 1)
    X = rng.poisson(1.0, (N, G)).astype(float)
    m = ref.scDEF(ref.adata(X), layer_sizes=list(Ks), seed=0)
    lp = [np.array(p, dtype=np.float64) for p in ref.to_numpy(m.local_params)]
    gp = [np.array(p, dtype=np.float64) for p in ref.to_numpy(m.global_params)]
    for p in lp + gp:
        p[0] = rng.normal(-0.5, 1.0, p[0].shape)
        p[1] = rng.normal(-1.0, 0.4, p[1].shape)
    m.local_params = [ref.jnp.array(p) for p in lp]
    m.global_params = [ref.jnp.array(p) for p in gp]
    m.set_posterior_means()
    m.set_posterior_variances()

    mine = scDEF(MiniAnnData(X.astype(np.float32)), layer_sizes=list(Ks), seed=0, logginglevel=30)
    mine.local_params = [p.astype(np.float32) for p in lp]
    mine.global_params = [p.astype(np.float32) for p in gp]
    mine.set_posterior_means()
    mine.set_posterior_variances()
    assert set(mine.pmeans) == set(m.pmeans) and set(mine.pvars) == set(m.pvars)
    for k in m.pmeans:
        assert np.allclose(np.asarray(mine.pmeans[k], float), np.asarray(ref.to_numpy(m.pmeans[k]), float), rtol=2e-6), k
    for k in m.pvars:
        assert np.allclose(np.asarray(mine.pvars[k], float), np.asarray(ref.to_numpy(m.pvars[k]), float), rtol=1e-5), k
