"""GPU: Adam / clip / step-order parity against the oracle, the `_learn` loop, the seam
functions, golden vectors, and the data-parallel decomposition emulated on one device."""
import os

import numpy as np
import pytest
import torch

from oracle import elbo_closed, host_ref
from oracle.elbo_ref import Noise
from tests.helpers import make_engine, noise_to_spec, rel_err, small_problem

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "elbo_small.npz")


def test_golden_vectors():
    g = np.load(GOLD)
    Ks = tuple(int(k) for k in g["layer_sizes"])
    X, batch, idx = g["X"], g["batch"], g["idx"]
    model = host_ref.build_model(X, list(Ks), batch_ids=batch)
    lp = [g["lp0"], g["lp1"]]
    gp = [g[f"gp{i}"] for i in range(4 + len(Ks))]
    eng = make_engine(X, batch, model, lp, gp)
    from scdef_b200.engine import NoiseSpec

    spec = NoiseSpec.explicit(c=g["eps_c"], z=g["eps_z"], s=g["eps_s"], tau=g["eps_tau"], omega=g["eps_omega"],
                              a=g["eps_a"], W=[g[f"eps_W{l}"] for l in range(len(Ks))])
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    kw = dict(num_samples=3, annealing=float(g["annealing"]), stop_gradients=[0, 0, 0], alpha=float(g["alpha"]))
    eng.elbo_grad("local", idx_d, spec, **kw)
    assert abs(eng.loss_value() - float(g["loss"])) / abs(float(g["loss"])) < 1e-4
    for i, t in enumerate(eng.local_grads(len(idx))):
        assert rel_err(t.cpu().numpy(), g[f"gl{i}"][:, idx]) < 1e-4
    eng.elbo_grad("global", idx_d, spec, **kw)
    for i, t in enumerate(eng.glob_grad.views):
        ref = g[f"gg{i}"]
        if np.abs(ref).max() > 0:
            assert rel_err(t.cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize("freeze_w,freeze_z", [(False, ()), (True, (1,))])
def test_train_steps_follow_the_oracle(freeze_w, freeze_z):
    """Several iterations of the hot loop (_scdef.py:3230-3271) with injected noise: parameters after
    local Adam, snap-back, global Adam, freeze_w restore and clip agree with the float64 oracle."""
    N, G, Ks, nb, B, S = 80, 50, (8, 4, 2), 2, 24, 3
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=3)
    eng = make_engine(X, batch, model, lp, gp)
    lo = host_ref.Adam(lp, 1e-2, dtype=np.float64)
    go = host_ref.Adam(gp, 1e-1, dtype=np.float64)
    z_orig = np.asarray(lp[1], dtype=np.float64)
    off = np.concatenate([[0], np.cumsum(Ks)])
    slices = [(int(off[l]), int(off[l + 1])) for l in freeze_z]
    ref_lp = [p.astype(np.float64) for p in lp]
    ref_gp = [p.astype(np.float64) for p in gp]
    stream = host_ref.data_stream_indices(N, B)
    sg = [0, 0, 0]
    for step in range(4):
        idx = next(stream)
        noise = Noise.draw(S, len(idx), nb, G, Ks, seed=100 + step, dtype=torch.float32)
        ref_loss, ref_lp, ref_gp = host_ref.learn_step(
            model, X[idx], idx, ref_lp, ref_gp, noise, lo, go, 1.1, sg, 0, 0, model.alpha,
            freeze_w=freeze_w, freeze_z_slices=slices, z_orig=z_orig)
        idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
        spec = noise_to_spec(noise)
        kw = dict(num_samples=S, annealing=1.1, stop_gradients=sg, alpha=model.alpha)
        eng.elbo_grad("local", idx_d, spec, **kw)
        eng.adam_local(idx_d, 1e-2, freeze_z)
        eng.elbo_grad("global", idx_d, spec, **kw)
        eng.adam_global(1e-1, freeze_w=freeze_w)
        assert abs(eng.loss_value() - ref_loss) / abs(ref_loss) < 1e-4
        got_lp, got_gp = eng.get_params()
        for a, b in zip(got_lp + got_gp, ref_lp + ref_gp):
            assert np.abs(a - b).max() < 2e-4 * max(1.0, np.abs(b).max()), step
    if freeze_w:
        for i in range(2, 2 + len(Ks)):
            np.testing.assert_array_equal(got_gp[i], gp[i])
    for (s, e) in slices:
        np.testing.assert_array_equal(got_lp[1][:, :, s:e], lp[1][:, :, s:e])


def test_dense_adam_moves_rows_outside_the_minibatch():
    """optax semantics: rows with zero gradient still move through their momentum tail, and the
    update is |lr| in the first step for every touched element (m_hat/sqrt(v_hat) = sign)."""
    X, batch, model, lp, gp = small_problem(64, 40, (6, 3), 1, seed=5)
    eng = make_engine(X, batch, model, lp, gp)
    from scdef_b200.engine import NoiseSpec

    idx0 = torch.arange(0, 16, dtype=torch.int64, device="cuda")
    idx1 = torch.arange(16, 32, dtype=torch.int64, device="cuda")
    kw = dict(num_samples=2, stop_gradients=[0, 0], alpha=model.alpha)
    eng.elbo_grad("local", idx0, NoiseSpec.philox(1, 1), **kw)
    eng.adam_local(idx0, 1e-2)
    p1 = eng.get_params()[0][1]
    assert np.allclose(np.abs(p1[:, :16] - lp[1][:, :16]), 1e-2, atol=1e-6)
    np.testing.assert_array_equal(p1[:, 16:], lp[1][:, 16:])
    eng.elbo_grad("local", idx1, NoiseSpec.philox(1, 2), **kw)
    eng.adam_local(idx1, 1e-2)
    p2 = eng.get_params()[0][1]
    assert np.all(p2[:, :16] != p1[:, :16])        # momentum tail keeps moving rows 0..15
    np.testing.assert_array_equal(p2[:, 32:], lp[1][:, 32:])
    assert int(eng.row_slot.max().item()) == -1     # scratch restored


def test_global_clip_bounds():
    X, batch, model, lp, gp = small_problem(48, 30, (5, 2), 1, seed=6)
    gp = [p.copy() for p in gp]
    gp[0][0, 0, :5] = 2e4      # gene-scale mu above 1e4 -> clipped
    gp[2][1, 0, :5] = 11.5     # W0 rho above 10 -> clipped
    gp[4][0, 0, 0] = 2e4       # ard (second-to-last) is never clipped
    eng = make_engine(X, batch, model, lp, gp)
    from scdef_b200.engine import NoiseSpec

    idx = torch.arange(0, 16, dtype=torch.int64, device="cuda")
    eng.elbo_grad("global", idx, NoiseSpec.philox(1, 1), num_samples=2, stop_gradients=[0, 0], alpha=model.alpha)
    eng.adam_global(1e-1)
    got = eng.get_params()[1]
    assert np.all(got[0][0, 0, :5] == 1e4) and np.all(got[2][1, 0, :5] == 10.0)
    assert got[4][0, 0, 0] > 1e4


def test_philox_noise_is_reproducible_and_standard_normal():
    X, batch, model, lp, gp = small_problem(64, 300, (16, 4), 1, seed=7)
    eng = make_engine(X, batch, model, lp, gp)
    from scdef_b200.engine import NoiseSpec

    S, B = 64, 32
    spec = NoiseSpec.philox(7, 3, coupling=True)
    w0 = eng.noise_fill(spec, 4, 0, S, B, (16, 300))
    assert torch.equal(w0, eng.noise_fill(spec, 4, 0, S, B, (16, 300)))
    x = w0.double().flatten()
    assert abs(float(x.mean())) < 0.01 and abs(float(x.std()) - 1.0) < 0.01
    assert abs(float((x**3).mean())) < 0.03 and abs(float((x**4).mean()) - 3.0) < 0.1
    # reference coupling: brd (K0,1) and ard (K0,1) share their draws; independent mode does not
    assert torch.equal(eng.noise_fill(spec, 3, 0, S, B, (16, 1)), eng.noise_fill(spec, 5, 0, S, B, (16, 1)))
    ind = NoiseSpec.philox(7, 3, coupling=False)
    assert not torch.equal(eng.noise_fill(ind, 3, 0, S, B, (16, 1)), eng.noise_fill(ind, 5, 0, S, B, (16, 1)))
    assert not torch.equal(w0, eng.noise_fill(NoiseSpec.philox(7, 4), 4, 0, S, B, (16, 300)))


def test_philox_mode_matches_oracle_fed_the_same_draws():
    """Production noise path: dump the draws the kernels use, feed them to the oracle."""
    N, G, Ks, nb, B, S = 90, 64, (8, 4, 1), 1, 32, 3
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=8)
    eng = make_engine(X, batch, model, lp, gp)
    from scdef_b200.engine import NoiseSpec

    spec = NoiseSpec.philox(11, 5, coupling=True)
    f = lambda kind, layer, shape: eng.noise_fill(spec, kind, layer, S, B, shape).cpu()
    noise = Noise(c=f(0, 0, (B, 1)), s=f(2, 0, (nb, G)), tau=f(3, 0, (Ks[0], 1)), omega=f(5, 0, (Ks[0], 1)),
                  a=f(6, 0, (1, 1)), W=[f(4, l, (Ks[l], G if l == 0 else Ks[l - 1])) for l in range(3)],
                  z=torch.cat([f(1, l, (B, Ks[l])) for l in range(3)], dim=2))
    idx = np.random.RandomState(0).permutation(N)[:B]
    ref_loss, ref_gl, ref_gg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, 1.0, [0, 0, 0], 0, 0, model.alpha)
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    kw = dict(num_samples=S, stop_gradients=[0, 0, 0], alpha=model.alpha)
    eng.elbo_grad("local", idx_d, spec, **kw)
    assert abs(eng.loss_value() - ref_loss) / abs(ref_loss) < 1e-4
    for t, r in zip(eng.local_grads(B), ref_gl):
        assert rel_err(t.cpu().numpy(), r[:, idx]) < 1e-4
    eng.elbo_grad("global", idx_d, spec, **kw)
    for t, r in zip(eng.glob_grad.views, ref_gg):
        if np.abs(r).max() > 0:
            assert rel_err(t.cpu().numpy(), r) < 1e-4


def test_two_rank_decomposition_equals_one_rank():
    """SURVEY.md §8e emulated on one GPU: two engines each holding a block of the cells, global
    gradients summed with global_weight = 1/2, equal the single-engine gradients."""
    N, G, Ks, nb, S = 100, 60, (8, 4, 2), 2, 3
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=9)
    full = make_engine(X, batch, model, lp, gp)
    cut = 55
    parts = []
    for sl in (slice(0, cut), slice(cut, N)):
        import copy

        sub = copy.copy(model)
        sub.n_cells = sl.stop - sl.start
        sub.batch_indices_onehot = model.batch_indices_onehot[sl]
        sub.batch_lib_ratio = model.batch_lib_ratio[sl]
        sub.cell_alpha_factor = model.cell_alpha_factor[sl]
        parts.append(make_engine(X[sl], batch[sl], sub, [p[:, sl] for p in lp], gp))
    gidx = np.random.RandomState(0).permutation(N)[:40]
    noise = Noise.draw(S, 40, nb, G, Ks, seed=21, dtype=torch.float32)
    spec = noise_to_spec(noise)
    kw = dict(num_samples=S, stop_gradients=[0, 0, 0], alpha=model.alpha, annealing=1.2)
    full.elbo_grad("global", torch.as_tensor(gidx, device="cuda"), spec, scale=N / 40, **kw)
    ref = full.glob_grad.flat.clone()
    ref_loss = full.loss_value()
    total = torch.zeros_like(ref)
    loss = 0.0
    from scdef_b200.engine import NoiseSpec

    for r, (eng, start, n_loc) in enumerate(zip(parts, (0, cut), (cut, N - cut))):
        m = (gidx >= start) & (gidx < start + n_loc)
        loc = gidx[m] - start
        sub_spec = NoiseSpec.explicit(c=noise.c[:, m], z=noise.z[:, m], s=noise.s, tau=noise.tau, omega=noise.omega,
                                      a=noise.a, W=noise.W)
        eng.elbo_grad("global", torch.as_tensor(loc, device="cuda"), sub_spec, scale=N / 40, global_weight=0.5, **kw)
        total += eng.glob_grad.flat
        loss += eng.loss_value()
    assert abs(loss - ref_loss) / abs(ref_loss) < 1e-5
    assert rel_err(total.cpu().numpy(), ref.cpu().numpy()) < 1e-5


def _fit_model(**kw):
    from scdef_b200 import MiniAnnData, scDEF
    from scdef_b200.synth import synth_counts_numpy

    X, b = synth_counts_numpy(200, 120, seed=0)
    m = scDEF(MiniAnnData(X), layer_sizes=[10, 5, 2], seed=1, logginglevel=30, **kw)
    return X, m


def test_fit_c1_config_runs_and_improves():
    """BASELINE config C1 (tests/test_integration.py scale): fit(n_epoch=10), full batch."""
    X, m = _fit_model()
    m.fit(n_epoch=10, num_samples=10)
    assert len(m.elbos) == 1 and len(m.elbos[0]) == 10
    assert np.all(np.isfinite(m.elbos[0])) and m.elbos[0][-1] < m.elbos[0][0]
    assert len(m.adata.uns["entropy_annealing_trace"]) == 10
    assert m.pmeans["L0z"].shape == (200, 10) and m.pmeans["L0W"].shape == (10, 120)
    for v in list(m.pmeans.values()) + list(m.pvars.values()):
        assert np.all(np.isfinite(np.asarray(v)))
    np.testing.assert_allclose(m.pmeans["L1W"], host_ref.lognormal_mean(*m.global_params[3]), rtol=2e-5)
    np.testing.assert_allclose(m.pvars["cell_scale"], host_ref.lognormal_var(*m.local_params[0]), rtol=2e-4)
    with pytest.raises(NotImplementedError):
        m.fit(n_epoch=1)


def test_learn_phases_minibatches_and_annealing():
    X, m = _fit_model()
    m.fit(n_epoch=4, num_samples=5, batch_size=64, root_epochs=2, pretraining=True, pretraining_n_epoch=2)
    assert len(m.elbos) == 4  # 2 pretraining passes + main + root-only
    assert [len(e) for e in m.elbos] == [2, 2, 4, 2]
    X, m = _fit_model()
    m.fit(n_epoch=20, num_samples=3, entropy_anneal=True, entropy_window=5, entropy_check_every=1,
          entropy_rel_change_low=1e9, entropy_rel_change_high=1e10)
    tr = m.adata.uns["entropy_annealing_trace"]
    assert len(tr) == len(m.elbos[0]) and tr.min() >= 1.0 and tr.max() <= 5.0 and tr.max() > 1.0


def test_streamed_host_counts_equal_resident_counts():
    _, m1 = _fit_model()
    _, m2 = _fit_model(x_placement="host")
    m1.fit(n_epoch=3, num_samples=4, batch_size=64)
    m2.fit(n_epoch=3, num_samples=4, batch_size=64)
    np.testing.assert_allclose(m1.elbos[0], m2.elbos[0], rtol=1e-5)
    for a, b in zip(m1.global_params, m2.global_params):
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=1e-4)


def test_seam_functions_return_reference_shaped_outputs():
    X, m = _fit_model()
    idx = np.random.RandomState(0).permutation(200)[:50]
    from scdef_b200.engine import NoiseSpec

    args = (X[idx], idx, m.local_params, m.global_params, NoiseSpec.philox(1, 1), 1.0, np.zeros(3), 0, 0, m.alpha)
    loss_l, gl = m.local_loss_grad(*args, num_samples=4)
    loss_g, gg = m.global_loss_grad(*args, num_samples=4)
    assert loss_l == pytest.approx(loss_g, rel=1e-6)
    assert [g.shape for g in gl] == [p.shape for p in m.local_params]
    assert [g.shape for g in gg] == [p.shape for p in m.global_params]
    mask = np.ones(200, bool)
    mask[idx] = False
    assert np.all(gl[1][:, mask] == 0) and np.any(gl[1][:, idx] != 0)


@pytest.mark.parametrize("freeze_z", [(), (1,)])
def test_lazy_adam_is_bit_identical_to_dense(freeze_z):
    """The lazy local Adam replays the zero-gradient steps of untouched rows on demand; fed the same
    gradients, parameters AND both moments equal the dense kernel's bit for bit at every sync point,
    over several epochs (gaps of different lengths), with frozen z layers, and across an optimiser reset."""
    N, G, Ks, nb, B, S = 150, 64, (8, 4, 2), 1, 32, 2
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=12)
    from scdef_b200.engine import NoiseSpec

    dense, lazy = [make_engine(X, batch, model, lp, gp, lazy_adam=lz) for lz in (False, True)]
    stream = host_ref.data_stream_indices(N, B)

    def same_state():
        a, b = dense.get_params(), lazy.get_params()   # get_params() syncs the lazy engine
        for x, y in zip(a[0], b[0]):
            np.testing.assert_array_equal(x, y)
        for x, y in zip(dense.local_m.views + dense.local_v.views, lazy.local_m.views + lazy.local_v.views):
            assert torch.equal(x, y)

    for step in range(1, 16):  # ~3 epochs
        idx = torch.as_tensor(next(stream), dtype=torch.int64, device="cuda")
        lazy.local_catchup(idx)
        # rows of the minibatch are current before they are read (what the objective would see)
        for blk_d, blk_l in zip(dense.local.views, lazy.local.views):
            assert torch.equal(blk_d[:, idx], blk_l[:, idx])
        dense.elbo_grad("local", idx, NoiseSpec.philox(3, step), num_samples=S, stop_gradients=[0, 0, 0], alpha=model.alpha)
        lazy._ensure_batch(int(idx.numel()))
        lazy.local_grad.flat.copy_(dense.local_grad.flat)     # identical gradients for both optimisers
        dense.adam_local(idx, 1e-2, freeze_z)
        lazy.adam_local(idx, 1e-2, freeze_z)
        if step in (1, 6, 15):
            same_state()
    # a long gap: 1500 more steps in which only two rows receive (zero) gradients; every other row replays 1500
    # zero-gradient steps at once (p freezes, m decays into a denormal that no longer changes, v keeps decaying)
    two = torch.as_tensor([3, 77], dtype=torch.int64, device="cuda")
    dense._ensure_batch(2); lazy._ensure_batch(2)
    dense.local_grad.flat.zero_(); lazy.local_grad.flat.zero_()
    for _ in range(1500):
        dense.adam_local(two, 1e-2, freeze_z)
        lazy.adam_local(two, 1e-2, freeze_z)
    same_state()
    assert float(lazy.local_m.views[1].abs().max()) < 1e-30      # the first moment has decayed into the denormals
    dense.reset_optimizer(); lazy.reset_optimizer()           # a reset with pending steps flushes them first
    same_state()
    assert int(lazy.row_step.max().item()) == 0


@pytest.mark.parametrize("local_updates", ["reference", "keep"])
def test_sscdef_fit_keeps_the_top_layer_pinned(local_updates):
    """/root/reference/tests/test_sscdef.py:36-49: after fit, the top layer's posterior mean is the label one-hot.
    Default = the reference's literal loop (`_sscdef.py:892-893`: every local update is discarded, shown on the
    reference itself by tests/test_ref_pin.py); `local_updates="keep"` trains the per-cell blocks below the pinned layer."""
    from scdef_b200 import MiniAnnData, sscDEF
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(80, 120, seed=0)
    types = np.random.RandomState(0).choice(["A", "B"], size=80)
    m = sscDEF(MiniAnnData(X, obs={"cell_type": types}), top_key="cell_type", n_layers=2, n_factors=4, seed=1,
               logginglevel=30, local_updates=local_updates)
    before = [p.copy() for p in m.local_params]
    gbefore = [p.copy() for p in m.global_params]
    m.fit(n_epoch=2, n_rounds=1, num_samples=5)
    top = m._supervised_top_z
    np.testing.assert_allclose(m.pmeans["cell_typez"], top, rtol=1e-5, atol=1e-5)
    start, end = m._layer_column_range(m.supervised_top_layer_idx)
    zs, zr = m.local_params[1]
    np.testing.assert_allclose(np.exp(zs[:, start:end] + 0.5 * np.exp(zr[:, start:end]) ** 2), top, rtol=1e-3, atol=1e-3)
    assert len(m.elbos) == 2 and all(np.all(np.isfinite(e)) for e in m.elbos)   # main fit + root phase (root_epochs=10)
    assert np.any(m.global_params[2] != gbefore[2])                             # the globals always learn
    moved = np.any(m.local_params[1][0][:, :start] != before[1][0][:, :start])
    assert moved == (local_updates == "keep")


# ---- data parallel with position-based ownership: two ranks == one rank with twice the batch ---------------------
def _dp_problem():
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(600, 64, seed=3)
    return X, [10, 5, 2], dict(n_epoch=3, batch_size=64, num_samples=5, lr=0.05, local_lr=0.02)


def _dp_worker(rank, world, port, out, reshard, placement="device"):
    import torch.distributed as td

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    td.init_process_group("gloo", rank=rank, world_size=world)   # one GPU here: gloo, payloads staged through the host
    from scdef_b200 import MiniAnnData, scDEF

    X, Ks, kw = _dp_problem()
    sl = slice(0, 300) if rank == 0 else slice(300, 600)
    init = torch.load(out + ".init", weights_only=False)
    m = scDEF(MiniAnnData(X[sl]), layer_sizes=Ks, seed=1, logginglevel=30, process_group="world", epoch_reshard=reshard,
              x_placement=placement)
    m.local_params = [p[:, sl] for p in init["lp"]]
    m.global_params = init["gp"]
    Xdev_before = m._get_engine().X.clone() if placement == "device" else None
    m._learn(**kw)
    st = m._last_stream
    torch.save(dict(elbos=m.elbos[-1], lp=m.local_params, gp=m.global_params, factor_lists=[f.tolist() for f in m.factor_lists], moves=0 if st.migrator is None else st.migrator.n_moves,
                    x_restored=True if Xdev_before is None else bool(torch.equal(m._get_engine().X, Xdev_before))), out + f".{rank}")
    td.destroy_process_group()


@pytest.mark.parametrize("reshard,placement", [(True, "device"), (False, "device"), (True, "host")])
def test_two_ranks_equal_one_rank_with_twice_the_batch(tmp_path, reshard, placement):
    """SURVEY.md §8e end to end: `_learn` on 2 ranks (batch b each) against `_learn` on 1 rank with batch 2b.
    With position-based ownership the per-cell Philox rows line up too, so the trajectories agree up to
    floating-point summation order; with static ownership only the global minibatches are the same."""
    import socket

    import torch.multiprocessing as mp

    from scdef_b200 import MiniAnnData, scDEF

    X, Ks, kw = _dp_problem()
    one = scDEF(MiniAnnData(X), layer_sizes=Ks, seed=1, logginglevel=30)
    out = str(tmp_path / "dp")
    torch.save(dict(lp=[np.array(p) for p in one.local_params], gp=[np.array(p) for p in one.global_params]), out + ".init")
    one._learn(**dict(kw, batch_size=128))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_dp_worker, args=(2, port, out, reshard, placement), nprocs=2, join=True)
    r0, r1 = (torch.load(out + f".{r}", weights_only=False) for r in (0, 1))
    assert r0["x_restored"] and r1["x_restored"]
    assert r0["moves"] == (3 + 1 if reshard else 0)          # one move per epoch + the way home
    np.testing.assert_allclose(r0["elbos"], r1["elbos"], rtol=1e-12)
    if reshard:
        np.testing.assert_allclose(r0["elbos"], one.elbos[-1], rtol=2e-5)
        for i, p in enumerate(one.local_params):
            got = np.concatenate([r0["lp"][i], r1["lp"][i]], axis=1)
            np.testing.assert_allclose(got, p, rtol=2e-3, atol=2e-4)
        for a, b_, p in zip(r0["gp"], r1["gp"], one.global_params):
            np.testing.assert_array_equal(a, b_)              # replicated blocks stay bit-identical across ranks
            np.testing.assert_allclose(a, p, rtol=2e-3, atol=2e-4)
        # the hard-assignment counts behind filter_factors are summed over the ranks
        assert r0["factor_lists"] == r1["factor_lists"] == [f.tolist() for f in one.factor_lists]
    else:
        np.testing.assert_allclose(r0["elbos"], one.elbos[-1], rtol=5e-2)   # same minibatches, other per-cell noise


def test_assignment_counts_match_numpy_argmax_including_ties():
    """`scdef_assignment_counts` against np.argmax / count_nonzero on the pulled posterior means (`_scdef.py:3540-3549`)."""
    N, G, Ks, nb = 777, 40, (33, 7, 3), 1
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=4)
    lp = [np.array(p) for p in lp]
    lp[1][:, :50, :] = lp[1][:, :1, :]          # 50 identical rows
    lp[1][0, 100:140, 33:40] = 0.25             # exact ties inside layer 1 -> the first maximum wins
    lp[1][1, 100:140, 33:40] = -1.0
    eng = make_engine(X, batch, model, lp, gp)
    counts = [c.cpu().numpy() for c in eng.assignment_counts()]
    (means_l, _), _ = eng.posterior()
    zm = means_l[1].cpu().numpy()
    off = 0
    for l, K in enumerate(Ks):
        a = np.argmax(zm[:, off:off + K], axis=1)
        np.testing.assert_array_equal(counts[l], np.bincount(a, minlength=K))
        assert counts[l].sum() == N
        off += K
    assert counts[1][0] >= 40


def test_learn_filters_upper_layers_like_the_reference():
    X, m = _fit_model()
    m._learn(n_epoch=3, num_samples=5, batch_size=64)
    Ks = m.layer_sizes
    pz = [np.asarray(m.pmeans[f"{n}z"]) for n in m.layer_names]
    pW = [np.asarray(m.pmeans[f"{n}W"]) for n in m.layer_names]
    want = host_ref.filter_factors_upper_only(pz, pW, Ks, X.shape[0])
    assert [f.tolist() for f in m.factor_lists] == [w.tolist() for w in want]
    assert [len(n) for n in m.factor_names] == [len(w) for w in want]
    m.filter_factors(upper_only=True, min_cells_upper=1, filter_up=False)     # recomputed from the engine's parameters
    want = host_ref.filter_factors_upper_only(pz, pW, Ks, X.shape[0], min_cells_upper=1, filter_up=False)
    assert [f.tolist() for f in m.factor_lists] == [w.tolist() for w in want]
    with pytest.raises(NotImplementedError):
        m.filter_factors()


def test_csr_resident_counts_equal_dense():
    """x_placement='csr' (north star: "count matrix (dense or CSR)"): the minibatch densified from the CSR rows is the
    dense minibatch bit for bit, so training follows the dense path exactly (up to the order of float atomics)."""
    import scipy.sparse as sp

    from scdef_b200 import MiniAnnData, scDEF
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(300, 120, seed=2)
    md = scDEF(MiniAnnData(X), layer_sizes=[10, 5, 2], seed=1, logginglevel=30)
    ms = scDEF(MiniAnnData(sp.csr_matrix(X)), layer_sizes=[10, 5, 2], seed=1, logginglevel=30, x_placement="csr")
    eng_d, eng_s = md._get_engine(), ms._get_engine()
    idx = torch.as_tensor(np.random.RandomState(1).permutation(300)[:77], dtype=torch.int64, device="cuda")
    assert torch.equal(eng_s.gather_rows(idx), eng_d.X[idx])
    np.testing.assert_allclose(eng_s.x_lgamma.cpu().numpy(), eng_d.x_lgamma.cpu().numpy(), rtol=1e-6, atol=1e-6)
    kw = dict(n_epoch=3, num_samples=5, batch_size=64)
    md._learn(**kw)
    ms._learn(**kw)
    np.testing.assert_allclose(ms.elbos[-1], md.elbos[-1], rtol=1e-6)
    for a, c in zip(ms.local_params + ms.global_params, md.local_params + md.global_params):
        np.testing.assert_allclose(a, c, rtol=1e-4, atol=1e-5)


def test_iscdef_fit_runs_with_matrix_priors():
    """iscDEF end to end (matrix-valued W_0 prior -> fp32 likelihood path; frozen-root second phase)."""
    from scdef_b200 import MiniAnnData, iscDEF
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(120, 80, seed=0)
    genes = [f"g{i}" for i in range(80)]
    m = iscDEF(MiniAnnData(X, var_names=genes), markers_dict={"T": genes[:3], "B": genes[3:6]}, markers_layer=1,
               add_other=1, seed=1, logginglevel=30)
    assert m.layer_sizes == [6, 3, 1]
    m.fit(n_epoch=3, num_samples=5, batch_size=64)
    assert len(m.elbos) == 2 and len(m.elbos[1]) == 10          # main phase + 10 root epochs
    assert all(np.all(np.isfinite(e)) for e in m.elbos)
    assert m.pmeans["markerW"].shape == (3, 6) and np.all(np.isfinite(m.pmeans["L0W"]))
    # the informed prior does its job: the T block loads its markers more than the B markers
    W = np.asarray(m.pmeans["L0W"])
    assert W[:2, :3].mean() > W[:2, 3:6].mean()
    flat = iscDEF(MiniAnnData(X, var_names=genes), markers_dict={"T": genes[:3], "B": genes[3:6]}, markers_layer=0,
                  add_other=1, n_layers=2, seed=1, logginglevel=30)
    flat.fit(n_epoch=2, num_samples=5)
    assert len(flat.elbos) == 1 and np.all(np.isfinite(flat.elbos[0]))


def test_prior_changes_after_the_first_learn_reach_the_device():
    """The engine copies the W priors and prior hyper-parameters into its device-side description once; editing them
    afterwards (`update_model_priors`, `set_geneset_prior`, `model.brd = ...`) must rebuild it.  Same parameters, same
    noise step: the loss must move when a prior moves, and come back when it is restored."""
    from scdef_b200 import MiniAnnData, scDEF
    from scdef_b200.synth import synth_counts_numpy

    X, _ = synth_counts_numpy(120, 80, seed=0)
    m = scDEF(MiniAnnData(X), layer_sizes=[8, 4, 2], seed=1, logginglevel=30)
    m._learn(n_epoch=1, num_samples=3, filter=False, annotate=False)
    idx = np.arange(64)
    args = (X[idx], idx, m.local_params, m.global_params, 7, 1.0, [0, 0, 0], 0.0, 0.0, m.alpha)
    base, _ = m.global_loss_grad(*args, num_samples=3)
    m.brd = 3.0                                     # a scalar hyper-parameter
    moved, _ = m.global_loss_grad(*args, num_samples=3)
    assert abs(moved - base) > 1e-6 * abs(base)
    m.brd = 1.0
    back, _ = m.global_loss_grad(*args, num_samples=3)
    assert back == pytest.approx(base, rel=1e-6)
    m.factor_shape = 0.5                            # W priors rebuilt by update_model_priors (`_scdef.py:1246-1287`)
    m.update_model_priors()
    moved2, _ = m.global_loss_grad(*args, num_samples=3)
    assert abs(moved2 - base) > 1e-6 * abs(base)
