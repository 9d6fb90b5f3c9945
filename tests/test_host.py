"""CPU: host logic of the drop-in class, the minibatch stream, the ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import host_ref
from scdef_b200 import MiniAnnData, scDEF, _lib
from scdef_b200.dist import ShardInfo, local_members
from scdef_b200.synth import synth_counts_numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(N=200, G=120, nb=1, **kw):
    X, b = synth_counts_numpy(N, G, n_batches=nb, seed=0)
    obs = {"batch": np.array([f"b{i}" for i in range(max(nb, 1))])[b]} if nb > 1 else None
    return X, b, scDEF(MiniAnnData(X, obs=obs), batch_key="batch" if nb > 1 else None, **kw)


def test_geometric_ladder_golden_values():
    # /root/reference/tests/test_sscdef.py:55-83 and SURVEY.md §8c
    _, _, m = _model(20, 10)
    assert m.layer_sizes == [100, 40, 16, 6, 3, 1]
    assert host_ref.geometric_layer_sizes(100, 1, 6) == [100, 40, 16, 6, 3, 1]
    _, _, m = _model(20, 10, n_factors=20, top_factors=5, n_layers=3)
    assert m.layer_sizes == [20, 10, 5, 1]
    assert host_ref.geometric_layer_sizes(20, 5, 3) == [20, 10, 5, 1]
    for nf, tf, nl in [(100, 1, 6), (50, 2, 4), (10, 1, 2), (8, 8, 3), (30, 3, 5)]:
        _, _, m = _model(20, 10, n_factors=nf, top_factors=tf, n_layers=nl)
        assert m.layer_sizes == host_ref.geometric_layer_sizes(nf, tf, nl)


def test_constructor_validation_matches_reference():
    X, _ = synth_counts_numpy(20, 10)
    with pytest.raises(TypeError, match="AnnData"):
        scDEF(X)
    with pytest.raises(ValueError, match="n_factors"):
        scDEF(MiniAnnData(X), n_factors=0)
    with pytest.raises(ValueError, match="top_factors must be >= 1"):
        scDEF(MiniAnnData(X), top_factors=0)
    with pytest.raises(ValueError, match="top_factors must be <= n_factors"):
        scDEF(MiniAnnData(X), n_factors=4, top_factors=5)


@pytest.mark.parametrize("nb", [1, 3])
def test_constants_match_oracle_restatement(nb):
    X, b, m = _model(150, 60, nb=nb, layer_sizes=[10, 5, 2], seed=1)
    ref = host_ref.build_model(X, [10, 5, 2], batch_ids=b if nb > 1 else None)
    np.testing.assert_allclose(m.batch_lib_ratio, ref.batch_lib_ratio, rtol=1e-10)
    np.testing.assert_allclose(np.broadcast_to(m.gene_ratio, (nb, 60)), np.broadcast_to(ref.gene_ratio, (nb, 60)), rtol=1e-10)
    np.testing.assert_allclose(m.cell_alpha_factor, ref.cell_alpha_factor, rtol=1e-12)
    np.testing.assert_allclose(m.gene_ratio_init, ref.gene_ratio_init, rtol=1e-10)
    assert m.alpha == pytest.approx(ref.alpha, rel=1e-12)
    np.testing.assert_array_equal(m.batch_indices_onehot, ref.batch_indices_onehot)
    for l in range(3):
        np.testing.assert_allclose(np.asarray(m.w_priors[l][0]), np.asarray(ref.w_priors[l][0]))
    # parameter layout
    assert [p.shape for p in m.local_params] == [(2, 150, 1), (2, 150, 17)]
    assert [p.shape for p in m.global_params] == [(2, nb, 60), (2, 10, 1), (2, 10, 60), (2, 5, 10), (2, 2, 5), (2, 10, 1), (2, 1, 1)]
    assert all(p.dtype == np.float32 for p in m.local_params + m.global_params)
    # posterior summaries follow _scdef.py:3394-3476
    np.testing.assert_allclose(m.pmeans["L0W"], host_ref.lognormal_mean(*m.global_params[2]), rtol=1e-6)
    np.testing.assert_allclose(m.pvars["L1z"], host_ref.lognormal_var(m.local_params[1][0][:, 10:15], m.local_params[1][1][:, 10:15]), rtol=1e-6)


def test_init_matches_reference_moments():
    # cold-start c and s are deterministic (App. B): exact agreement with the oracle restatement
    X, b, m = _model(150, 60, layer_sizes=[10, 5, 2], seed=1)
    ref = host_ref.build_model(X, [10, 5, 2])
    lp, gp = host_ref.init_var_params(ref, seed=1)
    np.testing.assert_allclose(m.local_params[0], lp[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.global_params[0], gp[0], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(m.global_params[5], gp[5], rtol=1e-5, atol=1e-6)  # ard
    np.testing.assert_allclose(m.global_params[6], gp[6], rtol=1e-5, atol=1e-6)  # alpha


def test_learn_kwargs_are_strict():
    _, _, m = _model(50, 20, layer_sizes=[4, 2])
    with pytest.raises(TypeError, match="Unexpected keyword arguments for _learn: bogus"):
        m._learn(bogus=1)
    assert scDEF.learn is scDEF._learn


def test_learn_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    _, _, m = _model(50, 20, layer_sizes=[4, 2])
    with pytest.raises(Exception, match="no CPU fallback"):
        m._learn(n_epoch=1)


@pytest.mark.parametrize("N,bs", [(10, 3), (257, 256), (100, None), (64, 64), (5, 8)])
def test_minibatch_stream_is_bit_exact(N, bs):
    """`data_stream` (_scdef.py:3167-3173): RandomState(0), permutation per epoch, short last slice."""
    import torch

    from scdef_b200.stream import MinibatchStream

    bsz = N if bs is None else bs
    st = MinibatchStream(N, bsz, ShardInfo(), torch.device("cpu"))
    ref = host_ref.data_stream_indices(N, bs)
    rng = np.random.RandomState(0)
    for epoch in range(3):
        st.new_epoch()
        perm = rng.permutation(N)
        for it in range(st.num_batches):
            expect = perm[it * bsz:(it + 1) * bsz]
            np.testing.assert_array_equal(next(ref), expect)
            idx, Xb, n_glob = st.batch(it)
            assert idx.dtype == torch.int64
            np.testing.assert_array_equal(idx.numpy(), expect)
            assert n_glob == len(expect) and Xb is None


def test_sharded_stream_partitions_the_reference_batch():
    import torch

    from scdef_b200.stream import MinibatchStream

    N, b, R = 103, 8, 4
    counts = [26, 26, 26, 25]
    starts = np.concatenate([[0], np.cumsum(counts)])
    streams = [MinibatchStream(N, b * R, ShardInfo(rank=r, world=R), torch.device("cpu"), n_local=counts[r], start=int(starts[r]))
               for r in range(R)]
    ref = host_ref.data_stream_indices(N, b * R)
    for epoch in range(2):
        for s in streams:
            s.new_epoch()
        for it in range(streams[0].num_batches):
            expect = next(ref)
            got = np.concatenate([s.batch(it)[0].numpy() + starts[r] for r, s in enumerate(streams)])
            assert sorted(got.tolist()) == sorted(expect.tolist())
            assert all(s.batch(it)[2] == len(expect) for s in streams)
            for r, s in enumerate(streams):
                np.testing.assert_array_equal(s.batch(it)[0].numpy(), local_members(expect, int(starts[r]), counts[r]))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "scdef_b200.h")).read()
    declared = set(re.findall(r"\b(scdef_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    path = _lib.lib_path()
    assert os.path.exists(path), "build the extension first (__graft_entry__.build())"
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.scdef_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call
    assert lib.scdef_ctx_create(None, None) == -1
    d = _lib.ModelDesc()
    d.abi_version = 999
    h = ctypes.c_void_p()
    assert lib.scdef_ctx_create(ctypes.byref(d), ctypes.byref(h)) == -1


def test_struct_layout_matches_header():
    """ctypes mirrors are the same size as the C structs (checked by compiling a probe)."""
    import subprocess
    import tempfile

    src = '#include <stdio.h>\n#include "scdef_b200.h"\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(scdef_model_desc), sizeof(scdef_blocks), sizeof(scdef_noise), sizeof(scdef_call));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "p.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.ModelDesc), ctypes.sizeof(_lib.Blocks), ctypes.sizeof(_lib.Noise), ctypes.sizeof(_lib.Call)]


def test_sscdef_constructor_matches_reference_test():
    """/root/reference/tests/test_sscdef.py:8-35 and :55-83 (layer sizes, names, pinned z)."""
    from scdef_b200 import sscDEF

    X, _ = synth_counts_numpy(80, 120, seed=0)
    rng = np.random.RandomState(0)
    types = rng.choice(["A", "B"], size=80)
    m = sscDEF(MiniAnnData(X, obs={"cell_type": types}), top_key="cell_type", n_layers=2, n_factors=4, seed=1)
    assert m.n_layers == 3 and m.layer_sizes == [4, 2, 1]
    assert m.layer_names[m.supervised_top_layer_idx] == "cell_type"
    assert len(m.w_priors) == 3 and np.asarray(m.w_priors[1][0]).shape == (2, 4)
    top = m._supervised_top_z
    assert top.shape == (80, 2)
    np.testing.assert_allclose(top.sum(1), 1.0 + m._Z_OFF_TYPE, rtol=1e-5)
    np.testing.assert_allclose(top.max(1), m._Z_ON)
    np.testing.assert_allclose(m.pmeans["cell_typez"], top, rtol=1e-5, atol=1e-5)
    start, end = m._layer_column_range(m.supervised_top_layer_idx)
    zs, zr = m.local_params[1]
    np.testing.assert_allclose(np.exp(zs[:, start:end] + 0.5 * np.exp(zr[:, start:end]) ** 2), top, rtol=1e-3, atol=1e-3)
    assert m.factor_names[m.supervised_top_layer_idx] == ["A", "B"]
    m2 = sscDEF(MiniAnnData(np.ones((10, 5), dtype=np.float32) + np.arange(5), obs={"t": np.array(list("aabbccddee"))}),
                top_key="t", n_factors=20, n_layers=3)
    assert m2.layer_sizes == sscDEF._geometric_layer_sizes_no_root(20, 5, 3) + [1] == [20, 10, 5, 1]
    with pytest.raises(KeyError):
        sscDEF(MiniAnnData(X), top_key="nope")


def test_filter_factors_upper_only_restatement_known_answer():
    """oracle/host_ref.filter_factors_upper_only on a hand-made case (`_scdef.py:3518-3568`)."""
    # 12 cells; layer 1 (3 factors): cells assigned 10 / 2 / 0 -> with min 2 cells keep {0, 1}; factor 2 is empty
    z1 = np.zeros((12, 3)); z1[:10, 0] = 1.0; z1[10:, 1] = 1.0
    z0 = np.ones((12, 4))
    # layer-0 factors attach (argmax over the kept rows of W_1) to: f0->0, f1->0, f2->0, f3->0  => filter_up drops 1
    W1 = np.array([[5.0, 4.0, 3.0, 2.0], [1.0, 1.0, 1.0, 1.0], [9.0, 9.0, 9.0, 9.0]])
    # layer 2 (2 factors): all cells on factor 1; W_2's kept row attaches layer-1 factor 0 to factor 1
    z2 = np.zeros((12, 2)); z2[:, 1] = 1.0
    W2 = np.array([[3.0, 0.0, 0.0], [1.0, 2.0, 0.0]])
    out = host_ref.filter_factors_upper_only([z0, z1, z2], [None, W1, W2], [4, 3, 2], 12, min_cells_upper=2)
    assert [o.tolist() for o in out] == [[0, 1, 2, 3], [0], [1]]
    out = host_ref.filter_factors_upper_only([z0, z1, z2], [None, W1, W2], [4, 3, 2], 12, min_cells_upper=2, filter_up=False)
    assert [o.tolist() for o in out] == [[0, 1, 2, 3], [0, 1], [1]]
    # fraction -> max(frac * n, 10): nothing reaches 10 cells in layer 2's factor 0, factor 1 has 12
    out = host_ref.filter_factors_upper_only([z0, z1, z2], [None, W1, W2], [4, 3, 2], 12, min_cells_upper=0.5, filter_up=False)
    assert [o.tolist() for o in out] == [[0, 1, 2, 3], [0], [1]]
    # an empty layer falls back to all of its factors
    out = host_ref.filter_factors_upper_only([z0, z1, z2], [None, W1, W2], [4, 3, 2], 12, min_cells_upper=11, filter_up=False)
    assert [o.tolist() for o in out] == [[0, 1, 2, 3], [0, 1, 2], [1]]


def test_count_statistics_of_sparse_counts_equal_dense():
    import scipy.sparse as sp

    from scdef_b200 import dist

    X, b = synth_counts_numpy(150, 60, n_batches=3, seed=5)
    labels = np.array(["p", "q", "r"])[b]
    d = dist.count_statistics(X, labels, dist.ShardInfo())
    s = dist.count_statistics(sp.csr_matrix(X), labels, dist.ShardInfo())
    for k in d:
        if isinstance(d[k], np.ndarray) and d[k].dtype.kind == "f":
            np.testing.assert_allclose(s[k], d[k], rtol=1e-12)
        elif isinstance(d[k], np.ndarray):
            np.testing.assert_array_equal(s[k], d[k])
        elif isinstance(d[k], float):
            assert s[k] == pytest.approx(d[k], rel=1e-12)
        else:
            assert s[k] == d[k] or list(s[k]) == list(d[k])
    m = scDEF(MiniAnnData(sp.csr_matrix(X), obs={"batch": labels}), batch_key="batch", layer_sizes=[6, 3], x_placement="csr")
    m2 = scDEF(MiniAnnData(X, obs={"batch": labels}), batch_key="batch", layer_sizes=[6, 3])
    assert m.alpha == pytest.approx(m2.alpha, rel=1e-12) and m.X.nnz == np.count_nonzero(X)
    for a, c in zip(m.global_params, m2.global_params):
        np.testing.assert_allclose(a, c, rtol=1e-6)


def test_iscdef_hierarchical_ladder_names_and_priors():
    """/root/reference/tests/test_integration.py:441-470 (block layout of the marker hierarchy) + the prior matrices
    of `_iscdef.py:272-450`."""
    from scdef_b200 import iscDEF

    X, _ = synth_counts_numpy(60, 80, seed=0)
    genes = [f"g{i}" for i in range(80)]
    markers = {"T": genes[:3], "B": genes[3:6]}
    m = iscDEF(MiniAnnData(X, var_names=genes), markers_dict=markers, markers_layer=2, add_other=0, decay_factor=2, seed=1)
    assert m.layer_sizes == [8, 4, 2, 1] and m.layer_names == ["L0", "L1", "marker", "root"]
    assert m.n_factors_per_marker == 4 and m.use_brd
    assert all(m.factor_names[0][s].startswith("T_L0_") for s in range(4))
    assert all(m.factor_names[0][s].startswith("B_L0_") for s in range(4, 8))
    assert m.factor_names[2] == ["T", "B"] and m.factor_names[3] == ["root_0"]
    shape0, rate0 = (np.asarray(a) for a in m.w_priors[0])
    assert shape0.shape == rate0.shape == (8, 80)
    scale0 = shape0 / rate0
    np.testing.assert_allclose(scale0[:4, :3], 10.0, rtol=1e-6)       # T block x T markers: gs_big_scale
    np.testing.assert_allclose(scale0[:4, 3:6], 1e-6, rtol=1e-5)      # T block x B markers: penalised
    np.testing.assert_allclose(scale0[4:, 3:6], 10.0, rtol=1e-6)
    np.testing.assert_allclose(scale0[:, 6:], 1.0, rtol=1e-6)         # everything else: gs_small_scale
    np.testing.assert_allclose(shape0[:4, :3], 1.0)                   # marker_strength
    np.testing.assert_allclose(shape0[:4, 3:6], 0.1, rtol=1e-6)       # other_strength
    np.testing.assert_allclose(shape0[:, 6:], 1.0)                    # nonmarker_strength -> 1.0 under BRD
    # connectivity prior of W_1 (4 x 8): factor u of L1 owns L0 factors {2u, 2u+1}
    s1, r1 = (np.asarray(a) for a in m.w_priors[1])
    own = np.zeros((4, 8), dtype=bool)
    for u in range(4):
        own[u, 2 * u: 2 * u + 2] = True
    base = float(m.factor_shapes[1])
    np.testing.assert_allclose(s1[own], base * 1.0, rtol=1e-6)        # cn_big_strength
    np.testing.assert_allclose(s1[~own], base * 0.1, rtol=1e-6)       # cn_small_strength
    np.testing.assert_allclose((s1 / r1)[own], 10.0, rtol=1e-5)       # mean = cn_big_mean
    np.testing.assert_allclose((s1 / r1)[~own], 1.0, rtol=1e-5)       # mean = cn_small_mean
    # the root's W keeps the plain scDEF prior
    np.testing.assert_allclose(np.asarray(m.w_priors[3][0]), base, rtol=1e-6)
    with pytest.raises(ValueError):
        iscDEF(MiniAnnData(X, var_names=genes), markers_dict=markers, markers_layer=0, add_root=True)


def test_iscdef_flat_markers_layer():
    """`markers_layer=0`: one factor per marker set plus `other`, halving ladder above (`_iscdef.py:169-177`)."""
    from scdef_b200 import iscDEF

    X, _ = synth_counts_numpy(60, 40, seed=0)
    genes = [f"g{i}" for i in range(40)]
    m = iscDEF(MiniAnnData(X, var_names=genes), markers_dict={"B": genes[:2], "Mono": genes[2:4]}, markers_layer=0,
               add_other=1, n_layers=2, seed=1)
    assert m.layer_sizes == [3, 1] and m.layer_names == ["marker", "L1"] and not m.add_root
    assert m.factor_names[0] == ["B", "Mono", "other0"]
    scale0 = np.asarray(m.w_priors[0][0]) / np.asarray(m.w_priors[0][1])
    np.testing.assert_allclose(scale0[0, :2], 10.0, rtol=1e-6)
    np.testing.assert_allclose(scale0[2, :4], 1e-6, rtol=1e-5)        # `other` factors are pushed off every typed marker
    m6 = iscDEF(MiniAnnData(X, var_names=genes), markers_dict={f"m{i}": [genes[i]] for i in range(9)}, markers_layer=0)
    assert m6.layer_sizes == [9, 4, 2, 1]


@pytest.mark.parametrize("markers_layer,add_other,penalize,use_brd", [(0, 0, True, True), (0, 2, True, False), (1, 1, True, True),
                                                                      (2, 0, False, True), (2, 1, True, True), (3, 0, True, True)])
def test_iscdef_priors_match_the_loop_restatement(markers_layer, add_other, penalize, use_brd):
    """The vectorised prior construction of scdef_b200.iscDEF against oracle/host_ref's loop-by-loop restatement of
    `_iscdef.py:167-450`, with overlapping marker lists, a marker missing from the data and `other` groups."""
    from scdef_b200 import iscDEF

    X, _ = synth_counts_numpy(40, 50, seed=1)
    genes = [f"g{i}" for i in range(50)]
    markers = {"T": genes[:4], "B": genes[3:8] + ["not_in_data"], "NK": genes[10:12]}
    kw = dict(gs_small_scale=0.7, gs_big_scale=12.0, marker_strength=3.0, nonmarker_strength=0.2, other_strength=0.05,
              cn_small_mean=0.5, cn_big_mean=7.0, cn_small_strength=0.3, cn_big_strength=1.5)
    m = iscDEF(MiniAnnData(X, var_names=genes), markers_dict=markers, markers_layer=markers_layer, add_other=add_other,
               penalize_other=penalize, use_brd=use_brd, seed=1, logginglevel=40, **kw)
    names = list(markers) + [f"other{i}" for i in range(add_other)]
    sizes, lnames, per = host_ref.iscdef_layer_sizes(len(names), markers_layer, m.add_root)
    assert m.layer_sizes == sizes and m.layer_names == lnames
    nonmarker = 1.0 if use_brd else kw["nonmarker_strength"]
    shape0, rate0 = host_ref.iscdef_geneset_prior(genes, markers, names, sizes, markers_layer, per, kw["gs_small_scale"],
                                                  kw["gs_big_scale"], kw["marker_strength"], nonmarker, kw["other_strength"],
                                                  penalize_other=penalize)
    np.testing.assert_allclose(np.asarray(m.w_priors[0][0]), shape0, rtol=1e-6)
    np.testing.assert_allclose(np.asarray(m.w_priors[0][1]), rate0, rtol=1e-6)
    if markers_layer > 0:
        mult = host_ref.iscdef_connectivity_prior(sizes, len(names), markers_layer, m.add_root, 2, kw["cn_small_mean"],
                                                  kw["cn_big_mean"], kw["cn_small_strength"], kw["cn_big_strength"])
        for l in range(1, len(sizes)):
            base = float(m.factor_shapes[l])
            want = mult.get(l, (1.0, 1.0))
            np.testing.assert_allclose(np.asarray(m.w_priors[l][0]), base * np.asarray(want[0]) * np.ones((sizes[l], sizes[l - 1])), rtol=1e-5)
            np.testing.assert_allclose(np.asarray(m.w_priors[l][1]), base * np.asarray(want[1]) * np.ones((sizes[l], sizes[l - 1])), rtol=1e-5)


def test_persistent_pair_kernel_work_decomposition():
    """Chunks of the flattened (row block, sample, pair-tile) space and the partial-buffer slots of the persistent
    CTA-pair kernel (csrc/lik_flat.cuh), both orientations, checked through the library's own arithmetic for 68 shapes
    (in a subprocess: plain ctypes, no torch).  No CUDA call."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_flat_plan_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "flat plan ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_hier_mma_fragment_maps():
    """NumPy replay of the tensor-core form of the hierarchy kernel's three small products (`k_hier_tiled<true>`,
    kernels_small.cuh): per-lane fragment index maps of mma.sync.m16n8k8 (A row-major, B column-major), the tile loops, the
    zero padding and the stores, on the kernel's shared-memory layouts ([k][cell] with stride 33; W_l [u][ld])."""
    HLD, HT, HNW = 33, 32, 16
    rng = np.random.default_rng(0)

    def mma_step(c, af, bf):          # c: (32 lanes, 4); one k-step of C[16 x 8] += A[16 x 8] B[8 x 8]
        A, Bm = np.zeros((16, 8)), np.zeros((8, 8))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = af(g, t), af(g + 8, t), af(g, t + 4), af(g + 8, t + 4)
            Bm[t, g], Bm[t + 4, g] = bf(t, g), bf(t + 4, g)
        D = A @ Bm
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            c[lane] += [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]

    for Ku, Kl in ((60, 100), (30, 60), (10, 30), (1, 10)):
        ld = (Kl + 3) // 4 * 4
        z = rng.normal(size=(Ku, HLD)); P = rng.normal(size=(Kl, HLD)); W = np.zeros((Ku, ld)); W[:, :Kl] = rng.normal(size=(Ku, Kl))
        z[:, HT:] = np.nan; P[:, HT:] = np.nan                        # the padding column must never be read
        # forward: m[k][cell] = sum_u z[u][cell] W[u][k]
        Pm = np.full((Kl, HLD), np.nan)
        n_nt = (Kl + 7) // 8
        for warp in range(HNW):
            for tile in range(warp, 2 * n_nt, HNW):
                m0, n0 = (tile & 1) * 16, (tile >> 1) * 8
                c = np.zeros((32, 4))
                for k0 in range(0, Ku, 8):
                    mma_step(c, lambda r, kk: z[k0 + kk, m0 + r] if k0 + kk < Ku else 0.0,
                             lambda kk, n: W[k0 + kk, n0 + n] if (k0 + kk < Ku and n0 + n < Kl) else 0.0)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    if n0 + 2 * t < Kl:
                        Pm[n0 + 2 * t, m0 + g], Pm[n0 + 2 * t, m0 + g + 8] = c[lane, 0], c[lane, 2]
                    if n0 + 2 * t + 1 < Kl:
                        Pm[n0 + 2 * t + 1, m0 + g], Pm[n0 + 2 * t + 1, m0 + g + 8] = c[lane, 1], c[lane, 3]
        assert np.allclose(Pm[:, :HT], (z[:, :HT].T @ W[:, :Kl]).T)
        # backward, LOCAL: gz[u][cell] += sum_k P[k][cell] W[u][k]
        gz = np.zeros((Ku, HLD))
        n_nt = (Ku + 7) // 8
        for warp in range(HNW):
            for tile in range(warp, 2 * n_nt, HNW):
                m0, n0 = (tile & 1) * 16, (tile >> 1) * 8
                c = np.zeros((32, 4))
                for k0 in range(0, Kl, 8):
                    mma_step(c, lambda r, kk: P[k0 + kk, m0 + r] if k0 + kk < Kl else 0.0,
                             lambda kk, n: W[n0 + n, k0 + kk] if (k0 + kk < Kl and n0 + n < Ku) else 0.0)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    if n0 + 2 * t < Ku:
                        gz[n0 + 2 * t, m0 + g] += c[lane, 0]; gz[n0 + 2 * t, m0 + g + 8] += c[lane, 2]
                    if n0 + 2 * t + 1 < Ku:
                        gz[n0 + 2 * t + 1, m0 + g] += c[lane, 1]; gz[n0 + 2 * t + 1, m0 + g + 8] += c[lane, 3]
        assert np.allclose(gz[:, :HT], W[:, :Kl] @ P[:, :HT])
        # backward, GLOBAL: GW[u][k] += sum_cell z[u][cell] P[k][cell]
        GW = np.zeros((Ku, ld))
        n_mt, n_nt = (Ku + 15) // 16, (Kl + 7) // 8
        for warp in range(HNW):
            for tile in range(warp, n_mt * n_nt, HNW):
                m0, n0 = (tile % n_mt) * 16, (tile // n_mt) * 8
                c = np.zeros((32, 4))
                for k0 in range(0, HT, 8):
                    mma_step(c, lambda r, kk: z[m0 + r, k0 + kk] if m0 + r < Ku else 0.0,
                             lambda kk, n: P[n0 + n, k0 + kk] if n0 + n < Kl else 0.0)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for i in range(4):
                        u, k = m0 + g + (i >> 1) * 8, n0 + 2 * t + (i & 1)
                        if u < Ku and k < Kl:
                            GW[u, k] += c[lane, i]
        assert np.allclose(GW[:, :Kl], z[:, :HT] @ P[:, :HT].T)
        assert np.all(GW[:, Kl:] == 0)
