"""Generates tests/golden/ref_elbo.npz and tests/golden/ref_learn.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

`oracle/ref_exec.py` imports the unmodified `/root/reference/src/scdef/{utils/jax_utils.py, models/_scdef.py,
models/_sscdef.py, models/_iscdef.py}` on top of the torch-backed jax / tfp / optax stand-ins of
`oracle/jaxshim.py` (float64) and this script records, for every case of `tests/ref_cases.py`:
  * what went in: counts, minibatch indices, the model constants the reference's constructor computed,
    variational parameters (float32 values), the N(0,1) draws its `lognormal_sample` calls consumed
    (float32 values), flags;
  * what came out of `local_loss_grad` / `global_loss_grad` (`_scdef.py:2848-2849`): loss and gradients;
  * for the `_learn` cases: per-epoch losses and the final parameters of the reference's hot loop
    (`_scdef.py:3215-3271` with `optax.adam`, `clip_params`, `data_stream`).
The vectors travel to the GPU box (where `/root/reference` does not exist); `tests/test_golden_ref.py`
compares the CUDA path and the restated oracle against them.

Run from the repo root (needs /root/reference):  python tests/golden/make_golden_ref.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import jaxshim, ref_exec  # noqa: E402
from tests import ref_cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def f32(a):
    return np.asarray(a, dtype=np.float64).astype(np.float32).astype(np.float64)


def constants(m):
    """What `elbo` reads off the reference object (`_scdef.py:1848, 1949, 1963, 2017-2043, 2068`)."""
    out = dict(
        batch_indices_onehot=np.asarray(m.batch_indices_onehot, dtype=np.float64),
        batch_lib_ratio=np.asarray(m.batch_lib_ratio, dtype=np.float64),
        gene_ratio=np.asarray(m.gene_ratio, dtype=np.float64),
        cell_alpha_factor=np.asarray(m.cell_alpha_factor, dtype=np.float64),
        layer_sizes=np.asarray(m.layer_sizes, dtype=np.int64),
        scalars=np.asarray([m.alpha, m.top_alpha, m.brd, m.brd_mean, m.shrinkage_shape, m.shrinkage_rate,
                            m.cell_scale_shape, m.gene_scale_shape, float(m.use_brd), float(m.marginalize_alpha)],
                           dtype=np.float64),
    )
    for l, (a, b) in enumerate(m.w_priors):
        out[f"w_prior_shape{l}"] = np.asarray(a, dtype=np.float64)
        out[f"w_prior_rate{l}"] = np.asarray(b, dtype=np.float64)
    return out


def evaluate(ref, m, case, X, key_seed):
    lp = [f32(p) for p in ref_cases.perturb(m.local_params, seed=11)]
    gp = [f32(p) for p in ref_cases.perturb(m.global_params, seed=12)]
    key = ref.random.PRNGKey(key_seed)
    idx = np.random.RandomState(0).permutation(case["N"])[: case["B"]]
    jnp = ref.jnp
    local_fn, global_fn = m._get_or_build_learn_grad_fns(case["S"])
    args = (jnp.array(X[idx]), jnp.array(idx), [jnp.array(p) for p in lp], [jnp.array(p) for p in gp], key,
            case.get("anneal", 1.0), jnp.array(np.asarray(case["sg"], dtype=np.float64)),
            case.get("stop_cell", 0.0), case.get("stop_gene", 0.0), m.alpha)
    loss, g_loc = local_fn(*args)
    loss_g, g_glob = global_fn(*args)
    assert float(loss) == float(loss_g)
    noise = ref_cases.noise_from_key(ref.random, key, case["S"], len(idx), case["nb"], case["G"], case["Ks"])
    out = dict(X=X.astype(np.float32), idx=idx.astype(np.int64), loss=np.float64(float(loss)))
    out.update(constants(m))
    out.update(ref_cases.noise_arrays(noise))
    for i, p in enumerate(lp):
        out[f"lp{i}"] = p.astype(np.float32)
        out[f"gl{i}"] = ref.to_numpy(g_loc)[i][:, idx]  # the rows outside the minibatch are exactly zero
        assert np.abs(np.delete(ref.to_numpy(g_loc)[i], idx, axis=1)).max(initial=0.0) == 0.0
    for i, p in enumerate(gp):
        out[f"gp{i}"] = p.astype(np.float32)
        out[f"gg{i}"] = ref.to_numpy(g_glob)[i]
    return out


def learn(ref, m, case, X, n_epoch, **learn_kw):
    # float32-valued starting point, as the CUDA path will hold it
    jnp = ref.jnp
    m.local_params = [jnp.array(f32(p)) for p in m.local_params]
    m.global_params = [jnp.array(f32(p)) for p in m.global_params]
    if hasattr(m, "_pin_supervised_top_z_local_params"):
        m._pin_supervised_top_z_local_params()
        m.local_params = [jnp.array(f32(p)) for p in m.local_params]
    out = dict(X=X.astype(np.float32))
    out.update(constants(m))
    for i, p in enumerate(m.local_params):
        out[f"lp{i}"] = np.asarray(p, dtype=np.float32)
    for i, p in enumerate(m.global_params):
        out[f"gp{i}"] = np.asarray(p, dtype=np.float32)
    m._learn(n_epoch=n_epoch, num_samples=case["S"], batch_size=case["B"], filter=False, annotate=False,
             min_epochs=n_epoch + 1, **learn_kw)
    out["epoch_losses"] = np.asarray(m.elbos[-1], dtype=np.float64)
    for i, p in enumerate(m.local_params):
        out[f"lp_final{i}"] = np.asarray(p, dtype=np.float64)
    for i, p in enumerate(m.global_params):
        out[f"gp_final{i}"] = np.asarray(p, dtype=np.float64)
    # the draws of every step, in order: PRNGKey(seed) -> split per step (3181, 3231) -> split(S) (2163)
    rng = ref.random.PRNGKey(m.seed)
    stream_len = -(-case["N"] // case["B"])
    perm_rng = np.random.RandomState(0)
    step = 0
    for _ in range(n_epoch):
        perm = perm_rng.permutation(case["N"])
        for it in range(stream_len):
            rng, rng_input = ref.random.split(rng)
            b = len(perm[it * case["B"]:(it + 1) * case["B"]])
            noise = ref_cases.noise_from_key(ref.random, rng_input, case["S"], b, case["nb"], case["G"], case["Ks"])
            for k, v in ref_cases.noise_arrays(noise).items():
                out[f"step{step}_{k}"] = v.astype(np.float32)
            step += 1
    out["n_steps"] = np.int64(step)
    return out


def main():
    assert ref_exec.available(), "needs /root/reference"
    jaxshim.set_noise_float32(True)
    elbo_out, learn_out, meta = {}, {}, {"elbo": {}, "learn": {}}
    with ref_exec.reference_runtime(torch.float64) as ref:
        src = {k: ref.source_lines(v) for k, v in dict(elbo=ref.scDEF.elbo, batch_elbo=ref.scDEF.batch_elbo,
                                                       learn=ref.scDEF._learn, ss_elbo=ref.sscDEF.elbo).items()}
        meta["source"] = {k: f"{p}:{a}-{b}" for k, (p, a, b) in src.items()}

        def add(group, store, name, data, info):
            for k, v in data.items():
                store[f"{name}/{k}"] = v
            meta[group][name] = info

        for name, case in ref_cases.ELBO_CASES.items():
            m, X, batch = ref_cases.build_reference_model(ref, case)
            data = evaluate(ref, m, case, X, key_seed=7)
            data["batch"] = np.asarray(batch, dtype=np.int64)
            add("elbo", elbo_out, name, data, dict(case, kind="scDEF"))

        # sscDEF: pinned layer, alpha floor 1.0 (`_sscdef.py:484-487, 505-507, 530-531`)
        for add_root in (True, False):
            case = dict(N=90, G=60, Ks=[8, 3, 1] if add_root else [8, 3], nb=1, B=30, S=2,
                        sg=[0, 0, 1] if add_root else [0, 0])
            X, _ = ref_cases.make_counts(case)
            ad = ref.adata(X)
            labels = np.asarray([i % 3 for i in range(case["N"])])
            ad.obs["celltype"] = [f"t{i}" for i in labels]
            m = ref.sscDEF(ad, "celltype", add_root=add_root, layer_sizes=list(case["Ks"]), seed=1)
            m.elbos, m.step_sizes = [], []
            data = evaluate(ref, m, case, X, key_seed=3)
            data["fixed_top_z"] = np.asarray(m._supervised_top_z, dtype=np.float64)
            data["labels"] = labels
            name = "sscdef_root" if add_root else "sscdef_noroot"
            add("elbo", elbo_out, name, data, dict(case, kind="sscDEF", supervised_layer=int(m.supervised_top_layer_idx),
                                                  alpha_floor=1.0))
            if add_root:
                m2 = ref.sscDEF(ad, "celltype", add_root=True, layer_sizes=list(case["Ks"]), seed=1)
                m2.elbos, m2.step_sizes = [], []
                d2 = learn(ref, m2, case, X, 2, optimize_layers=[0, 1])
                d2["fixed_top_z"] = np.asarray(m2._supervised_top_z, dtype=np.float64)
                d2["labels"] = labels
                add("learn", learn_out, "sscdef_root", d2,
                    dict(case, kind="sscDEF", n_epoch=2, learn_kw=dict(optimize_layers=[0, 1]),
                         supervised_layer=int(m2.supervised_top_layer_idx), alpha_floor=1.0, seed=1))

        # iscDEF: matrix-valued W priors (`_iscdef.py:272-450`)
        markers = {"A": ["g1", "g2", "g3"], "B": ["g10", "g11", "g3"], "C": ["g20", "g21", "g22", "g23"]}
        for ml, use_brd in ((0, True), (1, True), (0, False)):
            case = dict(N=90, G=60, nb=1, B=30, S=2)
            X, _ = ref_cases.make_counts(case)
            m = ref.iscDEF(ref.adata(X), markers, add_other=1, markers_layer=ml, use_brd=use_brd, seed=1)
            m.elbos, m.step_sizes = [], []
            case["Ks"] = [int(k) for k in m.layer_sizes]
            case["sg"] = [0] * len(case["Ks"])
            if ml > 0:
                case["sg"][-1] = 1
            data = evaluate(ref, m, case, X, key_seed=5)
            add("elbo", elbo_out, f"iscdef_ml{ml}_{'brd' if use_brd else 'nobrd'}", data,
                dict(case, kind="iscDEF", markers_layer=ml, use_brd=use_brd, markers=markers, add_other=1))

        for name, kw in (("base", {}), ("marg_alpha", {}), ("two_batches", {"freeze_w": True}),
                         ("root_frozen", {"optimize_layers": [0, 1]}), ("hyper", {"annealing": 1.7, "lr": 0.05, "local_lr": 0.02})):
            case = dict(ref_cases.ELBO_CASES[name])
            m, X, batch = ref_cases.build_reference_model(ref, case)
            data = learn(ref, m, case, X, 2, **kw)
            data["batch"] = np.asarray(batch, dtype=np.int64)
            add("learn", learn_out, name, data, dict(case, kind="scDEF", n_epoch=2, learn_kw=kw, seed=int(m.seed)))

    elbo_out["meta"] = np.asarray(json.dumps(meta))
    learn_out["meta"] = np.asarray(json.dumps(meta))
    for fname, store in (("ref_elbo.npz", elbo_out), ("ref_learn.npz", learn_out)):
        path = os.path.join(HERE, fname)
        np.savez_compressed(path, **store)
        print("wrote", path, os.path.getsize(path), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
