"""CPU baseline for bench.py: the reference's algorithm for one hot-loop iteration
(`_scdef.py:3230-3271`: value_and_grad of the objective w.r.t. locals, Adam, then w.r.t.
globals, Adam, clip) restated with torch CPU autograd in float32, all host threads.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py).  The reference's own JAX path cannot
run in this image (no jax); this is the "restated CPU oracle (PyTorch), not JAX" of BASELINE.md §3.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from . import elbo_ref, host_ref
from .elbo_ref import Noise


def time_train_steps(X, layer_sizes, batch_size=256, num_samples=100, steps=2, warmup=1, seed=0, threads=None,
                     budget_s=None):
    """Times `steps` reference-semantics train steps on the count matrix X (cells x genes).
    Returns dict(steps_per_s, cells_per_s, ms_per_step, cores, batch_used).

    budget_s: wall-clock bound for the timed steps.  If the warm-up steps predict that `steps` steps of `batch_size`
    cells would take longer, the timed steps use fewer cells per step (never fewer than 8): the per-sample costs that
    do not depend on the minibatch (sampling W_0, its prior) then weigh more, so the reported cells/s is a slight
    under-estimate of the full-batch rate; `batch_used` says what was run."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(int(threads))
    X = np.asarray(X, dtype=np.float32)
    N, G = X.shape
    model = host_ref.build_model(X, list(layer_sizes))
    lp, gp = host_ref.init_var_params(model, seed=seed)
    lo = host_ref.Adam(lp, 1e-2, dtype=np.float32)
    go = host_ref.Adam(gp, 1e-1, dtype=np.float32)
    stream = host_ref.data_stream_indices(N, batch_size)
    sg = [0.0] * len(layer_sizes)
    times, warm_times = [], []
    batch_used = int(batch_size)
    for it in range(warmup + steps):
        idx = next(stream)
        if it == warmup and budget_s is not None and warm_times:
            t_est = float(np.min(warm_times)) * steps
            if t_est > float(budget_s):
                batch_used = max(8, int(batch_size * float(budget_s) / t_est))
        idx = idx[:batch_used]
        noise = Noise.draw(num_samples, len(idx), 1, G, layer_sizes, seed=seed + it, dtype=torch.float32)
        t0 = time.perf_counter()
        Xb = X[idx]
        _, gl = elbo_ref.local_loss_grad(model, Xb, idx, lp, gp, noise, 1.0, sg, 0, 0, model.alpha, dtype=torch.float32)
        lp = host_ref.clip_params(lo.update(lp, [g.numpy() for g in gl]))
        loss, gg = elbo_ref.global_loss_grad(model, Xb, idx, lp, gp, noise, 1.0, sg, 0, 0, model.alpha, dtype=torch.float32)
        gp = host_ref.clip_params(go.update(gp, [g.numpy() for g in gg]))
        dt = time.perf_counter() - t0
        (times if it >= warmup else warm_times).append(dt)
    ms = 1e3 * float(np.mean(times))
    return dict(steps_per_s=1e3 / ms, cells_per_s=1e3 / ms * batch_used, ms_per_step=ms, cores=int(threads),
                loss=float(loss), batch_used=batch_used)
