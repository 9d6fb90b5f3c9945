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
    Returns dict(steps_per_s, cells_per_s, ms_per_step, cores, steps_used).

    budget_s: wall-clock bound for the timed steps.  If the warm-up steps predict that `steps` steps would take
    longer, fewer steps are timed (never fewer than 3; `steps_used` says how many).  Shrinking the minibatch instead
    helps little: a large part of a CPU step does not depend on the minibatch (sampling the S x K0 x G matrix W_0,
    its prior and entropy, and their backward passes) — measured on 8 cores: 5.8 s at B = 256, 4.1 s at B = 124."""
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
    steps_used = int(steps)
    for it in range(warmup + steps):
        if it == warmup and budget_s is not None and warm_times:
            t_step = float(np.min(warm_times))
            if t_step * steps > float(budget_s):
                steps_used = max(3, min(int(steps), int(float(budget_s) / t_step)))
        if it >= warmup + steps_used:
            break
        idx = next(stream)
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
    return dict(steps_per_s=1e3 / ms, cells_per_s=1e3 / ms * batch_size, ms_per_step=ms, cores=int(threads),
                loss=float(loss), steps_used=steps_used)
