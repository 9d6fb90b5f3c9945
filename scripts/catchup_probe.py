"""Steady-state cost of the lazy-Adam catch-up (C5, 1 GPU): times the catch-up of the next minibatch's rows alone."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import MiniAnnData, scDEF  # noqa: E402
from scdef_b200.synth import CONFIGS, synth_counts_torch  # noqa: E402

dev = torch.device("cuda:0")
N, G, Ks, nb, lib = CONFIGS["C5"]
X, _ = synth_counts_torch(N, G, n_batches=1, lib=lib, seed=1000, device=dev)
m = scDEF(MiniAnnData(X), layer_sizes=Ks, seed=42, logginglevel=30, device="cuda:0")
spe = math.ceil(N / 256)
times, upd = [], []


def hook(i, eng, stream):
    if i > spe + 5 and i < spe + 45 and (i % spe) < stream.num_batches:
        idx, _, _ = stream.batch(i % spe)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        eng.local_catchup(idx)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
        rs = eng.row_step[idx]
        upd.append(int((eng.local_step - rs.min()).item()))


m._step_hook = hook
m.max_steps_per_learn = spe + 50
m._learn(n_epoch=2, num_samples=100, batch_size=256, filter=False)
print("catch-up of one minibatch in steady state: median %.4f ms, min %.4f, max %.4f (n=%d)" %
      (np.median(times), np.min(times), np.max(times), len(times)))
