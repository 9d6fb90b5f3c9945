"""Host-side cost of one training step: a tiny workload (GPU time negligible) so wall time = enqueue cost.
Run on the GPU box:  timeout 300 python tests/host_overhead.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import MiniAnnData, scDEF  # noqa: E402
from scdef_b200.synth import synth_counts_numpy  # noqa: E402

X, _ = synth_counts_numpy(4096, 512, seed=0)
m = scDEF(MiniAnnData(X), layer_sizes=[100, 60, 30, 10], seed=1, logginglevel=30)
m.max_steps_per_learn = 64 * 30
t = {}


def hook(step, eng, stream):
    if step == 64 * 5:
        torch.cuda.synchronize(); t["a"] = time.perf_counter()
    if step == 64 * 25:
        torch.cuda.synchronize(); t["b"] = time.perf_counter()


m._step_hook = hook
pr = cProfile.Profile()
pr.enable()
m._learn(n_epoch=30, num_samples=1, batch_size=64)
pr.disable()
print("host+gpu ms/step (tiny workload, under cProfile):", (t["b"] - t["a"]) / (64 * 20) * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
m2 = scDEF(MiniAnnData(X), layer_sizes=[100, 60, 30, 10], seed=1, logginglevel=30)
m2.max_steps_per_learn = 64 * 30
m2._step_hook = hook
m2._learn(n_epoch=30, num_samples=1, batch_size=64)
print("host+gpu ms/step (tiny workload, no profiler):", (t["b"] - t["a"]) / (64 * 20) * 1e3)
