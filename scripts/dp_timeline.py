"""Kernel timeline of a few data-parallel training steps (torchrun, rank 0 prints). Debug tool."""
import os
import sys

import torch
import torch.distributed as td

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import MiniAnnData, scDEF  # noqa: E402
from scdef_b200.synth import CONFIGS, synth_counts_torch  # noqa: E402

lr = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
if world > 1:
    td.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
dev = torch.device(f"cuda:{lr}")
torch.cuda.set_device(dev)
N, G, Ks, nb, lib = CONFIGS["C5"]
N = int(os.environ.get("CELLS", 200000))
X, _ = synth_counts_torch(N // world, G, n_batches=1, lib=lib, seed=1000 + lr, device=dev)
m = scDEF(MiniAnnData(X), layer_sizes=Ks, seed=42, logginglevel=30, device=str(dev),
          process_group="world" if world > 1 else None, epoch_reshard=os.environ.get("RESHARD", "1") == "1")
prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU])
A = int(os.environ.get("START", 30))
B_ = A + 6


def hook(i, eng, stream):
    if i == A:
        torch.cuda.synchronize()
        prof.start()
    if i == B_:
        torch.cuda.synchronize()
        prof.stop()


m._step_hook = hook
m.max_steps_per_learn = B_ + 2
m._learn(n_epoch=3, num_samples=100, batch_size=256)
if lr == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    last_end = t0
    for e in evs:
        gap = e.time_range.start - last_end
        print(f"{(e.time_range.start - t0):9.1f} us  dur {e.time_range.end - e.time_range.start:8.1f}  gap {gap:7.1f}  {e.name[:70]}")
        last_end = max(last_end, e.time_range.end)
if world > 1:
    td.destroy_process_group()
