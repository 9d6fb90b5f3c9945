"""NCCL sanity probe (torchrun): all-reduce latency of the global-gradient buffer and all-to-all bandwidth."""
import os
import torch
import torch.distributed as td

lr = int(os.environ.get("LOCAL_RANK", 0))
td.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
dev = torch.device(f"cuda:{lr}")
torch.cuda.set_device(dev)
R, rank = td.get_world_size(), td.get_rank()


def timeit(fn, n=50, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for mb in (0.004, 0.5, 4.1, 32):
    t = torch.ones(int(mb * 1e6 / 4), device=dev)
    ms = timeit(lambda: td.all_reduce(t))
    if rank == 0:
        print(f"all_reduce {mb} MB: {ms*1e3:.1f} us", flush=True)
n = 1 << 28
s = torch.ones(n, device=dev); r = torch.empty(n, device=dev)
ms = timeit(lambda: td.all_to_all_single(r, s), n=5, w=2)
if rank == 0:
    print(f"all_to_all_single 1 GiB/rank: {ms:.2f} ms -> {n*4*(R-1)/R/ms/1e6:.1f} GB/s off-rank per GPU", flush=True)
x = torch.ones(1 << 28, device=dev)
idx = torch.randperm(1 << 16, device=dev)
xv = x.view(1 << 16, -1)
ms = timeit(lambda: xv.index_select(0, idx), n=5, w=2)
if rank == 0:
    print(f"index_select 1 GiB rows: {ms:.2f} ms", flush=True)
td.destroy_process_group()
