"""Latency of the all-reduce of the small gradient blocks (~27k floats) on N GPUs: NCCL vs torch symmetric memory's one-shot
kernel (peers' buffers read over NVLink).  torchrun --nproc-per-node N scripts/small_allreduce_probe.py"""
import os, sys, time
import torch, torch.distributed as td

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
td.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 26604
x = torch.full((n,), float(rank + 1), device="cuda")


def timeit(fn, iters=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


t_nccl = timeit(lambda: td.all_reduce(x))
big = torch.ones(1027000, device="cuda")
t_big = timeit(lambda: td.all_reduce(big))
msg = f"rank {rank}: NCCL all_reduce {n} floats {t_nccl:.1f} us;  4.1 MB {t_big:.1f} us"
try:
    import torch.distributed._symmetric_memory as sm

    buf = sm.empty(n, dtype=torch.float32, device=torch.device("cuda", lr))
    hdl = sm.rendezvous(buf, td.group.WORLD.group_name)
    buf.fill_(float(rank + 1))
    torch.cuda.synchronize(); td.barrier()
    out = torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", td.group.WORLD.group_name)
    torch.cuda.synchronize()
    want = world * (world + 1) / 2
    ok = bool(torch.all(out == want))
    t_sm = timeit(lambda: torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", td.group.WORLD.group_name))
    t_sm_copy = timeit(lambda: (buf.copy_(x), x.copy_(torch.ops.symm_mem.one_shot_all_reduce(buf, "sum", td.group.WORLD.group_name))))
    msg += f";  symm_mem one-shot {t_sm:.1f} us (correct: {ok}), with copies in/out {t_sm_copy:.1f} us"
except Exception as e:  # noqa: BLE001
    msg += f";  symm_mem unavailable: {type(e).__name__}: {str(e)[:200]}"
print(msg, flush=True)
td.destroy_process_group()
