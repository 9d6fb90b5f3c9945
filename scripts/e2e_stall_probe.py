"""Where does the host-placement (e2e) path lose time?  For every step: did the minibatch's H2D copy finish before the
compute stream needed it (CUDA-event timestamps), and how long did the host spend in stream.batch().
   python scripts/e2e_stall_probe.py [cells] [steps]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import MiniAnnData, scDEF, stream as _stream
from scdef_b200.synth import synth_counts_torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
X, _ = synth_counts_torch(N, 5000, lib=3000.0, device="cuda")
Xh = X.cpu().numpy()
del X
m = scDEF(MiniAnnData(Xh), layer_sizes=[100, 60, 30, 10], logginglevel=30, x_placement="host")
rec = []
orig = _stream.MinibatchStream.batch


def batch(self, it):
    dev = self.device
    t0 = time.perf_counter()
    before = torch.cuda.Event(enable_timing=True)
    before.record(torch.cuda.current_stream(dev))          # the compute stream's position when it asks for the batch
    fut = self._staged.get(it)
    done_already = fut.done() if fut is not None else None
    out = orig(self, it)
    after = torch.cuda.Event(enable_timing=True)
    after.record(torch.cuda.current_stream(dev))           # ... and right after the wait on the copy
    rec.append((before, after, time.perf_counter() - t0, done_already))
    return out


_stream.MinibatchStream.batch = batch
tg, tj = [], []
og, oj = _stream.MinibatchStream._gather, _stream.MinibatchStream._stage_job


def gather(self, rows, out):
    t = time.perf_counter(); og(self, rows, out); tg.append(time.perf_counter() - t)


def stage_job(self, rows, k, extra=None):
    t = time.perf_counter(); r = oj(self, rows, k, extra); tj.append(time.perf_counter() - t); return r


variant = sys.argv[3] if len(sys.argv) > 3 else ""
if "nogather" in variant:
    def gather(self, rows, out):   # noqa: F811 - variant: skip the host gather (stale pinned rows)
        tg.append(0.0)
if "slots" in variant:
    n_slots = int(variant.split("slots")[1][:1])
    oinit = _stream.MinibatchStream.__init__

    def init(self, *a, **k):
        k["pinned_slots"] = n_slots
        oinit(self, *a, **k)

    _stream.MinibatchStream.__init__ = init
marks = {}


def hook(step, eng, stream):
    if step in (100, K - 1):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(eng.device))
        marks[step] = e


m._step_hook = hook
_stream.MinibatchStream._gather = gather
_stream.MinibatchStream._stage_job = stage_job
m.max_steps_per_learn = K
t0 = time.perf_counter()
m._learn(n_epoch=3, num_samples=100, batch_size=256, filter=False, annotate=False)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
gaps = np.array([b.elapsed_time(a) for b, a, _, _ in rec[20:]])
host = np.array([h for _, _, h, _ in rec[20:]]) * 1e3
ready = np.array([bool(d) for _, _, _, d in rec[20:]])
print(f"steps {len(rec)}  wall ms/step {wall / len(rec) * 1e3:.4f}")
print(f"compute-stream time between 'asks for the batch' and 'has it' (ms): median {np.median(gaps):.4f}  p90 {np.percentile(gaps, 90):.4f}  max {gaps.max():.4f}")
print(f"host time in stream.batch() (ms): median {np.median(host):.4f}  p90 {np.percentile(host, 90):.4f}  max {host.max():.4f};  staging future already done: {ready.mean():.2f}")
print(f"helper thread per job (ms): whole job median {np.median(tj[20:]) * 1e3:.4f}, of which the row gather {np.median(tg[20:]) * 1e3:.4f}")
print(f"variant [{variant}]  GPU ms/step over steps 100..{K - 1}: {marks[100].elapsed_time(marks[K - 1]) / (K - 1 - 100):.4f}")
