"""Per-launch time of the fused likelihood kernel, LOCAL and GLOBAL pass separately, at C5 shapes (256 cells x 5000
genes, layers [100,60,30,10], 100 MC samples):   python tests/pair_timing.py [cells] [samples] [dump]
(`dump` needs a -DSCDEF_TC_TIMING build and prints where each role of the CTA-pair kernel spends its cycles.)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import MiniAnnData, scDEF
from scdef_b200.engine import NoiseSpec
from scdef_b200.synth import synth_counts_torch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
X, _ = synth_counts_torch(N, 5000, lib=3000.0, device="cuda")
m = scDEF(MiniAnnData(X), layer_sizes=[100, 60, 30, 10], logginglevel=30)
m.max_steps_per_learn = 3
m._learn(n_epoch=1, num_samples=S, batch_size=256, filter=False, annotate=False)
eng = m._get_engine()
idx = torch.arange(256, dtype=torch.int64, device="cuda")
kw = dict(num_samples=S, annealing=1.0, alpha=float(m.alpha), scale=N / 256.0)
def dump(which):   # needs a -DSCDEF_TC_TIMING build: where each role of the pair kernel spends its cycles
    import ctypes
    from scdef_b200 import _lib
    lib = _lib.load()
    buf = (ctypes.c_ulonglong * 128)()
    lib.scdef_debug_pair_timing.argtypes = [ctypes.c_void_p]
    lib.scdef_debug_pair_timing.restype = None
    lib.scdef_debug_pair_timing(buf)
    a = np.array(buf[:]).reshape(2, 4, 16)
    names = {0: ["w_empty", "z_empty"], 1: ["w_full", "pw_full", "r_empty", "q_full(net)", "issue MMA1", "issue MMA2", "z_full", "dz_empty"],
             2: ["r_full", "q_empty", "-", "tmem_ld+arrive", "math", "store+fence+arrive", "-", "dz_full", "drain/fold", "fold:issue W loads", "fold:tmem ld", "fold:math", "fold:tmem st"]}
    names[3] = names[2]
    print(f"  --- {which} kernel, cycles per role (one pair in the middle of the grid)", flush=True)
    for rank in range(2):
        for role, rn in enumerate(["COPY", "MMA/RELAY", "EPI warp0", "EPI warp15"]):
            print(f"  rank {rank} {rn:10s} total {int(a[rank, role, 15]):8d}  " +
                  " ".join(f"{n}={int(a[rank, role, i])}" for i, n in enumerate(names[role])), flush=True)


out = {}
for which in ("local", "global"):
    for rep in range(2):   # first round = warm-up
        eng.profile(True, ["lik_main", "lik_main_global"])
        eng.profile_read(reset=True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(20):
            eng.elbo_grad(which, idx, NoiseSpec.philox(0, 100 + i), want_loss=(which == "global"), **kw)
        ev1.record()
        torch.cuda.synchronize()
        prof, _ = eng.profile_read(reset=True)
        pk = prof["lik_main_global" if which == "global" else "lik_main"]
        out[which] = (pk[0] / max(pk[1], 1), ev0.elapsed_time(ev1) / 20)
    if len(sys.argv) > 3 and sys.argv[3] == "dump":
        dump(which)
    eng.profile(False)
mode = "monolithic kernels" if os.environ.get("SCDEF_TC_MONO") else "CTA-pair kernel"
print(f"[{mode}] S={S}  main kernel ms/launch: LOCAL {out['local'][0]:.4f}  GLOBAL {out['global'][0]:.4f}   "
      f"whole pass ms: LOCAL {out['local'][1]:.4f}  GLOBAL {out['global'][1]:.4f}", flush=True)
