"""Prints where each warp role of the pipelined tcgen05 kernel waits (needs a build with
SCDEF_EXTRA_FLAGS=-DSCDEF_TC_TIMING).  C5 shapes, one train step."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import MiniAnnData, scDEF, _lib
from scdef_b200.synth import synth_counts_torch
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
X, _ = synth_counts_torch(N, 5000, lib=3000.0, device="cuda")
m = scDEF(MiniAnnData(X), layer_sizes=[100, 60, 30, 10], logginglevel=30)
m.max_steps_per_learn = 4
m._learn(n_epoch=1, num_samples=100, batch_size=256)
lib = _lib.load()
out = (ctypes.c_ulonglong * 128)()
lib.scdef_debug_tc_timing(out)
a = np.array(out[:]).reshape(2, 4, 16)
names = {0: ["P: w_empty"], 1: ["MMA: w_full", "z_full", "r_empty", "q_full", "dw_empty"], 2: ["E: r_full", "q_empty", "dw_full", "t_tmem_ld", "t_math", "t_store+fence", "t_top(loads issue)", "t_wfull"], 3: ["COPY: z_empty"]}
for v, vn in enumerate(["LOCAL", "GLOBAL"]):
    print(vn)
    for role in range(4):
        tot = a[v, role, 15]
        print("   total cycles", tot, " waits:", {n: int(a[v, role, i]) for i, n in enumerate(names[role])})
