"""Run by tests/test_host.py in a subprocess: checks the work decomposition of the persistent CTA-pair likelihood kernel
(`csrc/lik_flat.cuh`: chunks of the flattened (row block, sample, pair-tile) space, partial-buffer slots) for random
shapes and both orientations, through the library's own arithmetic.  No CUDA call is made."""
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scdef_b200 import _lib  # noqa: E402

lib = ctypes.CDLL(_lib.lib_path())
lib.scdef_debug_tc_geom.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p]
lib.scdef_debug_tc_geom.restype = None
lib.scdef_debug_flat_chunk_of.argtypes = [ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int]
lib.scdef_debug_flat_chunk_of.restype = ctypes.c_int


def geom(B, S, K0, G):
    out = (ctypes.c_int * 10)()
    lib.scdef_debug_tc_geom(B, S, K0, G, out)
    v = list(out)
    return dict(KP=v[0], n_tiles=v[1], local=dict(n_rblk=v[2], n_pt=v[3], P=v[4], slots=v[5]),
                glob=dict(n_rblk=v[6], n_pt=v[7], P=v[8], slots=v[9]))


rng = np.random.default_rng(0)
shapes = [(256, 100, 100, 5000), (256, 1, 16, 128), (512, 3, 24, 200), (256, 100, 24, 200), (1024, 7, 100, 3000), (200, 4, 10, 120),
          (64, 10, 100, 5000), (77, 2, 33, 72)]
shapes += [(int(rng.integers(1, 1200)), int(rng.integers(1, 300)), int(rng.integers(1, 113)), 8 * int(rng.integers(1, 800))) for _ in range(60)]
for B, S, K0, G in shapes:
    g = geom(B, S, K0, G)
    n_gb = -(-G // 256)
    assert g["KP"] % 16 == 0 and K0 <= g["KP"] < K0 + 16
    assert g["n_tiles"] == 4 * n_gb                               # W_0 tile images per sample: whole 256-gene blocks
    assert g["local"]["n_rblk"] == -(-B // 256) and g["local"]["n_pt"] == 2 * n_gb
    assert g["glob"]["n_rblk"] == n_gb and g["glob"]["n_pt"] == -(-B // 128)
    for name, by_segment in (("local", True), ("glob", False)):
        o = g[name]
        n_pt, P = o["n_pt"], o["P"]
        T = o["n_rblk"] * S * n_pt
        assert 1 <= P <= 74 and P <= T
        bounds = [c * T // P for c in range(P + 1)]
        assert bounds[0] == 0 and bounds[-1] == T and all(b1 > b0 for b0, b1 in zip(bounds[:-1], bounds[1:])), "every pair has work"
        assert max(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:])) - min(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:])) <= 1
        # a pair's partial result of a group (LOCAL: segment (row block, sample); GLOBAL: gene block) goes to slot
        # chunk - chunk_of(first pair-tile of the group): distinct per group, below `slots`; the reduction reads exactly
        # the slots that were written
        per_group = n_pt if by_segment else S * n_pt
        written = {}
        for c in range(P):
            for f in range(bounds[c], bounds[c + 1]):
                assert lib.scdef_debug_flat_chunk_of(f, T, P) == c
                grp = f // per_group
                slot = c - lib.scdef_debug_flat_chunk_of(grp * per_group, T, P)
                assert 0 <= slot < o["slots"], (name, slot, o)
                written.setdefault(grp, set()).add(slot)
        for grp in range(T // per_group):
            nsl = lib.scdef_debug_flat_chunk_of((grp + 1) * per_group - 1, T, P) - lib.scdef_debug_flat_chunk_of(grp * per_group, T, P) + 1
            assert written[grp] == set(range(nsl)), (name, grp, written[grp], nsl)
print("flat plan ok", len(shapes))
