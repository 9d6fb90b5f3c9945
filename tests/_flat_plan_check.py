"""Run by tests/test_host.py in a subprocess with SCDEF_TC_PAIR=2 (the switch is read once per process): checks the
work decomposition of the persistent CTA-pair kernel (`lik_tc_pair.cuh`: chunks of the flattened (sample, pair-tile)
space, partial-buffer slots) for random shapes, through the library's own arithmetic.  No CUDA call is made."""
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
    keys = ["KP", "n_tiles", "n_chunks", "NG", "tiles_per_range", "n_cblocks", "NGp", "ptiles_per_range", "Pflat", "mode"]
    return dict(zip(keys, out))


rng = np.random.default_rng(0)
shapes = [(256, 100, 100, 5000), (256, 1, 16, 128), (512, 3, 24, 200), (256, 100, 24, 200), (1024, 7, 100, 3000)]
shapes += [(256 * int(rng.integers(1, 5)), int(rng.integers(1, 300)), int(rng.integers(1, 113)), 8 * int(rng.integers(1, 800))) for _ in range(60)]
for B, S, K0, G in shapes:
    g = geom(B, S, K0, G)
    assert g["mode"] == 2
    assert g["n_tiles"] % 2 == 0 and g["n_tiles"] * 64 >= G and (g["n_tiles"] - 2) * 64 < G   # padded to whole pair-tiles
    n_pt, P = g["n_tiles"] // 2, g["Pflat"]
    T = S * n_pt
    assert 1 <= P <= max(1, 74 // g["n_cblocks"]) and P <= T
    bounds = [c * T // P for c in range(P + 1)]
    assert bounds[0] == 0 and bounds[-1] == T and all(b1 > b0 for b0, b1 in zip(bounds[:-1], bounds[1:])), "every pair has work"
    assert max(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:])) - min(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:])) <= 1
    # the kernel's slot of (chunk c, sample s) = c - chunk_of(first pair-tile of s): distinct per sample, below NGp;
    # the reduction reads exactly the slots that were written
    written = {}
    for c in range(P):
        for f in range(bounds[c], bounds[c + 1]):
            assert lib.scdef_debug_flat_chunk_of(f, T, P) == c
            s = f // n_pt
            slot = c - lib.scdef_debug_flat_chunk_of(s * n_pt, T, P)
            assert 0 <= slot < g["NGp"], (slot, g)
            written.setdefault(s, set()).add(slot)
    for s in range(S):
        nsl = lib.scdef_debug_flat_chunk_of((s + 1) * n_pt - 1, T, P) - lib.scdef_debug_flat_chunk_of(s * n_pt, T, P) + 1
        assert written[s] == set(range(nsl)), (s, written[s], nsl)
print("flat plan ok", len(shapes))
