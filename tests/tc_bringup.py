"""Bring-up of the tcgen05 likelihood kernel: TC vs SIMT vs the oracle, plus a dump of the rate tile R.
Run on the GPU box:  timeout 300 python tests/tc_bringup.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import elbo_closed  # noqa: E402
from oracle.elbo_ref import Noise  # noqa: E402
from scdef_b200 import _lib  # noqa: E402
from tests.helpers import make_engine, noise_to_spec, rel_err, small_problem  # noqa: E402


def run(N, G, Ks, nb, B, S, sg=None, dump=True, **kw):
    sg = sg or [0] * len(Ks)
    X, batch, model, lp, gp = small_problem(N, G, Ks, nb, seed=0, **kw)
    eng = make_engine(X, batch, model, lp, gp)
    idx = np.random.RandomState(0).permutation(N)[:B]
    noise = Noise.draw(S, B, nb, G, Ks, seed=7, dtype=torch.float32)
    ref_loss, ref_gl, ref_gg = elbo_closed.loss_and_grads(model, X[idx], idx, lp, gp, noise, 1.3, sg, 0, 0, model.alpha)
    idx_d = torch.as_tensor(idx, dtype=torch.int64, device="cuda")
    spec = noise_to_spec(noise)
    args = dict(num_samples=S, annealing=1.3, stop_gradients=sg, alpha=model.alpha)
    lib = _lib.load()
    lib.scdef_debug_tc_dump.argtypes = [ctypes.c_void_p]
    lib.scdef_debug_tc_dump.restype = None
    Rd = torch.zeros(S, B, G, device="cuda")
    if dump:
        lib.scdef_debug_tc_dump(ctypes.c_void_p(Rd.data_ptr()))
    out = {}
    for name, impl in (("tc", 2), ("simt", 1)):
        eng.elbo_grad("local", idx_d, spec, lik_impl=impl, **args)
        torch.cuda.synchronize()
        ll = eng.loss_value()
        gl = [g.cpu().numpy().copy() for g in eng.local_grads(B)]
        eng.elbo_grad("global", idx_d, spec, lik_impl=impl, **args)
        torch.cuda.synchronize()
        lg = eng.loss_value()
        gg = [g.cpu().numpy().copy() for g in eng.glob_grad.views]
        out[name] = (ll, lg, gl, gg)
    lib.scdef_debug_tc_dump(None)
    if dump:
        mu = np.clip(lp[1][0][idx][:, :Ks[0]], np.log(1e-10), np.log(1e6))
        sig = np.exp(np.clip(lp[1][1][idx][:, :Ks[0]], np.log(1e-10), np.log(1e10)))
        z = np.exp(mu[None] + sig[None] * noise.z[:, :, :Ks[0]].numpy())
        wm = np.clip(gp[2][0], np.log(1e-10), np.log(1e6))
        ws = np.exp(np.clip(gp[2][1], np.log(1e-10), np.log(1e10)))
        W = np.exp(wm[None] + ws[None] * noise.W[0].numpy())
        Rref = np.einsum("sbk,skg->sbg", z.astype(np.float64), W.astype(np.float64))
        print("  R dump rel err:", rel_err(Rd.cpu().numpy(), Rref), " max|R|", np.abs(Rref).max(), flush=True)
    names_g = ["gene_scale", "brd"] + [f"W{l}" for l in range(len(Ks))] + ["ard", "alpha"]
    for name in ("tc", "simt"):
        ll, lg, gl, gg = out[name]
        errs = {"loss_l": abs(ll - ref_loss) / abs(ref_loss), "loss_g": abs(lg - ref_loss) / abs(ref_loss)}
        for n_, g, r in zip(("c", "z"), gl, ref_gl):
            errs[n_] = rel_err(g, r[:, idx]) if np.abs(r).max() > 0 else float(np.abs(g).max())
        for n_, g, r in zip(names_g, gg, ref_gg):
            errs[n_] = rel_err(g, r) if np.abs(r).max() > 0 else float(np.abs(g).max())
        print(f"  {name}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)
        out[name + "_max_err"] = max(errs.values())
    return out


def pair_cases():
    """The CTA-pair kernel (both passes; SCDEF_TC_MONO=1: the monolithic kernels).  On a trapped launch the record of the
    wait that timed out is printed: (code, blockIdx.x, warp or sub-unit, parity) -- codes in lik_flat.cuh."""
    worst = 0.0
    try:
        print("pair 1: one pair-tile, KP=16", flush=True); worst = max(worst, run(300, 128, (16, 4), 1, 256, 1)["tc_max_err"])
        print("pair 2: C5-like layers, phantom tile", flush=True); worst = max(worst, run(600, 312, (100, 60, 30, 10), 1, 256, 2)["tc_max_err"])
        print("pair 3: batches, two cell blocks", flush=True); worst = max(worst, run(700, 200, (24, 6, 1), 3, 512, 3)["tc_max_err"])
        # 100 samples x 2 pair-tiles over 74 pairs: chunks of 2-3 pair-tiles that cross sample boundaries (persistent variant)
        print("pair 4: many samples, chunks cross samples", flush=True)
        worst = max(worst, run(400, 200, (24, 6, 1), 1, 256, 100, dump=False)["tc_max_err"])
        print(f"pair: worst relative error of the tcgen05 path {worst:.2e} (bar 1e-4)", flush=True)
    finally:
        lib = _lib.load()
        rec = (ctypes.c_int * 4)()
        lib.scdef_debug_pair_stuck.argtypes = [ctypes.c_void_p]
        lib.scdef_debug_pair_stuck.restype = None
        lib.scdef_debug_pair_stuck(rec)
        print("pair stuck record (code, block, warp/sub-unit, parity):", list(rec), flush=True)
    return worst


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "pair":
        sys.exit(0 if pair_cases() <= 1e-4 else 1)
    print("case 1: tiny", flush=True); run(64, 64, (16, 4), 1, 32, 1)
    print("case 2: C1-like", flush=True); run(200, 120, (10, 5, 2), 1, 200, 3)
    print("case 3: ragged", flush=True); run(300, 200, (100, 60, 30, 10), 1, 257, 2)
    print("case 4: batches", flush=True); run(160, 72, (24, 6, 1), 3, 130, 2)
    print("case 5: no brd, stop", flush=True); run(100, 64, (12, 4), 1, 50, 2, sg=[0, 1], use_brd=False)
    print("case 6: lik off", flush=True); run(100, 64, (12, 4), 1, 50, 2, sg=[1, 0], dump=False)
