"""Operand-precision experiment for the two likelihood GEMMs at the real per-step shape (256 cells x 5000 genes x 100 factors):
which tensor-core operand formats keep the loss and the gradients inside the 1e-4 parity bar?   python scripts/numerics_probe.py
Arithmetic: operands rounded to the format, products and sums in float64 (the tensor core accumulates in fp32; that error
is ~1e-7 and the same for every variant).  Gradients are the full ones (likelihood part of dz0, dW0 including the rank-1
terms), error metric = max|g - g_ref| / max|g_ref| as in the parity tests."""
import numpy as np

rng = np.random.default_rng(0)
B, G, K, S = 256, 5000, 100, 3


def rnd(x, bits):   # round-to-nearest to `bits` explicit mantissa bits
    x = np.asarray(x, np.float64)
    m, e = np.frexp(x)
    return np.ldexp(np.round(m * 2.0 ** (bits + 1)) / 2.0 ** (bits + 1), e)


def split(x, bits):
    hi = rnd(x, bits)
    return hi, rnd(x - hi, bits)


def products(A, Bm, mode):
    """A @ Bm with operands in the given format."""
    if mode == "fp64":
        return A @ Bm
    if mode == "tf32":
        return rnd(A, 10) @ rnd(Bm, 10)
    if mode == "bf16x3":
        ah, al = split(A, 7); bh, bl = split(Bm, 7)
        return al @ bh + ah @ bl + ah @ bh
    if mode == "bf16x2a":   # A split, B hi only
        ah, al = split(A, 7); bh = rnd(Bm, 7)
        return al @ bh + ah @ bh
    if mode == "bf16x2b":   # B split, A hi only
        ah = rnd(A, 7); bh, bl = split(Bm, 7)
        return ah @ bl + ah @ bh
    if mode == "tf32x2a":
        ah, al = split(A, 10); bh = rnd(Bm, 10)
        return al @ bh + ah @ bh
    if mode == "tf32x3":
        ah, al = split(A, 10); bh, bl = split(Bm, 10)
        return al @ bh + ah @ bl + ah @ bh
    raise ValueError(mode)


def run(mode1, mode2, seed):
    r = np.random.default_rng(seed)
    z = np.exp(r.normal(-1.0, 1.0, (B, K)))
    W = np.exp(r.normal(-2.0, 1.5, (K, G)))
    c = np.exp(r.normal(0, 0.3, (B, 1))); s = np.exp(r.normal(0, 0.3, (1, G)))
    lam = c * s * (z @ W)
    x = r.poisson(np.minimum(lam * 3000.0 / lam.sum(1, keepdims=True), 50.0)).astype(np.float64)   # ~3000 counts per cell

    def grads(m1, m2):
        R = products(z, W, m1)
        Q = np.where(x > 0, x / R, 0.0)
        ll = np.sum(x * np.log(np.maximum(R, 1e-300)))
        Ws = (W * s).sum(1)                      # rank-1 terms: exact fp32 SIMT work in the kernels
        dz = products(Q, W.T, m2) - c * Ws[None, :]
        dW = products(z.T, Q, m2) - (c * z).sum(0)[:, None] * s
        return ll, dz * z, dW * W                # d/d mu of the LogNormal blocks: times the sample

    ref = grads("fp64", "fp64")
    got = grads(mode1, mode2)
    return (abs(got[0] - ref[0]) / abs(ref[0]),
            np.abs(got[1] - ref[1]).max() / np.abs(ref[1]).max(),
            np.abs(got[2] - ref[2]).max() / np.abs(ref[2]).max())


def oracle_run(modes, Ks=(100, 60, 30, 10), G=5000, nb=1, B=256, S=3):
    """The same question through the oracle's closed-form gradients on the parity tests' own problem generator at the C5
    step shape: errors of the FINAL (mu, rho) gradients of every block, as `tests/test_parity_shapes_gpu.py` measures them."""
    import os, sys
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import elbo_closed
    from oracle.elbo_ref import Noise
    from tests.helpers import rel_err, small_problem
    X, batch, model, lp, gp = small_problem(640, G, Ks, nb, seed=21)
    idx = np.random.RandomState(3).permutation(640)[:B]
    noise = Noise.draw(S, B, nb, G, Ks, seed=5, dtype=torch.float32, couple_like_reference=True)
    args = (model, X[idx], idx, lp, gp, noise, 1.2, [0] * len(Ks), 0.0, 0.0, model.alpha)
    elbo_closed.LIK_GEMM = None
    ref = elbo_closed.loss_and_grads(*args)
    names = ["cell_scale", "z", "gene_scale", "brd"] + [f"W{l}" for l in range(len(Ks))] + ["ard", "alpha"]
    for m1, m2 in modes:
        elbo_closed.LIK_GEMM = lambda A, Bm, which: products(A, Bm, m1 if which == "rate" else m2)
        got = elbo_closed.loss_and_grads(*args)
        errs = {"loss": abs(got[0] - ref[0]) / abs(ref[0])}
        for n, g, r in zip(names, [x[:, idx] for x in got[1]] + list(got[2]), [x[:, idx] for x in ref[1]] + list(ref[2])):
            if np.abs(r).max() > 0:
                errs[n] = rel_err(g, r)
        worst = max(errs, key=errs.get)
        print(f"{m1:8s} / {m2:8s}   loss {errs['loss']:.1e}  z {errs['z']:.1e}  c {errs['cell_scale']:.1e}  s {errs['gene_scale']:.1e}  W0 {errs['W0']:.1e}"
              f"   worst {worst} {errs[worst]:.1e}   {'ok' if errs[worst] < 1e-4 else ('marginal' if errs[worst] < 3e-4 else 'FAIL')}", flush=True)
    elbo_closed.LIK_GEMM = None


if __name__ == "__main__":
    print(f"# B={B} G={G} K0={K}; worst over {S} draws;  rate GEMM / second GEMMs ->  loss, dz0 (mu), dW0 (mu) relative errors")
    for m1, m2 in [("bf16x3", "bf16x3"), ("tf32", "tf32"), ("tf32", "bf16x3"), ("bf16x3", "tf32"), ("tf32x2a", "tf32x2a"), ("tf32x3", "tf32x3"),
                   ("bf16x2a", "bf16x3"), ("bf16x2b", "bf16x3"), ("bf16x3", "bf16x2a"), ("bf16x3", "bf16x2b")]:
        e = np.max([run(m1, m2, sd) for sd in range(S)], axis=0)
        print(f"{m1:8s} / {m2:8s}   loss {e[0]:.1e}   dz0 {e[1]:.1e}   dW0 {e[2]:.1e}   {'ok' if max(e) < 1e-4 else ('marginal' if max(e) < 3e-4 else 'FAIL')}")
    print("# the parity tests' problem (tests/helpers.small_problem, C5 step shape), final (mu, rho) gradients through oracle/elbo_closed.py")
    oracle_run([("bf16x3", "bf16x3"), ("tf32", "tf32"), ("tf32", "bf16x3"), ("bf16x3", "tf32"), ("tf32x2a", "tf32x2a"),
                ("bf16x2a", "bf16x3"), ("bf16x2b", "bf16x3"), ("bf16x3", "bf16x2a"), ("bf16x3", "bf16x2b")])
