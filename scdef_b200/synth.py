"""Synthetic Poisson count matrices of the BASELINE.json shapes.

K*=20 latent programs W* ~ Gamma(0.3,1) row-normalised; one dominant program per cell
(0.9*onehot + 0.1*Dirichlet(1)); library size LogNormal(ln lib, 0.5); optional batches with
per-batch gene scale LogNormal(0, 0.3); X ~ Poisson(rate).  NumPy version for tests / CPU
baselines, torch version (chunked, on device) for the large bench shapes.
"""
from __future__ import annotations

import math

import numpy as np

CONFIGS = {
    # name: (n_cells, n_genes, layer_sizes, n_batches, median library size)
    "C1": (200, 120, [10, 5, 2], 1, 1500.0),
    "C2": (2700, 2000, [100, 60, 30, 10], 1, 1500.0),
    "C3": (100_000, 3000, [48, 24, 1], 1, 3000.0),
    "C4": (500_000, 5000, [100, 60, 30, 10], 8, 3000.0),
    "C5": (1_000_000, 5000, [100, 60, 30, 10], 1, 3000.0),
}


def synth_counts_numpy(n_cells, n_genes, n_batches=1, lib=1500.0, n_programs=20, seed=0):
    rng = np.random.default_rng(seed)
    Wst = rng.gamma(0.3, 1.0, size=(n_programs, n_genes)) + 1e-8
    Wst /= Wst.sum(1, keepdims=True)
    t = rng.integers(0, n_programs, size=n_cells)
    zst = 0.1 * rng.dirichlet(np.ones(n_programs), size=n_cells)
    zst[np.arange(n_cells), t] += 0.9
    ell = np.exp(rng.normal(math.log(lib), 0.5, size=n_cells))
    batch = rng.integers(0, n_batches, size=n_cells) if n_batches > 1 else np.zeros(n_cells, dtype=np.int64)
    rate = ell[:, None] * (zst @ Wst)
    if n_batches > 1:
        gst = np.exp(rng.normal(0.0, 0.3, size=(n_batches, n_genes)))
        rate = rate * gst[batch]
    X = rng.poisson(rate).astype(np.float32)
    return X, batch.astype(np.int32)


def synth_counts_torch(n_cells, n_genes, n_batches=1, lib=3000.0, n_programs=20, seed=0,
                       device="cuda", chunk=65536, out_dtype=None):
    """Same generative model on the device, chunked over cells so the (N,G) rate matrix is
    never materialised.  Returns (X float32 (N,G), batch int32 (N,))."""
    import torch

    g = torch.Generator(device=device).manual_seed(int(seed))
    out_dtype = out_dtype or torch.float32
    conc = torch.full((n_programs, n_genes), 0.3, device=device)
    Wst = torch._standard_gamma(conc, generator=g) + 1e-8
    Wst /= Wst.sum(1, keepdim=True)
    gst = None
    if n_batches > 1:
        gst = torch.exp(0.3 * torch.randn(n_batches, n_genes, device=device, generator=g))
    X = torch.empty(n_cells, n_genes, dtype=out_dtype, device=device)
    batch = torch.zeros(n_cells, dtype=torch.int32, device=device)
    for s in range(0, n_cells, chunk):
        n = min(chunk, n_cells - s)
        t = torch.randint(0, n_programs, (n,), device=device, generator=g)
        d = torch._standard_gamma(torch.ones(n, n_programs, device=device), generator=g)
        zst = 0.1 * d / d.sum(1, keepdim=True)
        zst[torch.arange(n, device=device), t] += 0.9
        ell = torch.exp(math.log(lib) + 0.5 * torch.randn(n, device=device, generator=g))
        rate = ell[:, None] * (zst @ Wst)
        if gst is not None:
            b = torch.randint(0, n_batches, (n,), device=device, generator=g)
            batch[s : s + n] = b.to(torch.int32)
            rate = rate * gst[b]
        X[s : s + n] = torch.poisson(rate, generator=g).to(out_dtype)
    return X, batch
