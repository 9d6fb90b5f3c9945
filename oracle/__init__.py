"""CPU oracle for the scDEF training hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's per-minibatch
stochastic ELBO + gradient, its Adam/clip update and the minibatch stream
(`/root/reference/src/scdef/models/_scdef.py:1823-2181, 2808-2856, 3034-3173, 3215-3271`,
`/root/reference/src/scdef/utils/jax_utils.py:6-46`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / reported CPU baseline.  Nothing under
`scdef_b200/` imports it; the product path fails loudly when the CUDA library is missing.

PARITY UNPINNED: the reference is pure Python on top of jax / tensorflow-probability / optax,
none of which is importable in this image (no network, not in the wheelhouse), and the
reference's own tests hold no golden ELBO / gradient values for this path.
The oracle is therefore pinned only by (i) two independent restatements that must agree
(`elbo_ref`: line-by-line torch autograd; `elbo_closed`: hand-derived closed-form gradients),
(ii) central finite differences, and (iii) the integer invariants the reference's tests do
hold (geometric layer ladder values, minibatch permutation = numpy RandomState(0)).
The third-party formulas restated (tfd.LogNormal.sample/entropy, tfd.Gamma.log_prob,
jax.scipy.stats.poisson.logpmf, optax.adam) are from their published definitions:
tensorflow-probability ^0.25, jax >=0.4.31,<0.7, optax ^0.2.3 (reference pyproject.toml:12-28).
"""
