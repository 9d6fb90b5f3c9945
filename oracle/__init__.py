"""CPU oracle for the scDEF training hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's per-minibatch
stochastic ELBO + gradient, its Adam/clip update and the minibatch stream
(`/root/reference/src/scdef/models/_scdef.py:1823-2181, 2808-2856, 3034-3173, 3215-3271`,
`/root/reference/src/scdef/utils/jax_utils.py:6-46`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / reported CPU baseline.  Nothing under
`scdef_b200/` imports it; the product path fails loudly when the CUDA library is missing.

PARITY PIN: the reference is pure Python on top of jax / tensorflow-probability / optax, none of
which is importable in this image (no network, not in the wheelhouse), and its own tests hold no golden
ELBO / gradient values for this path.  The oracle is pinned to the reference's OWN SOURCE
instead: `oracle/ref_exec.py` imports the unmodified `/root/reference/src/scdef/{utils/jax_utils.py,
models/_scdef.py, _sscdef.py, _iscdef.py}` on top of a torch-backed stand-in for the third-party modules
(`oracle/jaxshim.py`) and runs the constructor, `local_loss_grad` / `global_loss_grad` and the whole
`_learn` loop; `tests/test_ref_pin.py` requires `elbo_ref`, `elbo_closed` and `host_ref` to reproduce those
executions to 1e-11 (float64, same counts, same parameters, the very draws the reference consumed) over
every flag combination, sscDEF and iscDEF included; `tests/golden/ref_*.npz` (written by
`tests/golden/make_golden_ref.py` from the same executions) carry the reference's outputs to the GPU box.
What remains restated rather than executed is the THIRD-PARTY arithmetic inside the stand-in
(tfd.LogNormal.sample/entropy, tfd.Gamma.log_prob, jax.scipy.stats.poisson.logpmf, jnp.clip's
sub-gradient, optax.adam), from the published definitions of tensorflow-probability ^0.25,
jax >=0.4.31,<0.7 and optax ^0.2.3 (reference pyproject.toml:12-28), each checked against scipy /
torch.distributions / torch.optim.Adam in `tests/test_oracle.py`; jax's threefry random stream is not
reproduced (noise is injected).  Further self-checks: two independent restatements that must agree
(`elbo_ref`: line-by-line torch autograd; `elbo_closed`: hand-derived closed-form gradients), central
finite differences, and the integer invariants the reference's tests hold (geometric layer ladder,
minibatch permutation = numpy RandomState(0)).
"""
