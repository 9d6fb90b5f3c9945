"""A torch-backed stand-in for the slice of jax / tensorflow-probability / optax that the
reference's hot path uses, so that the UNMODIFIED reference source can be executed here.

TEST INFRASTRUCTURE (see oracle/__init__.py): nothing under `scdef_b200/` may import this.

Why: `jax`, `optax` and `tensorflow_probability` are not installable in this image, and the
reference's tests hold no known-answer vectors for the hot path.  The restated
oracle (`oracle/elbo_ref.py`, `oracle/host_ref.py`) was therefore pinned only by a human reading.
`oracle/ref_exec.py` loads the reference's own `.py` files from `/root/reference` and runs them on top
of the modules built here; `oracle/elbo_ref.py` must reproduce that execution to 1e-12 in float64
(`tests/test_ref_pin.py`), and `tests/golden/ref_*.npz` (made by `tests/golden/make_golden_ref.py`)
carry the reference's outputs to the GPU box where `/root/reference` does not exist.

What is restated here is THIRD-PARTY arithmetic, each item from the library's published definition
(pins: `pyproject.toml:12-28` of the reference — jax >=0.4.31,<0.7, tensorflow-probability ^0.25,
optax ^0.2.3):
  * `jnp.clip(x, lo, hi)` = `minimum(maximum(x, lo), hi)` — ties pass half the gradient (torch's
    `maximum`/`minimum` do the same);
  * `jax.lax.cond(pred, t, f)` with a concrete predicate = `t() if pred else f()`;
    `jax.lax.stop_gradient` = `.detach()`; `jit` = identity; `vmap` = `torch.func.vmap`
    (a Python loop + stack when a mapped argument is a PRNG key, because keys must be concrete);
  * `value_and_grad(f, argnums)` = torch reverse-mode autograd over the leaves of that argument;
  * `tfd.LogNormal(loc, scale)`: `sample = exp(loc + scale * N(0,1))`,
    `entropy = loc + log(scale) + 1/2 + 1/2 log(2 pi)`, `log_prob`;
    `tfd.Gamma(concentration, rate).log_prob(x) = xlogy(a-1, x) - b x - lgamma(a) + a log(b)`;
  * `jax.scipy.stats.poisson.logpmf(k, mu) = xlogy(k, mu) - gammaln(k+1) - mu` (-inf for k < 0);
  * `optax.adam(lr, b1=.9, b2=.999, eps=1e-8, eps_root=0)` = `scale_by_adam` (bias-corrected first and
    second moments, `m_hat / (sqrt(v_hat + eps_root) + eps)`) then `scale(-lr)`; `apply_updates = p + u`.
Random numbers: jax's threefry stream is NOT reproduced (bit-reproducing it is not a goal).
A key is a pair of integers; `split` / `fold_in` derive children with `numpy.random.SeedSequence`, and
`normal(key, shape)` is `numpy.random.default_rng(key).standard_normal(shape)` — a deterministic
function of (key, shape), so the reference's "same key, same shape -> same draw" coupling happens by
itself, and `normal_like_reference(key, shape)` lets a test build exactly the draws the reference's
`lognormal_sample(rng, ...)` consumed.
"""
from __future__ import annotations

import contextlib
import math
import sys
import types
from typing import Any, Callable, Sequence

import numpy as np
import torch

# the float type of "jnp.float32" / default jnp floats: float64 to pin the float64 oracle,
# float32 to imitate the reference's precision
_STATE = {"dtype": torch.float64, "eps_f32": False}


def set_noise_float32(flag: bool):
    """Round the N(0,1) draws to float32 values (still carried in the float dtype): a float32 CUDA path
    can then be fed bit-identical noise.  Used when the golden vectors are made."""
    _STATE["eps_f32"] = bool(flag)


def set_float_dtype(dtype):
    _STATE["dtype"] = dtype


def float_dtype():
    return _STATE["dtype"]


# --------------------------------------------------------------------------------------------
# array type

def _conv(x):
    """numpy values entering a torch op."""
    if isinstance(x, np.ndarray):
        t = torch.as_tensor(x)
        if t.dtype in (torch.float32, torch.float64, torch.float16):
            t = t.to(_STATE["dtype"])
        return t
    if isinstance(x, np.generic):
        return x.item()
    if isinstance(x, (list, tuple)) and any(isinstance(v, (np.ndarray, np.generic)) for v in x):
        return type(x)(_conv(v) for v in x)
    return x


class _At:
    """`x.at[idx].set(v)` / `.add(v)` (functional updates)."""

    def __init__(self, arr):
        self.arr = arr
        self.idx = None

    def __getitem__(self, idx):
        out = _At(self.arr)
        out.idx = idx
        return out

    def _idx(self):
        idx = self.idx
        if isinstance(idx, tuple):
            return tuple(_conv(i) for i in idx)
        if isinstance(idx, list):
            return torch.as_tensor(idx, dtype=torch.long)
        return _conv(idx)

    def set(self, v):
        out = self.arr.clone()
        v = _conv(v)
        if isinstance(v, torch.Tensor):
            v = v.to(out.dtype)
        out[self._idx()] = v
        return out

    def add(self, v):
        out = self.arr.clone()
        out[self._idx()] = out[self._idx()] + _conv(v)
        return out


class Array(torch.Tensor):
    """torch.Tensor with the handful of jax.Array methods the reference calls."""

    __array_priority__ = 2000

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = tuple(_conv(a) for a in args)
        kwargs = {k: _conv(v) for k, v in kwargs.items()}
        return super().__torch_function__(func, types, args, kwargs)

    def __getitem__(self, idx):
        # numpy / jax allow negative slice steps (`x.argsort()[::-1][:k]`, _scdef.py:3989); torch does not
        items = idx if isinstance(idx, tuple) else (idx,)
        if any(isinstance(i, slice) and i.step is not None and i.step < 0 for i in items):
            out, dim = self, 0
            for i in items:
                if i is None:
                    out = torch.Tensor.unsqueeze(out, dim)
                elif isinstance(i, slice) and i.step is not None and i.step < 0:
                    sel = torch.as_tensor(list(range(*i.indices(out.shape[dim]))), dtype=torch.long)
                    out = torch.Tensor.index_select(out, dim, sel)
                elif isinstance(i, slice):
                    out = torch.Tensor.__getitem__(out, (slice(None),) * dim + (i,))
                else:
                    raise IndexError("jaxshim: a negative slice step mixed with non-slice indices is not supported")
                dim += 1
            return out
        return super().__getitem__(idx)

    # numpy_array <op> Array lands here because of __array_priority__
    def __rmul__(self, o):
        return torch.mul(self, _conv(o))

    def __radd__(self, o):
        return torch.add(self, _conv(o))

    def __rsub__(self, o):
        return torch.sub(_as_array(o), self)

    def __rtruediv__(self, o):
        return torch.div(_as_array(o), self)

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __array_wrap__(self, array, context=None, return_scalar=False):
        # np.exp(jax_array) returns a NumPy value in jax; same here
        return array[()] if return_scalar else array

    def dot(self, other):  # jax: matrix product for >= 1-D operands
        return torch.matmul(self, _as_array(other))

    def astype(self, dtype):
        return self.to(_dtype(dtype))

    @property
    def at(self):
        return _At(self)

    def block_until_ready(self):
        return self

    def copy(self):
        return self.clone()

    def tolist(self):
        return self.detach().as_subclass(torch.Tensor).tolist()

    # numpy-style reductions (np.mean(x) dispatches to x.mean(axis=None, dtype=None, out=None))
    def _np_reduce(self, fn, args, kwargs):
        axis = kwargs.pop("axis", kwargs.pop("dim", args[0] if args else None))
        keep = kwargs.pop("keepdims", kwargs.pop("keepdim", False))
        kwargs.pop("out", None)
        kwargs.pop("dtype", None)
        kwargs.pop("where", None)
        t = self.as_subclass(torch.Tensor)
        if axis is None:
            out = fn(t)
            if keep:
                out = out.reshape((1,) * t.dim())
        else:
            out = fn(t, dim=axis, keepdim=bool(keep))
        return _wrap(out)

    def mean(self, *a, **k):
        return self._np_reduce(torch.mean, a, k)

    def sum(self, *a, **k):
        return self._np_reduce(torch.sum, a, k)

    def max(self, *a, **k):
        return self._np_reduce(torch.amax, a, k)

    def min(self, *a, **k):
        return self._np_reduce(torch.amin, a, k)

    def std(self, *a, ddof=0, **k):
        return self._np_reduce(lambda t, **kw: torch.std(t, correction=ddof, **kw), a, k)

    def var(self, *a, ddof=0, **k):
        return self._np_reduce(lambda t, **kw: torch.var(t, correction=ddof, **kw), a, k)


class KeyArray(Array):
    """PRNG key(s): int64 tensor (..., 2)."""


def _wrap(t):
    if isinstance(t, torch.Tensor) and not isinstance(t, Array):
        return t.as_subclass(Array)
    return t


def _dtype(d):
    if d is None:
        return None
    if isinstance(d, torch.dtype):
        return d
    if d in (float, "float", "float32", "float64", np.float32, np.float64, _F32, _F64):
        return _STATE["dtype"]
    if d in (int, "int", "int32", "int64", np.int32, np.int64, _I32, _I64):
        return torch.int64
    if d in (bool, "bool", np.bool_, _BOOL):
        return torch.bool
    nd = np.dtype(d)
    if nd.kind == "f":
        return _STATE["dtype"]
    if nd.kind in "iu":
        return torch.int64
    if nd.kind == "b":
        return torch.bool
    raise TypeError(f"jaxshim: unsupported dtype {d!r}")


class _DType:
    """Callable dtype token, like `jnp.float32` (usable as dtype= and as a cast)."""

    def __init__(self, name):
        self.name = name

    def __call__(self, x):
        return _as_array(x, dtype=self)

    def __repr__(self):
        return f"jaxshim.{self.name}"


_F32, _F64, _I32, _I64, _BOOL = (_DType(n) for n in ("float32", "float64", "int32", "int64", "bool_"))


def _as_array(x, dtype=None):
    want = _dtype(dtype)
    if isinstance(x, torch.Tensor):
        t = x
    else:
        if isinstance(x, (list, tuple)) and len(x) and any(isinstance(v, torch.Tensor) for v in x):
            t = torch.stack([_as_array(v) for v in x])
        else:
            a = np.asarray(x)
            if a.dtype == object:
                raise TypeError("jaxshim: cannot convert an object array")
            t = torch.as_tensor(a)
        if t.dtype.is_floating_point:
            t = t.to(_STATE["dtype"])
        elif t.dtype in (torch.int8, torch.int16, torch.int32, torch.uint8):
            t = t.to(torch.int64)
    if want is not None and t.dtype != want:
        t = t.to(want)
    return _wrap(t)


def _unwrap(x):
    return x.as_subclass(torch.Tensor) if isinstance(x, Array) else x


# --------------------------------------------------------------------------------------------
# jax.numpy

def _build_numpy():
    m = types.ModuleType("jax.numpy")
    m.float32, m.float64, m.int32, m.int64, m.bool_ = _F32, _F64, _I32, _I64, _BOOL
    m.float_, m.int_ = _F64, _I64
    m.inf, m.pi, m.nan, m.newaxis = math.inf, math.pi, math.nan, None
    m.ndarray = Array

    def array(x, dtype=None, copy=True):
        out = _as_array(x, dtype)
        return out.clone() if isinstance(x, torch.Tensor) else out

    m.array = array
    m.asarray = lambda x, dtype=None: _as_array(x, dtype)

    def _unary(fn):
        return lambda x: _wrap(fn(_as_array(x)))

    for name in ("log", "exp", "sqrt", "abs", "log1p", "expm1", "sign", "square", "floor", "ceil",
                 "isnan", "isfinite", "logical_not", "tanh"):
        setattr(m, name, _unary(getattr(torch, name)))
    m.rint = _unary(torch.round)

    def _binary(fn):
        def f(a, b):
            a, b = _as_array(a), _as_array(b)
            if a.dtype != b.dtype and (a.dtype.is_floating_point or b.dtype.is_floating_point):
                a, b = a.to(_STATE["dtype"]), b.to(_STATE["dtype"])
            return _wrap(fn(a, b))
        return f

    m.maximum = _binary(torch.maximum)
    m.minimum = _binary(torch.minimum)
    m.logical_and = _binary(torch.logical_and)
    m.logical_or = _binary(torch.logical_or)
    m.power = _binary(torch.pow)
    m.multiply = _binary(torch.mul)
    m.add = _binary(torch.add)

    def clip(x, a_min=None, a_max=None, *, min=None, max=None):
        lo = a_min if a_min is not None else min
        hi = a_max if a_max is not None else max
        out = _as_array(x)
        if lo is not None:
            out = m.maximum(out, lo)
        if hi is not None:
            out = m.minimum(out, hi)
        return out

    m.clip = clip

    def _shape(shape):
        return (shape,) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)

    m.ones = lambda shape, dtype=None: _wrap(torch.ones(_shape(shape), dtype=_dtype(dtype) or _STATE["dtype"]))
    m.zeros = lambda shape, dtype=None: _wrap(torch.zeros(_shape(shape), dtype=_dtype(dtype) or _STATE["dtype"]))
    m.full = lambda shape, v, dtype=None: _wrap(torch.full(_shape(shape), float(v), dtype=_dtype(dtype) or _STATE["dtype"]))
    m.ones_like = lambda x, dtype=None: _wrap(torch.ones_like(_as_array(x), dtype=_dtype(dtype)))
    m.zeros_like = lambda x, dtype=None: _wrap(torch.zeros_like(_as_array(x), dtype=_dtype(dtype)))
    m.arange = lambda *a, dtype=None: _wrap(torch.arange(*a, dtype=_dtype(dtype) or torch.int64))
    m.eye = lambda n, dtype=None: _wrap(torch.eye(n, dtype=_dtype(dtype) or _STATE["dtype"]))

    def _axis_kw(axis, keepdims):
        kw = {}
        if axis is not None:
            kw["dim"] = axis
        if keepdims:
            kw["keepdim"] = True
        return kw

    def _reduce(fn):
        def f(x, axis=None, keepdims=False):
            x = _as_array(x)
            if axis is None and not keepdims:
                return _wrap(fn(x))
            if axis is None:
                axis = tuple(range(x.dim()))
            return _wrap(fn(x, **_axis_kw(axis, keepdims)))
        return f

    m.sum = _reduce(torch.sum)
    m.mean = _reduce(torch.mean)
    m.max = _reduce(torch.amax)
    m.min = _reduce(torch.amin)
    m.std = lambda x, axis=None, keepdims=False, ddof=0: _wrap(
        torch.std(_as_array(x), correction=ddof, **_axis_kw(axis, keepdims)))
    m.var = lambda x, axis=None, keepdims=False, ddof=0: _wrap(
        torch.var(_as_array(x), correction=ddof, **_axis_kw(axis, keepdims)))
    m.argmax = lambda x, axis=None: _wrap(torch.argmax(_as_array(x), dim=axis))
    m.argmin = lambda x, axis=None: _wrap(torch.argmin(_as_array(x), dim=axis))
    m.cumsum = lambda x, axis=0: _wrap(torch.cumsum(_as_array(x), dim=axis))

    def quantile(x, q, axis=None, method="linear"):
        x = _as_array(x)
        qt = torch.as_tensor(q, dtype=x.dtype)
        return _wrap(torch.quantile(x, qt, dim=axis, interpolation=method))

    m.quantile = quantile
    m.hstack = lambda xs: _wrap(torch.hstack([_as_array(v) for v in xs]))
    m.vstack = lambda xs: _wrap(torch.vstack([_as_array(v) for v in xs]))
    m.stack = lambda xs, axis=0: _wrap(torch.stack([_as_array(v) for v in xs], dim=axis))
    m.concatenate = lambda xs, axis=0: _wrap(torch.cat([_as_array(v) for v in xs], dim=axis))
    m.einsum = lambda spec, *ops: _wrap(torch.einsum(spec, *[_as_array(o) for o in ops]))
    m.dot = lambda a, b: _wrap(torch.matmul(_as_array(a), _as_array(b)))
    m.matmul = m.dot
    m.transpose = lambda x, axes=None: _wrap(_as_array(x).permute(*axes) if axes else _as_array(x).T)
    m.squeeze = lambda x, axis=None: _wrap(torch.squeeze(_as_array(x)) if axis is None else torch.squeeze(_as_array(x), axis))
    m.expand_dims = lambda x, axis: _wrap(torch.unsqueeze(_as_array(x), axis))
    m.reshape = lambda x, shape: _wrap(_as_array(x).reshape(shape))

    def where(c, a=None, b=None):
        c = _as_array(c)
        if a is None:
            return tuple(_wrap(t) for t in torch.where(c))
        a, b = _as_array(a), _as_array(b)
        if a.dtype != b.dtype:
            a, b = a.to(_STATE["dtype"]), b.to(_STATE["dtype"])
        return _wrap(torch.where(c.bool(), a, b))

    m.where = where
    m.isscalar = np.isscalar
    m.shape = lambda x: tuple(_as_array(x).shape)
    return m


# --------------------------------------------------------------------------------------------
# jax.random

def _key_ints(key):
    k = np.asarray(key).reshape(-1).astype(np.int64)
    if k.size != 2:
        raise ValueError(f"jaxshim: expected a single PRNG key, got shape {np.asarray(key).shape}")
    return [int(k[0]), int(k[1])]


def _mk_key(arr):
    return torch.as_tensor(np.asarray(arr, dtype=np.int64)).as_subclass(KeyArray)


def normal_like_reference(key, shape, dtype=None):
    """The N(0,1) draws that `tfd.LogNormal(m, s).sample(seed=key)` consumes for an event of this
    shape under the shim: a pure function of (key, shape)."""
    shape = tuple(int(s) for s in shape)
    eps = np.random.default_rng(_key_ints(key)).standard_normal(shape)
    if _STATE["eps_f32"]:
        eps = eps.astype(np.float32).astype(np.float64)
    return torch.as_tensor(eps, dtype=dtype or _STATE["dtype"])


def _build_random():
    m = types.ModuleType("jax.random")

    def PRNGKey(seed):
        return _mk_key([0, int(seed)])

    def split(key, num=2):
        st = np.random.SeedSequence(_key_ints(key)).generate_state(2 * int(num), dtype=np.uint32)
        return _mk_key(st.reshape(int(num), 2))

    def fold_in(key, data):
        st = np.random.SeedSequence(_key_ints(key) + [int(data)]).generate_state(2, dtype=np.uint32)
        return _mk_key(st)

    def normal(key, shape=(), dtype=None):
        return _wrap(normal_like_reference(key, shape, _dtype(dtype)))

    def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
        u = np.random.default_rng(_key_ints(key)).uniform(minval, maxval, tuple(shape))
        return _as_array(u, dtype)

    def gamma(key, a, shape=None, dtype=None):
        a_np = np.asarray(a, dtype=np.float64)
        out = np.random.default_rng(_key_ints(key)).standard_gamma(a_np, size=shape)
        return _as_array(out, dtype)

    def choice(key, a, shape=(), replace=True, p=None):
        n = int(a) if np.isscalar(a) or (isinstance(a, torch.Tensor) and a.dim() == 0) else np.asarray(a)
        pp = None if p is None else np.asarray(p, dtype=np.float64)
        if pp is not None:
            pp = pp / pp.sum()
        out = np.random.default_rng(_key_ints(key)).choice(n, size=tuple(shape), replace=replace, p=pp)
        return _as_array(out)

    def permutation(key, x):
        n = int(x) if np.isscalar(x) else np.asarray(x)
        return _as_array(np.random.default_rng(_key_ints(key)).permutation(n))

    m.PRNGKey, m.key, m.split, m.fold_in = PRNGKey, PRNGKey, split, fold_in
    m.normal, m.uniform, m.gamma, m.choice, m.permutation = normal, uniform, gamma, choice, permutation
    return m


# --------------------------------------------------------------------------------------------
# jax core: jit, vmap, value_and_grad, lax, scipy

def _tree_leaves(x):
    if isinstance(x, (list, tuple)):
        out = []
        for v in x:
            out.extend(_tree_leaves(v))
        return out
    return [x]


def _tree_unflatten(like, leaves):
    it = iter(leaves)

    def rec(x):
        if isinstance(x, (list, tuple)):
            return type(x)(rec(v) for v in x)
        return next(it)

    return rec(like)


def _tree_map(fn, tree, *rest):
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(fn, t, *[r[i] for r in rest]) for i, t in enumerate(tree))
    return fn(tree, *rest)


def _vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        if len(axes) != len(args):
            raise ValueError("jaxshim.vmap: in_axes does not match the arguments")
        has_key = any(ax is not None and isinstance(a, KeyArray) for a, ax in zip(args, axes))
        if has_key:
            n = next(a.shape[ax] for a, ax in zip(args, axes) if ax is not None)
            outs = []
            for i in range(n):
                call = [a if ax is None else torch.select(a, ax, i) for a, ax in zip(args, axes)]
                outs.append(fun(*call))
            if isinstance(outs[0], (tuple, list)):
                return type(outs[0])(_wrap(torch.stack([o[j] for o in outs], dim=out_axes))
                                     for j in range(len(outs[0])))
            return _wrap(torch.stack([_as_array(o) for o in outs], dim=out_axes))

        conv = [a if ax is None else _unwrap(_as_array(a)) for a, ax in zip(args, axes)]

        def inner(*xs):
            out = fun(*[_wrap(x) for x in xs])
            return _tree_map(_unwrap, out) if isinstance(out, (tuple, list)) else _unwrap(out)

        out = torch.func.vmap(inner, in_dims=tuple(axes), out_dims=out_axes)(*conv)
        return _tree_map(_wrap, out) if isinstance(out, (tuple, list)) else _wrap(out)

    return mapped


def _value_and_grad(fun, argnums=0, has_aux=False):
    def vg(*args, **kwargs):
        args = list(args)
        target = args[argnums]
        leaves = [_unwrap(_as_array(t)).detach().clone().requires_grad_(True).as_subclass(Array)
                  for t in _tree_leaves(target)]
        args[argnums] = _tree_unflatten(target, leaves)
        out = fun(*args, **kwargs)
        value, aux = (out if has_aux else (out, None))
        grads = torch.autograd.grad(value, leaves, allow_unused=True)
        grads = [_wrap(torch.zeros_like(p)) if g is None else _wrap(g.detach()) for p, g in zip(leaves, grads)]
        g_tree = _tree_unflatten(target, grads)
        value = _wrap(value.detach())
        return ((value, aux), g_tree) if has_aux else (value, g_tree)

    return vg


def _build_lax():
    m = types.ModuleType("jax.lax")

    def cond(pred, true_fun, false_fun, *operands):
        return true_fun(*operands) if bool(pred) else false_fun(*operands)

    def stop_gradient(x):
        return _tree_map(lambda t: _wrap(_as_array(t).detach()), x) if isinstance(x, (list, tuple)) \
            else _wrap(_as_array(x).detach())

    def scan(f, init, xs=None, length=None):
        n = length if xs is None else len(_tree_leaves(xs)[0])
        carry, ys = init, []
        for i in range(int(n)):
            x = None if xs is None else _tree_map(lambda t: t[i], xs)
            carry, y = f(carry, x)
            ys.append(y)
        if ys and ys[0] is not None:
            if isinstance(ys[0], (tuple, list)):
                ys = type(ys[0])(_wrap(torch.stack([y[j] for y in ys])) for j in range(len(ys[0])))
            else:
                ys = _wrap(torch.stack([_as_array(y) for y in ys]))
        else:
            ys = None
        return carry, ys

    def fori_loop(lo, hi, body, init):
        val = init
        for i in range(int(lo), int(hi)):
            val = body(i, val)
        return val

    m.cond, m.stop_gradient, m.scan, m.fori_loop = cond, stop_gradient, scan, fori_loop
    return m


def _build_scipy():
    sp = types.ModuleType("jax.scipy")
    stats = types.ModuleType("jax.scipy.stats")
    special = types.ModuleType("jax.scipy.special")
    poisson = types.ModuleType("jax.scipy.stats.poisson")

    def logpmf(k, mu, loc=0):
        k, mu = _as_array(k, _F32), _as_array(mu)
        x = k - loc
        lp = torch.xlogy(x, mu) - torch.lgamma(x + 1.0) - mu
        return _wrap(torch.where(x < 0, torch.full_like(lp, -math.inf), lp))

    poisson.logpmf = logpmf
    poisson.pmf = lambda k, mu, loc=0: _wrap(torch.exp(logpmf(k, mu, loc)))
    special.gammaln = lambda x: _wrap(torch.lgamma(_as_array(x)))
    special.digamma = lambda x: _wrap(torch.digamma(_as_array(x)))
    special.xlogy = lambda x, y: _wrap(torch.xlogy(_as_array(x), _as_array(y)))
    special.logsumexp = lambda x, axis=None: _wrap(torch.logsumexp(_as_array(x), dim=axis))
    stats.poisson = poisson
    sp.stats, sp.special = stats, special
    return sp, stats, special, poisson


def _build_nn():
    m = types.ModuleType("jax.nn")
    def one_hot(x, num_classes, dtype=None):
        return _wrap(torch.nn.functional.one_hot(_as_array(x).long(), int(num_classes)).to(_dtype(dtype) or _STATE["dtype"]))

    m.one_hot = one_hot
    m.softmax = lambda x, axis=-1: _wrap(torch.softmax(_as_array(x), dim=axis))
    m.relu = lambda x: _wrap(torch.relu(_as_array(x)))
    return m


# --------------------------------------------------------------------------------------------
# tensorflow_probability.substrates.jax.distributions

class _LogNormal:
    def __init__(self, loc, scale):
        self.loc, self.scale = _as_array(loc), _as_array(scale)

    def sample(self, sample_shape=(), seed=None):
        shape = tuple(sample_shape) + tuple(torch.broadcast_shapes(self.loc.shape, self.scale.shape))
        eps = normal_like_reference(seed, shape, self.loc.dtype if self.loc.dtype.is_floating_point else None)
        return _wrap(torch.exp(self.loc + self.scale * eps))

    def entropy(self):
        return _wrap(self.loc + 0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(self.scale))

    def log_prob(self, x):
        x = _as_array(x)
        z = (torch.log(x) - self.loc) / self.scale
        return _wrap(-0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi) - torch.log(x))

    def mean(self):
        return _wrap(torch.exp(self.loc + 0.5 * self.scale ** 2))


class _Gamma:
    def __init__(self, concentration, rate=None):
        self.concentration, self.rate = _as_array(concentration), _as_array(rate)

    def log_prob(self, x):
        x, a, b = _as_array(x), self.concentration, self.rate
        return _wrap(torch.xlogy(a - 1.0, x) - b * x - torch.lgamma(a) + a * torch.log(b))

    def sample(self, sample_shape=(), seed=None):
        if isinstance(sample_shape, (int, np.integer)):
            sample_shape = (int(sample_shape),)
        a = _unwrap(self.concentration).detach().numpy()
        b = _unwrap(self.rate).detach().numpy()
        a, b = np.broadcast_arrays(a, b)
        shape = tuple(sample_shape) + a.shape
        out = np.random.default_rng(_key_ints(seed)).standard_gamma(np.broadcast_to(a, shape)) / np.broadcast_to(b, shape)
        return _as_array(out)

    def mean(self):
        return _wrap(self.concentration / self.rate)


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = _as_array(loc), _as_array(scale)

    def sample(self, sample_shape=(), seed=None):
        shape = tuple(sample_shape) + tuple(torch.broadcast_shapes(self.loc.shape, self.scale.shape))
        return _wrap(self.loc + self.scale * normal_like_reference(seed, shape))

    def log_prob(self, x):
        z = (_as_array(x) - self.loc) / self.scale
        return _wrap(-0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi))


# --------------------------------------------------------------------------------------------
# optax

class _AdamState:
    def __init__(self, count, mu, nu):
        self.count, self.mu, self.nu = count, mu, nu


class _GradientTransformation:
    def __init__(self, init, update):
        self.init, self.update = init, update


def _build_optax():
    m = types.ModuleType("optax")

    def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
        def init(params):
            z = lambda p: _wrap(torch.zeros_like(_as_array(p)))
            return (_AdamState(0, _tree_map(z, params), _tree_map(z, params)), ())

        def update(updates, state, params=None):
            st = state[0]
            count = st.count + 1
            mu = _tree_map(lambda g, mm: _wrap((1.0 - b1) * _as_array(g) + b1 * mm), updates, st.mu)
            nu = _tree_map(lambda g, vv: _wrap((1.0 - b2) * (_as_array(g) * _as_array(g)) + b2 * vv), updates, st.nu)
            c1 = 1.0 - b1 ** count
            c2 = 1.0 - b2 ** count
            lr = learning_rate(count - 1) if callable(learning_rate) else learning_rate
            upd = _tree_map(
                lambda mm, vv: _wrap(-lr * ((mm / c1) / (torch.sqrt(vv / c2 + eps_root) + eps))), mu, nu)
            return upd, (_AdamState(count, mu, nu), ())

        return _GradientTransformation(init, update)

    def apply_updates(params, updates):
        return _tree_map(lambda p, u: _wrap(_as_array(p) + _as_array(u).to(_as_array(p).dtype)), params, updates)

    m.adam, m.apply_updates = adam, apply_updates
    m.GradientTransformation = _GradientTransformation
    return m


# --------------------------------------------------------------------------------------------
# inert stand-ins for the plotting / AnnData imports at the top of the reference's model files

class AnnData:
    """The few attributes `scDEF.__init__` / `load_adata` / `_learn` touch (`_scdef.py:482-578`)."""

    def __init__(self, X=None, obs=None, var=None, uns=None, obsm=None, layers=None):
        import pandas as pd

        self.X = X
        n, g = X.shape
        self.obs = obs if obs is not None else pd.DataFrame(index=[f"cell{i}" for i in range(n)])
        self.var = var if var is not None else pd.DataFrame(index=[f"g{i}" for i in range(g)])
        self.uns = dict(uns or {})
        self.obsm = dict(obsm or {})
        self.layers = dict(layers or {})
        self.raw = None

    @property
    def shape(self):
        return tuple(self.X.shape)

    n_obs = property(lambda self: self.X.shape[0])
    n_vars = property(lambda self: self.X.shape[1])
    obs_names = property(lambda self: self.obs.index)
    var_names = property(lambda self: self.var.index)

    def copy(self):
        import copy as _copy

        return AnnData(self.X.copy(), self.obs.copy(), self.var.copy(), _copy.deepcopy(self.uns),
                       {k: np.array(v) for k, v in self.obsm.items()},
                       {k: v.copy() for k, v in self.layers.items()})


def _build_stubs():
    out = {}
    ad = types.ModuleType("anndata")
    ad.AnnData = AnnData
    ad.read_h5ad = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("jaxshim: no h5ad support"))
    out["anndata"] = ad

    sns = types.ModuleType("seaborn")

    def color_palette(palette=None, n_colors=None, **kw):
        n = int(n_colors or 6)
        return [((i + 1) / (n + 1), 0.5, 1.0 - (i + 1) / (n + 1)) for i in range(n)]

    sns.color_palette = color_palette
    out["seaborn"] = sns

    mpl = types.ModuleType("matplotlib")
    colors = types.ModuleType("matplotlib.colors")

    def to_hex(c, keep_alpha=False):
        if isinstance(c, str):
            return c
        return "#" + "".join(f"{int(round(255 * float(v))):02x}" for v in tuple(c)[:3])

    def to_rgb(c):
        if isinstance(c, str):
            c = c.lstrip("#")
            return tuple(int(c[i:i + 2], 16) / 255.0 for i in (0, 2, 4))
        return tuple(c)[:3]

    colors.to_hex, colors.to_rgb = to_hex, to_rgb
    colors.cnames = {}
    mpl.colors = colors
    out["matplotlib"] = mpl
    out["matplotlib.colors"] = colors
    return out


# --------------------------------------------------------------------------------------------

def build_modules():
    jnp = _build_numpy()
    rnd = _build_random()
    lax = _build_lax()
    sp, stats, special, poisson = _build_scipy()
    nn = _build_nn()

    jax = types.ModuleType("jax")
    jax.__path__ = []  # a package: "import jax.numpy" must resolve through sys.modules
    jax.numpy, jax.random, jax.lax, jax.scipy, jax.nn = jnp, rnd, lax, sp, nn
    jax.jit = lambda f=None, **kw: (f if f is not None else (lambda g: g))
    jax.vmap = _vmap
    jax.value_and_grad = _value_and_grad
    jax.grad = lambda f, argnums=0: (lambda *a, **k: _value_and_grad(f, argnums)(*a, **k)[1])
    jax.clear_caches = lambda: None
    jax.device_get = lambda x: x
    jax.devices = lambda *a: ["cpu:0 (jaxshim/torch)"]
    jax.Array = Array
    tree_util = types.ModuleType("jax.tree_util")
    tree_util.tree_map = _tree_map
    tree_util.tree_leaves = _tree_leaves
    jax.tree_util = tree_util
    jax.tree_map = _tree_map

    tfp = types.ModuleType("tensorflow_probability")
    tfp.__path__ = []
    sub = types.ModuleType("tensorflow_probability.substrates")
    sub.__path__ = []
    sj = types.ModuleType("tensorflow_probability.substrates.jax")
    sj.__path__ = []
    tfd = types.ModuleType("tensorflow_probability.substrates.jax.distributions")
    tfd.LogNormal, tfd.Gamma, tfd.Normal = _LogNormal, _Gamma, _Normal
    tfp.substrates, sub.jax, sj.distributions = sub, sj, tfd

    mods = {
        "jax": jax, "jax.numpy": jnp, "jax.random": rnd, "jax.lax": lax, "jax.scipy": sp,
        "jax.scipy.stats": stats, "jax.scipy.special": special, "jax.scipy.stats.poisson": poisson,
        "jax.nn": nn, "jax.tree_util": tree_util,
        "tensorflow_probability": tfp, "tensorflow_probability.substrates": sub,
        "tensorflow_probability.substrates.jax": sj,
        "tensorflow_probability.substrates.jax.distributions": tfd,
        "optax": _build_optax(),
    }
    mods.update(_build_stubs())
    return mods


@contextlib.contextmanager
def installed(dtype=torch.float64, extra=None):
    """Put the stand-in modules into `sys.modules` for the duration of the block (the names that
    really exist in the interpreter — none of them does in this image — are restored afterwards)."""
    mods = build_modules()
    if extra:
        mods.update(extra)
    saved = {name: sys.modules.get(name) for name in mods}
    prev = _STATE["dtype"]
    _STATE["dtype"] = dtype
    sys.modules.update(mods)
    try:
        yield mods
    finally:
        _STATE["dtype"] = prev
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old
