"""A NumPy-backed stand-in for the subset of the jax / equinox API that the reference's hot-path modules use,
so that the REFERENCE'S OWN, UNMODIFIED code (imported from /root/reference by make_golden_hotpath.py) can be
executed in this container, where jax itself is not installable.  Test infrastructure only: nothing in the
product or in the oracle imports this file.

What this is and is not.  Every function below restates the *documented* semantics of the jax function of the
same name with NumPy (eager evaluation, no tracing): ``jit`` is the identity, ``vmap`` is a Python loop followed by
a stack, ``x.at[i].set(v)`` copies, ``lax.cond`` is an ``if``.  Golden vectors produced through it therefore pin
everything that lives in the reference's own source -- operator/term ordering, right-to-left application of
operator strings, NaN marking and compaction order, soft pseudo-inverse and SNR formulas, group closure order,
neighbour-table layout, phase conventions -- but they do NOT pin third-party numerics (XLA's eigh, gelu, PRNG),
which stay "parity unpinned" as DESIGN.md section 2 says.
"""
from __future__ import annotations

import functools
import sys
import types

import numpy as np
import scipy.linalg


# ---- arrays ------------------------------------------------------------------------------------------------
class Arr(np.ndarray):
    """ndarray with the functional-update property ``.at`` of jax arrays."""

    @property
    def at(self):
        return _At(self)


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def _new(self):
        return np.array(self.a, copy=True).view(Arr)

    def set(self, v):
        out = self._new()
        out[self.idx] = v
        return out

    def add(self, v):
        out = self._new()
        np.add.at(out, self.idx, v)
        return out

    def mul(self, v):
        out = self._new()
        np.multiply.at(out, self.idx, v)
        return out

    multiply = mul

    def get(self):
        return wrap(self.a[self.idx])


def wrap(x):
    if isinstance(x, np.ndarray) and not isinstance(x, Arr):
        return x.view(Arr)
    if isinstance(x, tuple):
        return tuple(wrap(v) for v in x)
    if isinstance(x, list):
        return [wrap(v) for v in x]
    return x


def _wrapped(fn):
    @functools.wraps(fn)
    def inner(*a, **k):
        return wrap(fn(*a, **k))

    return inner


# ---- jax.numpy ---------------------------------------------------------------------------------------------
def _nonzero(a, size=None, fill_value=None):
    idx = np.nonzero(np.asarray(a))
    if size is None:
        return tuple(i.view(Arr) for i in idx)
    fills = fill_value if isinstance(fill_value, (tuple, list)) else (0 if fill_value is None else fill_value,) * len(idx)
    out = []
    for i, f in zip(idx, fills):
        o = np.full(size, f, dtype=i.dtype)
        n = min(size, i.size)
        o[:n] = i[:n]
        out.append(o.view(Arr))
    return tuple(out)


def _flatnonzero(a, size=None, fill_value=None):
    return _nonzero(np.ravel(np.asarray(a)), size=size, fill_value=fill_value)[0]


def _argwhere(a, size=None, fill_value=None):
    idx = np.argwhere(np.asarray(a))
    if size is None:
        return idx.view(Arr)
    out = np.full((size, idx.shape[1]), 0 if fill_value is None else fill_value, dtype=idx.dtype)
    n = min(size, idx.shape[0])
    out[:n] = idx[:n]
    return out.view(Arr)


def _cumulative_sum(x, axis=None, dtype=None, include_initial=False):
    x = np.asarray(x)
    if axis is None:
        axis = 0
    c = np.cumsum(x, axis=axis, dtype=dtype)
    if include_initial:
        shape = list(c.shape)
        shape[axis] = 1
        c = np.concatenate([np.zeros(shape, dtype=c.dtype), c], axis=axis)
    return c.view(Arr)


def _argsort(a, axis=-1, stable=True, descending=False, kind=None):
    a = np.asarray(a)
    if descending:
        a = -a
    return np.argsort(a, axis=axis, kind="stable").view(Arr)


def make_jnp():
    jnp = types.ModuleType("jax.numpy")
    for name in dir(np):
        if name.startswith("_"):
            continue
        obj = getattr(np, name)
        if isinstance(obj, type) or not callable(obj):
            setattr(jnp, name, obj)  # dtypes, constants
        else:
            setattr(jnp, name, _wrapped(obj))
    jnp.ndarray = Arr

    def _drop_device(fn):
        def inner(*a, device=None, **k):  # placement arguments have no meaning here
            return wrap(fn(*a, **k))

        return inner

    for name in ("asarray", "array", "zeros", "ones", "arange", "full", "empty"):
        setattr(jnp, name, _drop_device(getattr(np, name)))
    jnp.nonzero = _nonzero
    jnp.flatnonzero = _flatnonzero
    jnp.argwhere = _argwhere
    jnp.cumulative_sum = _cumulative_sum
    jnp.argsort = _argsort
    jnp.bool_ = np.bool_
    jnp.issubdtype = np.issubdtype
    jnp.finfo = np.finfo
    jnp.dtype = np.dtype
    linalg = types.ModuleType("jax.numpy.linalg")
    for name in ("norm", "eigh", "solve", "inv", "det", "slogdet", "pinv", "svd", "qr"):
        setattr(linalg, name, _wrapped(getattr(np.linalg, name)))
    linalg.trace = _wrapped(np.trace)
    jnp.linalg = linalg
    return jnp


# ---- pytrees (dict / list / tuple of arrays) ------------------------------------------------------------------
def _tree_stack(items, axis=0):
    first = items[0]
    if _is_registered(first):
        flat = [tree_flatten(it) for it in items]
        stacked = [_tree_stack([f[0][i] for f in flat], axis) for i in range(len(flat[0][0]))]
        return tree_unflatten(flat[0][1], stacked)
    if isinstance(first, dict):
        return {k: _tree_stack([it[k] for it in items], axis) for k in first}
    if isinstance(first, (list, tuple)):
        return type(first)(_tree_stack([it[i] for it in items], axis) for i in range(len(first)))
    return np.stack([np.asarray(it) for it in items], axis=axis).view(Arr)


def _tree_index(x, i, axis):
    if _is_registered(x):
        leaves, treedef = tree_flatten(x)
        return tree_unflatten(treedef, [_tree_index(v, i, axis) for v in leaves])
    if isinstance(x, dict):
        return {k: _tree_index(v, i, axis) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_tree_index(v, i, axis) for v in x)
    return wrap(np.take(np.asarray(x), i, axis=axis)) if np.ndim(x) else x


def _tree_batch(x, axis):
    if _is_registered(x):
        return _tree_batch(tree_flatten(x)[0][0], axis)
    if isinstance(x, dict):
        x = next(iter(x.values()))
        return _tree_batch(x, axis)
    if isinstance(x, (list, tuple)):
        return _tree_batch(x[0], axis)
    return np.shape(x)[axis]


# registered pytree classes (jax.tree_util.register_pytree_node_class): objects with tree_flatten / tree_unflatten
_REGISTERED = []


def register_pytree_node_class(cls):
    _REGISTERED.append(cls)
    return cls


def _is_registered(x):
    return any(isinstance(x, c) for c in _REGISTERED)


def tree_flatten(tree):
    """(leaves, treedef); None is an empty subtree like in jax."""
    leaves = []

    def rec(x):
        if x is None:
            return ("none",)
        if _is_registered(x):
            children, aux = x.tree_flatten()
            return ("reg", type(x), aux, [rec(c) for c in children])
        if isinstance(x, dict):
            return ("dict", list(x.keys()), [rec(x[k]) for k in x])
        if isinstance(x, (list, tuple)):
            return ("seq", type(x), [rec(v) for v in x])
        leaves.append(x)
        return ("leaf",)

    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        if d[0] == "none":
            return None
        if d[0] == "leaf":
            return next(it)
        if d[0] == "reg":
            return d[1].tree_unflatten(d[2], [rec(c) for c in d[3]])
        if d[0] == "dict":
            return {k: rec(c) for k, c in zip(d[1], d[2])}
        return d[1](rec(c) for c in d[2])

    return rec(treedef)


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


class custom_jvp:
    """Only the primal function is ever evaluated here."""

    def __init__(self, fn, nondiff_argnums=()):
        self.fn = fn
        functools.update_wrapper(self, fn)

    def __call__(self, *a, **k):
        return self.fn(*a, **k)

    def defjvp(self, f, **_kw):
        return f


def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _tree_batch(a, ax)
                break
        if n is None:
            raise ValueError("vmap needs at least one mapped argument")
        outs = [fn(*[a if ax is None else _tree_index(a, i, ax) for a, ax in zip(args, axes)]) for i in range(n)]
        return _tree_stack(outs, out_axes)

    return mapped


def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def segment_sum(data, segment_ids, num_segments=None, **_kw):
    data, seg = np.asarray(data), np.asarray(segment_ids)
    if num_segments is None:
        num_segments = int(seg.max()) + 1
    out = np.zeros((num_segments,) + data.shape[1:], dtype=data.dtype)
    ok = (seg >= 0) & (seg < num_segments)  # out-of-range ids are dropped
    np.add.at(out, seg[ok], data[ok])
    return out.view(Arr)


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _gelu(x, approximate=True):
    x = np.asarray(x)
    if approximate:
        return wrap(0.5 * x * (1 + np.tanh(np.sqrt(2 / np.pi) * (x + 0.044715 * x ** 3))))
    from scipy.special import erf

    return wrap(0.5 * x * (1 + erf(x / np.sqrt(2))))


class _Device:
    platform = "cpu"
    id = 0


class _Anything:
    """Inert object for annotations and decorators that are never exercised."""

    def __call__(self, *a, **k):
        if a and callable(a[0]) and not isinstance(a[0], _Anything):
            return a[0]
        return self

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, item):
        return self

    def __or__(self, other):
        return self

    __ror__ = __or__

    def __mro_entries__(self, bases):
        return (object,)


class _AnyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


# ---- equinox stand-ins (third-party semantics restated from the equinox documentation) --------------------------------
class Module:
    """equinox.Module for classes that define their own __init__: a plain mutable base class."""

    def __init__(self, *a, **k):
        pass


class Conv(Module):
    """equinox.nn.Conv: cross-correlation over [channels, *spatial] inputs (no batch axis), weight
    [out, in, *kernel], bias [out, 1, ...]; padding="SAME" with padding_mode "CIRCULAR" (wrap) or "ZEROS"."""

    def __init__(self, num_spatial_dims, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 use_bias=True, padding_mode="ZEROS", dtype=None, *, key=None):
        nd = num_spatial_dims
        k = (kernel_size,) * nd if isinstance(kernel_size, int) else tuple(kernel_size)
        dtype = np.float32 if dtype is None else dtype
        if padding != "SAME" or stride != 1 or dilation != 1 or groups != 1:
            raise NotImplementedError("only the configuration the reference's ResConv uses")
        self.num_spatial_dims, self.kernel_size, self.padding_mode = nd, k, padding_mode
        self.weight = np.zeros((out_channels, in_channels) + k, dtype=dtype).view(Arr)
        self.bias = np.zeros((out_channels,) + (1,) * nd, dtype=dtype).view(Arr) if use_bias else None

    def __call__(self, x, *, key=None):
        import itertools

        x = np.asarray(x)
        k, nd = self.kernel_size, self.num_spatial_dims
        pads = [(0, 0)] + [((kk - 1) // 2, kk - 1 - (kk - 1) // 2) for kk in k]
        xp = np.pad(x, pads, mode="wrap" if self.padding_mode == "CIRCULAR" else "constant")
        w = np.asarray(self.weight)
        out = np.zeros((w.shape[0],) + x.shape[1:], dtype=np.result_type(x.dtype, w.dtype))
        for tap in itertools.product(*[range(kk) for kk in k]):
            sl = (slice(None),) + tuple(slice(t, t + n) for t, n in zip(tap, x.shape[1:]))
            out += np.tensordot(w[(slice(None), slice(None)) + tap], xp[sl], axes=([1], [0]))
        if self.bias is not None:
            out = out + np.asarray(self.bias)
        return out.view(Arr)


class EqxSequential(Module):
    def __init__(self, layers):
        self.layers = tuple(layers)


class Lambda(Module):
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x, *, key=None):
        return self.fn(x)


def _eqx_combine(*trees):
    """equinox.combine: leaf-wise first non-None of trees of equal structure (None = missing leaf)."""
    def rec(xs):
        xs = [x for x in xs if x is not None]
        if not xs:
            return None
        x0 = xs[0]
        if isinstance(x0, dict):
            return {k: rec([x.get(k) for x in xs]) for k in x0}
        if isinstance(x0, (list, tuple)) and not _is_registered(x0):
            return type(x0)(rec([x[i] for x in xs]) for i in range(len(x0)))
        return x0

    return rec(list(trees))


def install():
    """Put the stand-ins into sys.modules under the names the reference imports."""
    jnp = make_jnp()
    jax = _AnyModule("jax")
    jax.numpy = jnp
    jax.Array = Arr
    jax.jit = jit
    jax.vmap = vmap
    jax.device_count = lambda: 1
    jax.local_device_count = lambda: 1
    jax.process_count = lambda: 1
    jax.process_index = lambda: 0
    jax.devices = lambda *a: [_Device()]
    lax = _AnyModule("jax.lax")
    lax.cond = cond
    lax.stop_gradient = lambda x: x
    lax.complex = lambda re, im: wrap(np.asarray(re) + 1j * np.asarray(im))
    jax.lax = lax
    jax.custom_jvp = custom_jvp
    tu = types.ModuleType("jax.tree_util")
    tu.register_pytree_node_class = register_pytree_node_class
    tu.tree_flatten, tu.tree_unflatten, tu.tree_map = tree_flatten, tree_unflatten, tree_map
    jax.tree_util = tu
    tree = types.ModuleType("jax.tree")
    tree.flatten, tree.unflatten, tree.map = tree_flatten, tree_unflatten, tree_map
    jax.tree = tree
    ops = types.ModuleType("jax.ops")
    ops.segment_sum = segment_sum
    jax.ops = ops
    nn = _AnyModule("jax.nn")
    nn.relu = _wrapped(lambda x: np.maximum(x, 0))
    nn.gelu = _gelu
    jax.nn = nn
    jsl = types.ModuleType("jax.scipy.linalg")
    jsl.eigh = _wrapped(lambda a, **k: np.linalg.eigh(a))
    jsl.solve = _wrapped(lambda a, b, assume_a="gen", **k: scipy.linalg.solve(a, b, assume_a=assume_a))
    jsp = _AnyModule("jax.scipy")
    jsp.linalg = jsl
    jax.scipy = jsp
    mods = {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": jnp.linalg, "jax.lax": lax, "jax.ops": ops, "jax.nn": nn,
            "jax.scipy": jsp, "jax.scipy.linalg": jsl}
    mods["jax.tree_util"], mods["jax.tree"] = tu, tree
    for name in ("jax.random", "jax.flatten_util", "jax.sharding", "jax.typing", "jax.experimental",
                 "jax.experimental.multihost_utils", "jax.scipy.sparse", "jax.scipy.sparse.linalg", "jax.scipy.special",
                 "jaxtyping", "equinox", "equinox.nn", "lrux", "quspin"):
        mods[name] = _AnyModule(name)
    eqx = mods["equinox"]
    eqx.is_array_like = lambda x: isinstance(x, (int, float, complex, np.number))
    eqx.is_array = lambda x: isinstance(x, np.ndarray)
    eqx.filter_jit = jit
    # eqx.filter / partition / combine on the pytrees above (documented behaviour: leaves failing the predicate -> None)
    eqx.filter = lambda tree, pred, inverse=False: tree_map(lambda x: x if bool(pred(x)) != inverse else None, tree)
    eqx.partition = lambda tree, pred: (eqx.filter(tree, pred), eqx.filter(tree, pred, inverse=True))
    eqx.combine = _eqx_combine
    eqx.Module = Module
    eqx.field = lambda **k: None
    eqx.nn = mods["equinox.nn"]
    eqx.nn.Conv, eqx.nn.Sequential, eqx.nn.Lambda = Conv, EqxSequential, Lambda
    sys.modules.update(mods)
    return jax, jnp
