"""Amplitude containers and small host utilities (quantax/utils/big_array.py, data.py, basis.py).

psi travels between sampler / operator / state as a pair of float64 device vectors:
``LogArray(sign, logabs)`` for RBM_Dense and ``ScaleArray(significand, exponent)`` for ResConv;
both mean value = first * exp(second).  Only the members the hot path touches are provided.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


def _t(x, like: torch.Tensor) -> torch.Tensor:
    """Dense operand -> tensor on the device of ``like`` (Python scalars keep the dtype of ``like``)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (int, float)):
        return torch.tensor(x, dtype=like.dtype if like.is_floating_point() or like.is_complex() else torch.float64,
                            device=like.device)
    return torch.as_tensor(np.asarray(x), device=like.device)


def _real_dtype(dt):
    return {torch.complex128: torch.float64, torch.complex64: torch.float32}.get(dt, dt)


def _addexp(x1, x2, b1, b2):
    """b1 exp(x1) + b2 exp(x2) as (x, b) (big_array.py:14-26): the larger exponent is factored out exactly."""
    xmax = torch.maximum(x1, x2)
    r1 = torch.where(x1 != xmax, torch.exp(x1 - xmax), torch.ones_like(xmax))
    r2 = torch.where(x2 != xmax, torch.exp(x2 - xmax), torch.ones_like(xmax))
    return xmax, b1 * r1 + b2 * r2


def _sumexp(x, b, axis, keepdims, mean):
    """sum_i b_i exp(x_i) over ``axis`` as (x, b) (big_array.py:46-148)."""
    if axis is None:
        axis = tuple(range(x.ndim))
    elif isinstance(axis, int):
        axis = (axis,)
    if x.ndim == 0:
        return x, b
    xmax = torch.amax(x, dim=axis, keepdim=True)
    r = torch.where(x != xmax, torch.exp(x - xmax), torch.ones_like(x))
    bs = torch.sum(b * r, dim=axis, keepdim=True)
    if mean:
        n = 1
        for a in axis:
            n *= x.shape[a]
        xmax = xmax - float(np.log(n))
    if not keepdims:
        for a in sorted((a % x.ndim for a in axis), reverse=True):
            xmax, bs = xmax.squeeze(a), bs.squeeze(a)
    return xmax, bs


@dataclass
class LogArray:
    """value = sign * exp(logabs) (quantax/utils/big_array.py:154-402); zero is sign = 0, logabs = -inf."""

    sign: torch.Tensor
    logabs: torch.Tensor

    __array_priority__ = 1000
    mult = property(lambda self: self.sign)
    expo = property(lambda self: self.logabs)
    shape = property(lambda self: self.sign.shape)
    dtype = property(lambda self: self.sign.dtype)
    ndim = property(lambda self: self.sign.ndim)
    size = property(lambda self: self.sign.numel())

    @staticmethod
    def from_value(x) -> "LogArray":
        if isinstance(x, LogArray):
            return x
        if isinstance(x, ScaleArray):
            return LogArray(torch.sgn(x.significand), torch.log(x.significand.abs()) + x.exponent)
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64))
        return LogArray(torch.sgn(x), torch.log(x.abs()))

    def _other(self, other) -> "LogArray":
        if isinstance(other, (LogArray, ScaleArray)):
            return LogArray.from_value(other)
        return LogArray.from_value(_t(other, self.logabs))

    def value(self) -> torch.Tensor:
        return self.sign * torch.exp(self.logabs)

    def __getitem__(self, idx):
        return LogArray(self.sign[idx], self.logabs[idx])

    def __len__(self):
        return self.sign.shape[0]

    def reshape(self, *shape):
        return LogArray(self.sign.reshape(*shape), self.logabs.reshape(*shape))

    def flatten(self):
        return LogArray(self.sign.flatten(), self.logabs.flatten())

    def __neg__(self):
        return LogArray(-self.sign, self.logabs)

    def conj(self):
        return LogArray(torch.conj(self.sign).resolve_conj(), self.logabs)

    def abs(self):
        return LogArray(torch.ones_like(self.logabs), self.logabs)

    __abs__ = abs

    @property
    def real(self):
        if not self.sign.is_complex():
            return self
        re = self.sign.real
        return LogArray(torch.sgn(re), self.logabs + torch.log(re.abs()))

    @property
    def imag(self):
        if not self.sign.is_complex():
            return LogArray(torch.zeros_like(self.sign), torch.full_like(self.logabs, -float("inf")))
        im = self.sign.imag
        return LogArray(torch.sgn(im), self.logabs + torch.log(im.abs()))

    def astype(self, dtype):
        return LogArray(self.sign.to(dtype), self.logabs.to(_real_dtype(dtype)))

    def __mul__(self, other):
        o = self._other(other)
        return LogArray(self.sign * o.sign, self.logabs + o.logabs)

    __rmul__ = __mul__

    def __truediv__(self, other):
        o = self._other(other)
        return LogArray(self.sign / o.sign, self.logabs - o.logabs)

    def __rtruediv__(self, other):
        o = self._other(other)
        return LogArray(o.sign / self.sign, o.logabs - self.logabs)

    def __pow__(self, p):
        p = _t(p, self.logabs)
        return LogArray(torch.pow(self.sign, p), self.logabs * p)

    def __add__(self, other):
        o = self._other(other)
        x, b = _addexp(self.logabs, o.logabs, self.sign, o.sign)
        return LogArray(torch.sgn(b), x + torch.log(b.abs()))

    __radd__ = __add__

    def __sub__(self, other):
        return self.__add__(-self._other(other))

    def __rsub__(self, other):
        return self._other(other).__add__(-self)

    def sum(self, axis=None, keepdims: bool = False):
        x, b = _sumexp(self.logabs, self.sign, axis, keepdims, mean=False)
        return LogArray(torch.sgn(b), x + torch.log(b.abs()))

    def mean(self, axis=None, keepdims: bool = False):
        x, b = _sumexp(self.logabs, self.sign, axis, keepdims, mean=True)
        return LogArray(torch.sgn(b), x + torch.log(b.abs()))

    def prod(self, axis=None, keepdims: bool = False):
        if axis is None:
            sign, logabs = torch.prod(self.sign.flatten()), torch.sum(self.logabs)
            if keepdims:
                sign, logabs = sign.reshape([1] * self.ndim), logabs.reshape([1] * self.ndim)
            return LogArray(sign, logabs)
        return LogArray(torch.prod(self.sign, dim=axis, keepdim=keepdims), torch.sum(self.logabs, dim=axis, keepdim=keepdims))

    def __array__(self, dtype=None):
        return np.asarray(self.value().cpu().numpy(), dtype)


@dataclass
class ScaleArray:
    """value = significand * exp(exponent) (quantax/utils/big_array.py:407-688); the exponent is a scalar or has
    the shape of the significand."""

    significand: torch.Tensor
    exponent: torch.Tensor

    __array_priority__ = 2000
    mult = property(lambda self: self.significand)
    expo = property(lambda self: self.exponent)
    shape = property(lambda self: self.significand.shape)
    dtype = property(lambda self: self.significand.dtype)
    ndim = property(lambda self: self.significand.ndim)
    size = property(lambda self: self.significand.numel())

    def normalize(self) -> "ScaleArray":
        """Largest |significand| becomes 1 relative to the largest exponent (big_array.py:442-451)."""
        max_sig = self.significand.abs().max()
        max_exp = self.exponent.max()
        exponent = max_exp + torch.log(max_sig)
        exponent = torch.where(torch.isfinite(exponent), exponent, max_exp)
        diff = torch.where(self.exponent != exponent, self.exponent - exponent, torch.zeros_like(exponent))
        return ScaleArray(self.significand * torch.exp(diff), exponent)

    @staticmethod
    def from_value(x) -> "ScaleArray":
        if isinstance(x, ScaleArray):
            return x
        if isinstance(x, LogArray):
            return ScaleArray(x.sign, x.logabs)
        x = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
        zero = torch.zeros((), dtype=_real_dtype(x.dtype) if x.is_floating_point() or x.is_complex() else torch.float64,
                           device=x.device)
        return ScaleArray(x, zero).normalize()

    def _other(self, other) -> "ScaleArray":
        if isinstance(other, (LogArray, ScaleArray)):
            return ScaleArray.from_value(other)
        return ScaleArray.from_value(_t(other, self.exponent))

    def value(self) -> torch.Tensor:
        return self.significand * torch.exp(self.exponent)

    def __getitem__(self, idx):
        return ScaleArray(self.significand[idx], self.exponent[idx] if self.exponent.ndim else self.exponent)

    def __len__(self):
        return self.significand.shape[0]

    def reshape(self, *shape):
        return ScaleArray(self.significand.reshape(*shape),
                          self.exponent.reshape(*shape) if self.exponent.ndim else self.exponent)

    def flatten(self):
        return ScaleArray(self.significand.flatten(), self.exponent.flatten() if self.exponent.ndim else self.exponent)

    def __neg__(self):
        return ScaleArray(-self.significand, self.exponent)

    def conj(self):
        return ScaleArray(torch.conj(self.significand).resolve_conj(), self.exponent)

    def abs(self):
        return ScaleArray(self.significand.abs(), self.exponent)

    __abs__ = abs

    @property
    def real(self):
        return ScaleArray(self.significand.real if self.significand.is_complex() else self.significand, self.exponent)

    @property
    def imag(self):
        sig = self.significand.imag if self.significand.is_complex() else torch.zeros_like(self.significand)
        return ScaleArray(sig, self.exponent)

    def astype(self, dtype):
        return ScaleArray(self.significand.to(dtype), self.exponent.to(_real_dtype(dtype)))

    def __mul__(self, other):
        from .nn import SignPhase

        if isinstance(other, SignPhase):
            if not self.significand.is_complex():
                raise TypeError("a phase layer needs a complex-output model (out_dtype=torch.complex128)")
            return ScaleArray(other.apply_(self.significand.contiguous()), self.exponent)
        o = self._other(other)
        return ScaleArray(self.significand * o.significand, self.exponent + o.exponent)

    __rmul__ = __mul__

    def __truediv__(self, other):
        o = self._other(other)
        return ScaleArray(self.significand / o.significand, self.exponent - o.exponent)

    def __rtruediv__(self, other):
        o = self._other(other)
        return ScaleArray(o.significand / self.significand, o.exponent - self.exponent)

    def __pow__(self, p):
        p = _t(p, self.exponent)
        return ScaleArray(torch.pow(self.significand, p), self.exponent * p)

    def __add__(self, other):
        o = self._other(other)
        exponent, significand = _addexp(self.exponent, o.exponent, self.significand, o.significand)
        return ScaleArray(significand, exponent)

    __radd__ = __add__

    def __sub__(self, other):
        return self.__add__(-self._other(other))

    def __rsub__(self, other):
        return self._other(other).__add__(-self)

    def _reduce(self, axis, keepdims, mean):
        if self.exponent.ndim == 0:
            red = torch.mean if mean else torch.sum
            if axis is None:
                sig = red(self.significand)
                sig = sig.reshape([1] * self.ndim) if keepdims else sig
            else:
                sig = red(self.significand, dim=axis, keepdim=keepdims)
            return ScaleArray(sig, self.exponent)
        if self.exponent.shape != self.significand.shape:
            raise ValueError(f"Cannot reduce ScaleArray with significand shape {tuple(self.significand.shape)} "
                             f"and exponent shape {tuple(self.exponent.shape)}")
        exponent, significand = _sumexp(self.exponent, self.significand, axis, keepdims, mean)
        return ScaleArray(significand, exponent)

    def sum(self, axis=None, keepdims: bool = False):
        return self._reduce(axis, keepdims, mean=False)

    def mean(self, axis=None, keepdims: bool = False):
        return self._reduce(axis, keepdims, mean=True)

    def prod(self, axis=None, keepdims: bool = False):
        """big_array.py:654-685 for an exponent of the significand's shape."""
        if self.exponent.shape != self.significand.shape:
            raise NotImplementedError("ScaleArray.prod with a scalar exponent is not implemented")
        sign, logabs = torch.sgn(self.significand), torch.log(self.significand.abs())
        finite = torch.isfinite(logabs)
        logabs = torch.where(finite, logabs, torch.zeros_like(logabs))
        sig = torch.where(finite, sign, self.significand)
        exponent = self.exponent + logabs
        if axis is None:
            sig, exponent = torch.prod(sig.flatten()), torch.sum(exponent)
            if keepdims:
                sig, exponent = sig.reshape([1] * self.ndim), exponent.reshape([1] * self.ndim)
            return ScaleArray(sig, exponent)
        return ScaleArray(torch.prod(sig, dim=axis, keepdim=keepdims), torch.sum(exponent, dim=axis, keepdim=keepdims))

    def __array__(self, dtype=None):
        return np.asarray(self.value().cpu().numpy(), dtype)


PsiArray = (torch.Tensor, LogArray, ScaleArray)  # quantax/utils/big_array.py:692


def where(cond, x, y):
    """Element-wise selection that keeps the container type (big_array.py:749-767)."""
    if isinstance(x, ScaleArray) or isinstance(y, ScaleArray):
        ref = x if isinstance(x, ScaleArray) else y
        x, y = ref._other(x), ref._other(y)
        exponent = torch.where(cond, x.exponent, y.exponent)
        significand = torch.where(cond, x.significand, y.significand)
        max_exp = exponent.max()
        return ScaleArray(significand * torch.exp(exponent - max_exp), max_exp)
    if isinstance(x, LogArray) or isinstance(y, LogArray):
        ref = x if isinstance(x, LogArray) else y
        x, y = ref._other(x), ref._other(y)
        return LogArray(torch.where(cond, x.sign, y.sign), torch.where(cond, x.logabs, y.logabs))
    return torch.where(cond, x, y)


def log_abs(psi) -> torch.Tensor:
    """log|psi| of either container (used for reweighting, sampler.py:66-69)."""
    return torch.log(psi.mult.abs()) + psi.expo


class DataTracer:
    """quantax/utils/data.py:8-147 (append / mean / uncertainty / save; plotting omitted)."""

    def __init__(self):
        self._data = []
        self._time = []

    def append(self, data, time: Optional[float] = None):
        self._data.append(float(data))
        self._time.append(len(self._data) - 1 if time is None else time)

    @property
    def data(self):
        return np.asarray(self._data)

    def __len__(self):
        return len(self._data)

    def __getitem__(self, idx):
        return self.data[idx]

    def mean(self, start=None, end=None):
        return float(np.mean(self.data[start:end]))

    def uncertainty(self, start=None, end=None):
        d = self.data[start:end]
        return float(np.std(d) / np.sqrt(max(len(d), 1)))

    def save(self, file):
        np.save(file, self.data)


def rand_states(ns: Optional[int] = None, seed: Optional[int] = None) -> torch.Tensor:
    """Random basis states on the device (quantax/utils/basis.py:121-134,160-219): uniform +-1 when
    the particle number is not conserved, a random permutation of Nup ups otherwise."""
    from .global_defs import device, get_sites, get_subkeys

    sites = get_sites()
    n = 1 if ns is None else ns
    gen = torch.Generator(device="cpu")
    gen.manual_seed((get_subkeys() if seed is None else seed) & 0x7FFFFFFFFFFFFFFF)
    N = sites.Nmodes
    if isinstance(sites.Nparticles, int):
        s = torch.randint(0, 2, (n, N), generator=gen, dtype=torch.int8) * 2 - 1
    else:
        nup = sites.Nparticles[0]
        order = torch.rand((n, N), generator=gen).argsort(dim=1)
        s = torch.where(order < nup, 1, -1).to(torch.int8)
    s = s.to(device())
    return s[0] if ns is None else s


# ---- equinox leaf serialisation (eqx.tree_serialise_leaves / tree_deserialise_leaves) -----------------------
# quantax saves a state with eqx.tree_serialise_leaves(file, model) (quantax/state/variational.py:581-587): every
# array leaf, in jax.tree flatten order, is written back to back with np.save; Python bool / int / float leaves
# (non-static dataclass fields such as ``holomorphic`` or ``nblocks``) are written the same way as 0-d arrays;
# callables, dtypes and None are skipped.  (Restated from equinox's published behaviour, >= 0.11.4 -- equinox is
# not installable here, so this interchange format is unpinned, see DESIGN.md.)
def write_eqx_leaves(file, arrays, scalars=()):
    def dump(f):
        for a in arrays:
            np.save(f, np.asarray(a))
        for v in scalars:
            np.save(f, np.asarray(v))

    if hasattr(file, "write"):
        dump(file)
    else:
        with open(file, "wb") as f:
            dump(f)


def read_eqx_leaves(file):
    """All leaves of an equinox leaf file, in order (arrays and 0-d scalars)."""
    def load(f):
        out = []
        while True:
            pos = f.tell()
            if not f.read(1):
                break
            f.seek(pos)
            out.append(np.load(f, allow_pickle=False))
        return out

    if hasattr(file, "read"):
        return load(file)
    with open(file, "rb") as f:
        return load(f)


# ---- chunk_map (quantax/utils/function.py:12-146) -----------------------------------------------------------------
def _chunk_split(x: torch.Tensor, axis: int, chunk_size: int):
    """The reference's chunk composition on ONE device (this process owns one GPU, so the `devices` axis of
    utils/function.py:28-39 has length 1): a batch longer than ``chunk_size`` is zero-padded to a multiple of it,
    viewed as [chunk_size, nchunks] and chunk i takes column i -- samples i, i + nchunks, i + 2 nchunks, ... --
    otherwise the whole batch is the only chunk.  Returns [nchunks, ..., chunk, ...]."""
    n = x.shape[axis]
    if n > chunk_size:
        pad = (-n) % chunk_size
        if pad:
            shape = list(x.shape)
            shape[axis] = pad
            x = torch.cat([x, torch.zeros(shape, dtype=x.dtype, device=x.device)], dim=axis)
        before, after = x.shape[:axis], x.shape[axis + 1:]
        x = x.reshape(*before, chunk_size, -1, *after)
    else:
        before, after = x.shape[:axis], x.shape[axis + 1:]
        x = x.reshape(*before, n, 1, *after)
    return torch.movedim(x, axis + 1, 0)


def _chunk_combine(x: torch.Tensor, axis: int, batch: int) -> torch.Tensor:
    """Inverse of _chunk_split on stacked chunk outputs [nchunks, ..., chunk, ...] (utils/function.py:69-76): the
    chunk axis goes back behind the in-chunk axis, padding rows are cut."""
    x = torch.movedim(x, axis + 1, 0)  # [chunk, nchunks, ...]
    rest = x.shape[2:]
    x = x.reshape(-1, *rest)[:batch]
    return torch.movedim(x, 0, axis)


def chunk_map(f, in_axes=0, out_axes=0, chunk_size: Optional[int] = None, use_scan: bool = False):
    """``chunk_map`` of the reference (utils/function.py:88-146) for torch tensors: a per-sample (vmapped) function is
    evaluated chunk by chunk with the reference's interleaved chunk composition and zero padding, and the outputs are
    re-assembled in sample order.  ``use_scan`` is accepted for signature parity (there is no tracing here)."""
    all_none = isinstance(in_axes, (tuple, list)) and all(a is None for a in in_axes)
    if in_axes is None or all_none or chunk_size is None:
        return f
    if out_axes is None or (isinstance(out_axes, (tuple, list)) and any(a is None for a in out_axes)):
        raise NotImplementedError("`chunk_map` with `out_axes=None` not implemented")

    def chunked_f(*args):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        batch = next(a.shape[ax] for a, ax in zip(args, axes) if ax is not None)
        split = [None if ax is None else _chunk_split(a, ax, chunk_size) for a, ax in zip(args, axes)]
        nchunks = next(s.shape[0] for s in split if s is not None)
        outs = []
        for i in range(nchunks):
            outs.append(f(*[a if s is None else s[i] for a, s in zip(args, split)]))
        is_tuple = isinstance(outs[0], tuple)
        if not is_tuple:
            outs = [(o,) for o in outs]
        oaxes = tuple(out_axes) if isinstance(out_axes, (tuple, list)) else (out_axes,) * len(outs[0])
        res = tuple(_chunk_combine(torch.stack([o[k] for o in outs], dim=0), oaxes[k], batch) for k in range(len(outs[0])))
        return res if is_tuple else res[0]

    return chunked_f
