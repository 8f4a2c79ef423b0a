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


@dataclass
class LogArray:
    """quantax/utils/big_array.py:154-402."""

    sign: torch.Tensor
    logabs: torch.Tensor

    mult = property(lambda self: self.sign)
    expo = property(lambda self: self.logabs)
    shape = property(lambda self: self.sign.shape)
    dtype = property(lambda self: self.logabs.dtype)

    def value(self) -> torch.Tensor:
        return self.sign * torch.exp(self.logabs)

    def __getitem__(self, idx):
        return LogArray(self.sign[idx], self.logabs[idx])

    def __len__(self):
        return self.sign.shape[0]

    def abs(self):
        return LogArray(torch.ones_like(self.sign), self.logabs)

    __abs__ = abs

    def __truediv__(self, other):
        return LogArray(self.sign / other.sign, self.logabs - other.logabs)

    def __array__(self, dtype=None):
        return np.asarray(self.value().cpu().numpy(), dtype)


@dataclass
class ScaleArray:
    """quantax/utils/big_array.py:407-688."""

    significand: torch.Tensor
    exponent: torch.Tensor

    mult = property(lambda self: self.significand)
    expo = property(lambda self: self.exponent)
    shape = property(lambda self: self.significand.shape)
    dtype = property(lambda self: self.significand.dtype)

    def value(self) -> torch.Tensor:
        return self.significand * torch.exp(self.exponent)

    def __getitem__(self, idx):
        return ScaleArray(self.significand[idx], self.exponent[idx])

    def __len__(self):
        return self.significand.shape[0]

    def abs(self):
        return ScaleArray(self.significand.abs(), self.exponent)

    __abs__ = abs

    def __truediv__(self, other):
        return ScaleArray(self.significand / other.significand, self.exponent - other.exponent)

    def __mul__(self, other):
        from .nn import SignPhase

        if isinstance(other, SignPhase):
            if not self.significand.is_complex():
                raise TypeError("a phase layer needs a complex-output model (out_dtype=torch.complex128)")
            return ScaleArray(other.apply_(self.significand.contiguous()), self.exponent)
        return NotImplemented

    def __array__(self, dtype=None):
        return np.asarray(self.value().cpu().numpy(), dtype)


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
