"""Package-wide dtype / seed / sites globals, mirroring quantax/global_defs.py:13-137.

The reference keeps a jax threefry key and splits it for every consumer; here the global state
is a 64-bit Philox seed plus a monotonically increasing step counter (see sampler.py).
"""
from __future__ import annotations

from enum import Enum

import torch

DTYPE = torch.float64  # quantax/global_defs.py:13
_SEED = 42  # quantax/global_defs.py:88-89
_SUBKEY_COUNTER = 0


def set_default_dtype(dtype) -> None:
    """quantax/global_defs.py:16-30.  complex128 selects complex-output states with real parameters
    (VS_TYPE.real_to_complex); complex64 is not provided."""
    global DTYPE
    if dtype == torch.complex64:
        raise NotImplementedError("complex64 default dtype is not implemented (use complex128)")
    if dtype not in (torch.float32, torch.float64, torch.complex128):
        raise ValueError("'dtype' should be float or complex types")
    DTYPE = dtype


def get_default_dtype():
    return DTYPE


def get_real_dtype():
    return torch.float64 if DTYPE == torch.complex128 else DTYPE


def is_default_cpl() -> bool:
    return DTYPE == torch.complex128


def set_random_seed(seed: int) -> None:
    """quantax/global_defs.py:54-63."""
    global _SEED, _SUBKEY_COUNTER
    _SEED = int(seed)
    _SUBKEY_COUNTER = 0


def get_seed() -> int:
    return _SEED


def get_subkeys(num=None):
    """Stand-in for quantax/global_defs.py:75-91: returns fresh 64-bit sub-seeds derived from the
    global seed (splitmix64 of seed and a running counter)."""
    global _SUBKEY_COUNTER

    def mix(x):
        x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return x ^ (x >> 31)

    n = 1 if num is None else num
    keys = []
    for _ in range(n):
        _SUBKEY_COUNTER += 1
        keys.append(mix(mix(_SEED) ^ _SUBKEY_COUNTER))
    return keys[0] if num is None else keys


class PARTICLE_TYPE(Enum):
    """quantax/global_defs.py:94-109."""

    spin = 0
    spinful_fermion = 1
    spinless_fermion = 2


def get_sites():
    """quantax/global_defs.py:115-126."""
    from .sites import Sites

    if Sites._SITES is None:
        raise RuntimeError("The `Sites` hasn't been defined.")
    return Sites._SITES


def get_lattice():
    """quantax/global_defs.py:129-137."""
    from .sites import Lattice

    sites = get_sites()
    if not isinstance(sites, Lattice):
        raise RuntimeError("Require a `Lattice`, but got a general `Sites`")
    return sites


def device():
    """The CUDA device of this process (one process per GPU)."""
    return torch.device("cuda", torch.cuda.current_device())


def world():
    """(rank, world_size) of the data-parallel job; (0, 1) without torch.distributed."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


# ---- NVTX ranges per phase of the VMC step (QTX_NVTX=1; SURVEY section 5: tracing) --------------------------------
import contextlib as _contextlib
import os as _os

NVTX = _os.environ.get("QTX_NVTX", "0") == "1"


@_contextlib.contextmanager
def nvtx_range(name: str):
    """``with nvtx_range("sweep"):`` -- an NVTX range around a phase when QTX_NVTX=1 (ncu / nsys can filter on it),
    nothing otherwise."""
    if not NVTX:
        yield
        return
    import torch

    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
