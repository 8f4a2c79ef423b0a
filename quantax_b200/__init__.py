"""quantax_b200 -- B200-native implementation of the quantax VMC step hot path.

Drop-in for the path ``sampler.sweep() -> optimizer.get_step(samples) -> state.update(step)`` of
ChenAo-Phys/quantax (README.md:42-76): same sub-module names and class signatures
(``sites``, ``operator``, ``model``, ``state``, ``sampler``, ``optimizer``, ``utils``, ``nn``),
arrays are CUDA ``torch.Tensor`` objects, all arithmetic runs in hand-written sm_100a kernels
behind the C ABI of ``include/qtx_b200.h``.  There is no CPU fallback.
"""
from . import global_defs, sites, symmetry, utils, nn, operator, model, state, sampler, optimizer  # noqa: F401
from .global_defs import (  # noqa: F401
    PARTICLE_TYPE,
    get_default_dtype,
    get_lattice,
    get_real_dtype,
    get_sites,
    get_subkeys,
    is_default_cpl,
    set_default_dtype,
    set_random_seed,
)

__version__ = "0.1.0"
