"""Variational ansaetze of the hot path: RBM_Dense and ResConv parameter containers.

Mirrors quantax/model/shallow_nets.py:111-126 (RBM_Dense) and quantax/model/conv_nets.py:95-183
(ResConv).  A model owns ONE flat parameter vector on the device, laid out in the reference's
``ravel_pytree`` order (quantax/state/variational.py:236-243), so the Jacobian columns, the SR
step and ``Variational.update`` address parameters exactly as the reference does:
  RBM_Dense : [W.ravel() (M x N row-major), b (M)]
  ResConv   : per block conv1.weight [C,Cin,kh,kw], conv1.bias [C], conv2.weight, conv2.bias
              (the last conv has no bias, conv_nets.py:66,75).
All arithmetic happens in the CUDA kernels; these classes only hold parameters and metadata.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .global_defs import device, get_lattice, get_sites, get_subkeys

_TRUNC_STD = 0.87962566103423978  # std of a unit normal truncated to [-2, 2] (jax.nn.initializers)


def _truncated_normal(rng, shape):
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2
    return out / _TRUNC_STD


def _get_scale(features: int, nsites: int) -> float:
    """Scalar in [0, 0.99] such that std(sum_i log|cosh(scale*x_i)|) over 1000 Gaussian inputs is
    closest to 0.1*sqrt(N) (quantax/model/shallow_nets.py:17-32; the jax key(0) stream is not
    reproduced, a fixed NumPy stream is used instead)."""
    x = np.random.default_rng(0).standard_normal((1000, features))
    target = 0.1 * np.sqrt(nsites)
    best, arg = np.inf, 0.0
    for scale in np.arange(0, 1, 0.01):
        out = np.sum(np.log(np.abs(np.cosh(x * scale))), axis=1)
        err = (np.std(out) - target) ** 2
        if err < best:
            best, arg = err, scale
    return float(arg)


class RBM_Dense:
    r"""psi(s) = prod_i cosh(W s + b)  (quantax/model/shallow_nets.py:111-126)."""

    is_ref_model = True  # quantax.nn.RefModel: supports init_internal / ref_forward
    kind = "rbm"

    def __init__(self, features: int, use_bias: bool = True, dtype=torch.float32, params: Optional[torch.Tensor] = None):
        if dtype not in (torch.float32, torch.float64):
            raise NotImplementedError("complex RBM parameters are outside the B200 hot path")
        if not use_bias:
            raise NotImplementedError("RBM_Dense(use_bias=False) is not implemented")
        sites = get_sites()
        self.N = sites.Nmodes
        self.M = int(features)
        self.dtype = dtype
        self.holomorphic = False
        if params is None:
            rng = np.random.default_rng(get_subkeys() & 0xFFFFFFFF)
            w = _truncated_normal(rng, (self.M, self.N)) / np.sqrt(self.N)  # LeCun normal, fan_in = N
            w *= _get_scale(self.M, sites.Nsites)
            flat = np.concatenate([w.ravel(), np.zeros(self.M)])
            params = torch.from_numpy(flat).to(device=device(), dtype=dtype)
        self.params = params.to(device=device(), dtype=dtype).contiguous()
        assert self.params.numel() == self.M * self.N + self.M

    @property
    def nparams(self) -> int:
        return self.M * self.N + self.M

    @property
    def layer_param_sizes(self):
        """Parameter count of every ``Sequential`` layer that has parameters (block_pinv_eig, solver.py:241-249):
        one Linear layer."""
        return [self.nparams]

    def eqx_leaf_layout(self):
        """Array leaves in equinox order: Linear.weight [M, N], Linear.bias [M] (shallow_nets.py:71)."""
        return [("linear.weight", 0, (self.M, self.N)), ("linear.bias", self.M * self.N, (self.M,))]

    def eqx_trailing_scalars(self):
        return [False]  # Sequential.holomorphic (nn/modules.py:20-21)

    @property
    def W(self) -> torch.Tensor:
        return self.params[: self.M * self.N].view(self.M, self.N)

    @property
    def b(self) -> torch.Tensor:
        return self.params[self.M * self.N:]


class RBM_Conv:
    r"""psi(s) = prod cosh(Conv(s)) with one full-lattice circular convolution (quantax/model/shallow_nets.py:129-190).
    Evaluated as a dense RBM with M = channels * N tied hidden units: ``W`` / ``b`` are the expanded weights
    (rebuilt from the convolution kernel on every access, qtx_rbm_conv_expand), so the fused sweep / Oloc / forward
    kernels of RBM_Dense apply; the log-derivative has its own kernel (qtx_rbm_conv_jacobian)."""

    is_ref_model = True  # local updates of the equivalent dense RBM (same amplitudes as the full forward)
    kind = "rbm"
    tied = True

    def __init__(self, channels: int, use_bias: bool = True, dtype=torch.float32, params: Optional[torch.Tensor] = None):
        if dtype not in (torch.float32, torch.float64):
            raise NotImplementedError("complex RBM parameters are outside the B200 hot path")
        if not use_bias:
            raise NotImplementedError("RBM_Conv(use_bias=False) is not implemented")
        lattice = get_lattice()
        if lattice.shape[0] != 1 or lattice.ndim > 2:
            raise NotImplementedError("RBM_Conv is implemented for 1-D and 2-D lattices with one site per cell")
        if not all(bc != 0 for bc in lattice.boundary):
            raise NotImplementedError("open boundaries are outside the B200 hot path")
        ext = lattice.shape[1:]
        self.Lx, self.Ly = (1, ext[0]) if len(ext) == 1 else (ext[0], ext[1])
        self.channels = int(channels)
        self.N = lattice.Nsites
        self.M = self.channels * self.N
        self.dtype = dtype
        self.holomorphic = False
        if params is None:
            rng = np.random.default_rng(get_subkeys() & 0xFFFFFFFF)
            w = _truncated_normal(rng, (self.channels, self.N)) / np.sqrt(self.N)  # LeCun normal, fan_in = 1 * prod(kernel)
            w *= _get_scale(self.M, lattice.Nsites)
            params = torch.from_numpy(np.concatenate([w.ravel(), np.zeros(self.channels)]))
        self.params = params.to(device=device(), dtype=dtype).contiguous()
        assert self.params.numel() == self.nparams
        self._W = torch.empty((self.M, self.N), dtype=dtype, device=self.params.device)
        self._b = torch.empty(self.M, dtype=dtype, device=self.params.device)

    @property
    def nparams(self) -> int:
        return self.channels * self.N + self.channels

    @property
    def layer_param_sizes(self):
        """One Conv layer carries all parameters (block_pinv_eig, solver.py:241-249)."""
        return [self.nparams]

    def _expand(self):
        from . import _lib

        _lib.call("qtx_rbm_conv_expand", _lib.dtype_code(self.dtype), _lib.ptr(self.params),
                  _lib.ptr(self.params[self.channels * self.N:]), self.channels, self.Lx, self.Ly, _lib.ptr(self._W),
                  _lib.ptr(self._b), _lib.stream())

    @property
    def W(self) -> torch.Tensor:
        self._expand()
        return self._W

    @property
    def b(self) -> torch.Tensor:
        return self._b  # filled by the W access that precedes every use

    def eqx_leaf_layout(self):
        """Conv.weight [C, 1, Lx, Ly] (chains: [C, 1, L]), Conv.bias [C, 1, 1] ([C, 1])."""
        one_d = self.Lx == 1
        wshape = (self.channels, 1, self.Ly) if one_d else (self.channels, 1, self.Lx, self.Ly)
        bshape = (self.channels, 1) if one_d else (self.channels, 1, 1)
        return [("conv.weight", 0, wshape), ("conv.bias", self.channels * self.N, bshape)]

    def eqx_trailing_scalars(self):
        return [False]


class ResConv:
    """Deep convolutional residual network (quantax/model/conv_nets.py:95-183)."""

    is_ref_model = False
    kind = "resconv"

    def __init__(self, nblocks: int, channels: int, kernel_size: int, final_activation=None, trans_symm=None,
                 dtype=torch.float32, out_dtype=None, params: Optional[torch.Tensor] = None):
        from . import nn

        if dtype not in (torch.float32, torch.float64):
            raise ValueError("`ResSum` doesn't support complex dtypes.")
        if out_dtype is None:
            out_dtype = dtype
        if out_dtype not in (dtype, torch.complex128):
            raise NotImplementedError("out_dtype must equal dtype or be torch.complex128")
        self.out_dtype = out_dtype
        self.cplx = out_dtype == torch.complex128  # pair_cpl before the final activation (conv_nets.py:167-168)
        if self.cplx and channels % 2:
            raise ValueError("a complex-output ResConv needs an even number of channels (pair_cpl)")
        self.raw_layers = ()
        if trans_symm is not None:
            raise NotImplementedError("only the default translation symmetry (sector 0) is implemented")
        lattice = get_lattice()
        if not all(bc != 0 for bc in lattice.boundary):
            raise NotImplementedError("open boundaries are outside the B200 hot path")
        if lattice.ndim > 2:
            raise NotImplementedError("3D lattices are not implemented")
        if isinstance(kernel_size, (tuple, list)):
            raise NotImplementedError("anisotropic kernels are not implemented")
        if kernel_size % 2 != 1:
            raise NotImplementedError("even kernel sizes are not implemented")
        self.nblocks, self.channels, self.kernel_size = int(nblocks), int(channels), int(kernel_size)
        if final_activation is None:
            final_activation = nn.exp_by_scale
        if final_activation is nn.exp_by_scale:
            self.final = 0
        elif final_activation is nn.sinhp1_by_scale:
            self.final = 1
        else:
            raise NotImplementedError("final_activation must be exp_by_scale or sinhp1_by_scale")
        self.final_activation = final_activation
        self.dtype = dtype
        self.holomorphic = False
        ext = lattice.shape[1:]
        self.Lx, self.Ly = (1, ext[0]) if len(ext) == 1 else (ext[0], ext[1])
        self.kh = 1 if self.Lx == 1 and len(ext) == 1 else self.kernel_size
        self.kw = self.kernel_size
        self.N = lattice.Nsites
        # parameter layout
        self.layout = []  # (name, offset, shape)
        off = 0
        for i in range(self.nblocks):
            for name, cin, last in (("conv1", 1 if i == 0 else self.channels, False),
                                    ("conv2", self.channels, i == self.nblocks - 1)):
                shape = (self.channels, cin, self.kh, self.kw)
                self.layout.append((f"block{i}.{name}.weight", off, shape))
                off += int(np.prod(shape))
                if not last:
                    self.layout.append((f"block{i}.{name}.bias", off, (self.channels,)))
                    off += self.channels
        self._nparams = off
        if params is None:
            rng = np.random.default_rng(get_subkeys() & 0xFFFFFFFF)
            flat = np.zeros(off)
            for name, o, shape in self.layout:
                if name.endswith("weight"):
                    fan_in = shape[1] * shape[2] * shape[3]  # He normal (conv_nets.py:71)
                    flat[o:o + int(np.prod(shape))] = (_truncated_normal(rng, shape) * np.sqrt(2.0 / fan_in)).ravel()
            params = torch.from_numpy(flat)
        self.params = params.to(device=device(), dtype=dtype).contiguous()
        assert self.params.numel() == off

    @property
    def nparams(self) -> int:
        return self._nparams

    @property
    def layer_param_sizes(self):
        """Parameter count per ``Sequential`` layer with parameters = per residual block (conv_nets.py:164-181:
        ReshapeConv, the blocks, final_layer and ConvSymmetrize; only the blocks have parameters), in flat order."""
        sizes = [0] * self.nblocks
        for name, _, shape in self.layout:
            sizes[int(name.split(".")[0][len("block"):])] += int(np.prod(shape))
        return sizes

    def eqx_leaf_layout(self):
        """(name, offset, shape) of the array leaves as equinox stores them: Conv.weight [C, Cin, kh, kw] (2-D lattices)
        or [C, Cin, kw] (chains), Conv.bias [C, 1, 1] / [C, 1]."""
        one_d = self.Lx == 1 and self.kh == 1
        out = []
        for name, o, shape in self.layout:
            if name.endswith("weight"):
                out.append((name, o, (shape[0], shape[1], shape[3]) if one_d else shape))
            else:
                out.append((name, o, (shape[0], 1) if one_d else (shape[0], 1, 1)))
        return out

    def eqx_trailing_scalars(self):
        """Non-array leaves equinox serialises after the layers: ``holomorphic`` for every Sequential; a bare ResConv
        also carries nblocks, channels, kernel_size (dataclass field order, conv_nets.py:98-106)."""
        if getattr(self, "raw_layers", ()):
            return [False]
        return [False, self.nblocks, self.channels, self.kernel_size]

    @property
    def layers(self):
        """The model as a tuple of layers, to be extended with RawInputLayers and re-assembled by
        ``quantax_b200.nn.Sequential`` (tutorials/triangular.ipynb:120-128)."""
        return (self,)

    def named_parameters(self):
        for name, o, shape in self.layout:
            yield name, self.params[o:o + int(np.prod(shape))].view(shape)
