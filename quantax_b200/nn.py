"""Final activations selectable for ResConv (quantax/nn/activation.py:7-32).  The functions are
markers: the arithmetic is fused into the CUDA forward / backward kernels."""


def _as_tensor(x):
    import torch

    return x if torch.is_tensor(x) else torch.as_tensor(x)


def exp_by_scale(x):
    r"""f(x) = exp(x) as a ``ScaleArray`` with the scalar exponent max|x| (quantax/nn/activation.py:26-32).  Callable
    on any tensor like the reference's; inside ``ResConv`` the same function is evaluated by the CUDA kernels, which
    recognise it by identity (``ResConv(final_activation=exp_by_scale)``)."""
    import torch

    from .utils import ScaleArray

    x = _as_tensor(x)
    xmax = torch.nan_to_num(x.abs(), nan=0.0).max() if x.numel() else x.new_zeros(())
    return ScaleArray(torch.exp(x - xmax), xmax)


def sinhp1_by_scale(x):
    r"""f(x) = sinh(x) + 1 as a ``ScaleArray`` (quantax/nn/activation.py:7-14); see ``exp_by_scale``."""
    import torch

    from .utils import ScaleArray

    x = _as_tensor(x)
    xmax = torch.nan_to_num(x.abs(), nan=0.0).max() if x.numel() else x.new_zeros(())
    return ScaleArray((torch.exp(x - xmax) - torch.exp(-x - xmax)) / 2 + torch.exp(-xmax), xmax)


def pair_cpl(x):
    r"""f(x) = x_1 + i x_2 for x = (x_1, x_2) split along the first axis (quantax/nn/activation.py:75-81)."""
    import torch

    x = _as_tensor(x)
    h = x.shape[0] // 2
    return torch.complex(x[:h], x[h:])


class RawInputLayer:
    """Layer that also receives the raw spins: ``layer(x, s)`` (quantax/nn/modules.py:64-77).  Here ``x`` is the
    batched amplitude container and ``s`` the int8 spin batch [ns, N]."""

    def __call__(self, x, s):
        raise NotImplementedError


class SignPhase:
    """exp(i * dot(kernel, s)) for a batch of spins (quantax/nn/sign.py:8-43, output="phase"), evaluated lazily:
    ``ScaleArray * SignPhase`` multiplies the significands in place on the device (qtx_apply_sign_phase)."""

    def __init__(self, kernel, spins):
        self.kernel, self.spins = kernel, spins

    def apply_(self, mult):
        from . import _lib

        s = self.spins
        _lib.call("qtx_apply_sign_phase", _lib.ptr(self.kernel), _lib.ptr(s), s.shape[0], s.shape[1], _lib.ptr(mult),
                  _lib.stream())
        return mult

    def tensor(self):
        import torch

        one = torch.ones(self.spins.shape[0], dtype=torch.complex128, device=self.spins.device)
        return self.apply_(one)


def neel120_phase(s) -> SignPhase:
    """120-degree Neel phase for triangular lattices (quantax/nn/sign.py:62-75)."""
    import numpy as np
    import torch

    from .global_defs import device, get_lattice
    from .operator import _as_spins

    lattice = get_lattice()
    Lx, Ly = lattice.shape[1:]
    x = 2 * np.arange(Lx)
    y = np.zeros(Ly, dtype=x.dtype) if type(lattice).__name__ == "TriangularB" else np.arange(Ly)
    kernel = (x[:, None] + y[None, :]) % 3
    kernel = (np.pi / 3 * kernel - np.pi / 6).astype(np.float32).ravel()
    return SignPhase(torch.from_numpy(kernel).to(device()), _as_spins(s))


def Sequential(layers):
    """``qtx.nn.Sequential(model.layers + (MyRawInputLayer(),))`` (tutorials/triangular.ipynb:120-128): the model
    followed by parameter-free layers acting on (psi, s)."""
    import copy

    layers = tuple(layers)
    if not layers or not hasattr(layers[0], "kind"):
        raise NotImplementedError("Sequential expects the layers of one of this package's models first")
    extra = layers[1:]
    if not all(isinstance(l, RawInputLayer) for l in extra):
        raise NotImplementedError("only RawInputLayer instances can follow the model")
    model = copy.copy(layers[0])
    model.raw_layers = tuple(getattr(layers[0], "raw_layers", ())) + tuple(extra)
    return model
