"""Lattice symmetries: permutation groups, characters, composition.

Mirrors quantax/symmetry/: ``Symmetry`` (symmetry.py:87-432), ``Identity`` / ``Z2Inversion`` /
``SpinInverse`` / ``LinearTransform`` / ``Flip`` / ``Rotation`` / ``C4v`` / ``D6``
(common_symmetries.py) and ``Translation`` / ``TransND`` (translation.py).  These are host tables
(NumPy); the projection psi(s) = sum_g chi_g psi(T_g s) / |G| itself runs in the CUDA kernels
``qtx_symm_images`` / ``qtx_symm_combine`` / ``qtx_weighted_rowsum`` driven by
``state.Variational``.  Spin systems and real characters (real default dtype) only; the quspin
basis bridge is outside the hot path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np

from .global_defs import get_lattice, get_sites


def _closure(generator: np.ndarray, sector) -> tuple:
    """Group elements and characters from commuting generators (symmetry.py:11-57)."""
    nmodes = generator.shape[1]
    identity = np.arange(nmodes)
    perm, char = identity[None, :], np.ones(1)
    for g, sec in zip(generator, sector):
        orbit, cur = [identity], g
        while not np.array_equal(cur, identity):
            orbit.append(cur)
            cur = cur[g]
        orbit = np.stack(orbit)
        order = orbit.shape[0]
        if not 0 <= sec < order:
            raise ValueError(f"Sector {sec} out of range.")
        if (2 * sec) % order != 0:
            raise ValueError("Default dtype is real, but got complex characters.")
        chi = (-1.0 if sec else 1.0) ** np.arange(order)
        perm = perm[:, orbit].reshape(-1, nmodes)
        char = np.outer(char, chi).ravel()
    return perm, char


class Symmetry:
    def __init__(self, generator: Optional[np.ndarray] = None, sector: Union[int, Sequence] = 0,
                 generator_sign: Optional[np.ndarray] = None, Z2_inversion: int = 0,
                 perm: Optional[np.ndarray] = None, character: Optional[np.ndarray] = None, perm_sign=None):
        sites = get_sites()
        self._Nmodes = sites.Nmodes
        if generator_sign is not None and not np.all(np.asarray(generator_sign) == 1):
            raise NotImplementedError("generator signs (anti-periodic fermions) are outside the spin hot path")
        if generator is None:
            generator = np.arange(self._Nmodes)[None, :]
        generator = np.atleast_2d(np.asarray(generator, dtype=np.int64))
        if generator.shape[1] != self._Nmodes:
            raise ValueError(f"Got a generator with size {generator.shape[1]}, incompatible with the system size "
                             f"{self._Nmodes}.")
        self._generator = generator
        self._sector = [sector] * generator.shape[0] if isinstance(sector, (int, np.integer)) else \
            np.asarray(sector).ravel().tolist()
        if Z2_inversion not in (0, 1, -1):
            raise ValueError("Z2_inversion should be 0, 1 or -1")
        self._Z2_inversion = Z2_inversion
        if perm is None or character is None:
            p, c = _closure(generator, self._sector)
            perm = p if perm is None else perm
            character = c if character is None else character
        self._perm = np.ascontiguousarray(perm, dtype=np.int32)
        self._character = np.asarray(character, dtype=np.float64)
        self._device_tables = None

    Nmodes = property(lambda self: self._Nmodes)
    Nsites = property(lambda self: self._Nmodes)
    character = property(lambda self: self._character)
    Z2_inversion = property(lambda self: self._Z2_inversion)
    perm = property(lambda self: self._perm)

    @property
    def nsymm(self) -> int:
        n = self._character.size
        return n if self._Z2_inversion == 0 else 2 * n

    @property
    def is_identity(self) -> bool:
        return self.nsymm == 1 and np.array_equal(self._perm[0], np.arange(self._Nmodes))

    def weights(self) -> np.ndarray:
        """chi_g chi_0 / nsymm including the Z2 block (symmetry.py:389-391)."""
        c = self._character
        if self._Z2_inversion != 0:
            c = np.concatenate([c, self._Z2_inversion * c])
        return c * c[0] / c.size

    def get_symm_spins(self, spins: np.ndarray) -> np.ndarray:
        """Host version of symmetry.py:325-341 for a single configuration."""
        spins = np.asarray(spins)
        if spins.ndim > 1:
            raise ValueError(f"Input spins should be 1D, got dimension {spins.ndim}")
        out = spins[self._perm]
        return np.concatenate([out, -out], axis=-2) if self._Z2_inversion != 0 else out

    def __matmul__(self, other: "Symmetry") -> "Symmetry":
        """Superposition of two symmetries (symmetry.py:394-432)."""
        perm = self._perm[:, other._perm].reshape(-1, self._Nmodes)
        character = np.outer(self._character, other._character).ravel()
        if self._Z2_inversion == 0:
            z2 = other._Z2_inversion
        elif other._Z2_inversion in (0, self._Z2_inversion):
            z2 = self._Z2_inversion
        else:
            raise ValueError("Symmetry with different Z2_inversion can't be added")
        return Symmetry(np.concatenate([self._generator, other._generator], axis=0), [*self._sector, *other._sector],
                        None, z2, perm, character)

    def device_tables(self):
        """(perm int32 [nperm, N], weights float64 [nsymm]) on the current CUDA device."""
        import torch

        from .global_defs import device

        if self._device_tables is None or self._device_tables[0].device != device():
            self._device_tables = (torch.from_numpy(self._perm).to(device()),
                                   torch.from_numpy(self.weights()).to(device()))
        return self._device_tables


# Per-lattice singletons.  The cache holds the lattice object itself: keyed by id() alone, a new lattice that happened to
# get the address of a collected one was handed the old lattice's tables (wrong Nmodes).
_SINGLETONS = {"sites": None, "identity": None, "z2": {}}


def _singletons():
    sites = get_sites()
    if _SINGLETONS["sites"] is not sites:
        _SINGLETONS.update(sites=sites, identity=None, z2={})
    return _SINGLETONS


def Identity() -> Symmetry:
    c = _singletons()
    if c["identity"] is None:
        c["identity"] = Symmetry()
    return c["identity"]


def Z2Inversion(eigval: int = 1) -> Symmetry:
    if eigval not in (1, -1):
        raise ValueError("'eigval' of Z2Inversion should be 1 or -1.")
    c = _singletons()
    if eigval not in c["z2"]:
        c["z2"][eigval] = Symmetry(Z2_inversion=eigval)
    return c["z2"][eigval]


def SpinInverse(eigval: int = 1) -> Symmetry:
    """Global spin flip (common_symmetries.py:42-71, spin systems)."""
    if eigval == 0:
        return Identity()
    return Z2Inversion(eigval)


class Translation(Symmetry):
    """translation.py:7-59."""

    def __init__(self, vectors: Sequence, sector: int = 0):
        lattice = get_lattice()
        vectors = np.asarray(vectors, dtype=np.int64).reshape(-1, lattice.ndim)
        if np.any((vectors != 0) & (lattice.boundary[None, :] == 0)):
            raise ValueError("Translation symmetry can't be imposed on open boundary.")
        extent = np.asarray(lattice.shape[1:])
        gens = []
        for vec in vectors:
            for axis, vi in enumerate(vec):
                if vi != 0 and lattice.shape[axis + 1] % vi != 0:
                    raise ValueError("Translation vector must be compatible with lattice shape, "
                                     f"got lattice shape {lattice.shape[1:]} and vector {vec}.")
            cell = (lattice.xyz_from_index[:, 1:] + vec[None, :]) % extent
            gens.append(np.ravel_multi_index(tuple(cell.T), tuple(extent)))
        self._vectors = vectors
        super().__init__(np.stack(gens), sector)

    @property
    def vectors(self) -> np.ndarray:
        return self._vectors.copy()


def TransND(sector: Union[int, Sequence] = 0) -> Symmetry:
    """Translations along every lattice basis vector (translation.py, TransND)."""
    return Translation(np.eye(get_lattice().ndim, dtype=np.int64), sector)


def _wrap_into_cell(coord: np.ndarray) -> np.ndarray:
    """common_symmetries.py:90-101."""
    lattice = get_lattice()
    frac = np.linalg.solve(lattice.basis_vectors.T, coord.T).T
    periodic = lattice.boundary != 0
    ext = np.asarray(lattice.shape[1:])[periodic]
    wrapped = frac[:, periodic] % ext
    wrapped[np.isclose(wrapped, ext)] = 0.0
    frac[:, periodic] = wrapped
    return frac @ lattice.basis_vectors


def LinearTransform(matrix: np.ndarray, center: Optional[np.ndarray] = None, sector: int = 0,
                    character: Optional[np.ndarray] = None) -> Symmetry:
    """Point-group element as a site permutation (common_symmetries.py:104-145)."""
    lattice = get_lattice()
    if center is None:
        center = lattice.coord.mean(axis=0)
    ref = _wrap_into_cell(lattice.coord)
    moved = _wrap_into_cell((lattice.coord - center) @ np.asarray(matrix).T + center)
    match = np.isclose(ref[:, None, :], moved[None, :, :]).all(axis=-1)
    if not np.all(match.sum(axis=1) == 1):
        raise ValueError("The transformation does not map the lattice to itself.")
    return Symmetry(match.argmax(axis=1), sector, character=character)


def Flip(axis: Union[int, Sequence] = 0, center: Optional[np.ndarray] = None, sector: int = 0) -> Symmetry:
    diag = np.ones(get_lattice().ndim)
    diag[np.asarray(axis)] = -1
    return LinearTransform(np.diag(diag), center, sector)


def Rotation(angle: float, axes: Sequence = (0, 1), center: Optional[np.ndarray] = None, sector: int = 0,
             character: Optional[np.ndarray] = None) -> Symmetry:
    ndim = get_lattice().ndim
    if max(axes) >= ndim:
        raise ValueError(f"The rotated axis {max(axes)} is out-of-bound for a {ndim}-D system")
    m = np.eye(ndim)
    x, y = axes
    c, s = np.cos(angle), np.sin(angle)
    m[x, x], m[x, y], m[y, x], m[y, y] = c, -s, s, c
    return LinearTransform(m, center, sector, character)


def _dihedral(angle: float, half: int, center, repr: str, names) -> Symmetry:
    table = {names[0]: (0, 0), names[1]: (0, 1), names[2]: (half, 0), names[3]: (half, 1)}
    if repr not in table:
        raise NotImplementedError(f"representation '{repr}' (multi-dimensional) is not implemented; choose from {names}")
    rot_sector, flip_sector = table[repr]
    return Rotation(angle=angle, center=center, sector=rot_sector) @ Flip(center=center, sector=flip_sector)


def C4v(center: Optional[np.ndarray] = None, repr: str = "A1") -> Symmetry:
    """common_symmetries.py:208-240 (one-dimensional representations)."""
    return _dihedral(np.pi / 2, 2, center, repr, ("A1", "A2", "B1", "B2"))


def D6(center: Optional[np.ndarray] = None, repr: str = "A1") -> Symmetry:
    """common_symmetries.py:243-278 (one-dimensional representations)."""
    return _dihedral(np.pi / 3, 3, center, repr, ("A1", "A2", "B1", "B2"))
