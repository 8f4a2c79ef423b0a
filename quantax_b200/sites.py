"""Host-side lattice tables: site indexing, PBC distances, n-th neighbour bond lists.

Mirrors the public surface of quantax/sites/ that the hot path consumes
(Sites / Lattice / Grid / Chain / Square / Cube / Triangular, ``get_neighbor``): the bond ORDER
(quantax/sites/sites.py:267-283: shells by increasing distance, pairs i<j in lexicographic
order) defines the connected-configuration enumeration order of Operator.Oloc, and is checked
against tables produced by the reference's own code (tests/golden/ref_tables.npz).
"""
from __future__ import annotations

from itertools import product
from typing import Optional, Sequence, Tuple, Union
from warnings import warn

import numpy as np

from .global_defs import PARTICLE_TYPE


class Sites:
    _SITES = None  # quantax/sites/sites.py:13 -- global singleton

    def __init__(self, Nsites: int, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None, coord=None):
        if Sites._SITES is not None:
            warn("Quantax treats the `Sites` as a global constant. Defining multiple `Sites` might lead to "
                 "unexpected behaviors.")
        Sites._SITES = self
        if isinstance(particle_type, str):
            particle_type = PARTICLE_TYPE[particle_type.lower().replace(" ", "_")]
        if particle_type != PARTICLE_TYPE.spin:
            raise NotImplementedError("fermionic sites are outside the B200 hot path (spin systems only)")
        self._Nsites = int(Nsites)
        self._particle_type = particle_type
        if Nparticles is None:
            Nparticles = self._Nsites
        elif isinstance(Nparticles, int):
            if Nparticles != Nsites:
                raise ValueError("Specify spin conservation with an integer is ambiguous. "
                                 "Please use a tuple (Nup, Ndown).")
        else:
            Nparticles = tuple(int(n) for n in Nparticles)
            if sum(Nparticles) != Nsites:
                raise ValueError("The total number of spin-up and spin-down particles should be equal to the "
                                 "number of sites in spin systems.")
        self._Nparticles = Nparticles
        if double_occ:
            raise ValueError("Double occupancy is only for spinful fermions.")
        self._double_occ = False
        self._coord = None if coord is None else np.asarray(coord, dtype=float)
        self._dist = None
        self._neighbors = []

    Nsites = property(lambda self: self._Nsites)
    Nmodes = property(lambda self: self._Nsites)
    Nparticles = property(lambda self: self._Nparticles)
    particle_type = property(lambda self: self._particle_type)
    double_occ = property(lambda self: self._double_occ)
    is_fermion = property(lambda self: False)
    is_spinful = property(lambda self: True)

    @property
    def Ntotal(self):
        return self._Nparticles if isinstance(self._Nparticles, int) else sum(self._Nparticles)

    @property
    def coord(self) -> np.ndarray:
        if self._coord is None:
            raise RuntimeError("The coordinates are unavailable.")
        return self._coord

    @property
    def ndim(self) -> int:
        return self.coord.shape[1]

    @property
    def dist(self) -> np.ndarray:
        if self._dist is None:
            self._dist = self._compute_dist()
        return self._dist

    def _compute_dist(self) -> np.ndarray:
        return np.linalg.norm(self.coord[None, :, :] - self.coord[:, None, :], axis=2)

    def get_neighbor(self, n_neighbor: Union[int, Sequence[int]] = 1):
        """n-th nearest neighbour pairs (i<j), one [nbonds, 2] array per requested shell."""
        shells = [n_neighbor] if isinstance(n_neighbor, int) else list(n_neighbor)
        self._ensure_shells(max(shells))
        out = [self._neighbors[n - 1] for n in shells]
        return out[0] if isinstance(n_neighbor, int) else out

    def _ensure_shells(self, nmax: int) -> None:
        tol = 1e-6  # quantax/sites/sites.py:269
        d = self.dist
        while len(self._neighbors) < nmax:
            if self._neighbors:
                i, j = self._neighbors[-1][0]
                lower = d[i, j] * (1 + tol)
            else:
                lower = tol
            r = d[d > lower].min()
            ii, jj = np.nonzero(np.abs((d - r) / r) < tol)  # row-major = lexicographic (i, j)
            keep = ii < jj
            self._neighbors.append(np.stack([ii[keep], jj[keep]], axis=1))


class Lattice(Sites):
    """Periodic structure with one site per unit cell (quantax/sites/lattice.py:12-90)."""

    def __init__(self, extent, basis_vectors, site_offsets=None, boundary=1, particle_type=PARTICLE_TYPE.spin,
                 Nparticles=None, double_occ=None):
        if site_offsets is not None and np.asarray(site_offsets).shape[0] != 1:
            raise NotImplementedError("multi-site unit cells are outside the B200 hot path")
        extent = tuple(int(e) for e in extent)
        nd = len(extent)
        self._basis_vectors = np.asarray(basis_vectors, dtype=float)
        self._shape = (1,) + extent
        self._boundary = np.full(nd, boundary, dtype=int) if isinstance(boundary, int) else np.asarray(boundary, int)
        if np.any(self._boundary == -1):
            raise ValueError("Spin system can't have anti-periodic boundary conditions.")
        n = int(np.prod(extent))
        cell = np.stack(np.unravel_index(np.arange(n), extent), axis=1)  # row-major site index
        self._xyz_from_index = np.concatenate([np.zeros((n, 1), dtype=int), cell], axis=1)
        self._index_from_xyz = np.arange(n).reshape(self._shape)
        super().__init__(n, particle_type, Nparticles, double_occ, cell.astype(float) @ self._basis_vectors)

    shape = property(lambda self: self._shape)
    boundary = property(lambda self: self._boundary)
    basis_vectors = property(lambda self: self._basis_vectors)
    index_from_xyz = property(lambda self: self._index_from_xyz)
    xyz_from_index = property(lambda self: self._xyz_from_index)

    @property
    def ncells(self):
        return int(np.prod(self._shape[1:]))

    def _compute_dist(self) -> np.ndarray:
        """Shortest distance where each periodic axis may wrap once (quantax/sites/lattice.py:139-179)."""
        cell = self._xyz_from_index[:, 1:]
        delta = cell[None, :, :] - cell[:, None, :]
        ext = np.asarray(self._shape[1:])
        best = np.full(delta.shape[:2], np.inf)
        choices = [(0, 1) if bc != 0 else (0,) for bc in self._boundary]
        for wrap in product(*choices):
            d = delta.copy()
            for ax, w in enumerate(wrap):
                if w:
                    d[..., ax] = d[..., ax] - np.sign(d[..., ax]) * ext[ax]
            best = np.minimum(best, np.linalg.norm(d.astype(float) @ self._basis_vectors, axis=-1))
        return best


class Grid(Lattice):
    def __init__(self, extent, boundary=1, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None):
        super().__init__(extent, np.eye(len(extent)), None, boundary, particle_type, Nparticles, double_occ)


def Chain(L, boundary=1, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None):
    return Grid([L], boundary, particle_type, Nparticles, double_occ)


def Square(L, boundary=1, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None):
    return Grid([L, L], boundary, particle_type, Nparticles, double_occ)


def Cube(L, boundary=1, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None):
    return Grid([L, L, L], boundary, particle_type, Nparticles, double_occ)


class Triangular(Lattice):
    """quantax/sites/common_lattices.py:99-115."""

    def __init__(self, extent, boundary=1, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None):
        if isinstance(extent, int):
            extent = [extent] * 2
        super().__init__(extent, np.array([[1, 0], [0.5, np.sqrt(0.75)]]), None, boundary, particle_type,
                         Nparticles, double_occ)


class TriangularB(Lattice):
    """2D triangular lattice of type B, N = 3 extent^2 sites (quantax/sites/common_lattices.py:118-139)."""

    def __init__(self, extent: int, boundary=1, particle_type=PARTICLE_TYPE.spin, Nparticles=None, double_occ=None):
        super().__init__([extent * 3, extent], np.array([[1, 0], [1.5, np.sqrt(0.75)]]), None, boundary, particle_type,
                         Nparticles, double_occ)

