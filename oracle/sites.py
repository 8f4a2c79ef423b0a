"""Oracle: lattice geometry and n-th neighbour bond lists (test infrastructure).

Restates quantax/sites/lattice.py:12-90 (site indexing + coordinates),
quantax/sites/lattice.py:139-179 (minimum-image distance under PBC) and
quantax/sites/sites.py:267-283 (neighbour shells: tolerance 1e-6, i<j,
``np.argwhere`` lexicographic order).  The bond ORDER matters: it defines the
connected-configuration enumeration order of Operator.Oloc.
"""
from __future__ import annotations

import itertools
import numpy as np


class Lattice:
    def __init__(self, extent, basis_vectors=None, boundary=1, Nparticles=None):
        extent = [int(e) for e in extent]
        ndim = len(extent)
        if basis_vectors is None:
            basis_vectors = np.eye(ndim)
        self.extent = tuple(extent)
        self.ndim = ndim
        self.basis_vectors = np.asarray(basis_vectors, dtype=float)
        if isinstance(boundary, int):
            boundary = [boundary] * ndim
        self.boundary = np.asarray(boundary, dtype=int)
        self.shape = (1,) + tuple(extent)
        self.Nsites = int(np.prod(extent))
        self.Nmodes = self.Nsites
        if Nparticles is None:
            Nparticles = self.Nsites  # lattice.py -> sites.py:67-69 (spin, unconserved)
        self.Nparticles = Nparticles
        # sites/lattice.py:73-79: row-major index over the extent
        idx = np.arange(self.Nsites)
        xyz = []
        for i in range(ndim):
            later = int(np.prod(extent[i + 1:], dtype=int))
            xyz.append(idx // later % extent[i])
        self.xyz = np.stack(xyz, axis=1)
        self.coord = self.xyz.astype(float) @ self.basis_vectors
        self._dist = None
        self._neighbors = []

    @property
    def dist(self):
        """sites/lattice.py:139-179: per axis with bc != 0 the displacement d may be
        replaced by d -+ L; the distance is the minimum over the 2^ndim choices."""
        if self._dist is None:
            d = self.xyz[None, :, :] - self.xyz[:, None, :]  # [i, j, axis], j - i
            alts = []
            for ax, L in enumerate(self.extent):
                if self.boundary[ax] != 0:
                    wrapped = np.where(d[..., ax] > 0, d[..., ax] - L,
                                       np.where(d[..., ax] < 0, d[..., ax] + L, d[..., ax]))
                    alts.append([d[..., ax], wrapped])
                else:
                    alts.append([d[..., ax]])
            best = None
            for combo in itertools.product(*alts):
                disp = np.stack(combo, axis=-1).astype(float) @ self.basis_vectors
                dist = np.linalg.norm(disp, axis=-1)
                best = dist if best is None else np.minimum(best, dist)
            self._dist = best
        return self._dist

    def get_neighbor(self, n_neighbor=1):
        """sites/sites.py:224-283."""
        many = not isinstance(n_neighbor, int)
        nmax = max(n_neighbor) if many else n_neighbor
        tol = 1e-6
        if len(self._neighbors) < nmax:
            self._neighbors = []
            min_dist = tol
            for _ in range(nmax):
                min_dist = np.min(self.dist[self.dist > min_dist])
                nb = np.argwhere(np.abs((self.dist - min_dist) / min_dist) < tol)
                nb = nb[nb[:, 0] < nb[:, 1]]
                self._neighbors.append(nb)
                min_dist *= 1 + tol
        if many:
            return [self._neighbors[n - 1] for n in n_neighbor]
        return self._neighbors[n_neighbor - 1]


def Chain(L, boundary=1, Nparticles=None):
    return Lattice([L], None, boundary, Nparticles)


def Square(L, boundary=1, Nparticles=None):
    return Lattice([L, L], None, boundary, Nparticles)


def Triangular(L, boundary=1, Nparticles=None):
    """sites/common_lattices.py:99-115."""
    ext = [L, L] if isinstance(L, int) else list(L)
    return Lattice(ext, np.array([[1, 0], [0.5, np.sqrt(0.75)]]), boundary, Nparticles)


def TriangularB(L, boundary=1, Nparticles=None):
    """common_lattices.py:118-139."""
    return Lattice([3 * L, L], np.array([[1, 0], [1.5, np.sqrt(0.75)]]), boundary, Nparticles)


def Cube(L, boundary=1, Nparticles=None):
    """common_lattices.py:51-57."""
    return Lattice([L, L, L], np.eye(3), boundary, Nparticles)


def site_neighbor_table(lattice, n_neighbor=1):
    """sampler/common_samplers.py:36-51: [N, max_nb] int32 table of neighbours of each
    site in ascending site order, padded with -1 (``flatnonzero(size=, fill_value=-1)``)."""
    nn = [n_neighbor] if isinstance(n_neighbor, int) else list(n_neighbor)
    nbs = np.concatenate(lattice.get_neighbor(nn), axis=0)
    N = lattice.Nsites
    mat = np.zeros((N, N), dtype=bool)
    mat[nbs[:, 0], nbs[:, 1]] = True
    mat |= mat.T
    max_nb = int(mat.sum(axis=1).max())
    table = -np.ones((N, max_nb), dtype=np.int32)
    for i in range(N):
        nz = np.flatnonzero(mat[i])
        table[i, : nz.size] = nz
    return table
