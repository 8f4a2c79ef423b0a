"""Oracle: permutation symmetry groups, characters and the projected amplitude (test infrastructure).

Restates quantax/symmetry/symmetry.py:11-57 (_get_perm: group closure from generators, characters
from sectors), :325-341 (get_symm_spins), :344-392 (symmetrize), :394-432 (composition `@`),
quantax/symmetry/translation.py:10-55, quantax/symmetry/common_symmetries.py:90-278
(LinearTransform / Flip / Rotation / C4v / D6 / SpinInverse), and the projected log-derivative of
quantax/state/variational.py:438-491.  Real default dtype only (characters +-1)."""
from __future__ import annotations

import numpy as np


def get_perm(generator, sector):
    """symmetry.py:11-57 for spin systems (perm signs are trivially +1)."""
    generator = np.atleast_2d(generator)
    N = generator.shape[1]
    if np.array_equiv(generator, np.arange(N)):
        return np.arange(N)[None], np.array([1.0])
    s0 = np.arange(N)
    perm = s0.reshape(1, -1)
    character = np.array([1.0])
    for g, sec in zip(generator, sector):
        new_perm = [s0]
        s_perm = g
        while not np.array_equal(s0, s_perm):
            new_perm.append(s_perm)
            s_perm = s_perm[g]
        new_perm = np.stack(new_perm, axis=0)
        n = new_perm.shape[0]
        if not 0 <= sec < n:
            raise ValueError(f"Sector {sec} out of range.")
        if (sec * 2) % n != 0:
            raise ValueError("Default dtype is real, but got complex characters.")
        chi = -1.0 if sec else 1.0
        new_char = chi ** np.arange(n)
        perm = perm[:, new_perm].reshape(-1, N)
        character = np.einsum("i,j->ij", character, new_char).flatten()
    return perm, character


class Symmetry:
    def __init__(self, generator=None, sector=0, Z2_inversion=0, perm=None, character=None, N=None):
        if generator is None:
            generator = np.arange(N)[None]
        generator = np.atleast_2d(generator)
        self.generator = generator
        self.sector = [sector] * generator.shape[0] if np.isscalar(sector) else list(sector)
        self.Z2 = Z2_inversion
        if perm is None or character is None:
            p, c = get_perm(generator, self.sector)
            perm = p if perm is None else perm
            character = c if character is None else character
        self.perm, self.character = np.asarray(perm), np.asarray(character, dtype=np.float64)

    @property
    def nsymm(self):
        return self.character.size * (1 if self.Z2 == 0 else 2)

    def __matmul__(self, other):
        """symmetry.py:394-432."""
        N = self.perm.shape[1]
        perm = self.perm[:, other.perm].reshape(-1, N)
        character = np.einsum("i,j->ij", self.character, other.character).flatten()
        if self.Z2 == 0:
            z2 = other.Z2
        elif other.Z2 == 0 or other.Z2 == self.Z2:
            z2 = self.Z2
        else:
            raise ValueError("Symmetry with different Z2_inversion can't be added")
        return Symmetry(np.concatenate([self.generator, other.generator]), [*self.sector, *other.sector], z2, perm,
                        character)

    def get_symm_spins(self, s):
        """symmetry.py:325-341, batched: [ns, N] -> [ns, nsymm, N]."""
        out = s[:, self.perm]
        if self.Z2 != 0:
            out = np.concatenate([out, -out], axis=1)
        return out

    def weights(self):
        """symmetry.py:389-391: chi_g chi_0 / nsymm (with the Z2 block)."""
        c = self.character
        if self.Z2 != 0:
            c = np.concatenate([c, self.Z2 * c])
        return c * c[0] / c.size


def translation_generators(lattice, vectors):
    """translation.py:25-47 (periodic spin lattices)."""
    vectors = np.asarray(vectors, dtype=np.int64).reshape(-1, lattice.ndim)
    gens = []
    for vec in vectors:
        xyz = lattice.xyz.copy() + vec[None, :]
        xyz %= np.asarray(lattice.extent)
        idx = np.zeros(lattice.Nsites, dtype=np.int64)
        for ax in range(lattice.ndim):
            idx = idx * lattice.extent[ax] + xyz[:, ax]
        gens.append(idx)
    return np.stack(gens)


def Translation(lattice, vectors, sector=0):
    return Symmetry(translation_generators(lattice, vectors), sector)


def TransND(lattice, sector=0):
    return Symmetry(translation_generators(lattice, np.eye(lattice.ndim, dtype=np.int64)), sector)


def _standardize(lattice, coord):
    """common_symmetries.py:90-101."""
    basis = lattice.basis_vectors.T
    xyz = np.linalg.solve(basis, coord.T).T
    per = lattice.boundary != 0
    sh = xyz[:, per]
    ext = np.asarray(lattice.extent)[per]
    sh %= ext
    sh[np.isclose(sh, ext)] = 0.0
    xyz[:, per] = sh
    return np.einsum("ij,nj->ni", lattice.basis_vectors, xyz)


def LinearTransform(lattice, matrix, center=None, sector=0, character=None):
    """common_symmetries.py:104-145."""
    if center is None:
        center = np.mean(lattice.coord, axis=0)
    coord = _standardize(lattice, lattice.coord)
    new = _standardize(lattice, np.einsum("ij,nj->ni", matrix, lattice.coord - center) + center)
    match = np.all(np.isclose(coord[:, None, :], new[None, :, :]), axis=-1)
    if not np.all(match.sum(axis=1) == 1):
        raise ValueError("The transformation does not map the lattice to itself.")
    return Symmetry(np.argmax(match, axis=1), sector, character=character)


def Flip(lattice, axis=0, center=None, sector=0):
    m = np.ones(lattice.ndim)
    m[np.asarray(axis)] = -1
    return LinearTransform(lattice, np.diag(m), center, sector)


def Rotation(lattice, angle, axes=(0, 1), center=None, sector=0):
    m = np.eye(lattice.ndim)
    x, y = axes
    m[x, x] = m[y, y] = np.cos(angle)
    m[x, y] = -np.sin(angle)
    m[y, x] = np.sin(angle)
    return LinearTransform(lattice, m, center, sector)


def SpinInverse(lattice, eigval=1):
    return Symmetry(Z2_inversion=eigval, N=lattice.Nsites)


def project(symm, forward, s):
    """variational.py:262-266 + symmetry.py:386-392: psi_proj = sum_g w_g psi(T_g s) in the container
    ((mult, expo) with a signed log-sum-exp); returns (mult, expo) in the LogArray convention
    (sign, logabs) so that both containers can be compared through log|psi| and sign."""
    imgs = symm.get_symm_spins(np.asarray(s))
    ns, g, N = imgs.shape
    m, e = forward(imgs.reshape(-1, N))
    m, e = m.reshape(ns, g), e.reshape(ns, g)
    w = symm.weights()
    emax = e.max(axis=1, keepdims=True)
    b = np.sum(m * w[None, :] * np.exp(e - emax), axis=1)
    return np.sign(b), emax[:, 0] + np.log(np.abs(b)), (m, e, w, b, emax[:, 0])


def projected_jacobian(symm, forward, jacobian, s):
    """variational.py:438-491: O_proj(s) = sum_g (w_g psi_g / psi_proj) O(T_g s)."""
    imgs = symm.get_symm_spins(np.asarray(s))
    ns, g, N = imgs.shape
    _, _, (m, e, w, b, emax) = project(symm, forward, s)
    coef = m * w[None, :] * np.exp(e - emax[:, None]) / b[:, None]
    J = jacobian(imgs.reshape(-1, N)).reshape(ns, g, -1)
    return np.einsum("sg,sgk->sk", coef, J)
