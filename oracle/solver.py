"""Oracle: SR / MinSR linear solve (test infrastructure).

Restates
  quantax/optimizer/sr.py:74-88    (Obar = (O - mean O) * sqrt(rw/Ns); _Omean),
  quantax/optimizer/sr.py:180-195  (Ebar, energy, VarE),
  quantax/optimizer/solver.py:12-21,94-101 (soft pseudo-inverse of eigenvalues),
  quantax/optimizer/solver.py:128-149 (minnorm_pinv_eig, MinSR),
  quantax/optimizer/solver.py:152-164 (lstsq_pinv_eig, SR),
  quantax/optimizer/solver.py:167-201 (auto_pinv_eig),
  quantax/optimizer/solver.py:262-294 (minsr_pinv_eig),
  quantax/optimizer/solver.py:24-90 (lstsq_shift_cg, minnorm/lstsq/auto_shift_eig), :114-125 (_sum_without_noise),
  quantax/optimizer/solver.py:204-259 (block_pinv_eig), :297-302 (sgd_solver),
  quantax/state/variational.py:558-579 (update: theta <- theta - step, skip if non-finite).
``eigh`` is LAPACK syevd here and cuSOLVER/XLA in the reference (third party, unpinned):
eigenvectors are only defined up to sign / rotation inside degenerate subspaces, so parity
is asserted on T, on the spectrum and on the final step x, never on U itself.
"""
from __future__ import annotations

import numpy as np


def get_rtol(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return 1e-12
    if dtype == np.float32:
        return 1e-6
    raise ValueError(dtype)


def obar(Omat, rw):
    ns = Omat.shape[0]
    omean = np.mean(Omat * rw[:, None], axis=0)
    factor = np.sqrt(rw / ns)[:, None]
    return (Omat - np.mean(Omat, axis=0, keepdims=True)) * factor, omean


def ebar(Eloc, rw):
    ns = Eloc.shape[0]
    emean = np.mean(Eloc * rw)
    var = np.mean(np.abs(Eloc - emean) ** 2 * rw)
    eb = (Eloc - np.mean(Eloc)) * np.sqrt(rw / ns)
    return eb, float(np.real(emean)), float(np.real(var))


def eigs_inv(vals, rtol=None, atol=0.0):
    a = np.abs(vals)
    if rtol is None:
        rtol = get_rtol(a.dtype)
    with np.errstate(divide="ignore", over="ignore", invalid="ignore"):
        inv_factor = 1 + ((rtol * np.max(a) + atol) / a) ** 6
        inv = 1 / (vals * inv_factor)
    return np.where(a > 0.0, inv, 0.0)


def _sum_without_noise(inputs, tol_snr):
    x = np.sum(inputs, axis=0)
    if tol_snr > 1e-6:
        mean = x / inputs.shape[0]
        var = np.sqrt(np.mean(np.abs(inputs - mean[None, :]) ** 2, axis=0) / inputs.shape[0])
        snr = np.abs(mean) / var
        x = x / (1 + (tol_snr / snr) ** 6)
    return x


def minnorm_pinv_eig(A, b, rtol=None, atol=0.0, tol_snr=0.0, return_parts=False):
    T = A @ A.conj().T
    vals, U = np.linalg.eigh(T)
    inv = eigs_inv(vals, rtol, atol)
    rho = _sum_without_noise(U.conj() * b[:, None], tol_snr)
    x = A.conj().T @ (U @ (inv * rho))
    if return_parts:
        return x, T, vals
    return x


def lstsq_pinv_eig(A, b, rtol=None, atol=0.0, tol_snr=0.0):
    S = A.conj().T @ A
    vals, V = np.linalg.eigh(S)
    inv = eigs_inv(vals, rtol, atol)
    rho_sk = (A.conj() @ V.conj()) * b[:, None]
    rho = _sum_without_noise(rho_sk, tol_snr)
    return V @ (inv * rho)


def auto_pinv_eig(A, b, rtol=None, atol=0.0, tol_snr=0.0):
    if A.shape[0] < A.shape[1]:
        return minnorm_pinv_eig(A, b, rtol, atol, tol_snr)
    return lstsq_pinv_eig(A, b, rtol, atol, tol_snr)


def minsr_pinv_eig(T, b, rtol=None, atol=0.0, tol_snr=0.0):
    vals, U = np.linalg.eigh(T)
    inv = eigs_inv(vals, rtol, atol)
    rho = _sum_without_noise(U.conj() * b[:, None], tol_snr)
    return U @ (inv * rho)


def _shift(trace, rshift, ashift):
    rel = get_rtol(np.asarray(trace).dtype) if rshift is None else rshift
    return rel * trace + ashift


def minnorm_shift_eig(A, b, rshift=None, ashift=1e-4):
    """solver.py:50-62: x = A^+ (A A^+ + shift I)^-1 b, shift = rshift tr(T) + ashift (Cholesky solve in the
    reference; any exact solve of the SPD system is the same restatement)."""
    T = A @ A.conj().T
    T = T + _shift(np.trace(T).real, rshift, ashift) * np.identity(T.shape[0], T.dtype)
    return A.conj().T @ np.linalg.solve(T, b)


def lstsq_shift_eig(A, b, rshift=None, ashift=1e-4):
    """solver.py:65-77."""
    S = A.conj().T @ A
    F = A.conj().T @ b
    S = S + _shift(np.trace(S).real, rshift, ashift) * np.identity(S.shape[0], S.dtype)
    return np.linalg.solve(S, F)


def auto_shift_eig(A, b, rshift=None, ashift=1e-4):
    """solver.py:80-90."""
    if A.shape[0] < A.shape[1]:
        return minnorm_shift_eig(A, b, rshift, ashift)
    return lstsq_shift_eig(A, b, rshift, ashift)


def lstsq_shift_cg(A, b, diag_shift=0.01, rtol=1e-5, atol=0.0, maxiter=None, return_iterations=False):
    """solver.py:24-47 for real A.  ``jax.scipy.sparse.linalg.cg`` (third party, unpinned; restated from its
    published algorithm): x0 = 0, no preconditioner, loop while |r|^2 > max(tol^2 |b|^2, atol^2) and k < maxiter,
    maxiter defaults to 10 * size."""
    diag = np.einsum("sk,sk->k", A, A)

    def S_apply(x):
        return A.T @ (A @ x) + diag_shift * diag * x

    F = A.conj().T @ b
    if maxiter is None:
        maxiter = 10 * F.size
    atol2 = max(rtol ** 2 * float(F @ F), atol ** 2)
    x = np.zeros_like(F)
    r = F.copy()
    p = r.copy()
    gamma = float(r @ r)
    k = 0
    while gamma > atol2 and k < maxiter:
        Ap = S_apply(p)
        alpha = gamma / float(p @ Ap)
        x = x + alpha * p
        r = r - alpha * Ap
        gamma_new = float(r @ r)
        p = r + (gamma_new / gamma) * p
        gamma = gamma_new
        k += 1
    return (x, k) if return_iterations else x


def block_pinv_eig(Obar, Ebar, layer_sizes, rtol=None, atol=0.0, tol_snr=0.0):
    """solver.py:204-259: split the parameter axis at the layer boundaries, solve every block with
    auto_pinv_eig against Ebar / nlayers, concatenate."""
    sizes = [n for n in layer_sizes if n > 0]
    cuts = np.cumsum(sizes)[:-1]
    Eb = Ebar / len(sizes)
    return np.concatenate([auto_pinv_eig(Oi, Eb, rtol, atol, tol_snr) for Oi in np.split(Obar, cuts, axis=1)])


def sgd_solver(A, b):
    """solver.py:297-302."""
    return A.conj().T @ b / b.shape[0]


def sr_step(Omat, Eloc, rw, rtol=None, atol=0.0, real_to_complex=False, imag_time=True):
    """optimizer/sr.py:90-123.  ``real_to_complex`` (real parameters, complex output, sr.py:99-104): the real and
    imaginary parts of Obar and Ebar are stacked as 2 Ns rows; real-time evolution (``imag_time=False``) stacks
    [-Im Ebar; Re Ebar] instead (sr.py:103-104)."""
    eb, energy, var = ebar(Eloc, rw)
    ob, _ = obar(Omat, rw)
    if real_to_complex:
        ob = np.concatenate([ob.real, ob.imag], axis=0)
        eb = np.concatenate([eb.real, eb.imag]) if imag_time else np.concatenate([-eb.imag, eb.real])
    elif not imag_time:
        raise NotImplementedError("real output: Ebar *= 1j needs complex parameters (out of scope)")
    return auto_pinv_eig(ob, eb, rtol, atol), energy, var


def pinvh_solve(H, b, rtol=None, atol=0.0):
    """solver.py:104-111."""
    vals, U = np.linalg.eigh(H)
    return U @ (eigs_inv(vals, rtol, atol) * (U.conj().T @ b))


def time_evol_step(Omat, Eloc, max_parallel=None, rtol=None, atol=0.0):
    """TimeEvol.get_step (optimizer/time_evol.py:55-134) for real parameters / complex output
    (VS_TYPE.real_to_complex): S = Re(Obar^+ Obar), F = -Im(Obar^+ Ebar), step = pinvh(S) F.
    ``max_parallel``: the chunked accumulation of time_evol.py:76-115 (un-centred sums, corrected at the end)."""
    ns = Eloc.shape[0]
    if max_parallel is None or ns <= max_parallel:
        eb, energy, var = ebar(Eloc, np.ones(ns))
        ob, _ = obar(Omat, np.ones(ns))
        S, F = ob.conj().T @ ob, ob.conj().T @ eb
    else:
        Emean = np.mean(Eloc)
        energy, var = float(Emean.real), float(np.mean(np.abs(Eloc - Emean) ** 2))
        S = np.zeros((Omat.shape[1],) * 2, dtype=complex)
        F = np.zeros(Omat.shape[1], dtype=complex)
        Om = np.zeros(Omat.shape[1], dtype=complex)
        for lo in range(0, ns, max_parallel):
            O, e = Omat[lo:lo + max_parallel], Eloc[lo:lo + max_parallel]
            Om += O.sum(axis=0)
            S += O.conj().T @ O
            F += O.conj().T @ e
        S, F, Om = S / ns, F / ns, Om / ns
        S = S - np.outer(Om.conj(), Om)
        F = F - Om.conj() * Emean
    return pinvh_solve(S.real, -F.imag, rtol, atol), energy, var, S.real, -F.imag


def update_params(params, step):
    """variational.py:570-579."""
    if not np.all(np.isfinite(step)):
        return params
    return params + (-step.real).astype(params.dtype)


# ---- momentum variants (quantax/optimizer/sr.py:198-429), real parameters -------------------------
class SpringOracle:
    def __init__(self, nparams, mu=0.9):
        self.mu, self.last = mu, np.zeros(nparams)

    def solve(self, Obar, Ebar):
        Ebar = Ebar - self.mu * (Obar @ self.last)
        step = auto_pinv_eig(Obar, Ebar) + self.mu * self.last
        self.last = step
        return step


class MarchOracle:
    def __init__(self, nparams, mu=0.95, beta=0.995):
        self.mu, self.beta = mu, beta
        self.last, self.V, self.t = np.zeros(nparams), np.zeros(nparams), 0

    def solve(self, Obar, Ebar):
        self.t += 1
        Ebar = Ebar - self.mu * (Obar @ self.last)
        if np.allclose(self.V, 0):
            V = np.ones_like(self.V)
        else:
            V = (self.V / (1 - self.beta ** self.t)) ** 0.25 + 1e-8
        step = auto_pinv_eig(Obar / V[None, :], Ebar)
        step = step / V + self.mu * self.last
        self.V = self.beta * self.V + (1 - self.beta) * np.abs(step - self.last) ** 2
        self.last = step
        return step


class AdamSROracle:
    def __init__(self, nparams, mu=0.95, beta=0.995):
        self.mu, self.beta = mu, beta
        self.m, self.v, self.t = np.zeros(nparams), np.zeros(nparams), 0

    def solve(self, Obar, Ebar):
        self.t += 1
        g = auto_pinv_eig(Obar, Ebar)
        self.m = self.mu * self.m + (1 - self.mu) * g
        self.v = self.beta * self.v + (1 - self.beta) * np.abs(g) ** 2
        m = self.m / (1 - self.mu ** self.t)
        V = (self.v / (1 - self.beta ** self.t)) ** 0.25 + 1e-8
        step = auto_pinv_eig(Obar / V[None, :], Ebar - Obar @ m)
        return step / V + m
