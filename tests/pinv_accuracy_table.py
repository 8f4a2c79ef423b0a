#!/usr/bin/env python
"""CPU measurement behind DESIGN.md 4.0b: error of the MinSR step x = A^T f(A A^T) b against a 50-digit evaluation of
the same float64 T = A A^T, for (a) the eigenvalue route (LAPACK syevd, the reference's formula), (b) a second LAPACK
driver (syevr) -- how reproducible the eigenvalue route is --, (c) the three shifted LU solves in plain float64,
(d) the same with refinement in double-double (the product's arithmetic, oracle/pinv_rational.py).
Writes profiles/r1_cpu_pinv_rational_accuracy.md.  Needs mpmath; runs in about a minute.  (Lives under tests/ because it
uses the oracle, which only test code may import.)"""
import os
import sys

import mpmath as mp
import numpy as np
import scipy.linalg as sla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pinv_rational as pr, solver as osolver  # noqa: E402


def problem(ns, npar, decay, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
    V, _ = np.linalg.qr(rng.standard_normal((npar, ns)))
    A = (U * np.exp(-decay * np.arange(ns) / ns)) @ V.T
    A -= A.mean(axis=0, keepdims=True)
    return A / np.sqrt(ns), rng.standard_normal(ns) / np.sqrt(ns)


def exact(T, b, rtol):
    mp.mp.dps = 50
    n = T.shape[0]
    E, Q = mp.eigsy(mp.matrix(T.tolist()))
    c = mp.mpf(rtol) * max(abs(e) for e in E)
    bb = mp.matrix(b.tolist())
    coef = [sum(Q[i, k] * bb[i] for i in range(n)) * (E[k] ** 5 / (E[k] ** 6 + c ** 6)) for k in range(n)]
    return np.array([float(sum(Q[i, k] * coef[k] for k in range(n))) for i in range(n)])


def eig_route(T, b, rtol, driver):
    w, Q = sla.eigh(T, driver=driver)
    return Q @ (osolver.eigs_inv(w, rtol) * (Q.T @ b))


def band_route(T, b, rtol, lam, nb=8):
    """Candidate for a later round: orthogonal reduction of T to bandwidth nb (blocked Householder, what a BLAS-3 sy2sb
    does), then the three shifted solves on the band in 34-digit arithmetic (standing in for double-double)."""
    A = T.copy()
    n = A.shape[0]
    Qf = np.eye(n)
    for j0 in range(0, n - nb - 1, nb):
        r0 = j0 + nb
        Q, _ = np.linalg.qr(A[r0:, j0:j0 + nb], mode="complete")
        A[r0:, :] = Q.T @ A[r0:, :]
        A[:, r0:] = A[:, r0:] @ Q
        Qf[:, r0:] = Qf[:, r0:] @ Q
    B = np.zeros_like(A)
    for i in range(n):
        lo, hi = max(0, i - nb), min(n, i + nb + 1)
        B[i, lo:hi] = A[i, lo:hi]
    B = (B + B.T) / 2
    mp.mp.dps = 34
    c = mp.mpf(rtol) * mp.mpf(lam)
    M0, bp, y = mp.matrix(B.tolist()), mp.matrix((Qf.T @ b).tolist()), mp.matrix(n, 1)
    for k in range(3):
        z = c * mp.e ** (1j * mp.pi * (2 * k + 1) / 6)
        M = M0.copy()
        for i in range(n):
            M[i, i] -= z
        x = mp.lu_solve(M, bp)
        for i in range(n):
            y[i] += mp.re(x[i]) / 3
    return Qf @ np.array([float(v) for v in y])


def main():
    rows = []
    for ns, npar, decay, rtol in ((60, 240, 3, 1e-12), (60, 240, 20, 1e-12), (60, 240, 40, 1e-12), (60, 240, 20, 1e-8),
                                  (60, 240, 20, 1e-4)):
        A, b = problem(ns, npar, decay, 11)
        T = A @ A.T
        w = np.linalg.eigvalsh(T)
        lam = float(np.abs(w).max())
        xt = A.T @ exact(T, b, rtol)
        err = lambda y: float(np.linalg.norm(A.T @ y - xt) / np.linalg.norm(xt))
        corr = []
        y_ref = pr.dd_sum_scale([pr.pinv_rational_partial(T, b, rtol=rtol, lam=lam, corrections=corr)])
        rows.append((decay, rtol, int((w < rtol * lam).sum()), err(eig_route(T, b, rtol, "evd")),
                     err(eig_route(T, b, rtol, "evr")), err(pr.pinv_rational_solve(T, b, rtol=rtol, lam=lam, refine_steps=0)),
                     err(y_ref), err(band_route(T, b, rtol, lam)), corr[:4]))
    out = ["# Accuracy of the soft pseudo-inverse routes (CPU, `tests/pinv_accuracy_table.py`)", "",
           "Relative error of the MinSR step `x = Aᵀ f(AAᵀ) b` against a 50-digit evaluation of `f(T) b` for the same float64",
           "`T` (60 × 60, singular values of `A` = exp(−decay·i/60), centred columns).  `eigh` = the reference's formula on",
           "LAPACK `syevd`; `syevr` = a second LAPACK driver; `LU` = three shifted complex LU solves in float64; `LU + dd` = the",
           "same with 4 refinement steps on double-double residuals (the arithmetic of `csrc/pinv_rational.cu`); `band` = a",
           "candidate for a later round: orthogonal reduction to bandwidth 8, then the shifted solves on the band in 34-digit",
           "arithmetic -- cheaper ((4/3)n³ real BLAS-3 flops), but the reduction perturbs T by eps|T| like `eigh` does.", "",
           "| decay | rtol | eigenvalues below the cut-off | eigh (syevd) | eigh (syevr) | LU | LU + dd | band | size of the corrections |",
           "|---|---|---|---|---|---|---|---|---|"]
    for d, r, nb, e1, e2, e3, e4, e5, corr in rows:
        out.append(f"| {d} | {r:g} | {nb} | {e1:.1e} | {e2:.1e} | {e3:.1e} | {e4:.1e} | {e5:.1e} | "
                   + " → ".join(f"{c:.0e}" for c in corr) + " |")
    path = os.path.join(ROOT, "profiles", "r1_cpu_pinv_rational_accuracy.md")
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
