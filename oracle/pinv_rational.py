"""Soft pseudo-inverse WITHOUT an eigendecomposition -- TEST INFRASTRUCTURE ONLY (restates quantax_b200/csrc/
pinv_rational.cu step by step; the reference formula it reproduces is quantax/optimizer/solver.py:94-111).

The reference applies ``f(lambda) = 1 / (lambda (1 + (c/|lambda|)^6))``, ``c = rtol max|lambda| + atol``, to the
eigenvalues of a Hermitian matrix.  ``f(lambda) = lambda^5 / (lambda^6 + c^6) = P'(lambda) / (6 P(lambda))`` with
``P = lambda^6 + c^6``, so by partial fractions over the six roots ``z_k = c exp(i pi (2k+1)/6)`` of P

    f(T) b = (1/6) sum_{k=0..5} (T - z_k)^-1 b = (1/3) Re sum_{k=0,1,2} (T - z_k I)^-1 b      (T, b real)

-- an exact identity, no approximation: three complex-symmetric linear solves (LU) replace ``eigh``.  Every
``T - z_k I`` is at distance ``Im z_k >= c/2`` from singular, i.e. its condition number is at most ``2 |T| / c``:
the same ``eps |T| / c`` sensitivity the eigenvalue route has for eigenvalues near the cut-off
(tests/test_pinv_rational_cpu.py measures both).  ``max|lambda|`` comes from a Lanczos recurrence.
"""
import numpy as np

LANCZOS_STEPS = 512  # quantax_b200/optimizer.py LANCZOS_STEPS: upper bound of the adaptive run
REFINE_STEPS = 4  # quantax_b200/optimizer.py REFINE_STEPS


def start_vector(n: int) -> np.ndarray:
    """Deterministic pseudo-random start vector (the kernel's integer hash), entries in (-1, 1), not normalised."""
    i = np.arange(n, dtype=np.uint64)
    x = (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(32)
    x = (x * np.uint64(0xD6E8FEB86659FD93)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(32)
    return (x >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0


def lanczos_tridiagonal(T: np.ndarray, steps: int = LANCZOS_STEPS):
    """(alpha [m], beta [m]) of the three-term recurrence, no reorthogonalisation (ghost copies of converged Ritz
    values do not move the extreme ones).  A breakdown (beta below 1e-13 of the running scale) zeroes the
    remaining vectors, which appends zeros to the spectrum of the tridiagonal matrix."""
    n = T.shape[0]
    m = min(steps, n)
    v = start_vector(n)
    nrm = np.sqrt(v @ v)
    v = v / nrm
    v_prev = np.zeros(n)
    alpha, beta = np.zeros(m), np.zeros(m)
    beta_prev, scale = 0.0, 0.0
    for j in range(m):
        w = T @ v
        a = w @ v
        w = w - a * v - beta_prev * v_prev
        a2 = w @ v  # second Gram-Schmidt pass against the current vector
        w = w - a2 * v
        a += a2
        bnew = np.sqrt(w @ w)
        scale = max(scale, abs(a), bnew)
        alpha[j] = a
        if not (bnew > 1e-13 * scale):
            bnew = 0.0
            v_prev, v = v, np.zeros(n)
        else:
            v_prev, v = v, w / bnew
        beta[j] = bnew
        beta_prev = bnew
    return alpha, beta


def _sturm_count(alpha, beta, x):
    """Number of eigenvalues of the tridiagonal matrix below x."""
    cnt, d = 0, 1.0
    for i in range(len(alpha)):
        off = beta[i - 1] ** 2 if i > 0 else 0.0
        d = (alpha[i] - x) - off / d
        if d == 0.0:
            d = 1e-300
        if d < 0.0:
            cnt += 1
    return cnt


def tridiagonal_extremes(alpha, beta):
    """(smallest, largest) eigenvalue by bisection on the Sturm count."""
    m = len(alpha)
    b = np.abs(beta[: m - 1]) if m > 1 else np.zeros(0)
    r = np.zeros(m)
    r[: m - 1] += b
    r[1:] += b
    lo0, hi0 = float((alpha - r).min()), float((alpha + r).max())
    out = []
    for target in (1, m):  # first x with count >= target
        lo, hi = lo0, hi0
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if mid <= lo or mid >= hi:
                break
            if _sturm_count(alpha, beta, mid) >= target:
                hi = mid
            else:
                lo = mid
        out.append(0.5 * (lo + hi))
    return out[0], out[1]


def lanczos_stages(n: int, max_steps: int = LANCZOS_STEPS):
    """quantax_b200.optimizer.lanczos_stages: 32, 64, 128, ... capped by n and max_steps."""
    out, k = [], 32
    cap = max(1, min(n, max_steps))
    while k < cap:
        out.append(k)
        k *= 2
    out.append(cap)
    return out


LANCZOS_AGREE = 1e-7  # quantax_b200.optimizer.LANCZOS_AGREE


def abs_max_eigenvalue(T: np.ndarray, steps=None) -> float:
    """max|lambda|.  ``steps=None``: the recurrence is evaluated after 32, 64, 128, ... steps until two consecutive
    values agree to 1e-7 -- the error of an extreme Ritz value after 2k steps is about the square of its error after k
    steps, so the later value is then good to 1e-14 (quantax_b200.optimizer.sym_absmax_eig); the recurrence is
    deterministic, so evaluating its first k steps equals running k steps."""
    n = T.shape[0]
    if steps is not None:
        alpha, beta = lanczos_tridiagonal(T, steps)
        lo, hi = tridiagonal_extremes(alpha, beta)
        return max(abs(lo), abs(hi))
    stages = lanczos_stages(n)
    alpha, beta = lanczos_tridiagonal(T, stages[-1])
    prev = None
    for k in stages:
        lo, hi = tridiagonal_extremes(alpha[:k], beta[:k])
        cur = max(abs(lo), abs(hi))
        if prev is not None and abs(cur - prev) <= LANCZOS_AGREE * abs(cur):
            break
        prev = cur
    return cur


def shifts(c: float):
    return [c * np.exp(1j * np.pi * (2 * k + 1) / 6) for k in range(3)]


# ---- double-double helpers (error-free transformations on NumPy arrays) ---------------------------------------
def _two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def _split(a):
    c = 134217729.0 * a
    hi = c - (c - a)
    return hi, a - hi


def _two_prod(a, b):
    p = a * b
    ah, al = _split(a)
    bh, bl = _split(b)
    return p, ((ah * bh - p) + ah * bl + al * bh) + al * bl


def dd_add(ah, al, bh, bl):
    s, e = _two_sum(ah, bh)
    t, f = _two_sum(al, bl)
    e = e + t
    s, e = _two_sum(s, e)
    e = e + f
    return _two_sum(s, e)


def _dd_mul_d(ah, al, b):
    p, e = _two_prod(ah, b)
    return _two_sum(p, e + al * b)


def _dd_matvec(T, xh, xl):
    """T (x_hi + x_lo) accumulated in double-double along the rows."""
    ph, pl = _two_prod(T, xh[None, :])
    pl = pl + T * xl[None, :]
    sh, sl = np.zeros(T.shape[0]), np.zeros(T.shape[0])
    for j in range(T.shape[1]):
        sh, sl = dd_add(sh, sl, ph[:, j], pl[:, j])
    return sh, sl


def shifted_solve_refined(T, b, z, refine_steps=REFINE_STEPS, corrections=None):
    """Re (T - z I)^-1 b as a double-double vector: complex LU in float64, then ``refine_steps`` corrections whose
    residual b - (T - z I) x is evaluated in double-double with x kept in double-double (csrc/pinv_rational.cu
    dd_residual_kernel / dd_correct_kernel)."""
    import scipy.linalg as sla

    n = T.shape[0]
    lu = sla.lu_factor(T.astype(np.complex128) - z * np.eye(n))
    d = sla.lu_solve(lu, b.astype(np.complex128))
    zero = np.zeros(n)
    xrh, xrl, xih, xil = d.real.copy(), zero.copy(), d.imag.copy(), zero.copy()
    zr, zi = float(np.real(z)), float(np.imag(z))
    for _ in range(refine_steps):
        trh, trl = _dd_matvec(T, xrh, xrl)
        tih, til = _dd_matvec(T, xih, xil)
        zxr = dd_add(*_dd_mul_d(xrh, xrl, zr), *_dd_mul_d(xih, xil, -zi))
        zxi = dd_add(*_dd_mul_d(xih, xil, zr), *_dd_mul_d(xrh, xrl, zi))
        rr = dd_add(*dd_add(-trh, -trl, b, zero), *zxr)
        ri = dd_add(-tih, -til, *zxi)
        d = sla.lu_solve(lu, (rr[0] + rr[1]) + 1j * (ri[0] + ri[1]))
        xrh, xrl = dd_add(xrh, xrl, d.real, zero)
        xih, xil = dd_add(xih, xil, d.imag, zero)
        if corrections is not None:
            corrections.append(float(np.linalg.norm(d) / np.linalg.norm(xrh + 1j * xih)))
    return xrh, xrl


def pinv_rational_partial(T, b, rtol=None, atol=0.0, lam=None, which=(0, 1, 2), refine_steps=REFINE_STEPS,
                          corrections=None):
    """Double-double sum over the shifts in ``which`` of Re (T - z_k I)^-1 b (qtx_pinv_rational_partial)."""
    T = np.asarray(T, dtype=np.float64)
    n = T.shape[0]
    if rtol is None:
        rtol = 1e-12
    if rtol == 0.0 and atol == 0.0:
        raise ValueError("rtol = atol = 0 is the plain inverse: use the eigenvalue route")
    if lam is None:
        lam = abs_max_eigenvalue(T)
    c = rtol * lam + atol
    yh, yl = np.zeros(n), np.zeros(n)
    if not c > 0.0:
        return yh, yl  # T = 0: every eigenvalue is zero and maps to zero
    for k in which:
        h, l = shifted_solve_refined(T, np.asarray(b, dtype=np.float64), shifts(c)[k], refine_steps, corrections)
        yh, yl = dd_add(yh, yl, h, l)
    return yh, yl


def dd_sum_scale(parts, scale=1.0 / 3.0):
    """scale * sum of double-double vectors, summed in order and rounded once (qtx_dd_sum_scale)."""
    yh, yl = np.zeros_like(parts[0][0]), np.zeros_like(parts[0][0])
    for h, l in parts:
        yh, yl = dd_add(yh, yl, h, l)
    return (yh + yl) * scale


def pinv_rational_solve(T, b, rtol=None, atol=0.0, lam=None, refine_steps=REFINE_STEPS, masks=((0, 1, 2),)):
    """y = f(T) b; ``masks`` lists the shifts of every rank of a replicated solve (quantax_b200.optimizer
    rational_shift_masks), whose partial sums are added in rank order."""
    if lam is None:
        lam = abs_max_eigenvalue(np.asarray(T, dtype=np.float64))
    return dd_sum_scale([pinv_rational_partial(T, b, rtol, atol, lam, which, refine_steps) for which in masks])
