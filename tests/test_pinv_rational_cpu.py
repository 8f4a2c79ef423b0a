"""The eigendecomposition-free soft pseudo-inverse (oracle/pinv_rational.py restates csrc/pinv_rational.cu):
partial fractions of f(lambda) = lambda^5 / (lambda^6 + c^6) over three complex shifts, LU solves refined with
double-double residuals.  Checked against the reference formula through eigh (oracle/solver.py, which is held to the
reference's own `_get_eigs_inv` / `minnorm_pinv_eig` outputs in tests/test_golden_hotpath_cpu.py) and against an
exact evaluation in 50-digit arithmetic."""
import numpy as np
import pytest

from oracle import pinv_rational as pr, solver as osolver


def _problem(ns, npar, decay, kind, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "svd":  # prescribed singular values exp(-decay i / ns): eigenvalues of T down to exp(-2 decay)
        U, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        V, _ = np.linalg.qr(rng.standard_normal((npar, ns)))
        A = (U * np.exp(-decay * np.arange(ns) / ns)) @ V.T
    else:
        A = rng.standard_normal((ns, npar)) * np.exp(-decay * rng.random((1, npar)))
    A -= A.mean(axis=0, keepdims=True)  # like Obar: T has the exact null vector (1, ..., 1)
    A /= np.sqrt(ns)
    return A, rng.standard_normal(ns) / np.sqrt(ns)


def _rel(x, ref):
    return float(np.linalg.norm(x - ref) / np.linalg.norm(ref))


def _exact(T, b, rtol):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    n = T.shape[0]
    E, Q = mp.eigsy(mp.matrix(T.tolist()))
    c = mp.mpf(rtol) * max(abs(e) for e in E)
    bb = mp.matrix(b.tolist())
    coef = [sum(Q[i, k] * bb[i] for i in range(n)) * (E[k] ** 5 / (E[k] ** 6 + c ** 6)) for k in range(n)]
    return np.array([float(sum(Q[i, k] * coef[k] for k in range(n))) for i in range(n)])


def test_partial_fraction_identity_on_scalars():
    c = 0.37
    lam = np.array([-5.0, -0.4, -1e-9, 0.0, 1e-7, 0.2, 0.37, 3.0, 1e6])
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(lam != 0, 1 / (lam * (1 + (c / np.abs(lam)) ** 6)), 0.0)  # solver.py:94-101
    g = sum((1 / (lam - z)).real for z in pr.shifts(c)) / 3
    assert np.allclose(g, f, rtol=1e-13, atol=1e-13)
    assert np.allclose(osolver.eigs_inv(lam, rtol=c / 1e6), f, rtol=1e-15, atol=0)


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (50, 3), (300, 4)])
def test_lanczos_gives_the_largest_eigenvalue_magnitude(n, seed):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, n))
    for T in (B @ B.T, -(B @ B.T), B + B.T, np.zeros((n, n)), np.eye(n) * 3.0):
        ref = np.abs(np.linalg.eigvalsh(T)).max()
        assert abs(pr.abs_max_eigenvalue(T) - ref) <= 1e-12 * max(ref, 1e-300)


@pytest.mark.parametrize("ns,npar,decay,kind,rtol", [(96, 700, 3, "col", None), (130, 333, 1, "col", 1e-10),
                                                     (64, 640, 6, "svd", 1e-8), (40, 400, 2, "svd", 1e-3)])
def test_equals_the_eigenvalue_route_where_that_one_is_accurate(ns, npar, decay, kind, rtol):
    A, b = _problem(ns, npar, decay, kind, seed=ns)
    T = A @ A.T
    y_ref = osolver.minsr_pinv_eig(T, b, rtol=rtol)
    y = pr.pinv_rational_solve(T, b, rtol=rtol)
    assert _rel(A.T @ y, A.T @ y_ref) < 1e-10  # the MinSR step x = A^T y (solver.py:146)
    # y itself carries an absolute error ~1e-20 cond(T - z I) |b_null| / |T| along EXACT null directions of T (here the
    # constant vector, and b is not centred): the 1/c-sized terms of the three shifts cancel there.  A^T removes it.
    assert _rel(y, y_ref) < 1e-5
    u = np.ones(ns) / np.sqrt(ns)
    assert _rel(y - u * (u @ y), y_ref - u * (u @ y_ref)) < 1e-10


def test_shift_split_over_ranks_gives_the_same_bits_in_any_grouping():
    A, b = _problem(48, 300, 4, "svd", seed=7)
    T = A @ A.T
    lam = pr.abs_max_eigenvalue(T)
    y1 = pr.pinv_rational_solve(T, b, lam=lam)
    y2 = pr.pinv_rational_solve(T, b, lam=lam, masks=((0, 2), (1,)))
    y8 = pr.pinv_rational_solve(T, b, lam=lam, masks=((0,), (1,), (2,), (), (), (), (), ()))
    assert _rel(y2, y1) < 1e-13 and _rel(y8, y1) < 1e-13


def test_more_accurate_than_eigh_when_eigenvalues_sit_at_the_cutoff():
    """Spectrum running through the default cut-off 1e-12 |T|: against the exact f(T) b of the SAME float64 matrix,
    the eigenvalue route is only good to ~eps |T| / c (the computed eigenvalues near c carry that relative error),
    the refined shifted solves keep ten digits; without refinement the shifted solves have the same weakness."""
    A, b = _problem(60, 240, 20, "svd", seed=11)
    T = A @ A.T
    xt = A.T @ _exact(T, b, 1e-12)
    lam = float(np.abs(np.linalg.eigvalsh(T)).max())
    e_eigh = _rel(A.T @ osolver.minsr_pinv_eig(T, b), xt)
    corrections = []
    e_refined = _rel(A.T @ pr.dd_sum_scale([pr.pinv_rational_partial(T, b, lam=lam, corrections=corrections)]), xt)
    e_plain = _rel(A.T @ pr.pinv_rational_solve(T, b, lam=lam, refine_steps=0), xt)
    assert e_refined < 1e-9
    assert e_eigh > 100 * e_refined and e_plain > 100 * e_refined
    # every refinement step gains the factor eps cond(T - z I) ~ 1e-16 * 2 / 1e-12
    assert corrections[0] < 1e-3 and corrections[1] < 1e-3 * corrections[0] and corrections[3] < 1e-17


def test_degenerate_inputs():
    b = np.arange(5.0)
    assert not pr.pinv_rational_solve(np.zeros((5, 5)), b).any()
    with pytest.raises(ValueError):
        pr.pinv_rational_solve(np.eye(5), b, rtol=0.0, atol=0.0)
    # atol alone sets the cut-off
    T = np.diag([4.0, 1.0, 1e-3, 0.0, 2.0])
    y = pr.pinv_rational_solve(T, b, rtol=0.0, atol=1e-2)
    assert np.allclose(y, osolver.eigs_inv(np.diag(T), rtol=0.0, atol=1e-2) * b, rtol=1e-13, atol=1e-13)


def test_adaptive_lanczos_on_a_dense_spectral_edge():
    """Wishart matrix: the largest eigenvalues are packed at the edge (relative gaps ~ n^-2/3), where 64 steps are not
    enough; the staged run continues until two consecutive stages agree."""
    rng = np.random.default_rng(5)
    B = rng.standard_normal((700, 1400))
    T = B @ B.T
    ref = np.linalg.eigvalsh(T).max()
    assert abs(pr.abs_max_eigenvalue(T, steps=32) - ref) > 1e-9 * ref  # too few steps: not converged
    assert abs(pr.abs_max_eigenvalue(T) - ref) <= 1e-12 * ref
    assert pr.lanczos_stages(700) == [32, 64, 128, 256, 512] and pr.lanczos_stages(50) == [32, 50]
    assert pr.lanczos_stages(32) == [32] and pr.lanczos_stages(5000, 100) == [32, 64, 100]
    from quantax_b200.optimizer import lanczos_stages

    for n, mx in ((700, 512), (50, 512), (64, 512), (5000, 100), (1, 512), (4096, 1024)):
        assert lanczos_stages(n, mx) == pr.lanczos_stages(n, mx)


GOLD = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "ref_hotpath.npz"))


@pytest.mark.parametrize("rtol,key,tol", [(1e-10, "auto_pinv_eig_snr0.0", 1e-9), (None, "auto_pinv_eig_default", 1e-6)])
def test_against_the_references_own_solver_code(rtol, key, tol):
    """The MinSR / SR steps the reference's OWN `minnorm_pinv_eig` / `lstsq_pinv_eig` produced
    (tests/golden/make_golden_hotpath.py) from the shifted solves: x = A^T f(A A^T) b and x = f(A^T A) A^T b."""
    A, b = GOLD["solver/minnorm/A"], GOLD["solver/minnorm/b"]
    x = A.T @ pr.pinv_rational_solve(A @ A.T, b, rtol=rtol)
    ref = GOLD[f"solver/minnorm/{key}"]
    assert np.linalg.norm(x - ref) <= tol * np.linalg.norm(ref)
    A, b = GOLD["solver/lstsq/A"], GOLD["solver/lstsq/b"]
    x = pr.pinv_rational_solve(A.T @ A, A.T @ b, rtol=rtol)
    ref = GOLD[f"solver/lstsq/{key}"]
    assert np.linalg.norm(x - ref) <= tol * np.linalg.norm(ref)
