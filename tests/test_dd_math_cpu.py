"""The double-double primitives and the Sturm bisection of csrc/pinv_rational.cu, compiled for the HOST from the same
header (quantax_b200/csrc/dd_math.cuh) and checked against exact rational arithmetic / LAPACK."""
import ctypes as C
import os
import shutil
import subprocess
from fractions import Fraction

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ddmath") / "libddmath.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "quantax_b200", "csrc"), "-x", "c++",
                    os.path.join(ROOT, "tests", "native", "dd_math_host.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    d, p = C.c_double, C.POINTER(C.c_double)
    L.h_two_sum.argtypes = [d, d, p]
    L.h_two_prod.argtypes = [d, d, p]
    L.h_dd_add.argtypes = [d, d, d, d, p]
    L.h_dd_add_d.argtypes = [d, d, d, p]
    L.h_dd_mul_d.argtypes = [d, d, d, p]
    L.h_dd_dot.argtypes = [p, p, p, C.c_int, p]
    L.h_sturm_count.argtypes = [p, p, C.c_int, d]
    L.h_sturm_count.restype = C.c_int
    L.h_tridiag_eigenvalue.argtypes = [p, p, C.c_int, C.c_int]
    L.h_tridiag_eigenvalue.restype = d
    return L


def _call(fn, *args):
    out = (C.c_double * 2)()
    fn(*args, out)
    return out[0], out[1]


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _dd_pairs(rng, n):
    hi = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 9, n)
    lo = hi * 2.0 ** -53 * rng.uniform(-1, 1, n)  # a normalised double-double: |lo| <= ulp(hi)/2
    return hi, lo


def test_two_sum_and_two_prod_are_error_free(lib):
    rng = np.random.default_rng(0)
    a = rng.standard_normal(400) * 10.0 ** rng.integers(-12, 13, 400)
    b = rng.standard_normal(400) * 10.0 ** rng.integers(-12, 13, 400)
    for x, y in zip(a, b):
        s, e = _call(lib.h_two_sum, x, y)
        assert Fraction(s) + Fraction(e) == Fraction(x) + Fraction(y) and s == x + y
        p, q = _call(lib.h_two_prod, x, y)
        assert Fraction(p) + Fraction(q) == Fraction(x) * Fraction(y) and p == x * y


def test_dd_add_and_mul_keep_about_106_bits(lib):
    rng = np.random.default_rng(1)
    ah, al = _dd_pairs(rng, 300)
    bh, bl = _dd_pairs(rng, 300)
    for i in range(300):
        exact = Fraction(ah[i]) + Fraction(al[i]) + Fraction(bh[i]) + Fraction(bl[i])
        h, l = _call(lib.h_dd_add, ah[i], al[i], bh[i], bl[i])
        scale = abs(Fraction(ah[i]) + Fraction(al[i])) + abs(Fraction(bh[i]) + Fraction(bl[i]))
        assert abs(Fraction(h) + Fraction(l) - exact) <= scale * Fraction(1, 2 ** 102)
        assert abs(l) <= abs(h) * 2.0 ** -52 or h == 0.0
        h, l = _call(lib.h_dd_add_d, ah[i], al[i], bh[i])
        exact = Fraction(ah[i]) + Fraction(al[i]) + Fraction(bh[i])
        assert abs(Fraction(h) + Fraction(l) - exact) <= (abs(Fraction(ah[i])) + abs(Fraction(bh[i]))) * Fraction(1, 2 ** 102)
        h, l = _call(lib.h_dd_mul_d, ah[i], al[i], bh[i])
        exact = (Fraction(ah[i]) + Fraction(al[i])) * Fraction(bh[i])
        assert abs(Fraction(h) + Fraction(l) - exact) <= abs(exact) * Fraction(1, 2 ** 100)


def test_residual_row_sum_survives_twelve_digits_of_cancellation(lib):
    """The situation of dd_residual_kernel: products of size 1e12 that cancel to O(1)."""
    rng = np.random.default_rng(2)
    n = 2000
    t = rng.standard_normal(n)
    xh, xl = _dd_pairs(rng, n)
    xh *= 1e12 / np.abs(xh).max()
    xl = xh * 2.0 ** -53 * rng.uniform(-1, 1, n)
    # make the exact sum tiny: last element fixed so that the sum nearly cancels
    part = sum(Fraction(t[j]) * (Fraction(xh[j]) + Fraction(xl[j])) for j in range(n - 1))
    xh[-1] = float(-part / Fraction(t[-1]))
    xl[-1] = float(-part / Fraction(t[-1]) - Fraction(xh[-1]))
    exact = sum(Fraction(t[j]) * (Fraction(xh[j]) + Fraction(xl[j])) for j in range(n))
    h, l = _call(lib.h_dd_dot, _ptr(t), _ptr(xh), _ptr(xl), n)
    mag = sum(abs(Fraction(t[j]) * Fraction(xh[j])) for j in range(n))
    assert abs(exact) < mag * Fraction(1, 10 ** 25)  # the test really cancels
    assert abs(Fraction(h) + Fraction(l) - exact) <= mag * Fraction(1, 2 ** 98)
    assert abs(float(np.dot(t, xh)) - float(exact)) > 1e3 * abs(float(Fraction(h) + Fraction(l) - exact)) or float(exact) == 0


@pytest.mark.parametrize("m,seed", [(1, 0), (2, 1), (17, 2), (128, 3)])
def test_sturm_bisection_matches_lapack_and_the_oracle(lib, m, seed):
    from oracle import pinv_rational as pr

    rng = np.random.default_rng(seed)
    alpha = np.ascontiguousarray(rng.standard_normal(m) * 3)
    beta = np.ascontiguousarray(np.abs(rng.standard_normal(max(m, 1))))
    if m > 4:
        beta[m // 2] = 0.0  # a breakdown of the Lanczos recurrence decouples the matrix
    Tm = np.diag(alpha) + np.diag(beta[: m - 1], 1) + np.diag(beta[: m - 1], -1)
    w = np.linalg.eigvalsh(Tm)
    scale = np.abs(w).max()
    for target in (1, m, (m + 1) // 2):
        ev = lib.h_tridiag_eigenvalue(_ptr(alpha), _ptr(beta), m, target)
        assert abs(ev - w[target - 1]) <= 4e-15 * scale
    lo, hi = pr.tridiagonal_extremes(alpha, beta)
    assert lib.h_tridiag_eigenvalue(_ptr(alpha), _ptr(beta), m, 1) == lo
    assert lib.h_tridiag_eigenvalue(_ptr(alpha), _ptr(beta), m, m) == hi
    x = float(w[m // 2]) + 1e-9
    assert lib.h_sturm_count(_ptr(alpha), _ptr(beta), m, x) == int((w < x).sum())
