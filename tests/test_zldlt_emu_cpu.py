"""The kernels of quantax_b200/csrc/zldlt.cu (complex-symmetric LDL^T without pivoting, wavefront triangular solves)
executed on the CPU under tests/native/cuda_emu.h with 8 x 8 blocks: indexing, ragged last blocks, the ticket / flag
protocol in dependency order, against NumPy."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libzldlt_emu.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++20", "-pthread", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "quantax_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "native"),
                    "-I", os.path.join(ROOT, "include"), "-x", "c++",
                    os.path.join(ROOT, "tests", "native", "zldlt_emu.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    p = C.POINTER(C.c_double)
    L.emu_zldlt_scratch_bytes.restype = C.c_size_t
    L.emu_zldlt_scratch_bytes.argtypes = [C.c_int64]
    L.emu_zldlt_factor.argtypes = [p, C.c_int64, C.c_void_p, C.POINTER(C.c_int32)]
    L.emu_zldlt_solve.argtypes = [p, C.c_int64, p, C.c_void_p]
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _shifted(n, seed, c=1e-3, theta=np.pi / 6):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, max(1, n // 2)))  # rank deficient: eigenvalues 0 below the shift
    T = B @ B.T
    z = c * np.abs(np.linalg.eigvalsh(T)).max() * np.exp(1j * theta)
    return T, z, (T - z * np.eye(n)).astype(np.complex128)


def _ldlt_ref(M):
    n = M.shape[0]
    A = M.copy()
    L, d = np.eye(n, dtype=np.complex128), np.zeros(n, dtype=np.complex128)
    for k in range(n):
        d[k] = A[k, k]
        L[k + 1:, k] = A[k + 1:, k] / d[k]
        A[k + 1:, k + 1:] -= np.outer(L[k + 1:, k], A[k + 1:, k])
    return L, d


@pytest.mark.parametrize("n", [1, 5, 15, 16, 17, 33, 50, 71])
@pytest.mark.parametrize("theta", [np.pi / 6, np.pi / 2, 5 * np.pi / 6])
def test_factor_and_solve(emu, n, theta):
    assert emu.emu_zldlt_block_size() == 16
    T, z, M = _shifted(n, n, theta=theta)
    work = np.ascontiguousarray(np.tril(M) + np.triu(np.full((n, n), 7.5 + 3j), 1))  # the upper triangle is never read
    scratch = np.zeros(emu.emu_zldlt_scratch_bytes(n) + 64, dtype=np.uint8)
    guard = scratch[-64:]
    guard[:] = 0xAB
    info = np.zeros(1, dtype=np.int32)
    assert emu.emu_zldlt_factor(_p(work), n, scratch.ctypes.data, info.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert info[0] == 0 and (guard == 0xAB).all()
    L, d = _ldlt_ref(M)
    assert np.allclose(np.tril(work, -1), np.tril(L, -1), rtol=1e-9, atol=1e-12)
    assert np.allclose(np.diag(work), d, rtol=1e-9)
    assert np.array_equal(np.triu(work, 1), np.triu(np.full((n, n), 7.5 + 3j), 1))
    rng = np.random.default_rng(100 + n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x = b.copy()
    assert emu.emu_zldlt_solve(_p(work), n, _p(x), scratch.ctypes.data) == 0
    assert (guard == 0xAB).all()
    ref = np.linalg.solve(M, b)
    assert np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)
    # a second solve with the same factors (the refinement loop reuses them)
    x2 = (2 * b).copy()
    assert emu.emu_zldlt_solve(_p(work), n, _p(x2), scratch.ctypes.data) == 0
    assert np.linalg.norm(x2 - 2 * ref) <= 1e-9 * np.linalg.norm(ref)


def test_zero_pivot_is_reported(emu):
    n = 12
    M = np.zeros((n, n), dtype=np.complex128)
    M[np.arange(n), np.arange(n)] = 1.0
    M[3, 3] = 0.0
    scratch = np.zeros(emu.emu_zldlt_scratch_bytes(n), dtype=np.uint8)
    info = np.zeros(1, dtype=np.int32)
    work = M.copy()
    assert emu.emu_zldlt_factor(_p(work), n, scratch.ctypes.data, info.ctypes.data_as(C.POINTER(C.c_int32))) == 0
    assert info[0] == 4 and np.isfinite(work.view(np.float64)).all()
