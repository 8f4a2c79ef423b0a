"""The KERNELS of quantax_b200/csrc/pinv_rational.cu executed on the CPU (tests/native/cuda_emu.h: one std::thread per
CUDA thread, std::barrier for __syncthreads / warp shuffles) in the order of qtx_sym_absmax_eig and
qtx_pinv_rational_partial, with SciPy's complex LU standing in for cuSOLVER's Zgetrf / Zgetrs.  Checks the kernel
source itself -- indexing, reductions, double-double bookkeeping -- against the NumPy restatement
(oracle/pinv_rational.py) and the eigenvalue route, without a GPU."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import pinv_rational as pr, solver as osolver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")

COS = [0.86602540378443864676, 0.0, -0.86602540378443864676]  # qtx_pinv_rational_partial kCos / kSin
SIN = [0.5, 1.0, 0.5]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libpinv_emu.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++20", "-pthread", "-shared", "-fPIC",
                    "-I", os.path.join(ROOT, "quantax_b200", "csrc"), "-I", os.path.join(ROOT, "tests", "native"),
                    "-I", os.path.join(ROOT, "include"), "-x", "c++",
                    os.path.join(ROOT, "tests", "native", "pinv_rational_emu.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    p, d, i64, i32, u32 = C.POINTER(C.c_double), C.c_double, C.c_int64, C.c_int, C.c_uint
    L.emu_sym_absmax_eig.argtypes = [p, i64, i32, i32, u32, p, p]
    L.emu_shift_build.argtypes = [p, i64, p, p, d, d, d, d, u32, p, p]
    L.emu_dd_set.argtypes = [i64, p, p]
    L.emu_dd_correct.argtypes = [i64, p, p]
    L.emu_dd_residual.argtypes = [p, i64, p, p, d, d, d, d, p, p]
    L.emu_dd_zero.argtypes = [i64, p]
    L.emu_dd_accum_real.argtypes = [i64, p, p]
    L.emu_dd_sum_scale.argtypes = [p, i32, i64, d, p]
    L.emu_lanczos_max_steps.restype = i32
    return L


def _p(a):
    assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.float64, np.complex128)
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _lanczos(emu, T, stages, block=64):
    n = T.shape[0]
    work = np.zeros(3 * n + 2 * emu.emu_lanczos_max_steps() + 2)
    lam = np.zeros(1)
    done, out = 0, []
    for upto in stages:
        emu.emu_sym_absmax_eig(_p(T), n, done, upto, block, _p(work), _p(lam))
        done = upto
        out.append(float(lam[0]))
    return out


def _partial(emu, T, b, rtol, atol, lam, mask, refine_steps, ydd, grid=3):
    """qtx_pinv_rational_partial, kernel for kernel (accumulate = 1 semantics on ydd)."""
    n = T.shape[0]
    lam_a = np.array([lam])
    for k in range(3):
        if not (mask >> k) & 1:
            continue
        M = np.zeros((n, n), dtype=np.complex128)
        rhs = np.zeros(n, dtype=np.complex128)
        x = np.zeros(4 * n)
        emu.emu_shift_build(_p(T), n, _p(b), _p(lam_a), rtol, atol, COS[k], SIN[k], grid, _p(M), _p(rhs))
        lu = sla.lu_factor(M)  # cusolverDnZgetrf (M is complex symmetric: row-major = column-major)
        for it in range(refine_steps + 1):
            if it > 0:
                emu.emu_dd_residual(_p(T), n, _p(b), _p(lam_a), rtol, atol, COS[k], SIN[k], _p(x), _p(rhs))
            rhs = np.ascontiguousarray(sla.lu_solve(lu, rhs))  # cusolverDnZgetrs
            (emu.emu_dd_set if it == 0 else emu.emu_dd_correct)(n, _p(rhs), _p(x))
        emu.emu_dd_accum_real(n, _p(x), _p(ydd))
    return ydd


def _solve(emu, T, b, rtol, atol=0.0, masks=(7,), refine_steps=4, lam=None):
    n = T.shape[0]
    if lam is None:
        lam = pr.abs_max_eigenvalue(T)
    parts = np.zeros((len(masks), 2, n))
    for q, mask in enumerate(masks):
        emu.emu_dd_zero(2 * n, _p(parts[q]))
        _partial(emu, T, b, rtol, atol, lam, mask, refine_steps, parts[q])
    y = np.zeros(n)
    emu.emu_dd_sum_scale(_p(parts), len(masks), n, 1.0 / 3.0, _p(y))
    return y


def _problem(ns, npar, decay, seed):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
    V, _ = np.linalg.qr(rng.standard_normal((npar, ns)))
    A = (U * np.exp(-decay * np.arange(ns) / ns)) @ V.T
    A -= A.mean(axis=0, keepdims=True)
    A /= np.sqrt(ns)
    return A, rng.standard_normal(ns) / np.sqrt(ns)


def _rel(x, ref):
    return float(np.linalg.norm(x - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("n,seed", [(1, 0), (3, 1), (40, 2), (150, 3)])
def test_lanczos_kernels_match_the_oracle_and_lapack(emu, n, seed):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, n))
    for T in (B @ B.T, B + B.T, np.zeros((n, n)))[: 3 if n < 100 else 1]:
        T = np.ascontiguousarray(T)
        ref = np.abs(np.linalg.eigvalsh(T)).max()
        stages = pr.lanczos_stages(n)
        got = _lanczos(emu, T, stages)  # staged continuation of the recurrence
        assert abs(got[-1] - ref) <= 1e-12 * max(ref, 1e-300)
        one_go = _lanczos(emu, T, [stages[-1]])[0]  # the same number of steps in one call
        assert one_go == got[-1]
        if n == 40:  # another block size changes the summation order only
            assert abs(_lanczos(emu, T, [stages[-1]], block=256)[0] - got[-1]) <= 1e-13 * max(ref, 1e-300)
        for k, g in zip(stages, got):
            assert abs(g - pr.abs_max_eigenvalue(T, steps=k)) <= 1e-10 * max(ref, 1e-300)


def test_lanczos_start_vector_is_the_oracles(emu):
    n = 100
    work = np.zeros(3 * n + 2 * emu.emu_lanczos_max_steps() + 2)
    lam = np.zeros(1)
    T = np.eye(n)
    emu.emu_sym_absmax_eig(_p(T), n, 0, 1, 1024, _p(work), _p(lam))  # the library's block size
    v0 = pr.start_vector(n)
    # after one step on the identity: alpha_0 = 1, breakdown, v_prev = normalised start vector
    assert np.allclose(work[2 * n:3 * n], v0 / np.linalg.norm(v0), rtol=1e-15, atol=0)
    assert abs(work[3 * n] - 1.0) < 1e-15 and work[3 * n + emu.emu_lanczos_max_steps()] == 0.0 and lam[0] == 1.0


@pytest.mark.parametrize("ns,npar,decay,rtol", [(48, 300, 3, 1e-12), (33, 333, 1, 1e-10), (40, 400, 6, 1e-8)])
def test_kernel_chain_equals_oracle_and_eigenvalue_route(emu, ns, npar, decay, rtol):
    A, b = _problem(ns, npar, decay, seed=ns)
    T = np.ascontiguousarray(A @ A.T)
    y = _solve(emu, T, b, rtol)
    y_or = pr.pinv_rational_solve(T, b, rtol=rtol)
    y_eig = osolver.minsr_pinv_eig(T, b, rtol=rtol)
    assert _rel(A.T @ y, A.T @ y_or) < 1e-12
    assert _rel(A.T @ y, A.T @ y_eig) < 1e-10


def test_kernel_chain_rank_split_and_grid_sizes(emu):
    A, b = _problem(20, 100, 4, seed=3)
    T = np.ascontiguousarray(A @ A.T)
    lam = pr.abs_max_eigenvalue(T)
    y1 = _solve(emu, T, b, 1e-12, lam=lam)
    y2 = _solve(emu, T, b, 1e-12, masks=(0b101, 0b010), lam=lam)
    y8 = _solve(emu, T, b, 1e-12, masks=(1, 2, 4, 0, 0, 0, 0, 0), lam=lam)
    assert _rel(A.T @ y2, A.T @ y1) < 1e-13 and _rel(A.T @ y8, A.T @ y1) < 1e-13
    # the build kernel with more blocks than rows / a single block gives the same matrix
    n = T.shape[0]
    out = []
    for grid in (1, 7, 64):
        M = np.zeros((n, n), dtype=np.complex128)
        rhs = np.zeros(n, dtype=np.complex128)
        emu.emu_shift_build(_p(T), n, _p(b), _p(np.array([lam])), 1e-12, 0.0, COS[0], SIN[0], grid, _p(M), _p(rhs))
        out.append((M.copy(), rhs.copy()))
    z = (1e-12 * lam) * (COS[0] + 1j * SIN[0])
    assert np.array_equal(out[0][0], T - z * np.eye(n)) and np.array_equal(out[0][1], b.astype(complex))
    assert all(np.array_equal(o[0], out[0][0]) and np.array_equal(o[1], out[0][1]) for o in out[1:])


def test_kernel_chain_refinement_removes_the_pollution(emu):
    """Cut-off inside the spectrum: without refinement the float64 LU is only good to eps / rtol."""
    A, b = _problem(40, 160, 20, seed=11)
    T = np.ascontiguousarray(A @ A.T)
    lam = float(np.abs(np.linalg.eigvalsh(T)).max())
    x_ref = A.T @ pr.pinv_rational_solve(T, b, lam=lam)  # held to a 50-digit evaluation in test_pinv_rational_cpu.py
    e4 = _rel(A.T @ _solve(emu, T, b, 1e-12, lam=lam), x_ref)
    e0 = _rel(A.T @ _solve(emu, T, b, 1e-12, lam=lam, refine_steps=0), x_ref)
    assert e4 < 1e-9 and e0 > 1e-7 and e0 > 100 * e4


def test_kernel_chain_degenerate_zero_matrix(emu):
    n = 9
    T = np.zeros((n, n))
    b = np.arange(1.0, n + 1)
    y = _solve(emu, T, b, 1e-12, lam=0.0)
    assert not y.any()


# ---- the ENTRY POINTS themselves (host orchestration, workspace layout, cuSOLVER call sequence) --------------------
def _entry(emu):
    vp, d, i64, i32, sz = C.c_void_p, C.c_double, C.c_int64, C.c_int, C.c_size_t
    emu.qtx_pinv_rational_workspace_size.restype = sz
    emu.qtx_pinv_rational_workspace_size.argtypes = [i64]
    emu.qtx_sym_absmax_eig.argtypes = [vp, i64, i32, i32, vp, vp, sz, vp]
    emu.qtx_pinv_rational_partial.argtypes = [vp, i64, vp, d, d, vp, i32, i32, vp, i32, vp, vp, sz, vp]
    emu.qtx_dd_sum_scale.argtypes = [vp, i32, i64, d, vp, vp]
    emu.qtx_pinv_ldlt_workspace_size.restype = sz
    emu.qtx_pinv_ldlt_workspace_size.argtypes = [i64, i32]
    emu.qtx_sym_absmax_eig_ws.argtypes = [vp, i64, i32, i32, vp, vp, sz, i32, vp]
    emu.qtx_pinv_ldlt_partial.argtypes = emu.qtx_pinv_rational_partial.argtypes
    emu.emu_last_error.restype = C.c_char_p
    return emu


def _vp(a):
    return a.ctypes.data


def _entry_solve(emu, T, b, rtol, atol=0.0, masks=(7,), refine=4, accumulate_split=False, route="rational"):
    """quantax_b200.optimizer.pinv_rational_solve through the C entry points of the emulated library; route "ldlt" =
    the library's own LDL^T kernels (csrc/zldlt.cu, the default of the product), "rational" = the cuSOLVER LU calls."""
    n = T.shape[0]
    ldlt = route == "ldlt"
    nsh = max(1, max(bin(m).count("1") for m in masks))
    wsz = emu.qtx_pinv_ldlt_workspace_size(n, nsh) if ldlt else emu.qtx_pinv_rational_workspace_size(n)
    assert wsz > 0
    ws = np.full(wsz + 8, 0x5A, dtype=np.uint8)  # guard bytes behind the promised size
    lam = np.zeros(1)
    done = 0
    for upto in pr.lanczos_stages(n):
        if ldlt:
            rc = emu.qtx_sym_absmax_eig_ws(_vp(T), n, done, upto, _vp(lam), _vp(ws), wsz, nsh, None)
        else:
            rc = emu.qtx_sym_absmax_eig(_vp(T), n, done, upto, _vp(lam), _vp(ws), wsz, None)
        assert rc == 0, emu.emu_last_error()
        done = upto
    partial = emu.qtx_pinv_ldlt_partial if ldlt else emu.qtx_pinv_rational_partial
    parts = np.zeros((len(masks), 2, n))
    info = np.full(1, 77, dtype=np.int32)
    for q, mask in enumerate(masks):
        if mask == 0:
            continue
        if accumulate_split:  # one call per shift, accumulating into the same double-double vector
            first = True
            for k in range(3):
                if (mask >> k) & 1:
                    rc = partial(_vp(T), n, _vp(b), rtol, atol, _vp(lam), 1 << k, refine,
                                 _vp(parts[q]), 0 if first else 1, _vp(info), _vp(ws), wsz, None)
                    assert rc == 0, emu.emu_last_error()
                    first = False
        else:
            parts[q] = 1e300  # accumulate = 0 must overwrite
            rc = partial(_vp(T), n, _vp(b), rtol, atol, _vp(lam), mask, refine, _vp(parts[q]), 0,
                         _vp(info), _vp(ws), wsz, None)
            assert rc == 0, emu.emu_last_error()
        assert info[0] == 0
    assert (ws[wsz:] == 0x5A).all(), "wrote behind the workspace"
    y = np.zeros(n)
    assert emu.qtx_dd_sum_scale(_vp(parts), len(masks), n, 1.0 / 3.0, _vp(y), None) == 0
    return y, float(lam[0])


@pytest.mark.parametrize("route", ["rational", "ldlt"])
@pytest.mark.parametrize("ns,npar,decay,rtol", [(40, 300, 3, -1.0), (33, 333, 1, 1e-10), (17, 100, 6, 1e-8)])
def test_entry_points_equal_oracle_and_eigenvalue_route(emu, ns, npar, decay, rtol, route):
    emu = _entry(emu)
    A, b = _problem(ns, npar, decay, seed=ns + 1)
    T = np.ascontiguousarray(A @ A.T)
    y, lam = _entry_solve(emu, T, b, rtol, route=route)
    r = None if rtol < 0 else rtol  # rtol < 0 selects the float64 default 1e-12
    assert abs(lam - np.abs(np.linalg.eigvalsh(T)).max()) <= 1e-12 * lam
    assert _rel(A.T @ y, A.T @ pr.pinv_rational_solve(T, b, rtol=r)) < 1e-11
    assert _rel(A.T @ y, A.T @ osolver.minsr_pinv_eig(T, b, rtol=r)) < (1e-10 if rtol > 0 else 1e-9)


def test_ldlt_route_rank_split_accumulate_zero_matrix_and_errors(emu):
    """qtx_pinv_ldlt_partial (own LDL^T kernels): rank split, accumulate, degenerate and refused inputs."""
    emu = _entry(emu)
    A, b = _problem(20, 100, 4, seed=3)
    T = np.ascontiguousarray(A @ A.T)
    y0, _ = _entry_solve(emu, T, b, 1e-12)  # the LU route
    y1, _ = _entry_solve(emu, T, b, 1e-12, route="ldlt")
    y2, _ = _entry_solve(emu, T, b, 1e-12, masks=(0b101, 0b010), route="ldlt")
    y8, _ = _entry_solve(emu, T, b, 1e-12, masks=(1, 2, 4, 0, 0, 0, 0, 0), route="ldlt")
    ya, _ = _entry_solve(emu, T, b, 1e-12, accumulate_split=True, route="ldlt")
    assert _rel(A.T @ y1, A.T @ y0) < 1e-12
    assert _rel(A.T @ y2, A.T @ y1) < 1e-13 and _rel(A.T @ y8, A.T @ y1) < 1e-13 and np.array_equal(ya, y1)
    n = T.shape[0]
    wsz = emu.qtx_pinv_ldlt_workspace_size(n, 3)
    assert emu.qtx_pinv_ldlt_workspace_size(n, 4) == 0 and emu.qtx_pinv_ldlt_workspace_size(n, 1) < wsz
    ws, lam, ydd, info = np.zeros(wsz, dtype=np.uint8), np.ones(1), np.zeros((2, n)), np.zeros(1, dtype=np.int32)
    args = lambda rtol, atol, mask, size: (_vp(T), n, _vp(b), rtol, atol, _vp(lam), mask, 4, _vp(ydd), 0, _vp(info),
                                           _vp(ws), size, None)
    assert emu.qtx_pinv_ldlt_partial(*args(0.0, 0.0, 7, wsz)) == -3 and b"plain inverse" in emu.emu_last_error()
    assert emu.qtx_pinv_ldlt_partial(*args(1e-12, 0.0, 7, emu.qtx_pinv_ldlt_workspace_size(n, 1))) == -1
    assert emu.qtx_pinv_ldlt_partial(*args(1e-12, 0.0, 8, wsz)) == -1
    Z = np.zeros((6, 6))
    yz, lz = _entry_solve(emu, Z, np.arange(1.0, 7.0), 1e-12, route="ldlt")
    assert lz == 0.0 and not yz.any()


def test_entry_points_rank_split_accumulate_and_errors(emu):
    emu = _entry(emu)
    A, b = _problem(20, 100, 4, seed=3)
    T = np.ascontiguousarray(A @ A.T)
    y1, _ = _entry_solve(emu, T, b, 1e-12)
    y2, _ = _entry_solve(emu, T, b, 1e-12, masks=(0b101, 0b010))
    y8, _ = _entry_solve(emu, T, b, 1e-12, masks=(1, 2, 4, 0, 0, 0, 0, 0))
    ya, _ = _entry_solve(emu, T, b, 1e-12, accumulate_split=True)
    assert _rel(A.T @ y2, A.T @ y1) < 1e-13 and _rel(A.T @ y8, A.T @ y1) < 1e-13 and np.array_equal(ya, y1)
    # error behaviour: plain inverse refused, small workspace refused, bad mask refused
    n = T.shape[0]
    wsz = emu.qtx_pinv_rational_workspace_size(n)
    ws, lam, ydd, info = np.zeros(wsz, dtype=np.uint8), np.ones(1), np.zeros((2, n)), np.zeros(1, dtype=np.int32)
    args = lambda rtol, atol, mask, size: (_vp(T), n, _vp(b), rtol, atol, _vp(lam), mask, 4, _vp(ydd), 0, _vp(info),
                                           _vp(ws), size, None)
    assert emu.qtx_pinv_rational_partial(*args(0.0, 0.0, 7, wsz)) == -3 and b"plain inverse" in emu.emu_last_error()
    assert emu.qtx_pinv_rational_partial(*args(1e-12, 0.0, 7, wsz - 1)) == -1 and b"workspace" in emu.emu_last_error()
    assert emu.qtx_pinv_rational_partial(*args(1e-12, 0.0, 8, wsz)) == -1
    assert emu.qtx_sym_absmax_eig(_vp(T), n, 5, 5, _vp(lam), _vp(ws), wsz, None) == -1
    # zero matrix: zero cut-off, y = 0, info = 0
    Z = np.zeros((6, 6))
    yz, lz = _entry_solve(emu, Z, np.arange(1.0, 7.0), 1e-12)
    assert lz == 0.0 and not yz.any()


def test_entry_points_against_the_references_own_solver_code(emu):
    """Kernel source + host orchestration of csrc/pinv_rational.cu (emulated) against the steps the reference's own
    `minnorm_pinv_eig` / `lstsq_pinv_eig` code produced (tests/golden/ref_hotpath.npz, rtol = 1e-10)."""
    emu = _entry(emu)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_hotpath.npz"))
    A, b = gold["solver/minnorm/A"], gold["solver/minnorm/b"]
    for route in ("rational", "ldlt"):
        A, b = gold["solver/minnorm/A"], gold["solver/minnorm/b"]
        y, _ = _entry_solve(emu, np.ascontiguousarray(A @ A.T), np.ascontiguousarray(b), 1e-10, route=route)
        assert _rel(A.T @ y, gold["solver/minnorm/auto_pinv_eig_snr0.0"]) < 1e-9
        A, b = gold["solver/lstsq/A"], gold["solver/lstsq/b"]
        x, _ = _entry_solve(emu, np.ascontiguousarray(A.T @ A), np.ascontiguousarray(A.T @ b), 1e-10, route=route)
        assert _rel(x, gold["solver/lstsq/auto_pinv_eig_snr0.0"]) < 1e-9
