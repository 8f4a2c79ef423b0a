"""GPU parity tests of the remaining solver factories of quantax/optimizer/solver.py against the CPU oracle:
signal-to-noise damping (tol_snr), diagonal-shift Cholesky solvers, conjugate gradients, the layer-wise block
solver and the plain gradient.  (Named zz so that it runs after the hot-path suites.)"""
import numpy as np
import pytest
import torch

from oracle import solver as osolver
from tests.gpu_util import check,  to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    return q


def _problem(ns, npar, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((ns, npar)) * np.exp(0.5 * rng.standard_normal((1, npar)))
    A -= A.mean(axis=0, keepdims=True)
    return A / np.sqrt(ns), rng.standard_normal(ns) / np.sqrt(ns)


def _rel(x, ref):
    return float(np.linalg.norm(x - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("ns,npar", [(96, 700), (130, 333), (200, 64)])
@pytest.mark.parametrize("tol_snr", [0.5, 2.0])
def test_snr_damped_solvers(qtx, ns, npar, tol_snr):
    A, b = _problem(ns, npar, seed=ns)
    At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    x = to_np(qtx.optimizer.auto_pinv_eig(rtol=1e-10, tol_snr=tol_snr)(At.clone(), bt))
    ref = osolver.auto_pinv_eig(A, b, rtol=1e-10, tol_snr=tol_snr)
    assert _rel(x, ref) < 1e-8
    if ns < npar:
        T = A @ A.T
        y = to_np(qtx.optimizer.minsr_pinv_eig(rtol=1e-10, tol_snr=tol_snr)(torch.from_numpy(T).cuda(), bt))
        assert _rel(y, osolver.minsr_pinv_eig(T, b, rtol=1e-10, tol_snr=tol_snr)) < 1e-8


def test_rows_dot_snr_kernel(qtx):
    from quantax_b200.optimizer import rows_dot_snr

    rng = np.random.default_rng(5)
    M = rng.standard_normal((37, 501))
    b = rng.standard_normal(501)
    pad = torch.zeros((37, 512), dtype=torch.float64, device="cuda")
    pad[:, :501] = torch.from_numpy(M).cuda()
    for tol in (0.0, 1.0):
        rho = to_np(rows_dot_snr(pad[:, :501], torch.from_numpy(b).cuda(), tol))  # padded rows: ld = 512
        ref = osolver._sum_without_noise((M * b[None, :]).T, tol)
        assert np.allclose(rho, ref, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("ns,npar", [(96, 700), (257, 1000), (200, 64), (2, 5)])
def test_shift_cholesky_solvers(qtx, ns, npar):
    A, b = _problem(ns, npar, seed=ns + 1)
    At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    for rshift, ashift in ((None, 1e-4), (1e-3, 0.0)):
        solver = qtx.optimizer.auto_shift_eig(rshift, ashift)
        x = to_np(solver(At.clone(), bt))
        assert _rel(x, osolver.auto_shift_eig(A, b, rshift, ashift)) < 1e-9
    mn, ls = qtx.optimizer.minnorm_shift_eig(1e-3, 1e-4), qtx.optimizer.lstsq_shift_eig(1e-3, 1e-4)
    assert _rel(to_np(mn(At.clone(), bt)), to_np(ls(At.clone(), bt))) < 1e-8
    assert int(mn.last_info.item()) == 0


def test_shift_cholesky_reports_indefinite_matrix(qtx):
    from quantax_b200.optimizer import shift_chol_solve

    T = torch.tensor([[1.0, 2.0], [2.0, 1.0]], dtype=torch.float64, device="cuda")
    _, info = shift_chol_solve(T, torch.ones(2, dtype=torch.float64, device="cuda"), 0.0, 0.0)
    assert int(info.item()) > 0


def test_col_sumsq(qtx):
    from quantax_b200.optimizer import col_sumsq

    rng = np.random.default_rng(7)
    for dt, tol in ((torch.float64, 1e-13), (torch.float32, 1e-6)):
        for ns, npar in ((37, 101), (64, 1000), (5, 3)):
            At = torch.from_numpy(rng.standard_normal((ns, npar))).to("cuda", dt)
            ref = (to_np(At).astype(np.float64) ** 2).sum(axis=0)
            assert np.allclose(to_np(col_sumsq(At)), ref, rtol=tol * 10, atol=tol)


def test_conjugate_gradient_solver(qtx):
    A, b = _problem(300, 80, seed=11)
    At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    cg = qtx.optimizer.lstsq_shift_cg(diag_shift=0.01, rtol=1e-10)
    x = to_np(cg(At, bt))
    ref, k = osolver.lstsq_shift_cg(A, b, 0.01, 1e-10, return_iterations=True)
    assert _rel(x, ref) < 1e-7
    assert abs(cg.last_iterations - k) <= 2
    S = A.T @ A
    direct = np.linalg.solve(S + 0.01 * np.diag(np.diag(S)), A.T @ b)
    assert _rel(x, direct) < 1e-6
    capped = qtx.optimizer.lstsq_shift_cg(rtol=1e-12, maxiter=3)
    capped(At, bt)
    assert capped.last_iterations == 3


def test_block_solver_and_sgd(qtx):
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(4, Nparticles=(8, 8))
    model = qtx.model.ResConv(3, 4, 3, dtype=torch.float64)
    state = qtx.state.Variational(model)
    sizes = model.layer_param_sizes
    assert len(sizes) == 3 and sum(sizes) == model.nparams
    assert sizes[0] == 4 * 1 * 9 + 4 + 4 * 4 * 9 + 4 and sizes[2] == 2 * 4 * 4 * 9 + 4
    A, b = _problem(48, model.nparams, seed=13)
    At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    x = to_np(qtx.optimizer.block_pinv_eig(state, rtol=1e-10)(At, bt))
    assert _rel(x, osolver.block_pinv_eig(A, b, sizes, rtol=1e-10)) < 1e-8
    g = to_np(qtx.optimizer.sgd_solver()(At, bt))
    assert np.allclose(g, osolver.sgd_solver(A, b), rtol=1e-12, atol=1e-14)
    rbm = qtx.state.Variational(qtx.model.RBM_Dense(8))
    assert rbm.model.layer_param_sizes == [rbm.nparams]


def test_sr_with_shift_solver_runs_a_vmc_step(qtx):
    """The solver argument of SR is any callable (sr.py:32,106): a VMC step with auto_shift_eig against the oracle."""
    from oracle import models as omodels, operator as oop, sites as osites
    from tests.gpu_util import lattice_pair, make_rbm

    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, net = make_rbm(qtx, 16, 24, torch.float64, seed=3)
    state = qtx.state.Variational(model)
    H = qtx.operator.Heisenberg(msr=True)
    sampler = qtx.sampler.SpinExchange(state, nsamples=128)
    opt = qtx.optimizer.SR(state, H, solver=qtx.optimizer.auto_shift_eig(1e-3, 1e-4))
    samples = sampler.sweep()
    step = to_np(opt.get_step(samples))
    s = to_np(samples.spins)
    oH = oop.to_array_op_list(oop.heisenberg_op_list(olat, msr=True))
    Eo = oop.oloc(oH, net.forward, s)
    rw = np.ones(len(s))
    ob, _ = osolver.obar(net.jacobian(s), rw)
    eb, _, _ = osolver.ebar(Eo, rw)
    assert _rel(step, osolver.auto_shift_eig(ob, eb, 1e-3, 1e-4)) < 1e-8


# ---- API surface: dense apply_off_diag layout, Metropolis.propose, RandomSampler, digit-exact Gram ------------------

def test_apply_off_diag_dense_layout(qtx):
    from oracle import operator as oop, sampler as osmp
    from tests.gpu_util import lattice_pair

    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    aop = oop.to_array_op_list(oop.heisenberg_op_list(olat, J=[1, 0.5], n_neighbor=[1, 2], msr=True))
    s = osmp.rand_states(20, 16, 8, seed=5)
    got = H.apply_off_diag(torch.from_numpy(s).cuda())
    ref = oop.apply_off_diag(s, aop)
    assert sorted(got) == sorted(ref)
    for nflips, (sc, Hc) in got.items():
        sc, Hc = to_np(sc), to_np(Hc)
        rs, rH = ref[nflips]
        assert np.array_equal(np.isnan(Hc), np.isnan(rH))
        ok = ~np.isnan(rH)
        assert np.array_equal(Hc[ok], rH[ok]) and np.array_equal(sc[ok], rs[ok])


def test_propose_method_draws_from_the_sweep_stream(qtx):
    from oracle import sampler as osmp, sites as osites
    from tests.gpu_util import lattice_pair, make_rbm

    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, _ = make_rbm(qtx, 16, 8, torch.float64)
    state = qtx.state.Variational(model)
    sampler = qtx.sampler.SpinExchange(state, nsamples=64, thermal_steps=0)
    s = osmp.rand_states(64, 16, 8, seed=9)
    new = to_np(sampler.propose((1234, 7), torch.from_numpy(s)))
    table = osites.site_neighbor_table(olat)
    pos, slot, _ = osmp.philox_proposal("exchange", 1234, 7, np.arange(64), s, 1, table.shape[1])
    assert np.array_equal(new, osmp.propose_exchange(s, pos, slot, table))
    flip = qtx.sampler.LocalFlip(qtx.state.Variational(model), nsamples=64, thermal_steps=0)
    s1 = (2 * np.random.default_rng(3).integers(0, 2, (64, 16)) - 1).astype(np.int8)
    pos, _, _ = osmp.philox_proposal("localflip", 99, 0, np.arange(64), s1, 1, 0)
    assert np.array_equal(to_np(flip.propose(99, torch.from_numpy(s1))), osmp.propose_localflip(s1, pos))


def test_random_sampler_reweighting(qtx):
    from tests.gpu_util import lattice_pair, make_rbm

    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, net = make_rbm(qtx, 16, 8, torch.float64)
    state = qtx.state.Variational(model)
    smp = qtx.sampler.RandomSampler(state, 128).sweep()
    s = to_np(smp.spins)
    assert (s.sum(axis=1) == 0).all()  # Nparticles = (8, 8)
    sign, logabs = net.forward(s)
    w = np.exp(2 * logabs)
    assert np.allclose(to_np(smp.reweight_factor), w / w.mean(), rtol=1e-10)
    assert np.allclose(to_np(smp.psi.logabs), logabs, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("ns,npar,dtype,nslices", [(100, 333, np.float64, 0), (257, 1000, np.float64, 8),
                                                   (130, 4099, np.float64, 7), (64, 70001, np.float64, 8),
                                                   (200, 500, np.float32, 0), (300, 777, np.float64, 5)])
def test_gram_kernel_equals_its_digit_arithmetic_bit_for_bit(qtx, ns, npar, dtype, nslices):
    """oracle/gram_digits.py restates the digit split, the exact integer level sums and the order of the rounded
    additions of gram_split_kernel + gram_tc2_kernel: the tensor-core result must equal it exactly (including two
    K chunks at 70001 columns and 8 digits)."""
    from oracle import gram_digits as gd
    from quantax_b200.optimizer import gram

    rng = np.random.default_rng(ns + npar)
    A = (rng.standard_normal((ns, npar)) * np.exp(3 * rng.standard_normal((ns, 1)))).astype(dtype)
    A[:, ::7] *= 1e-3
    A[2] = 0.0
    T = to_np(gram(torch.from_numpy(A).cuda(), nslices=nslices))
    ref = gd.gram(A, nslices)
    assert np.array_equal(T, ref), float(np.abs(T - ref).max())
    T2 = to_np(gram(torch.from_numpy(A).cuda(), out=torch.from_numpy(ref.copy()).cuda(), nslices=nslices, accumulate=True))
    assert np.array_equal(T2, gd.gram(A, nslices, T=ref))


# ---- eigendecomposition-free pseudo-inverse (csrc/pinv_rational.cu, QTX_PINV=rational) -----------------------------
@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (7, 2), (130, 3), (1000, 4)])
def test_lanczos_absmax_eigenvalue(qtx, n, seed):
    from oracle import pinv_rational as pr
    from quantax_b200.optimizer import sym_absmax_eig

    rng = np.random.default_rng(seed)
    B = rng.standard_normal((n, n))
    for T in (B @ B.T, -(B @ B.T), B + B.T, np.zeros((n, n)), np.eye(n) * 3.0):
        ref = np.abs(np.linalg.eigvalsh(T)).max()
        lam = float(sym_absmax_eig(torch.from_numpy(np.ascontiguousarray(T)).cuda()).item())
        assert abs(lam - ref) <= 1e-12 * max(ref, 1e-300)
        assert abs(lam - pr.abs_max_eigenvalue(T)) <= 1e-12 * max(ref, 1e-300)


def _centred_problem(ns, npar, decay, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "svd":
        U, _ = np.linalg.qr(rng.standard_normal((ns, ns)))
        V, _ = np.linalg.qr(rng.standard_normal((npar, ns)))
        A = (U * np.exp(-decay * np.arange(ns) / ns)) @ V.T
    else:
        A = rng.standard_normal((ns, npar)) * np.exp(-decay * rng.random((1, npar)))
    A -= A.mean(axis=0, keepdims=True)
    return A / np.sqrt(ns), rng.standard_normal(ns) / np.sqrt(ns)


@pytest.mark.parametrize("method", ["ldlt", "rational"])  # own LDL^T kernels (default) / cuSOLVER LU (cross-check)
@pytest.mark.parametrize("ns,npar,decay,kind,rtol", [(96, 700, 3, "col", None), (130, 333, 1, "col", 1e-10),
                                                     (64, 640, 6, "svd", 1e-8), (300, 1200, 2, "svd", 1e-3),
                                                     (1000, 3000, 4, "col", 1e-9)])
def test_rational_pseudo_inverse_equals_the_eigenvalue_route(qtx, ns, npar, decay, kind, rtol, method):
    from oracle import pinv_rational as pr
    from quantax_b200.optimizer import pinv_eig_solve, pinv_rational_solve

    A, b = _centred_problem(ns, npar, decay, kind, ns)
    T = A @ A.T
    Tt, bt = torch.from_numpy(T).cuda(), torch.from_numpy(b).cuda()
    y, info = pinv_rational_solve(Tt, bt, rtol, 0.0, method=method)
    assert int(info.item()) == 0
    assert np.array_equal(to_np(Tt), T)  # T is not overwritten
    y = to_np(y)
    y_eig = to_np(pinv_eig_solve(Tt.clone(), bt, rtol, 0.0)[0])
    y_ref = osolver.minsr_pinv_eig(T, b, rtol=rtol)
    assert _rel(A.T @ y, A.T @ y_ref) < 1e-10 and _rel(A.T @ y, A.T @ y_eig) < 1e-10
    assert _rel(A.T @ y, A.T @ pr.pinv_rational_solve(T, b, rtol=rtol)) < 1e-10


def test_rational_pseudo_inverse_at_the_cutoff_beats_eigh(qtx):
    """Against the exact f(T) b (50 digits) of a spectrum running through the default cut-off."""
    from quantax_b200.optimizer import pinv_eig_solve, pinv_rational_solve
    from tests.test_pinv_rational_cpu import _exact

    A, b = _centred_problem(60, 240, 20, "svd", 11)
    T = A @ A.T
    xt = A.T @ _exact(T, b, 1e-12)
    Tt, bt = torch.from_numpy(T).cuda(), torch.from_numpy(b).cuda()
    e_eig = _rel(A.T @ to_np(pinv_eig_solve(Tt.clone(), bt, None, 0.0)[0]), xt)
    for method in ("ldlt", "rational"):
        e_rat = _rel(A.T @ to_np(pinv_rational_solve(Tt, bt, None, 0.0, method=method)[0]), xt)
        assert e_rat < 1e-8 and e_rat < e_eig, (method, e_rat, e_eig)


def test_minsr_step_through_the_rational_route(qtx, monkeypatch):
    import quantax_b200.optimizer as qopt

    A, b = _problem(200, 3000, seed=5)
    At, bt = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda()
    ref = osolver.auto_pinv_eig(A, b, rtol=1e-10)
    monkeypatch.setattr(qopt, "PINV_METHOD", "eigh")
    x_eig = to_np(qopt.auto_pinv_eig(rtol=1e-10)(At.clone(), bt))
    for method in ("ldlt", "rational"):
        monkeypatch.setattr(qopt, "PINV_METHOD", method)
        x_rat = to_np(qopt.auto_pinv_eig(rtol=1e-10)(At.clone(), bt))
        assert _rel(x_rat, ref) < 1e-9 and _rel(x_rat, x_eig) < 1e-9, method
    # SNR damping needs the eigen-directions: it stays on the eigenvalue route
    x_snr = to_np(qopt.auto_pinv_eig(rtol=1e-10, tol_snr=1.0)(At.clone(), bt))
    assert _rel(x_snr, osolver.auto_pinv_eig(A, b, rtol=1e-10, tol_snr=1.0)) < 1e-8


def test_ldlt_factorisation_and_solves_at_ragged_and_large_sizes(qtx):
    """csrc/zldlt.cu through qtx_pinv_ldlt_partial: sizes off the 64-row block, single shifts (the rank split), more
    block rows than SMs' worth of resident CTAs is not needed -- the ticket order makes the wavefront deadlock-free."""
    from quantax_b200 import _lib
    from quantax_b200.optimizer import _pinv_workspace, pinv_rational_solve, sym_absmax_eig

    for n, npar, seed in ((2, 4, 0), (63, 200, 1), (65, 300, 2), (129, 500, 3), (777, 2500, 4), (2048, 6000, 5)):
        A, b = _centred_problem(n, npar, 3, "col", seed)
        T = A @ A.T
        Tt, bt = torch.from_numpy(T).cuda(), torch.from_numpy(b).cuda()
        y, info = pinv_rational_solve(Tt, bt, 1e-9, 0.0, method="ldlt")
        assert int(info.item()) == 0
        ref = osolver.minsr_pinv_eig(T, b, rtol=1e-9)
        check(f"ldlt route n={n}", _rel(A.T @ to_np(y), A.T @ ref), 1e-10)
        # one shift per call, accumulated: what the ranks of a replicated solve do
        lam = sym_absmax_eig(Tt, method="ldlt", nshifts=1)
        ws, wsz = _pinv_workspace(n, "ldlt", 1)
        ydd = torch.zeros((2, n), dtype=torch.float64, device="cuda")
        inf = torch.zeros(1, dtype=torch.int32, device="cuda")
        for k in range(3):
            _lib.call("qtx_pinv_ldlt_partial", _lib.ptr(Tt), n, _lib.ptr(bt), 1e-9, 0.0, _lib.ptr(lam), 1 << k, 3,
                      _lib.ptr(ydd), int(k > 0), _lib.ptr(inf), _lib.ptr(ws), wsz, _lib.stream())
            assert int(inf.item()) == 0
        y1 = (ydd[0] + ydd[1]) / 3.0
        check(f"ldlt route, one shift per call n={n}", _rel(A.T @ to_np(y1), A.T @ to_np(y)), 1e-13)
