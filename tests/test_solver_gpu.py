"""GPU parity tests of the SR / MinSR dense algebra against the CPU oracle (float64: 1e-10)."""
import numpy as np
import pytest
import torch

from oracle import models as omodels, operator as oop, sampler as osmp, sites as osites, solver as osolver
from tests.gpu_util import check, lattice_pair, make_rbm, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    return q


def _call(name, *args):
    from quantax_b200 import _lib

    _lib.call(name, *args)


def test_dense_helpers(qtx):
    from quantax_b200 import _lib
    from quantax_b200.optimizer import matvec, matvec_t

    rng = np.random.default_rng(0)
    for dt, tol in ((torch.float64, 1e-13), (torch.float32, 1e-6)):
        for ns, npar in ((37, 101), (64, 1000), (5, 3)):
            A = rng.standard_normal((ns, npar))
            At = torch.from_numpy(A).to("cuda", dt)
            A = to_np(At).astype(np.float64)
            w = rng.random(ns)
            mean = torch.empty(npar, dtype=torch.float64, device="cuda")
            _call("qtx_colmean", _lib.dtype_code(dt), _lib.ptr(At), ns, npar, npar, None, _lib.ptr(mean), _lib.stream())
            assert np.allclose(to_np(mean), A.mean(axis=0), atol=tol)
            wt = torch.from_numpy(w).cuda()
            _call("qtx_colmean", _lib.dtype_code(dt), _lib.ptr(At), ns, npar, npar, _lib.ptr(wt), _lib.ptr(mean), _lib.stream())
            assert np.allclose(to_np(mean), (A * w[:, None]).mean(axis=0), atol=tol)
            y = rng.standard_normal(ns); x = rng.standard_normal(npar)
            assert np.allclose(to_np(matvec_t(At, torch.from_numpy(y).cuda())), A.T @ y, atol=tol * npar)
            assert np.allclose(to_np(matvec(At, torch.from_numpy(x).cuda())), A @ x, atol=tol * npar)
            B = At.clone()
            sc = torch.from_numpy(rng.random(ns)).cuda()
            _call("qtx_center_scale", _lib.dtype_code(dt), _lib.ptr(B), ns, npar, npar, _lib.ptr(mean), _lib.ptr(sc), _lib.stream())
            ref = (A - to_np(mean)[None]) * to_np(sc)[:, None]
            assert np.allclose(to_np(B), ref, atol=tol * 10)


def test_ebar_energy_variance(qtx):
    from quantax_b200 import _lib

    rng = np.random.default_rng(1)
    for ns in (1, 64, 4097):
        E = rng.standard_normal(ns) * 3 - 20
        rw = rng.random(ns) + 0.5
        rw /= rw.mean()
        eb_o, e_o, v_o = osolver.ebar(E, rw)
        Et, rwt = torch.from_numpy(E).cuda(), torch.from_numpy(rw).cuda()
        eb = torch.empty(ns, dtype=torch.float64, device="cuda")
        st = torch.empty(2, dtype=torch.float64, device="cuda")
        _call("qtx_ebar", _lib.ptr(Et), _lib.ptr(rwt), ns, _lib.ptr(eb), _lib.ptr(st), _lib.stream())
        assert np.allclose(to_np(eb), eb_o, rtol=1e-12, atol=1e-13)
        assert np.allclose(to_np(st), [e_o, v_o], rtol=1e-12)


@pytest.mark.parametrize("nslices", [-1, 0])
def test_gram_fma(qtx, nslices):
    from quantax_b200.optimizer import gram

    rng = np.random.default_rng(2)
    for ns, npar in ((64, 256), (100, 333), (130, 50), (1, 7)):
        A = rng.standard_normal((ns, npar)) * np.exp(rng.standard_normal((ns, 1)))
        T = to_np(gram(torch.from_numpy(A).cuda(), nslices=nslices))
        ref = A @ A.T
        assert np.abs(T - ref).max() <= 1e-12 * np.abs(ref).max()
        assert np.array_equal(T, T.T)


def test_pinv_eig_solve(qtx):
    from quantax_b200.optimizer import pinv_eig_solve

    rng = np.random.default_rng(3)
    for n, rank in ((50, 50), (128, 100), (300, 299)):
        B = rng.standard_normal((n, rank))
        T = B @ B.T
        b = rng.standard_normal(n)
        y, evals, info = pinv_eig_solve(torch.from_numpy(T).cuda(), torch.from_numpy(b).cuda(), None, 0.0, want_evals=True)
        assert int(info.item()) == 0
        yo = osolver.minsr_pinv_eig(T, b)
        w = np.linalg.eigvalsh(T)
        assert np.allclose(to_np(evals), w, atol=1e-10 * w.max())
        assert np.linalg.norm(to_np(y) - yo) <= 1e-7 * np.linalg.norm(yo)
    # explicit tolerances (solver.py:94-101)
    T = np.diag([4.0, 1.0, 1e-3, 0.0])
    b = np.ones(4)
    y, _ = pinv_eig_solve(torch.from_numpy(T).cuda(), torch.from_numpy(b).cuda(), 1e-2, 1e-3)
    assert np.allclose(to_np(y), osolver.minsr_pinv_eig(T, b, rtol=1e-2, atol=1e-3), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("case", ["minsr", "sr"])
def test_sr_step_matches_oracle(qtx, case):
    """optimizer.get_step: MinSR when Ns < Np (solver.py:196), SR otherwise."""
    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    M = 24 if case == "minsr" else 2
    ns = 128
    model, net = make_rbm(qtx, 16, M, torch.float64, seed=5)
    state = qtx.state.Variational(model)
    H = qtx.operator.Heisenberg(msr=True)
    s = osmp.rand_states(ns, 16, 8, seed=6)
    samples = qtx.sampler.Samples(torch.from_numpy(s).cuda(), state(torch.from_numpy(s)), None,
                                  torch.ones(ns, dtype=torch.float64, device="cuda"))
    opt = qtx.optimizer.SR(state, H)
    step = to_np(opt.get_step(samples))
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat, msr=True))
    Eo = oop.oloc(aol, net.forward, s)
    xo, eo, vo = osolver.sr_step(net.jacobian(s), Eo, np.ones(ns))
    assert (ns < model.nparams) == (case == "minsr")
    assert abs(opt.energy - eo) <= 1e-10 * abs(eo) and abs(opt.VarE - vo) <= 1e-9 * abs(vo)
    # the spectrum of this system has a gap at the default cut-off (smallest non-zero eigenvalue 1.6e-3 lambda_max,
    # exact null directions at 1e-17): the float64 bar of 1e-10 applies
    check(f"{case} step vs oracle", np.linalg.norm(step - xo) / np.linalg.norm(xo), 1e-10)
    # MinSR alias forces the Ns x Ns solver
    if case == "minsr":
        step2 = to_np(qtx.optimizer.MinSR(state, H).get_step(samples))
        check("MinSR alias step vs oracle", np.linalg.norm(step2 - xo) / np.linalg.norm(xo), 1e-10)
    p0 = to_np(model.params).copy()
    state.update(torch.from_numpy(step).cuda() * 0.01)
    assert np.allclose(to_np(model.params), osolver.update_params(p0, step * 0.01), rtol=1e-14)
    assert state.check_last_update()
    bad = torch.from_numpy(step).cuda().clone()
    bad[3] = float("nan")
    p1 = to_np(model.params).copy()
    state.update(bad)
    with pytest.warns(UserWarning):
        assert not state.check_last_update()
    assert np.array_equal(to_np(model.params), p1)  # variational.py:570-573: update skipped


def test_quick_start_converges_to_ed(qtx):
    """README quick start (config A): Chain(8) Ising h=1, RBM_Dense(16), LocalFlip nsamples=64... run with
    more samples for a tight check against ED (E0 = -10.2516617910, oracle.ed_lowest)."""
    qtx.set_random_seed(42)
    lat, olat = lattice_pair(qtx, "chain", 8)
    H = qtx.operator.Ising(h=1.0)
    model = qtx.model.RBM_Dense(features=16)
    state = qtx.state.Variational(model)
    sampler = qtx.sampler.LocalFlip(state, nsamples=1024)
    optimizer = qtx.optimizer.SR(state, H)
    hist = []
    for i in range(300):  # README learning rate 1e-2 (README.md:70); 3e-2 sits at the edge of stability of plain SR
        samples = sampler.sweep()
        step = optimizer.get_step(samples)
        state.update(step * 1e-2)
        hist.append(optimizer.energy)
    e0 = oop.ed_lowest(oop.to_array_op_list(oop.ising_op_list(olat, h=1.0)), 8, k=1)[0]
    e = np.mean(hist[-20:])
    assert e > e0 - 0.02 and abs(e - e0) < 0.01 * abs(e0), (e, e0)
    assert sampler.check_local_updates(samples) == 0


@pytest.mark.parametrize("name", ["SPRING", "MARCH", "AdamSR"])
def test_momentum_optimizers_match_oracle(qtx, name):
    """SPRING / MARCH / AdamSR (quantax/optimizer/sr.py:198-429): three consecutive steps on fixed samples
    must reproduce the oracle's momentum recursion."""
    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    ns = 96
    model, net = make_rbm(qtx, 16, 20, torch.float64, seed=41)
    state = qtx.state.Variational(model)
    H = qtx.operator.Heisenberg(msr=True)
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat, msr=True))
    opt = getattr(qtx.optimizer, name)(state, H)
    orc = {"SPRING": osolver.SpringOracle, "MARCH": osolver.MarchOracle, "AdamSR": osolver.AdamSROracle}[name](model.nparams)
    for it in range(3):
        s = osmp.rand_states(ns, 16, 8, seed=50 + it)
        st = torch.from_numpy(s).cuda()
        samples = qtx.sampler.Samples(st, state(st), None, torch.ones(ns, dtype=torch.float64, device="cuda"))
        step = to_np(opt.get_step(samples))
        Eo = oop.oloc(aol, net.forward, s)
        eb, _, _ = osolver.ebar(Eo, np.ones(ns))
        ob, _ = osolver.obar(net.jacobian(s), np.ones(ns))
        xo = orc.solve(ob, eb)
        assert np.linalg.norm(step - xo) <= 1e-5 * np.linalg.norm(xo), (name, it)


def test_momentum_optimizers_save_and_resume(qtx, tmp_path):
    """SPRING / MARCH / AdamSR write their internal quantities in the reference's leaf order (sr.py:256-262, 343-349,
    423-429: one np.save blob per leaf of (mu, last_step) / (mu, beta, last_step, V, t) / (mu, beta, m, v, t)) and a
    resumed optimizer continues exactly like the one that kept running."""
    from quantax_b200.utils import read_eqx_leaves

    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, net = make_rbm(qtx, 16, 6, torch.float64, seed=9)
    state = qtx.state.Variational(model)
    H = qtx.operator.Heisenberg(msr=True)
    s = osmp.rand_states(96, 16, 8, seed=10)
    st = torch.from_numpy(s).cuda()
    samples = qtx.sampler.Samples(st, state(st), None, torch.ones(96, dtype=torch.float64, device="cuda"))
    for cls, nleaves, kw in ((qtx.optimizer.SPRING, 2, {"mu": 0.8}), (qtx.optimizer.MARCH, 5, {"mu": 0.9, "beta": 0.99}),
                             (qtx.optimizer.AdamSR, 5, {"mu": 0.9, "beta": 0.99})):
        a = cls(state, H, **kw)
        for _ in range(2):
            a.get_step(samples)
        f = str(tmp_path / f"{cls.__name__}.eqx")  # no suffix games: the path is used as given
        a.save(f)
        leaves = read_eqx_leaves(f)
        assert len(leaves) == nleaves and leaves[0].shape == () and float(leaves[0]) == kw["mu"]
        assert leaves[-1].shape == ((model.nparams,) if nleaves == 2 else ())
        b = cls(state, H, file=f)  # mu / beta come from the file
        xa, xb = a.get_step(samples), b.get_step(samples)
        # same state, same inputs; A^T y sums with atomics, so equal to rounding rather than bit for bit
        assert (xa - xb).norm().item() <= 1e-12 * xa.norm().item(), cls.__name__
