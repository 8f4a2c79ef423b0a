"""Complex-output states with real parameters (VS_TYPE.real_to_complex):
ResConv(out_dtype=complex128) + 120-degree Neel phase layer on the triangular lattice
(tutorials/triangular.ipynb:100-128,236-239), against the NumPy oracle.  float64 models: 1e-10; float32: 1e-5."""
import numpy as np
import pytest
import torch

from oracle import models as omodels, operator as oop, sampler as osmp, sites as osites, solver as osolver
from oracle import symmetry as osym
from tests.gpu_util import check, lattice_pair, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture()
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    q.set_default_dtype(torch.complex128)
    yield q
    q.set_default_dtype(torch.float64)


def neel120_kernel(Lx, Ly):
    """nn/sign.py:62-75."""
    x = 2 * np.arange(Lx)
    y = np.arange(Ly)
    k = (x[:, None] + y[None, :]) % 3
    return (np.pi / 3 * k - np.pi / 6).astype(np.float32).ravel()


def make_model(qtx, L, nb, C, dtype, final="exp", seed=0, phase=True):
    npdt = np.float32 if dtype == torch.float32 else np.float64
    net = omodels.ResConv.random((L, L), nb, C, 3, npdt, seed=seed, final=final, bias_std=0.1, out_complex=True,
                                 phase_kernel=neel120_kernel(L, L) if phase else None)
    fa = qtx.nn.exp_by_scale if final == "exp" else qtx.nn.sinhp1_by_scale
    model0 = qtx.model.ResConv(nb, C, 3, final_activation=fa, dtype=dtype, out_dtype=torch.complex128,
                               params=torch.from_numpy(net.params().copy()))
    if not phase:
        return model0, net

    class PhaseLayer(qtx.nn.RawInputLayer):  # verbatim from tutorials/triangular.ipynb cell 9
        def __call__(self, x, s):
            phase = qtx.nn.neel120_phase(s)
            return x * phase

    model = qtx.nn.Sequential(model0.layers + (PhaseLayer(),))
    return model, net


def logpsi(psi):
    m = to_np(psi.mult)
    return np.log(np.abs(m)) + to_np(psi.expo), np.angle(m)


def logpsi_np(sig, ex):
    return np.log(np.abs(sig)) + ex, np.angle(sig)


def angle_diff(a, b):
    return np.abs(np.angle(np.exp(1j * (a - b))))


def test_neel120_phase_matches_tutorial_formula(qtx):
    """tutorials/triangular.ipynb cell 5: the reference checks neel120_phase against this product formula."""
    lattice_pair(qtx, "triangular", 6, (18, 18))
    s = osmp.rand_states(9, 36, 18, seed=1)
    sub = ((np.arange(6) % 3)[:, None] + ((-np.arange(6)) % 3)[None, :]).ravel() % 3
    ph = np.exp(1j * sub * 2 * np.pi / 3)
    expect = np.prod(np.where(s > 0, 1.0, ph[None, :]), axis=1)
    got = to_np(qtx.nn.neel120_phase(torch.from_numpy(s)).tensor())
    assert np.abs(got - expect).max() < 1e-5  # float32 dot product inside (nn/sign.py:73)


@pytest.mark.parametrize("final", ["exp", "sinhp1"])
@pytest.mark.parametrize("dtype,tol,phase", [(torch.float64, 1e-10, False), (torch.float64, 2e-7, True),
                                             (torch.float32, 2e-5, True)])
def test_complex_forward_and_jacobian(qtx, final, dtype, tol, phase):
    # with the phase layer |psi| itself carries complex64 rounding (|exp(i phi)| = 1 +- 6e-8 in the reference's
    # float32 phase, nn/sign.py:36,73), so 1e-10 is only reachable without it
    lattice_pair(qtx, "triangular", 6, (18, 18))
    model, net = make_model(qtx, 6, 2, 8, dtype, final, seed=2, phase=phase)
    state = qtx.state.Variational(model, max_parallel=(64, 7))
    assert state.vs_type == qtx.state.VS_TYPE.real_to_complex
    s = osmp.rand_states(23, 36, 18, seed=3)
    la, ph = logpsi(state(torch.from_numpy(s)))
    lo, po = logpsi_np(*net.forward(s))
    assert np.abs(la - lo).max() <= tol * max(1.0, np.abs(lo).max())
    assert angle_diff(ph, po).max() <= max(tol, 2e-6)  # the phase layer is float32 in the reference too
    O = to_np(state.jacobian(torch.from_numpy(s)))
    Oo = net.jacobian(s)
    assert O.shape == Oo.shape and np.iscomplexobj(O)
    assert np.abs(O - Oo).max() <= 10 * tol * np.abs(Oo).max()


def test_complex_jacobian_on_the_tensor_cores(qtx):
    """Complex-output float32 ResConv on a 16x16 lattice: both backward passes (seeds d Re log psi, d Im log psi,
    variational.py:461-487) run through the tensor-core towers (csrc/resconv_tc.cu); rows against the oracle."""
    from quantax_b200 import _lib

    lattice_pair(qtx, "square", 16)
    model, net = make_model(qtx, 16, 2, 24, torch.float32, "sinhp1", seed=4, phase=False)
    state = qtx.state.Variational(model)
    assert _lib.lib().qtx_resconv_tc_backward_available(_lib.dtype_code(torch.float32), 24, 16, 16, 3, 3)
    s = osmp.rand_states(9, 256, seed=5)
    O = to_np(state.jacobian(torch.from_numpy(s)))
    Oo = net.jacobian(s)
    assert O.shape == Oo.shape and np.iscomplexobj(O)
    check("complex tc jacobian rows vs oracle", (np.linalg.norm(O - Oo, axis=1) / np.linalg.norm(Oo, axis=1)).max(), 1e-5)


@pytest.mark.parametrize("phase,tol", [(True, 1e-6), (False, 1e-10)])
def test_complex_oloc_sweep_and_sr_step(qtx, phase, tol):
    """Heisenberg on the 6x6 triangular lattice: exchange sweep with injected randoms (bit-exact accept pattern),
    complex local energies, stacked [Re; Im] SR step (sr.py:99-104).  With the phase layer the reference's own
    phases are complex64 (nn/sign.py:36,73), which bounds the agreement at 1e-6; without it the float64 bar holds."""
    lat, olat = lattice_pair(qtx, "triangular", 6, (18, 18))
    model, net = make_model(qtx, 6, 2, 4, torch.float64, "exp", seed=4, phase=phase)
    state = qtx.state.Variational(model)
    ns, T = 40, 25
    sampler = qtx.sampler.SpinExchange(state, ns, thermal_steps=0)
    spins0 = to_np(sampler._spins).copy()
    rng = np.random.default_rng(5)
    table = osites.site_neighbor_table(olat)
    u = rng.random((T, ns)); pos = rng.integers(0, 36, size=(T, ns)); slot = rng.integers(0, table.shape[1], size=(T, ns))
    sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), torch.from_numpy(slot))
    samples = sampler.sweep(T, record=True)
    ref = osmp.sweep(osmp.FullForwardChainModel(net), spins0, T, "exchange", neighbors=table, pos=pos, slot=slot, u=u,
                     record=True)
    assert np.array_equal(to_np(sampler.last_accept_log), ref["accept_log"])
    assert np.array_equal(to_np(samples.spins), ref["spins"])
    s = ref["spins"]
    H = qtx.operator.Heisenberg()
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat))
    Eo = oop.oloc(aol, net.forward, s)
    assert np.iscomplexobj(Eo)
    opt = qtx.optimizer.SR(state, H)
    step = to_np(opt.get_step(samples))
    check(f"complex Oloc phase={phase}", np.abs(to_np(opt._Eloc) - Eo).max() / np.abs(Eo).max(), tol)
    xo, eo, vo = osolver.sr_step(net.jacobian(s), Eo, np.ones(ns), real_to_complex=True)
    assert abs(opt.energy - eo.real) <= tol * abs(eo) and abs(opt.VarE - vo) <= 10 * tol * abs(vo)
    assert np.isrealobj(step)
    check(f"complex stacked SR step phase={phase}", np.linalg.norm(step - xo.real) / np.linalg.norm(xo), 10 * tol)
    # real-time evolution of the same state: the solver sees [-Im Ebar; Re Ebar] (sr.py:103-104)
    step_rt = to_np(qtx.optimizer.SR(state, H, imag_time=False).get_step(samples))
    xo_rt, _, _ = osolver.sr_step(net.jacobian(s), Eo, np.ones(ns), real_to_complex=True, imag_time=False)
    check(f"complex stacked SR step, imag_time=False, phase={phase}", np.linalg.norm(step_rt - xo_rt.real) / np.linalg.norm(xo_rt), 10 * tol)
    assert np.linalg.norm(step_rt - step) > 0.1 * np.linalg.norm(step)
    p0 = to_np(state.get_params_flatten()).copy()
    state.update(torch.from_numpy(step).cuda() * 0.01)
    assert np.allclose(to_np(state.get_params_flatten()), p0 - 0.01 * step, rtol=1e-12, atol=1e-14)


def test_complex_projected_state(qtx):
    """tutorials/triangular.ipynb cell 17: D6(center=(0, 0)) @ SpinInverse() projection of the complex state."""
    lat, olat = lattice_pair(qtx, "triangular", 6, (18, 18))
    model, net = make_model(qtx, 6, 2, 4, torch.float64, "exp", seed=6)
    S = qtx.symmetry
    symm = S.D6(center=(0, 0)) @ S.SpinInverse()
    osymm = osym.Rotation(olat, np.pi / 3, center=(0, 0)) @ osym.Flip(olat, center=(0, 0)) @ osym.SpinInverse(olat)
    state = qtx.state.Variational(model, symm=symm, max_parallel=(4096, 48))
    assert state.symm.nsymm == osymm.nsymm == 24
    s = osmp.rand_states(7, 36, 18, seed=7)
    _, _, (m, e, w, b, emax) = osym.project(osymm, net.forward, s)
    psi = state(torch.from_numpy(s))
    # the projection sums 24 images whose phases carry float32 rounding (in the reference too): compare the
    # projected amplitude on the scale of the summands, sum_g |w_g psi_g|
    got = to_np(psi.mult) * np.exp(to_np(psi.expo) - emax)
    scale = np.sum(np.abs(m * w[None, :]) * np.exp(e - emax[:, None]), axis=1)
    assert np.abs(got - b).max() <= 1e-6 * scale.max(), (np.abs(got - b).max(), scale.max(), np.abs(b).min())
    O = to_np(state.jacobian(torch.from_numpy(s)))
    Oo = osym.projected_jacobian(osymm, net.forward, net.jacobian, s)
    amp = (scale / np.abs(b)).max()  # cancellation factor of the projection
    assert np.abs(O - Oo).max() <= 1e-6 * amp * max(1.0, np.abs(Oo).max()), (np.abs(O - Oo).max(), amp)


def test_time_evol_step_and_heun_driver(qtx):
    """TimeEvol (quantax/optimizer/time_evol.py:26-134) for the complex state: S = Re(Obar^+ Obar), F = -Im(Obar^+ Ebar),
    direct and chunked accumulation, against the oracle; then one adaptive Heun step (driver.py:44-102) runs and
    keeps the parameters finite."""
    lat, olat = lattice_pair(qtx, "triangular", 6, (18, 18))
    model, net = make_model(qtx, 6, 1, 2, torch.float64, "exp", seed=8, phase=False)
    ns = 160
    assert ns > model.nparams
    s = osmp.rand_states(ns, 36, 18, seed=9)
    H = qtx.operator.Heisenberg()
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat))
    Eo = oop.oloc(aol, net.forward, s)
    Oo = net.jacobian(s)
    for mp in (None, 48):
        state = qtx.state.Variational(model, max_parallel=(4096, mp) if mp else None)
        st = torch.from_numpy(s).cuda()
        samples = qtx.sampler.Samples(st, state(st), None, torch.ones(ns, dtype=torch.float64, device="cuda"))
        tdvp = qtx.optimizer.TimeEvol(state, H)
        S, F = tdvp.get_SF(samples)
        xo, eo, vo, So, Fo = osolver.time_evol_step(Oo, Eo, max_parallel=mp)
        assert np.abs(to_np(S) - So).max() <= 1e-10 * np.abs(So).max()
        assert np.abs(to_np(F) - Fo).max() <= 1e-10 * np.abs(Fo).max()
        assert abs(tdvp.energy - eo) <= 1e-10 * abs(eo) and abs(tdvp.VarE - vo) <= 1e-9 * abs(vo)
        step = to_np(tdvp.get_step(samples))
        check("TimeEvol step", np.linalg.norm(step - xo) / np.linalg.norm(xo), 1e-10)
    with pytest.raises(ValueError):
        bad = qtx.sampler.Samples(st, state(st), None, torch.full((ns,), 2.0, dtype=torch.float64, device="cuda"))
        tdvp.get_step(bad)
    state = qtx.state.Variational(model)
    sampler = qtx.sampler.SpinExchange(state, 256, thermal_steps=20)
    drv = qtx.optimizer.AdaptiveHeunEvolution(state, sampler, qtx.optimizer.TimeEvol(state, H), step_length=1e-3)
    p0 = to_np(state.get_params_flatten()).copy()
    drv.step()
    p1 = to_np(state.get_params_flatten())
    assert np.all(np.isfinite(p1)) and not np.array_equal(p0, p1)
    assert 1e-4 <= drv._step_length <= 1e-2 and len(drv.energy) == 1
