"""Symmetry-projected states (variational.py:262-266,438-491; symmetry.py:325-392) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import models as omodels, operator as oop, sampler as osmp, sites as osites, solver as osolver
from oracle import symmetry as osym
from tests.gpu_util import check, lattice_pair, make_rbm, to_np
from tests.test_resconv_gpu import make_resconv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    return q


def _symms(qtx, olat, which):
    S = qtx.symmetry
    if which == "c4v_z2":  # tutorials/J1J2.ipynb: Rotation(pi/2) @ Flip() @ SpinInverse()
        return (S.Rotation(np.pi / 2) @ S.Flip() @ S.SpinInverse(),
                osym.Rotation(olat, np.pi / 2) @ osym.Flip(olat) @ osym.SpinInverse(olat))
    if which == "b1_odd":  # non-trivial characters and odd spin inversion
        return (S.C4v(repr="B1") @ S.SpinInverse(-1),
                osym.Rotation(olat, np.pi / 2, sector=2) @ osym.Flip(olat) @ osym.SpinInverse(olat, -1))
    if which == "trans":
        return S.TransND(), osym.TransND(olat)
    raise ValueError


def _logpsi(psi):
    return np.log(np.abs(to_np(psi.mult))) + to_np(psi.expo), np.sign(to_np(psi.mult))


@pytest.mark.parametrize("which", ["c4v_z2", "b1_odd", "trans"])
@pytest.mark.parametrize("model_kind", ["rbm", "resconv"])
def test_projected_amplitude_and_jacobian(qtx, which, model_kind):
    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    if model_kind == "rbm":
        model, net = make_rbm(qtx, 16, 24, torch.float64, seed=31)
    else:
        model, net = make_resconv(qtx, (4, 4), 2, 4, 3, torch.float64, "sinhp1", seed=32)
    symm, osymm = _symms(qtx, olat, which)
    state = qtx.state.Variational(model, symm=symm, max_parallel=(4096, 64))
    assert state.symm.nsymm == osymm.nsymm and not state.use_ref
    s = osmp.rand_states(19, 16, 8, seed=33)
    sign_o, la_o, _ = osym.project(osymm, net.forward, s)
    la, sign = _logpsi(state(torch.from_numpy(s)))
    ok = np.isfinite(la_o)
    assert np.array_equal(sign[ok], sign_o[ok])
    assert np.abs(la[ok] - la_o[ok]).max() <= 1e-10 * max(1.0, np.abs(la_o[ok]).max())
    # covariance: psi(T_g s) = chi_g psi(s) for every group element (spot-check three of them)
    for g in (1, osymm.perm.shape[0] // 2, osymm.perm.shape[0] - 1):
        la_g, sign_g = _logpsi(state(torch.from_numpy(np.ascontiguousarray(s[:, osymm.perm[g]]))))
        assert np.allclose(la_g[ok], la[ok], rtol=1e-9, atol=1e-9)
        assert np.array_equal(sign_g[ok], sign[ok] * np.sign(osymm.character[g]))
    O = to_np(state.jacobian(torch.from_numpy(s)))
    Oo = osym.projected_jacobian(osymm, net.forward, net.jacobian, s)
    check("projected jacobian", np.abs(O[ok] - Oo[ok]).max() / max(1.0, np.abs(Oo[ok]).max()), 1e-10)


def test_projected_state_full_step(qtx):
    """Sweep (generic path), Oloc and an SR step with a C4v x Z2 projected RBM; injected randoms make the
    accept/reject pattern comparable bit for bit."""
    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, net = make_rbm(qtx, 16, 12, torch.float64, seed=34)
    symm, osymm = _symms(qtx, olat, "c4v_z2")
    state = qtx.state.Variational(model, symm=symm)
    fwd = lambda x: osym.project(osymm, net.forward, x)[:2]
    ns, T = 48, 30
    sampler = qtx.sampler.SpinExchange(state, ns, thermal_steps=0)
    spins0 = to_np(sampler._spins).copy()
    rng = np.random.default_rng(35)
    table = osites.site_neighbor_table(olat)
    u = rng.random((T, ns)); pos = rng.integers(0, 16, size=(T, ns)); slot = rng.integers(0, 4, size=(T, ns))
    sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), torch.from_numpy(slot))
    samples = sampler.sweep(T, record=True)

    class Proj:
        def forward(self, x):
            return fwd(x)

    ref = osmp.sweep(osmp.FullForwardChainModel(Proj()), spins0, T, "exchange", neighbors=table, pos=pos, slot=slot, u=u,
                     record=True)
    assert np.array_equal(to_np(sampler.last_accept_log), ref["accept_log"])
    assert np.array_equal(to_np(samples.spins), ref["spins"])
    H = qtx.operator.Heisenberg(msr=True)
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat, msr=True))
    s = ref["spins"]
    Eo = oop.oloc(aol, fwd, s)
    opt = qtx.optimizer.SR(state, H)
    step = to_np(opt.get_step(samples))
    check("projected Oloc", np.abs(to_np(opt._Eloc) - Eo).max() / np.abs(Eo).max(), 1e-10)
    Oo = osym.projected_jacobian(osymm, net.forward, net.jacobian, s)
    xo, eo, vo = osolver.sr_step(Oo, Eo, np.ones(ns))
    assert abs(opt.energy - eo) <= 1e-10 * abs(eo)
    check("projected SR step", np.linalg.norm(step - xo) / np.linalg.norm(xo), 1e-10)
