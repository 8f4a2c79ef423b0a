"""Full-size (BASELINE.json configs B / C / E shapes) checks through size-independent properties:
conservation laws of the sampler, agreement of independent code paths, translation covariance."""
import numpy as np
import pytest
import torch

from tests.gpu_util import to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    return q


def test_config_b_full_size_step_properties(qtx):
    """10x10 Heisenberg, RBM_Dense alpha=4, SpinExchange, Ns=4096 (BASELINE configs[1])."""
    qtx.set_random_seed(7)
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(10, Nparticles=(50, 50))
    H = qtx.operator.Heisenberg(msr=True)
    model = qtx.model.RBM_Dense(features=400)
    state = qtx.state.Variational(model)
    sampler = qtx.sampler.SpinExchange(state, nsamples=4096)
    samples = sampler.sweep(record=True)
    s = samples.spins
    assert s.shape == (4096, 100) and s.dtype == torch.int8
    assert bool((s.abs() == 1).all()) and bool((s.sum(dim=1) == 0).all())  # Sz = 0 sector conserved
    acc = sampler.last_accept_log.float().mean().item()
    assert 0.05 < acc < 0.8
    assert bool((sampler.last_naccept == sampler.last_accept_log.sum(dim=0)).all())
    assert sampler.check_local_updates(samples) == 0  # metropolis.py:201-212 drift check
    # psi returned by the sweep == direct forward; fused Oloc == enumerate/forward/reduce path
    assert torch.allclose(samples.psi.logabs, state(s).logabs, rtol=1e-5)  # float32 model: summation order differs
    E_fused = H.Oloc(state, samples)
    E_generic = H.Oloc(qtx.state.Variational(model, use_ref=False), samples)
    assert (E_fused - E_generic).abs().max().item() <= 2e-4 * E_fused.abs().max().item()
    # connected-configuration counts: every antiparallel nearest-neighbour bond, nothing else
    seg, cidx, sc, Hc, n_nonnan = H.get_conn(s, 2, with_spins=False)
    bonds = qtx.get_sites().get_neighbor(1)
    sn = to_np(s)
    anti = (sn[:, bonds[:, 0]] != sn[:, bonds[:, 1]]).sum()
    assert n_nonnan == anti == seg.numel()
    assert bool((Hc == -2.0).all())  # Marshall sign rule on nearest neighbours: 2 * (-1) * J
    # SR step: finite, and the update decreases nothing catastrophically (energy stays finite)
    opt = qtx.optimizer.SR(state, H)
    step = opt.get_step(samples)
    assert step.shape == (40400,) and bool(torch.isfinite(step).all())
    assert np.isfinite(opt.energy) and opt.VarE > 0
    # Obar is centred: column sums vanish; T = Obar Obar^T has a null vector of ones
    Obar = opt.get_Obar(samples)
    assert Obar.sum(dim=0).abs().max().item() < 1e-10
    assert Obar.shape == (4096, 40400) and Obar.dtype == torch.float64


@pytest.mark.parametrize("L,C,nb,ns", [(10, 32, 8, 512), (16, 88, 8, 64)])
def test_resconv_translation_invariance_full_size(qtx, L, C, nb, ns):
    """Configs C / E network shapes: the translation-symmetrised ResConv amplitude (nn/conv.py:61-68) and its
    local energies are invariant under lattice translations of the input; Jacobian rows too."""
    qtx.set_random_seed(11)
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(L, Nparticles=(L * L // 2, L * L // 2))
    model = qtx.model.ResConv(nb, C, 3, final_activation=qtx.nn.sinhp1_by_scale)
    state = qtx.state.Variational(model)
    s = qtx.utils.rand_states(ns)
    shifted = torch.roll(s.view(ns, L, L), shifts=(3, L - 2), dims=(1, 2)).reshape(ns, L * L).contiguous()
    a, b = state(s), state(shifted)
    la = torch.log(a.significand.abs()) + a.exponent
    lb = torch.log(b.significand.abs()) + b.exponent
    assert bool(torch.isfinite(la).all())
    assert (la - lb).abs().max().item() < 2e-4
    assert bool((torch.sign(a.significand) == torch.sign(b.significand)).all())
    if L == 10:
        H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
        Ea, Eb = H.Oloc(state, s[:64]), H.Oloc(state, shifted[:64])
        assert (Ea - Eb).abs().max().item() < 1e-2 * Ea.abs().max().item()
        Ja, Jb = state.jacobian(s[:16]), state.jacobian(shifted[:16])
        assert (Ja - Jb).abs().max().item() < 1e-3 * Ja.abs().max().item()
        assert Ja.shape == (16, model.nparams)


def test_config_e_slice_tensor_core_tower(qtx, monkeypatch):
    """Config E network (16x16, ResConv 8 blocks x 88 channels, sinh+1, 1 047 552 parameters) on one GPU's 2048 chains:
    the CTA-pair tensor-core tower, the single-CTA tower and (on a subset) the float64 model agree within the float32
    bar; a short exchange sweep conserves Sz and carries the same psi as a direct forward of the final chains."""
    qtx.set_random_seed(13)
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(16, Nparticles=(128, 128))
    model = qtx.model.ResConv(8, 88, 3, final_activation=qtx.nn.sinhp1_by_scale)
    assert model.nparams == 1047552
    state = qtx.state.Variational(model)
    s = qtx.utils.rand_states(2048)
    lg = lambda psi: torch.log(psi.significand.abs()) + psi.exponent
    a = state(s)
    monkeypatch.setenv("QTX_TC_2CTA", "0")
    b = state(s)
    monkeypatch.delenv("QTX_TC_2CTA")
    assert bool(torch.isfinite(lg(a)).all())
    assert (lg(a) - lg(b)).abs().max().item() < 2e-6  # same arithmetic, different tiling
    assert bool((torch.sign(a.significand) == torch.sign(b.significand)).all())
    m64 = qtx.model.ResConv(8, 88, 3, final_activation=qtx.nn.sinhp1_by_scale, dtype=torch.float64,
                            params=model.params.double())
    c = qtx.state.Variational(m64)(s[:48])
    assert (lg(a)[:48] - lg(c)).abs().max().item() < 1e-5 * max(1.0, lg(c).abs().max().item())
    sampler = qtx.sampler.SpinExchange(state, nsamples=2048, thermal_steps=0, initial_spins=s)
    samples = sampler.sweep(8)
    assert bool((samples.spins.sum(dim=1) == 0).all())
    direct = state(samples.spins)
    assert (lg(samples.psi) - lg(direct)).abs().max().item() < 1e-6
    assert 0 < int(sampler.last_naccept.sum().item()) < 8 * 2048
