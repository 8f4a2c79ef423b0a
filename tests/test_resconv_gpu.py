"""GPU parity tests of the ResConv path (forward, Jacobian, generic sweep, generic Oloc) against the
CPU oracle.  float64 models: 1e-10 relative; float32 models: 1e-5 (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import models as omodels, operator as oop, sampler as osmp, sites as osites, solver as osolver
from tests.gpu_util import check, lattice_pair, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    return q


def make_resconv(qtx, shape, nblocks, C, k, dtype, final="exp", seed=0, bias_std=0.1):
    net = omodels.ResConv.random(shape, nblocks, C, k, np.float32 if dtype == torch.float32 else np.float64,
                                 seed=seed, final=final, bias_std=bias_std)
    fa = qtx.nn.exp_by_scale if final == "exp" else qtx.nn.sinhp1_by_scale
    model = qtx.model.ResConv(nblocks, C, k, final_activation=fa, dtype=dtype, params=torch.from_numpy(net.params().copy()))
    assert model.nparams == net.nparams
    return model, net


CASES = [
    ("square", 4, (4, 4), 2, 8, 3, "exp"),       # tutorials/J1J2.ipynb:140 ResConv(2, 8, 3) on 4x4
    ("square", 4, (4, 4), 2, 8, 3, "sinhp1"),    # tutorials/J1J2.ipynb:403
    ("square", 6, (6, 6), 3, 5, 3, "exp"),       # channel count not a multiple of the tiles
    ("chain", 12, (1, 12), 2, 4, 3, "exp"),      # 1-D lattice -> Conv1d
    ("square", 10, (10, 10), 2, 36, 3, "exp"),   # two out-channel tiles, N=100 (config C lattice)
    ("square", 6, (6, 6), 2, 4, 5, "sinhp1"),    # 5x5 kernel
]


# lattices / widths that exercise both rasters of the tensor-core forward (csrc/resconv_tc.cu): SEG (16x16),
# RASTER with one tile per sample (10x10) and two (12x12), channel counts off the 16-channel K step
TC_CASES = [
    ("square", 16, (16, 16), 2, 24, 3, "exp"),
    ("square", 16, (16, 16), 3, 40, 3, "sinhp1"),
    ("square", 12, (12, 12), 2, 20, 3, "exp"),
    ("square", 10, (10, 10), 3, 32, 3, "sinhp1"),
    ("square", 8, (8, 8), 2, 12, 3, "exp"),
]


@pytest.mark.parametrize("kind,L,shape,nb,C,k,final", CASES + TC_CASES)
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-11), (torch.float32, 1e-5)])
def test_forward_matches_oracle(qtx, kind, L, shape, nb, C, k, final, dtype, tol):
    lattice_pair(qtx, kind, L)
    model, net = make_resconv(qtx, shape, nb, C, k, dtype, final, seed=1)
    state = qtx.state.Variational(model, max_parallel=7)  # forces several chunks
    N = shape[0] * shape[1]
    s = osmp.rand_states(37, N, seed=2)
    psi = state(torch.from_numpy(s))
    sig, ex = net.forward(s)
    lo = np.log(np.abs(sig)) + ex
    lg = np.log(np.abs(to_np(psi.significand))) + to_np(psi.exponent)
    assert np.array_equal(np.sign(sig), np.sign(to_np(psi.significand)))
    check(f"forward log psi {shape} C={C} {dtype}", np.abs(lg - lo).max() / max(1.0, np.abs(lo).max()), tol)
    assert np.allclose(to_np(psi.exponent), ex, rtol=tol, atol=tol)


def test_tensor_core_forward_is_used_and_matches_cuda_core_path(qtx, monkeypatch):
    """float32 3x3 models run the tcgen05 tower; the CUDA-core float32 path (QTX_RESCONV_TC=0) must agree to
    1e-5 in log psi on a batch large enough for several work items per CTA and a ragged tail."""
    from quantax_b200 import _lib

    lattice_pair(qtx, "square", 16)
    model, net = make_resconv(qtx, (16, 16), 3, 88, 3, torch.float32, "exp", seed=11)
    state = qtx.state.Variational(model)
    s = torch.from_numpy(osmp.rand_states(701, 256, seed=12)).cuda()
    n0 = _lib.lib().qtx_launch_count()
    psi = state(s)
    n_tc = _lib.lib().qtx_launch_count() - n0
    monkeypatch.setenv("QTX_RESCONV_TC", "0")
    n0 = _lib.lib().qtx_launch_count()
    ref = state(s)
    n_fp = _lib.lib().qtx_launch_count() - n0
    monkeypatch.delenv("QTX_RESCONV_TC")
    assert n_tc < n_fp  # one persistent kernel instead of one launch per convolution
    lg = to_np(torch.log(psi.significand.abs()) + psi.exponent)
    lr = to_np(torch.log(ref.significand.abs()) + ref.exponent)
    assert np.abs(lg - lr).max() <= 1e-5 * max(1.0, np.abs(lr).max())
    # and the oracle on a few of the samples
    sig, ex = net.forward(to_np(s[:5]))
    check("tc forward vs float32 oracle", np.abs(lg[:5] - (np.log(np.abs(sig)) + ex)).max() / max(1.0, np.abs(lr).max()), 1e-5)


@pytest.mark.parametrize("kind,L,shape,nb,C,k,final", CASES[:5] + TC_CASES[:1] + TC_CASES[3:4])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_jacobian_matches_oracle(qtx, kind, L, shape, nb, C, k, final, dtype, tol):
    lattice_pair(qtx, kind, L)
    model, net = make_resconv(qtx, shape, nb, C, k, dtype, final, seed=3)
    state = qtx.state.Variational(model, max_parallel=(64, 5))
    N = shape[0] * shape[1]
    s = osmp.rand_states(12, N, seed=4)
    O = to_np(state.jacobian(torch.from_numpy(s)))
    Oo = net.jacobian(s)
    assert O.shape == Oo.shape
    check(f"jacobian {shape} C={C} {dtype}", np.abs(O - Oo).max() / np.abs(Oo).max(), tol)


# shapes served by the tensor-core Jacobian (csrc/resconv_tc.cu: backward-data tower + per-sample weight gradients):
# both rasters, one and two tiles per sample, two and four samples per CTA pair, channel counts that are odd (scalar
# store path), not a multiple of 8 and above one 32-row quarter
TC_BWD_CASES = [
    ((10, 10), 3, 32, "sinhp1"),   # RASTER layout, one tile per sample (config C lattice and width)
    ((12, 12), 2, 20, "exp"),      # RASTER, two tiles per sample (config D lattice)
    ((6, 10), 2, 5, "exp"),        # RASTER, rectangular, odd channel count
    ((4, 4), 2, 8, "sinhp1"),      # RASTER, four samples per CTA pair, K padded from 22 to 32 slots
    ((16, 16), 2, 24, "exp"),
    ((16, 16), 3, 40, "sinhp1"),
    ((16, 16), 2, 5, "exp"),
    ((16, 8), 2, 20, "sinhp1"),
    ((32, 8), 2, 12, "exp"),
    ((16, 16), 1, 16, "exp"),
]


@pytest.mark.parametrize("shape,nb,C,final", TC_BWD_CASES)
def test_tensor_core_jacobian_matches_oracle(qtx, monkeypatch, shape, nb, C, final):
    """variational.py:429-491 on the tensor cores: rows of the log-derivative matrix against the float32 oracle AND the
    float64 evaluation of the same weights (relative 2-norm per sample, bar 1e-5), and the launch list shows that the
    CUDA-core backward pass did not run."""
    from quantax_b200 import _lib

    qtx.sites.Sites._SITES = None
    qtx.sites.Grid(list(shape))
    model, net = make_resconv(qtx, shape, nb, C, 3, torch.float32, final, seed=5)
    state = qtx.state.Variational(model)
    N = shape[0] * shape[1]
    s = osmp.rand_states(11, N, seed=6)
    st = torch.from_numpy(s).cuda()
    assert _lib.lib().qtx_resconv_tc_backward_available(_lib.dtype_code(torch.float32), C, shape[0], shape[1], 3, 3)
    n0 = _lib.lib().qtx_launch_count()
    O = to_np(state.jacobian(st))
    n_tc = _lib.lib().qtx_launch_count() - n0
    monkeypatch.setenv("QTX_RESCONV_TC_BWD", "0")
    n0 = _lib.lib().qtx_launch_count()
    Oc = to_np(state.jacobian(st))
    n_fp = _lib.lib().qtx_launch_count() - n0
    monkeypatch.delenv("QTX_RESCONV_TC_BWD")
    if nb >= 2:  # two persistent towers + one weight-gradient launch instead of four launches per block
        assert n_tc < n_fp
    assert n_tc != n_fp
    Oo = net.jacobian(s)
    blocks = [{k: (None if v is None else v.astype(np.float64)) for k, v in blk.items()} for blk in net.blocks]
    O64 = omodels.ResConv(blocks, net.shape, net.final).jacobian(s)
    nrm = np.linalg.norm(O64, axis=1)
    check(f"tc jacobian {shape} C={C} rows vs float64 evaluation", (np.linalg.norm(O - O64, axis=1) / nrm).max(), 1e-5)
    check(f"tc jacobian {shape} C={C} rows vs float32 oracle", (np.linalg.norm(O - Oo, axis=1) / nrm).max(), 1e-5)
    check(f"tc jacobian {shape} C={C} rows vs CUDA-core backward", (np.linalg.norm(O - Oc, axis=1) / nrm).max(), 1e-5)
    check(f"tc jacobian {shape} C={C} worst entry vs float64 evaluation", np.abs(O - O64).max() / np.abs(O64).max(), 1e-5)


def test_tensor_core_jacobian_float32_rows_and_chunks(qtx):
    """float32 output rows (the C ABI's out_dtype) and several backward chunks give the same matrix."""
    qtx.sites.Sites._SITES = None
    qtx.sites.Grid([16, 16])
    model, net = make_resconv(qtx, (16, 16), 2, 24, 3, torch.float32, "exp", seed=7)
    s = torch.from_numpy(osmp.rand_states(13, 256, seed=8)).cuda()
    O = qtx.state.Variational(model).jacobian(s)
    O32 = torch.empty((13, model.nparams), dtype=torch.float32, device="cuda")
    qtx.state.Variational(model).jacobian(s, out=O32)
    assert torch.equal(O32, O.to(torch.float32))  # the same float32 values, rounded once
    Oc = qtx.state.Variational(model, max_parallel=(64, 5)).jacobian(s)  # chunks of 5, 5, 3 samples
    assert torch.equal(O, Oc)


@pytest.mark.parametrize("kind", ["localflip", "exchange"])
def test_generic_sweep_bit_exact_f64(qtx, kind):
    """ResConv has no local-update path: full forward per proposal; accept/reject pattern and chains
    identical to the oracle for injected randoms (float64 parameters)."""
    nup = (8, 8) if kind == "exchange" else None
    lat, olat = lattice_pair(qtx, "square", 4, nup)
    model, net = make_resconv(qtx, (4, 4), 2, 4, 3, torch.float64, "exp", seed=5)
    state = qtx.state.Variational(model)
    ns, T = 64, 40
    cls = qtx.sampler.LocalFlip if kind == "localflip" else qtx.sampler.SpinExchange
    sampler = cls(state, ns, thermal_steps=0)
    assert not state.use_ref
    spins0 = to_np(sampler._spins).copy()
    rng = np.random.default_rng(6)
    table = osites.site_neighbor_table(olat) if kind == "exchange" else None
    u = rng.random((T, ns)); pos = rng.integers(0, 16, size=(T, ns))
    slot = None if table is None else rng.integers(0, table.shape[1], size=(T, ns))
    sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), None if slot is None else torch.from_numpy(slot))
    samples = sampler.sweep(T, record=True)
    ref = osmp.sweep(osmp.FullForwardChainModel(net), spins0, T, kind, neighbors=table, pos=pos, slot=slot, u=u, record=True)
    assert np.array_equal(to_np(sampler.last_accept_log), ref["accept_log"])
    assert np.array_equal(to_np(samples.spins), ref["spins"])
    lg = np.log(np.abs(to_np(samples.psi.significand))) + to_np(samples.psi.exponent)
    lo = np.log(np.abs(ref["psi"][0])) + ref["psi"][1]
    assert np.allclose(lg, lo, rtol=1e-11, atol=1e-11)
    # the production Philox stream gives the same chains as the oracle's restatement
    spins1 = to_np(sampler._spins).copy()
    samples2 = sampler.sweep(24)
    ref2 = osmp.sweep(osmp.FullForwardChainModel(net), spins1, 24, kind, neighbors=table, seed=sampler._seed, step0=T)
    assert np.array_equal(to_np(samples2.spins), ref2["spins"])


@pytest.mark.parametrize("final", ["exp", "sinhp1"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_resconv_oloc_and_sr_step(qtx, final, dtype, tol):
    """Oloc through enumerate -> forward -> reduce for a J1-J2 Hamiltonian, then a full SR step."""
    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, net = make_resconv(qtx, (4, 4), 2, 6, 3, dtype, final, seed=7)
    state = qtx.state.Variational(model)
    H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat, J=[1, 0.5], n_neighbor=[1, 2], msr=True))
    ns = 48
    s = osmp.rand_states(ns, 16, 8, seed=8)
    st = torch.from_numpy(s).cuda()
    E = to_np(H.Oloc(state, st))
    Eo = oop.oloc(aol, net.forward, s)
    check(f"ResConv Oloc {final} {dtype}", np.abs(E - Eo).max() / np.abs(Eo).max(), tol)
    samples = qtx.sampler.Samples(st, state(st), None, torch.ones(ns, dtype=torch.float64, device="cuda"))
    opt = qtx.optimizer.SR(state, H)
    step = to_np(opt.get_step(samples))
    xo, eo, vo = osolver.sr_step(net.jacobian(s), Eo, np.ones(ns))
    assert abs(opt.energy - eo) <= tol * abs(eo)
    if dtype == torch.float64:
        check(f"ResConv SR step {final}", np.linalg.norm(step - xo) / np.linalg.norm(xo), 1e-10)


def test_resconv_vmc_converges_on_4x4_heisenberg(qtx):
    """tutorials/J1J2.ipynb: ResConv(2, 8, 3), SpinExchange, SR on the 4x4 Heisenberg model approaches the
    ED energy -44.91393283 (printed at J1J2.ipynb:90); here a shortened run must get within 1 %."""
    qtx.set_random_seed(42)
    lattice_pair(qtx, "square", 4, (8, 8))
    H = qtx.operator.Heisenberg(msr=True)
    model = qtx.model.ResConv(nblocks=2, channels=8, kernel_size=3)
    state = qtx.state.Variational(model)
    sampler = qtx.sampler.SpinExchange(state, nsamples=512)
    optimizer = qtx.optimizer.SR(state, H)
    hist = []
    for i in range(80):
        samples = sampler.sweep()
        step = optimizer.get_step(samples)
        state.update(step * 0.02)
        hist.append(optimizer.energy)
    e = np.mean(hist[-10:])
    assert e > -44.913932833715506 - 0.3 and e < -44.913932833715506 * 0.99, (e, hist[::10])


def test_tutorial_run_reaches_the_exact_ground_state_energy(qtx):
    """tutorials/J1J2.ipynb cells 7-19 end to end: ResConv(2, 8, 3) on the 4x4 Heisenberg model, 200 SR steps of
    1024 SpinExchange samples at rate 0.01, then 200 steps with the Rotation @ Flip @ SpinInverse projected state.
    The tutorial evaluates <psi|H|psi> of the dense projected state and prints a relative error of 6.3e-6
    (J1J2.ipynb:331) against ED, -44.913932833715506 (J1J2.ipynb:90).  Here the same quantity is evaluated WITHOUT
    sampling noise through the product path itself: all 12 870 states of the S_z = 0 sector, psi and Oloc on the GPU."""
    import itertools

    from quantax_b200.symmetry import Flip, Rotation, SpinInverse

    E0 = -44.913932833715506
    qtx.set_random_seed(7)
    lattice_pair(qtx, "square", 4, (8, 8))
    H = qtx.operator.Heisenberg(msr=True)
    model = qtx.model.ResConv(nblocks=2, channels=8, kernel_size=3)
    state = qtx.state.Variational(model, max_parallel=25000)
    sampler = qtx.sampler.SpinExchange(state, nsamples=1024)
    optimizer = qtx.optimizer.SR(state, H)
    for _ in range(200):
        state.update(optimizer.get_step(sampler.sweep()) * 0.01)
    symm = Rotation(np.pi / 2) @ Flip() @ SpinInverse()
    symm_state = qtx.state.Variational(model, symm=symm, max_parallel=2048)
    sampler = qtx.sampler.SpinExchange(symm_state, nsamples=1024)
    optimizer = qtx.optimizer.SR(symm_state, H)
    hist = []
    for _ in range(200):
        symm_state.update(optimizer.get_step(sampler.sweep()) * 0.01)
        hist.append(float(optimizer.energy))
    # exact variational energy of the projected state over the whole S_z = 0 sector
    basis = np.full((12870, 16), -1, dtype=np.int8)
    for r, up in enumerate(itertools.combinations(range(16), 8)):
        basis[r, list(up)] = 1
    sb = torch.from_numpy(basis).cuda()
    psi = symm_state(sb)
    w = torch.exp(2.0 * (qtx.utils.log_abs(psi) - qtx.utils.log_abs(psi).max()))
    samples = qtx.sampler.Samples(sb, psi, None, torch.ones(len(basis), dtype=torch.float64, device="cuda"))
    El = H.Oloc(symm_state, samples)
    e = float((w * El.real).sum() / w.sum())
    assert e >= E0 - 1e-9 * abs(E0)  # variational
    check("4x4 Heisenberg: exact energy of the trained projected ResConv vs ED (tutorial: 6.3e-6)", abs(e - E0) / abs(E0), 2e-5)
    check("4x4 Heisenberg: sampled energy of the last 20 steps vs ED", abs(np.mean(hist[-20:]) - E0) / abs(E0), 1e-3)


def test_state_save_load_eqx_layout(qtx, tmp_path):
    """Variational.save / load (variational.py:162-163,581-587) through the equinox leaf layout."""
    from quantax_b200.utils import read_eqx_leaves

    lattice_pair(qtx, "square", 4)
    model, net = make_resconv(qtx, (4, 4), 2, 4, 3, torch.float32, "exp", seed=21)
    state = qtx.state.Variational(model)
    f = tmp_path / "resconv.eqx"
    state.save(f)
    leaves = read_eqx_leaves(f)
    shapes = [l.shape for l in leaves if l.ndim]
    assert shapes == [(4, 1, 3, 3), (4, 1, 1), (4, 4, 3, 3), (4, 1, 1), (4, 4, 3, 3), (4, 1, 1), (4, 4, 3, 3)]
    model2 = qtx.model.ResConv(2, 4, 3)
    state2 = qtx.state.Variational(model2, param_file=f)
    assert torch.equal(state2.get_params_flatten(), state.get_params_flatten())
    lattice_pair(qtx, "chain", 8)
    from tests.gpu_util import make_rbm

    rbm, _ = make_rbm(qtx, 8, 16, torch.float32, seed=22)
    st = qtx.state.Variational(rbm)
    g = tmp_path / "rbm.eqx"
    st.save(g)
    assert [l.shape for l in read_eqx_leaves(g)] == [(16, 8), (16,), ()]
    st2 = qtx.state.Variational(qtx.model.RBM_Dense(16), param_file=g)
    assert torch.equal(st2.get_params_flatten(), st.get_params_flatten())


def test_projected_sweep_of_moved_proposals_only_equals_full_sweep(qtx, monkeypatch):
    """The same for a symmetry-projected float32 ResConv (C4v x Z2, 16 images per proposal): only the images of the
    moved chains are forwarded."""
    lat, olat = lattice_pair(qtx, "square", 6, (18, 18))
    model, net = make_resconv(qtx, (6, 6), 2, 8, 3, torch.float32, "sinhp1", seed=81)
    S = qtx.symmetry
    state = qtx.state.Variational(model, symm=S.C4v() @ S.SpinInverse())
    assert state.symm.nsymm == 16
    ns, T = 64, 12
    rng = np.random.default_rng(82)
    u = rng.random((T, ns)); pos = rng.integers(0, 36, size=(T, ns)); slot = rng.integers(0, 4, size=(T, ns))
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("QTX_SWEEP_COMPACT", mode)
        sampler = qtx.sampler.SpinExchange(state, ns, thermal_steps=0,
                                           initial_spins=torch.from_numpy(osmp.rand_states(ns, 36, 18, seed=83)))
        sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), torch.from_numpy(slot))
        samples = sampler.sweep(T, record=True)
        out[mode] = (to_np(samples.spins), to_np(sampler.last_accept_log), to_np(samples.psi.mult), to_np(samples.psi.expo))
    monkeypatch.delenv("QTX_SWEEP_COMPACT")
    for a, b in zip(out["1"], out["0"]):
        assert np.array_equal(a, b)
    assert 0.1 < out["1"][1].mean() < 0.9


@pytest.mark.parametrize("cplx", [False, True])
def test_sweep_of_moved_proposals_only_equals_full_sweep(qtx, monkeypatch, cplx):
    """Exchange proposals of two equal spins can never be accepted (metropolis.py:314-316); the float32 tensor-core
    path evaluates psi of the moved proposals only (device-side batch size).  Chains, accept pattern and carried psi
    must be identical to the sweep that evaluates every proposal, and match the oracle away from near-ties."""
    if cplx:
        qtx.set_default_dtype(torch.complex128)
    try:
        lat, olat = lattice_pair(qtx, "square", 6, (18, 18))
        net = omodels.ResConv.random((6, 6), 2, 8, 3, np.float32, seed=71, bias_std=0.1, out_complex=cplx)
        model = qtx.model.ResConv(2, 8, 3, out_dtype=torch.complex128 if cplx else None,
                                  params=torch.from_numpy(net.params().copy()))
        state = qtx.state.Variational(model)
        ns, T = 96, 24
        rng = np.random.default_rng(72)
        table = osites.site_neighbor_table(olat)
        u = rng.random((T, ns)); pos = rng.integers(0, 36, size=(T, ns)); slot = rng.integers(0, 4, size=(T, ns))
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("QTX_SWEEP_COMPACT", mode)
            sampler = qtx.sampler.SpinExchange(state, ns, thermal_steps=0,
                                               initial_spins=torch.from_numpy(osmp.rand_states(ns, 36, 18, seed=73)))
            spins0 = to_np(sampler._spins).copy()
            sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), torch.from_numpy(slot))
            samples = sampler.sweep(T, record=True)
            out[mode] = (to_np(samples.spins), to_np(sampler.last_accept_log), to_np(samples.psi.mult), to_np(samples.psi.expo))
        monkeypatch.delenv("QTX_SWEEP_COMPACT")
        for a, b in zip(out["1"], out["0"]):
            assert np.array_equal(a, b)
        assert 0.2 < out["1"][1].mean() < 0.9
        ref = osmp.sweep(osmp.FullForwardChainModel(net), spins0, T, "exchange", neighbors=table, pos=pos, slot=slot, u=u,
                         record=True)
        same = (ref["spins"] == out["1"][0]).all(axis=1)
        assert same.mean() > 0.9  # float32 model: chains can only differ after a provable near-tie
        first = np.array([np.argmax(ref["accept_log"][:, c] != out["1"][1][:, c]) for c in np.flatnonzero(~same)])
        for c, t in zip(np.flatnonzero(~same), first):
            assert ref["margin"][t, c] < 1e-3
    finally:
        qtx.set_default_dtype(torch.float64)


@pytest.mark.parametrize("kind,dtype", [("exchange", torch.float32), ("localflip", torch.float32), ("exchange", torch.float64)])
def test_one_call_sweep_equals_the_step_by_step_sweep(qtx, kind, dtype):
    """qtx_resconv_sweep (the whole sweep behind one C-ABI call: what a jax.ffi binder would register) against the
    step-by-step entry points driven from Python, on the production Philox stream: identical chains, accept counts
    and amplitudes."""
    nup = (32, 32) if kind == "exchange" else None
    outs = []
    for record in (False, True):  # record=True keeps the Python loop (it needs the per-step accept log)
        qtx.set_random_seed(77)
        lattice_pair(qtx, "square", 8, nup)
        model, _ = make_resconv(qtx, (8, 8), 2, 16, 3, dtype, "sinhp1", seed=9)
        state = qtx.state.Variational(model)
        cls = qtx.sampler.SpinExchange if kind == "exchange" else qtx.sampler.LocalFlip
        sampler = cls(state, 96, thermal_steps=0)
        samples = sampler.sweep(40, record=record)
        outs.append((to_np(samples.spins), to_np(samples.psi.significand), to_np(samples.psi.exponent),
                     to_np(sampler.last_naccept)))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
