"""GPU parity tests of the RBM hot path against the CPU oracle, through the C ABI.

Tolerances (BASELINE.json north_star): enumeration and accept/reject bit-exact given the same
random stream; float64 models 1e-10 relative; float32 models 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import models as omodels, operator as oop, sampler as osmp, sites as osites, solver as osolver
from tests.gpu_util import chains_equal, check, lattice_pair, make_rbm, sr_step_tolerance, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qtx():
    import quantax_b200 as q

    torch.cuda.set_device(0)
    return q


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-6)])
def test_forward_and_internal(qtx, dtype, tol):
    lattice_pair(qtx, "square", 6)
    model, net = make_rbm(qtx, 36, 100, dtype, seed=1)
    state = qtx.state.Variational(model)
    s = osmp.rand_states(300, 36, seed=2)
    psi = state(torch.from_numpy(s))
    sign, la = net.forward(s)
    assert np.allclose(to_np(psi.logabs), la, rtol=tol, atol=tol * 10)
    assert np.array_equal(to_np(psi.sign), sign)
    theta = state.init_internal(torch.from_numpy(s))
    assert np.allclose(to_np(theta), net.init_internal(s), rtol=tol * 10, atol=tol * 10)


def _inject(rng, nsweeps, ns, N, nb_table=None, spins=None):
    u = rng.random((nsweeps, ns))
    pos = rng.integers(0, N, size=(nsweeps, ns))
    slot = None if nb_table is None else rng.integers(0, nb_table.shape[1], size=(nsweeps, ns))
    return pos, slot, u


@pytest.mark.parametrize("kind", ["localflip", "exchange"])
def test_sweep_injected_randoms_bit_exact_f64(qtx, kind):
    """float64 parameters: accept/reject pattern and final chains identical to the oracle for the
    same injected proposal sites and uniforms (metropolis.py:291-322)."""
    nup = (18, 18) if kind == "exchange" else None
    lat, olat = lattice_pair(qtx, "square", 6, nup)
    model, net = make_rbm(qtx, 36, 72, torch.float64, seed=3)
    state = qtx.state.Variational(model)
    ns, T = 512, 90
    cls = qtx.sampler.LocalFlip if kind == "localflip" else qtx.sampler.SpinExchange
    sampler = cls(state, ns, thermal_steps=0)
    spins0 = to_np(sampler._spins).copy()
    rng = np.random.default_rng(5)
    table = osites.site_neighbor_table(olat) if kind == "exchange" else None
    pos, slot, u = _inject(rng, T, ns, 36, table)
    sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), None if slot is None else torch.from_numpy(slot))
    samples = sampler.sweep(T, record=True)
    ref = osmp.sweep(osmp.RBMChainModel(net), spins0, T, kind, neighbors=table, pos=pos, slot=slot, u=u, record=True)
    log = to_np(sampler.last_accept_log)
    assert np.array_equal(log, ref["accept_log"]), f"{(log != ref['accept_log']).sum()} accept decisions differ"
    assert np.array_equal(to_np(samples.spins), ref["spins"])
    assert np.array_equal(to_np(sampler.last_naccept), ref["naccept"])
    assert np.allclose(to_np(samples.psi.logabs), ref["psi"][1], rtol=1e-12)
    assert np.allclose(to_np(sampler.last_psi_chain.logabs), ref["psi_chain"][1], rtol=1e-11)
    if kind == "exchange":
        assert (to_np(samples.spins).sum(axis=1) == 0).all()  # magnetisation conserved
    assert 0.05 < log.mean() < 0.95


@pytest.mark.parametrize("kind,hopcase", [("localflip", 0), ("exchange", 1), ("exchange", -1)])
def test_sweep_philox_stream_matches_oracle_f64(qtx, kind, hopcase):
    """Production RNG path: in-kernel Philox4x32-10 == the oracle's restatement, incl. the k-th
    hopping-particle search and the neighbour slot draw; hop=-1 when Nup > N/2."""
    nup = None if kind == "localflip" else ((12, 24) if hopcase == 1 else (24, 12))
    lat, olat = lattice_pair(qtx, "square", 6, nup)
    model, net = make_rbm(qtx, 36, 40, torch.float64, seed=4)
    state = qtx.state.Variational(model)
    ns, T = 300, 70  # T > 64 exercises two Philox refills + a partial one
    cls = qtx.sampler.LocalFlip if kind == "localflip" else qtx.sampler.SpinExchange
    sampler = cls(state, ns, thermal_steps=0)
    assert kind == "localflip" or sampler._hopping_particle == hopcase
    spins0 = to_np(sampler._spins).copy()
    samples = sampler.sweep(T, record=True)
    samples2 = sampler.sweep(T)  # the stream continues with step0 = T
    table = osites.site_neighbor_table(olat) if kind == "exchange" else None
    ref = osmp.sweep(osmp.RBMChainModel(net), spins0, T, kind, neighbors=table, hop=hopcase or 1,
                     seed=sampler._seed, step0=0, record=True)
    assert np.array_equal(to_np(sampler.last_accept_log) if False else to_np(samples.spins), ref["spins"])
    ref2 = osmp.sweep(osmp.RBMChainModel(net), ref["spins"], T, kind, neighbors=table, hop=hopcase or 1,
                      seed=sampler._seed, step0=T)
    assert np.array_equal(to_np(samples2.spins), ref2["spins"])


@pytest.mark.parametrize("kind", ["localflip", "exchange"])
def test_sweep_f32_matches_up_to_near_ties(qtx, kind):
    """float32 parameters (reference default): log|psi| differs from the oracle in the last float32
    bits, so an accept decision may flip only where |ratio - (1-u)| is within float32 rounding.
    Every differing chain must diverge at such a near-tie; all others must be identical."""
    nup = (18, 18) if kind == "exchange" else None
    lat, olat = lattice_pair(qtx, "square", 6, nup)
    model, net = make_rbm(qtx, 36, 72, torch.float32, seed=6)
    state = qtx.state.Variational(model)
    ns, T = 1024, 72
    cls = qtx.sampler.LocalFlip if kind == "localflip" else qtx.sampler.SpinExchange
    sampler = cls(state, ns, thermal_steps=0)
    spins0 = to_np(sampler._spins).copy()
    table = osites.site_neighbor_table(olat) if kind == "exchange" else None
    samples = sampler.sweep(T, record=True)
    ref = osmp.sweep(osmp.RBMChainModel(net), spins0, T, kind, neighbors=table, seed=sampler._seed, step0=0,
                     record=True)
    log, rlog = to_np(sampler.last_accept_log), ref["accept_log"]
    differ = np.flatnonzero((log != rlog).any(axis=0))
    assert differ.size <= max(2, ns // 100), f"{differ.size} chains diverged"
    for c in differ:
        t = int(np.flatnonzero(log[:, c] != rlog[:, c])[0])
        assert ref["margin"][t, c] < 1e-4, f"chain {c} diverged at step {t} with margin {ref['margin'][t, c]}"
    same = np.setdiff1d(np.arange(ns), differ)
    assert np.array_equal(to_np(samples.spins)[same], ref["spins"][same])
    assert np.allclose(to_np(samples.psi.logabs)[same], ref["psi"][1][same], rtol=1e-5)


def _hamiltonians(qtx, which):
    if which == "ising":
        lat, olat = lattice_pair(qtx, "chain", 16)
        return qtx.operator.Ising(h=0.8, J=1.3), oop.ising_op_list(olat, h=0.8, J=1.3), 16, None
    if which == "heis":
        lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
        return qtx.operator.Heisenberg(msr=True), oop.heisenberg_op_list(olat, msr=True), 16, 8
    if which == "j1j2":
        lat, olat = lattice_pair(qtx, "square", 6, (18, 18))
        return (qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True),
                oop.heisenberg_op_list(olat, J=[1, 0.5], n_neighbor=[1, 2], msr=True), 36, 18)
    if which == "tri":
        lat, olat = lattice_pair(qtx, "triangular", 6, (18, 18))
        return qtx.operator.Heisenberg(), oop.heisenberg_op_list(olat), 36, 18
    raise ValueError


@pytest.mark.parametrize("which", ["ising", "heis", "j1j2", "tri"])
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-5)])
def test_oloc_matches_oracle(qtx, which, dtype, tol):
    H, ol, N, nup = _hamiltonians(qtx, which)
    model, net = make_rbm(qtx, N, 3 * N, dtype, seed=7)
    state = qtx.state.Variational(model)
    s = osmp.rand_states(257, N, nup, seed=8)
    E = to_np(H.Oloc(state, torch.from_numpy(s)))
    Eo = oop.oloc(oop.to_array_op_list(ol), net.forward, s)
    scale = np.abs(Eo).max()
    check(f"Oloc fused {which} {dtype}", np.abs(E - Eo).max() / scale, tol)
    # generic path (enumerate -> forward -> reduce) == fused local-update path (local_updates.ipynb:354)
    state_direct = qtx.state.Variational(model, use_ref=False)
    E2 = to_np(H.Oloc(state_direct, torch.from_numpy(s)))
    check(f"Oloc generic {which} {dtype}", np.abs(E2 - Eo).max() / scale, tol)


@pytest.mark.parametrize("which", ["ising", "heis", "j1j2", "tri"])
def test_connected_enumeration_bit_exact(qtx, which):
    """(segment, conn index order, s', H) identical to _apply_off_diag + _get_conn, incl. padding."""
    H, ol, N, nup = _hamiltonians(qtx, which)
    aol = oop.to_array_op_list(ol)
    s = osmp.rand_states(130, N, nup, seed=9)
    st = torch.from_numpy(s)
    assert np.array_equal(to_np(H.apply_diag(st)), oop.apply_diag(s, aol)) or np.allclose(
        to_np(H.apply_diag(st)), oop.apply_diag(s, aol), rtol=1e-14)
    off = oop.apply_off_diag(s, aol)
    assert list(off) == list(H.group_tables)
    for nflips, (s_conn, H_conn) in off.items():
        size = oop.get_conn_size(H_conn)
        seg_o, sc_o, Hc_o = oop.get_conn(s_conn, H_conn, size)
        seg, cidx, sc, Hc, n_nonnan = H.get_conn(st, nflips)
        assert n_nonnan == size == seg.numel()
        assert np.array_equal(to_np(seg), seg_o)
        assert np.array_equal(to_np(sc), sc_o)
        assert np.array_equal(to_np(Hc), Hc_o)
        # a larger, chunk-rounded conn_size pads exactly like jnp.nonzero(size=..., fill_value=-1)
        seg2, _, sc2, Hc2, _ = H.get_conn(st, nflips, conn_size=size + 37)
        seg_p, sc_p, Hc_p = oop.get_conn(s_conn, H_conn, size + 37)
        assert np.array_equal(to_np(seg2), seg_p) and np.array_equal(to_np(sc2), sc_p)
        assert np.array_equal(to_np(Hc2), Hc_p)


def test_ref_forward_matches_direct(qtx):
    """state.ref_forward(s_new, s_old, nflips, idx_segment, internal) == state(s_new)
    (tutorials/local_updates.ipynb:233), including the reference's padding of missing flips."""
    lattice_pair(qtx, "chain", 64)
    model, net = make_rbm(qtx, 64, 256, torch.float64, seed=10)
    state = qtx.state.Variational(model)
    s_old = osmp.rand_states(1000, 64, seed=11)
    s_new = s_old.copy()
    s_new[:, 0] *= -1
    so, sn = torch.from_numpy(s_old), torch.from_numpy(s_new)
    internal = state.init_internal(so)
    seg = torch.arange(1000)
    psi = state.ref_forward(sn, so, 1, seg, internal)
    assert np.allclose(to_np(psi.logabs), to_np(state(sn).logabs), rtol=1e-12)
    (sg, la), _ = net.ref_forward(s_new, s_old, 1, net.init_internal(s_old))
    assert np.allclose(to_np(psi.logabs), la, rtol=1e-12)
    # two flips, shuffled parents
    perm = np.random.default_rng(0).permutation(1000)
    s2 = s_old[perm].copy()
    s2[:, 5] *= -1
    s2[:, 40] *= -1
    psi2 = state.ref_forward(torch.from_numpy(s2), so, 2, torch.from_numpy(perm), internal)
    assert np.allclose(to_np(psi2.logabs), net.forward(s2)[1], rtol=1e-12)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 1e-6)])
def test_jacobian_and_obar(qtx, dtype, tol):
    lattice_pair(qtx, "square", 6)
    model, net = make_rbm(qtx, 36, 50, dtype, seed=12)
    state = qtx.state.Variational(model)
    s = osmp.rand_states(96, 36, seed=13)
    st = torch.from_numpy(s)
    O = to_np(state.jacobian(st))
    Oo = net.jacobian(s)
    assert O.shape == Oo.shape == (96, 36 * 50 + 50)
    assert np.abs(O - Oo).max() <= tol * 10
    mean = to_np(state.jacobian_colmean(st))
    assert np.abs(mean - Oo.mean(axis=0)).max() <= tol * 10
    rw = np.ones(96)
    ob_o, _ = osolver.obar(Oo, rw)
    scale = torch.sqrt(torch.ones(96, dtype=torch.float64, device="cuda") / 96)
    Ob = to_np(state.jacobian(st, col_mean=torch.from_numpy(Oo.mean(axis=0)).cuda(), row_scale=scale))
    assert np.abs(Ob - ob_o).max() <= tol * 10
    # odd leading dimension -> scalar-store path
    buf = torch.zeros((96, model.nparams + 3), dtype=torch.float64, device="cuda")
    state.jacobian(st, out=buf[:, : model.nparams])
    assert np.abs(to_np(buf[:, : model.nparams]) - Oo).max() <= tol * 10
    assert float(buf[:, model.nparams:].abs().max()) == 0.0


def test_empty_and_tiny_batches(qtx):
    lattice_pair(qtx, "chain", 8)
    model, net = make_rbm(qtx, 8, 16, torch.float32, seed=14)
    state = qtx.state.Variational(model)
    H = qtx.operator.Ising(h=1.0)
    s1 = osmp.rand_states(1, 8, seed=15)
    assert np.allclose(to_np(state(torch.from_numpy(s1)).logabs), net.forward(s1)[1], rtol=1e-5)
    E = to_np(H.Oloc(state, torch.from_numpy(s1)))
    Eo = oop.oloc(oop.to_array_op_list(oop.ising_op_list(osites.Chain(8), h=1.0)), net.forward, s1)
    assert np.allclose(E, Eo, rtol=1e-5)
    empty = torch.zeros((0, 8), dtype=torch.int8)
    assert state(empty).logabs.numel() == 0
    assert H.Oloc(state, empty).numel() == 0


def test_jacobian_odd_site_count_and_wide_rows(qtx):
    """N odd -> scalar-store path; N = 256 -> two hidden units per block iteration."""
    for kind, L, N, M in (("chain", 9, 9, 20), ("square", 16, 256, 40)):
        lattice_pair(qtx, kind, L)
        model, net = make_rbm(qtx, N, M, torch.float32, seed=21)
        state = qtx.state.Variational(model)
        s = osmp.rand_states(33, N, seed=22)
        st = torch.from_numpy(s)
        Oo = net.jacobian(s)
        assert np.abs(to_np(state.jacobian(st)) - Oo).max() <= 1e-5
        assert np.abs(to_np(state.jacobian_colmean(st)) - Oo.mean(axis=0)).max() <= 1e-5


def test_mix_sampler_matches_oracle(qtx):
    """MixSampler (quantax/sampler/metropolis.py:325-428, tutorials/samples.ipynb): a LocalFlip + SpinExchange mixture;
    with injected component choices and randoms the accept pattern and the chains equal the oracle's bit for bit,
    and the production streams reproduce the oracle's restatement."""
    lat, olat = lattice_pair(qtx, "square", 4, (8, 8))
    model, net = make_rbm(qtx, 16, 24, torch.float64, seed=51)
    state = qtx.state.Variational(model)
    a = qtx.sampler.LocalFlip(state, 24, thermal_steps=0)
    b = qtx.sampler.SpinExchange(state, 40, thermal_steps=0)
    mix = qtx.sampler.MixSampler([a, b], thermal_steps=0)
    assert mix.nsamples == 64 and mix.nflips == 2
    with pytest.raises(ValueError):
        qtx.sampler.MixSampler([a, qtx.sampler.LocalFlip(qtx.state.Variational(model), 8, thermal_steps=0)])
    spins0 = to_np(mix._spins).copy()
    assert np.array_equal(spins0, np.concatenate([to_np(a._spins), to_np(b._spins)]))
    T, ns = 37, 64
    rng = np.random.default_rng(52)
    table = osites.site_neighbor_table(olat)
    choice = rng.integers(0, 2, size=T)
    u = rng.random((T, ns)); pos = rng.integers(0, 16, size=(T, ns)); slot = rng.integers(0, 4, size=(T, ns))
    mix.inject_choice(choice)
    mix.inject(torch.from_numpy(pos), torch.from_numpy(u), torch.from_numpy(slot))
    samples = mix.sweep(T, record=True)
    ref = osmp.mix_sweep(osmp.RBMChainModel(net), spins0, choice, ["localflip", "exchange"], [None, table], [1, 1],
                         pos=pos, slot=slot, u=u, record=True)
    assert np.array_equal(to_np(mix.last_accept_log), ref["accept_log"])
    assert np.array_equal(to_np(samples.spins), ref["spins"])
    assert np.array_equal(to_np(mix.last_naccept), ref["naccept"])
    # production streams: component choice from the sampler's NumPy Philox stream, proposals from the device Philox
    spins1 = to_np(mix._spins).copy()
    choice2 = mix._draw_choice(20)
    samples2 = mix.sweep(20)
    ref2 = osmp.mix_sweep(osmp.RBMChainModel(net), spins1, choice2, ["localflip", "exchange"], [None, table], [1, 1],
                          seed=mix._seed, step0=T)
    assert np.array_equal(to_np(samples2.spins), ref2["spins"])
    assert abs(np.mean(choice2) - 40 / 64) < 0.35


@pytest.mark.parametrize("kind,L,shape,nup", [("square", 4, (4, 4), (8, 8)), ("chain", 10, (1, 10), (5, 5))])
def test_rbm_conv_matches_oracle(qtx, kind, L, shape, nup, tmp_path):
    """RBM_Conv (quantax/model/shallow_nets.py:129-190, the model of examples/RBM.ipynb): full-lattice circular
    convolution + prod cosh, evaluated through the equivalent tied dense RBM.  Forward, log-derivatives, an exchange
    sweep with injected randoms (against the oracle's FULL-forward Metropolis: the reference has no local updates for
    this model), Oloc and the SR step."""
    lat, olat = lattice_pair(qtx, kind, L, nup)
    N = shape[0] * shape[1]
    net = omodels.RBMConv.random(shape, 3, np.float64, seed=61)
    model = qtx.model.RBM_Conv(3, dtype=torch.float64, params=torch.from_numpy(net.params().copy()))
    assert model.nparams == net.nparams == 3 * N + 3
    state = qtx.state.Variational(model)
    s = osmp.rand_states(21, N, nup[0], seed=62)
    psi = state(torch.from_numpy(s))
    assert np.allclose(to_np(psi.logabs), net.forward(s)[1], rtol=1e-12, atol=1e-12)
    O = to_np(state.jacobian(torch.from_numpy(s)))
    assert np.abs(O - net.jacobian(s)).max() <= 1e-11 * np.abs(net.jacobian(s)).max()
    ns, T = 32, 30
    sampler = qtx.sampler.SpinExchange(state, ns, thermal_steps=0)
    spins0 = to_np(sampler._spins).copy()
    rng = np.random.default_rng(63)
    table = osites.site_neighbor_table(olat)
    u = rng.random((T, ns)); pos = rng.integers(0, N, size=(T, ns)); slot = rng.integers(0, table.shape[1], size=(T, ns))
    sampler.inject(torch.from_numpy(pos), torch.from_numpy(u), torch.from_numpy(slot))
    samples = sampler.sweep(T, record=True)
    ref = osmp.sweep(osmp.FullForwardChainModel(net), spins0, T, "exchange", neighbors=table, pos=pos, slot=slot, u=u,
                     record=True)
    assert np.array_equal(to_np(sampler.last_accept_log), ref["accept_log"])
    assert np.array_equal(to_np(samples.spins), ref["spins"])
    H = qtx.operator.Heisenberg(msr=(kind == "square"))
    aol = oop.to_array_op_list(oop.heisenberg_op_list(olat, msr=(kind == "square")))
    sc = ref["spins"]
    Eo = oop.oloc(aol, net.forward, sc)
    opt = qtx.optimizer.SR(state, H)
    step = to_np(opt.get_step(samples))
    assert np.abs(to_np(opt._Eloc) - Eo).max() <= 1e-10 * np.abs(Eo).max()
    xo, eo, vo = osolver.sr_step(net.jacobian(sc), Eo, np.ones(ns))
    assert abs(opt.energy - eo) <= 1e-10 * abs(eo)
    # 1e-10 unless the kept eigenvalues of this 32-row system reach below 2e-6 lambda_max (then eps * condition)
    check("RBM_Conv SR step vs oracle", np.linalg.norm(step - xo) / np.linalg.norm(xo),
          sr_step_tolerance(osolver.obar(net.jacobian(sc), np.ones(ns))[0]))
    f = tmp_path / "rbmconv.eqx"
    state.save(f)
    state2 = qtx.state.Variational(qtx.model.RBM_Conv(3, dtype=torch.float64), param_file=f)
    assert torch.equal(state2.get_params_flatten(), state.get_params_flatten())
