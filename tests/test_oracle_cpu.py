"""Pins the CPU oracle against the reference's own tables (tests/golden/ref_tables.npz, produced by
tests/golden/make_golden.py from /root/reference) and against the known answers printed in the
reference tutorials."""
import os

import numpy as np
import pytest

from oracle import models, operator as oop, sampler as osmp, sites as osites, solver as osolver

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_tables.npz"))

LATTICES = {
    "chain8": lambda: osites.Chain(8),
    "square4": lambda: osites.Square(4, Nparticles=(8, 8)),
    "square6": lambda: osites.Square(6),
    "square10": lambda: osites.Square(10),
    "square16": lambda: osites.Square(16),
    "triangular6": lambda: osites.Triangular(6),
    "triangular12": lambda: osites.Triangular(12),
}


@pytest.mark.parametrize("name", list(LATTICES))
def test_bond_tables_match_reference(name):
    lat = LATTICES[name]()
    assert np.allclose(lat.coord, GOLD[f"{name}/coord"])
    for n in (1, 2):
        assert np.array_equal(lat.get_neighbor(n), GOLD[f"{name}/nb{n}"])


def _flat(op_list):
    J = np.array([t[0] for _, ts in op_list for t in ts], dtype=np.float64)
    idx = [list(t[1:]) for _, ts in op_list for t in ts]
    return [o for o, _ in op_list], J, idx


@pytest.mark.parametrize("name", list(LATTICES))
def test_op_lists_match_reference(name):
    lat = LATTICES[name]()
    if name == "chain8":
        ops = {"ising_h1": oop.ising_op_list(lat, h=1.0), "ising_h0.5_J2": oop.ising_op_list(lat, h=0.5, J=2.0)}
    else:
        ops = {"heis": oop.heisenberg_op_list(lat), "heis_msr": oop.heisenberg_op_list(lat, msr=True),
               "j1j2_msr": oop.heisenberg_op_list(lat, J=[1, 0.5], n_neighbor=[1, 2], msr=True)}
    for oname, ol in ops.items():
        names, J, idx = _flat(ol)
        assert names == list(GOLD[f"{name}/{oname}/names"])
        assert np.array_equal(J, GOLD[f"{name}/{oname}/J"])
        for a, b in zip(idx, GOLD[f"{name}/{oname}/idx"]):
            assert a == [v for v in b if v >= 0]


def test_ed_energies_of_reference_tutorials():
    lat = osites.Square(4, Nparticles=(8, 8))
    e = oop.ed_lowest(oop.to_array_op_list(oop.heisenberg_op_list(lat, msr=True)), 16, nup=8, k=2)
    assert abs(e[0] - (-44.913932833715506)) < 1e-9  # tutorials/exact_diag.ipynb:191
    assert abs(e[1] - (-42.599539490653896)) < 1e-9
    e = oop.ed_lowest(oop.to_array_op_list(oop.heisenberg_op_list(lat, J=[1, 0.5], n_neighbor=[1, 2], msr=True)),
                      16, nup=8, k=1)
    assert abs(e[0] - (-33.831693405579394)) < 1e-9  # tutorials/J1J2.ipynb:374


def test_local_update_equals_direct_forward():
    """tutorials/local_updates.ipynb:189,233."""
    rng = np.random.default_rng(0)
    net = models.RBM.random(64, 256, np.float64, seed=1)
    s_old = osmp.rand_states(128, 64, seed=2)
    theta = net.init_internal(s_old)
    s_new = s_old.copy()
    s_new[:, 0] *= -1
    (sg, la), _ = net.ref_forward(s_new, s_old, 1, theta)
    sg2, la2 = net.forward(s_new)
    assert np.allclose(la, la2, rtol=1e-12, atol=1e-12) and np.array_equal(sg, sg2)


def test_oloc_local_updates_equals_direct():
    """tutorials/local_updates.ipynb:354: Oloc through local updates == through full forwards."""
    lat = osites.Chain(16)
    H = oop.to_array_op_list(oop.ising_op_list(lat, h=1.0))
    net = models.RBM.random(16, 32, np.float64, seed=3)
    s = osmp.rand_states(32, 16, seed=4)
    direct = oop.oloc(H, net.forward, s)
    theta = net.init_internal(s)
    diag = oop.apply_diag(s, H)
    # local-update path: every flip i, ratio from theta
    out = diag.copy()
    la = net.forward(s)[1]
    for i in range(16):
        s2 = s.copy()
        s2[:, i] *= -1
        (_, la2), _ = net.ref_forward(s2, s, 1, theta)
        out += -1.0 * np.exp(la2 - la)
    assert np.allclose(out, direct, rtol=1e-12)


def test_variational_energy_is_above_ground_state_and_unbiased():
    """<Eloc> over exact sampling of |psi|^2 equals <psi|H|psi>/<psi|psi> (4-site chain)."""
    lat = osites.Chain(4)
    H = oop.to_array_op_list(oop.ising_op_list(lat, h=0.7))
    net = models.RBM.random(4, 6, np.float64, seed=5)
    import itertools

    s = np.array(list(itertools.product([1, -1], repeat=4)), dtype=np.int8)
    psi = models.dense_value(net.forward(s))
    E = oop.oloc(H, net.forward, s)
    evar = np.sum(psi ** 2 * E) / np.sum(psi ** 2)
    e0 = oop.ed_lowest(H, 4, k=1)[0]
    assert evar >= e0 - 1e-12


def test_resconv_jacobian_finite_differences():
    rng = np.random.default_rng(1)
    for final in ("exp", "sinhp1"):
        net = models.ResConv.random((4, 4), 2, 3, 3, dtype=np.float64, seed=3, final=final, bias_std=0.1)
        s = rng.choice([-1, 1], size=(2, 16)).astype(np.int8)
        J = net.jacobian(s)
        p0 = net.params()

        def setp(p):
            o = 0
            for blk in net.blocks:
                for k in ("w1", "b1", "w2", "b2"):
                    if blk[k] is not None:
                        n = blk[k].size
                        blk[k] = p[o:o + n].reshape(blk[k].shape).copy()
                        o += n

        def logpsi():
            m, e = net.forward(s)
            return np.log(np.abs(m)) + e

        eps = 1e-6
        for k in rng.choice(p0.size, 40, replace=False):
            p = p0.copy(); p[k] += eps; setp(p); lp = logpsi()
            p[k] -= 2 * eps; setp(p); lm = logpsi()
            assert np.allclose((lp - lm) / (2 * eps), J[:, k], atol=1e-7)
        setp(p0)


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    r = osmp.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    r = osmp.philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v) for v in r] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    r = osmp.philox4x32(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0)
    assert [int(v) for v in r] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_solver_minnorm_and_lstsq_agree_with_lstsq():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((20, 50)); b = rng.standard_normal(20)
    x = osolver.minnorm_pinv_eig(A, b)
    assert np.allclose(x, np.linalg.lstsq(A, b, rcond=None)[0], atol=1e-9)
    A = rng.standard_normal((50, 20)); b = rng.standard_normal(50)
    x = osolver.lstsq_pinv_eig(A, b)
    assert np.allclose(x, np.linalg.lstsq(A, b, rcond=None)[0], atol=1e-9)


def test_oracle_vmc_converges_to_ed_quick_start():
    """README quick start (config A): Ising chain L=8, h=1, RBM_Dense(16), LocalFlip, SR."""
    lat = osites.Chain(8)
    H = oop.to_array_op_list(oop.ising_op_list(lat, h=1.0))
    net = models.RBM.random(8, 16, np.float32, seed=1, scale=0.3)
    cm = osmp.RBMChainModel(net)
    ns = 256
    spins = osmp.sweep(cm, osmp.rand_states(ns, 8, seed=2), 160, "localflip", seed=7, step0=0)["spins"]
    step0, hist = 160, []
    for it in range(60):
        out = osmp.sweep(cm, spins, 16, "localflip", seed=7, step0=step0)
        step0 += 16
        spins = out["spins"]
        E = oop.oloc(H, net.forward, spins, out["psi"])
        x, e, v = osolver.sr_step(net.jacobian(spins), E, np.ones(ns))
        p = osolver.update_params(net.params(), x * 0.03)
        net.W = p[: net.W.size].reshape(net.W.shape).copy()
        net.b = p[net.W.size:].copy()
        hist.append(e)
    e0 = oop.ed_lowest(H, 8, k=1)[0]
    assert abs(np.mean(hist[-10:]) - e0) < 0.05 * abs(e0) and np.mean(hist[-10:]) > e0 - 0.05


def test_complex_resconv_jacobian_is_the_log_derivative():
    """Complex-output ResConv with a phase layer (conv_nets.py:165-170, nn/sign.py:62-75): the two-pass
    Jacobian of variational.py:461-487 equals the finite-difference derivative of log psi (pins the oracle's
    complex path, for which the reference holds no stored vectors)."""
    from oracle import models as om

    for final in ("exp", "sinhp1"):
        net = om.ResConv.random((4, 4), 2, 4, 3, np.float64, seed=1, final=final, bias_std=0.1, out_complex=True,
                                phase_kernel=np.linspace(0, 1, 16))
        rng = np.random.default_rng(0)
        s = (2 * rng.integers(0, 2, (3, 16)) - 1).astype(np.int8)
        J = net.jacobian(s)
        p0 = net.params()

        def setp(p):
            o = 0
            for blk in net.blocks:
                for k in ("w1", "b1", "w2", "b2"):
                    if blk[k] is not None:
                        n = blk[k].size
                        blk[k] = p[o:o + n].reshape(blk[k].shape)
                        o += n

        for k in rng.integers(0, p0.size, 8):
            h = 1e-6
            pp = p0.copy(); pp[k] += h; setp(pp); s1, e1 = net.forward(s)
            pm = p0.copy(); pm[k] -= h; setp(pm); s2, e2 = net.forward(s)
            d = ((np.log(np.abs(s1)) + e1) - (np.log(np.abs(s2)) + e2)) / (2 * h) + 1j * np.angle(s1 / s2) / (2 * h)
            assert np.abs(d - J[:, k]).max() < 1e-8
        setp(p0)


def test_sr_step_real_to_complex_stacking():
    """sr.py:99-104: with real parameters and complex Obar the solver sees [Re; Im]; the solution solves the
    complex least-squares problem restricted to real steps."""
    from oracle import solver as osolver

    rng = np.random.default_rng(3)
    O = rng.standard_normal((12, 40)) + 1j * rng.standard_normal((12, 40))
    E = rng.standard_normal(12) + 1j * rng.standard_normal(12)
    x, e, v = osolver.sr_step(O, E, np.ones(12), real_to_complex=True)
    ob, _ = osolver.obar(O, np.ones(12))
    eb, _, _ = osolver.ebar(E, np.ones(12))
    assert np.isrealobj(x)
    assert np.allclose(ob @ x, eb, atol=1e-9)  # 24 real equations, 40 unknowns: exact min-norm solution


def test_rbm_conv_oracle_is_the_same_convolution_as_resconv_layers():
    """RBM_Conv (shallow_nets.py:129-173) uses eqx.nn.Conv with the lattice-sized kernel, SAME + CIRCULAR padding:
    the oracle's direct definition equals the generic circular cross-correlation restated for ResConv, and its
    log-derivative equals finite differences of log psi."""
    from oracle import models as om

    net = om.RBMConv.random((4, 6), 3, np.float64, seed=1)
    rng = np.random.default_rng(0)
    s = (2 * rng.integers(0, 2, (5, 24)) - 1).astype(np.int8)
    x = s.astype(np.float64).reshape(-1, 1, 4, 6)
    assert np.abs(net.theta(s) - om.conv_circ(x, net.K, net.b)).max() < 1e-14
    J, p0 = net.jacobian(s), net.params()
    for k in rng.integers(0, p0.size, 8):
        def logpsi(p):
            return om.RBMConv(p[:net.K.size].reshape(net.K.shape), p[net.K.size:], net.shape).forward(s)[1]
        pp, pm = p0.copy(), p0.copy()
        pp[k] += 1e-6
        pm[k] -= 1e-6
        assert np.abs((logpsi(pp) - logpsi(pm)) / 2e-6 - J[:, k]).max() < 1e-8


def test_time_evol_chunked_accumulation_equals_direct():
    """time_evol.py:55-115: the chunked S / F accumulation (un-centred sums corrected by the means) equals the direct
    Obar^+ Obar, Obar^+ Ebar; and S, F are the stacked-real products used by the B200 path."""
    from oracle import solver as osolver

    rng = np.random.default_rng(5)
    O = rng.standard_normal((90, 12)) + 1j * rng.standard_normal((90, 12))
    E = rng.standard_normal(90) + 1j * rng.standard_normal(90)
    x1, e1, v1, S1, F1 = osolver.time_evol_step(O, E)
    x2, e2, v2, S2, F2 = osolver.time_evol_step(O, E, max_parallel=32)
    assert np.allclose(S1, S2, atol=1e-12) and np.allclose(F1, F2, atol=1e-12) and np.allclose(x1, x2, atol=1e-9)
    assert abs(e1 - e2) < 1e-12 and abs(v1 - v2) < 1e-12
    ob, _ = osolver.obar(O, np.ones(90))
    eb, _, _ = osolver.ebar(E, np.ones(90))
    A = np.concatenate([ob.real, ob.imag], axis=0)
    b = np.concatenate([-eb.imag, eb.real])
    assert np.allclose(A.T @ A, S1, atol=1e-12) and np.allclose(A.T @ b, F1, atol=1e-12)


def test_mix_sweep_reduces_to_plain_sweep_for_one_component():
    from oracle import models as om, sampler as osmp, sites as osites

    lat = osites.Square(4, Nparticles=(8, 8))
    net = om.RBM.random(16, 8, np.float64, seed=2, scale=0.5)
    spins = osmp.rand_states(12, 16, 8, seed=3)
    table = osites.site_neighbor_table(lat)
    a = osmp.sweep(osmp.RBMChainModel(net), spins, 9, "exchange", neighbors=table, seed=11, step0=4)
    b = osmp.mix_sweep(osmp.RBMChainModel(net), spins, [0] * 9, ["exchange"], [table], [1], seed=11, step0=4)
    assert np.array_equal(a["spins"], b["spins"]) and np.array_equal(a["naccept"], b["naccept"])


# ---- solver variants (quantax/optimizer/solver.py:24-90,114-125,204-259,297-302) ---------------------------
def _lsq_problem(ns, npar, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((ns, npar)) * np.exp(0.5 * rng.standard_normal((1, npar)))
    A -= A.mean(axis=0, keepdims=True)
    return A / np.sqrt(ns), rng.standard_normal(ns) / np.sqrt(ns)


@pytest.mark.parametrize("ns,npar", [(24, 60), (60, 24)])
def test_oracle_shift_solvers_solve_the_shifted_normal_equations(ns, npar):
    A, b = _lsq_problem(ns, npar)
    x = osolver.auto_shift_eig(A, b, rshift=1e-3, ashift=1e-4)
    # both branches are the ridge solution with lambda = rshift tr + ashift; tr(A A^T) == tr(A^T A)
    lam = 1e-3 * np.trace(A @ A.T) + 1e-4
    ref = np.linalg.solve(A.T @ A + lam * np.eye(npar), A.T @ b)
    assert np.allclose(x, ref, rtol=1e-9, atol=1e-12)
    assert np.allclose(osolver.minnorm_shift_eig(A, b, 1e-3, 1e-4), osolver.lstsq_shift_eig(A, b, 1e-3, 1e-4), rtol=1e-9)


def test_oracle_cg_converges_to_the_direct_solution():
    A, b = _lsq_problem(80, 30, seed=1)
    x, k = osolver.lstsq_shift_cg(A, b, diag_shift=0.01, rtol=1e-10, return_iterations=True)
    S = A.T @ A
    ref = np.linalg.solve(S + 0.01 * np.diag(np.diag(S)), A.T @ b)
    assert 0 < k <= 300
    assert np.allclose(x, ref, rtol=1e-7, atol=1e-10)
    # loose tolerance stops early; maxiter caps the iteration count
    _, k2 = osolver.lstsq_shift_cg(A, b, rtol=1e-2, return_iterations=True)
    assert k2 < k
    assert osolver.lstsq_shift_cg(A, b, rtol=1e-12, maxiter=3, return_iterations=True)[1] == 3


def test_oracle_block_solver_and_sgd():
    A, b = _lsq_problem(20, 45, seed=2)
    sizes = [10, 0, 20, 15]
    x = osolver.block_pinv_eig(A, b, sizes)
    assert x.shape == (45,)
    assert np.allclose(x[:10], osolver.auto_pinv_eig(A[:, :10], b / 3))
    assert np.allclose(x[30:], osolver.auto_pinv_eig(A[:, 30:], b / 3))
    assert np.allclose(osolver.block_pinv_eig(A, b, [45]), osolver.auto_pinv_eig(A, b))
    assert np.allclose(osolver.sgd_solver(A, b), A.T @ b / 20)


def test_oracle_snr_damping_limits():
    A, b = _lsq_problem(16, 50, seed=3)
    x0 = osolver.minnorm_pinv_eig(A, b)
    assert np.allclose(osolver.minnorm_pinv_eig(A, b, tol_snr=1e-7), x0)  # below the reference's 1e-6 switch
    xs = osolver.minnorm_pinv_eig(A, b, tol_snr=1.0)
    assert np.linalg.norm(xs) < np.linalg.norm(x0)  # every eigen-direction is damped by a factor in (0, 1]
    # the damping factor per direction, recomputed independently
    vals, U = np.linalg.eigh(A @ A.T)
    r = U * b[:, None]
    mean = r.mean(axis=0)
    snr = np.abs(mean) / np.sqrt(((r - mean) ** 2).mean(axis=0) / 16)
    rho = r.sum(axis=0) / (1 + (1.0 / snr) ** 6)
    assert np.allclose(xs, A.T @ (U @ (osolver.eigs_inv(vals) * rho)), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("dtype,final", [(np.float32, "sinhp1"), (np.float64, "exp")])
def test_torch_cpu_resconv_equals_the_numpy_oracle(dtype, final):
    """oracle/resconv_torch.py (the timed CPU leg of bench.py and the fast oracle of the BASELINE-shape GPU tests)
    against the einsum oracle."""
    from oracle.resconv_torch import TorchResConv

    net = models.ResConv.random((6, 6), 3, 10, 3, dtype, seed=3, final=final, bias_std=0.2)
    s = osmp.rand_states(9, 36, 18, seed=4)
    a, b = net.forward(s), TorchResConv(net).forward(s)
    tol = 2e-6 if dtype == np.float32 else 1e-13
    assert np.array_equal(np.sign(a[0]), np.sign(b[0]))
    assert np.abs(np.log(np.abs(a[0])) + a[1] - np.log(np.abs(b[0])) - b[1]).max() < tol
