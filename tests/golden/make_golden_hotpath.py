"""Golden vectors of the hot path produced by the REFERENCE'S OWN code under a NumPy stand-in for jax.

Run in the build container only (needs /root/reference; it is never read by the tests):
    python tests/golden/make_golden_hotpath.py

jax is not installable here, so ``tests/golden/minijax.py`` provides eager NumPy restatements of the jax
functions the reference's hot-path modules call (see its header for what that does and does not pin).  The
reference modules themselves are imported UNMODIFIED from /root/reference:

  quantax/operator/operator.py      _apply_diag, _apply_off_diag, _get_conn_size, _get_conn, _get_Olocx,
                                    Operator.jax_op_list (lines 30-184, 221-238)
  quantax/operator/common_operators.py, site_operator.py   Heisenberg / Ising op lists
  quantax/optimizer/solver.py       _get_eigs_inv, _sum_without_noise, pinvh_solve, minnorm/lstsq/auto/minsr_pinv_eig,
                                    minnorm/lstsq/auto_shift_eig, sgd_solver (lines 50-201, 262-302)
  quantax/symmetry/symmetry.py      _get_perm (group closure, characters; lines 11-57)
  quantax/sampler/common_samplers.py  _get_site_neighbors (lines 36-55)
  quantax/nn/sign.py                neel120_phase, marshall_sign, stripe_sign kernels
  quantax/utils/array.py            array_extend, local_to_replicate, to_distribute_array

Inputs are seeded NumPy arrays stored next to the outputs in ``tests/golden/ref_hotpath.npz``.
"""
import enum
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import minijax  # noqa: E402


def install_reference():
    jax, jnp = minijax.install()
    jax.lax.with_sharding_constraint = lambda x, s: x
    pkg = types.ModuleType("quantax")
    pkg.__path__ = [os.path.join(REF, "quantax")]
    sys.modules["quantax"] = pkg

    gd = types.ModuleType("quantax.global_defs")

    class PARTICLE_TYPE(enum.Enum):
        spin = 0
        spinful_fermion = 1
        spinless_fermion = 2

    state = {"dtype": np.float64}
    gd.PARTICLE_TYPE = PARTICLE_TYPE
    gd.get_default_dtype = lambda: state["dtype"]
    gd.get_real_dtype = lambda: np.float64
    gd.is_default_cpl = lambda: np.issubdtype(state["dtype"], np.complexfloating)
    gd.set_default_dtype = lambda dt: state.__setitem__("dtype", dt)
    gd.get_subkeys = lambda num=None: None  # PRNG keys are never consumed: random draws are injected
    sys.modules["quantax.global_defs"] = gd
    pkg.global_defs = gd
    sites = importlib.import_module("quantax.sites")  # pure NumPy reference code
    gd.get_sites = lambda: sites.Sites._SITES
    gd.get_lattice = lambda: sites.Sites._SITES

    # quantax.utils: the real array helpers, the rest inert
    utils = minijax._AnyModule("quantax.utils")
    utils.__path__ = [os.path.join(REF, "quantax", "utils")]
    sys.modules["quantax.utils"] = utils
    importlib.import_module("quantax.utils.sharding")
    arr = importlib.import_module("quantax.utils.array")
    for name in ("local_to_replicate", "to_distribute_array", "to_replicate_array", "to_replicate_numpy", "array_extend",
                 "array_set"):
        setattr(utils, name, getattr(arr, name))
    tree = importlib.import_module("quantax.utils.tree")        # filter_tree_map
    big = importlib.import_module("quantax.utils.big_array")    # LogArray / ScaleArray / where (plain dataclasses)
    utils.filter_tree_map = tree.filter_tree_map
    utils.LogArray, utils.ScaleArray, utils.PsiArray, utils.where = big.LogArray, big.ScaleArray, big.PsiArray, big.where
    for sub in ("state", "sampler", "nn"):
        sys.modules[f"quantax.{sub}"] = minijax._AnyModule(f"quantax.{sub}")
    symm_pkg = minijax._AnyModule("quantax.symmetry")
    symm_pkg.__path__ = [os.path.join(REF, "quantax", "symmetry")]
    sys.modules["quantax.symmetry"] = symm_pkg
    operator = importlib.import_module("quantax.operator")
    opmod = importlib.import_module("quantax.operator.operator")
    symmod = importlib.import_module("quantax.symmetry.symmetry")
    opt_pkg = minijax._AnyModule("quantax.optimizer")
    opt_pkg.__path__ = [os.path.join(REF, "quantax", "optimizer")]
    sys.modules["quantax.optimizer"] = opt_pkg
    solver = importlib.import_module("quantax.optimizer.solver")
    smp_pkg = minijax._AnyModule("quantax.sampler")
    smp_pkg.__path__ = [os.path.join(REF, "quantax", "sampler")]
    sys.modules["quantax.sampler"] = smp_pkg
    for sub in ("sampler", "samples", "metropolis"):
        sys.modules[f"quantax.sampler.{sub}"] = minijax._AnyModule(f"quantax.sampler.{sub}")
    csamp = importlib.import_module("quantax.sampler.common_samplers")
    nn_pkg = minijax._AnyModule("quantax.nn")
    nn_pkg.__path__ = [os.path.join(REF, "quantax", "nn")]
    sys.modules["quantax.nn"] = nn_pkg
    sys.modules["quantax.nn.modules"] = minijax._AnyModule("quantax.nn.modules")
    sign = importlib.import_module("quantax.nn.sign")
    act = importlib.import_module("quantax.nn.activation")
    # the real Samples / Metropolis classes (only _update is called, unbound, with a stand-in self)
    for sub in ("sampler", "samples", "metropolis"):
        del sys.modules[f"quantax.sampler.{sub}"]
    samples = importlib.import_module("quantax.sampler.samples")
    metro = importlib.import_module("quantax.sampler.metropolis")

    class VS_TYPE(enum.Enum):  # state/variational.py:40-62 (the module itself needs equinox)
        real_or_holomorphic = 0
        non_holomorphic = 1
        real_to_complex = 2

    sys.modules["quantax.state"].VS_TYPE = VS_TYPE
    srmod = importlib.import_module("quantax.optimizer.sr")
    temod = importlib.import_module("quantax.optimizer.time_evol")
    shallow = None
    try:
        model_pkg = minijax._AnyModule("quantax.model")
        model_pkg.__path__ = [os.path.join(REF, "quantax", "model")]
        sys.modules["quantax.model"] = model_pkg
        nn_pkg.prod_by_log = act.prod_by_log
        nn_pkg.Sequential, nn_pkg.RefModel = type("Sequential", (), {}), type("RefModel", (), {})  # base classes only
        shallow = importlib.import_module("quantax.model.shallow_nets")
    except Exception as exc:  # class machinery of equinox: the local-update vectors are then skipped
        print("shallow_nets not importable under the stand-in:", exc)
    # ResConv: the reference's own model / layer / symmetry classes on the equinox stand-ins of minijax
    convnets = None
    try:
        for name in ("translation", "common_symmetries"):
            m = importlib.import_module(f"quantax.symmetry.{name}")
            for attr in ("Translation", "TransND", "Identity", "Z2Inversion", "SpinInverse", "LinearTransform", "Flip",
                         "Rotation", "C4v", "D6"):
                if hasattr(m, attr):
                    setattr(symm_pkg, attr, getattr(m, attr))
        symm_pkg.Symmetry = symmod.Symmetry
        del sys.modules["quantax.nn.modules"]
        modules = importlib.import_module("quantax.nn.modules")
        nn_pkg.Sequential, nn_pkg.RefModel, nn_pkg.RawInputLayer = modules.Sequential, modules.RefModel, modules.RawInputLayer
        nnconv = importlib.import_module("quantax.nn.conv")
        for attr in ("ReshapeConv", "ConvSymmetrize", "Reshape_TriangularB", "ReshapeTo_TriangularB", "Gconv"):
            setattr(nn_pkg, attr, getattr(nnconv, attr))
        nn_pkg.apply_he_normal = lambda key, conv: conv  # weights are injected after construction
        nn_pkg.exp_by_scale, nn_pkg.sinhp1_by_scale, nn_pkg.pair_cpl = act.exp_by_scale, act.sinhp1_by_scale, act.pair_cpl
        convnets = importlib.import_module("quantax.model.conv_nets")
    except Exception as exc:
        import traceback

        traceback.print_exc()
        print("conv_nets not importable under the stand-in:", exc)
    return dict(gd=gd, sites=sites, operator=operator, opmod=opmod, symmod=symmod, solver=solver, csamp=csamp, sign=sign,
                convnets=convnets, symm_pkg=symm_pkg, temod=temod,
                act=act, big=big, samples=samples, metro=metro, jax=jax, srmod=srmod, VS_TYPE=VS_TYPE, shallow=shallow)


def rand_spins(rng, ns, N, nup=None):
    if nup is None:
        return (2 * rng.integers(0, 2, size=(ns, N)) - 1).astype(np.int8)
    s = -np.ones((ns, N), dtype=np.int8)
    for r in range(ns):
        s[r, rng.permutation(N)[:nup]] = 1
    return s


def gen_operator(ref, out):
    sites, operator, opmod = ref["sites"], ref["operator"], ref["opmod"]
    rng = np.random.default_rng(20261017)
    cases = {
        "chain8_ising": (lambda: sites.Chain(8), lambda: operator.Ising(h=1.0), None, 12),
        "chain8_ising_h0.5_J2": (lambda: sites.Chain(8), lambda: operator.Ising(h=0.5, J=2.0), None, 12),
        "square4_heis_msr": (lambda: sites.Square(4, Nparticles=(8, 8)), lambda: operator.Heisenberg(msr=True), 8, 16),
        "square4_j1j2_msr": (lambda: sites.Square(4, Nparticles=(8, 8)),
                             lambda: operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True), 8, 16),
        "square6_j1j2_msr": (lambda: sites.Square(6, Nparticles=(18, 18)),
                             lambda: operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True), 18, 8),
        "triangular6_heis": (lambda: sites.Triangular(6, Nparticles=(18, 18)), lambda: operator.Heisenberg(), 18, 8),
        "square4_heis_unconstrained": (lambda: sites.Square(4), lambda: operator.Heisenberg(), None, 8),
    }
    for name, (mk_lat, mk_op, nup, ns) in cases.items():
        sites.Sites._SITES = None
        lat = mk_lat()
        H = mk_op()
        s = rand_spins(rng, ns, lat.Nsites, nup)
        sj = minijax.wrap(s.copy())
        jl = H.jax_op_list
        out[f"op/{name}/spins"] = s
        out[f"op/{name}/diag"] = np.asarray(opmod._apply_diag(sj, jl), dtype=np.float64)
        off = opmod._apply_off_diag(sj, jl)
        out[f"op/{name}/nflips"] = np.array(sorted(off.keys()), dtype=np.int64)
        for nflips, (s_conn, H_conn) in off.items():
            out[f"op/{name}/{nflips}/H_raw"] = np.asarray(H_conn, dtype=np.float64)
            for chunk in (None, 16):
                size = int(opmod._get_conn_size(H_conn, chunk))
                seg, sc, Hc = opmod._get_conn(s_conn, H_conn, size)
                tag = "none" if chunk is None else str(chunk)
                out[f"op/{name}/{nflips}/chunk_{tag}/segment"] = np.asarray(seg, dtype=np.int64)
                out[f"op/{name}/{nflips}/chunk_{tag}/s_conn"] = np.asarray(sc, dtype=np.int8)
                out[f"op/{name}/{nflips}/chunk_{tag}/H"] = np.asarray(Hc, dtype=np.float64)
            # _get_Olocx with a synthetic amplitude table psi(s) = exp(sum_i a_i s_i) (dense arrays, not PsiArray)
            a = rng.standard_normal(lat.Nsites) * 0.3
            psi = np.exp(s @ a)
            size = int(opmod._get_conn_size(H_conn, None))
            seg, sc, Hc = opmod._get_conn(s_conn, H_conn, size)
            psi_conn = np.exp(np.asarray(sc, dtype=np.float64) @ a)
            out[f"op/{name}/{nflips}/amp_a"] = a
            out[f"op/{name}/{nflips}/Olocx"] = np.asarray(
                opmod._get_Olocx(minijax.wrap(psi), seg, minijax.wrap(psi_conn), Hc), dtype=np.float64)


def gen_solver(ref, out):
    solver = ref["solver"]
    rng = np.random.default_rng(7)
    vals = np.concatenate([[0.0, 1e-30, -1e-14, 3e-13], np.exp(rng.uniform(-30, 2, size=20)), -np.exp(rng.uniform(-20, 0, 4))])
    out["solver/eigs_inv/vals"] = vals
    for tag, rtol, atol in (("default", None, 0.0), ("r1e-8_a1e-10", 1e-8, 1e-10)):
        out[f"solver/eigs_inv/{tag}"] = np.asarray(solver._get_eigs_inv(minijax.wrap(vals.copy()), rtol, atol))
    inputs = rng.standard_normal((17, 9)) * np.exp(rng.standard_normal((1, 9)))
    out["solver/snr/inputs"] = inputs
    for tol in (0.0, 1e-7, 0.5, 3.0):
        out[f"solver/snr/tol_{tol}"] = np.asarray(solver._sum_without_noise(minijax.wrap(inputs.copy()), tol))
    for tag, (ns, npar) in {"minnorm": (12, 40), "lstsq": (40, 12)}.items():
        A = rng.standard_normal((ns, npar)) * np.exp(0.5 * rng.standard_normal((1, npar)))
        A -= A.mean(axis=0, keepdims=True)
        A /= np.sqrt(ns)
        b = rng.standard_normal(ns) / np.sqrt(ns)
        out[f"solver/{tag}/A"], out[f"solver/{tag}/b"] = A, b
        Aj, bj = minijax.wrap(A.copy()), minijax.wrap(b.copy())
        for tol_snr in (0.0, 1.0):
            out[f"solver/{tag}/auto_pinv_eig_snr{tol_snr}"] = np.asarray(
                solver.auto_pinv_eig(rtol=1e-10, tol_snr=tol_snr)(Aj, bj))
        out[f"solver/{tag}/auto_pinv_eig_default"] = np.asarray(solver.auto_pinv_eig()(Aj, bj))
        out[f"solver/{tag}/auto_shift_eig_default"] = np.asarray(solver.auto_shift_eig()(Aj, bj))
        out[f"solver/{tag}/auto_shift_eig_r1e-3_a0"] = np.asarray(solver.auto_shift_eig(1e-3, 0.0)(Aj, bj))
        out[f"solver/{tag}/sgd"] = np.asarray(solver.sgd_solver()(Aj, bj))
        T = A @ A.T
        out[f"solver/{tag}/minsr_pinv_eig_T"] = np.asarray(solver.minsr_pinv_eig(rtol=1e-10)(minijax.wrap(T.copy()), bj))
        out[f"solver/{tag}/pinvh_T"] = np.asarray(solver.pinvh_solve(rtol=1e-10)(minijax.wrap(T.copy()), bj))


def gen_symmetry(ref, out):
    sites, symmod, gd = ref["sites"], ref["symmod"], ref["gd"]
    gens = np.load(os.path.join(HERE, "ref_symm_generators.npz"))  # generators from the reference's NumPy code
    for lat_name, mk in {"square4": lambda: sites.Square(4), "square6": lambda: sites.Square(6),
                         "triangular6": lambda: sites.Triangular(6), "chain8": lambda: sites.Chain(8)}.items():
        sites.Sites._SITES = None
        mk()
        combos = {"trans": ["trans"], "flip0": ["flip0"]}
        if f"{lat_name}/rot" in gens:
            combos["rot"] = ["rot"]
            combos["rot_flip0"] = ["rot", "flip0"]
        for cname, parts in combos.items():
            g = np.concatenate([np.atleast_2d(gens[f"{lat_name}/{p}"]) for p in parts], axis=0)
            for sec in (0, 1):
                sector = [sec if i == 0 else 0 for i in range(g.shape[0])]
                try:
                    perm, character, perm_sign = symmod._get_perm(g, sector, np.ones(g.shape[0], dtype=np.int8))
                except ValueError:
                    continue  # complex characters with the real default dtype
                key = f"symm/{lat_name}/{cname}/sec{sec}"
                out[f"{key}/generator"] = g.astype(np.int64)
                out[f"{key}/perm"] = np.asarray(perm, dtype=np.int64)
                out[f"{key}/character"] = np.asarray(character, dtype=np.float64)


def gen_sampler_tables(ref, out):
    sites, csamp = ref["sites"], ref["csamp"]
    for lat_name, mk in {"chain8": lambda: sites.Chain(8), "square4": lambda: sites.Square(4),
                         "square10": lambda: sites.Square(10), "triangular6": lambda: sites.Triangular(6)}.items():
        for nn_arg, tag in ((1, "n1"), ([1, 2], "n12")):
            sites.Sites._SITES = None
            mk()
            out[f"nbr/{lat_name}/{tag}"] = np.asarray(csamp._get_site_neighbors(nn_arg), dtype=np.int64)


def gen_more_lattices(ref, out):
    """sites/common_lattices.py: TriangularB and Cube geometry, neighbour shells and Heisenberg op lists from the
    reference's own NumPy code (the other lattices are in ref_tables.npz)."""
    sites, operator = ref["sites"], ref["operator"]
    for name, mk in {"triangularB2": lambda: sites.TriangularB(2), "cube3": lambda: sites.Cube(3)}.items():
        sites.Sites._SITES = None
        lat = mk()
        out[f"lat/{name}/coord"] = lat.coord
        for n in (1, 2):
            out[f"lat/{name}/nb{n}"] = np.asarray(lat.get_neighbor(n), dtype=np.int64)
        H = operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2])
        out[f"lat/{name}/j1j2/names"] = np.array([o for o, _ in H.op_list])
        out[f"lat/{name}/j1j2/J"] = np.array([t[0] for _, ts in H.op_list for t in ts], dtype=np.float64)
        out[f"lat/{name}/j1j2/idx"] = np.array([list(t[1:]) for _, ts in H.op_list for t in ts], dtype=np.int64)


ALGEBRA_EXPRESSIONS = {
    # name -> expression over the module `O` (the reference's or the product's `operator` package)
    "pm": "O.sigma_p(0) @ O.sigma_m(1)",
    "pm_H": "(O.sigma_p(0) @ O.sigma_m(1)).H",
    "mixed": "2 * (O.sigma_p(0) @ O.sigma_m(1)) + O.sigma_z(2) @ O.sigma_z(3) - O.sigma_p(0) @ O.sigma_m(1)",
    "scaled": "(O.S_x(1) @ O.S_x(2) + 0.5 * O.S_z(3)) / 4",
    "neg": "-(O.sigma_x(1, 2) + O.sigma_z(0, 3))",
    "coord_wrap": "O.sigma_z(5, 1) @ O.sigma_z(-1, 0)",
    "triple": "O.S_p(0) @ O.S_m(5) @ O.S_z(10)",
    "sum_same_opstr": "O.sigma_z(0) @ O.sigma_z(1) + O.sigma_z(1) @ O.sigma_z(2) + O.sigma_z(2) @ O.sigma_z(3)",
    "heis_plus_field": "O.Heisenberg(msr=True) + 0.3 * sum((O.sigma_z(i) for i in range(16)), start=0 * O.sigma_z(0))",
    "rsub": "O.Ising(h=1.0) - O.Heisenberg()",
}


def _op_list_arrays(op_list):
    names = np.array([o for o, _ in op_list])
    width = np.array([len(ts) for _, ts in op_list], dtype=np.int64)
    J = np.array([complex(t[0]) for _, ts in op_list for t in ts])
    idx = np.array([list(t[1:]) + [-1] * (4 - len(t[1:])) for _, ts in op_list for t in ts], dtype=np.int64)
    return names, width, J, idx


def gen_operator_algebra(ref, out):
    """operator/operator.py:325-470 (`@ + - * / .H` of Operator) and site_operator.py: op lists of compound
    expressions from the reference's own pure-Python algebra on a 4x4 square lattice."""
    sites, O = ref["sites"], ref["operator"]
    sites.Sites._SITES = None
    sites.Square(4)
    for name, expr in ALGEBRA_EXPRESSIONS.items():
        names, width, J, idx = _op_list_arrays(eval(expr, {"O": O, "sum": sum, "range": range}).op_list)
        out[f"algebra/{name}/names"], out[f"algebra/{name}/width"] = names, width
        out[f"algebra/{name}/J"], out[f"algebra/{name}/idx"] = J, idx


def gen_sign(ref, out):
    sites, sign = ref["sites"], ref["sign"]
    rng = np.random.default_rng(5)
    sites.Sites._SITES = None
    lat = sites.Triangular(6)
    s = rand_spins(rng, 6, lat.Nsites, 18)
    out["sign/triangular6/spins"] = s
    out["sign/triangular6/neel120_phase"] = np.stack([np.asarray(sign.neel120_phase(minijax.wrap(r.copy()))) for r in s])
    sites.Sites._SITES = None
    lat = sites.TriangularB(2)
    s = rand_spins(rng, 6, lat.Nsites, 6)
    out["sign/triangularB2/spins"] = s
    out["sign/triangularB2/neel120_phase"] = np.stack([np.asarray(sign.neel120_phase(minijax.wrap(r.copy()))) for r in s])


def _parts(x):
    """(first, second) leaves of a LogArray / ScaleArray as plain arrays."""
    a, b = x.tree_flatten()[0]
    return np.asarray(a), np.asarray(b)


def gen_containers(ref, out):
    """utils/big_array.py: arithmetic of the psi containers (the Oloc ratio, the Metropolis rate, symmetrize)."""
    big = ref["big"]
    rng = np.random.default_rng(11)
    n = 13
    W = minijax.wrap
    sa, sb = rng.choice([-1.0, 1.0], n), rng.choice([-1.0, 1.0], n)
    la, lb = rng.standard_normal(n) * 300, rng.standard_normal(n) * 300
    lb[:4] = la[:4] + rng.standard_normal(4)  # comparable magnitudes: additions that do not degenerate
    out["cont/log/a_sign"], out["cont/log/a_logabs"], out["cont/log/b_sign"], out["cont/log/b_logabs"] = sa, la, sb, lb
    A, B = big.LogArray(W(sa.copy()), W(la.copy())), big.LogArray(W(sb.copy()), W(lb.copy()))
    dense = rng.standard_normal(n)
    out["cont/dense"] = dense
    ops = {"div": A / B, "mul": A * B, "add": A + B, "sub": A - B, "neg": -A, "abs": abs(A), "pow2": A ** 2, "pow1.3": abs(A) ** 1.3,
           "mul_dense": A * W(dense.copy()), "rdiv": 2.0 / A,
           "sum": A.sum(), "mean": A.mean(), "prod": A.prod(), "from_value": big.LogArray.from_value(W(dense.copy())),
           "where": big.where(W(dense > 0), A, B)}
    for k, v in ops.items():
        out[f"cont/log/{k}/0"], out[f"cont/log/{k}/1"] = _parts(v)
    # per sample: sum over the image axis with characters, as symmetrize does (symmetry.py:386-392)
    s2, l2 = rng.choice([-1.0, 1.0], (5, 4)), rng.standard_normal((5, 4)) * 40
    chi = np.array([1.0, -1.0, 1.0, -1.0])
    out["cont/log/img_sign"], out["cont/log/img_logabs"], out["cont/chi"] = s2, l2, chi
    char = chi * chi[0] / chi.size  # symmetry.py:391
    rows = [_parts((big.LogArray(W(s2[r].copy()), W(l2[r].copy())) * W(char.copy())).sum()) for r in range(5)]
    out["cont/log/proj/0"], out["cont/log/proj/1"] = np.array([r[0] for r in rows]), np.array([r[1] for r in rows])

    ma, mb = rng.standard_normal(n), rng.standard_normal(n)
    ea, eb = rng.standard_normal(n) * 200, rng.standard_normal(n) * 200
    eb[:4] = ea[:4] + rng.standard_normal(4)
    out["cont/scale/a_sig"], out["cont/scale/a_exp"], out["cont/scale/b_sig"], out["cont/scale/b_exp"] = ma, ea, mb, eb
    A, B = big.ScaleArray(W(ma.copy()), W(ea.copy())), big.ScaleArray(W(mb.copy()), W(eb.copy()))
    ops = {"div": A / B, "mul": A * B, "add": A + B, "sub": A - B, "neg": -A, "abs": abs(A), "pow2": A ** 2,
           "pow1.3": abs(A) ** 1.3, "mul_dense": A * W(dense.copy()), "rdiv": 2.0 / A, "sum": A.sum(), "mean": A.mean(),
           "prod": A.prod(), "normalize": A.normalize(), "from_value": big.ScaleArray.from_value(W(dense.copy())),
           "where": big.where(W(dense > 0), A, B), "to_log": big.LogArray.from_value(A)}
    for k, v in ops.items():
        out[f"cont/scale/{k}/0"], out[f"cont/scale/{k}/1"] = _parts(v)
    m2, e2 = rng.standard_normal((5, 4)), rng.standard_normal((5, 4)) * 40
    out["cont/scale/img_sig"], out["cont/scale/img_exp"] = m2, e2
    rows = [_parts((big.ScaleArray(W(m2[r].copy()), W(e2[r].copy())) * W(char.copy())).sum()) for r in range(5)]
    out["cont/scale/proj/0"], out["cont/scale/proj/1"] = np.array([r[0] for r in rows]), np.array([r[1] for r in rows])
    # complex significands (config D)
    mc = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    out["cont/scale/c_sig"] = mc
    C = big.ScaleArray(W(mc.copy()), W(ea.copy()))
    for k, v in {"cdiv": C / A, "cabs": abs(C), "cconj": C.conj(), "csum": C.sum()}.items():
        out[f"cont/scale/{k}/0"], out[f"cont/scale/{k}/1"] = _parts(v)


def gen_activations(ref, out):
    """nn/activation.py: final activations of ResConv / RBM in container form."""
    act = ref["act"]
    rng = np.random.default_rng(13)
    for dt in (np.float32, np.float64):
        x = (rng.standard_normal((6, 16)) * 30).astype(dt)
        tag = np.dtype(dt).name
        out[f"act/{tag}/x"] = x
        for name in ("exp_by_scale", "sinhp1_by_scale"):
            r = getattr(act, name)(minijax.wrap(x.copy()))
            out[f"act/{tag}/{name}/0"], out[f"act/{tag}/{name}/1"] = _parts(r)
        th = (rng.standard_normal(24) * 3).astype(dt)
        out[f"act/{tag}/theta"] = th
        r = act.prod_by_log(minijax.wrap(np.cosh(th)))
        out[f"act/{tag}/prod_by_log_cosh/0"], out[f"act/{tag}/prod_by_log_cosh/1"] = _parts(r)
        xc = (rng.standard_normal((8, 5)) * 3).astype(dt)
        out[f"act/{tag}/pair_in"] = xc
        out[f"act/{tag}/pair_cpl"] = np.asarray(act.pair_cpl(minijax.wrap(xc.copy())))


def gen_sampler_steps(ref, out):
    """common_samplers.py:28-33,58-82 (proposals) and metropolis.py:291-322 (_update) on INJECTED picks / uniforms:
    ``jr.split`` hands out the injected numbers, ``jr.choice(key, n, p=...)`` returns the injected index,
    ``jr.choice(key, array)`` the array element at the injected index, ``jr.uniform`` the injected uniforms -- the
    mapping from PRNG bits to these numbers is third-party code and stays unpinned."""
    sites, csamp, metro, samples, big = ref["sites"], ref["csamp"], ref["metro"], ref["samples"], ref["big"]
    W = minijax.wrap
    rng = np.random.default_rng(17)

    def choice(key, a, shape=None, p=None, **kw):
        if np.ndim(a) == 0:
            return key
        return W(np.asarray(a))[key]

    for mod in (csamp, metro):
        mod.jr.split = lambda key, n: key
        mod.jr.choice = choice
        mod.jr.uniform = lambda key, shape=None, dtype=None, **kw: key
    sites.Sites._SITES = None
    lat = sites.Square(4, Nparticles=(8, 8))
    N, ns = lat.Nsites, 40
    s = rand_spins(rng, ns, N, 8)
    nbr = np.asarray(csamp._get_site_neighbors([1, 2]))
    nbr[::3, -1] = -1  # some empty table slots, as on lattices with fewer neighbours
    pos = np.array([rng.choice(np.flatnonzero(r == 1)) for r in s])
    slot = rng.integers(0, nbr.shape[1], ns)
    key = W(np.concatenate([pos, slot]))
    new = csamp._propose_exchange(key, W(s.copy()), 1, W(nbr.copy()))
    out["smp/exchange/spins"], out["smp/exchange/nbr"], out["smp/exchange/pos"], out["smp/exchange/slot"] = s, nbr, pos, slot
    out["smp/exchange/new"] = np.asarray(new, dtype=np.int8)
    s1 = rand_spins(rng, ns, N)
    pos1 = rng.integers(0, N, ns)
    out["smp/localflip/spins"], out["smp/localflip/pos"] = s1, pos1
    out["smp/localflip/new"] = np.asarray(csamp.LocalFlip.propose(None, W(pos1.copy()), W(s1.copy())), dtype=np.int8)

    # _update: LogArray amplitudes with cached internals, several reweight exponents; rows with a zero old
    # amplitude, with an unmoved proposal, and with an exact tie rate == 1 - u
    import types as _t

    new_s = np.asarray(new, dtype=np.int8)
    for rw in (2.0, 1.3):
        sg0, sg1 = rng.choice([-1.0, 1.0], ns), rng.choice([-1.0, 1.0], ns)
        l0 = rng.standard_normal(ns) * 2
        l1 = l0 + rng.standard_normal(ns) * 0.7
        u = rng.random(ns)
        moved = np.flatnonzero((new_s != s).any(axis=1))
        i_zero, i_same, i_tie = int(moved[0]), int(moved[1]), int(moved[2])
        l0[i_zero], sg0[i_zero] = -np.inf, 0.0            # |psi_old| == 0 -> accepted whatever u is
        u[i_zero] = 0.0
        new_s2 = new_s.copy()
        new_s2[i_same] = s[i_same]                        # unmoved -> never accepted
        l1[i_same], u[i_same] = l0[i_same] + 5.0, 0.999
        l1[i_tie] = l0[i_tie]
        u[i_tie] = 0.0                                    # rate = 1 == 1 - u: strict '>' rejects
        th0, th1 = rng.standard_normal((ns, 5)), rng.standard_normal((ns, 5))
        old = samples.Samples(W(s.copy()), big.LogArray(W(sg0.copy()), W(l0.copy())), W(th0.copy()))
        prop = samples.Samples(W(new_s2.copy()), big.LogArray(W(sg1.copy()), W(l1.copy())), W(th1.copy()))
        res = metro.Metropolis._update(_t.SimpleNamespace(_reweight=rw), W(u.copy()), None, old, prop)
        k = f"smp/update/rw{rw}"
        out[f"{k}/old_spins"], out[f"{k}/new_spins"] = s, new_s2
        out[f"{k}/special_rows"] = np.array([i_zero, i_same, i_tie])
        out[f"{k}/old_sign"], out[f"{k}/old_logabs"], out[f"{k}/new_sign"], out[f"{k}/new_logabs"] = sg0, l0, sg1, l1
        out[f"{k}/u"], out[f"{k}/old_theta"], out[f"{k}/new_theta"] = u, th0, th1
        out[f"{k}/res_spins"] = np.asarray(res.spins, dtype=np.int8)
        out[f"{k}/res_sign"], out[f"{k}/res_logabs"] = _parts(res.psi)
        out[f"{k}/res_theta"] = np.asarray(res.state_internal)


def gen_full_sweep(ref, out):
    """sampler/metropolis.py:246-322 + common_samplers.py:14-162: the reference's own sweep loop
    (_partial_sweep -> _single_sweep -> propose -> ref_forward_with_updates -> _update -> _get_reweight_factor) for
    SpinExchange and LocalFlip on a stand-in state with a closed-form RBM amplitude.  Random draws are served
    by the stand-ins below and LOGGED (site picked among the allowed ones, neighbour-table slot, uniform), so that
    the oracle can replay exactly the same draws."""
    import types as _t

    sites, big, samples_mod, gd = ref["sites"], ref["big"], ref["samples"], ref["gd"]
    W = minijax.wrap
    sys.modules.pop("quantax.sampler.common_samplers", None)
    cs = importlib.import_module("quantax.sampler.common_samplers")  # now on top of the real Metropolis class
    rng = np.random.default_rng(41)

    class State:
        use_ref = True

        def __init__(self, Wm, b):
            self.Wm, self.b = Wm, b

        def _psi(self, theta):
            c = np.cosh(theta)
            return big.LogArray(W(np.prod(np.sign(c), axis=-1)), W(np.sum(np.log(np.abs(c)), axis=-1)))

        def init_internal(self, s):
            return W(np.asarray(s, dtype=np.float64) @ self.Wm.T + self.b)

        def __call__(self, s):
            return self._psi(np.asarray(self.init_internal(s)))

        def ref_forward_with_updates(self, s, s_old, nflips, internal):
            theta = self.init_internal(s)  # direct evaluation: the local-update formula is pinned separately
            return self._psi(np.asarray(theta)), theta

    for kind, reweight in (("exchange", 2.0), ("exchange", 1.5), ("localflip", 2.0)):
        sites.Sites._SITES = None
        lat = sites.Square(4, Nparticles=(8, 8)) if kind == "exchange" else sites.Chain(10)
        N, ns, nsweeps = lat.Nsites, 12, 25
        Wm, b = rng.standard_normal((6, N)) * 0.4, rng.standard_normal(6) * 0.1
        spins = rand_spins(rng, ns, N, 8 if kind == "exchange" else None)
        draws = {"pos": rng.integers(0, 1 << 30, (nsweeps, ns)), "slot": rng.integers(0, 1 << 30, (nsweeps, ns)),
                 "u": rng.random((nsweeps, ns))}
        log = {"pos": [], "slot": []}
        calls = {"n": 0}

        def get_subkeys(num=None):
            calls["n"] += 1
            if calls["n"] % 2 == 1:  # keys_propose, then keys_update (metropolis.py:253-254)
                if kind == "exchange":
                    return [W(np.concatenate([draws["pos"][t], draws["slot"][t]])) for t in range(num)]
                return [W(draws["pos"][t].copy()) for t in range(num)]
            return [W(draws["u"][t].copy()) for t in range(num)]

        def choice(key, a, shape=None, p=None, **kw):
            if np.ndim(a) == 0:
                if shape is not None:  # LocalFlip: one site per chain in one call
                    val = np.asarray(key) % int(a)
                    log["pos"].extend(val.tolist())
                    return W(val)
                valid = np.flatnonzero(np.asarray(p))
                val = int(valid[int(key) % valid.size])
                log["pos"].append(val)
                return np.int64(val)
            idx = int(key) % len(a)
            log["slot"].append(idx)
            return W(np.asarray(a))[idx]

        metro = ref["metro"]
        for mod in (cs, metro):
            mod.jr.split = lambda key, n: key
            mod.jr.choice = choice
            mod.jr.uniform = lambda key, shape=None, dtype=None, **kw: key
            mod.get_subkeys = get_subkeys
        cls = cs.SpinExchange if kind == "exchange" else cs.LocalFlip
        smp = object.__new__(cls)
        smp._state, smp._nsamples, smp._reweight = State(Wm, b), ns, W(np.asarray(reweight))
        if kind == "exchange":
            smp._hopping_particle, smp._neighbors = 1, cs._get_site_neighbors(1)
        res = smp._partial_sweep(nsweeps, W(spins.copy()))
        k = f"sweep/{kind}_rw{reweight}"
        out[f"{k}/W"], out[f"{k}/b"], out[f"{k}/spins0"] = Wm, b, spins
        out[f"{k}/pos"] = np.array(log["pos"], dtype=np.int64).reshape(nsweeps, ns)
        if kind == "exchange":
            out[f"{k}/slot"] = np.array(log["slot"], dtype=np.int64).reshape(nsweeps, ns)
        out[f"{k}/u"] = draws["u"]
        out[f"{k}/spins"] = np.asarray(res.spins, dtype=np.int8)
        out[f"{k}/sign"], out[f"{k}/logabs"] = _parts(res.psi)
        out[f"{k}/reweight_factor"] = np.asarray(res.reweight_factor)
        assert res.state_internal is None


def gen_chunk_and_mix_sweeps(ref, out):
    """sampler/metropolis.py:217-244 (_chunk_sweep: states without local updates evaluate psi only for the chains whose
    proposal moved, through _get_update_size / _get_updated_spins / _get_new_psi) and :325-428 (MixSampler: every
    step is proposed by a randomly chosen component).  Draws are served and logged as in gen_full_sweep."""
    sites, big, metro = ref["sites"], ref["big"], ref["metro"]
    W = minijax.wrap
    cs = sys.modules["quantax.sampler.common_samplers"]
    rng = np.random.default_rng(43)
    sites.Sites._SITES = None
    lat = sites.Square(4, Nparticles=(8, 8))
    N, ns, nsweeps = lat.Nsites, 12, 20
    Wm, b = rng.standard_normal((6, N)) * 0.4, rng.standard_normal(6) * 0.1

    class State:
        use_ref = False
        forward_calls = []

        def __call__(self, s):
            s = np.asarray(s, dtype=np.float64)
            State.forward_calls.append(s.shape[0])
            c = np.cosh(s @ Wm.T + b)
            return big.LogArray(W(np.prod(np.sign(c), axis=-1)), W(np.sum(np.log(np.abs(c)), axis=-1)))

        def init_internal(self, s):
            return None

        def ref_forward_with_updates(self, s, s_old, nflips, internal):
            return self(s), None

    for scenario in ("chunk", "mix"):
        spins = rand_spins(rng, ns, N, 8)
        draws = {"pos": rng.integers(0, 1 << 30, (nsweeps, ns)), "slot": rng.integers(0, 1 << 30, (nsweeps, ns)),
                 "u": rng.random((nsweeps, ns)), "comp": rng.integers(0, 1 << 30, nsweeps)}
        log = {"pos": [], "slot": [], "comp": []}
        order = (["comp"] if scenario == "mix" else []) + ["propose", "update"]
        calls = {"n": 0}

        def get_subkeys(num=None):
            what = order[calls["n"] % len(order)]
            calls["n"] += 1
            if what == "comp":
                return W(draws["comp"].copy())
            if what == "propose":
                return [W(np.concatenate([draws["pos"][t], draws["slot"][t]])) for t in range(num)]
            return [W(draws["u"][t].copy()) for t in range(num)]

        def choice(key, a, shape=None, p=None, **kw):
            if np.ndim(a) == 0 and shape is not None:  # MixSampler._rand_sampler_idx
                val = np.asarray(key) % int(a)
                log["comp"].extend(val.tolist())
                return W(val)
            if np.ndim(a) == 0:
                valid = np.flatnonzero(np.asarray(p))
                val = int(valid[int(key) % valid.size])
                log["pos"].append(val)
                return np.int64(val)
            idx = int(key) % len(a)
            log["slot"].append(idx)
            return W(np.asarray(a))[idx]

        for mod in (cs, metro):
            mod.jr.split = lambda key, n: key
            mod.jr.choice = choice
            mod.jr.uniform = lambda key, shape=None, dtype=None, **kw: key
            mod.get_subkeys = get_subkeys
        state = State()
        State.forward_calls = []

        def component(n_neighbor):
            c = object.__new__(cs.SpinExchange)
            c._state, c._nsamples, c._reweight = state, ns, W(np.asarray(2.0))
            c._hopping_particle, c._neighbors = 1, cs._get_site_neighbors(n_neighbor)
            return c

        k = f"sweep/{scenario}"
        if scenario == "chunk":
            smp = component(1)
            smp._spins = W(spins.copy())
            res = smp._chunk_sweep(nsweeps, 4)
            out[f"{k}/forward_batch_sizes"] = np.array(State.forward_calls, dtype=np.int64)
        else:
            comps = (component(1), component(2))
            smp = object.__new__(metro.MixSampler)
            smp._state, smp._nsamples, smp._reweight = state, ns, W(np.asarray(2.0))
            smp._samplers, smp._ratio = comps, W(np.array([0.5, 0.5]))
            res = smp._partial_sweep(nsweeps, W(spins.copy()))
            out[f"{k}/comp"] = np.array(log["comp"], dtype=np.int64)
            out[f"{k}/nbr2"] = np.asarray(cs._get_site_neighbors(2), dtype=np.int64)
        out[f"{k}/W"], out[f"{k}/b"], out[f"{k}/spins0"] = Wm, b, spins
        out[f"{k}/pos"] = np.array(log["pos"], dtype=np.int64).reshape(nsweeps, ns)
        out[f"{k}/slot"] = np.array(log["slot"], dtype=np.int64).reshape(nsweeps, ns)
        out[f"{k}/u"] = draws["u"]
        out[f"{k}/spins"] = np.asarray(res.spins, dtype=np.int8)
        out[f"{k}/sign"], out[f"{k}/logabs"] = _parts(res.psi)


def gen_optimizer(ref, out):
    """optimizer/sr.py: SR.get_step (Ebar, Obar, _Omean, energy, VarE, real_to_complex stacking) with a stand-in
    state / Hamiltonian that return given Jacobians and local energies, and three consecutive solves of SPRING,
    MARCH and AdamSR (lines 74-123, 180-195, 247-255, 321-341, 405-421)."""
    import types as _t

    srmod, solver, samples, VS = ref["srmod"], ref["solver"], ref["samples"], ref["VS_TYPE"]
    W = minijax.wrap
    rng = np.random.default_rng(23)
    ns, npar = 14, 37

    def make(cls, vs_type, Omat, Eloc, **attrs):
        opt = object.__new__(cls)
        opt._state = _t.SimpleNamespace(jacobian=lambda spins: W(Omat.copy()), vs_type=vs_type, _holomorphic=False)
        opt._hamiltonian = _t.SimpleNamespace(Oloc=lambda state, smp: W(Eloc.copy()))
        opt._imag_time, opt._solver = True, solver.auto_pinv_eig(rtol=1e-10)
        opt._energy = opt._VarE = None
        for k, v in attrs.items():
            setattr(opt, k, v)
        return opt

    rw = rng.random(ns) + 0.5
    rw /= rw.mean()
    spins = rand_spins(rng, ns, 8)
    for tag, cplx in (("real", False), ("real_to_complex", True)):
        Omat = rng.standard_normal((ns, npar))
        Eloc = rng.standard_normal(ns) * 2 - 7
        if cplx:
            Omat = Omat + 1j * rng.standard_normal((ns, npar))
            Eloc = Eloc + 1j * rng.standard_normal(ns)
            ref["gd"].set_default_dtype(np.complex128)
        smp = samples.Samples(W(spins.copy()), W(np.ones(ns)), None, W(rw.copy()))
        opt = make(srmod.SR, VS.real_to_complex if cplx else VS.real_or_holomorphic, Omat, Eloc)
        step = opt.get_step(smp)
        k = f"opt/sr_{tag}"
        out[f"{k}/Omat"], out[f"{k}/Eloc"], out[f"{k}/rw"] = Omat, Eloc, rw
        out[f"{k}/Ebar"] = np.asarray(opt.get_Ebar(smp))
        out[f"{k}/Obar"] = np.asarray(opt.get_Obar(smp))
        out[f"{k}/Omean"] = np.asarray(opt._Omean)
        out[f"{k}/energy"], out[f"{k}/VarE"] = np.asarray(opt._energy), np.asarray(opt._VarE)
        out[f"{k}/step"] = np.asarray(step)
        ref["gd"].set_default_dtype(np.float64)
    Obars = rng.standard_normal((3, ns, npar)) / np.sqrt(ns)
    Ebars = rng.standard_normal((3, ns)) / np.sqrt(ns)
    out["opt/momentum/Obar"], out["opt/momentum/Ebar"] = Obars, Ebars
    zeros = lambda: W(np.zeros(npar))
    opts = {"spring": make(srmod.SPRING, VS.real_or_holomorphic, Obars[0], Ebars[0], _mu=0.9, _last_step=zeros()),
            "march": make(srmod.MARCH, VS.real_or_holomorphic, Obars[0], Ebars[0], _mu=0.95, _beta=0.995,
                          _last_step=zeros(), _V=zeros(), _t=0),
            "adamsr": make(srmod.AdamSR, VS.real_or_holomorphic, Obars[0], Ebars[0], _mu=0.95, _beta=0.995, _m=zeros(),
                           _v=zeros(), _t=0)}
    for name, opt in opts.items():
        out[f"opt/momentum/{name}"] = np.stack([np.asarray(opt.solve(W(Obars[i].copy()), W(Ebars[i].copy())))
                                                for i in range(3)])


def gen_time_evol(ref, out):
    """optimizer/time_evol.py:55-134: TimeEvol.get_step for a real-parameter / complex-output state, direct
    (S = Obar^+ Obar, F = Obar^+ Ebar) and chunked (_get_SF_indirect: un-centred sums corrected at the end)."""
    import types as _t

    temod, solver, samples, VS, gd = ref["temod"], ref["solver"], ref["samples"], ref["VS_TYPE"], ref["gd"]
    W = minijax.wrap
    rng = np.random.default_rng(37)
    ns, npar, N = 12, 7, 10
    spins = rand_spins(rng, ns, N)
    while len({r.tobytes() for r in spins}) < ns:
        spins = rand_spins(rng, ns, N)
    Omat = rng.standard_normal((ns, npar)) + 1j * rng.standard_normal((ns, npar))
    Eloc = rng.standard_normal(ns) * 2 - 5 + 1j * rng.standard_normal(ns)
    row = {r.tobytes(): i for i, r in enumerate(spins)}
    gd.set_default_dtype(np.complex128)
    out["tevol/Omat"], out["tevol/Eloc"] = Omat, Eloc
    for tag, maxp in (("direct", None), ("chunked", 4)):
        opt = object.__new__(temod.TimeEvol)
        opt._state = _t.SimpleNamespace(
            jacobian=lambda s: W(Omat[[row[np.asarray(r, dtype=np.int8).tobytes()] for r in np.asarray(s)]].copy()),
            vs_type=VS.real_to_complex, _holomorphic=False, nparams=npar)
        opt._hamiltonian = _t.SimpleNamespace(Oloc=lambda state, smp: W(Eloc.copy()))
        opt._imag_time, opt._solver, opt._max_parallel = False, solver.pinvh_solve(rtol=1e-10), maxp
        opt._energy = opt._VarE = None
        smp = samples.Samples(W(spins.copy()), W(np.ones(ns)), None, W(np.ones(ns)))
        S, F = opt.get_SF(smp)
        out[f"tevol/{tag}/S"], out[f"tevol/{tag}/F"] = np.asarray(S), np.asarray(F)
        out[f"tevol/{tag}/energy"], out[f"tevol/{tag}/VarE"] = np.asarray(opt._energy), np.asarray(opt._VarE)
        out[f"tevol/{tag}/step"] = np.asarray(opt.get_step(smp))
    gd.set_default_dtype(np.float64)


def gen_local_updates(ref, out):
    """model/shallow_nets.py:87-108: SingleDense.ref_forward (local update of theta from the flipped sites, then
    prod_by_log(cosh(theta))), called unbound with a stand-in that carries the weight matrix."""
    import types as _t

    shallow, act = ref["shallow"], ref["act"]
    if shallow is None:
        return
    W = minijax.wrap
    rng = np.random.default_rng(29)
    N, M = 16, 12
    import jax.numpy as jnp

    for dt in (np.float32, np.float64):
        Wm = (rng.standard_normal((M, N)) * 0.4).astype(dt)
        b = (rng.standard_normal(M) * 0.1).astype(dt)
        stub = _t.SimpleNamespace(layers=[_t.SimpleNamespace(weight=W(Wm.copy())), lambda x: jnp.cosh(x), act.prod_by_log])
        tag = np.dtype(dt).name
        out[f"rbm/{tag}/W"], out[f"rbm/{tag}/b"] = Wm, b
        for nflips in (1, 2):
            s_old = rand_spins(rng, 9, N, 8)
            s_new = s_old.copy()
            for r in range(9):
                if nflips == 1:
                    s_new[r, rng.integers(N)] *= -1
                else:
                    i, j = rng.choice(np.flatnonzero(s_old[r] == 1)), rng.choice(np.flatnonzero(s_old[r] == -1))
                    s_new[r, i], s_new[r, j] = -1, 1
            theta = (s_old.astype(dt) @ Wm.T + b).astype(dt)
            signs, logs, thetas = [], [], []
            for r in range(9):
                psi, th = shallow.SingleDense.ref_forward(stub, W(s_new[r].copy()), W(s_old[r].copy()), nflips,
                                                          W(theta[r].copy()), return_update=True)
                sg, lg = _parts(psi)
                signs.append(sg), logs.append(lg), thetas.append(np.asarray(th))
            k = f"rbm/{tag}/nflips{nflips}"
            out[f"{k}/s_old"], out[f"{k}/s_new"], out[f"{k}/theta_old"] = s_old, s_new, theta
            out[f"{k}/sign"], out[f"{k}/logabs"], out[f"{k}/theta_new"] = np.array(signs), np.array(logs), np.array(thetas)


def gen_resconv(ref, out):
    """model/conv_nets.py:26-183 + nn/conv.py:13-68 + nn/activation.py + symmetry.symmetrize: the reference's ResConv
    forward (block scalings, gelu placement, residual tiling, bias on all but the last convolution, final scaling,
    pair_cpl, final activation, channel mean, translation sum) with injected weights, per sample; and the
    symmetry-projected amplitude  symmetrize(vmap(model)(get_symm_spins(s)))  of state/variational.py:262-266.
    equinox.nn.Conv and jax.nn.gelu are the documented stand-ins of minijax (third party)."""
    convnets, sites, act, symm_pkg, gd = ref["convnets"], ref["sites"], ref["act"], ref["symm_pkg"], ref["gd"]
    if convnets is None:
        return
    W = minijax.wrap
    rng = np.random.default_rng(31)
    cases = {
        "sq4_f64_exp": dict(lat=lambda: sites.Square(4, Nparticles=(8, 8)), nb=2, C=4, k=3, dt=np.float64, final="exp", cplx=False),
        "sq4_f64_sinhp1": dict(lat=lambda: sites.Square(4, Nparticles=(8, 8)), nb=3, C=4, k=3, dt=np.float64, final="sinhp1", cplx=False),
        "sq6_f32_sinhp1": dict(lat=lambda: sites.Square(6, Nparticles=(18, 18)), nb=2, C=8, k=3, dt=np.float32, final="sinhp1", cplx=False),
        "chain8_f64_exp": dict(lat=lambda: sites.Chain(8), nb=2, C=4, k=3, dt=np.float64, final="exp", cplx=False),
        "tri6_f64_cplx": dict(lat=lambda: sites.Triangular(6, Nparticles=(18, 18)), nb=2, C=4, k=3, dt=np.float64, final="exp", cplx=True),
    }
    for name, c in cases.items():
        sites.Sites._SITES = None
        lat = c["lat"]()
        gd.set_default_dtype(np.complex128 if c["cplx"] else np.float64)
        fa = act.exp_by_scale if c["final"] == "exp" else act.sinhp1_by_scale
        model = convnets.ResConv(c["nb"], c["C"], c["k"], final_activation=fa, dtype=c["dt"],
                                 out_dtype=np.complex128 if c["cplx"] else None)
        blocks = [l for l in model.layers if isinstance(l, convnets._ConvBlock)]
        assert len(blocks) == c["nb"]
        for i, blk in enumerate(blocks):
            for cname in ("conv1", "conv2"):
                conv = getattr(blk, cname)
                fan_in = int(np.prod(conv.weight.shape[1:]))
                conv.weight = W((rng.standard_normal(conv.weight.shape) * np.sqrt(2.0 / fan_in)).astype(c["dt"]))
                out[f"resconv/{name}/block{i}.{cname}.weight"] = np.asarray(conv.weight)
                if conv.bias is not None:
                    conv.bias = W((rng.standard_normal(conv.bias.shape) * 0.1).astype(c["dt"]))
                    out[f"resconv/{name}/block{i}.{cname}.bias"] = np.asarray(conv.bias).reshape(-1)
        nup = None if name.startswith("chain") else lat.Nsites // 2
        s = rand_spins(rng, 6, lat.Nsites, nup)
        out[f"resconv/{name}/spins"] = s
        res = [_parts(model(W(r.copy()))) for r in s]
        out[f"resconv/{name}/significand"] = np.array([r[0] for r in res])
        out[f"resconv/{name}/exponent"] = np.array([r[1] for r in res])
        # symmetry-projected amplitude
        if name == "sq4_f64_exp":
            symm = symm_pkg.Rotation(np.pi / 2, sector=2) @ symm_pkg.Flip() @ symm_pkg.SpinInverse(-1)
        elif name == "tri6_f64_cplx":
            symm = symm_pkg.D6(center=(0, 0)) @ symm_pkg.SpinInverse()
        else:
            symm = None
        if symm is not None:
            import jax

            rows = []
            for r in s:
                imgs = symm.get_symm_spins(W(r.copy()))
                psi = jax.vmap(model)(imgs)
                rows.append(_parts(symm.symmetrize(psi, W(r.copy()))))
            out[f"resconv/{name}/symm_perm"] = np.asarray(symm._perm, dtype=np.int64)
            out[f"resconv/{name}/symm_character"] = np.asarray(symm._character)
            out[f"resconv/{name}/symm_Z2"] = np.asarray(symm.Z2_inversion)
            out[f"resconv/{name}/proj_significand"] = np.array([r[0] for r in rows])
            out[f"resconv/{name}/proj_exponent"] = np.array([r[1] for r in rows])
        gd.set_default_dtype(np.float64)


def gen_chunk_map(ref, out):
    """utils/function.py:12-146: the reference's own chunk_map (interleaved chunk composition [devices, chunk, nchunks],
    zero padding, re-assembly) on one device: which samples every chunk sees and what comes back."""
    fn = importlib.import_module("quantax.utils.function")
    W = minijax.wrap
    cases = [(10, 4), (8, 4), (5, 8), (7, 7), (13, 5), (1, 3)]
    out["chunk/cases"] = np.asarray(cases)
    for ci, (B, cs) in enumerate(cases):
        x = (np.arange(B * 3, dtype=np.float64).reshape(B, 3) + 1.0)
        w = np.linspace(0.5, 1.5, 3)
        seen = []

        def f(xc, ww):  # per-sample function of the chunk; records the chunk it was given
            seen.append(np.asarray(xc).copy())
            return W(np.asarray(xc) @ np.asarray(ww)), W(np.asarray(xc).T * 2.0)  # batch axis 0 and batch axis 1

        y, z = fn.chunk_map(f, in_axes=(0, None), out_axes=(0, 1), chunk_size=cs)(W(x.copy()), W(w.copy()))
        out[f"chunk/{ci}/x"], out[f"chunk/{ci}/w"] = x, w
        out[f"chunk/{ci}/nchunks"] = np.asarray(len(seen))
        for k, c in enumerate(seen):
            out[f"chunk/{ci}/seen{k}"] = c
        out[f"chunk/{ci}/y"], out[f"chunk/{ci}/z"] = np.asarray(y), np.asarray(z)


def main():
    import warnings

    warnings.simplefilter("ignore")
    ref = install_reference()
    out = {}
    gen_operator(ref, out)
    gen_solver(ref, out)
    gen_symmetry(ref, out)
    gen_sampler_tables(ref, out)
    gen_sign(ref, out)
    gen_more_lattices(ref, out)
    gen_operator_algebra(ref, out)
    gen_containers(ref, out)
    gen_activations(ref, out)
    gen_sampler_steps(ref, out)
    gen_optimizer(ref, out)
    gen_local_updates(ref, out)
    gen_resconv(ref, out)
    gen_time_evol(ref, out)
    gen_full_sweep(ref, out)
    gen_chunk_and_mix_sweeps(ref, out)
    gen_chunk_map(ref, out)
    path = os.path.join(HERE, "ref_hotpath.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
