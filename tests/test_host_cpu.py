"""CPU tests of the product's host layer: lattice tables, operator lists, term-table compilation,
and that the C-ABI library loads and exports every symbol declared in include/qtx_b200.h."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import operator as oop, sites as osites

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_tables.npz"))


def _mk(name):
    from quantax_b200 import sites

    return {
        "chain8": lambda: sites.Chain(8),
        "square4": lambda: sites.Square(4, Nparticles=(8, 8)),
        "square10": lambda: sites.Square(10, Nparticles=(50, 50)),
        "square16": lambda: sites.Square(16, Nparticles=(128, 128)),
        "triangular12": lambda: sites.Triangular(12, Nparticles=(72, 72)),
    }[name]()


@pytest.mark.parametrize("name", ["chain8", "square4", "square10", "square16", "triangular12"])
def test_product_bond_tables_match_reference(name):
    lat = _mk(name)
    assert np.allclose(lat.coord, GOLD[f"{name}/coord"])
    for n in (1, 2):
        assert np.array_equal(lat.get_neighbor(n), GOLD[f"{name}/nb{n}"])
    assert [np.array_equal(a, b) for a, b in zip(lat.get_neighbor([1, 2]), [GOLD[f"{name}/nb1"], GOLD[f"{name}/nb2"]])]


@pytest.mark.parametrize("name", ["square4", "square10"])
def test_product_op_lists_match_reference(name):
    from quantax_b200 import operator

    _mk(name)
    ops = {"heis": operator.Heisenberg(), "heis_msr": operator.Heisenberg(msr=True),
           "j1j2_msr": operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)}
    for oname, op in ops.items():
        assert [o for o, _ in op.op_list] == list(GOLD[f"{name}/{oname}/names"])
        J = np.array([t[0] for _, ts in op.op_list for t in ts], dtype=np.float64)
        assert np.array_equal(J, GOLD[f"{name}/{oname}/J"])
        idx = [list(t[1:]) for _, ts in op.op_list for t in ts]
        for a, b in zip(idx, GOLD[f"{name}/{oname}/idx"]):
            assert a == [v for v in b if v >= 0]


def test_product_ising_op_list_matches_reference():
    from quantax_b200 import operator

    _mk("chain8")
    op = operator.Ising(h=1.0)
    assert [o for o, _ in op.op_list] == list(GOLD["chain8/ising_h1/names"])
    J = np.array([t[0] for _, ts in op.op_list for t in ts])
    assert np.array_equal(J, GOLD["chain8/ising_h1/J"])


def test_operator_algebra():
    from quantax_b200 import operator as O

    _mk("square4")
    a = O.sigma_p(0) @ O.sigma_m(1)
    assert a.op_list == [["+-", [[1.0, 0, 1]]]]
    assert a.H.op_list == [["+-", [[1.0, 1, 0]]]]  # (S+_0 S-_1)^dagger = S+_1 S-_0 (operator.py:361-375)
    b = 2 * a + O.sigma_z(2) @ O.sigma_z(3) - a
    names = [o for o, _ in b.op_list]
    assert names == ["+-", "zz"]
    assert [t[0] for t in b.op_list[0][1]] == [2.0, -1.0]
    with pytest.raises(ValueError):
        a + 1.0
    assert (a + 0) is a
    assert (a / 2).op_list[0][1][0][0] == 0.5
    assert O.sigma_x(1, 2).op_list == [["x", [[2.0, 6]]]]  # coordinate indexing (site_operator.py:19-32)


def test_compile_terms_matches_oracle_semantics():
    from quantax_b200 import operator as O

    _mk("square4")
    H = O.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    coef, sites, ops, nfl = O.compile_terms(H.op_list)
    assert coef.shape == (192,) and sites.shape == (192, 4) and ops.shape == (192, 4)
    assert list(nfl[:128]) == [2] * 128 and list(nfl[128:]) == [0] * 64
    assert np.array_equal(ops[0], [3, 4, 0, 0]) and np.array_equal(ops[64], [4, 3, 0, 0])
    assert np.array_equal(ops[128], [1, 1, 0, 0])
    with pytest.raises(NotImplementedError):
        O.compile_terms([["y", [[1.0, 0]]]])
    with pytest.raises(NotImplementedError):
        O.compile_terms([["+-", [[1.0, 0, 0]]]])


def test_site_neighbor_table_matches_oracle():
    from quantax_b200.sampler import _site_neighbors

    _mk("square10")
    assert np.array_equal(_site_neighbors(1), osites.site_neighbor_table(osites.Square(10), 1))
    assert np.array_equal(_site_neighbors([1, 2]), osites.site_neighbor_table(osites.Square(10), [1, 2]))


def test_error_behaviour_of_constructors():
    """Same exceptions as the reference: sampler.py:29-33, common_samplers.py:132-136, sites.py:71-81."""
    from quantax_b200 import sites

    with pytest.raises(ValueError):
        sites.Square(4, Nparticles=(8, 7))
    sites.Sites._SITES = None
    with pytest.raises(ValueError):
        sites.Square(4, Nparticles=8)


def test_c_abi_exports_every_declared_symbol():
    """The shared library must load without a GPU and export exactly the header's entry points."""
    from quantax_b200 import _lib

    header = open(os.path.join(ROOT, "include", "qtx_b200.h")).read()
    declared = set(re.findall(r"\b(qtx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert os.path.exists(_lib.LIB_PATH), "libqtx_b200.so is not built (run __graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert handle.qtx_abi_version() == 1


def test_c_abi_rejects_null_pointers_with_a_status_and_a_message():
    """Error behaviour of the boundary (include/qtx_b200.h conventions): every int-returning entry point validates
    its arguments before touching the device -- null pointers with non-empty sizes give a negative qtx_status and a
    message from qtx_last_error(); empty batches are no-ops or rejected, never a crash.  No compute call is made."""
    from quantax_b200 import _lib

    L = _lib.lib()
    # free / close / destroy of NULL is a no-op, like cudaFree(0); size / rank of no communicator are plain queries
    no_pointer_check = {"qtx_peer_free", "qtx_peer_close", "qtx_comm_destroy", "qtx_comm_size", "qtx_comm_rank"}
    checked = 0
    for name, (res, args) in _lib.SIGNATURES.items():
        if res is not ctypes.c_int or ctypes.c_void_p not in args:
            continue
        for fill in (1, 0):
            vals = [None if a is ctypes.c_void_p else (0.0 if a is ctypes.c_double else fill) for a in args]
            rc = getattr(L, name)(*vals)
            if fill == 1 and name not in no_pointer_check:
                assert rc in (-1, -3), f"{name}(null pointers) returned {rc}"
                assert name in L.qtx_last_error().decode(), f"{name}: message does not name the entry point"
                checked += 1
            elif name not in ("qtx_comm_size", "qtx_comm_rank"):
                assert rc <= 0
    assert checked >= 55


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "quantax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_no_cuda_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from quantax_b200 import _lib

    with pytest.raises(_lib.QtxError):
        _lib.ptr(torch.zeros(4))


SYMM_GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_symm_generators.npz"))


@pytest.mark.parametrize("name", ["square4", "square6", "square10", "triangular6", "chain8"])
def test_symmetry_generators_match_reference(name):
    """Translation / Flip / Rotation permutations equal those of the reference's own NumPy code
    (quantax/symmetry/translation.py:25-47, common_symmetries.py:104-205)."""
    from quantax_b200 import sites, symmetry

    lat = {"square4": lambda: sites.Square(4), "square6": lambda: sites.Square(6), "square10": lambda: sites.Square(10),
           "triangular6": lambda: sites.Triangular(6), "chain8": lambda: sites.Chain(8)}[name]()
    nd = lat.ndim
    c = np.zeros(nd) if name.startswith("tri") else None
    assert np.array_equal(symmetry.Translation(np.eye(nd, dtype=int))._generator, SYMM_GOLD[f"{name}/trans"])
    assert np.array_equal(symmetry.Flip(0, center=c)._generator, SYMM_GOLD[f"{name}/flip0"])
    if nd == 2:
        ang = np.pi / 3 if name.startswith("tri") else np.pi / 2
        assert np.array_equal(symmetry.Rotation(ang, center=c)._generator, SYMM_GOLD[f"{name}/rot"])
        assert np.array_equal(symmetry.Flip(1, center=c)._generator, SYMM_GOLD[f"{name}/flip1"])


def test_symmetry_group_tables_match_oracle():
    from oracle import symmetry as osym
    from quantax_b200 import sites, symmetry

    sites.Square(4, Nparticles=(8, 8))
    olat = osites.Square(4)
    full = symmetry.TransND() @ symmetry.Rotation(np.pi / 2) @ symmetry.Flip() @ symmetry.SpinInverse()
    ofull = osym.TransND(olat) @ osym.Rotation(olat, np.pi / 2) @ osym.Flip(olat) @ osym.SpinInverse(olat)
    assert full.nsymm == ofull.nsymm == 256
    assert np.array_equal(full.perm, ofull.perm) and np.array_equal(full.character, ofull.character)
    assert np.allclose(full.weights(), ofull.weights())
    b1 = symmetry.C4v(repr="B1")
    ob1 = osym.Rotation(olat, np.pi / 2, sector=2) @ osym.Flip(olat, sector=0)
    assert np.array_equal(b1.perm, ob1.perm) and np.array_equal(b1.character, ob1.character)
    assert set(b1.character) == {1.0, -1.0}
    # group property: the permutation set is closed under composition
    perms = {tuple(p) for p in b1.perm}
    assert all(tuple(p[q]) in perms for p in b1.perm for q in b1.perm)
    with pytest.raises(ValueError):
        symmetry.Rotation(np.pi / 2, sector=1)  # complex character under a real default dtype
    with pytest.raises(ValueError):
        symmetry.Z2Inversion(1) @ symmetry.Z2Inversion(-1)
    s = np.arange(16)
    assert np.array_equal(symmetry.Identity().get_symm_spins(s), s[None])


def test_eqx_leaf_file_roundtrip(tmp_path):
    """The parameter files follow eqx.tree_serialise_leaves (variational.py:581-587): np.save blobs back to back,
    arrays in flatten order, then the bool / int dataclass leaves."""
    import io

    import numpy as np

    from quantax_b200.utils import read_eqx_leaves, write_eqx_leaves

    rng = np.random.default_rng(0)
    arrays = [rng.standard_normal((4, 1, 3, 3)).astype(np.float32), rng.standard_normal((4, 1, 1)).astype(np.float32),
              rng.standard_normal((4, 4, 3, 3)).astype(np.float32)]
    path = tmp_path / "model.eqx"
    write_eqx_leaves(path, arrays, [False, 1, 4, 3])
    leaves = read_eqx_leaves(path)
    assert len(leaves) == 7
    for a, b in zip(arrays, leaves[:3]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert [l.ndim for l in leaves[3:]] == [0, 0, 0, 0] and int(leaves[5]) == 4
    # the same bytes as a sequence of np.save calls (what jnp.save / np.save write for equinox)
    buf = io.BytesIO()
    for a in arrays:
        np.save(buf, a)
    for v in (False, 1, 4, 3):
        np.save(buf, np.asarray(v))
    assert path.read_bytes() == buf.getvalue()


def test_rational_shift_masks_cover_every_shift_once():
    """Rank split of the eigendecomposition-free pseudo-inverse (optimizer.rational_shift_masks)."""
    from quantax_b200.optimizer import rational_shift_masks

    for P in range(1, 9):
        masks = rational_shift_masks(P)
        assert len(masks) == P and all(0 <= m < 8 for m in masks)
        assert sum(masks) == 7 and all((a & b) == 0 for i, a in enumerate(masks) for b in masks[i + 1:])
        assert max(bin(m).count("1") for m in masks) == -(-3 // min(P, 3))  # balanced: ceil(3 / min(P, 3)) shifts


def test_apply_off_diag_scatter_logic_on_cpu(monkeypatch):
    """Operator.apply_off_diag (the reference's dense layout, operator.py:96-119) scatters the compacted enumeration
    of get_conn back to [ns, nterms]; here get_conn is replaced by a compaction of the ORACLE's dense tensors, so the
    Python scatter is checked without a GPU (the GPU test of the same method is gated until its first run)."""
    import types

    import torch

    from oracle import operator as oop, sampler as osmp, sites as osites
    from quantax_b200 import operator as qop, sites

    sites.Sites._SITES = None
    sites.Square(4, Nparticles=(8, 8))
    olat = osites.Square(4, Nparticles=(8, 8))
    H = qop.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    aop = oop.to_array_op_list(oop.heisenberg_op_list(olat, J=[1, 0.5], n_neighbor=[1, 2], msr=True))
    s = osmp.rand_states(12, 16, 8, seed=5)
    ref = oop.apply_off_diag(s, aop)

    def fake_get_conn(spins, nflips, conn_size=None, with_spins=True):
        sc, Hc = ref[nflips]
        keep = ~np.isnan(Hc) & (np.abs(Hc) > 1e-8)
        seg, idx = np.nonzero(keep)
        return (torch.from_numpy(seg.astype(np.int32)), torch.from_numpy(idx.astype(np.int32)),
                torch.from_numpy(sc[seg, idx]), torch.from_numpy(Hc[seg, idx]), int(keep.sum()))

    monkeypatch.setattr(qop, "_as_spins", lambda x: torch.as_tensor(np.asarray(x), dtype=torch.int8))
    H._group_tables = {nf: types.SimpleNamespace(nterms=ref[nf][1].shape[1]) for nf in ref}
    monkeypatch.setattr(H, "get_conn", fake_get_conn)
    got = H.apply_off_diag(s)
    assert sorted(got) == sorted(ref)
    for nflips, (sc, Hc) in got.items():
        sc, Hc = sc.numpy(), Hc.numpy()
        rs, rH = ref[nflips]
        ok = ~np.isnan(rH) & (np.abs(rH) > 1e-8)
        assert np.array_equal(np.isnan(Hc), ~ok)
        assert np.array_equal(Hc[ok], rH[ok]) and np.array_equal(sc[ok], rs[ok])
        assert np.array_equal(sc[~ok], np.repeat(s[:, None, :], rH.shape[1], axis=1)[~ok])
