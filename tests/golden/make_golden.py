"""Generate golden fixtures by running the REFERENCE's own NumPy/Python code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

jax / equinox / lrux / quspin are not installable here, so the reference package cannot
be imported as a whole.  The pieces of the hot path that are pure NumPy / pure Python --
``quantax.sites`` (lattice geometry, neighbour shells) and the operator algebra in
``quantax.operator`` (op lists of Heisenberg / Ising) -- are imported with inert stand-ins
for the jax-dependent sibling modules.  Only data produced by unmodified reference code is
written to ``tests/golden/ref_tables.npz``; nothing from /root/reference is copied.
"""
import enum
import importlib
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class _Inert:
    """Stands in for jax/equinox objects: usable as decorator, attribute bag, base class arg."""

    def __call__(self, *a, **k):
        if a and callable(a[0]) and not isinstance(a[0], _Inert):
            return a[0]
        return self

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __getitem__(self, item):
        return self

    def __mro_entries__(self, bases):
        return (object,)


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()


def _install_stubs():
    for name in ["jax", "jax.numpy", "jax.random", "jax.tree_util", "jax.flatten_util", "jax.lax",
                 "jax.sharding", "jax.scipy", "jax.scipy.linalg", "jax.typing", "jaxtyping", "equinox",
                 "equinox.nn", "lrux", "quspin"]:
        sys.modules[name] = _InertModule(name)
    eqx = sys.modules["equinox"]
    eqx.is_array_like = lambda x: isinstance(x, (int, float, complex, np.number))
    eqx.is_array = lambda x: False
    eqx.filter_jit = lambda f=None, **k: f if f is not None else (lambda g: g)

    pkg = types.ModuleType("quantax")
    pkg.__path__ = [os.path.join(REF, "quantax")]
    sys.modules["quantax"] = pkg

    gd = types.ModuleType("quantax.global_defs")

    class PARTICLE_TYPE(enum.Enum):
        spin = 0
        spinful_fermion = 1
        spinless_fermion = 2

    gd.PARTICLE_TYPE = PARTICLE_TYPE
    gd.get_default_dtype = lambda: np.float64
    gd.is_default_cpl = lambda: False
    sys.modules["quantax.global_defs"] = gd
    pkg.global_defs = gd

    sites = importlib.import_module("quantax.sites")  # the real, pure-NumPy reference code

    def get_sites():
        return sites.Sites._SITES

    gd.get_sites = get_sites
    gd.get_lattice = get_sites
    for sub in ["state", "sampler", "symmetry", "utils"]:
        sys.modules[f"quantax.{sub}"] = _InertModule(f"quantax.{sub}")
    operator = importlib.import_module("quantax.operator")
    return sites, operator


def _flatten_op_list(op_list):
    """op list -> (names, offsets, J, idx) arrays for npz storage."""
    names, J, idx, width = [], [], [], []
    for opstr, terms in op_list:
        names.append(opstr)
        width.append(len(terms))
        for t in terms:
            J.append(float(t[0]))
            idx.append([int(v) for v in t[1:]] + [-1] * (4 - len(t[1:])))
    return (np.array(names), np.array(width, dtype=np.int64), np.array(J, dtype=np.float64),
            np.array(idx, dtype=np.int64))


def main():
    import warnings

    warnings.simplefilter("ignore")
    sites, operator = _install_stubs()
    out = {}
    cases = {
        "chain8": lambda: sites.Chain(8),
        "square4": lambda: sites.Square(4, Nparticles=(8, 8)),
        "square6": lambda: sites.Square(6, Nparticles=(18, 18)),
        "square10": lambda: sites.Square(10, Nparticles=(50, 50)),
        "square16": lambda: sites.Square(16, Nparticles=(128, 128)),
        "triangular6": lambda: sites.Triangular(6, Nparticles=(18, 18)),
        "triangular12": lambda: sites.Triangular(12, Nparticles=(72, 72)),
    }
    for name, make in cases.items():
        sites.Sites._SITES = None
        lat = make()
        out[f"{name}/coord"] = lat.coord
        for n in (1, 2):
            out[f"{name}/nb{n}"] = np.asarray(lat.get_neighbor(n), dtype=np.int64)
        if name == "chain8":
            ops = {"ising_h1": operator.Ising(h=1.0), "ising_h0.5_J2": operator.Ising(h=0.5, J=2.0)}
        else:
            ops = {
                "heis": operator.Heisenberg(),
                "heis_msr": operator.Heisenberg(msr=True),
                "j1j2_msr": operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True),
            }
        for oname, op in ops.items():
            names, width, J, idx = _flatten_op_list(op.op_list)
            out[f"{name}/{oname}/names"] = names
            out[f"{name}/{oname}/width"] = width
            out[f"{name}/{oname}/J"] = J
            out[f"{name}/{oname}/idx"] = idx
    path = os.path.join(HERE, "ref_tables.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


def make_symm(sites):
    """Generators of Translation / Flip / Rotation from the reference's NumPy code
    (quantax/symmetry/translation.py:25-47, common_symmetries.py:104-205).  The group closure
    (_get_perm) is jnp code and cannot run here, so Symmetry.__init__ is replaced by a recorder."""
    for k in list(sys.modules):
        if k.startswith("quantax.symmetry"):
            del sys.modules[k]
    symmod = importlib.import_module("quantax.symmetry.symmetry")

    def fake_init(self, generator=None, sector=0, generator_sign=None, Z2_inversion=0, perm=None, character=None,
                  perm_sign=None):
        self.generator = None if generator is None else np.atleast_2d(np.asarray(generator))

    symmod.Symmetry.__init__ = fake_init
    tr = importlib.import_module("quantax.symmetry.translation")
    cs = importlib.import_module("quantax.symmetry.common_symmetries")
    out = {}
    cases = {"square4": lambda: sites.Square(4), "square6": lambda: sites.Square(6), "square10": lambda: sites.Square(10),
             "triangular6": lambda: sites.Triangular(6), "chain8": lambda: sites.Chain(8)}
    for name, mk in cases.items():
        sites.Sites._SITES = None
        lat = mk()
        nd = lat.ndim
        c = np.zeros(nd) if name.startswith("tri") else None
        out[f"{name}/trans"] = tr.Translation(np.eye(nd, dtype=int)).generator
        out[f"{name}/flip0"] = cs.Flip(0, center=c).generator
        if nd == 2:
            ang = np.pi / 3 if name.startswith("tri") else np.pi / 2
            out[f"{name}/rot"] = cs.Rotation(ang, center=c).generator
            out[f"{name}/flip1"] = cs.Flip(1, center=c).generator
    path = os.path.join(HERE, "ref_symm_generators.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    sys.path.insert(0, REF)
    main()
    make_symm(sys.modules["quantax.sites"])
