"""CPU oracle for the quantax VMC hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm
(ChenAo-Phys/quantax v0.2.1) for the path Metropolis sweep -> Operator.Oloc ->
Variational.jacobian -> SR/MinSR solve.  Every function cites the reference
file:line it follows.

Rules (enforced by tests/test_host_cpu.py::test_product_never_imports_oracle):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * the product (``quantax_b200``) never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * lattice bond tables and operator lists are checked against the reference's
    own NumPy/Python code imported in the build container
    (tests/golden/make_golden.py -> tests/golden/*.npz);
  * Hamiltonian matrix elements / connected-configuration enumeration are
    checked against the ED energies the reference prints in its tutorials;
  * local-update psi == direct psi and Oloc(local updates) == Oloc(direct)
    follow the reference's notebook asserts;
  * enumeration / compaction / Oloc reduction, the solver formulas, group
    closure, neighbour tables, phase kernels, final activations, proposals and
    accept/reject are checked against outputs of the reference's OWN modules
    executed under a NumPy stand-in for jax (tests/golden/minijax.py,
    tests/golden/make_golden_hotpath.py -> tests/golden/ref_hotpath.npz,
    tests/test_golden_hotpath_cpu.py);
  * ``gram_digits`` and ``pinv_rational`` restate the ARITHMETIC of two product kernels (the int8 digit scheme of
    the tensor-core Gram, the eigendecomposition-free pseudo-inverse) so that their accuracy claims are checked on
    the CPU; the reference formulas they reproduce are solver.py:139 and solver.py:94-111;
  * everything that lives in un-vendored third-party code (jax PRNG streams,
    ``jax.nn.gelu`` form, ``equinox.nn.Conv`` padding semantics,
    ``ravel_pytree`` leaf order, ``eigh``) is restated from its published
    behaviour: **parity unpinned** for those pieces (jax>=0.6.1 and
    equinox>=0.11.4 are not installable here).
"""
