"""The digit (Ozaki) arithmetic of the tensor-core Gram kernel, restated in oracle/gram_digits.py, against an
extended-precision product: pins the accuracy figures of DESIGN.md 4.1 without a GPU.  The GPU kernel is compared
bit for bit with the same restatement in tests/test_zz_solver_variants_gpu.py."""
import numpy as np
import pytest

from oracle import gram_digits as gd

pytestmark = pytest.mark.skipif(np.finfo(np.longdouble).eps > 2e-19, reason="needs 80-bit long double")


def _matrix(ns, npar, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((ns, npar)) * np.exp(3 * rng.standard_normal((ns, 1)))  # rows of very different size
    A[:, ::7] *= 1e-3  # wide dynamic range inside a row
    A[2] = 0.0
    return A


def _rel_err(T, A):
    ref = gd.exact_gram(A)
    nrm = np.linalg.norm(A, axis=1)
    den = np.maximum(np.outer(nrm, nrm), 1e-300)
    return float((np.abs(T - ref) / den).max())


@pytest.mark.parametrize("s", [1, 2, 4, 7, 8])
def test_digits_reconstruct_the_row_to_seven_bits_per_digit(s):
    A = _matrix(40, 500, 0)
    Q, scale = gd.split_digits(A, s)
    assert Q.dtype == np.int8 and np.abs(Q.astype(np.int32)).max() <= 64
    # after s digits the residual is at most half a unit of the last digit: 2^(-7 s) of the row scale
    assert (np.abs(A - gd.digits_value(Q, scale)) <= scale[:, None] * 2.0 ** (-7 * s)).all()
    assert np.all(scale[np.abs(A).max(axis=1) > 0] > np.abs(A).max(axis=1)[np.abs(A).max(axis=1) > 0])
    assert not Q[:, 2].any()


# error relative to |a_i||a_j| at 3000 columns; DESIGN.md 4.1 quotes 4.8e-16 / 1.7e-15 / 7e-12 at 40400 columns
# (the truncation error falls like 1/sqrt(K) against the norms) and 7.9e-15 for cuBLAS DGEMM
@pytest.mark.parametrize("s,bound", [(8, 1.5e-15), (7, 2e-14), (6, 2e-11), (4, 3e-7), (2, 5e-3)])
def test_gram_accuracy_per_number_of_digits(s, bound):
    A = _matrix(96, 3000, 1)
    T = gd.gram(A, s)
    assert _rel_err(T, A) <= bound
    assert np.array_equal(T, T.T)
    assert not T[2].any() and not T[:, 2].any()


def test_default_digits_beat_float64_gemm_at_config_b_width():
    # Np = 40400 columns (config B), few rows: 7 digits are at least as accurate as a float64 BLAS product
    A = _matrix(24, 40400, 2)
    e7 = _rel_err(gd.gram(A), A)
    assert e7 <= 4e-15
    assert gd.default_slices(np.float64) == 7 and gd.default_slices(np.float32) == 4


def test_float32_input_is_exact_with_four_digits_up_to_the_dropped_levels():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((50, 700)).astype(np.float32)
    assert _rel_err(gd.gram(A), A.astype(np.float64)) <= 3e-7
    # 24-bit significands fit in four 7-bit digits only for |x| close to the row maximum: 8 digits are exact
    Q, scale = gd.split_digits(A, 8)
    assert np.array_equal(gd.digits_value(Q, scale), A.astype(np.float64))


def test_k_chunks_accumulate_and_accumulate_flag():
    s = 8
    kmax = gd.chunk_columns(1 << 40, s)  # the largest chunk the exact int32 accumulation allows
    assert kmax % 64 == 0 and kmax * s * 4096 < 2 ** 31 <= (kmax + 64) * s * 4096 + s * 4096 * 64
    kc = gd.chunk_columns(10 ** 6, s)  # equal chunks: 16 of 62 528 columns, not 15 of 65 472 and a short one
    assert kc % 64 == 0 and kc <= kmax and -(-10 ** 6 // kc) == -(-10 ** 6 // kmax) and 10 ** 6 - 15 * kc > kc // 2
    A = _matrix(6, kmax + 777, 4)  # two K chunks
    T = gd.gram(A, s)
    assert _rel_err(T, A) <= 2e-15
    # T_accum: second call adds to the first (qtx_gram T_accum = 1)
    T2 = gd.gram(A[:, 300:], s, T=gd.gram(A[:, :300], s))
    assert _rel_err(T2, A) <= 2e-15


def test_non_finite_row_poisons_only_its_row_and_column():
    A = _matrix(8, 100, 5)
    A[5, 17] = np.inf
    with np.errstate(invalid="ignore"):
        T = gd.gram(A, 7)
    bad = np.isnan(T)
    assert bad[5].all() and bad[:, 5].all() and not np.delete(np.delete(bad, 5, 0), 5, 1).any()
