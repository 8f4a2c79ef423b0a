"""NumPy restatement of the ARITHMETIC of the tensor-core Gram kernel -- TEST INFRASTRUCTURE ONLY.

The reference computes ``T = Adag.conj().T @ Adag`` in float64 (quantax/optimizer/solver.py:139); the product
computes the same matrix on the int8 tensor pipe with an error-free digit scheme
(quantax_b200/csrc/gram_tc.cu ``gram_split_kernel``, quantax_b200/csrc/gram_tc2.cu epilogue).  This module follows
the kernel step by step so that

  * the accuracy claims of DESIGN.md 4.1 (error per number of digits) are checked on the CPU against an extended
    precision product, and
  * the GPU kernel can be compared BIT FOR BIT with this emulation (every floating-point operation of the kernel is
    either exact -- powers of two, integers below 2^53 -- or a single rounded addition whose order is restated here).

Scheme: row i is scaled by 2^-e_i (|x| 2^-e_i < 1), and cut into s signed digits q_a, |q_a| <= 64,
``x 2^-e = sum_a q_a 2^(-7a+1)`` (a = 1..s); digit products of equal level d = a + b are summed exactly in int32;
levels d > s + 1 are dropped; the epilogue adds ``2^(-7d+2) * level_d`` from the highest kept level down.
"""
import numpy as np

K_LEVELS_PER_PASS = 4  # TMEM holds four 128-column level accumulators (gram_tc_common.cuh kLevelsPerPass)


def default_slices(dtype) -> int:
    """gram_tc.cu default_slices: 7 digits for float64 input, 4 for float32."""
    return 7 if np.dtype(dtype) == np.float64 else 4


def chunk_columns(npar: int, s: int) -> int:
    """K chunk keeping the int32 level sums exact (gram_tc.cu gram_tc_sizes): K * s * 4096 < 2^31."""
    kmax = (1 << 31) // (4096 * s) - 1
    kmax = kmax // 64 * 64
    nchunks = (npar + kmax - 1) // kmax  # equal chunks, rounded up to the 64-column padding unit
    kc = ((npar + nchunks - 1) // nchunks + 63) // 64 * 64
    return min(kc, kmax)


def split_digits(A: np.ndarray, s: int):
    """(Q int8 [s, ns, np], rowscale float64 [ns]) as gram_split_kernel writes them."""
    A = np.asarray(A)
    ns, npar = A.shape
    mx = np.abs(A.astype(np.float64)).max(axis=1) if npar else np.zeros(ns)
    _, e = np.frexp(mx)  # mx = m 2^e, m in [0.5, 1)
    pos = mx > 0
    scale = np.where(pos, np.ldexp(1.0, e), 1.0)
    inv = np.where(pos, np.ldexp(1.0, -e), 0.0)
    bad = ~np.isfinite(mx)
    scale = np.where(bad, np.nan, scale)
    inv = np.where(bad, 0.0, inv)
    with np.errstate(invalid="ignore"):
        res = A.astype(np.float64) * inv[:, None] * 64.0
    res[bad] = 0.0  # the kernel's (int)NaN is 0; the NaN row scale poisons the whole row of T instead
    Q = np.zeros((s, ns, npar), dtype=np.int8)
    for a in range(s):
        q = np.rint(res)  # round half to even, like CUDA rint()
        assert np.abs(q).max(initial=0.0) <= 64
        Q[a] = q.astype(np.int8)
        res = (res - q) * 128.0
    return Q, scale


def digits_value(Q: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """The float64 matrix the digits stand for: scale * sum_a q_a 2^(-7a+1)."""
    s = Q.shape[0]
    v = np.zeros(Q.shape[1:], dtype=np.float64)
    for a in range(s, 0, -1):
        v += np.ldexp(Q[a - 1].astype(np.float64), -7 * a + 1)
    return v * np.where(np.isfinite(scale), scale, 0.0)[:, None]


def level_sums(Q: np.ndarray, d: int) -> np.ndarray:
    """Exact integer matrix sum_{a+b=d, 1<=a,b<=s} Q_a Q_b^T (the content of one TMEM accumulator)."""
    s, ns, npar = Q.shape
    assert npar * s * 4096 < 2 ** 31, "the int32 accumulator of the kernel would overflow: chunk K first"
    out = np.zeros((ns, ns), dtype=np.float64)
    for a in range(max(1, d - s), min(d - 1, s) + 1):
        b = d - a
        # integers below 2^31: the float64 product is exact
        out += Q[a - 1].astype(np.float64) @ Q[b - 1].astype(np.float64).T
    assert np.abs(out).max(initial=0.0) < 2 ** 31
    return out


def gram(A: np.ndarray, nslices: int = 0, T: np.ndarray = None) -> np.ndarray:
    """T (+)= A A^T with the operation order of the kernel: K chunks outermost, passes of four levels, levels
    summed from the highest down, one rounded addition into T per pass."""
    A = np.asarray(A)
    ns, npar = A.shape
    s = nslices if nslices > 0 else default_slices(A.dtype)
    Q, scale = split_digits(A, s)
    pair_scale = np.outer(scale, scale)
    kc = chunk_columns(npar, s)
    npasses = (s + K_LEVELS_PER_PASS - 1) // K_LEVELS_PER_PASS
    accumulate = T is not None
    if T is None:
        T = np.zeros((ns, ns), dtype=np.float64)
    first = True
    for k0 in range(0, npar, kc):
        Qc = Q[:, :, k0:k0 + kc]
        for p in range(npasses):
            d_lo = 2 + p * K_LEVELS_PER_PASS
            d_hi = min(d_lo + K_LEVELS_PER_PASS - 1, s + 1)
            acc = np.zeros((ns, ns), dtype=np.float64)
            for d in range(d_hi, d_lo - 1, -1):
                acc += np.ldexp(level_sums(Qc, d), -7 * d + 2)
            val = acc * pair_scale
            T = T + val if (accumulate or not first) else val
            first = False
    return T


def exact_gram(A: np.ndarray) -> np.ndarray:
    """Extended-precision product (x87 80-bit where numpy has it) used as the accuracy reference."""
    L = np.asarray(A, dtype=np.longdouble)
    return L @ L.T
