"""Tensor-core Gram kernel (tcgen05 int8 slices + TMA) against float64 references."""
import numpy as np
import pytest
import torch

from tests.gpu_util import to_np

pytestmark = pytest.mark.gpu


def _rel_err(T, A):
    ref = A @ A.T
    nrm = np.linalg.norm(A, axis=1)
    den = np.maximum(np.outer(nrm, nrm), 1e-300)
    return (np.abs(T - ref) / den).max()


@pytest.mark.parametrize("ns,npar", [(128, 64), (100, 333), (257, 1000), (1, 7), (130, 4099), (64, 70001)])
@pytest.mark.parametrize("nslices,tol", [(8, 2e-13), (0, 2e-12), (7, 1e-11), (4, 2e-6), (2, 5e-2)])
def test_gram_tc_f64(ns, npar, nslices, tol):
    from quantax_b200.optimizer import gram

    rng = np.random.default_rng(ns * 1000 + npar)
    A = rng.standard_normal((ns, npar)) * np.exp(3 * rng.standard_normal((ns, 1)))
    A[:, ::7] *= 1e-3  # wide dynamic range inside rows
    if ns > 2:
        A[2] = 0.0  # an all-zero row must give exact zeros
    T = to_np(gram(torch.from_numpy(A).cuda(), nslices=nslices))
    err = _rel_err(T, A)
    assert err <= tol * np.sqrt(npar), err
    assert np.array_equal(T, T.T)
    if ns > 2:
        assert not T[2].any() and not T[:, 2].any()


def test_gram_tc_f32_input_and_accumulate():
    from quantax_b200.optimizer import gram

    rng = np.random.default_rng(5)
    A = rng.standard_normal((200, 500)).astype(np.float32)
    At = torch.from_numpy(A).cuda()
    T = gram(At, nslices=0)  # float32 input: 4 slices by default
    ref = A.astype(np.float64) @ A.astype(np.float64).T
    assert np.abs(to_np(T) - ref).max() <= 1e-5 * np.abs(ref).max()
    T8 = gram(At, nslices=8)
    assert np.abs(to_np(T8) - ref).max() <= 1e-13 * np.abs(ref).max()
    # accumulate over two column shards == full product (the distributed MinSR sum, solver.py:139)
    T2 = gram(At[:, :300].contiguous(), nslices=8)
    gram(At[:, 300:].contiguous(), out=T2, nslices=8, accumulate=True)
    assert np.abs(to_np(T2) - ref).max() <= 1e-13 * np.abs(ref).max()
    # row-padded input (leading dimension > np)
    T3 = gram(At[:, :301], nslices=8)
    ref3 = A[:, :301].astype(np.float64) @ A[:, :301].astype(np.float64).T
    assert np.abs(to_np(T3) - ref3).max() <= 1e-13 * np.abs(ref3).max()


def test_gram_tc_full_size_against_cublas():
    """BASELINE config B shape (4096 x 40400 float64): compare with torch/cuBLAS float64 and check the
    size-independent properties (symmetry, PSD-ness via a random quadratic form, trace identity)."""
    from quantax_b200.optimizer import gram

    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((4096, 40400), dtype=torch.float64, device="cuda", generator=g) / 200
    A -= A.mean(dim=0, keepdim=True)
    T = gram(A, nslices=0)
    ref = A @ A.T
    nrm = A.norm(dim=1)
    err = ((T - ref).abs() / torch.outer(nrm, nrm)).max().item()
    assert err < 1e-12, err
    assert torch.equal(T, T.T)
    assert abs(T.diagonal().sum().item() - (A * A).sum().item()) < 1e-10 * (A * A).sum().item()
    x = torch.randn(4096, dtype=torch.float64, device="cuda", generator=g)
    assert (x @ (T @ x)).item() >= -1e-9
    assert (T.sum(dim=0).abs().max() / T.abs().max()).item() < 1e-9  # centred columns: T 1 = 0
