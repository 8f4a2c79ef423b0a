"""world_size-2 gloo tests (CPU) of the host-side exchange logic of the distributed MinSR solve.
The dense kernels are replaced by CPU stand-ins defined HERE (test infrastructure); what is under
test is the layout / collective choreography of quantax_b200.optimizer.distributed_minnorm:
row-sharded -> column-sharded all-to-all with parameter-axis padding, Gram all-reduce, b all-gather,
step all-gather (quantax/optimizer/solver.py:134-147)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import solver as osolver


class CpuOps:
    @staticmethod
    def gram(A):
        return A @ A.T

    @staticmethod
    def pinv_eig_solve(T, b, rtol, atol):
        y = osolver.minsr_pinv_eig(T.numpy(), b.numpy(), rtol, atol)
        return torch.from_numpy(y), torch.zeros(1, dtype=torch.int32)

    @staticmethod
    def matvec_t(A, y):
        return A.T @ y


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ns, npar, out, shift=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quantax_b200.optimizer import distributed_minnorm

        rng = np.random.default_rng(0)
        A = rng.standard_normal((ns, npar))
        A -= A.mean(axis=0, keepdims=True)
        b = rng.standard_normal(ns)
        nl = ns // world
        Al = torch.from_numpy(A[rank * nl:(rank + 1) * nl].copy())
        bl = torch.from_numpy(b[rank * nl:(rank + 1) * nl].copy())
        if shift:
            # pluggable T-solver (diagonal-shift Cholesky on the GPU; a dense solve stands in here)
            def tsolve(T, bfull):
                lam = 1e-3 * torch.trace(T) + 1e-4
                return torch.linalg.solve(T + lam * torch.eye(T.shape[0], dtype=T.dtype), bfull), None

            x, _ = distributed_minnorm(Al, bl, None, 0.0, CpuOps, tsolve)
            xo = osolver.minnorm_shift_eig(A, b, 1e-3, 1e-4)
        else:
            x, _ = distributed_minnorm(Al, bl, None, 0.0, CpuOps)
            xo = osolver.minnorm_pinv_eig(A, b)
        err = float(np.linalg.norm(x.numpy() - xo) / np.linalg.norm(xo))
        # every rank must hold the same full step
        gathered = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(gathered, x)
        same = all(torch.equal(g, gathered[0]) for g in gathered)
        if rank == 0:
            out.put((err, same, tuple(x.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ns,npar", [(16, 40), (16, 41), (8, 9)])
def test_distributed_minnorm_two_ranks(ns, npar):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ns, npar, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err, same, shape = q.get(timeout=10)
    assert shape == (npar,)
    assert same
    assert err < 1e-8, err


def test_distributed_minnorm_with_shift_solver_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 16, 41, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err, same, shape = q.get(timeout=10)
    assert shape == (41,) and same and err < 1e-9, err


def _gather_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from quantax_b200.optimizer import _gather_rows_as_columns

        k, nl = 3, 4
        full = torch.arange(k * nl * world, dtype=torch.float64).reshape(k, nl * world)  # columns = global samples
        mine = full[:, rank * nl:(rank + 1) * nl].contiguous()
        got = _gather_rows_as_columns(mine)
        if rank == 0:
            out.put(bool(torch.equal(got, full)))
    finally:
        dist.destroy_process_group()


def test_gather_rows_as_columns_restores_the_global_sample_order():
    """lstsq_pinv_eig(tol_snr > 0) on several ranks: (A V)^T [np, nl] per rank -> [np, Ns] with rank-major columns."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10)


def test_sampler_rejects_indivisible_sample_count(monkeypatch):
    """sampler.py:29-33: nsamples must be a multiple of the device count."""
    from quantax_b200 import global_defs, sampler, sites

    sites.Sites._SITES = None
    sites.Chain(8)
    monkeypatch.setattr(sampler, "world", lambda: (1, 2))

    class FakeState:
        Nsites = Nmodes = 8

    with pytest.raises(ValueError):
        sampler.Sampler(FakeState(), 7)
    s = sampler.Sampler(FakeState(), 8)
    assert s.nlocal == 4 and s._rank == 1


# ---- rank split of the eigendecomposition-free pseudo-inverse (optimizer.pinv_rational_solve) ---------
class _FakeLib:
    """CPU stand-ins for the three C entry points, taking the arguments in the order of include/qtx_b200.h, built on
    the oracle restatement (test infrastructure): what is under test is the Python glue -- shift masks per rank,
    all-gather of the double-double partial sums, rank-ordered final sum."""

    QtxError = RuntimeError

    def __init__(self):
        from oracle import pinv_rational as pr

        self.pr = pr
        self.calls = []

    def lib(self):
        class L:
            @staticmethod
            def qtx_pinv_rational_workspace_size(n):
                return 64

            @staticmethod
            def qtx_pinv_ldlt_workspace_size(n, nshifts):
                assert 1 <= nshifts <= 3
                return 64 * nshifts

        return L()

    @staticmethod
    def ptr(t):
        return t

    @staticmethod
    def stream():
        return 0

    def call(self, name, *a):
        pr = self.pr
        self.calls.append(name)
        if name in ("qtx_sym_absmax_eig", "qtx_sym_absmax_eig_ws"):
            T, n, first, steps, lam, ws, wsz = a[:7]
            assert T.shape == (n, n) and 0 <= first < steps
            lam[0] = pr.abs_max_eigenvalue(T.numpy(), steps=steps)
        elif name in ("qtx_pinv_rational_partial", "qtx_pinv_ldlt_partial"):
            T, n, b, rtol, atol, lam, mask, refine, ydd, accumulate, info, ws, wsz, stream = a
            assert ydd.shape == (2, n) and accumulate == 0 and 0 < mask < 8
            which = [k for k in range(3) if (mask >> k) & 1]
            yh, yl = pr.pinv_rational_partial(T.numpy(), b.numpy(), None if rtol < 0 else rtol, atol, float(lam[0]), which,
                                              refine)
            ydd[0], ydd[1] = torch.from_numpy(yh), torch.from_numpy(yl)
            info[0] = 0
        elif name == "qtx_dd_sum_scale":
            ydd, count, n, scale, y, stream = a
            parts = ydd.reshape(count, 2, n).numpy()
            y.copy_(torch.from_numpy(pr.dd_sum_scale([(p[0], p[1]) for p in parts], scale)))
        else:
            raise AssertionError(name)


def _rational_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pinv_rational as pr
        from quantax_b200 import optimizer as qopt

        fake = _FakeLib()
        qopt._lib = fake

        class WS:
            @staticmethod
            def get(key, nbytes):
                return torch.empty(nbytes, dtype=torch.uint8)

        qopt._WS = WS
        rng = np.random.default_rng(3)  # the same replicated T, b on every rank
        B = rng.standard_normal((24, 90)) * np.exp(-6 * rng.random((1, 90)))
        B -= B.mean(axis=0, keepdims=True)
        T, b = B @ B.T, rng.standard_normal(24)
        y, info = qopt.pinv_rational_solve(torch.from_numpy(T), torch.from_numpy(b), 1e-9, 0.0, replicated=True)
        y_single = pr.pinv_rational_solve(T, b, rtol=1e-9)
        err = float(np.linalg.norm(B.T @ (y.numpy() - y_single)) / np.linalg.norm(B.T @ y_single))
        gathered = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(gathered, y)
        same = all(torch.equal(g, gathered[0]) for g in gathered)
        nparts = fake.calls.count("qtx_pinv_ldlt_partial")  # the default route: the library's own LDL^T kernels
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([nparts]))
        if rank == 0:
            out.put((err, same, [int(c.item()) for c in counts], int(info.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_rational_pseudo_inverse_rank_split(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rational_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    err, same, calls, info = q.get(timeout=10)
    assert same and info == 0 and err < 1e-12, err
    assert calls == ([1, 1] if world == 2 else [1, 1, 1, 0])  # every rank with a non-empty mask makes one call
