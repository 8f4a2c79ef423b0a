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
