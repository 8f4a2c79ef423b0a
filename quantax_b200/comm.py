"""Communicator of the C-ABI collectives (qtx_comm_*, csrc/comm.cu) for the torch host: one NCCL communicator per
process group, created from a unique id that rank 0 draws and ``torch.distributed`` broadcasts (the only thing
torch.distributed does for this path; a jax host would broadcast the 128 bytes by its own means or hand over its
ncclComm_t with ``qtx_comm_adopt``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .global_defs import device, world

_COMM = None


def communicator():
    """The process-wide qtx communicator (created on first use; ``None`` on a single process)."""
    global _COMM
    rank, P = world()
    if P == 1:
        return None
    if _COMM is None:
        import torch.distributed as dist

        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.call("qtx_comm_unique_id", C.cast(buf, C.c_void_p))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(device())
        dist.broadcast(ident, src=0)
        raw = bytes(ident.cpu().tolist())
        handle = C.c_void_p()
        _lib.call("qtx_comm_init", C.cast(C.pointer(handle), C.c_void_p), P, rank, C.cast(C.c_char_p(raw), C.c_void_p))
        _COMM = handle
    return _COMM


def minsr_solve_dist(A: torch.Tensor, b: torch.Tensor, rtol: float, atol: float, nslices: int, lanczos_steps: int,
                     refine_steps: int, workspace):
    """x = A^+ b with the rows of A sharded over the ranks: one call of qtx_minsr_solve_dist (all collectives inside
    the library).  Returns (x [Np] float64, info int32 [1])."""
    comm = communicator()
    nl, npar = A.shape
    dt = _lib.dtype_code(A.dtype)
    wsz = _lib.lib().qtx_minsr_solve_dist_workspace_size(comm, dt, nl, npar, int(nslices))
    if wsz == 0:
        raise _lib.QtxError(f"qtx_minsr_solve_dist_workspace_size failed: {_lib.lib().qtx_last_error().decode()}")
    ws = workspace("minsr_dist", wsz)
    x = torch.empty(npar, dtype=torch.float64, device=A.device)
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    _lib.call("qtx_minsr_solve_dist", comm, dt, _lib.ptr2d(A), nl, npar, A.stride(0), _lib.ptr(b.contiguous()),
              float(rtol), float(atol), int(nslices), int(lanczos_steps), int(refine_steps), _lib.ptr(x), _lib.ptr(info),
              _lib.ptr(ws), wsz, _lib.stream())
    return x, info
