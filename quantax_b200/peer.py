"""Fused Gram + exchange over NVLink peer memory for the distributed MinSR solve.

The reference sums the partial Gram matrices of the column shards with a GSPMD all-reduce
(quantax/optimizer/solver.py:134-139).  Here the Gram kernel itself stores every finished tile into a
staging area on every peer while the tensor cores work on the next tile (csrc/gram_tc2.cu, PUSH variant),
a flag per rank announces completion, and a reduce kernel sums the P partials in rank order, so that the
matrix handed to the replicated ``eigh`` is bit-identical on all ranks (csrc/peer.cu).  ``torch.distributed``
is used once, at construction, to exchange the CUDA IPC handles of the staging areas; no collective is on
the data path.  Selected by ``QTX_GRAM_P2P=1`` (see optimizer._CudaOps.gram_allreduce).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _lib

_FLAG_BYTES = 256  # flags[QTX_MAX_PEERS] uint64 at the start of the allocation, staging area behind it


class PeerGram:
    """Staging areas and flags of one process group for Gram matrices of a fixed size ``ns``."""

    def __init__(self, ns: int, group=None, timeout_s: float = float(os.environ.get("QTX_P2P_TIMEOUT_S", "60"))):
        import torch.distributed as dist

        self.group = group
        self.rank, self.P = dist.get_rank(group), dist.get_world_size(group)
        if self.P > 8:
            raise _lib.QtxError("PeerGram supports up to 8 ranks (one NVSwitch node)")
        self.ns, self.timeout_s, self.epoch = int(ns), float(timeout_s), 0
        self._slot_bytes = self.ns * self.ns * 8
        base = C.c_void_p()
        _lib.call("qtx_peer_alloc", _FLAG_BYTES + self.P * self._slot_bytes, C.byref(base))
        self._base = base.value
        handle = (C.c_char * 64)()
        _lib.call("qtx_peer_export", self._base, handle)
        everyone = [None] * self.P
        dist.all_gather_object(everyone, (bytes(handle), self.ns, os.getpid()), group=group)
        self._bases = []
        for q, (h, ns_q, _pid) in enumerate(everyone):
            if ns_q != self.ns:
                raise _lib.QtxError("PeerGram: ranks disagree on the Gram size")
            if q == self.rank:
                self._bases.append(self._base)
                continue
            ptr = C.c_void_p()
            _lib.call("qtx_peer_open", (C.c_char * 64).from_buffer_copy(h), C.byref(ptr))
            self._bases.append(ptr.value)
        vp = C.c_void_p * self.P
        # where rank q expects MY partial / the flag arrays of all ranks / the partials as seen from here
        self._slots = vp(*[b + _FLAG_BYTES + self.rank * self._slot_bytes for b in self._bases])
        self._flags = vp(*self._bases)
        self._mine = [self._base + _FLAG_BYTES + q * self._slot_bytes for q in range(self.P)]
        dist.barrier(group=group)  # every rank has mapped every staging area before the first push

    def gram_allreduce(self, A: torch.Tensor, nslices=None) -> torch.Tensor:
        """sum over ranks of A_r A_r^T for the local column shard A_r [ns, np_r]; float64 [ns, ns], the same bits
        on every rank."""
        from .optimizer import DEFAULT_NSLICES, _WS

        ns, npar = A.shape
        if ns != self.ns:
            raise _lib.QtxError(f"PeerGram was built for ns = {self.ns}, got {ns}")
        if nslices is None:
            nslices = DEFAULT_NSLICES
        self.epoch += 1
        T = torch.empty((ns, ns), dtype=torch.float64, device=A.device)
        dt = _lib.dtype_code(A.dtype)
        wsz = _lib.lib().qtx_gram_workspace_size(dt, ns, npar, nslices)
        ws = _WS.get("gram", wsz)
        st = _lib.stream()
        _lib.call("qtx_gram_push", dt, _lib.ptr2d(A), ns, npar, A.stride(0), int(nslices), _lib.ptr(T), self.P, self.rank,
                  self._slots, _lib.ptr(ws), wsz, st)
        _lib.call("qtx_peer_signal", self._flags, self.P, self.rank, self.epoch, st)
        vp = C.c_void_p * self.P
        partials = vp(*[T.data_ptr() if q == self.rank else self._mine[q] for q in range(self.P)])
        _lib.call("qtx_gram_reduce", partials, self.P, ns, _lib.ptr(T), self._base, self.epoch, self.timeout_s, st)
        return T

    def close(self) -> None:
        if self._base is None:
            return
        torch.cuda.synchronize()
        for q, b in enumerate(self._bases):
            if q != self.rank:
                _lib.call("qtx_peer_close", b)
        import torch.distributed as dist

        if dist.is_initialized():
            dist.barrier(group=self.group)  # nobody frees memory a peer still has mapped and may be reading
        _lib.call("qtx_peer_free", self._base)
        self._base, self._bases = None, []


_CACHE = {}


def peer_gram(ns: int) -> PeerGram:
    """One PeerGram per Gram size for the default process group (allocated on first use, kept for the run)."""
    pg = _CACHE.get(ns)
    if pg is None:
        pg = _CACHE[ns] = PeerGram(ns)
    return pg


def release_all() -> None:
    for pg in list(_CACHE.values()):
        pg.close()
    _CACHE.clear()
