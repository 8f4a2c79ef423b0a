"""Stochastic reconfiguration (SR / MinSR) and its linear solvers.

Mirrors quantax/optimizer/sr.py:17-195 (``QNGD``, ``SR``) and quantax/optimizer/solver.py:94-294
(``auto_pinv_eig``, ``minnorm_pinv_eig``, ``lstsq_pinv_eig``, ``minsr_pinv_eig``).  ``MinSR`` is
the pre-0.2 name of ``SR`` with the min-norm solver forced.

Data-parallel layout (one process per GPU): each rank owns ``Ns / P`` rows of the Jacobian.
Cross-rank traffic is exactly the reference's implicit GSPMD traffic made explicit:
  * all-gather of the local energies (Ns float64),
  * all-reduce of the Jacobian column means ([Np]),
  * row-sharded -> column-sharded exchange of Obar (all-to-all, solver.py:134-137),
  * all-reduce of the partial Gram matrices (solver.py:139),
  * all-gather of the column shards of the step (solver.py:146).
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch

from . import _lib
from .global_defs import device, get_default_dtype, get_real_dtype, world
from .state import VS_TYPE, Variational


def _dist():
    import torch.distributed as dist

    return dist


class _Workspaces:
    def __init__(self):
        self._bufs = {}

    def get(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device())
            self._bufs[key] = buf
        return buf


_WS = _Workspaces()

# Gram algorithm: 0 = tcgen05 int8-sliced tensor-core kernel with the dtype's default slice count,
# k > 0 = k slices, -1 = FP64 FMA cross-check kernel.  QTX_GRAM_NSLICES overrides (dev knob).
DEFAULT_NSLICES = int(os.environ.get("QTX_GRAM_NSLICES", "0"))


def model_gram_nslices(state) -> Optional[int]:
    """Digit count of the Gram for the default solver of ``state``: 5 for float32 models (their Jacobian is float32
    data in a float64 container), the dtype default (7 for float64) otherwise; QTX_GRAM_NSLICES overrides."""
    if DEFAULT_NSLICES != 0:
        return None
    mdt = getattr(getattr(state, "model", None), "dtype", None)
    return 5 if mdt == torch.float32 else None


def gram_nslices_for(dtype) -> int:
    """Digit count qtx_gram uses for an input of ``dtype`` (csrc/gram_tc.cu default_slices unless overridden)."""
    if DEFAULT_NSLICES > 0:
        return DEFAULT_NSLICES
    return 7 if dtype == torch.float64 else 4


# Per-kernel timing inside a running step (bench.py): when PHASE_EVENTS is a dict, the dense building blocks record
# a CUDA-event pair around their launches on the current stream; nothing is recorded (and nothing costs) otherwise.
PHASE_EVENTS = None


def _phase_tic(name: str):
    if PHASE_EVENTS is None:
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    return name, e0, e1


def _phase_toc(t) -> None:
    if t is not None and PHASE_EVENTS is not None:
        t[2].record()
        PHASE_EVENTS.setdefault(t[0], []).append((t[1], t[2]))


# ---- dense building blocks -----------------------------------------------------------------------
def gram(A: torch.Tensor, out: Optional[torch.Tensor] = None, nslices: Optional[int] = None, accumulate: bool = False):
    """T = A A^T (float64 [ns, ns]) on the tensor cores (qtx_gram)."""
    if nslices is None:
        nslices = DEFAULT_NSLICES
    ns, npar = A.shape
    if out is None:
        out = torch.empty((ns, ns), dtype=torch.float64, device=A.device)
    dt = _lib.dtype_code(A.dtype)
    wsz = _lib.lib().qtx_gram_workspace_size(dt, ns, npar, nslices)
    ws = _WS.get("gram", wsz)
    t = _phase_tic("gram")
    _lib.call("qtx_gram", dt, _lib.ptr2d(A), ns, npar, A.stride(0), int(nslices), _lib.ptr(out), int(accumulate),
              _lib.ptr(ws), wsz, _lib.stream())
    _phase_toc(t)
    return out


def _eig_workspace(n: int):
    wsz = _lib.lib().qtx_pinv_eig_workspace_size(n)
    if wsz == 0:
        raise _lib.QtxError(f"qtx_pinv_eig_workspace_size failed: {_lib.lib().qtx_last_error().decode()}")
    return _WS.get("eig", wsz), wsz


# How the soft pseudo-inverse y = f(T) b is evaluated when no eigen-quantity is asked for:
#   "ldlt"     (default) three complex-symmetric shifted solves by the library's OWN blocked LDL^T kernels
#              (csrc/zldlt.cu, qtx_pinv_ldlt_partial), refined in double-double: the same function of T by partial
#              fractions, no eigendecomposition, no cuSOLVER call on the step
#   "rational" the same identity with cuSOLVER's complex LU (qtx_pinv_rational_partial; cross-check)
#   "eigh"     cuSOLVER syevd + the pseudo-inverse epilogue (qtx_pinv_eig_solve; cross-check, and what the
#              eigen-quantity callers -- SNR damping, minsr_pinv_eig / pinvh_solve, rtol = atol = 0 -- keep using)
PINV_METHOD = os.environ.get("QTX_PINV", "ldlt")
LANCZOS_STEPS = max(1, min(1024, int(os.environ.get("QTX_LANCZOS_STEPS", "512"))))  # upper bound of the adaptive run
REFINE_STEPS = int(os.environ.get("QTX_PINV_REFINE", "2"))  # corrections 2e-5 -> 4e-10 -> 7e-15 (measured)
# two consecutive stages (k, 2k steps) of the Lanczos run must agree to this: the error of an extreme Ritz value after
# 2k steps is about the square of its error after k steps, so the accepted value is good to ~1e-14; max|lambda| only
# enters through the cut-off c = rtol max|lambda| + atol, to which f is insensitive away from the cut-off
LANCZOS_AGREE = 1e-7


def _shift_count(mask: int) -> int:
    return max(1, bin(int(mask)).count("1"))


def _pinv_workspace(n: int, method: str, nshifts: int):
    """(buffer, bytes) of the shifted-solve workspace of ``method`` ("ldlt": own kernels, "rational": cuSOLVER LU)."""
    if method == "ldlt":
        wsz = _lib.lib().qtx_pinv_ldlt_workspace_size(n, nshifts)
        what = "qtx_pinv_ldlt_workspace_size"
    else:
        wsz = _lib.lib().qtx_pinv_rational_workspace_size(n)
        what = "qtx_pinv_rational_workspace_size"
    if wsz == 0:
        raise _lib.QtxError(f"{what} failed: {_lib.lib().qtx_last_error().decode()}")
    return _WS.get("pinv_" + method, wsz), wsz


def sym_absmax_eig(T: torch.Tensor, steps: Optional[int] = None, method: Optional[str] = None,
                   nshifts: int = 3) -> torch.Tensor:
    """max|lambda| of a symmetric float64 matrix as a device scalar [1] (Lanczos, qtx_sym_absmax_eig[_ws]).  With
    ``steps=None`` the recurrence is continued to 32, 64, 128, ... steps (at most QTX_LANCZOS_STEPS) until two
    consecutive values agree to LANCZOS_AGREE -- the error after 2k steps is about the square of the error after k --
    which costs one scalar read-back per stage; an explicit ``steps`` runs exactly that many without synchronising."""
    n = T.shape[0]
    method = "rational" if method is None and PINV_METHOD == "rational" else ("ldlt" if method is None else method)
    ws, wsz = _pinv_workspace(n, method, nshifts)
    lam = torch.empty(1, dtype=torch.float64, device=T.device)

    def run(first, upto):
        if method == "ldlt":
            _lib.call("qtx_sym_absmax_eig_ws", _lib.ptr(T), n, first, int(upto), _lib.ptr(lam), _lib.ptr(ws), wsz,
                      int(nshifts), _lib.stream())
        else:
            _lib.call("qtx_sym_absmax_eig", _lib.ptr(T), n, first, int(upto), _lib.ptr(lam), _lib.ptr(ws), wsz,
                      _lib.stream())

    if steps is not None:
        run(0, steps)
        return lam
    done, prev = 0, None
    for upto in lanczos_stages(n, LANCZOS_STEPS):
        run(done, upto)
        done = upto
        cur = float(lam.item())
        if prev is not None and abs(cur - prev) <= LANCZOS_AGREE * abs(cur):
            break
        prev = cur
    return lam


def lanczos_stages(n: int, max_steps: int):
    """Step counts at which the adaptive Lanczos run is evaluated: 32, 64, 128, ... capped by n and max_steps."""
    out, k = [], 32
    cap = max(1, min(n, max_steps))
    while k < cap:
        out.append(k)
        k *= 2
    out.append(cap)
    return out


def rational_shift_masks(P: int):
    """Which of the three shifts every rank of a replicated solve takes (bit k = shift k)."""
    if P <= 1:
        return [7]
    if P == 2:
        return [0b101, 0b010]
    return [1, 2, 4] + [0] * (P - 3)


def pinv_rational_solve(T: torch.Tensor, b: torch.Tensor, rtol: Optional[float], atol: float, replicated: bool = False,
                        method: Optional[str] = None):
    """y = f(T) b, f(lambda) = lambda^5 / (lambda^6 + c^6), c = rtol max|lambda| + atol -- the soft pseudo-inverse
    of solver.py:94-111 -- as (1/3) Re sum_k (T - z_k I)^-1 b over the three roots of lambda^6 + c^6 in the upper
    half plane (exact partial fractions).  T is not overwritten.  ``replicated=True`` states that T and b are
    identical on every rank of the default process group: the ranks then take different shifts and exchange the
    double-double partial sums (one all-gather of 2 n doubles); otherwise every process does all three.
    ``method``: "ldlt" (own kernels, default) or "rational" (cuSOLVER LU)."""
    n = T.shape[0]
    if method is None:
        method = "rational" if PINV_METHOD == "rational" else "ldlt"
    rank, P = world() if replicated else (0, 1)
    mask = rational_shift_masks(P)[rank]
    nsh = _shift_count(mask)
    ws, wsz = _pinv_workspace(n, method, nsh)
    t = _phase_tic("pinv")
    tl = _phase_tic("pinv.lanczos")
    lam = sym_absmax_eig(T, method=method, nshifts=nsh)
    _phase_toc(tl)
    ydd = torch.zeros((2, n), dtype=torch.float64, device=T.device)
    info = torch.zeros(1, dtype=torch.int32, device=T.device)
    if mask:
        tf = _phase_tic("pinv.shifted_solves")
        _lib.call("qtx_pinv_ldlt_partial" if method == "ldlt" else "qtx_pinv_rational_partial", _lib.ptr(T), n,
                  _lib.ptr(b.contiguous()), -1.0 if rtol is None else float(rtol), float(atol), _lib.ptr(lam), int(mask),
                  int(REFINE_STEPS), _lib.ptr(ydd), 0, _lib.ptr(info), _lib.ptr(ws), wsz, _lib.stream())
        _phase_toc(tf)
    count = 1
    if P > 1:
        tc = _phase_tic("comm.all_gather_ydd")
        allydd = torch.empty((P, 2, n), dtype=torch.float64, device=T.device)
        _dist().all_gather_into_tensor(allydd.view(P * 2, n), ydd)
        info = info.abs()  # > 0: zero pivot, < 0: refinement did not contract -- any non-zero is a failure
        _dist().all_reduce(info, op=_dist().ReduceOp.MAX)
        _phase_toc(tc)
        ydd, count = allydd, P
    y = torch.empty(n, dtype=torch.float64, device=T.device)
    _lib.call("qtx_dd_sum_scale", _lib.ptr(ydd), count, n, 1.0 / 3.0, _lib.ptr(y), _lib.stream())
    _phase_toc(t)
    return y, info


def _use_rational(rtol, atol, tol_snr, want_evals) -> bool:
    if PINV_METHOD not in ("ldlt", "rational") or want_evals or tol_snr > 1e-6:
        return False
    return not (rtol is not None and float(rtol) == 0.0 and float(atol) == 0.0)  # plain inverse: eigenvalue route


def pinv_eig_solve(T: torch.Tensor, b: torch.Tensor, rtol: Optional[float], atol: float, want_evals: bool = False,
                   tol_snr: float = 0.0, replicated: bool = False, rational_ok: bool = False):
    """y = U (lambda^+ o rho) from eigh(T), rho = U^T b, optionally damped by the signal-to-noise ratio
    (``_sum_without_noise``, solver.py:114-125); T is overwritten by the eigenvectors
    (qtx_pinv_eig_solve / qtx_pinv_eig_solve_snr).  With ``QTX_PINV=rational`` the plain (no SNR, no eigenvalues)
    case of the callers that pass ``rational_ok`` (the SR / MinSR solves, whose result x = A^T y or f(S) A^T b has no
    component along null directions of T) goes through ``pinv_rational_solve`` instead and T is left untouched."""
    if rational_ok and _use_rational(rtol, atol, tol_snr, want_evals):
        return pinv_rational_solve(T, b, rtol, atol, replicated=replicated)
    n = T.shape[0]
    y = torch.empty(n, dtype=torch.float64, device=T.device)
    evals = torch.empty(n, dtype=torch.float64, device=T.device) if want_evals else None
    info = torch.empty(1, dtype=torch.int32, device=T.device)
    ws, wsz = _eig_workspace(n)
    rt = -1.0 if rtol is None else float(rtol)
    t = _phase_tic("pinv")
    if tol_snr > 1e-6:
        _lib.call("qtx_pinv_eig_solve_snr", _lib.ptr(T), n, _lib.ptr(b.contiguous()), rt, float(atol), float(tol_snr),
                  _lib.ptr(evals), _lib.ptr(y), _lib.ptr(info), _lib.ptr(ws), wsz, _lib.stream())
    else:
        _lib.call("qtx_pinv_eig_solve", _lib.ptr(T), n, _lib.ptr(b.contiguous()), rt, float(atol), _lib.ptr(evals),
                  _lib.ptr(y), _lib.ptr(info), _lib.ptr(ws), wsz, _lib.stream())
    _phase_toc(t)
    return (y, evals, info) if want_evals else (y, info)


def eigh(T: torch.Tensor):
    """(evals, info) = eigh(T); T [n, n] float64 is overwritten by the eigenvectors as ROWS (qtx_eigh)."""
    n = T.shape[0]
    evals = torch.empty(n, dtype=torch.float64, device=T.device)
    info = torch.empty(1, dtype=torch.int32, device=T.device)
    ws, wsz = _eig_workspace(n)
    _lib.call("qtx_eigh", _lib.ptr(T), n, _lib.ptr(evals), _lib.ptr(info), _lib.ptr(ws), wsz, _lib.stream())
    return evals, info


def rows_dot_snr(M: torch.Tensor, b: torch.Tensor, tol_snr: float) -> torch.Tensor:
    """rho[k] = sum_without_noise_i(M[k, i] b[i]) (solver.py:114-125) for a float64 matrix M [nrows, n]."""
    nrows, n = M.shape
    rho = torch.empty(nrows, dtype=torch.float64, device=M.device)
    _lib.call("qtx_rows_dot_snr", _lib.ptr2d(M), nrows, n, M.stride(0), _lib.ptr(b.contiguous()), float(tol_snr),
              _lib.ptr(rho), _lib.stream())
    return rho


def pinv_apply(Ut: torch.Tensor, evals: torch.Tensor, rho: torch.Tensor, rtol: Optional[float], atol: float):
    """y = U (lambda^+ o rho) for eigenvectors stored as rows (qtx_pinv_apply); rho is overwritten."""
    n = Ut.shape[0]
    y = torch.empty(n, dtype=torch.float64, device=Ut.device)
    _lib.call("qtx_pinv_apply", _lib.ptr(Ut), n, _lib.ptr(evals), _lib.ptr(rho), -1.0 if rtol is None else float(rtol),
              float(atol), _lib.ptr(y), _lib.stream())
    return y


def shift_chol_solve(T: torch.Tensor, b: torch.Tensor, rshift: Optional[float], ashift: float):
    """y = (T + (rshift tr T + ashift) I)^-1 b by Cholesky; T is overwritten (qtx_shift_chol_solve)."""
    n = T.shape[0]
    y = torch.empty(n, dtype=torch.float64, device=T.device)
    info = torch.empty(1, dtype=torch.int32, device=T.device)
    wsz = _lib.lib().qtx_shift_chol_workspace_size(n)
    if wsz == 0:
        raise _lib.QtxError(f"qtx_shift_chol_workspace_size failed: {_lib.lib().qtx_last_error().decode()}")
    ws = _WS.get("chol", wsz)
    _lib.call("qtx_shift_chol_solve", _lib.ptr(T), n, _lib.ptr(b.contiguous()), -1.0 if rshift is None else float(rshift),
              float(ashift), _lib.ptr(y), _lib.ptr(info), _lib.ptr(ws), wsz, _lib.stream())
    return y, info


def col_sumsq(A: torch.Tensor) -> torch.Tensor:
    """out[k] = sum_s A[s, k]^2 (float64 [np])."""
    ns, npar = A.shape
    out = torch.empty(npar, dtype=torch.float64, device=A.device)
    _lib.call("qtx_col_sumsq", _lib.dtype_code(A.dtype), _lib.ptr2d(A), ns, npar, A.stride(0), _lib.ptr(out),
              _lib.stream())
    return out


def matvec_t(A: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x = A^T y (float64 [np])."""
    ns, npar = A.shape
    x = torch.empty(npar, dtype=torch.float64, device=A.device)
    _lib.call("qtx_matvec_t", _lib.dtype_code(A.dtype), _lib.ptr2d(A), ns, npar, A.stride(0), _lib.ptr(y.contiguous()),
              _lib.ptr(x), 0, _lib.stream())
    return x


def matvec(A: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    ns, npar = A.shape
    v = torch.empty(ns, dtype=torch.float64, device=A.device)
    _lib.call("qtx_matvec", _lib.dtype_code(A.dtype), _lib.ptr2d(A), ns, npar, A.stride(0), _lib.ptr(x.contiguous()),
              _lib.ptr(v), _lib.stream())
    return v


class _CudaOps:
    """The dense kernels used by the distributed solve (tests substitute CPU stand-ins to exercise
    the exchange logic over gloo)."""

    def __init__(self, nslices):
        self.nslices = nslices

    def gram(self, A):
        return gram(A, nslices=self.nslices)

    def gram_allreduce(self, A):
        """Sum over the ranks of the Gram matrices of the column shards.  QTX_GRAM_P2P=1: one fused kernel pushes
        the tiles over NVLink peer memory and a rank-ordered reduce follows (peer.py); otherwise Gram + NCCL
        all-reduce."""
        if os.environ.get("QTX_GRAM_P2P", "0") == "1":
            from .peer import peer_gram

            return peer_gram(A.shape[0]).gram_allreduce(A, self.nslices)
        T = gram(A, nslices=self.nslices)
        t = _phase_tic("comm.all_reduce_T")
        _dist().all_reduce(T)
        _phase_toc(t)
        return T

    @staticmethod
    def pinv_eig_solve(T, b, rtol, atol):
        # T (all-reduced Gram) and b (all-gathered) are identical on every rank here
        return pinv_eig_solve(T, b, rtol, atol, replicated=True, rational_ok=True)

    matvec_t = staticmethod(matvec_t)


def _all_to_all(recv: torch.Tensor, send: torch.Tensor) -> None:
    """recv[p] <- send[rank] of rank p.  gloo has no all-to-all for device tensors: gather everything and pick."""
    dist = _dist()
    if dist.get_backend() != "gloo" or not send.is_cuda:
        dist.all_to_all_single(recv, send)
        return
    P, rank = dist.get_world_size(), dist.get_rank()
    allsend = torch.empty((P,) + tuple(send.shape), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(allsend.view(P * send.shape[0], *send.shape[1:]), send)
    recv.copy_(allsend[:, rank])


def distributed_minnorm(A: torch.Tensor, b: torch.Tensor, rtol, atol, ops, tsolve=None):
    """MinSR solve with the rows of A sharded over the ranks (solver.py:131-147 under GSPMD):
    row-sharded -> column-sharded all-to-all (the parameter axis is zero-padded to a multiple of
    the world size like ``array_extend(Adag, ndevices)``, solver.py:136), local Gram of the column
    shard, all-reduce of the partial Ns x Ns Grams, replicated eigh + pseudo-inverse, column shard
    of x = A^T y, all-gather.  ``tsolve(T, b_full) -> (y, info)`` replaces the eigh pseudo-inverse (SNR damping,
    diagonal-shift Cholesky).  Returns (x [Np], info)."""
    dist = _dist()
    P = dist.get_world_size()
    nl, npar = A.shape
    npc = (npar + P - 1) // P
    if npc * P == npar:
        send = A.view(nl, P, npc).permute(1, 0, 2).contiguous()
    else:
        flat = torch.zeros((nl, P * npc), dtype=A.dtype, device=A.device)
        flat[:, :npar] = A
        send = flat.view(nl, P, npc).permute(1, 0, 2).contiguous()
    recv = torch.empty_like(send)
    t = _phase_tic("comm.all_to_all_Obar")
    _all_to_all(recv, send)
    _phase_toc(t)
    Ac = recv.view(P * nl, npc)  # all Ns rows (rank-major = global sample order), this rank's columns
    if hasattr(ops, "gram_allreduce"):
        T = ops.gram_allreduce(Ac)
    else:
        T = ops.gram(Ac)
        t = _phase_tic("comm.all_reduce_T")
        dist.all_reduce(T)
        _phase_toc(t)
    bfull = torch.empty(P * nl, dtype=b.dtype, device=b.device)
    t = _phase_tic("comm.all_gather_b")
    dist.all_gather_into_tensor(bfull, b.contiguous())
    _phase_toc(t)
    y, info = ops.pinv_eig_solve(T, bfull, rtol, atol) if tsolve is None else tsolve(T, bfull)
    t = _phase_tic("matvec_t")
    xc = ops.matvec_t(Ac, y)
    _phase_toc(t)
    x = torch.empty(P * npc, dtype=xc.dtype, device=xc.device)
    t = _phase_tic("comm.all_gather_x")
    dist.all_gather_into_tensor(x, xc.contiguous())
    _phase_toc(t)
    return x[:npar].contiguous(), info


# The distributed MinSR solve runs entirely behind the C ABI (csrc/comm.cu: NCCL collectives bound by the library)
# when the process group is NCCL and the default pseudo-inverse route is in use; QTX_DIST_C=0 keeps the collectives in
# torch.distributed (the path the gloo tests exercise on CPU), which is also what SNR damping and the other T-solvers use.
DIST_IN_LIBRARY = os.environ.get("QTX_DIST_C", "1") == "1"
DIST_LANCZOS_STEPS = int(os.environ.get("QTX_DIST_LANCZOS_STEPS", str(LANCZOS_STEPS)))  # > 0 adaptive cap, < 0 exactly


def _dist_in_library(tol_snr: float) -> bool:
    if not DIST_IN_LIBRARY or PINV_METHOD != "ldlt" or tol_snr > 1e-6 or os.environ.get("QTX_GRAM_P2P", "0") == "1":
        return False
    dist = _dist()
    return dist.is_initialized() and dist.get_backend() == "nccl"


# ---- solvers (callables (A, b) -> x; A is the rank-local row block of Obar) -----------------------
def _dtype_rtol(rtol: Optional[float], dtype) -> float:
    """The reference takes the default cut-off from the dtype of the eigenvalues, i.e. of A (solver.py:12-21):
    1e-6 for float32, 1e-12 for float64.  The Gram is accumulated in float64 here whatever A is, so the default is
    resolved from A's dtype before it crosses the C ABI (whose own default is the float64 one)."""
    if rtol is not None:
        return float(rtol)
    return 1e-6 if dtype in (torch.float32, torch.complex64) else 1e-12


def _snr_tsolve(rtol, atol, tol_snr):
    if tol_snr > 1e-6:
        return lambda T, bfull: pinv_eig_solve(T, bfull, rtol, atol, tol_snr=tol_snr)
    return None


def minnorm_pinv_eig(rtol: Optional[float] = None, atol: float = 0.0, tol_snr: float = 0.0, nslices: Optional[int] = None):
    """x = A^+ b through T = A A^+ (MinSR, solver.py:128-149)."""

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        rank, P = world()
        rt = _dtype_rtol(rtol, A.dtype)
        if P == 1:
            T = gram(A, nslices=nslices)
            y, info = pinv_eig_solve(T, b, rt, atol, tol_snr=tol_snr, rational_ok=True)
            solve.last_info = info
            return matvec_t(A, y)
        if _dist_in_library(tol_snr):
            # every collective inside the library (qtx_minsr_solve_dist over the qtx_comm_* communicator)
            from .comm import minsr_solve_dist

            t = _phase_tic("minsr_dist(all-to-all + gram + all-reduce + pinv + A^T y + all-gather)")
            x, info = minsr_solve_dist(A, b, rt, atol, 0 if nslices is None else nslices, DIST_LANCZOS_STEPS, REFINE_STEPS,
                                       _WS.get)
            _phase_toc(t)
        else:
            x, info = distributed_minnorm(A, b, rt, atol, _CudaOps(nslices), _snr_tsolve(rt, atol, tol_snr))
        solve.last_info = info
        return x

    solve.last_info = None
    return solve


def _gather_rows_as_columns(M: torch.Tensor) -> torch.Tensor:
    """[k, nl] per rank -> [k, P nl] with rank-major columns (the global sample order)."""
    dist = _dist()
    P = dist.get_world_size()
    k, nl = M.shape
    buf = torch.empty((P * k, nl), dtype=M.dtype, device=M.device)  # concatenation along dim 0 (any backend)
    dist.all_gather_into_tensor(buf, M.contiguous())
    return buf.view(P, k, nl).permute(1, 0, 2).reshape(k, P * nl).contiguous()


def lstsq_pinv_eig(rtol: Optional[float] = None, atol: float = 0.0, tol_snr: float = 0.0, nslices: Optional[int] = None):
    """x = (A^+ A)^-1 A^+ b through S = A^+ A (SR, solver.py:152-164)."""

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        rank, P = world()
        rt = _dtype_rtol(rtol, A.dtype)
        At = A.t().contiguous()  # [np, nl]: S = At At^T sums over the local samples
        S = gram(At, nslices=nslices)
        if P > 1:
            _dist().all_reduce(S)
        if tol_snr > 1e-6:
            # rho_sk = (A V)[s, k] b[s] is damped per eigen-direction k over ALL samples s (solver.py:160-161)
            evals, info = eigh(S)  # S now holds V^T (eigenvectors as rows)
            M = torch.matmul(S, At.to(torch.float64))  # [np, nl] = (A V)^T, library GEMM (np is small here)
            bl = b
            if P > 1:
                M = _gather_rows_as_columns(M)
                bl = torch.empty(P * b.shape[0], dtype=b.dtype, device=b.device)
                _dist().all_gather_into_tensor(bl, b.contiguous())
            rho = rows_dot_snr(M, bl, tol_snr)
            x = pinv_apply(S, evals, rho, rt, atol)
        else:
            F = matvec_t(A, b)
            if P > 1:
                _dist().all_reduce(F)
            # S, F are all-reduced (the same on all ranks) and F = A^T b lies in the range of S
            x, info = pinv_eig_solve(S, F, rt, atol, replicated=True, rational_ok=True)
        solve.last_info = info
        return x

    solve.last_info = None
    return solve


def auto_pinv_eig(rtol: Optional[float] = None, atol: float = 0.0, tol_snr: float = 0.0, nslices: Optional[int] = None):
    """SR when Ns >= Np, MinSR otherwise (solver.py:167-201); Ns is the GLOBAL sample count."""
    mn = minnorm_pinv_eig(rtol, atol, tol_snr, nslices)
    ls = lstsq_pinv_eig(rtol, atol, tol_snr, nslices)

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        _, P = world()
        return mn(A, b) if A.shape[0] * P < A.shape[1] else ls(A, b)

    return solve


def minsr_pinv_eig(rtol: Optional[float] = None, atol: float = 0.0, tol_snr: float = 0.0):
    """Solver of T x = b for a given Hermitian T (solver.py:262-294)."""

    def solve(T: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        y, _ = pinv_eig_solve(T.clone(), b, rtol, atol, tol_snr=tol_snr)
        return y

    return solve


def minnorm_shift_eig(rshift: Optional[float] = None, ashift: float = 1e-4, nslices: Optional[int] = None):
    """x = A^+ (A A^+ + shift I)^-1 b, shift = rshift tr(T) + ashift, by Cholesky (solver.py:50-62)."""

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        rank, P = world()
        rs = _dtype_rtol(rshift, A.dtype)
        if P == 1:
            T = gram(A, nslices=nslices)
            y, info = shift_chol_solve(T, b, rs, ashift)
            solve.last_info = info
            return matvec_t(A, y)
        x, info = distributed_minnorm(A, b, None, 0.0, _CudaOps(nslices),
                                      lambda T, bfull: shift_chol_solve(T, bfull, rs, ashift))
        solve.last_info = info
        return x

    solve.last_info = None
    return solve


def lstsq_shift_eig(rshift: Optional[float] = None, ashift: float = 1e-4, nslices: Optional[int] = None):
    """x = (A^+ A + shift I)^-1 A^+ b by Cholesky (solver.py:65-77)."""

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        rank, P = world()
        S = gram(A.t().contiguous(), nslices=nslices)
        F = matvec_t(A, b)
        if P > 1:
            _dist().all_reduce(S)
            _dist().all_reduce(F)
        x, info = shift_chol_solve(S, F, _dtype_rtol(rshift, A.dtype), ashift)
        solve.last_info = info
        return x

    solve.last_info = None
    return solve


def auto_shift_eig(rshift: Optional[float] = None, ashift: float = 1e-4, nslices: Optional[int] = None):
    """Diagonal-shift SR when Ns >= Np, MinSR otherwise (solver.py:80-90); Ns is the GLOBAL sample count."""
    mn = minnorm_shift_eig(rshift, ashift, nslices)
    ls = lstsq_shift_eig(rshift, ashift, nslices)

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        _, P = world()
        return mn(A, b) if A.shape[0] * P < A.shape[1] else ls(A, b)

    return solve


def _dot(x: torch.Tensor, y: torch.Tensor) -> float:
    """x . y on the device (one row of qtx_matvec), read back for the CG recurrence."""
    return float(matvec(x.view(1, -1), y)[0])


class lstsq_shift_cg:
    """Conjugate-gradient solve of (A^+ A + diag_shift diag(A^+ A)) x = A^+ b (solver.py:24-47).  The iteration
    is ``jax.scipy.sparse.linalg.cg`` with x0 = 0 and no preconditioner: stop when |r|^2 <= max(rtol^2 |F|^2,
    atol^2) or after ``maxiter`` (default 10 Np) iterations.  S is never formed: two passes over A per iteration."""

    def __init__(self, diag_shift: float = 0.01, rtol: float = 1e-5, atol: float = 0.0, maxiter: Optional[int] = None):
        self.diag_shift, self.rtol, self.atol, self.maxiter = diag_shift, rtol, atol, maxiter
        self.last_iterations = None

    def S_apply(self, A: torch.Tensor, x: torch.Tensor, diag: torch.Tensor) -> torch.Tensor:
        _, P = world()
        out = matvec_t(A, matvec(A, x))
        if P > 1:
            _dist().all_reduce(out)
        return out.add_(diag * x, alpha=self.diag_shift)

    def __call__(self, A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        _, P = world()
        F = matvec_t(A, b)
        diag = col_sumsq(A)
        if P > 1:
            _dist().all_reduce(F)
            _dist().all_reduce(diag)
        maxiter = 10 * F.numel() if self.maxiter is None else self.maxiter
        x = torch.zeros_like(F)
        r = F.clone()
        p = r.clone()
        gamma = _dot(r, r)
        atol2 = max(self.rtol ** 2 * _dot(F, F), self.atol ** 2)
        k = 0
        while gamma > atol2 and k < maxiter:
            Ap = self.S_apply(A, p, diag)
            alpha = gamma / _dot(p, Ap)
            _axpby(alpha, p, 1.0, x)
            _axpby(-alpha, Ap, 1.0, r)
            gamma_new = _dot(r, r)
            _axpby(1.0, r, gamma_new / gamma, p)
            gamma = gamma_new
            k += 1
        self.last_iterations = k
        return x


def block_pinv_eig(state: Variational, rtol: Optional[float] = None, atol: float = 0.0, tol_snr: float = 0.0,
                   nslices: Optional[int] = None):
    """Layer-wise solver (solver.py:204-259): the parameter axis is cut at the layer boundaries of a
    ``Sequential`` model, every block is solved with ``auto_pinv_eig`` against Ebar / nlayers and the block
    solutions are concatenated."""
    sizes = getattr(state.model, "layer_param_sizes", None)
    if sizes is None:
        raise ValueError("`block_pinv_eig` solver only works for `Sequential` models.")
    sizes = [int(n) for n in sizes if int(n) > 0]
    bounds = [0]
    for n in sizes:
        bounds.append(bounds[-1] + n)
    if bounds[-1] != state.nparams:
        raise ValueError("layer_param_sizes do not add up to the number of parameters")
    nlayers = len(sizes)
    solver0 = auto_pinv_eig(rtol, atol, tol_snr, nslices)

    def solve(Obar: torch.Tensor, Ebar: torch.Tensor) -> torch.Tensor:
        Eb = _axpby(1.0 / nlayers, Ebar, 0.0, torch.empty_like(Ebar))
        parts = []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            parts.append(solver0(Obar[:, lo:hi].contiguous(), Eb))
        return torch.cat(parts)

    return solve


def sgd_solver():
    """x = A^+ b / Ns (solver.py:297-302; Ns is the global sample count)."""

    def solve(A: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        _, P = world()
        x = matvec_t(A, b)
        if P > 1:
            _dist().all_reduce(x)
        return _axpby(1.0 / (A.shape[0] * P), x, 0.0, torch.empty_like(x))

    return solve


# ---- optimizers ---------------------------------------------------------------------------------
class QNGD:
    """Quantum natural gradient descent base class (sr.py:17-128)."""

    def __init__(self, state: Variational, imag_time: bool = True, solver: Optional[Callable] = None):
        if not imag_time and state.vs_type != VS_TYPE.real_to_complex:
            raise NotImplementedError("real-time evolution needs a complex-output state (set_default_dtype(complex128))")
        self._state = state
        self._imag_time = imag_time
        # float32 model: the Jacobian entries carry float32 rounding (6e-8 relative), so the Gram digits beyond that are
        # spent on noise -- 5 digits (35 bits, 3e-11 of |a_i||a_j|) instead of 7 cut the tensor-core work from 28 to 15
        # digit products (tests/test_gram_tc_gpu.py::test_five_digits_suffice_for_float32_models)
        self._solver = auto_pinv_eig(nslices=model_gram_nslices(state)) if solver is None else solver
        self._Omean = None
        self.timers = None  # optional dict name -> [(start_event, end_event)]

    state = property(lambda self: self._state)
    holomorphic = property(lambda self: False)
    vs_type = property(lambda self: self._state.vs_type)
    imag_time = property(lambda self: self._imag_time)

    def _tic(self, name):
        if self.timers is None:
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        self.timers.setdefault(name, []).append(ev)
        return ev

    @staticmethod
    def _toc(ev):
        if ev is not None:
            ev[1].record()

    def get_Ebar(self, samples) -> torch.Tensor:
        raise NotImplementedError

    def get_Obar(self, samples) -> torch.Tensor:
        r"""Obar = (O - <O>) sqrt(rw / Ns) for the rank-local samples (sr.py:79-88)."""
        rank, P = world()
        state = self._state
        ns_glob = samples.nsamples * P
        rw = samples.reweight_factor
        scale = torch.sqrt(rw / ns_glob)
        ev = self._tic("jacobian")
        if state.vs_type == VS_TYPE.real_to_complex:
            # complex O, real parameters: the solver sees [Re Obar; Im Obar] (sr.py:99-104); centring and scaling
            # act on the real and the imaginary block separately ((O - <O>) sqrt(rw/Ns) is linear)
            nl = samples.nsamples
            Omat = state.jacobian_stacked(samples.spins)
            dt = _lib.dtype_code(Omat.dtype)
            self._Omean = []
            for blk in (Omat[:nl], Omat[nl:]):
                mean = torch.empty(Omat.shape[1], dtype=torch.float64, device=Omat.device)
                _lib.call("qtx_colmean", dt, _lib.ptr(blk), nl, blk.shape[1], blk.stride(0), None, _lib.ptr(mean),
                          _lib.stream())
                if P > 1:
                    _dist().all_reduce(mean)
                    mean /= P
                _lib.call("qtx_center_scale", dt, _lib.ptr(blk), nl, blk.shape[1], blk.stride(0), _lib.ptr(mean),
                          _lib.ptr(scale), _lib.stream())
                self._Omean.append(mean)
            self._toc(ev)
            return Omat
        mean, table = state.jacobian_colmean(samples.spins, return_table=True)
        if mean is not None:
            # one-pass path: column means first (without materialising O), then write Obar directly
            if P > 1:
                _dist().all_reduce(mean)
                mean /= P
            self._Omean = mean  # == mean(O * rw) for reweight 2 (rw = 1)
            Obar = state.jacobian(samples.spins, col_mean=mean, row_scale=scale, tanh_table=table)
        else:
            t = _phase_tic("jacobian.rows")
            Omat = state.jacobian(samples.spins)
            _phase_toc(t)
            mean = torch.empty(Omat.shape[1], dtype=torch.float64, device=Omat.device)
            dt = _lib.dtype_code(Omat.dtype)
            t = _phase_tic("jacobian.colmean")
            _lib.call("qtx_colmean", dt, _lib.ptr(Omat), Omat.shape[0], Omat.shape[1], Omat.stride(0), None,
                      _lib.ptr(mean), _lib.stream())
            if P > 1:
                _dist().all_reduce(mean)
                mean /= P
            _phase_toc(t)
            self._Omean = mean
            t = _phase_tic("jacobian.center_scale")
            _lib.call("qtx_center_scale", dt, _lib.ptr(Omat), Omat.shape[0], Omat.shape[1], Omat.stride(0),
                      _lib.ptr(mean), _lib.ptr(scale), _lib.stream())
            _phase_toc(t)
            Obar = Omat
        self._toc(ev)
        return Obar

    def solve(self, Obar: torch.Tensor, Ebar: torch.Tensor) -> torch.Tensor:
        """Solve Obar x = Ebar (sr.py:90-113).  Real parameters: for a complex-output state Obar and Ebar arrive
        stacked as [Re; Im]; real-time evolution (``imag_time=False``) solves against [-Im Ebar; Re Ebar]
        (sr.py:99-104)."""
        if self._state.vs_type == VS_TYPE.real_to_complex and not self._imag_time:
            n = Ebar.shape[0] // 2
            Ebar = torch.cat([-Ebar[n:], Ebar[:n]])
        ev = self._tic("solve")
        step = self._solver(Obar, Ebar)
        self._toc(ev)
        # real parameters: the reference casts the step to the (complex) default dtype and update() takes its real
        # part again (sr.py:110, variational.py:577); the real step is returned directly
        return step.to(get_real_dtype())

    def save(self, file) -> None:
        """Save the optimizer internal quantities (sr.py:125-128): plain SR has none."""

    def get_step(self, samples, **kw) -> torch.Tensor:
        from .global_defs import nvtx_range

        with nvtx_range("qtx.Ebar(Oloc)"):
            Ebar = self.get_Ebar(samples, **kw)
        with nvtx_range("qtx.Obar(jacobian)"):
            Obar = self.get_Obar(samples)
        with nvtx_range("qtx.solve"):
            return self.solve(Obar, Ebar)


class SR(QNGD):
    """Stochastic reconfiguration; picks SR or MinSR by shape (sr.py:131-195)."""

    def __init__(self, state: Variational, hamiltonian, imag_time: bool = True, solver: Optional[Callable] = None):
        super().__init__(state, imag_time, solver)
        self._hamiltonian = hamiltonian
        self._stats = None

    hamiltonian = property(lambda self: self._hamiltonian)

    @property
    def energy(self) -> Optional[float]:
        return None if self._stats is None else float(self._stats[0].item())

    @property
    def VarE(self) -> Optional[float]:
        return None if self._stats is None else float(self._stats[1].item())

    def get_Ebar(self, samples, Eloc: Optional[torch.Tensor] = None) -> torch.Tensor:
        r"""Ebar = (Eloc - <Eloc>) sqrt(rw / Ns); also stores energy and VarE (sr.py:180-195).
        ``Eloc`` may be passed when the local energies of these samples are already known."""
        rank, P = world()
        if Eloc is None:
            ev = self._tic("oloc")
            Eloc = self._hamiltonian.Oloc(self._state, samples)
            self._toc(ev)
        if Eloc.is_complex():
            Eloc = Eloc.to(torch.complex128).contiguous()
            nl = Eloc.shape[0]
            rw = samples.reweight_factor.contiguous()
            if P > 1:
                full = torch.empty(nl * P, dtype=torch.complex128, device=Eloc.device)
                _dist().all_gather_into_tensor(torch.view_as_real(full), torch.view_as_real(Eloc))
                rwf = torch.empty(nl * P, dtype=torch.float64, device=Eloc.device)
                _dist().all_gather_into_tensor(rwf, rw)
                Eloc, rw = full, rwf
            n = Eloc.shape[0]
            ebar = torch.empty(2 * n, dtype=torch.float64, device=Eloc.device)  # [Re; Im] (sr.py:102)
            stats = torch.empty(2, dtype=torch.float64, device=Eloc.device)
            _lib.call("qtx_ebar_cplx", _lib.ptr(Eloc), _lib.ptr(rw), n, _lib.ptr(ebar), n, _lib.ptr(stats), _lib.stream())
            self._stats = stats
            self._Eloc = Eloc
            if P > 1:  # this rank's rows, stacked like its block of Obar
                return torch.cat([ebar[rank * nl:(rank + 1) * nl], ebar[n + rank * nl:n + (rank + 1) * nl]])
            return ebar
        Eloc = Eloc.to(torch.float64)
        rw = samples.reweight_factor
        nl = Eloc.shape[0]
        if P > 1:
            full = torch.empty(nl * P, dtype=torch.float64, device=Eloc.device)
            _dist().all_gather_into_tensor(full, Eloc.contiguous())
            rwf = torch.empty(nl * P, dtype=torch.float64, device=Eloc.device)
            _dist().all_gather_into_tensor(rwf, rw.contiguous())
            Eloc, rw = full, rwf
        ebar = torch.empty_like(Eloc)
        stats = torch.empty(2, dtype=torch.float64, device=Eloc.device)
        _lib.call("qtx_ebar", _lib.ptr(Eloc), _lib.ptr(rw.contiguous()), Eloc.shape[0], _lib.ptr(ebar),
                  _lib.ptr(stats), _lib.stream())
        self._stats = stats
        self._Eloc = Eloc
        return ebar[rank * nl:(rank + 1) * nl].contiguous() if P > 1 else ebar


def pinvh_solve(rtol: Optional[float] = None, atol: float = 0.0):
    """x = H^+ b for a Hermitian H through eigh and the soft cut-off (solver.py:104-111)."""
    return minsr_pinv_eig(rtol, atol)


def _axpby(a: float, x: torch.Tensor, b: float, y: torch.Tensor) -> torch.Tensor:
    """y <- a x + b y on float64 device vectors (qtx_axpby)."""
    _lib.call("qtx_axpby", x.numel(), float(a), _lib.ptr(x.contiguous()), float(b), _lib.ptr(y), _lib.stream())
    return y


class TimeEvol(SR):
    r"""Real-time evolution (TDVP), quantax/optimizer/time_evol.py:26-134, for states with real parameters and complex
    output (VS_TYPE.real_to_complex): S = Re(Obar^+ Obar) and F = -Im(Obar^+ Ebar) (time_evol.py:121-122) are
    A^T A and A^T b of the stacked real matrix A = [Re Obar; Im Obar] with b = [-Im Ebar; Re Ebar].  Assumes more
    samples than parameters.  With ``max_parallel`` (the state's backward chunk) smaller than the sample count S and
    F are accumulated chunk by chunk from un-centred Jacobians (time_evol.py:76-115)."""

    def __init__(self, state: Variational, hamiltonian, solver: Optional[Callable] = None):
        super().__init__(state, hamiltonian, imag_time=False, solver=pinvh_solve() if solver is None else solver)
        self._max_parallel = state.backward_chunk

    def _stats_ebar(self, samples):
        """Stacked Ebar [Re; Im] of the GLOBAL samples' statistics, local rows (SR.get_Ebar)."""
        return self.get_Ebar(samples)

    def get_SF(self, samples):
        """(S float64 [Np, Np], F float64 [Np]) with F = -Im(Obar^+ Ebar) already taken."""
        rank, P = world()
        nl = samples.nsamples
        np_ = self._state.nparams
        if self._max_parallel is None or nl <= self._max_parallel:
            eb = self.get_Ebar(samples)          # [Re; Im], centred and scaled by sqrt(1/Ns)
            A = self.get_Obar(samples)           # [Re Obar; Im Obar]
            b = torch.cat([-eb[nl:], eb[:nl]])
            S = gram(A.t().contiguous())
            F = matvec_t(A, b)
        else:
            # chunked: S = sum_c O_c^+ O_c / Ns - outer(conj(Om), Om),  F = sum_c O_c^+ E_c / Ns - conj(Om) <E>
            Eloc = self._hamiltonian.Oloc(self._state, samples).to(torch.complex128).contiguous()
            ns_glob = nl * P
            stats = torch.empty(2, dtype=torch.float64, device=Eloc.device)
            Eg = Eloc
            if P > 1:
                Eg = torch.empty(ns_glob, dtype=torch.complex128, device=Eloc.device)
                _dist().all_gather_into_tensor(torch.view_as_real(Eg), torch.view_as_real(Eloc))
            _lib.call("qtx_ebar_cplx", _lib.ptr(Eg), None, ns_glob, None, ns_glob, _lib.ptr(stats), _lib.stream())
            self._stats, self._Eloc = stats, Eg
            S = torch.zeros((np_, np_), dtype=torch.float64, device=Eloc.device)
            F = torch.zeros(np_, dtype=torch.float64, device=Eloc.device)
            msum = torch.zeros((2, np_), dtype=torch.float64, device=Eloc.device)  # sum Re O, sum Im O
            er = torch.view_as_real(Eloc)
            for lo in range(0, nl, self._max_parallel):
                hi = min(nl, lo + self._max_parallel)
                n = hi - lo
                A = self._state.jacobian_stacked(samples.spins[lo:hi])  # [2 n, Np]
                gram(A.t().contiguous(), out=S, accumulate=True)
                b = torch.cat([-er[lo:hi, 1], er[lo:hi, 0]]).contiguous()
                _lib.call("qtx_matvec_t", _lib.dtype_code(A.dtype), _lib.ptr2d(A), 2 * n, np_, A.stride(0), _lib.ptr(b),
                          _lib.ptr(F), 1, _lib.stream())
                ones = torch.ones(n, dtype=torch.float64, device=A.device)
                for k in range(2):
                    _lib.call("qtx_matvec_t", _lib.dtype_code(A.dtype), _lib.ptr2d(A[k * n:(k + 1) * n]), n, np_,
                              A.stride(0), _lib.ptr(ones), _lib.ptr(msum[k]), 1, _lib.stream())
            if P > 1:
                _dist().all_reduce(S)
                _dist().all_reduce(F)
                _dist().all_reduce(msum)
            S /= ns_glob
            F /= ns_glob
            msum /= ns_glob
            self._Omean = [msum[0], msum[1]]
            for k in range(2):  # Re outer(conj(Om), Om) = mr mr^T + mi mi^T
                _lib.call("qtx_rank1_update", np_, -1.0, _lib.ptr(msum[k]), _lib.ptr(S), _lib.stream())
            # -Im(conj(Om) <E>) = -(mr Ei - mi Er): F holds -Im(sum O^+ E)/Ns already
            Em = torch.view_as_real(Eg).mean(dim=0)
            _axpby(float(Em[1]), msum[0], 1.0, F)
            _axpby(-float(Em[0]), msum[1], 1.0, F)
            return S, F
        if P > 1:
            _dist().all_reduce(S)
            _dist().all_reduce(F)
        return S, F

    def solve(self, Smat: torch.Tensor, Fvec: torch.Tensor) -> torch.Tensor:
        ev = self._tic("solve")
        step = self._solver(Smat, Fvec)
        self._toc(ev)
        return step.to(get_real_dtype())

    def get_step(self, samples, **kw) -> torch.Tensor:
        if not torch.allclose(samples.reweight_factor, torch.ones_like(samples.reweight_factor)):
            raise ValueError("TimeEvol is only for non-reweighted samples")
        Smat, Fvec = self.get_SF(samples)
        return self.solve(Smat, Fvec)


class Driver:
    """quantax/optimizer/driver.py:10-29."""

    def __init__(self, state, sampler, optimizer, step_length: float):
        from .utils import DataTracer

        self._state, self._sampler, self._optimizer = state, sampler, optimizer
        self._step_length = step_length
        self._time = 0.0
        self.energy = DataTracer()
        self.VarE = DataTracer()

    def step(self) -> None:
        raise NotImplementedError


class Euler(Driver):
    """First order Euler driver (driver.py:32-41)."""

    def step(self) -> None:
        samples = self._sampler.sweep()
        step = self._optimizer.get_step(samples)
        self._state.update(step * self._step_length)
        self._time += self._step_length
        self.energy.append(self._optimizer.energy, self._time)
        self.VarE.append(self._optimizer.VarE, self._time)


class AdaptiveHeunEvolution(Driver):
    """Adaptive second order Heun driver for unitary time evolution (driver.py:44-102)."""

    def __init__(self, state, sampler, tdvp: TimeEvol, step_length: float = 1e-3, integ_threshold: float = 1e-3):
        from .utils import DataTracer

        super().__init__(state, sampler, tdvp, step_length)
        self._integ_threshold = integ_threshold
        self.step_size = DataTracer()

    def step(self) -> None:
        import numpy as np

        tdvp, dt, st, spl = self._optimizer, self._step_length, self._state, self._sampler

        def comb(pairs):  # sum_i a_i x_i on the device
            acc = torch.zeros_like(pairs[0][1])
            for a, x in pairs:
                _axpby(a, x, 1.0, acc)
            return acc

        stepi = tdvp.get_step(spl.sweep())
        st.update(stepi, lr=dt)
        stepf = tdvp.get_step(spl.sweep())
        step1 = comb([(0.5, stepi), (0.5, stepf)])
        st.update(stepi, lr=-dt / 2)
        stepm = tdvp.get_step(spl.sweep())
        st.update(comb([(1.0, stepm), (-1.0, stepi)]), lr=dt / 4)
        stepmm = tdvp.get_step(spl.sweep())
        st.update(stepmm, lr=dt / 2)
        Smat, Fvec = tdvp.get_SF(spl.sweep())
        stepff = tdvp.solve(Smat, Fvec)
        st.update(comb([(1.0, stepff), (-1.0, stepmm)]), lr=dt / 4)
        step2 = comb([(0.25, stepi), (0.25, stepm), (0.25, stepmm), (0.25, stepff)])
        diff = comb([(1.0, step1), (-1.0, step2)])
        new_err = float(np.sqrt(max(float(torch.dot(diff, matvec(Smat, diff))), 0.0))) * dt
        ratio = np.clip((self._integ_threshold / max(new_err, 1e-300)) ** (1 / 3), 0.2, 2)
        new_step_length = float(np.clip(dt * ratio, 1e-4, 1e-2))
        self._time += new_step_length
        self._step_length = new_step_length
        self.step_size.append(new_step_length, self._time)
        self.energy.append(tdvp.energy, self._time)
        self.VarE.append(tdvp.VarE, self._time)


class MinSR(SR):
    """SR with the min-norm (Ns x Ns) solver forced -- the name used by BASELINE.json and by
    older quantax releases (docs/.doctrees/optimizer/quantax.optimizer.MinSR)."""

    def __init__(self, state: Variational, hamiltonian, imag_time: bool = True, solver: Optional[Callable] = None):
        super().__init__(state, hamiltonian, imag_time, minnorm_pinv_eig() if solver is None else solver)


# ---- momentum variants (quantax/optimizer/sr.py:198-429) --------------------------------------------
def _vec(n: int) -> torch.Tensor:
    return torch.zeros(n, dtype=torch.float64, device=device())


def _save_leaves(file, leaves) -> None:
    """The optimizer's internal quantities in the leaf order of ``eqx.tree_serialise_leaves(file, (...))`` (one
    ``np.save`` blob per leaf, scalars included), written by rank 0 only (sr.py:256-262, 343-349, 423-429)."""
    from .utils import write_eqx_leaves

    if world()[0] == 0:
        write_eqx_leaves(file, [v.detach().cpu().numpy() if torch.is_tensor(v) else v for v in leaves])


def _load_leaves(file, like):
    """Leaves of a file written by ``_save_leaves`` / the reference, validated against ``like`` (scalars and vectors)."""
    from .utils import read_eqx_leaves

    got = read_eqx_leaves(file)
    if len(got) != len(like):
        raise ValueError(f"optimizer file holds {len(got)} leaves, expected {len(like)}")
    out = []
    for g, l in zip(got, like):
        if torch.is_tensor(l):
            if g.size != l.numel():
                raise ValueError(f"optimizer file: vector of {g.size} entries, expected {l.numel()}")
            out.append(torch.from_numpy(np.ascontiguousarray(np.real(g).astype(np.float64)).reshape(-1)).to(l.device))
        else:
            out.append(type(l)(np.asarray(g).reshape(())))
    return out


def _scale_columns(A: torch.Tensor, d: torch.Tensor) -> None:
    _lib.call("qtx_scale_columns", _lib.dtype_code(A.dtype), _lib.ptr2d(A), A.shape[0], A.shape[1], A.stride(0),
              _lib.ptr(d), _lib.stream())


class SPRING(SR):
    """SR with momentum (sr.py:198-262): Ebar -= mu Obar step_prev; step = solve + mu step_prev."""

    def __init__(self, state, hamiltonian, imag_time: bool = True, solver=None, mu: float = 0.9, file=None):
        super().__init__(state, hamiltonian, imag_time, solver)
        self._mu = mu
        self._last_step = _vec(state.nparams)
        if file is not None:
            self.load(file)

    def solve(self, Obar: torch.Tensor, Ebar: torch.Tensor) -> torch.Tensor:
        Ebar = _axpby(-self._mu, matvec(Obar, self._last_step), 1.0, Ebar.clone())
        step = super().solve(Obar, Ebar)
        step = _axpby(self._mu, self._last_step, 1.0, step)
        self._last_step = step.clone()
        return step

    def save(self, file) -> None:
        """sr.py:256-262: leaves (mu, last_step)."""
        _save_leaves(file, (float(self._mu), self._last_step))

    def load(self, file) -> None:
        self._mu, self._last_step = _load_leaves(file, (float(self._mu), self._last_step))


class MARCH(SR):
    """SR with first and second order momentum (sr.py:265-349)."""

    def __init__(self, state, hamiltonian, imag_time: bool = True, solver=None, mu: float = 0.95, beta: float = 0.995,
                 file=None):
        super().__init__(state, hamiltonian, imag_time, solver)
        self._mu, self._beta = mu, beta
        self._last_step = _vec(state.nparams)
        self._V = _vec(state.nparams)
        self._t = 0
        self._V_is_zero = True
        if file is not None:
            self.load(file)

    def save(self, file) -> None:
        """sr.py:343-349: leaves (mu, beta, last_step, V, t)."""
        _save_leaves(file, (float(self._mu), float(self._beta), self._last_step, self._V, int(self._t)))

    def load(self, file) -> None:
        self._mu, self._beta, self._last_step, self._V, self._t = _load_leaves(
            file, (float(self._mu), float(self._beta), self._last_step, self._V, int(self._t)))
        self._V_is_zero = not bool(torch.any(self._V.abs() > 1e-8))  # `jnp.allclose(self._V, 0)` (sr.py:298)

    def solve(self, Obar: torch.Tensor, Ebar: torch.Tensor) -> torch.Tensor:
        self._t += 1
        n = self._last_step.numel()
        Ebar = _axpby(-self._mu, matvec(Obar, self._last_step), 1.0, Ebar.clone())
        V = torch.ones(n, dtype=torch.float64, device=Obar.device)
        if not self._V_is_zero:  # `jnp.allclose(self._V, 0)` is true exactly before the first update (sr.py:298)
            _lib.call("qtx_fourth_root", n, _lib.ptr(self._V), 1.0 - self._beta ** self._t, 1e-8, _lib.ptr(V),
                      _lib.stream())
        _scale_columns(Obar, V)
        raw = super().solve(Obar, Ebar)
        step = torch.empty_like(raw)
        _lib.call("qtx_div_add", n, _lib.ptr(raw), _lib.ptr(V), float(self._mu), _lib.ptr(self._last_step),
                  _lib.ptr(step), _lib.stream())
        _lib.call("qtx_second_moment", n, float(self._beta), _lib.ptr(step), _lib.ptr(self._last_step),
                  _lib.ptr(self._V), _lib.stream())
        self._V_is_zero = False
        self._last_step = step.clone()
        return step


class AdamSR(SR):
    """Adam-like SR (sr.py:352-429); two solves per step."""

    def __init__(self, state, hamiltonian, imag_time: bool = True, solver=None, mu: float = 0.95, beta: float = 0.995,
                 file=None):
        super().__init__(state, hamiltonian, imag_time, solver)
        self._mu, self._beta = mu, beta
        self._m = _vec(state.nparams)
        self._v = _vec(state.nparams)
        self._t = 0
        if file is not None:
            self.load(file)

    def save(self, file) -> None:
        """sr.py:423-429: leaves (mu, beta, m, v, t)."""
        _save_leaves(file, (float(self._mu), float(self._beta), self._m, self._v, int(self._t)))

    def load(self, file) -> None:
        self._mu, self._beta, self._m, self._v, self._t = _load_leaves(
            file, (float(self._mu), float(self._beta), self._m, self._v, int(self._t)))

    def solve(self, Obar: torch.Tensor, Ebar: torch.Tensor) -> torch.Tensor:
        self._t += 1
        n = self._m.numel()
        g = super().solve(Obar, Ebar)
        _axpby(1.0 - self._mu, g, self._mu, self._m)
        _lib.call("qtx_second_moment", n, float(self._beta), _lib.ptr(g), None, _lib.ptr(self._v), _lib.stream())
        m = _axpby(1.0 / (1.0 - self._mu ** self._t), self._m, 0.0, torch.empty_like(self._m))
        V = torch.empty_like(self._v)
        _lib.call("qtx_fourth_root", n, _lib.ptr(self._v), 1.0 - self._beta ** self._t, 1e-8, _lib.ptr(V), _lib.stream())
        Ebar = _axpby(-1.0, matvec(Obar, m), 1.0, Ebar.clone())
        _scale_columns(Obar, V)
        raw = super().solve(Obar, Ebar)
        step = torch.empty_like(raw)
        _lib.call("qtx_div_add", n, _lib.ptr(raw), _lib.ptr(V), 1.0, _lib.ptr(m), _lib.ptr(step), _lib.stream())
        return step
