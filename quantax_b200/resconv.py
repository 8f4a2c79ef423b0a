"""ResConv entry points of ``Variational``: batched forward, generic Metropolis sweep (full forward
per proposal, quantax/state/variational.py:383-384) and the per-sample Jacobian.  All arithmetic
is in libqtx_b200 (csrc/resconv.cu); this module only chunks batches so the activation workspace
stays bounded (the role of ``max_parallel`` / ``forward_chunk`` in the reference,
quantax/state/variational.py:166-174, quantax/utils/function.py:88-146)."""
from __future__ import annotations

import torch

from . import _lib
from .utils import ScaleArray

# default per-launch sample caps when the state has no max_parallel (bytes of activation workspace)
_FWD_BUDGET = 12 << 30  # ~30 K configurations per tower launch at config E (the last round of a launch is partly idle)
_BWD_BUDGET = 24 << 30


def _shape_args(m):
    return (m.nblocks, m.channels, m.Lx, m.Ly, m.kh, m.kw)


def _chunk(state, ns, grad):
    m = state.model
    mdt = _lib.dtype_code(m.dtype)
    # marginal bytes per sample (the size-1 query also holds the fixed part: repacked weights, guards)
    ws = _lib.lib().qtx_resconv_workspace_size
    per = (ws(mdt, 1025, *_shape_args(m), int(grad)) - ws(mdt, 1, *_shape_args(m), int(grad))) // 1024
    cap = max(1, (_BWD_BUDGET if grad else _FWD_BUDGET) // max(per, 1))
    user = state.backward_chunk if grad else state.forward_chunk
    if user is not None:
        cap = min(cap, int(user))
    return max(1, min(ns, cap))


def resconv_forward(state, s: torch.Tensor) -> ScaleArray:
    m = state.model
    mdt = _lib.dtype_code(m.dtype)
    ns = s.shape[0]
    cplx = getattr(m, "cplx", False)
    sig = torch.empty(ns, dtype=torch.complex128 if cplx else torch.float64, device=s.device)
    ex = torch.empty(ns, dtype=torch.float64, device=s.device)
    if ns == 0:
        return ScaleArray(sig, ex)
    chunk = _chunk(state, ns, False)
    wsz = _lib.lib().qtx_resconv_workspace_size(mdt, chunk, *_shape_args(m), 0)
    ws = state._workspace("resconv_fwd", wsz)
    fn = "qtx_resconv_forward_cplx" if cplx else "qtx_resconv_forward"
    for lo in range(0, ns, chunk):
        hi = min(ns, lo + chunk)
        _lib.call(fn, mdt, _lib.ptr(m.params), *_shape_args(m), m.final, _lib.ptr(s[lo:hi]),
                  hi - lo, _lib.ptr(sig[lo:hi]), _lib.ptr(ex[lo:hi]), _lib.ptr(ws), wsz, _lib.stream())
    psi = ScaleArray(sig, ex)
    for layer in getattr(m, "raw_layers", ()):  # parameter-free layers acting on (psi, s), e.g. a phase layer
        psi = layer(psi, s)
    return psi


def _moved_only_ok(state, ns: int) -> bool:
    """The sweep may evaluate only the MOVED proposals when the state is a bare float32 ResConv served by the
    tensor-core tower and the batch fits one launch (the batch size then stays on the device)."""
    import os

    m = state.model
    if os.environ.get("QTX_SWEEP_COMPACT", "1") == "0":  # dev knob: evaluate every proposal like the reference
        return False
    if getattr(m, "kind", None) != "resconv":
        return False
    if not _lib.lib().qtx_resconv_tc_available(_lib.dtype_code(m.dtype), m.channels, m.Lx, m.Ly, m.kh, m.kw):
        return False
    nimg = ns * (1 if state.symm.is_identity else state.symm.nsymm)  # a projected state forwards all symmetry images
    return _chunk(state, nimg, False) >= nimg


def state_forward_n(state, s: torch.Tensor, count: torch.Tensor):
    """psi(s[:count]) of a (possibly symmetry-projected) ResConv state with the batch size on the device; entries
    beyond ``count`` are undefined."""
    if state.symm.is_identity:
        return resconv_forward_n(state, s, count)
    img = state._images(s)  # [ns * nsymm, N]: the images of sample i are rows i*nsymm .. (i+1)*nsymm - 1
    return state._combine(resconv_forward_n(state, img, count * state.symm.nsymm), s.shape[0])


def resconv_forward_n(state, s: torch.Tensor, count: torch.Tensor) -> ScaleArray:
    """psi of the first ``count`` (device int64 [1]) rows of ``s``; the other entries of the result are undefined."""
    m = state.model
    mdt = _lib.dtype_code(m.dtype)
    ns = s.shape[0]
    cplx = getattr(m, "cplx", False)
    sig = torch.empty(ns, dtype=torch.complex128 if cplx else torch.float64, device=s.device)
    ex = torch.empty(ns, dtype=torch.float64, device=s.device)
    wsz = _lib.lib().qtx_resconv_workspace_size(mdt, ns, *_shape_args(m), 0)
    ws = state._workspace("resconv_fwd", wsz)
    _lib.call("qtx_resconv_forward_n", mdt, _lib.ptr(m.params), *_shape_args(m), m.final, int(cplx), _lib.ptr(s), ns,
              _lib.ptr(count), _lib.ptr(sig), _lib.ptr(ex), _lib.ptr(ws), wsz, _lib.stream())
    psi = ScaleArray(sig, ex)
    for layer in getattr(m, "raw_layers", ()):
        psi = layer(psi, s)
    return psi


def resconv_jacobian(state, s: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """Real-output model: out [ns, Np].  Complex-output model: out [2 ns, Np] holds Re O in rows [0, ns) and Im O in
    rows [ns, 2 ns) (the stacking of sr.py:99-104).  Parameter-free phase layers do not change the log-derivative."""
    m = state.model
    mdt = _lib.dtype_code(m.dtype)
    ns = s.shape[0]
    if ns == 0:
        return out
    cplx = getattr(m, "cplx", False)
    chunk = _chunk(state, ns, True)
    wsz = _lib.lib().qtx_resconv_workspace_size(mdt, chunk, *_shape_args(m), 1)
    ws = state._workspace("resconv_bwd", wsz)
    for lo in range(0, ns, chunk):
        hi = min(ns, lo + chunk)
        if cplx:
            _lib.call("qtx_resconv_jacobian_cplx", mdt, _lib.ptr(m.params), *_shape_args(m), m.final,
                      _lib.ptr(s[lo:hi]), hi - lo, _lib.dtype_code(out.dtype), _lib.ptr2d(out[lo:]), out.stride(0), ns,
                      None, None, _lib.ptr(ws), wsz, _lib.stream())
        else:
            _lib.call("qtx_resconv_jacobian", mdt, _lib.ptr(m.params), *_shape_args(m), m.final, _lib.ptr(s[lo:hi]),
                      hi - lo, _lib.dtype_code(out.dtype), _lib.ptr2d(out[lo:hi]), out.stride(0), None, None,
                      _lib.ptr(ws), wsz, _lib.stream())
    return out


def generic_sweep(state, spins, nsweeps, kind, nbr, max_nb, hop, reweight, seed, step0, chain0, injected, record):
    """``_partial_sweep`` for a state without local updates (metropolis.py:246-275): every step is
    propose -> full forward of the proposed chains -> accept, all enqueued on the stream with no
    host synchronisation; psi of the current chains is carried along (Samples.psi)."""
    ns, N = spins.shape
    dev = spins.device
    m = state.model
    import os

    if (injected is None and not record and state.symm.is_identity and getattr(m, "kind", None) == "resconv"
            and os.environ.get("QTX_SWEEP_COMPACT", "1") != "0" and not getattr(m, "cplx", False) and not getattr(m, "raw_layers", ()) and ns > 0
            and _chunk(state, ns, False) >= ns):
        # bare real ResConv: the whole step loop runs behind one C-ABI call (qtx_resconv_sweep)
        mdt = _lib.dtype_code(m.dtype)
        wsz = _lib.lib().qtx_resconv_sweep_workspace_size(mdt, ns, *_shape_args(m))
        if wsz == 0:
            raise _lib.QtxError(f"qtx_resconv_sweep_workspace_size failed: {_lib.lib().qtx_last_error().decode()}")
        ws = state._workspace("resconv_sweep", wsz)
        sig = torch.empty(ns, dtype=torch.float64, device=dev)
        ex = torch.empty(ns, dtype=torch.float64, device=dev)
        nacc = torch.empty(ns, dtype=torch.int32, device=dev)
        _lib.call("qtx_resconv_sweep", mdt, _lib.ptr(m.params), *_shape_args(m), m.final, _lib.ptr(spins), ns,
                  int(nsweeps), int(kind), _lib.ptr(nbr), int(max_nb), int(hop), float(reweight),
                  int(seed) & 0xFFFFFFFFFFFFFFFF, int(step0), int(chain0), _lib.ptr(sig), _lib.ptr(ex), _lib.ptr(nacc),
                  _lib.ptr(ws), wsz, _lib.stream())
        out = ScaleArray(sig, ex)
        return out, out, nacc, None
    psi = state(spins)  # bare model or symmetry-projected, LogArray or ScaleArray
    mult, expo = psi.mult.contiguous(), psi.expo.contiguous()
    accept = "qtx_metropolis_accept_cplx" if mult.is_complex() else "qtx_metropolis_accept"
    new_spins = torch.empty_like(spins)
    moved = torch.empty(ns, dtype=torch.uint8, device=dev)
    nacc = torch.zeros(ns, dtype=torch.int32, device=dev)
    log = torch.empty((nsweeps, ns), dtype=torch.uint8, device=dev) if record else None
    pos = slot = u = None
    if injected is not None:
        pos, slot, u = injected
        pos = pos.to(device=dev, dtype=torch.int32).contiguous()
        slot = None if slot is None else slot.to(device=dev, dtype=torch.int32).contiguous()
        u = u.to(device=dev, dtype=torch.float64).contiguous()
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    st = _lib.stream()
    # exchange proposals of two equal spins are no-ops that can never be accepted (metropolis.py:314-316); the
    # reference still evaluates their psi.  Evaluate the moved proposals only when the forward allows it.
    compact = int(kind) == _lib.QTX_SPIN_EXCHANGE and _moved_only_ok(state, ns)
    if compact:
        rank = torch.empty(ns, dtype=torch.int32, device=dev)
        cspins = torch.zeros_like(spins)
        count = torch.empty(1, dtype=torch.int64, device=dev)
    for t in range(nsweeps):
        _lib.call("qtx_metropolis_propose", int(kind), _lib.ptr(spins), ns, N, _lib.ptr(nbr), int(max_nb), int(hop),
                  None if pos is None else _lib.ptr(pos[t]), None if slot is None else _lib.ptr(slot[t]), seed,
                  int(step0) + t, int(chain0), _lib.ptr(new_spins), _lib.ptr(moved), st)
        if compact:
            _lib.call("qtx_compact_moved", _lib.ptr(moved), _lib.ptr(new_spins), ns, N, _lib.ptr(rank), _lib.ptr(cspins),
                      _lib.ptr(count), st)
            psi_new = state_forward_n(state, cspins, count)
            _lib.call("qtx_metropolis_accept_compact", _lib.ptr(spins), _lib.ptr(new_spins), _lib.ptr(moved),
                      _lib.ptr(rank), ns, N, _lib.ptr(mult), _lib.ptr(expo), _lib.ptr(psi_new.mult.contiguous()),
                      _lib.ptr(psi_new.expo.contiguous()), int(mult.is_complex()), float(reweight),
                      None if u is None else _lib.ptr(u[t]), seed, int(step0) + t, int(chain0), _lib.ptr(nacc),
                      None if log is None else _lib.ptr(log[t]), st)
            continue
        psi_new = state(new_spins)
        _lib.call(accept, _lib.ptr(spins), _lib.ptr(new_spins), _lib.ptr(moved), ns, N,
                  _lib.ptr(mult), _lib.ptr(expo), _lib.ptr(psi_new.mult.contiguous()), _lib.ptr(psi_new.expo.contiguous()),
                  float(reweight), None if u is None else _lib.ptr(u[t]), seed, int(step0) + t, int(chain0),
                  _lib.ptr(nacc), None if log is None else _lib.ptr(log[t]), st)
    out = type(psi)(mult, expo)
    return out, out, nacc, log
