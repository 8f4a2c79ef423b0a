"""Variational state: psi(s), cached internals, local updates, per-sample log-derivatives.

Mirrors quantax/state/variational.py:97-587 (``Variational``) and the ``State`` protocol of
quantax/state/state.py:19-161 for the models of the hot path.  Every method routes to the C ABI
(include/qtx_b200.h); there is no PyTorch implementation of the arithmetic.
"""
from __future__ import annotations

from enum import Enum
from typing import Optional, Tuple
from warnings import warn

import numpy as np
import torch

from . import _lib
from .global_defs import device, get_default_dtype, get_real_dtype, get_sites, is_default_cpl
from .utils import LogArray, ScaleArray


class VS_TYPE(Enum):
    """quantax/state/variational.py:40-94."""

    real_or_holomorphic = 0
    non_holomorphic = 1
    real_to_complex = 2


class State:
    """Abstract state (quantax/state/state.py:19-100): default no-op internals."""

    def __init__(self, symm=None):
        from .symmetry import Identity

        sites = get_sites()
        self._Nsites, self._Nmodes = sites.Nsites, sites.Nmodes
        self._symm = Identity() if symm is None else symm

    symm = property(lambda self: self._symm)

    Nsites = property(lambda self: self._Nsites)
    Nmodes = property(lambda self: self._Nmodes)
    use_ref = property(lambda self: False)

    def init_internal(self, s):
        return None


    def ref_forward_with_updates(self, s, s_old, nflips: int, internal):
        """psi(s) from psi(s_old) and its internal quantities, returning the updated internals (state.py:77-88)."""
        raise NotImplementedError

    def ref_forward(self, s, s_old, nflips: int, idx_segment, internal):
        """psi of connected configurations s that differ from s_old[idx_segment] by nflips sites (state.py:90-100)."""
        raise NotImplementedError


class Variational(State):
    def __init__(self, model, param_file=None, symm=None, max_parallel=None, use_ref: bool = True):
        super().__init__(symm)
        self._model = model
        if param_file is not None:
            self.load(param_file)
        if max_parallel is None or isinstance(max_parallel, int):
            self._forward_chunk = self._backward_chunk = self._ref_chunk = max_parallel
        elif len(max_parallel) == 2:
            self._forward_chunk, self._backward_chunk = max_parallel
            self._ref_chunk = self._forward_chunk
        else:
            self._forward_chunk, self._backward_chunk, self._ref_chunk = max_parallel
        # local updates are implemented for the un-projected RefModel; a projected state evaluates
        # full forwards of all symmetry images (the reference's use_ref=False code path)
        self._use_ref = bool(use_ref) and getattr(model, "is_ref_model", False) and self._symm.is_identity
        cplx_model = bool(getattr(model, "cplx", False))
        if cplx_model != is_default_cpl():
            # variational.py:244-257 allows a real model under a complex default dtype (the output is cast); here the
            # two must agree: complex-output model <-> complex128 default dtype
            raise NotImplementedError("a complex-output model needs set_default_dtype(torch.complex128), and vice versa")
        self._vs_type = VS_TYPE.real_to_complex if cplx_model else VS_TYPE.real_or_holomorphic
        self._ws = {}

    # ---- properties (variational.py:180-226) ------------------------------------------------
    use_ref = property(lambda self: self._use_ref)
    model = property(lambda self: self._model)
    holomorphic = property(lambda self: False)
    forward_chunk = property(lambda self: self._forward_chunk)
    backward_chunk = property(lambda self: self._backward_chunk)
    ref_chunk = property(lambda self: self._ref_chunk)
    nparams = property(lambda self: self._model.nparams)
    dtype = property(lambda self: self._model.dtype)
    vs_type = property(lambda self: self._vs_type)

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device())
            self._ws[key] = buf
        return buf

    def _mdt(self) -> int:
        return _lib.dtype_code(self._model.dtype)

    @staticmethod
    def _spins(s) -> torch.Tensor:
        from .operator import _as_spins

        return _as_spins(s)

    # ---- forward (variational.py:325-347) -----------------------------------------------------
    def _forward_model(self, s: torch.Tensor):
        """psi of the bare model (no projection) for a batch [ns, N]."""
        m = self._model
        ns = s.shape[0]
        if m.kind == "rbm":
            logabs = torch.empty(ns, dtype=torch.float64, device=s.device)
            _lib.call("qtx_rbm_forward", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), ns,
                      None, _lib.ptr(logabs), _lib.stream())
            return LogArray(torch.ones_like(logabs), logabs)
        from .resconv import resconv_forward

        return resconv_forward(self, s)

    def _images(self, s: torch.Tensor) -> torch.Tensor:
        """All symmetry images [ns * nsymm, N] (symmetry.py:325-341)."""
        perm, _ = self._symm.device_tables()
        nsymm = self._symm.nsymm
        out = torch.empty((s.shape[0] * nsymm, s.shape[1]), dtype=torch.int8, device=s.device)
        _lib.call("qtx_symm_images", _lib.ptr(s), s.shape[0], s.shape[1], _lib.ptr(perm), perm.shape[0],
                  int(self._symm.Z2_inversion != 0), _lib.ptr(out), _lib.stream())
        return out

    def _combine(self, psi_img, ns: int, want_coef: bool = False):
        """psi = sum_g w_g psi_g (symmetry.py:386-392) in container arithmetic."""
        _, w = self._symm.device_tables()
        nsymm = self._symm.nsymm
        kind = 0 if isinstance(psi_img, LogArray) else 1
        dev = psi_img.mult.device
        if psi_img.mult.is_complex():
            mult = torch.empty(ns, dtype=torch.complex128, device=dev)
            expo = torch.empty(ns, dtype=torch.float64, device=dev)
            coef = torch.empty((ns, nsymm), dtype=torch.complex128, device=dev) if want_coef else None
            _lib.call("qtx_symm_combine_cplx", _lib.ptr(psi_img.mult.contiguous()), _lib.ptr(psi_img.expo.contiguous()),
                      ns, nsymm, _lib.ptr(w), _lib.ptr(mult), _lib.ptr(expo), _lib.ptr(coef), _lib.stream())
            psi = ScaleArray(mult, expo)
            return (psi, coef) if want_coef else psi
        mult = torch.empty(ns, dtype=torch.float64, device=dev)
        expo = torch.empty(ns, dtype=torch.float64, device=dev)
        coef = torch.empty((ns, nsymm), dtype=torch.float64, device=dev) if want_coef else None
        _lib.call("qtx_symm_combine", _lib.ptr(psi_img.mult.contiguous()), _lib.ptr(psi_img.expo.contiguous()), ns,
                  nsymm, _lib.ptr(w), kind, _lib.ptr(mult), _lib.ptr(expo), _lib.ptr(coef), _lib.stream())
        psi = LogArray(mult, expo) if kind == 0 else ScaleArray(mult, expo)
        return (psi, coef) if want_coef else psi

    def __call__(self, s):
        s = self._spins(s).reshape(-1, self.Nmodes)
        if self._symm.is_identity:
            return self._forward_model(s)
        ns = s.shape[0]
        chunk = max(1, (1 << 22) // max(self._symm.nsymm, 1))
        if ns <= chunk:
            return self._combine(self._forward_model(self._images(s)), ns)
        parts = [self._combine(self._forward_model(self._images(s[lo:lo + chunk])), min(chunk, ns - lo))
                 for lo in range(0, ns, chunk)]
        cls = type(parts[0])
        return cls(torch.cat([p.mult for p in parts]), torch.cat([p.expo for p in parts]))

    def init_internal(self, s):
        """theta = W s + b for RefModels, None otherwise (variational.py:349-356)."""
        if not self._use_ref:
            return None
        s = self._spins(s)
        m = self._model
        theta = torch.empty((s.shape[0], m.M), dtype=m.dtype, device=s.device)
        _lib.call("qtx_rbm_forward", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), s.shape[0],
                  _lib.ptr(theta), None, _lib.stream())
        return theta

    def ref_forward(self, s, s_old, nflips: int, idx_segment, internal):
        """psi of configurations `s` that differ from s_old[idx_segment] by `nflips` flips
        (variational.py:387-422)."""
        s = self._spins(s)
        if not self._use_ref:
            return self(s)
        m = self._model
        s_old = self._spins(s_old)
        if internal is None:
            internal = self.init_internal(s_old)
        seg = idx_segment.to(device=s.device, dtype=torch.int32).contiguous()
        logabs = torch.empty(s.shape[0], dtype=torch.float64, device=s.device)
        _lib.call("qtx_rbm_ref_forward", self._mdt(), _lib.ptr(m.W), m.N, m.M, _lib.ptr(internal.contiguous()),
                  _lib.ptr(s_old), s_old.shape[0], _lib.ptr(s), _lib.ptr(seg), s.shape[0], int(nflips),
                  _lib.ptr(logabs), _lib.stream())
        return LogArray(torch.ones_like(logabs), logabs)

    def ref_forward_with_updates(self, s, s_old, nflips: int, internal):
        """One proposal for every chain (variational.py:358-385).  The persistent sweep kernel
        (``fused_sweep``) replaces the per-step use of this method; it is kept for API parity."""
        s = self._spins(s)
        if not self._use_ref:
            return self(s), None
        seg = torch.arange(s.shape[0], dtype=torch.int32, device=s.device)
        psi = self.ref_forward(s, s_old, nflips, seg, internal)
        return psi, self.init_internal(s)

    # ---- fused hot-path entry points ------------------------------------------------------------
    def fused_sweep(self, spins: torch.Tensor, nsweeps: int, kind: int, nbr, max_nb: int, hop: int, reweight: float,
                    seed: int, step0: int, chain0: int, injected=None, record: bool = False):
        """Whole Metropolis sweep in one launch (RBM) -- see qtx_rbm_sweep."""
        m = self._model
        if m.kind != "rbm" or not self._use_ref:
            from .resconv import generic_sweep

            return generic_sweep(self, spins, nsweeps, kind, nbr, max_nb, hop, reweight, seed, step0, chain0,
                                 injected, record)
        ns = spins.shape[0]
        dev = spins.device
        logabs = torch.empty(ns, dtype=torch.float64, device=dev)
        logabs_chain = torch.empty(ns, dtype=torch.float64, device=dev)
        nacc = torch.empty(ns, dtype=torch.int32, device=dev)
        log = torch.empty((nsweeps, ns), dtype=torch.uint8, device=dev) if record else None
        wsz = _lib.lib().qtx_rbm_workspace_size(self._mdt(), m.N, m.M)
        ws = self._workspace("rbm", wsz)
        pos = slot = u = None
        if injected is not None:
            pos, slot, u = injected
            pos = pos.to(device=dev, dtype=torch.int32).contiguous()
            slot = None if slot is None else slot.to(device=dev, dtype=torch.int32).contiguous()
            u = u.to(device=dev, dtype=torch.float64).contiguous()
        _lib.call("qtx_rbm_sweep", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(spins), ns,
                  int(nsweeps), int(kind), _lib.ptr(nbr), int(max_nb), int(hop), float(reweight), _lib.ptr(pos),
                  _lib.ptr(slot), _lib.ptr(u), int(seed) & 0xFFFFFFFFFFFFFFFF, int(step0), int(chain0),
                  _lib.ptr(logabs), _lib.ptr(logabs_chain), _lib.ptr(nacc), _lib.ptr(log), _lib.ptr(ws), wsz,
                  _lib.stream())
        one = torch.ones_like(logabs)
        return LogArray(one, logabs), LogArray(one, logabs_chain), nacc, log

    @property
    def fused_oloc(self):
        """Fused local-energy kernel when the model supports local updates, else None."""
        if self._model.kind == "rbm" and self._use_ref:
            return self._rbm_oloc
        return None

    def _rbm_oloc(self, operator, s: torch.Tensor) -> torch.Tensor:
        m = self._model
        t = operator.term_table
        ns = s.shape[0]
        eloc = torch.empty(ns, dtype=torch.float64, device=s.device)
        nconn = torch.empty(ns, dtype=torch.int32, device=s.device)
        wsz = _lib.lib().qtx_rbm_workspace_size(self._mdt(), m.N, m.M)
        ws = self._workspace("rbm", wsz)
        _lib.call("qtx_rbm_oloc", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), ns,
                  _lib.ptr(t.coef), _lib.ptr(t.sites), _lib.ptr(t.ops), t.nterms, _lib.ptr(eloc), _lib.ptr(nconn),
                  _lib.ptr(ws), wsz, _lib.stream())
        operator._connectivity = nconn
        return eloc

    # ---- jacobian (variational.py:424-511) ----------------------------------------------------
    def jacobian(self, fock_states, out: Optional[torch.Tensor] = None, col_mean=None, row_scale=None,
                 tanh_table=None):
        r"""O[s, k] = (1/psi) d psi / d theta_k, [ns, nparams] in the default dtype; column order =
        ``get_params_flatten``.  With ``col_mean`` / ``row_scale`` the centred and scaled matrix
        Obar of sr.py:74-88 is written directly."""
        s = self._spins(fock_states)
        m = self._model
        ns = s.shape[0]
        if self._vs_type == VS_TYPE.real_to_complex:
            # complex [ns, Np], assembled from the stacked real matrix the optimizer works with
            st = self.jacobian_stacked(s)
            return torch.complex(st[:ns], st[ns:])
        odt = get_default_dtype()
        if out is None:
            out = torch.empty((ns, m.nparams), dtype=odt, device=s.device)
        if not self._symm.is_identity:
            self._projected_jacobian(s, out)
        elif m.kind == "rbm" and not getattr(m, "tied", False):
            _lib.call("qtx_rbm_jacobian", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), ns,
                      _lib.dtype_code(out.dtype), _lib.ptr2d(out), out.stride(0), _lib.ptr(col_mean),
                      _lib.ptr(row_scale), _lib.ptr(tanh_table), _lib.stream())
            return out
        else:
            self._model_jacobian(s, out)
        if col_mean is not None or row_scale is not None:
            _lib.call("qtx_center_scale", _lib.dtype_code(out.dtype), _lib.ptr2d(out), ns, m.nparams, out.stride(0),
                      _lib.ptr(col_mean), _lib.ptr(row_scale), _lib.stream())
        return out

    def jacobian_stacked(self, fock_states, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Complex-output state with real parameters: [2 ns, Np] real, rows [0, ns) = Re O, rows [ns, 2 ns) = Im O --
        the matrix QNGD.solve hands to the real solver (sr.py:99-104)."""
        s = self._spins(fock_states)
        m = self._model
        ns = s.shape[0]
        if out is None:
            out = torch.empty((2 * ns, m.nparams), dtype=get_real_dtype(), device=s.device)
        if self._symm.is_identity:
            self._model_jacobian(s, out)
        else:
            self._projected_jacobian(s, out)
        return out

    def _model_jacobian(self, s: torch.Tensor, out: torch.Tensor) -> None:
        m = self._model
        if m.kind == "rbm" and getattr(m, "tied", False):
            # RBM_Conv: theta of the equivalent dense RBM, then the tied-weight contraction (shallow_nets.py:129-173)
            theta = torch.empty((s.shape[0], m.M), dtype=m.dtype, device=s.device)
            _lib.call("qtx_rbm_forward", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), s.shape[0],
                      _lib.ptr(theta), None, _lib.stream())
            _lib.call("qtx_rbm_conv_jacobian", self._mdt(), _lib.ptr(theta), _lib.ptr(s), s.shape[0], m.channels, m.Lx,
                      m.Ly, _lib.dtype_code(out.dtype), _lib.ptr2d(out), out.stride(0), _lib.stream())
        elif m.kind == "rbm":
            _lib.call("qtx_rbm_jacobian", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), s.shape[0],
                      _lib.dtype_code(out.dtype), _lib.ptr2d(out), out.stride(0), None, None, None, _lib.stream())
        else:
            from .resconv import resconv_jacobian

            resconv_jacobian(self, s, out)

    def _projected_jacobian(self, s: torch.Tensor, out: torch.Tensor) -> None:
        """O(s) = sum_g (w_g psi_g / psi) O(T_g s)  (variational.py:438-491), in sample chunks."""
        m = self._model
        nsymm = self._symm.nsymm
        ns = s.shape[0]
        esz = out.element_size()
        cplx = self._vs_type == VS_TYPE.real_to_complex
        nre = 2 if cplx else 1
        chunk = max(1, min(ns, (2 << 30) // max(nre * nsymm * m.nparams * esz, 1)))
        if self._backward_chunk is not None:
            chunk = max(1, min(chunk, self._backward_chunk // nsymm if self._backward_chunk >= nsymm else 1))
        buf = torch.empty((nre * chunk * nsymm, m.nparams), dtype=out.dtype, device=s.device)
        for lo in range(0, ns, chunk):
            hi = min(ns, lo + chunk)
            img = self._images(s[lo:hi])
            _, coef = self._combine(self._forward_model(img), hi - lo, want_coef=True)
            if cplx:
                nimg = (hi - lo) * nsymm
                J = buf[: 2 * nimg]
                self._model_jacobian(img, J)  # rows [0, nimg) = Re, [nimg, 2 nimg) = Im
                _lib.call("qtx_weighted_rowsum_cplx", _lib.dtype_code(out.dtype), _lib.ptr2d(J), J.stride(0), nimg,
                          _lib.ptr(coef), hi - lo, nsymm, m.nparams, _lib.ptr2d(out[lo:]), out.stride(0), ns,
                          _lib.stream())
                continue
            J = buf[: (hi - lo) * nsymm]
            self._model_jacobian(img, J)
            _lib.call("qtx_weighted_rowsum", _lib.dtype_code(out.dtype), _lib.ptr2d(J), J.stride(0), _lib.ptr(coef),
                      hi - lo, nsymm, m.nparams, _lib.ptr2d(out[lo:hi]), out.stride(0), _lib.stream())

    def jacobian_colmean(self, fock_states, weight=None, return_table: bool = False):
        """Column mean of the Jacobian without materialising it (RBM), else None.  With ``return_table`` also
        returns tanh(theta) [ns, M], which ``jacobian(..., tanh_table=)`` reuses so that the centred rows are
        bitwise consistent with these means."""
        m = self._model
        if m.kind != "rbm" or getattr(m, "tied", False) or not self._symm.is_identity:
            return (None, None) if return_table else None
        s = self._spins(fock_states)
        ns = s.shape[0]
        mean = torch.empty(m.nparams, dtype=torch.float64, device=s.device)
        table = torch.empty((ns, m.M), dtype=m.dtype, device=s.device) if return_table else None
        wsz = _lib.lib().qtx_rbm_colmean_workspace_size(self._mdt(), m.N, m.M, ns)
        ws = self._workspace("colmean", wsz)
        _lib.call("qtx_rbm_jacobian_colmean", self._mdt(), _lib.ptr(m.W), _lib.ptr(m.b), m.N, m.M, _lib.ptr(s), ns,
                  _lib.ptr(weight), _lib.ptr(mean), _lib.ptr(table), _lib.ptr(ws), wsz, _lib.stream())
        return (mean, table) if return_table else mean

    # ---- parameters (variational.py:545-587) ----------------------------------------------------
    def get_params_flatten(self) -> torch.Tensor:
        return self._model.params

    def update(self, step: torch.Tensor, lr: float = 1.0) -> None:
        r"""theta' = theta - step (skipped, with a warning, when the step is not finite)."""
        step = step.to(device=device(), dtype=torch.float64).contiguous()
        flag = torch.empty(1, dtype=torch.int32, device=step.device)
        _lib.call("qtx_apply_update", self._mdt(), _lib.ptr(self._model.params), _lib.ptr(step), float(lr),
                  step.numel(), _lib.ptr(flag), _lib.stream())
        self._last_update_flag = flag  # checked lazily: no host sync on the hot path

    def check_last_update(self) -> bool:
        flag = getattr(self, "_last_update_flag", None)
        ok = True if flag is None else bool(flag.item())
        if not ok:
            warn("Got invalid update step. The update is interrupted.")
        return ok

    def save(self, file) -> None:
        """Write the model in the layout of ``eqx.tree_serialise_leaves`` (variational.py:581-587) so that the file
        loads into the reference's ``Variational(model, param_file=...)`` and vice versa."""
        from .utils import write_eqx_leaves

        from .global_defs import world

        if world()[0] != 0:  # the reference writes on process 0 only (variational.py:581-587)
            return
        m = self._model
        flat = m.params.detach().cpu().numpy()
        arrays = [flat[o:o + int(np.prod(shape))].reshape(shape) for _, o, shape in m.eqx_leaf_layout()]
        write_eqx_leaves(file, arrays, m.eqx_trailing_scalars())

    def load(self, file) -> None:
        """Read an equinox leaf file (or the flat .npy written by earlier versions of this package)."""
        from .utils import read_eqx_leaves

        leaves = read_eqx_leaves(file)
        n = self._model.nparams
        if len(leaves) == 1 and leaves[0].ndim == 1 and leaves[0].size == n:
            flat = leaves[0]
        else:
            arrays = [a for a in leaves if a.ndim >= 1]
            layout = self._model.eqx_leaf_layout()
            if len(arrays) != len(layout) or any(a.size != int(np.prod(sh)) for a, (_, _, sh) in zip(arrays, layout)):
                raise ValueError("parameter file does not match the model (leaf count / shapes)")
            flat = np.concatenate([a.ravel() for a in arrays])
        self._model.params.copy_(torch.from_numpy(np.ascontiguousarray(flat)).to(self._model.params))
