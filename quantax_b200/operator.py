"""Spin operators in the QuSpin list format and their local estimator on the GPU.

Mirrors quantax/operator/: ``Operator`` (algebra ``+ - * / @ .H``, ``Oloc``, ``expectation``,
``apply_diag``), the site operators ``sigma_x/z/p/m``, ``S_x/z/p/m`` and the Hamiltonians
``Ising`` / ``Heisenberg`` (quantax/operator/operator.py:187-595, site_operator.py,
common_operators.py:23-71).  ED helpers (quspin bridge) are outside the hot path.
"""
from __future__ import annotations

import copy
from numbers import Number
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from .global_defs import PARTICLE_TYPE, device, get_sites
from .utils import LogArray, ScaleArray

_OPCODE = {"z": 1, "x": 2, "+": 3, "-": 4, "I": 5}


class TermTable:
    """Flat device table of operator terms (include/qtx_b200.h, 'Hamiltonian term table')."""

    def __init__(self, coef: np.ndarray, sites: np.ndarray, ops: np.ndarray, nflips: np.ndarray):
        dev = device()
        self.nterms = int(coef.shape[0])
        self.nflips_host = nflips
        self.coef = torch.from_numpy(np.ascontiguousarray(coef, dtype=np.float64)).to(dev)
        self.sites = torch.from_numpy(np.ascontiguousarray(sites, dtype=np.uint16).view(np.int16)).to(dev)
        self.ops = torch.from_numpy(np.ascontiguousarray(ops, dtype=np.uint8)).to(dev)

    def select(self, mask: np.ndarray) -> "TermTable":
        return TermTable(self.coef.cpu().numpy()[mask], self.sites.cpu().numpy().view(np.uint16)[mask],
                         self.ops.cpu().numpy()[mask], self.nflips_host[mask])


def compile_terms(op_list) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """op list -> (coef f64 [T], sites u16 [T,4], ops u8 [T,4], nflips [T]) in op-list order."""
    coef, sites, ops, nfl = [], [], [], []
    for opstr, interaction in op_list:
        if len(opstr) > 4:
            raise NotImplementedError(f"operator string '{opstr}' acts on more than 4 sites")
        for ch in opstr:
            if ch not in _OPCODE:
                raise NotImplementedError(f"operator '{ch}' is outside the real spin hot path (supported: I z x + -)")
        for J, *index in interaction:
            if isinstance(J, complex):
                if J.imag != 0:
                    raise NotImplementedError("complex couplings are outside the B200 hot path")
                J = J.real
            acting = [i for ch, i in zip(opstr, index) if ch != "I"]
            if len(set(acting)) != len(acting):
                raise NotImplementedError("terms acting twice on one site are not supported")
            coef.append(float(J))
            sites.append(list(index) + [0] * (4 - len(index)))
            ops.append([_OPCODE[ch] for ch in opstr] + [0] * (4 - len(opstr)))
            nfl.append(sum(ch in "x+-" for ch in opstr))
    return (np.asarray(coef, dtype=np.float64), np.asarray(sites, dtype=np.uint16).reshape(-1, 4),
            np.asarray(ops, dtype=np.uint8).reshape(-1, 4), np.asarray(nfl, dtype=np.int32))


class Operator:
    """Quantum operator (quantax/operator/operator.py:187)."""

    def __init__(self, op_list: list):
        self._op_list = op_list
        self._table = None
        self._group_tables = None
        self._connectivity = None
        self.last_conn_count = 0

    @property
    def op_list(self) -> list:
        return self._op_list

    # ---- device tables --------------------------------------------------------------------
    @property
    def term_table(self) -> TermTable:
        if self._table is None:
            self._table = TermTable(*compile_terms(self._op_list))
        return self._table

    @property
    def group_tables(self) -> dict:
        """{nflips: TermTable} of the off-diagonal terms, in op-list order inside each group
        (the grouping of _apply_off_diag, operator.py:110-117)."""
        if self._group_tables is None:
            t = self.term_table
            self._group_tables = {}
            for nf in sorted(set(int(v) for v in t.nflips_host if v > 0), key=lambda v: list(t.nflips_host).index(v)):
                self._group_tables[nf] = t.select(t.nflips_host == nf)
        return self._group_tables

    # ---- algebra (operator.py:294-488) ------------------------------------------------------
    @property
    def expression(self) -> str:
        SUB = str.maketrans("0123456789", "₀₁₂₃₄₅₆₇₈₉")
        OP = str.maketrans({"x": "Sˣ", "y": "Sʸ", "z": "Sᶻ", "+": "S⁺", "-": "S⁻"})
        out = []
        for opstr, interaction in self.op_list:
            for J, *index in interaction:
                out.append(f"{J:+}")
                for op, i in zip(opstr, index):
                    out.append(f"{op.translate(OP)}{str(i).translate(SUB)}")
        return " ".join(out)

    def __repr__(self) -> str:
        return self.expression

    def __matmul__(self, other):
        if isinstance(other, Operator):
            op_list = []
            for s1, t1 in self.op_list:
                for s2, t2 in other.op_list:
                    op = [s1 + s2, []]
                    for J1, *i1 in t1:
                        for J2, *i2 in t2:
                            op[1].append([J1 * J2, *i1, *i2])
                    op_list.append(op)
            return Operator(op_list)
        return NotImplemented

    @property
    def H(self) -> "Operator":
        op_list = copy.deepcopy(self.op_list)
        trans = str.maketrans("+-", "-+")
        for i, (opstr, interaction) in enumerate(op_list):
            op_list[i][0] = opstr.translate(trans)[::-1]
            for term in interaction:
                term[0] = term[0].conjugate() if hasattr(term[0], "conjugate") else term[0]
                term[1:] = term[-1:0:-1]
        return Operator(op_list)

    @staticmethod
    def _is_zero(x) -> bool:
        return isinstance(x, Number) and bool(np.isclose(x, 0.0))

    def _merged(self, other: "Operator", inplace: bool) -> "Operator":
        """operator.py:384-393 / 409-418: terms of an opstr already present are appended to its group."""
        op_list = self.op_list if inplace else _copy_op_list(self.op_list)
        names = tuple(op for op, _ in op_list)
        for opstr2, interaction in other.op_list:
            terms = interaction if inplace else [list(t) for t in interaction]
            if opstr2 in names:
                op_list[names.index(opstr2)][1] += terms
            else:
                op_list.append([opstr2, terms])
        return Operator(op_list)

    def __add__(self, other):
        if isinstance(other, Number):
            if not self._is_zero(other):
                raise ValueError("Constant shift is not implemented for Operator.")
            return self
        if isinstance(other, Operator):
            return self._merged(other, inplace=False)
        return NotImplemented

    def __radd__(self, other):
        return self + other if isinstance(other, Number) else NotImplemented

    def __iadd__(self, other):
        if isinstance(other, Number):
            if not self._is_zero(other):
                raise ValueError("Constant shift is not implemented for Operator.")
            return self
        if isinstance(other, Operator):
            return self._merged(other, inplace=True)
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, Number):
            if not self._is_zero(other):
                raise ValueError("Constant shift is not implemented for Operator.")
            return self
        if isinstance(other, Operator):
            return self + (-other)
        return NotImplemented

    def __rsub__(self, other):
        if isinstance(other, Number):
            if not self._is_zero(other):
                raise ValueError("Constant shift is not implemented for Operator.")
            return -self
        return NotImplemented

    def __isub__(self, other):
        return self.__iadd__(-other)

    def __mul__(self, other):
        if isinstance(other, (Number, np.number)) or (torch.is_tensor(other) and other.numel() == 1):
            c = other.item() if hasattr(other, "item") else other
            op_list = _copy_op_list(self.op_list)
            for _, interaction in op_list:
                for term in interaction:
                    term[0] *= c
            return Operator(op_list)
        return NotImplemented

    __rmul__ = __mul__

    def __neg__(self):
        return (-1) * self

    def __truediv__(self, other):
        return self * (1 / other) if isinstance(other, Number) else NotImplemented

    # ---- applying the operator ----------------------------------------------------------------
    def apply_diag(self, s: torch.Tensor) -> torch.Tensor:
        """Diagonal matrix elements <s|O|s> (operator.py:81-93,490-491)."""
        s = _as_spins(s)
        t = self.term_table
        out = torch.empty(s.shape[0], dtype=torch.float64, device=s.device)
        _lib.call("qtx_apply_diag", _lib.ptr(s), s.shape[0], s.shape[1], _lib.ptr(t.coef), _lib.ptr(t.sites),
                  _lib.ptr(t.ops), t.nterms, _lib.ptr(out), _lib.stream())
        return out

    def apply_off_diag(self, s: torch.Tensor) -> dict:
        """Off-diagonal elements in the reference's dense layout (operator.py:96-119,493-497):
        ``{nflips: [s_conn int8 [ns, nconn, N], H_conn f64 [ns, nconn]]}`` with ``nconn`` = the number of terms of
        the group and ``H_conn = NaN`` where a term does not connect.  Built by scattering the compacted
        enumeration of ``get_conn`` (the hot path never materialises this tensor); entries the reference keeps with
        |H| <= 1e-8 are NaN here, and the configuration stored under a NaN entry is the input configuration --
        both are dropped by every consumer (operator.py:152-153)."""
        s = _as_spins(s)
        ns, N = s.shape
        out = {}
        for nflips, t in self.group_tables.items():
            segment, conn_idx, s_compact, H, _ = self.get_conn(s, nflips)
            H_conn = torch.full((ns, t.nterms), float("nan"), dtype=torch.float64, device=s.device)
            s_conn = s[:, None, :].repeat(1, t.nterms, 1)
            if segment.numel() > 0:
                seg, ci = segment.long(), conn_idx.long()
                H_conn[seg, ci] = H
                s_conn[seg, ci] = s_compact
            out[nflips] = [s_conn, H_conn]
        return out

    def get_conn(self, s: torch.Tensor, nflips: int, conn_size: Optional[int] = None, with_spins: bool = True):
        """Compacted connected configurations of one nflips group: the device-side equivalent of
        _apply_off_diag + _get_conn_size + _get_conn (operator.py:96-165) for ONE device range.
        Returns (segment i32, conn_idx i32, s_conn int8 [conn_size, N] or None, H f64, n_nonnan_max)."""
        s = _as_spins(s)
        t = self.group_tables[nflips]
        ns, N = s.shape
        nonnan = torch.empty(ns, dtype=torch.int32, device=s.device)
        valid = torch.empty(ns, dtype=torch.int32, device=s.device)
        st = _lib.stream()
        _lib.call("qtx_conn_count", _lib.ptr(s), ns, N, _lib.ptr(t.coef), _lib.ptr(t.sites), _lib.ptr(t.ops),
                  t.nterms, nflips, _lib.ptr(nonnan), _lib.ptr(valid), st)
        offsets = torch.empty(ns, dtype=torch.int64, device=s.device)
        total = torch.empty(1, dtype=torch.int64, device=s.device)
        _lib.call("qtx_exclusive_scan_i32", _lib.ptr(valid), ns, _lib.ptr(offsets), _lib.ptr(total), st)
        n_nonnan = int(nonnan.sum().item())  # the reference's .item() host sync (operator.py:549)
        if conn_size is None:
            conn_size = n_nonnan  # _get_conn_size with forward_chunk=None on one device
        segment = torch.empty(conn_size, dtype=torch.int32, device=s.device)
        conn_idx = torch.empty(conn_size, dtype=torch.int32, device=s.device)
        H = torch.empty(conn_size, dtype=torch.float64, device=s.device)
        s_conn = torch.empty((conn_size, N), dtype=torch.int8, device=s.device) if with_spins else None
        if conn_size > 0:
            _lib.call("qtx_conn_fill", _lib.ptr(s), ns, N, _lib.ptr(t.coef), _lib.ptr(t.sites), _lib.ptr(t.ops),
                      t.nterms, nflips, _lib.ptr(offsets), _lib.ptr(total), conn_size, _lib.ptr(segment),
                      _lib.ptr(conn_idx), _lib.ptr(H), _lib.ptr(s_conn), st)
        return segment, conn_idx, s_conn, H, n_nonnan

    def Oloc(self, state, samples) -> torch.Tensor:
        r"""Local operator O_loc(s) = sum_s' psi(s')/psi(s) <s|O|s'>  (operator.py:510-562)."""
        from .sampler import Samples

        fc, rc = getattr(state, "forward_chunk", None), getattr(state, "ref_chunk", None)
        if fc is not None and rc is not None and fc < rc:
            raise ValueError("Unsupported chunk size: forward_chunk < ref_chunk.")
        if isinstance(samples, Samples):
            s, psi = samples.spins, samples.psi
        else:
            s = _as_spins(samples)
            psi = None
        fused = getattr(state, "fused_oloc", None)
        if fused is not None:
            return fused(self, s)
        if psi is None:
            psi = state(s)
        out = self.apply_diag(s)
        cplx = psi.mult.is_complex()
        if cplx:  # complex amplitudes: complex128 local energies
            diag = out
            out = torch.empty(diag.shape[0], dtype=torch.complex128, device=diag.device)
            _lib.call("qtx_real_to_cplx", _lib.ptr(diag), diag.shape[0], _lib.ptr(out), _lib.stream())
        self.last_conn_count = 0  # connected configurations forwarded by this call (bench.py)
        for nflips in self.group_tables:
            segment, _, s_conn, H, _ = self.get_conn(s, nflips)
            if segment.numel() == 0:
                continue
            self.last_conn_count += segment.numel()
            psi_conn = state.ref_forward(s_conn, s, nflips, segment, None)
            _lib.call("qtx_oloc_reduce_cplx" if cplx else "qtx_oloc_reduce", _lib.ptr(segment), _lib.ptr(H), _lib.ptr(psi_conn.mult.contiguous()),
                      _lib.ptr(psi_conn.expo.contiguous()), segment.numel(), _lib.ptr(psi.mult.contiguous()),
                      _lib.ptr(psi.expo.contiguous()), s.shape[0], _lib.ptr(out), _lib.stream())
        return out

    def expectation(self, state, samples, return_var: bool = False):
        """operator.py:564-595 (single process; the optimizer does the cross-rank reduction)."""
        from .sampler import Samples

        rw = samples.reweight_factor if isinstance(samples, Samples) else 1.0
        Oloc = self.Oloc(state, samples)
        Omean = torch.mean(Oloc * rw)
        if return_var:
            Ovar = torch.mean(Oloc.abs() ** 2 * rw) - Omean.abs() ** 2
            return Omean.item(), Ovar.item()
        return Omean.item()


def _copy_op_list(op_list):
    return [[opstr, [list(t) for t in interaction]] for opstr, interaction in op_list]


def _as_spins(s) -> torch.Tensor:
    if not torch.is_tensor(s):
        s = torch.as_tensor(np.asarray(s))
    s = s.to(device=device(), dtype=torch.int8)
    if s.ndim == 1:
        s = s[None]
    return s.contiguous()


# ---- site operators (quantax/operator/site_operator.py) ------------------------------------
def _site_operator(index: tuple, opstr: str, strength: float = 1.0) -> Operator:
    sites = get_sites()
    if len(index) == 1 and 0 <= index[0] < sites.Nsites:
        idx = int(index[0])
    else:
        shape = sites.shape
        if len(index) == len(shape):
            xyz, rest = [index[0]], index[1:]
        elif len(index) == len(shape) - 1 and shape[0] == 1:
            xyz, rest = [0], index
        else:
            raise ValueError("The input index doesn't match the shape of lattice.")
        sign = 1
        for x, l, bc in zip(rest, shape[1:], sites.boundary):
            xyz.append(x % l)
            sign *= bc ** abs(x // l)
        idx = int(sites.index_from_xyz[tuple(xyz)])
        strength *= sign
    return Operator([[opstr, [[strength, idx]]]])


def sigma_x(*index) -> Operator:
    return _site_operator(index, "x", 2.0)


def sigma_z(*index) -> Operator:
    return _site_operator(index, "z", 2.0)


def sigma_p(*index) -> Operator:
    return _site_operator(index, "+")


def sigma_m(*index) -> Operator:
    return _site_operator(index, "-")


def S_x(*index) -> Operator:
    return _site_operator(index, "x")


def S_z(*index) -> Operator:
    return _site_operator(index, "z")


S_p, S_m = sigma_p, sigma_m


# ---- Hamiltonians (quantax/operator/common_operators.py:23-71) ------------------------------
def Heisenberg(J: Union[Number, Sequence[Number]] = 1.0, n_neighbor: Union[int, Sequence[int]] = 1,
               msr: bool = False) -> Operator:
    r"""H = sum_n J_n sum_<ij>_n sigma_i . sigma_j; ``msr`` applies the Marshall sign rule to
    nearest-neighbour bonds."""
    sites = get_sites()
    if sites.particle_type != PARTICLE_TYPE.spin:
        raise ValueError("The Heisenberg model is only implemented in the spin system.")
    J = [J] if isinstance(J, Number) else list(J)
    n_neighbor = [n_neighbor] if isinstance(n_neighbor, Number) else list(n_neighbor)
    if len(J) != len(n_neighbor):
        raise ValueError("'J' and 'n_neighbor' should have the same length.")
    neighbors = sites.get_neighbor(n_neighbor)

    def hij(i, j, sign):
        hx = 2 * sign * (sigma_p(i) @ sigma_m(j) + sigma_m(i) @ sigma_p(j))
        return hx + sigma_z(i) @ sigma_z(j)

    H = 0
    for k, bonds in enumerate(neighbors):
        sign = -1 if msr and n_neighbor[k] == 1 else 1
        H = H + J[k] * sum(hij(int(i), int(j), sign) for i, j in bonds)
    return H


def Ising(h: Number = 0.0, J: Number = 1.0) -> Operator:
    r"""H = -J sum_<ij> sigma^z_i sigma^z_j - h sum_i sigma^x_i."""
    sites = get_sites()
    if sites.particle_type != PARTICLE_TYPE.spin:
        raise ValueError("The Ising model is only implemented in the spin system.")
    H = -h * sum(sigma_x(i) for i in range(sites.Nmodes))
    H += -J * sum(sigma_z(int(i)) @ sigma_z(int(j)) for i, j in sites.get_neighbor())
    return H
