"""Oracle: QuSpin-format operator lists, connected configurations, local estimator.

Test infrastructure.  Restates
  quantax/operator/site_operator.py:8-36,39-130 (site operators and strengths),
  quantax/operator/operator.py:377-395,444-455,294-310 (operator algebra on op lists),
  quantax/operator/common_operators.py:23-71 (Heisenberg, Ising),
  quantax/operator/operator.py:30-78 (_apply_site_operator, spin branches),
  quantax/operator/operator.py:81-119 (_apply_diag/_apply_off_diag),
  quantax/operator/operator.py:122-165 (_get_conn_size/_get_conn),
  quantax/operator/operator.py:168-184,510-562 (_get_Olocx/Oloc).
"""
from __future__ import annotations

import copy
import numpy as np


# ----------------------------------------------------------------------------
# operator algebra on op lists  [[opstr, [[J, i, j, ...], ...]], ...]
# ----------------------------------------------------------------------------
def _site(opstr, strength, i):
    return [[opstr, [[strength, int(i)]]]]


def op_matmul(a, b):
    """operator.py:301-310."""
    out = []
    for s1, t1 in a:
        for s2, t2 in b:
            terms = []
            for J1, *i1 in t1:
                for J2, *i2 in t2:
                    terms.append([J1 * J2, *i1, *i2])
            out.append([s1 + s2, terms])
    return out


def op_add(a, b):
    """operator.py:384-393: terms of an already present opstr are appended to that group."""
    out = copy.deepcopy(a)
    names = tuple(s for s, _ in out)
    for s2, t2 in b:
        if s2 in names:
            out[names.index(s2)][1] += copy.deepcopy(t2)
        else:
            out.append([s2, copy.deepcopy(t2)])
    return out


def op_scale(a, c):
    """operator.py:444-453."""
    out = copy.deepcopy(a)
    for _, terms in out:
        for t in terms:
            t[0] *= c
    return out


def _sum_ops(ops):
    total = None
    for o in ops:  # python ``sum``: 0 + op -> op (operator.py:377-382,397-400)
        total = o if total is None else op_add(total, o)
    return total


def heisenberg_op_list(lattice, J=1.0, n_neighbor=1, msr=False):
    """common_operators.py:23-53."""
    J = [J] if np.isscalar(J) else list(J)
    n_neighbor = [n_neighbor] if np.isscalar(n_neighbor) else list(n_neighbor)
    neighbors = lattice.get_neighbor(n_neighbor)

    def hij(i, j, sign):
        pm = op_matmul(_site("+", 1.0, i), _site("-", 1.0, j))
        mp = op_matmul(_site("-", 1.0, i), _site("+", 1.0, j))
        hx = op_scale(op_add(pm, mp), 2 * sign)
        hz = op_matmul(_site("z", 2.0, i), _site("z", 2.0, j))
        return op_add(hx, hz)

    H = None
    for k, nbs in enumerate(neighbors):
        sign = -1 if (msr and n_neighbor[k] == 1) else 1
        part = op_scale(_sum_ops(hij(i, j, sign) for i, j in nbs), J[k])
        H = part if H is None else op_add(H, part)
    return H


def ising_op_list(lattice, h=0.0, J=1.0):
    """common_operators.py:56-71."""
    H = op_scale(_sum_ops(_site("x", 2.0, i) for i in range(lattice.Nmodes)), -h)
    zz = _sum_ops(op_matmul(_site("z", 2.0, i), _site("z", 2.0, j))
                  for i, j in lattice.get_neighbor())
    return op_add(H, op_scale(zz, -J))


def to_array_op_list(op_list, dtype=np.float64):
    """operator.py:220-236: J in the default dtype, indices uint16."""
    out = []
    for opstr, terms in op_list:
        J = np.asarray([t[0] for t in terms], dtype=dtype)
        idx = np.asarray([t[1:] for t in terms], dtype=np.uint16)
        out.append([opstr, J, idx])
    return out


# ----------------------------------------------------------------------------
# applying operators to configurations
# ----------------------------------------------------------------------------
def _apply_site(x, op, J, idx):
    """operator.py:30-78, spin system.  x: [ns, nterm, N] int8 (may be broadcast views,
    never written in place), J: [ns, nterm], idx: [nterm]."""
    if op == "I":
        return x, J
    ar = np.arange(idx.size)
    xi = x[:, ar, idx]
    if op == "z":
        return x, J * xi / 2
    if op in ("x", "y"):
        J = J / 2
    x = x.copy()
    if op == "+":
        J = np.where(xi < 0, J, np.nan)
        x[:, ar, idx] = 1
    elif op == "-":
        J = np.where(xi > 0, J, np.nan)
        x[:, ar, idx] = -1
    elif op == "x":
        x[:, ar, idx] = -xi
    elif op == "y":
        J = J * 1j * xi
        x[:, ar, idx] = -xi
    else:
        raise ValueError(f"operator '{op}' is outside the spin hot path")
    return x, J


def apply_diag(s, aop_list):
    """operator.py:81-93."""
    s = np.asarray(s, dtype=np.int8)
    Hz = np.zeros(s.shape[0], dtype=np.float64)
    for opstr, J, index in aop_list:
        if all(op in ("I", "n", "z") for op in opstr):
            x = np.broadcast_to(s[:, None, :], (s.shape[0], J.size, s.shape[1]))
            Jc = np.broadcast_to(J[None, :], (s.shape[0], J.size)).astype(J.dtype)
            for op, idx in zip(opstr, index.T.astype(np.int64)):
                _, Jc = _apply_site(x, op, Jc, idx)
            Hz = Hz + Jc.sum(axis=1)
    return Hz


def apply_off_diag(s, aop_list):
    """operator.py:96-119: {nflips: (s_conn [ns, nconn, N] int8, H_conn [ns, nconn])};
    invalid '+'/'-' applications carry H = NaN."""
    s = np.asarray(s, dtype=np.int8)
    out = {}
    for opstr, J, index in aop_list:
        nflips = sum(1 for c in opstr if c not in ("I", "n", "z"))
        if nflips == 0:
            continue
        x = np.repeat(s[:, None, :], J.size, axis=1)
        Jc = np.broadcast_to(J[None, :], (s.shape[0], J.size)).astype(J.dtype)
        idxT = index.T.astype(np.int64)
        for op, idx in zip(reversed(opstr), idxT[::-1]):  # right-most operator acts first
            x, Jc = _apply_site(x, op, Jc, idx)
        out.setdefault(nflips, [[], []])
        out[nflips][0].append(x)
        out[nflips][1].append(Jc)
    return {k: (np.concatenate(v[0], axis=1), np.concatenate(v[1], axis=1)) for k, v in out.items()}


def array_extend(a, multiple, axis=0, padding_values=0):
    """utils/array.py:104-130."""
    r = a.shape[axis] % multiple
    if r == 0:
        return a
    pad = [(0, 0)] * a.ndim
    pad[axis] = (0, multiple - r)
    return np.pad(a, pad, constant_values=padding_values)


def get_conn_size(H_conn, forward_chunk=None, ndevices=1):
    """operator.py:122-141."""
    ns, nconn = H_conn.shape
    if forward_chunk is None:
        H = H_conn.reshape(ndevices, -1, 1, nconn)
    else:
        H = H_conn.reshape(ndevices, -1, nconn)
        H = array_extend(H, forward_chunk, axis=1, padding_values=np.nan)
        H = H.reshape(ndevices, forward_chunk, -1, nconn)
    size = int(np.max(np.sum(~np.isnan(H), axis=(1, 3))))
    if forward_chunk is not None:
        size = ((size - 1) // forward_chunk + 1) * forward_chunk
    return size


def get_conn(s_conn, H_conn, conn_size, ndevices=1):
    """operator.py:144-165: per device, row-major (sample, conn) compaction of the valid
    entries (not NaN and |H| > 1e-8), padded with segment=-1 / H=0 up to conn_size."""
    ns, nconn, N = s_conn.shape
    Hd = H_conn.reshape(ndevices, -1, nconn)
    sd = s_conn.reshape(ndevices, -1, nconn, N)
    segs, ss, Hs = [], [], []
    for d in range(ndevices):
        valid = ~(np.isnan(Hd[d]) | np.isclose(Hd[d], 0))
        seg, cidx = np.nonzero(valid)
        if seg.size > conn_size:
            raise ValueError("conn_size smaller than the number of valid connections")
        npad = conn_size - seg.size
        seg = np.concatenate([seg, -np.ones(npad, dtype=seg.dtype)])
        cidx = np.concatenate([cidx, -np.ones(npad, dtype=cidx.dtype)])
        sc = sd[d][seg, cidx]  # index -1 wraps like jnp: the last sample/conn (unused)
        Hc = np.where(seg == -1, 0, np.nan_to_num(Hd[d][seg, cidx], nan=0.0))
        segs.append(seg), ss.append(sc), Hs.append(Hc)
    return np.concatenate(segs), np.concatenate(ss, axis=0), np.concatenate(Hs)


def oloc(aop_list, forward, s, psi=None, ndevices=1, forward_chunk=None):
    """operator.py:510-562 + 168-184.

    ``forward(spins) -> (mult, expo)`` with psi = mult * exp(expo) (LogArray: (sign,
    logabs); ScaleArray: (significand, exponent)).  The ratio is formed in container
    arithmetic then densified: (mult'/mult) * exp(expo' - expo)  (operator.py:179,
    utils/big_array.py:330-335,574-579)."""
    s = np.asarray(s, dtype=np.int8)
    if psi is None:
        psi = forward(s)
    mult, expo = psi
    out = apply_diag(s, aop_list).astype(np.result_type(np.float64, np.asarray(mult).dtype))
    nper = s.shape[0] // ndevices
    for nflips, (s_conn, H_conn) in apply_off_diag(s, aop_list).items():
        size = get_conn_size(H_conn, forward_chunk, ndevices)
        seg, sc, Hc = get_conn(s_conn, H_conn, size, ndevices)
        m2, e2 = forward(sc)
        dev = np.repeat(np.arange(ndevices), size)
        g = np.where(seg >= 0, seg + dev * nper, 0)
        ratio = (m2 / mult[g]) * np.exp(e2 - expo[g])
        contrib = np.where(seg >= 0, ratio * Hc, 0)
        np.add.at(out, g, contrib)  # segment_sum (order unspecified in the reference)
    return out


# ----------------------------------------------------------------------------
# exact diagonalisation through the SAME apply functions (pins the conventions
# against the ED energies printed in the reference tutorials)
# ----------------------------------------------------------------------------
def ed_lowest(aop_list, N, nup=None, k=2):
    import itertools
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    if nup is None:
        states = np.array(list(itertools.product([1, -1], repeat=N)), dtype=np.int8)
    else:
        states = []
        for ups in itertools.combinations(range(N), nup):
            v = -np.ones(N, dtype=np.int8)
            v[list(ups)] = 1
            states.append(v)
        states = np.asarray(states, dtype=np.int8)
    weights = (1 << np.arange(N, dtype=np.int64))
    key = ((states.astype(np.int64) + 1) // 2) @ weights
    order = np.argsort(key)
    key_sorted = key[order]
    dim = states.shape[0]
    rows, cols, vals = [np.arange(dim)], [np.arange(dim)], [apply_diag(states, aop_list)]
    for _, (s_conn, H_conn) in apply_off_diag(states, aop_list).items():
        valid = ~np.isnan(H_conn)
        r, c = np.nonzero(valid)
        k2 = ((s_conn[r, c].astype(np.int64) + 1) // 2) @ weights
        pos = np.searchsorted(key_sorted, k2)
        assert np.all(key_sorted[pos] == k2)
        rows.append(r), cols.append(order[pos]), vals.append(H_conn[r, c])
    Hm = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                       shape=(dim, dim))
    assert abs(Hm - Hm.T).max() < 1e-12
    if dim <= 600:
        return np.linalg.eigvalsh(Hm.toarray())[:k]
    w = spla.eigsh(Hm, k=k, which="SA", return_eigenvectors=False, tol=1e-12)
    return np.sort(w)
