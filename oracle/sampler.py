"""Oracle: Metropolis sweep (LocalFlip / SpinExchange) with injected or Philox randoms.

Test infrastructure.  Restates
  quantax/sampler/metropolis.py:246-275 (_partial_sweep/_single_sweep),
  quantax/sampler/metropolis.py:291-322 (_update: accept rule, select),
  quantax/sampler/common_samplers.py:28-33 (LocalFlip.propose),
  quantax/sampler/common_samplers.py:58-82 (_propose_exchange),
  quantax/sampler/sampler.py:66-69 (reweight factor).

The reference draws its randoms from jax threefry streams (third-party, unpinned).  Parity
is therefore defined on INJECTED proposal indices and uniforms; production kernels use the
Philox4x32-10 stream restated in ``philox4x32`` below, so the production path is bit-checkable
against this oracle as well.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Philox4x32-10 (Salmon et al., SC'11).  All arguments broadcastable uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for r in range(rounds):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def philox_draws(seed, step, chain_ids):
    """The production stream: counter = (chain, step_lo, step_hi, 0), key = seed (64 bit).
    Returns r0, r1 (uint32) and a 53-bit uniform u in [0,1)."""
    chain_ids = np.asarray(chain_ids, dtype=np.uint32)
    r0, r1, r2, r3 = philox4x32(chain_ids, np.uint32(step & 0xFFFFFFFF), np.uint32((step >> 32) & 0xFFFFFFFF),
                                np.uint32(0), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = (r2.astype(np.uint64) << np.uint64(32)) | r3.astype(np.uint64)
    u = (bits >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return r0, r1, u


def mulhi32(r, n):
    return ((r.astype(np.uint64) * np.asarray(n).astype(np.uint64)) >> np.uint64(32)).astype(np.int64)


def kth_hop_site(spins, hop, k):
    """Index of the k-th (0-based) site with spins == hop, per chain."""
    mask = spins == hop
    cs = np.cumsum(mask, axis=1)
    return np.argmax(cs > k[:, None], axis=1)


def philox_proposal(kind, seed, step, chain_ids, spins, hop=1, max_nb=0):
    r0, r1, u = philox_draws(seed, step, chain_ids)
    N = spins.shape[1]
    if kind == "localflip":
        return mulhi32(r0, N), None, u
    nhop = (spins == hop).sum(axis=1)  # per chain: a MixSampler with LocalFlip does not conserve the magnetisation
    pos = kth_hop_site(spins, hop, mulhi32(r0, nhop))
    return pos, mulhi32(r1, max_nb), u


def propose_localflip(spins, pos):
    """common_samplers.py:28-33."""
    new = spins.copy()
    ar = np.arange(spins.shape[0])
    new[ar, pos] *= -1
    return new


def propose_exchange(spins, pos, slot, neighbors):
    """common_samplers.py:71-82: neighbour = table[pos, slot]; -1 -> the particle itself."""
    ar = np.arange(spins.shape[0])
    nb = neighbors[pos, slot]
    nb = np.where(nb == -1, pos, nb)
    new = spins.copy()
    p, q = spins[ar, pos], spins[ar, nb]
    new[ar, pos] = q
    new[ar, nb] = p
    return new


def accept_mask(psi_old, psi_new, u, reweight, spins_old, spins_new):
    """metropolis.py:299-316.  ratio = |psi'/psi| formed in the container, densified, then
    raised to ``reweight``; accept iff ratio > 1-u or |psi| == 0; and the proposal moved."""
    m0, e0 = psi_old
    m1, e1 = psi_new
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        rate = np.abs((m1 / m0) * np.exp(e1 - e0)) ** reweight
        zero_old = np.abs(m0 * np.exp(e0)) == 0.0
    accepted = (rate > 1.0 - u) | zero_old
    updated = np.any(spins_old != spins_new, axis=1)
    return accepted & updated, rate


class RBMChainModel:
    """use_ref path: cached theta, local updates (variational.py:285-299)."""

    def __init__(self, rbm):
        self.net = rbm

    def init(self, spins):
        theta = self.net.init_internal(spins)
        return self.net.psi_from_theta(theta), theta

    def step(self, new, old, nflips, theta):
        return self.net.ref_forward(new, old, nflips, theta)

    def final_psi(self, spins, psi):
        return self.net.forward(spins)  # metropolis.py:201-213: direct psi replaces local-update psi


class FullForwardChainModel:
    """non-RefModel path (ResConv): full forward per proposal (variational.py:383-384)."""

    def __init__(self, net):
        self.net = net

    def init(self, spins):
        return self.net.forward(spins), None

    def step(self, new, old, nflips, internal):
        return self.net.forward(new), None

    def final_psi(self, spins, psi):
        return psi


def sweep(model, spins, nsweeps, kind, reweight=2.0, neighbors=None, hop=1,
          pos=None, slot=None, u=None, seed=None, step0=0, chain0=0, record=False):
    """One ``_partial_sweep``.  Either inject pos/slot/u arrays of shape [nsweeps, ns] or give
    a Philox seed.  Returns dict(spins, psi, psi_chain, naccept, accept_log, near_tie)."""
    spins = np.array(spins, dtype=np.int8)
    ns = spins.shape[0]
    psi, internal = model.init(spins)
    nflips = 1 if kind == "localflip" else 2
    chain_ids = np.arange(chain0, chain0 + ns)
    naccept = np.zeros(ns, dtype=np.int64)
    log = np.zeros((nsweeps, ns), dtype=np.uint8) if record else None
    margin = np.full((nsweeps, ns), np.inf) if record else None
    for t in range(nsweeps):
        if seed is not None:
            p_t, s_t, u_t = philox_proposal(kind, seed, step0 + t, chain_ids, spins, hop,
                                            0 if neighbors is None else neighbors.shape[1])
        else:
            p_t, u_t = pos[t], u[t]
            s_t = None if slot is None else slot[t]
        new = propose_localflip(spins, p_t) if kind == "localflip" else propose_exchange(spins, p_t, s_t, neighbors)
        psi_new, internal_new = model.step(new, spins, nflips, internal)
        acc, rate = accept_mask(psi, psi_new, u_t, reweight, spins, new)
        if record:
            log[t] = acc
            with np.errstate(invalid="ignore"):
                margin[t] = np.where(np.any(spins != new, axis=1), np.abs(rate - (1.0 - u_t)) / np.maximum(rate, 1e-300), np.inf)
        spins = np.where(acc[:, None], new, spins)
        psi = (np.where(acc, psi_new[0], psi[0]), np.where(acc, psi_new[1], psi[1]))
        if internal is not None:
            internal = np.where(acc[:, None], internal_new, internal)
        naccept += acc
    psi_final = model.final_psi(spins, psi)
    return dict(spins=spins, psi=psi_final, psi_chain=psi, naccept=naccept, accept_log=log, margin=margin)


def mix_sweep(model, spins, choice, kinds, neighbors, hops, reweight=2.0, pos=None, slot=None, u=None, seed=None,
              step0=0, chain0=0, record=False):
    """MixSampler._partial_sweep (metropolis.py:411-428): step t is proposed by component ``choice[t]`` and applied to
    all chains.  ``kinds`` / ``neighbors`` / ``hops`` are per-component lists.  Runs of equal components are one
    ``sweep`` call (with a RefModel the cached internals are re-initialised at run boundaries, which does not change
    the chains: psi(local update) == psi(direct), tutorials/local_updates.ipynb:189)."""
    spins = np.array(spins, dtype=np.int8)
    logs, t, out = [], 0, None
    nacc = np.zeros(spins.shape[0], dtype=np.int64)
    choice = np.asarray(choice)
    while t < len(choice):
        i, n = int(choice[t]), 1
        while t + n < len(choice) and int(choice[t + n]) == i:
            n += 1
        sl = slice(t, t + n)
        out = sweep(model, spins, n, kinds[i], reweight, neighbors[i], hops[i],
                    None if pos is None else pos[sl], None if (slot is None or kinds[i] == "localflip") else slot[sl],
                    None if u is None else u[sl], seed, step0 + t, chain0, record)
        spins = out["spins"]
        nacc += out["naccept"]
        if record:
            logs.append(out["accept_log"])
        t += n
    out = dict(out)
    out["naccept"] = nacc
    out["accept_log"] = np.concatenate(logs, axis=0) if record else None
    return out


def reweight_factor(psi, reweight):
    """sampler.py:66-69, evaluated in log space as the containers do."""
    mult, expo = psi
    with np.errstate(divide="ignore"):
        la = (np.log(np.abs(mult)) + expo) * (2.0 - reweight)
    mx = la.max()
    w = np.exp(la - mx)
    return w / w.mean()


def rand_states(ns, N, nup=None, seed=0):
    """utils/basis.py:121-134 (distribution only; the jax stream is unpinned)."""
    rng = np.random.default_rng(seed)
    if nup is None:
        return (rng.integers(0, 2, size=(ns, N)) * 2 - 1).astype(np.int8)
    base = -np.ones(N, dtype=np.int8)
    base[:nup] = 1
    return np.stack([rng.permutation(base) for _ in range(ns)])
