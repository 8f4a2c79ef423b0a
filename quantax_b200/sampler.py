"""Metropolis samplers: LocalFlip and SpinExchange (a.k.a. NeighborExchange).

Mirrors quantax/sampler/: ``Samples`` (samples.py:9-74), ``Sampler`` (sampler.py:14-69),
``Metropolis`` (metropolis.py:76-322), ``LocalFlip`` / ``SpinExchange``
(common_samplers.py:14-162).  One process drives one GPU and owns ``nsamples / world_size``
chains; the whole ``nsweeps`` loop runs in a single kernel launch (``state.fused_sweep``).

Random numbers: the reference consumes jax threefry keys; here every (chain, step) pair owns
one Philox4x32-10 counter, keyed by the global seed.  Chains are numbered globally
(rank * local + i), so results do not depend on how many GPUs share the chains.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Union
from warnings import warn

import numpy as np
import torch

from . import _lib
from .global_defs import PARTICLE_TYPE, device, get_seed, get_sites, get_subkeys, world
from .utils import LogArray, ScaleArray, log_abs, rand_states


# QTX_DRIFT_CHECK=1: run the reference's post-sweep drift check (metropolis.py:201-213) after every sweep of a state
# with local updates; off by default because it costs a device -> host read-back per sweep (check_local_updates()
# does the same on demand)
import os as _os

DRIFT_CHECK = _os.environ.get("QTX_DRIFT_CHECK", "0") == "1"


@dataclass(frozen=True)
class Samples:
    """quantax/sampler/samples.py:11-74."""

    spins: torch.Tensor
    psi: object
    state_internal: object = None
    reweight_factor: Optional[torch.Tensor] = None

    @property
    def nsamples(self) -> int:
        return self.spins.shape[0]

    def __getitem__(self, idx):
        f = lambda x: x if x is None else x[idx]
        return Samples(f(self.spins), f(self.psi), f(self.state_internal), f(self.reweight_factor))


class Sampler:
    """quantax/sampler/sampler.py:14-69."""

    def __init__(self, state, nsamples: int, reweight: float = 2.0):
        rank, nworld = world()
        if nsamples % nworld != 0:
            raise ValueError("`nsamples` should be a multiple of the number of devices, but got "
                             f"{nsamples} samples and {nworld} devices.")
        self._state, self._nsamples, self._reweight = state, nsamples, float(reweight)
        self._rank, self._world = rank, nworld
        self._nlocal = nsamples // nworld

    state = property(lambda self: self._state)
    Nsites = property(lambda self: self._state.Nsites)
    Nmodes = property(lambda self: self._state.Nmodes)
    nsamples = property(lambda self: self._nsamples)
    nlocal = property(lambda self: self._nlocal)
    reweight = property(lambda self: self._reweight)

    def sweep(self) -> Samples:
        return NotImplemented

    def _get_reweight_factor(self, psi) -> torch.Tensor:
        """|psi|^(2-n) / <|psi|^(2-n)> evaluated in log space (sampler.py:66-69); the mean runs
        over ALL chains of the job."""
        if self._reweight == 2.0:
            return torch.ones(psi.mult.shape[0], dtype=torch.float64, device=psi.mult.device)
        la = log_abs(psi) * (2.0 - self._reweight)
        mx = la.max()
        if self._world > 1:
            import torch.distributed as dist

            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        w = torch.exp(la - mx)
        tot = w.sum()
        if self._world > 1:
            dist.all_reduce(tot)
        return w / (tot / self._nsamples)


class RandomSampler(Sampler):
    r"""Random configurations with equal probability (quantax/sampler/sampler.py:125-143): the reweight exponent is
    0, so the samples carry the factor |psi|^2 / <|psi|^2>.  One batched forward of the state per call."""

    def __init__(self, state, nsamples: int):
        super().__init__(state, nsamples, reweight=0.0)

    def sweep(self) -> Samples:
        full = rand_states(self.nsamples)
        lo = self._rank * self._nlocal
        spins = full[lo:lo + self._nlocal].contiguous()
        psi = self._state(spins)
        return Samples(spins, psi, None, self._get_reweight_factor(psi))


class Metropolis(Sampler):
    """quantax/sampler/metropolis.py:76-322."""

    _kind = None

    def __init__(self, state, nsamples: int, reweight: float = 2.0, thermal_steps: Optional[int] = None,
                 sweep_steps: Optional[int] = None, initial_spins: Optional[torch.Tensor] = None):
        super().__init__(state, nsamples, reweight)
        ptype = get_sites().particle_type
        if ptype not in tuple(self.particle_type):
            raise ValueError(f"Particle type {ptype.name} is not supported by {self.__class__.__name__}.")
        self._thermal_steps = 20 * self.Nmodes if thermal_steps is None else thermal_steps
        self._sweep_steps = 2 * self.Nmodes if sweep_steps is None else sweep_steps
        if initial_spins is not None:
            initial_spins = torch.as_tensor(initial_spins)
            if initial_spins.ndim == 1:
                initial_spins = initial_spins.repeat(self.nsamples, 1)
            else:
                initial_spins = initial_spins.reshape(self.nsamples, self.Nmodes)
            lo = self._rank * self._nlocal
            initial_spins = initial_spins[lo:lo + self._nlocal].to(device=device(), dtype=torch.int8).contiguous()
        self._initial_spins = initial_spins
        self._seed = get_subkeys()
        self._step = 0
        self._injected = None
        self.last_accept_log = None
        self.last_naccept = None
        self.reset()

    @property
    def particle_type(self):
        return (PARTICLE_TYPE.spin,)

    @property
    def nflips(self) -> Optional[int]:
        return None

    def reset(self) -> None:
        """Reset all chains to ``initial_spins`` and thermalise them (metropolis.py:159-169)."""
        if self._initial_spins is None:
            full = rand_states(self.nsamples)
            lo = self._rank * self._nlocal
            self._spins = full[lo:lo + self._nlocal].contiguous()
        else:
            self._spins = self._initial_spins.clone()
        if self._thermal_steps > 0:
            self.sweep(self._thermal_steps)

    def propose(self, key, old_spins: torch.Tensor) -> torch.Tensor:
        """Propose new configurations for a batch of chains (metropolis.py:277-289, common_samplers.py:28-33,
        157-162).  ``key`` is a ``(seed, step)`` pair -- or an int seed with step 0 -- naming the Philox
        counter the draw comes from (the reference passes a jax PRNG key); row r uses chain id r.  One launch
        of qtx_metropolis_propose; ``sweep`` does not go through this method (its proposals are drawn inside
        the fused sweep kernels from the same stream)."""
        from . import _lib

        if self._kind is None:
            raise NotImplementedError
        seed, step = (key, 0) if isinstance(key, int) else key
        s = old_spins.to(device=device(), dtype=torch.int8).contiguous()
        ns, N = s.shape
        nbr, max_nb, hop = self._proposal_tables()
        new = torch.empty_like(s)
        moved = torch.empty(ns, dtype=torch.uint8, device=s.device)
        _lib.call("qtx_metropolis_propose", int(self._kind), _lib.ptr(s), ns, N, _lib.ptr(nbr), int(max_nb), int(hop), None,
                  None, int(seed), int(step), 0, _lib.ptr(new), _lib.ptr(moved), _lib.stream())
        return new

    def inject(self, pos: torch.Tensor, u: torch.Tensor, slot: Optional[torch.Tensor] = None) -> None:
        """Parity hook: the NEXT sweep uses these proposal sites [nsweeps, nlocal], neighbour-table
        columns (exchange) and acceptance uniforms instead of the Philox stream."""
        self._injected = (pos, slot, u)

    def _proposal_tables(self):
        return None, 0, 1

    def sweep(self, nsweeps: Optional[int] = None, record: bool = False) -> Samples:
        """Generate new samples (metropolis.py:171-215)."""
        from .global_defs import nvtx_range

        with nvtx_range("qtx.sweep"):
            return self._sweep(nsweeps, record)

    def _sweep(self, nsweeps: Optional[int] = None, record: bool = False) -> Samples:
        if nsweeps is None:
            nsweeps = self._sweep_steps
        state = self._state
        nbr, max_nb, hop = self._proposal_tables()
        injected, self._injected = self._injected, None
        psi, psi_chain, nacc, log = state.fused_sweep(
            self._spins, nsweeps, self._kind, nbr, max_nb, hop, self._reweight, self._seed, self._step,
            self._rank * self._nlocal, injected, record)
        self._step += nsweeps
        self.last_accept_log, self.last_naccept, self.last_psi_chain = log, nacc, psi_chain
        samples = Samples(self._spins.clone(), psi, None, self._get_reweight_factor(psi))
        if DRIFT_CHECK and state.use_ref:
            # the reference compares the local-update amplitudes with a direct forward pass after EVERY sweep and
            # warns (metropolis.py:201-213); that read-back synchronises the host, so it is behind QTX_DRIFT_CHECK=1
            self.check_local_updates(samples)
        return samples

    def check_local_updates(self, samples: Samples) -> int:
        """The reference's post-sweep drift check (metropolis.py:201-212), on demand: it needs a
        host sync, which the fused sweep otherwise avoids."""
        a, b = self.last_psi_chain.value(), samples.psi.value()
        close = ((a - b).abs() < 1e-8) | ((a / b - 1).abs() < 1e-3)
        ndiff = int((~close).sum().item())
        if ndiff > 0 and self._rank == 0:
            warn(f"{ndiff} out of {self.nsamples} wavefunctions are not close in direct forward pass and local "
                 "updates. This may indicate inaccurate local updates.")
        return ndiff


class LocalFlip(Metropolis):
    """Single spin flips (quantax/sampler/common_samplers.py:14-33)."""

    _kind = _lib.QTX_LOCAL_FLIP

    @property
    def nflips(self) -> int:
        return 1


def _site_neighbors(n_neighbor) -> np.ndarray:
    """[N, max_nb] int32 neighbour table, ascending, -1 padded (common_samplers.py:36-51)."""
    sites = get_sites()
    shells = [n_neighbor] if isinstance(n_neighbor, int) else list(n_neighbor)
    pairs = np.concatenate(sites.get_neighbor(shells), axis=0)
    adj = np.zeros((sites.Nsites, sites.Nsites), dtype=bool)
    adj[pairs[:, 0], pairs[:, 1]] = True
    adj |= adj.T
    width = int(adj.sum(axis=1).max())
    table = np.full((sites.Nsites, width), -1, dtype=np.int32)
    for i, row in enumerate(adj):
        nz = np.flatnonzero(row)
        table[i, :nz.size] = nz
    return table


class SpinExchange(Metropolis):
    """Neighbour spin exchange at fixed magnetisation (quantax/sampler/common_samplers.py:85-162)."""

    _kind = _lib.QTX_SPIN_EXCHANGE

    def __init__(self, state, nsamples: int, reweight: float = 2.0, thermal_steps: Optional[int] = None,
                 sweep_steps: Optional[int] = None, initial_spins: Optional[torch.Tensor] = None,
                 n_neighbor: Union[int, Sequence[int]] = 1):
        sites = get_sites()
        if isinstance(sites.Nparticles, int):
            raise ValueError("The number spin-up and spin-down particles should be specified in sites for "
                             "`SpinExchange` sampler.")
        self._hopping_particle = 1 if 2 * sites.Nparticles[0] <= state.Nmodes else -1
        table = _site_neighbors(n_neighbor)
        self._neighbors_host = table
        self._neighbors = torch.from_numpy(table).to(device()).contiguous()
        super().__init__(state, nsamples, reweight, thermal_steps, sweep_steps, initial_spins)

    @property
    def nflips(self) -> int:
        return 2

    def _proposal_tables(self):
        return self._neighbors, self._neighbors.shape[1], self._hopping_particle


class MixSampler(Metropolis):
    """A mixture of Metropolis samplers (quantax/sampler/metropolis.py:325-428): every sweep step is proposed by
    ONE component sampler, drawn with probability proportional to its ``nsamples``, and applied to all chains.
    Consecutive steps of the same component run as one fused sweep.  The component choice comes from a NumPy Philox
    stream keyed on (seed, step) -- the same on every rank -- or from ``inject_choice`` (parity hook)."""

    def __init__(self, samplers: Sequence[Metropolis], reweight: float = 2.0, thermal_steps: Optional[int] = None,
                 sweep_steps: Optional[int] = None, initial_spins: Optional[torch.Tensor] = None):
        state = samplers[0].state
        for sampler in samplers[1:]:
            if sampler.state is not state:
                raise ValueError("The states of component samplers should be the same in `MixSampler`.")
        self._samplers = tuple(samplers)
        nsamples = np.array([sampler.nsamples for sampler in samplers])
        total = int(nsamples.sum())
        self._ratio = nsamples / total
        self._choice = None
        super().__init__(state, total, reweight, thermal_steps, sweep_steps, initial_spins)

    @property
    def nflips(self) -> Optional[int]:
        nflips = tuple(sampler.nflips for sampler in self._samplers)
        return None if None in nflips else max(nflips)

    def reset(self) -> None:
        """metropolis.py:364-375: the first reset concatenates the (thermalised) chains of the components."""
        if hasattr(self, "_spins") or self._initial_spins is not None:
            super().reset()
            return
        self._spins = torch.cat([spl._spins for spl in self._samplers], dim=0).contiguous()
        if self._thermal_steps > 0:
            self.sweep(self._thermal_steps)

    def inject_choice(self, idx) -> None:
        """Parity hook: component index of every step of the NEXT sweep."""
        self._choice = np.asarray(idx, dtype=np.int64)

    def _draw_choice(self, nsweeps: int) -> np.ndarray:
        rng = np.random.Generator(np.random.Philox(key=int(self._seed) & 0xFFFFFFFFFFFFFFFF, counter=int(self._step)))
        return rng.choice(len(self._samplers), size=nsweeps, p=self._ratio)

    def sweep(self, nsweeps: Optional[int] = None, record: bool = False) -> Samples:
        if nsweeps is None:
            nsweeps = self._sweep_steps
        choice, self._choice = self._choice, None
        if choice is None:
            choice = self._draw_choice(nsweeps)
        if len(choice) != nsweeps:
            raise ValueError("inject_choice needs one component index per sweep step")
        injected, self._injected = self._injected, None
        state = self._state
        logs, nacc_tot, psi, psi_chain = [], None, None, None
        t = 0
        while t < nsweeps:
            i = int(choice[t])
            n = 1
            while t + n < nsweeps and int(choice[t + n]) == i:
                n += 1
            spl = self._samplers[i]
            nbr, max_nb, hop = spl._proposal_tables()
            inj = None
            if injected is not None:
                pos, slot, u = injected
                inj = (pos[t:t + n], None if (slot is None or spl._kind == _lib.QTX_LOCAL_FLIP) else slot[t:t + n],
                       u[t:t + n])
            psi, psi_chain, nacc, log = state.fused_sweep(self._spins, n, spl._kind, nbr, max_nb, hop, self._reweight,
                                                           self._seed, self._step, self._rank * self._nlocal, inj, record)
            self._step += n
            nacc_tot = nacc if nacc_tot is None else nacc_tot + nacc
            if record:
                logs.append(log)
            t += n
        if psi is None:  # nsweeps == 0
            psi = psi_chain = state(self._spins)
        self.last_accept_log = torch.cat(logs, dim=0) if record and logs else None
        self.last_naccept, self.last_psi_chain = nacc_tot, psi_chain
        return Samples(self._spins.clone(), psi, None, self._get_reweight_factor(psi))


NeighborExchange = SpinExchange  # pre-0.2 name (docs/.doctrees/sampler/quantax.sampler.NeighborExchange)
