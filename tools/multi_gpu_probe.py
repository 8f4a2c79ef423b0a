#!/usr/bin/env python
"""Multi-GPU first-run probe (one process per GPU; env RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*), started by
bench.py --gpus N as one SUBPROCESS per rank after all measurements, or by torchrun on its own.  Rank 0 prints ONE
JSON line; nothing here is a bench value.

  1. fused Gram + exchange over NVLink peer memory (quantax_b200/peer.py; verified on 2 GPUs only so
     far) against Gram + NCCL all-reduce at this world size: equality, symmetry, same bits on all ranks, and the
     time of both at the config B shard shape (4096 rows x 40400 / P columns);
  2. the eigendecomposition-free pseudo-inverse with its three shifts split over the ranks against
     the replicated cuSOLVER eigh route at n = 4096: time of both, difference of y projected on the range."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, iters=4, warmup=2):
    ts = []
    for i in range(warmup + iters):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sum(ts) / len(ts)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def p2p_gram(rank, world):
    from quantax_b200 import peer
    from quantax_b200.optimizer import gram

    out = {"checks": []}
    ok = True
    for ns, npc in ((300, 1000), (512, 4096)):
        pg = peer.peer_gram(ns)
        g = torch.Generator(device="cuda").manual_seed(1000 * ns + rank)
        A = torch.randn((ns, npc), dtype=torch.float64, device="cuda", generator=g)
        T = pg.gram_allreduce(A)
        ref = gram(A)
        dist.all_reduce(ref)
        err = float((T - ref).abs().max() / ref.abs().max())
        sym = bool(torch.equal(T, T.t()))
        gathered = [torch.empty_like(T) for _ in range(world)]
        dist.all_gather(gathered, T)
        same = all(torch.equal(x, gathered[0]) for x in gathered)
        out["checks"].append({"ns": ns, "cols_per_rank": npc, "rel_err_vs_nccl": err, "symmetric": sym,
                              "bit_identical_on_all_ranks": same})
        ok &= err < 1e-13 and sym and same
    out["ok"] = bool(ok)
    ns, npc = 4096, 40400 // world
    A = torch.randn((ns, npc), dtype=torch.float64, device="cuda")
    pg = peer.peer_gram(ns)
    T = torch.empty((ns, ns), dtype=torch.float64, device="cuda")

    def nccl():
        gram(A, out=T)
        dist.all_reduce(T)

    out["shape"] = [ns, npc]
    out["gram_alone_ms"] = timed(lambda: gram(A, out=T))
    out["gram_plus_nccl_allreduce_ms"] = timed(nccl)
    out["fused_push_signal_reduce_ms"] = timed(lambda: pg.gram_allreduce(A))
    peer.release_all()
    return out


def rational_split(rank, world, n=4096, npar=8192):
    from quantax_b200 import optimizer as qopt

    g = torch.Generator(device="cuda").manual_seed(5)  # the same matrix on every rank
    A = torch.randn((n, npar), dtype=torch.float64, device="cuda", generator=g)
    A *= torch.exp(-14.0 * torch.rand((1, npar), dtype=torch.float64, device="cuda", generator=g))
    A -= A.mean(dim=0, keepdim=True)
    A /= n ** 0.5
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    T = qopt.gram(A)
    out = {"n": n, "shift_masks": qopt.rational_shift_masks(world)}
    out["eigh_route_replicated_ms"] = timed(lambda: qopt.pinv_eig_solve(T.clone(), b, 1e-6, 0.0), iters=2, warmup=1)
    out["rational_route_split_ms"] = timed(lambda: qopt.pinv_rational_solve(T, b, 1e-6, 0.0, replicated=True),
                                           iters=2, warmup=1)
    y_e, _ = qopt.pinv_eig_solve(T.clone(), b, 1e-6, 0.0)
    y_r, info = qopt.pinv_rational_solve(T, b, 1e-6, 0.0, replicated=True)
    x_e, x_r = qopt.matvec_t(A, y_e), qopt.matvec_t(A, y_r)
    out["x_rel_diff_rtol_1e-6"] = float((x_e - x_r).norm() / x_e.norm())
    gathered = [torch.empty_like(y_r) for _ in range(world)]
    dist.all_gather(gathered, y_r)
    out["bit_identical_on_all_ranks"] = all(torch.equal(v, gathered[0]) for v in gathered)
    out["info"] = int(info.item())
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = {"n_gpus": world}
    t0 = time.time()
    for name, fn in (("p2p_gram", p2p_gram), ("pinv_rational_split", rational_split)):
        try:
            res[name] = fn(rank, world)
        except Exception as e:  # noqa: BLE001 -- a probe: report, do not raise
            res[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            break  # the ranks may be out of step after a failure: stop here
    res["probe_s"] = round(time.time() - t0, 1)
    if rank == 0:
        print(json.dumps(res), flush=True)
    os._exit(0)  # skip collective teardown: a failed rank must not make the others wait


if __name__ == "__main__":
    main()
