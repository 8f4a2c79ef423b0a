"""Multi-GPU check of the fused Gram + exchange path (launched by torchrun, one rank per GPU): the sum over
ranks of the column-shard Gram matrices computed by qtx_gram_push / qtx_peer_signal / qtx_gram_reduce
(quantax_b200/peer.py) must equal Gram + NCCL all-reduce, be exactly symmetric and carry the same bits on every
rank; a distributed MinSR step through it must reproduce the NCCL step.  Prints P2P_GRAM_OK on success."""
import os
import sys
import warnings

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantax_b200 as qtx  # noqa: E402
from quantax_b200 import peer  # noqa: E402
from quantax_b200.optimizer import gram, minnorm_pinv_eig  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (ns, columns per rank): ragged tiles, several 256-row blocks, two K chunks (> 74880 columns per rank)
    for ns, npc in ((300, 1000), (1000, 777), (512, 4096), (256, 80000)):
        pg = peer.peer_gram(ns)
        for rep in range(3):
            g = torch.Generator(device="cuda").manual_seed(1000 * ns + 10 * rep + rank)
            A = torch.randn((ns, npc), dtype=torch.float64, device="cuda", generator=g)
            A *= torch.exp(torch.randn((ns, 1), dtype=torch.float64, device="cuda", generator=g))
            T = pg.gram_allreduce(A)
            ref = gram(A)
            dist.all_reduce(ref)
            err = float((T - ref).abs().max() / ref.abs().max())
            sym = bool(torch.equal(T, T.t()))
            gathered = [torch.empty_like(T) for _ in range(world)]
            dist.all_gather(gathered, T)
            same = all(torch.equal(x, gathered[0]) for x in gathered)
            if rank == 0:
                print(f"ns {ns} cols/rank {npc} rep {rep}: rel err vs NCCL {err:.2e}, symmetric {sym}, "
                      f"bit-identical on all ranks {same}", flush=True)
            ok &= err < 1e-13 and sym and same
    # a distributed MinSR solve through the fused path against the NCCL path
    nl, npar = 96, 3001
    g = torch.Generator(device="cuda").manual_seed(77 + rank)
    A = torch.randn((nl, npar), dtype=torch.float64, device="cuda", generator=g) / (nl * world) ** 0.5
    b = torch.randn(nl, dtype=torch.float64, device="cuda", generator=g)
    solver = minnorm_pinv_eig(rtol=1e-10)
    x_nccl = solver(A, b)
    os.environ["QTX_GRAM_P2P"] = "1"
    x_p2p = solver(A, b)
    os.environ["QTX_GRAM_P2P"] = "0"
    dx = float((x_p2p - x_nccl).norm() / x_nccl.norm())
    if rank == 0:
        print(f"distributed MinSR step, fused vs NCCL: rel diff {dx:.2e}", flush=True)
    ok &= dx < 1e-9
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    peer.release_all()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("P2P_GRAM_OK" if int(flag.item()) == 1 else "P2P_GRAM_FAILED", flush=True)


if __name__ == "__main__":
    main()
