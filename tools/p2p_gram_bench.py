"""Times the sum over ranks of the column-shard Gram matrices three ways (torchrun, one rank per GPU):
local Gram alone, Gram + NCCL all-reduce, and the fused push/signal/reduce path over NVLink peer memory
(quantax_b200/peer.py).  CUDA events on the launching stream, barrier before every iteration, max over ranks."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantax_b200 import peer  # noqa: E402
from quantax_b200.optimizer import gram  # noqa: E402


def timed(fn, iters=8, warmup=3):
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


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    shapes = [(4096, 40400 // world), (8192, 32768)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
    for ns, npc in shapes:
        A = torch.randn((ns, npc), dtype=torch.float64, device="cuda")
        pg = peer.peer_gram(ns)
        T = torch.empty((ns, ns), dtype=torch.float64, device="cuda")

        def nccl():
            gram(A, out=T)
            dist.all_reduce(T)

        import ctypes as C

        from quantax_b200 import _lib
        from quantax_b200.optimizer import _WS

        wsz = _lib.lib().qtx_gram_workspace_size(1, ns, npc, 0)
        ws = _WS.get("gram", wsz)

        def push_only():
            _lib.call("qtx_gram_push", 1, _lib.ptr2d(A), ns, npc, A.stride(0), 0, _lib.ptr(T), pg.P, pg.rank, pg._slots,
                      _lib.ptr(ws), wsz, _lib.stream())

        def reduce_only():  # flags not waited for: the kernel alone
            vp = C.c_void_p * pg.P
            partials = vp(*[T.data_ptr() if q == pg.rank else pg._mine[q] for q in range(pg.P)])
            _lib.call("qtx_gram_reduce", partials, pg.P, ns, _lib.ptr(T), None, 0, 0.0, _lib.stream())

        t_push = timed(push_only)
        t_reduce = timed(reduce_only)
        t_gram = timed(lambda: gram(A, out=T))
        t_ar = timed(lambda: dist.all_reduce(T))
        t_nccl = timed(nccl)
        t_fused = timed(lambda: pg.gram_allreduce(A))
        if rank == 0:
            print(json.dumps({"n_gpus": world, "ns": ns, "cols_per_rank": npc, "gram_alone_ms": t_gram,
                              "nccl_allreduce_alone_ms": t_ar, "gram_plus_nccl_allreduce_ms": t_nccl,
                              "fused_push_signal_reduce_ms": t_fused, "push_gram_kernel_alone_ms": t_push,
                              "reduce_kernel_alone_ms": t_reduce,
                              "T_bytes": ns * ns * 8}), flush=True)
    peer.release_all()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
