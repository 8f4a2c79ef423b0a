"""Timing probe for the ResConv kernels at the BASELINE config C / E shapes (dev tool, GPU box)."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx  # noqa: E402


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    warnings.simplefilter("ignore")
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, L, C, nb, ns in (("C 10x10 C=32", 10, 32, 8, 8192), ("E 16x16 C=88", 16, 88, 8, 2048)):
        if only and not name.startswith(only):
            continue
        qtx.sites.Sites._SITES = None
        qtx.sites.Square(L, Nparticles=(L * L // 2, L * L // 2))
        model = qtx.model.ResConv(nb, C, 3)
        state = qtx.state.Variational(model)
        s = qtx.utils.rand_states(ns)
        N = L * L
        flops = 2.0 * N * (C * 9 + (2 * nb - 1) * C * C * 9) * ns
        t = timeit(lambda: state(s))
        print(f"{name}: forward ns={ns}: {t:8.2f} ms  {flops / t / 1e9:7.2f} TFLOP/s (fp32 FMA)  Np={model.nparams}", flush=True)
        nj = ns // 8
        out = torch.empty((nj, model.nparams), dtype=torch.float64, device="cuda")
        tj = timeit(lambda: state.jacobian(s[:nj], out=out), 2)
        print(f"{name}: jacobian ns={nj}: {tj:8.2f} ms  {3 * flops / 8 / tj / 1e9:7.2f} TFLOP/s-equivalent (3x forward flops)", flush=True)


if __name__ == "__main__":
    main()
