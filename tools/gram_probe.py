"""Timing / accuracy probe for qtx_gram at the config-B shape (dev tool, run on the GPU box)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantax_b200.optimizer import gram, pinv_eig_solve  # noqa: E402


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ns, npar = 4096, 40400
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((ns, npar), dtype=torch.float64, device="cuda", generator=g) * torch.rand((ns, 1), dtype=torch.float64, device="cuda", generator=g)
    A -= A.mean(dim=0, keepdim=True)
    rows = torch.arange(0, ns, 64, device="cuda")
    Asub = A[rows].cpu().numpy().astype(np.longdouble)
    exact = (Asub @ Asub.T).astype(np.float64)  # 80-bit accumulate reference on a 64 x 64 sub-block
    nrm = np.linalg.norm(Asub.astype(np.float64), axis=1)
    den = np.outer(nrm, nrm)
    ref = A @ A.T
    print(f"cuBLAS dgemm: {timeit(lambda: A @ A.T):8.3f} ms   err vs exact {np.abs(ref[rows][:, rows].cpu().numpy() - exact).max() / den.max():.2e}")
    for s in (8, 7, 6, 5, 4, -1):
        T = gram(A, nslices=s)
        err = np.abs(T[rows][:, rows].cpu().numpy() - exact) / den
        print(f"nslices {s:2d}: {timeit(lambda: gram(A, nslices=s)):8.3f} ms   max err/(|ai||aj|) {err.max():.2e}  rel-to-diag {np.abs(T[rows][:, rows].cpu().numpy() - exact).max() / exact.diagonal().max():.2e}", flush=True)
    T = gram(A, nslices=8)
    b = torch.randn(ns, dtype=torch.float64, device="cuda", generator=g)
    print(f"qtx eigh+pinv (cusolver Dsyevd): {timeit(lambda: pinv_eig_solve(T.clone(), b, None, 0.0), 2):8.3f} ms")
    print(f"torch.linalg.eigh:               {timeit(lambda: torch.linalg.eigh(T), 2):8.3f} ms")
    Tf = T.float()
    print(f"torch.linalg.eigh float32:       {timeit(lambda: torch.linalg.eigh(Tf), 2):8.3f} ms")


if __name__ == "__main__":
    main()
