"""Timing / accuracy probe of the soft pseudo-inverse routes (dev tool, GPU box): own LDL^T kernels (csrc/zldlt.cu)
against cuSOLVER syevd and cuSOLVER LU at the MinSR sizes of configs B (4096) and E (2048 per GPU slice, 16384 at
N = 8).  Prints one JSON line per size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantax_b200 import _lib  # noqa: E402
from quantax_b200 import optimizer as qopt  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def problem(n, npar, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((n, npar), dtype=torch.float64, device="cuda", generator=g)
    A *= torch.exp(-8.0 * torch.rand((1, npar), dtype=torch.float64, device="cuda", generator=g))
    A -= A.mean(dim=0, keepdim=True)
    A /= n ** 0.5
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    return A, b


def main():
    sizes = [int(v) for v in sys.argv[1:]] or [2048, 4096]
    for n in sizes:
        npar = min(4 * n, 16384)
        A, b = problem(n, npar, n)
        T = qopt.gram(A)
        res = {"n": n, "npar": npar}
        if n <= 8192:
            res["eigh_route_ms"], (y_e, _) = timed(lambda: qopt.pinv_eig_solve(T.clone(), b, None, 0.0), 2)
            x_e = qopt.matvec_t(A, y_e)
        res["lanczos_ms"], lam = timed(lambda: qopt.sym_absmax_eig(T, method="ldlt", nshifts=1))
        masks = [7, 1] if n <= 8192 else [1]
        for mask in masks:
            nsh = bin(mask).count("1")
            ws, wsz = qopt._pinv_workspace(n, "ldlt", nsh)
            ydd = torch.zeros((2, n), dtype=torch.float64, device="cuda")
            info = torch.zeros(1, dtype=torch.int32, device="cuda")
            for refine in (qopt.REFINE_STEPS, 0):
                def run():
                    _lib.call("qtx_pinv_ldlt_partial", _lib.ptr(T), n, _lib.ptr(b), -1.0, 0.0, _lib.ptr(lam), mask, refine,
                              _lib.ptr(ydd), 0, _lib.ptr(info), _lib.ptr(ws), wsz, _lib.stream())
                res[f"ldlt_partial_mask{mask}_refine{refine}_ms"], _ = timed(run, 2)
            res[f"info_mask{mask}"] = int(info.item())
        if n <= 8192:
            res["ldlt_route_ms"], (y_l, info) = timed(lambda: qopt.pinv_rational_solve(T, b, None, 0.0, method="ldlt"), 2)
            x_l = qopt.matvec_t(A, y_l)
            res["x_rel_diff_ldlt_vs_eigh"] = float((x_l - x_e).norm() / x_e.norm())
            if n <= 4096:
                res["lu_route_ms"], (y_r, _) = timed(lambda: qopt.pinv_rational_solve(T, b, None, 0.0, method="rational"), 2)
                res["x_rel_diff_ldlt_vs_lu"] = float((x_l - qopt.matvec_t(A, y_r)).norm() / x_e.norm())
            # a gapped problem: rtol well inside a gap -> both routes agree to rounding
            y_l2, _ = qopt.pinv_rational_solve(T, b, 1e-5, 0.0, method="ldlt")
            y_e2, _ = qopt.pinv_eig_solve(T.clone(), b, 1e-5, 0.0)
            res["x_rel_diff_rtol_1e-5"] = float((qopt.matvec_t(A, y_l2) - qopt.matvec_t(A, y_e2)).norm()
                                                / qopt.matvec_t(A, y_e2).norm())
        print(json.dumps(res), flush=True)
        del A, T


if __name__ == "__main__":
    main()
