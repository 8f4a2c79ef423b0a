#!/usr/bin/env python
"""First-run probe of code written without GPU access (run by bench.py in a SUBPROCESS after all measurements, so
that a crash or a time-out here cannot touch the bench line; also usable on its own under gpurun).

  1. the eigendecomposition-free pseudo-inverse (csrc/pinv_rational.cu) against the cuSOLVER eigh route at the
     MinSR size of config B (n = 4096): time of each (CUDA events), difference of the MinSR step x = A^T y;
  2. the GPU tests gated by QTX_UNVERIFIED=1 (tests/test_zz_solver_variants_gpu.py).

Prints ONE JSON line.  Nothing here is a bench value."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rational_probe(n=4096, npar=8192):
    import torch

    from quantax_b200 import optimizer as qopt

    torch.cuda.set_device(0)
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((n, npar), dtype=torch.float64, device="cuda", generator=g)
    A *= torch.exp(-14.0 * torch.rand((1, npar), dtype=torch.float64, device="cuda", generator=g))  # 12 decades in T
    A -= A.mean(dim=0, keepdim=True)
    A /= n ** 0.5
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) / n ** 0.5
    T = qopt.gram(A)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {"n": n, "npar": npar}

    def timed(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return r, e0.elapsed_time(e1) / reps

    (y_e, evals, info_e), out["eigh_route_ms"] = timed(lambda: qopt.pinv_eig_solve(T.clone(), b, None, 0.0, want_evals=True))
    lam, out["lanczos_ms"] = timed(lambda: qopt.sym_absmax_eig(T))
    out["lanczos_rel_err"] = float((lam[0] - evals.abs().max()).abs() / evals.abs().max())
    (y_r, info_r), out["rational_route_ms"] = timed(lambda: qopt.pinv_rational_solve(T, b, None, 0.0))
    keep = qopt.REFINE_STEPS
    qopt.REFINE_STEPS = 0  # Lanczos + 3 x (build, Zgetrf, one Zgetrs): what the refinement adds is the difference
    _, out["rational_route_no_refinement_ms"] = timed(lambda: qopt.pinv_rational_solve(T, b, None, 0.0))
    qopt.REFINE_STEPS = keep
    out["refine_steps"] = keep
    out["info"] = [int(info_e.item()), int(info_r.item())]
    x_e, x_r = qopt.matvec_t(A, y_e), qopt.matvec_t(A, y_r)
    out["x_rel_diff_default_rtol"] = float((x_e - x_r).norm() / x_e.norm())
    out["eigenvalues_below_cutoff"] = int((evals.abs() < 1e-12 * evals.abs().max()).sum().item())
    # a cut-off well inside the spectrum gap-free region but far from eps |T|: both routes are accurate there
    y_e2, _ = qopt.pinv_eig_solve(T.clone(), b, 1e-6, 0.0)
    y_r2, _ = qopt.pinv_rational_solve(T, b, 1e-6, 0.0)
    x_e2, x_r2 = qopt.matvec_t(A, y_e2), qopt.matvec_t(A, y_r2)
    out["x_rel_diff_rtol_1e-6"] = float((x_e2 - x_r2).norm() / x_e2.norm())
    return out


def main():
    res = {}
    t0 = time.time()
    try:
        res["pinv_rational"] = rational_probe()
    except Exception as e:  # noqa: BLE001 -- a probe: report, do not raise
        res["pinv_rational"] = {"error": f"{type(e).__name__}: {e}"[:400]}
    res["probe_s"] = round(time.time() - t0, 1)
    try:
        env = dict(os.environ, QTX_UNVERIFIED="1")
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_zz_solver_variants_gpu.py"),
                            "-m", "gpu", "-q", "--no-header", "-p", "no:cacheprovider", "-k",
                            "apply_off_diag or propose_method or random_sampler or bit_for_bit or lanczos or rational",
                            "--tb=line"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=150)
        tail = [ln for ln in r.stdout.strip().splitlines() if ln.strip()][-12:]
        res["unverified_tests"] = {"rc": r.returncode, "tail": [ln[:300] for ln in tail]}
    except Exception as e:  # noqa: BLE001
        res["unverified_tests"] = {"error": f"{type(e).__name__}: {e}"[:400]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
