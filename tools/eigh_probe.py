"""Wall time of the cuSOLVER eigh + pseudo-inverse step (qtx_pinv_eig_solve) at the MinSR sizes of the configs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantax_b200.optimizer import pinv_eig_solve
def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
g = torch.Generator(device="cuda").manual_seed(0)
sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]
for n in sizes:
    A = torch.randn((n, 2 * n), dtype=torch.float64, device="cuda", generator=g)
    T = A @ A.T
    del A
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y, info = pinv_eig_solve(T.clone(), b, None, 0.0)
    res = (T @ y - b).norm() / b.norm()
    print(f"n={n}: {timeit(lambda: pinv_eig_solve(T.clone(), b, None, 0.0), reps=1 if n > 8000 else 2):8.2f} ms  residual {res:.1e} info {int(info)}", flush=True)
