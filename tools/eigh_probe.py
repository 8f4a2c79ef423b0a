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
for n in (1024, 2048, 4096):
    A = torch.randn((n, 3 * n), dtype=torch.float64, device="cuda", generator=g)
    T = A @ A.T
    b = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    y, info = pinv_eig_solve(T.clone(), b, None, 0.0)
    res = (T @ y - b).norm() / b.norm()
    print(f"n={n} algo={os.environ.get('QTX_EIGH_ALGO','0')}: {timeit(lambda: pinv_eig_solve(T.clone(), b, None, 0.0)):8.2f} ms  residual {res:.1e} info {int(info)}", flush=True)
