"""One small pass over the hot-path entry points for compute-sanitizer (memcheck / racecheck / synccheck):
RBM sweep + Oloc + Jacobian, ResConv tensor-core forward + sweep + Jacobian, tensor-core Gram, the own LDL^T
pseudo-inverse, one full SR step.  Sizes are small: the sanitizer slows kernels by 10-100x.
Usage (GPU box): compute-sanitizer --tool memcheck python tools/sanitize_probe.py [part ...]"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx  # noqa: E402
from quantax_b200 import optimizer as qopt  # noqa: E402


def rbm():
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(6, Nparticles=(18, 18))
    H = qtx.operator.Heisenberg(msr=True)
    state = qtx.state.Variational(qtx.model.RBM_Dense(features=72))
    sampler = qtx.sampler.SpinExchange(state, nsamples=96, thermal_steps=10)
    opt = qtx.optimizer.SR(state, H)
    step = opt.get_step(sampler.sweep(20))
    state.update(step * 0.01)
    torch.cuda.synchronize()
    return float(opt.energy)


def resconv():
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(8, Nparticles=(32, 32))
    H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    state = qtx.state.Variational(qtx.model.ResConv(2, 16, 3, final_activation=qtx.nn.sinhp1_by_scale))
    sampler = qtx.sampler.SpinExchange(state, nsamples=32, thermal_steps=2)
    opt = qtx.optimizer.SR(state, H)
    step = opt.get_step(sampler.sweep(4))
    state.update(step * 0.01)
    torch.cuda.synchronize()
    return float(opt.energy)


def gram():
    A = torch.randn((300, 4100), dtype=torch.float64, device="cuda")
    T = qopt.gram(A)
    torch.cuda.synchronize()
    return float((T - A @ A.T).abs().max())


def ldlt():
    A = torch.randn((200, 700), dtype=torch.float64, device="cuda")
    A -= A.mean(dim=0, keepdim=True)
    T = qopt.gram(A, nslices=-1)
    b = torch.randn(200, dtype=torch.float64, device="cuda")
    y, info = qopt.pinv_rational_solve(T, b, 1e-9, 0.0, method="ldlt")
    y2, _ = qopt.pinv_eig_solve(T.clone(), b, 1e-9, 0.0)
    torch.cuda.synchronize()
    return float((A.T @ (y - y2)).norm() / (A.T @ y2).norm()), int(info.item())


def main():
    warnings.simplefilter("ignore")
    parts = sys.argv[1:] or ["rbm", "resconv", "gram", "ldlt"]
    for p in parts:
        print(p, globals()[p](), flush=True)
    print("SANITIZE_PROBE_DONE", flush=True)


if __name__ == "__main__":
    main()
