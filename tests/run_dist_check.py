"""Multi-GPU consistency check (launched by torchrun, one rank per GPU): a VMC step with the chains
sharded over P GPUs must reproduce the single-GPU step on the same global chains -- identical chains
(global Philox chain ids), same energies, same SR/MinSR step.  Prints DIST_CHECK_OK on success."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quantax_b200 as qtx  # noqa: E402


def vmc(model_kind, nsamples):
    qtx.set_random_seed(123)
    qtx.sites.Sites._SITES = None
    qtx.set_default_dtype(torch.complex128 if model_kind == "cplx" else torch.float64)
    if model_kind == "cplx":  # config D in miniature: triangular lattice, complex ResConv + Neel-120 phase, D6 x Z2
        qtx.sites.Triangular(6, Nparticles=(18, 18))
        H = qtx.operator.Heisenberg()
    else:
        qtx.sites.Square(4, Nparticles=(8, 8))
        H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    symm = None
    if model_kind == "rbm":
        model = qtx.model.RBM_Dense(features=48, dtype=torch.float64)
    elif model_kind == "cplx":
        class PhaseLayer(qtx.nn.RawInputLayer):
            def __call__(self, x, s):
                return x * qtx.nn.neel120_phase(s)

        model0 = qtx.model.ResConv(2, 4, 3, dtype=torch.float64, out_dtype=torch.complex128)
        model = qtx.nn.Sequential(model0.layers + (PhaseLayer(),))
        symm = qtx.symmetry.D6(center=(0, 0)) @ qtx.symmetry.SpinInverse()
    else:
        model = qtx.model.ResConv(2, 4, 3, dtype=torch.float64)
    state = qtx.state.Variational(model, symm=symm)
    sampler = qtx.sampler.SpinExchange(state, nsamples=nsamples, thermal_steps=40)
    opt = qtx.optimizer.SR(state, H)
    samples = sampler.sweep()
    step = opt.get_step(samples)
    state.update(step * 0.01)
    samples2 = sampler.sweep()
    return samples.spins, opt.energy, opt.VarE, step, samples2.spins, model.params.clone()


def main():
    warnings.simplefilter("ignore")
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    # QTX_DIST_CHECK_BACKEND=gloo: all ranks share GPU 0 (a 1-GPU box): the CUDA kernels, the sharded layout and the
    # rank split of the shifted solves run for real, the collectives go through gloo instead of NCCL
    backend = os.environ.get("QTX_DIST_CHECK_BACKEND", "nccl")
    local = int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ok = True
    cases = (("rbm", 64 * world), ("resconv", 32 * world), ("cplx", 16 * world))
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group("gloo")
    results = []
    for kind, ns in cases:
        spins, e, v, step, spins2, params = vmc(kind, ns)
        g1 = [torch.empty_like(spins) for _ in range(world)]
        g2 = [torch.empty_like(spins2) for _ in range(world)]
        dist.all_gather(g1, spins)
        dist.all_gather(g2, spins2)
        results.append((torch.cat(g1), torch.cat(g2), e, step, params))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        for (kind, ns), (s1, s2, e, step, params) in zip(cases, results):
            s_ref, e_ref, v_ref, step_ref, s2_ref, p_ref = vmc(kind, ns)  # world() == (0, 1) now
            same1, same2 = torch.equal(s1, s_ref), torch.equal(s2, s2_ref)
            de = abs(e - e_ref) / abs(e_ref)
            ds = float((step - step_ref).norm() / step_ref.norm())
            dp = float((params - p_ref).abs().max())
            print(f"{kind}: chains equal {same1}/{same2}, dE {de:.2e}, dstep {ds:.2e}, dparams {dp:.2e}", flush=True)
            ok &= same1 and same2 and de < 1e-12 and ds < 1e-9 and dp < 1e-10
    if rank == 0:
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)


if __name__ == "__main__":
    main()
