"""Dev probe (GPU box): one H.Oloc call at the config E shape (2048 samples, ~1.06 M connected configurations) and the
same number of plain forwards, timed with CUDA events; run under `ncu --metrics gpu__time_duration.sum` for the
kernel list of the Oloc call."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(16, Nparticles=(128, 128))
    H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
    model = qtx.model.ResConv(8, 88, 3, final_activation=qtx.nn.sinhp1_by_scale)
    state = qtx.state.Variational(model)
    s = qtx.utils.rand_states(2048)
    psi = state(s)
    samples = qtx.sampler.Samples(s, psi, None, torch.ones(2048, dtype=torch.float64, device="cuda"))
    H.Oloc(state, samples)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    H.Oloc(state, samples)
    b.record(); torch.cuda.synchronize()
    n = H.last_conn_count
    print(f"Oloc: {a.elapsed_time(b):9.2f} ms for {n} connected configurations ({a.elapsed_time(b) * 1e3 / n:6.3f} us each)", flush=True)
    for batch in (2048, 16384):
        big = qtx.utils.rand_states(batch)
        state(big); torch.cuda.synchronize()
        reps = max(1, n // batch)
        a.record()
        for _ in range(reps):
            state(big)
        b.record(); torch.cuda.synchronize()
        print(f"forward in batches of {batch}: {a.elapsed_time(b) / (reps * batch) * 1e3:6.3f} us per configuration", flush=True)


if __name__ == "__main__":
    main()
