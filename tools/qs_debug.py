import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx
torch.cuda.set_device(0)
qtx.set_random_seed(42)
qtx.sites.Sites._SITES = None
qtx.sites.Chain(8)
H = qtx.operator.Ising(h=1.0)
model = qtx.model.RBM_Dense(features=16)
state = qtx.state.Variational(model)
sampler = qtx.sampler.LocalFlip(state, nsamples=1024)
optimizer = qtx.optimizer.SR(state, H)
hist = []
for i in range(300):
    samples = sampler.sweep()
    step = optimizer.get_step(samples)
    state.update(step * 1e-2)
    hist.append(optimizer.energy)
    if i < 6 or i % 20 == 0:
        print(i, optimizer.energy, float(step.norm()), float(model.params.norm()), flush=True)
print("final", np.mean(hist[-20:]))
