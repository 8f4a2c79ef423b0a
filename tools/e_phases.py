"""Phase timing of the config-E slice (2048 chains, 16x16 J1-J2 ResConv 8x88): where sweep + Oloc time goes."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx

torch.cuda.set_device(0)
L, NS = 16, 2048
qtx.set_random_seed(42)
qtx.sites.Square(L, Nparticles=(L * L // 2, L * L // 2))
H = qtx.operator.Heisenberg(J=[1, 0.5], n_neighbor=[1, 2], msr=True)
model = qtx.model.ResConv(8, 88, 3, final_activation=qtx.nn.sinhp1_by_scale)
state = qtx.state.Variational(model)
sampler = qtx.sampler.SpinExchange(state, nsamples=NS, thermal_steps=32)

def timed(fn, reps=1):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, (time.perf_counter() - t0) * 1e3 / reps, out

s = sampler._spins
fwd, _, _ = timed(lambda: state(s), 10)
sw, sw_wall, samples = timed(lambda: sampler.sweep(128))
ol, ol_wall, _ = timed(lambda: H.Oloc(state, samples))
def enum():
    return [H.get_conn(samples.spins, nf) for nf in H.group_tables]
en, _, conn = timed(enum)
nconn = sum(c[0].numel() for c in conn)
big = conn[0][2]
f_big, _, _ = timed(lambda: state(big))
print(f"forward(2048) {fwd:.3f} ms | sweep(128 steps) {sw:.1f} ms = {sw/128:.3f} ms/step (wall {sw_wall:.1f}) | "
      f"Oloc {ol:.1f} ms (wall {ol_wall:.1f}): {nconn} connected configs = {nconn/NS:.0f} per sample, enumeration {en:.1f} ms, "
      f"forward of all {big.shape[0]} configs {f_big:.1f} ms = {f_big/big.shape[0]*2048:.3f} ms per 2048")
