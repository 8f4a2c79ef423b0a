"""Debug / timing probe of the tensor-core ResConv forward against the FP32 CUDA-core path.
usage: python tools/tc_debug.py L C nblocks ns [final]   (one case per process)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx  # noqa: E402


def main():
    L, C, nb, ns = (int(a) for a in sys.argv[1:5])
    final = sys.argv[5] if len(sys.argv) > 5 else "exp"
    torch.cuda.set_device(0)
    qtx.sites.Sites._SITES = None
    qtx.sites.Square(L, Nparticles=(L * L // 2, L * L - L * L // 2))
    fa = qtx.nn.exp_by_scale if final == "exp" else qtx.nn.sinhp1_by_scale
    qtx.set_random_seed(1)
    model = qtx.model.ResConv(nb, C, 3, final_activation=fa)
    # non-zero biases
    g = torch.Generator().manual_seed(0)
    for name, o, shape in model.layout:
        if name.endswith("bias"):
            model.params[o:o + shape[0]] = (0.1 * torch.randn(shape[0], generator=g)).to(model.params)
    state = qtx.state.Variational(model)
    rng = np.random.default_rng(3)
    s = torch.from_numpy((2 * rng.integers(0, 2, size=(ns, L * L)) - 1).astype(np.int8)).cuda()

    def run():
        psi = state(s)
        return (torch.log(psi.significand.abs()) + psi.exponent).cpu().numpy(), torch.sign(psi.significand).cpu().numpy()

    if os.environ.get("TC_ONLY"):
        os.environ["QTX_RESCONV_TC"] = "1"
        for _ in range(3):
            state(s)
        torch.cuda.synchronize()
        return
    os.environ["QTX_RESCONV_TC"] = "0"
    ref, sref = run()
    if os.environ.get("TC_F64REF"):
        m64 = qtx.model.ResConv(nb, C, 3, final_activation=fa, dtype=torch.float64, params=model.params.double())
        st64 = qtx.state.Variational(m64)
        p64 = st64(s)
        ref64 = (torch.log(p64.significand.abs()) + p64.exponent).cpu().numpy()
        d = np.abs(ref - ref64)
        print(f"fp32 CUDA-core path vs float64 model: max={d.max():.3e} mean={d.mean():.3e} std of diff={np.std(ref - ref64):.3e}")
        ref = ref64
    os.environ["QTX_RESCONV_TC"] = "1"
    for mode in (os.environ.get("TC_MODES", "1,all")).split(","):
        if mode == "all":
            os.environ.pop("QTX_TC_LAYERS_PER_LAUNCH", None)
        else:
            os.environ["QTX_TC_LAYERS_PER_LAUNCH"] = mode
        out, sg = run()
        d = np.abs(out - ref)
        print(f"L={L} C={C} nb={nb} ns={ns} {final} layers/launch={mode}: max|dlogpsi|={d.max():.3e} "
              f"mean={d.mean():.3e} signs_equal={np.array_equal(sg, sref)} nan={np.isnan(out).sum()} "
              f"worst_sample={int(d.argmax())}", flush=True)
    # timing
    for tc in ("0", "1"):
        os.environ["QTX_RESCONV_TC"] = tc
        os.environ.pop("QTX_TC_LAYERS_PER_LAUNCH", None)
        for _ in range(2):
            state(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            state(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * ns * L * L * (9 * C + (2 * nb - 1) * 9 * C * C)
        print(f"  tc={tc}: {ms:.3f} ms per forward of {ns} samples, {flops / ms * 1e-9:.1f} TFLOP/s (useful f32 flops)", flush=True)


if __name__ == "__main__":
    main()
