"""Dev probe (GPU box): the tensor-core Jacobian (csrc/resconv_tc.cu: backward-data tower + per-sample weight
gradients) against the CUDA-core backward pass of csrc/resconv.cu on the same state, and its timing at the config E
shape.  Usage: python tools/tc_bwd_probe.py [small|E|time]"""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx  # noqa: E402


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def compare(name, shape, C, nb, ns, final="exp", seed=0):
    qtx.sites.Sites._SITES = None
    qtx.sites.Grid(list(shape))
    torch.manual_seed(seed)
    fa = qtx.nn.exp_by_scale if final == "exp" else qtx.nn.sinhp1_by_scale
    model = qtx.model.ResConv(nb, C, 3, final_activation=fa)
    state = qtx.state.Variational(model)
    s = qtx.utils.rand_states(ns)
    os.environ["QTX_RESCONV_TC_BWD"] = "1"
    O = state.jacobian(s).clone()
    torch.cuda.synchronize()
    os.environ["QTX_RESCONV_TC_BWD"] = "0"
    R = state.jacobian(s).clone()
    torch.cuda.synchronize()
    os.environ["QTX_RESCONV_TC_BWD"] = "1"
    rows = ((O - R).norm(dim=1) / R.norm(dim=1)).max().item()
    worst = ((O - R).abs().max() / R.abs().max()).item()
    nan = int(torch.isnan(O).sum().item())
    # per-layer breakdown of the worst entry
    col = (O - R).abs().max(dim=0).values
    top = int(col.argmax().item())
    print(f"{name}: shape={shape} C={C} nb={nb} ns={ns}: rows rel 2-norm {rows:.3e}  worst entry {worst:.3e}  nan={nan} "
          f"worst col {top} of {O.shape[1]}", flush=True)
    return rows, worst


def main():
    warnings.simplefilter("ignore")
    what = sys.argv[1] if len(sys.argv) > 1 else "small"
    if what in ("small", "all"):
        compare("seg16 small", (16, 16), 24, 2, 5)
        compare("seg16 sinhp1", (16, 16), 40, 3, 7, final="sinhp1")
        compare("seg16 C=88 nb=2", (16, 16), 88, 2, 9)
        compare("seg16 C=16 nb=1", (16, 16), 16, 1, 3)
        compare("seg 16x8 (one tile per sample)", (16, 8), 24, 2, 6)
        compare("seg 32x8", (32, 8), 20, 2, 5)
    if what in ("raster", "all"):  # RASTER layout (any width): one and two tiles per sample
        compare("raster 10x10 C=32 nb=3", (10, 10), 32, 3, 9, final="sinhp1")
        compare("raster 10x10 C=36 nb=2", (10, 10), 36, 2, 5)
        compare("raster 12x12 C=20 nb=2", (12, 12), 20, 2, 7)
        compare("raster 8x8 C=12 nb=2", (8, 8), 12, 2, 6)
        compare("raster 6x6 C=5 nb=3", (6, 6), 5, 3, 4)
        compare("raster 4x4 C=8 nb=2", (4, 4), 8, 2, 11)
        compare("raster 6x10 C=16 nb=2", (6, 10), 16, 2, 5)
    if what == "sanitize":  # small enough for compute-sanitizer
        compare("seg16 C=24 nb=2", (16, 16), 24, 2, 3)
        compare("seg 16x8 C=20 nb=2", (16, 8), 20, 2, 5, final="sinhp1")
        compare("raster 10x10 C=16 nb=2", (10, 10), 16, 2, 3)
        print("SANITIZE_PROBE_DONE", flush=True)
    if what in ("E", "all"):
        compare("config E", (16, 16), 88, 8, 64, final="sinhp1")
    if what in ("time", "all"):
        qtx.sites.Sites._SITES = None
        qtx.sites.Square(16)
        model = qtx.model.ResConv(8, 88, 3, final_activation=qtx.nn.sinhp1_by_scale)
        state = qtx.state.Variational(model)
        ns = 2048
        s = qtx.utils.rand_states(ns)
        out = torch.empty((ns, model.nparams), dtype=torch.float64, device="cuda")
        for flag in ("1", "0"):
            os.environ["QTX_RESCONV_TC_BWD"] = flag
            t = timeit(lambda: state.jacobian(s, out=out), 2)
            print(f"config E jacobian ns={ns} QTX_RESCONV_TC_BWD={flag}: {t:8.2f} ms", flush=True)
        os.environ["QTX_RESCONV_TC_BWD"] = "1"
        tf = timeit(lambda: state(s))
        print(f"config E forward ns={ns}: {tf:8.2f} ms", flush=True)
    if what in ("timeC", "all"):  # config C lattice (10x10, RASTER layout), 8 blocks of 32 channels
        qtx.sites.Sites._SITES = None
        qtx.sites.Square(10)
        model = qtx.model.ResConv(8, 32, 3)
        state = qtx.state.Variational(model)
        ns = 8192
        s = qtx.utils.rand_states(ns)
        out = torch.empty((ns, model.nparams), dtype=torch.float64, device="cuda")
        for flag in ("1", "0"):
            os.environ["QTX_RESCONV_TC_BWD"] = flag
            t = timeit(lambda: state.jacobian(s, out=out), 2)
            print(f"config C shape jacobian ns={ns} Np={model.nparams} QTX_RESCONV_TC_BWD={flag}: {t:8.2f} ms", flush=True)
        os.environ["QTX_RESCONV_TC_BWD"] = "1"


if __name__ == "__main__":
    main()
