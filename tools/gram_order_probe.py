"""Dev probe (GPU box): qtx_gram with the supertile tile order against row-major order (QTX_GRAM_SUPERTILE=0) at the
shapes of the benchmark: config B (4096 x 40400, 7 digits), the config E slice (2048 x 1047552, 5 digits) and the
per-rank shard of config E on 8 GPUs (16384 x 130944, 5 digits).  Results must be bit-identical."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantax_b200.optimizer import gram  # noqa: E402


def timeit(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, ns, npar, s, dt in (("config B", 4096, 40400, 7, torch.float64), ("config E slice", 2048, 1047552, 5, torch.float64),
                                  ("config E shard at N=8", 16384, 130944, 5, torch.float64)):
        if only and only not in name:
            continue
        A = torch.randn((ns, npar), dtype=dt, device="cuda", generator=g)
        A *= torch.rand((ns, 1), dtype=dt, device="cuda", generator=g)
        out = {}
        for flag in ("1", "0"):
            os.environ["QTX_GRAM_SUPERTILE"] = flag
            T = gram(A, nslices=s).clone()
            t = timeit(lambda: gram(A, nslices=s))
            out[flag] = T
            ops = s * (s + 1) / 2 * ns * (ns + 1) * npar
            print(f"{name}: {ns} x {npar}, {s} digits, supertile={flag}: {t:9.3f} ms  ({ops / t / 1e9:8.1f} int8 TOP/s executed)", flush=True)
        print(f"{name}: identical {torch.equal(out['0'], out['1'])}", flush=True)
        del A, out
        torch.cuda.empty_cache()
    os.environ.pop("QTX_GRAM_SUPERTILE", None)


if __name__ == "__main__":
    main()
