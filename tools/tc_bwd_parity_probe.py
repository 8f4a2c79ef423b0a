"""Dev probe (GPU box): the config-E Jacobian against the float64 evaluation of the same weights, for the three
backward variants (tensor-core tower + tensor-core wgrad | tensor-core tower + CUDA-core wgrad | CUDA cores)."""
import os
import sys
import warnings

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import quantax_b200 as qtx  # noqa: E402
from oracle import models as omodels, sampler as osmp  # noqa: E402
from tests.gpu_util import lattice_pair, to_np  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    NB = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    CH = int(sys.argv[3]) if len(sys.argv) > 3 else 88
    comps = sys.argv[4].split(",") if len(sys.argv) > 4 else ["0", "1.5", "2.2", "3.0", "4.4"]
    lattice_pair(qtx, "square", 16, (128, 128))
    net = omodels.ResConv.random((16, 16), NB, CH, 3, np.float32, seed=11, final="sinhp1")
    model = qtx.model.ResConv(NB, CH, 3, final_activation=qtx.nn.sinhp1_by_scale, dtype=torch.float32,
                              params=torch.from_numpy(net.params().copy()))
    state = qtx.state.Variational(model)
    s = osmp.rand_states(ns, 256, 128, seed=12)
    st = torch.from_numpy(s).cuda()
    blocks = [{k: (None if v is None else v.astype(np.float64)) for k, v in blk.items()} for blk in net.blocks]
    O64 = omodels.ResConv(blocks, net.shape, net.final).jacobian(s)
    sig64, ex64 = omodels.ResConv(blocks, net.shape, net.final).forward(s)
    l64 = np.log(np.abs(sig64)) + ex64
    base = {"QTX_RESCONV_TC_BWD": "1", "QTX_TC_WGRAD": "1", "QTX_TC_PRECISE_GELU": "0", "QTX_TC_TRUNC_COMP": "2.2"}
    variants = [("tc all, comp %s" % c, {"QTX_TC_TRUNC_COMP": c}) for c in comps]
    variants += [("tc tower + cuda wgrad, comp 2.2", {"QTX_TC_WGRAD": "0"}), ("cuda backward, comp 2.2", {"QTX_RESCONV_TC_BWD": "0"}),
                 ("cuda backward, comp 0", {"QTX_RESCONV_TC_BWD": "0", "QTX_TC_TRUNC_COMP": "0"})]
    for name, delta in variants:
        env = dict(base)
        env.update(delta)
        os.environ.update(env)
        psi = state(st)
        lg = np.log(np.abs(to_np(psi.significand))) + to_np(psi.exponent)
        ferr = np.abs(lg - l64).max() / max(1.0, np.abs(l64).max())
        O = to_np(state.jacobian(st))
        rows = np.linalg.norm(O - O64, axis=1) / np.linalg.norm(O64, axis=1)
        worst = np.abs(O - O64).max() / np.abs(O64).max()
        # error by parameter group (first layers have the longest backward path)
        err = np.linalg.norm(O - O64, axis=0)
        ref = np.linalg.norm(O64, axis=0)
        per_block = []
        off = 0
        for i in range(NB):
            n = CH * (1 if i == 0 else CH) * 9 + CH + CH * CH * 9 + (0 if i == NB - 1 else CH)
            per_block.append(float(np.linalg.norm(err[off:off + n]) / np.linalg.norm(ref[off:off + n])))
            off += n
        print(f"{name:32s} fwd {ferr:.2e} rows max {rows.max():.3e} mean {rows.mean():.3e} worst entry {worst:.3e} per block "
              + " ".join(f"{x:.1e}" for x in per_block), flush=True)


if __name__ == "__main__":
    main()
