#!/bin/bash
set -u
OUT=gpurun_out
for g in 0 1; do
  QTX_TC_EPI_GENERIC=$g timeout 300 python tools/resconv_probe.py E > $OUT/s35_fwd_g$g.log 2>&1; echo "generic=$g"; grep "forward" $OUT/s35_fwd_g$g.log | head -1
  QTX_TC_EPI_GENERIC=$g QTX_TC_DEBUG=1 timeout 300 python tools/resconv_probe.py E 2>&1 | grep "tc dbg" | head -1 | cut -c1-330
done
timeout 600 python - <<'PY' > $OUT/s35_equal.log 2>&1
import os, sys, torch
sys.path.insert(0, '.')
import quantax_b200 as qtx
qtx.sites.Sites._SITES = None
qtx.sites.Square(16)
model = qtx.model.ResConv(8, 88, 3, final_activation=qtx.nn.sinhp1_by_scale)
state = qtx.state.Variational(model)
s = qtx.utils.rand_states(701)
os.environ["QTX_TC_EPI_GENERIC"] = "0"; a = state(s)
os.environ["QTX_TC_EPI_GENERIC"] = "1"; b = state(s)
print("bit-identical:", torch.equal(a.significand, b.significand) and torch.equal(a.exponent, b.exponent))
PY
cat $OUT/s35_equal.log | tail -2
