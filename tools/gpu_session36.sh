#!/bin/bash
set -u
OUT=gpurun_out
timeout 300 python tools/tc_bwd_probe.py small > $OUT/s36_probe_small.log 2>&1; echo "small rc=$?"; tail -n 6 $OUT/s36_probe_small.log | cut -c1-150
timeout 300 python tools/tc_bwd_probe.py raster > $OUT/s36_probe_raster.log 2>&1; echo "raster rc=$?"; tail -n 7 $OUT/s36_probe_raster.log | cut -c1-150
timeout 600 python tools/tc_bwd_parity_probe.py 8 8 88 1.8 > $OUT/s36_parity_E.log 2>&1; echo "parity rc=$?"; head -n 3 $OUT/s36_parity_E.log | cut -c1-150
for k in 1 0; do
  QTX_TC_PAIRK=$k timeout 300 python tools/resconv_probe.py E > $OUT/s36_fwd_k$k.log 2>&1; echo "pairk=$k"; grep "forward" $OUT/s36_fwd_k$k.log | head -1
done
timeout 900 python -m pytest tests/test_resconv_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s36_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 $OUT/s36_pytest.log
