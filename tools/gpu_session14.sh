#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/tc_bwd_probe.py small > $OUT/s14_probe_small.log 2>&1
echo "small rc=$?"; tail -n 20 $OUT/s14_probe_small.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/s14_launches_jac_E.csv \
  python tools/tc_bwd_probe.py time > $OUT/s14_ncu_launch.log 2>&1
echo "ncu rc=$?"; tail -n 5 $OUT/s14_ncu_launch.log
timeout 600 python -m pytest tests/test_resconv_gpu.py tests/test_baseline_shapes_gpu.py tests/test_complex_gpu.py tests/test_symmetry_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s14_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 $OUT/s14_pytest.log
