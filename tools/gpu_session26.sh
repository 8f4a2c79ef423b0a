#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gram_tc_gpu.py tests/test_solver_gpu.py tests/test_zz_solver_variants_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s26_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 $OUT/s26_pytest.log
timeout 600 python tools/gram_order_probe.py > $OUT/s26_gram_order.log 2>&1
echo "probe rc=$?"; cat $OUT/s26_gram_order.log | tail -12
