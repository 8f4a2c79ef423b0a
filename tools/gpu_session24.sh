#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python tools/gram_order_probe.py > $OUT/s24_gram_order.log 2>&1
echo "probe rc=$?"; cat $OUT/s24_gram_order.log | tail -12
timeout 600 python -m pytest tests/test_gram_tc_gpu.py tests/test_solver_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s24_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/s24_pytest.log
timeout 900 python -m pytest tests/test_resconv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x -k "tutorial or tensor_core_jac" > $OUT/s24_pytest2.log 2>&1
echo "pytest2 rc=$?"; tail -n 8 $OUT/s24_pytest2.log; grep "4x4 Heisenberg" $OUT/parity_report.jsonl | tail -2
