#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_resconv_gpu.py tests/test_complex_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/s23_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 $OUT/s23_pytest.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/tc_bwd_probe.py sanitize > $OUT/s23_sanitizer_memcheck_tc_bwd.log 2>&1
echo "memcheck rc=$?"; tail -n 4 $OUT/s23_sanitizer_memcheck_tc_bwd.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/tc_bwd_probe.py sanitize > $OUT/s23_sanitizer_racecheck_tc_bwd.log 2>&1
echo "racecheck rc=$?"; tail -n 4 $OUT/s23_sanitizer_racecheck_tc_bwd.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/tc_bwd_probe.py sanitize > $OUT/s23_sanitizer_synccheck_tc_bwd.log 2>&1
echo "synccheck rc=$?"; tail -n 4 $OUT/s23_sanitizer_synccheck_tc_bwd.log
