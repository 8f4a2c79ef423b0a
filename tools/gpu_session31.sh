#!/bin/bash
# RASTER-layout tensor-core Jacobian: whole ResConv / complex / symmetry / baseline-shape suites, timing at the config C
# lattice, sanitizer
set -u
OUT=gpurun_out
timeout 1200 python -m pytest tests/test_resconv_gpu.py tests/test_complex_gpu.py tests/test_symmetry_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fullsize_gpu.py tests/test_solver_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/s31_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 $OUT/s31_pytest.log
timeout 300 python tools/tc_bwd_probe.py timeC > $OUT/s31_probe_timeC.log 2>&1
echo "timeC rc=$?"; tail -n 3 $OUT/s31_probe_timeC.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/tc_bwd_probe.py sanitize > $OUT/s31_sanitizer_memcheck_tc_bwd.log 2>&1
echo "memcheck rc=$?"; tail -n 3 $OUT/s31_sanitizer_memcheck_tc_bwd.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/tc_bwd_probe.py sanitize > $OUT/s31_sanitizer_racecheck_tc_bwd.log 2>&1
echo "racecheck rc=$?"; grep -c "Race reported" $OUT/s31_sanitizer_racecheck_tc_bwd.log; grep "Race reported" -A2 $OUT/s31_sanitizer_racecheck_tc_bwd.log | grep -v tmem_alloc | grep -v "^--" | head -6
