#!/bin/bash
# Round-2 GPU session 3: fused look-ahead DMMA factorisation -- parity, timing, ncu of the step kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 900 python -m pytest tests/test_zz_solver_variants_gpu.py tests/test_solver_gpu.py tests/test_complex_gpu.py \
  "tests/test_rbm_gpu.py::test_rbm_conv_matches_oracle" tests/test_baseline_shapes_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/s3_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s3_pytest.log
tail -n 30 $OUT/s3_pytest.log
timeout 600 python tools/pinv_probe.py 2048 4096 8192 16384 > $OUT/s3_pinv_probe.jsonl 2> $OUT/s3_pinv_probe.err
echo "probe rc=$?"; cat $OUT/s3_pinv_probe.jsonl; tail -n 5 $OUT/s3_pinv_probe.err
timeout 300 python bench.py --workload B --steps 5 --warmup 3 --no-cpu --no-peaks > $OUT/s3_bench_B.json 2> $OUT/s3_bench_B.err
echo "bench B rc=$?"; cat $OUT/s3_bench_B.json; tail -n 5 $OUT/s3_bench_B.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/s3_launches_pinv4096.csv \
  python tools/pinv_probe.py 4096 > $OUT/s3_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zldlt_step -s 20 -c 2 -o $OUT/s3_prof_zldlt_step \
  python tools/pinv_probe.py 4096 > $OUT/s3_ncu_zldlt.log 2>&1
echo "ncu rc=$?"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py ldlt > $OUT/s3_sanitizer_memcheck_ldlt.log 2>&1
echo "memcheck rc=$?"; tail -n 4 $OUT/s3_sanitizer_memcheck_ldlt.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py ldlt > $OUT/s3_sanitizer_racecheck_ldlt.log 2>&1
echo "racecheck rc=$?"; tail -n 4 $OUT/s3_sanitizer_racecheck_ldlt.log
ls -la $OUT | tail -n 12
