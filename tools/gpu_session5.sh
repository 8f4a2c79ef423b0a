#!/bin/bash
# Round-2 GPU session 5 (1 GPU): fast pivot reciprocal + trsv prefetch timing; failing tests; one-call sweep test; E bench.
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 900 python -m pytest tests/test_zz_solver_variants_gpu.py tests/test_gram_tc_gpu.py tests/test_resconv_gpu.py tests/test_multigpu_gpu.py \
  tests/test_baseline_shapes_gpu.py tests/test_solver_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/s5_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s5_pytest.log
tail -n 25 $OUT/s5_pytest.log
timeout 600 python tools/pinv_probe.py 2048 4096 16384 > $OUT/s5_pinv_probe.jsonl 2> $OUT/s5_pinv_probe.err
echo "probe rc=$?"; cat $OUT/s5_pinv_probe.jsonl; tail -n 5 $OUT/s5_pinv_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zldlt|ztrsv|dd_|shift_build" -c 700 --csv --log-file $OUT/s5_launches_ldlt4096.csv \
  python tools/pinv_probe.py 4096 > $OUT/s5_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-peaks > $OUT/s5_bench.json 2> $OUT/s5_bench.err
echo "bench rc=$?"; tail -c 2500 $OUT/s5_bench.json; tail -n 5 $OUT/s5_bench.err
ls -la $OUT | tail -n 8
