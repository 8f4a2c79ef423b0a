#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 900 python -m pytest tests/test_zz_solver_variants_gpu.py tests/test_solver_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/s11_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s11_pytest.log
tail -n 8 $OUT/s11_pytest.log
timeout 600 python tools/pinv_probe.py 2048 4096 16384 > $OUT/s11_pinv_probe.jsonl 2> $OUT/s11_pinv_probe.err
echo "probe rc=$?"; cat $OUT/s11_pinv_probe.jsonl; tail -n 5 $OUT/s11_pinv_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zldlt|ztrsv" -c 140 --csv --log-file $OUT/s11_launches_ldlt4096.csv \
  python tools/pinv_probe.py 4096 > $OUT/s11_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python bench.py --workload B --steps 5 --warmup 3 --no-cpu --no-peaks > $OUT/s11_bench_B.json 2> $OUT/s11_bench_B.err
echo "bench B rc=$?"; cat $OUT/s11_bench_B.json | cut -c1-1700; tail -n 5 $OUT/s11_bench_B.err
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py ldlt > $OUT/s11_sanitizer_racecheck_ldlt.log 2>&1
echo "racecheck rc=$?"; tail -n 3 $OUT/s11_sanitizer_racecheck_ldlt.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py ldlt > $OUT/s11_sanitizer_memcheck_ldlt.log 2>&1
echo "memcheck rc=$?"; tail -n 3 $OUT/s11_sanitizer_memcheck_ldlt.log
