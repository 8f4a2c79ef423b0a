#!/bin/bash
# Round-2 GPU session 6 (1 GPU): fragment skipping in the in-register diag / panel; ncu of the step kernel at large T.
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 900 python -m pytest tests/test_zz_solver_variants_gpu.py tests/test_solver_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/s6_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s6_pytest.log
tail -n 8 $OUT/s6_pytest.log
timeout 600 python tools/pinv_probe.py 2048 4096 16384 > $OUT/s6_pinv_probe.jsonl 2> $OUT/s6_pinv_probe.err
echo "probe rc=$?"; cat $OUT/s6_pinv_probe.jsonl; tail -n 5 $OUT/s6_pinv_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zldlt|ztrsv" -c 140 --csv --log-file $OUT/s6_launches_ldlt4096.csv \
  python tools/pinv_probe.py 4096 > $OUT/s6_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zldlt_step -s 3 -c 2 -o $OUT/s6_prof_zldlt_step_8192 \
  python tools/pinv_probe.py 8192 > $OUT/s6_ncu_zldlt.log 2>&1
echo "ncu rc=$?"
timeout 300 python bench.py --workload B --steps 5 --warmup 3 --no-cpu --no-peaks > $OUT/s6_bench_B.json 2> $OUT/s6_bench_B.err
echo "bench B rc=$?"; cat $OUT/s6_bench_B.json | cut -c1-1800; tail -n 5 $OUT/s6_bench_B.err
