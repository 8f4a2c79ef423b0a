#!/bin/bash
# Round-2 GPU session 4 (1 GPU): full parity suite after the solver / ADVICE changes, pipelined LDL^T timing, launch list.
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s4_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s4_pytest.log
tail -n 30 $OUT/s4_pytest.log
timeout 600 python tools/pinv_probe.py 2048 4096 16384 > $OUT/s4_pinv_probe.jsonl 2> $OUT/s4_pinv_probe.err
echo "probe rc=$?"; cat $OUT/s4_pinv_probe.jsonl; tail -n 5 $OUT/s4_pinv_probe.err
timeout 300 python bench.py --workload B --steps 5 --warmup 3 --no-cpu --no-peaks > $OUT/s4_bench_B.json 2> $OUT/s4_bench_B.err
echo "bench B rc=$?"; cat $OUT/s4_bench_B.json; tail -n 5 $OUT/s4_bench_B.err
QTX_PROBE_ONLY_LDLT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zldlt|ztrsv|dd_|shift_build|lanczos|matvec|zero_sync|tridiag" -c 1500 --csv --log-file $OUT/s4_launches_ldlt4096.csv \
  python tools/pinv_probe.py 4096 > $OUT/s4_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zldlt_step -s 20 -c 2 -o $OUT/s4_prof_zldlt_step \
  python tools/pinv_probe.py 4096 > $OUT/s4_ncu_zldlt.log 2>&1
echo "ncu rc=$?"
ls -la $OUT | tail -n 10
