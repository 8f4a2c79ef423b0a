#!/bin/bash
# Round-2 GPU session 1: un-gated parity suite with the measured-error report, the new default bench line, the
# reference arm, a launch list and one full ncu capture of the forward tower.
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/s1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s1_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s1_pytest.log
tail -n 40 $OUT/s1_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s1_bench.json 2> $OUT/s1_bench.err
echo "bench rc=$?"; tail -c 3000 $OUT/s1_bench.json; tail -n 5 $OUT/s1_bench.err
QTX_REF_BUDGET_S=40 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/s1_ref.json 2> $OUT/s1_ref.err
echo "ref rc=$?"; cat $OUT/s1_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/s1_launches_E.csv \
  python bench.py --workload E --steps 1 --warmup 0 --no-cpu --no-peaks > $OUT/s1_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resconv_tc -s 2 -c 2 -o $OUT/s1_prof_resconv_tc \
  python tools/resconv_probe.py E > $OUT/s1_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
