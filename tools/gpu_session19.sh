#!/bin/bash
set -u
OUT=gpurun_out
rm -f $OUT/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s19_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s19_pytest.log
tail -n 15 $OUT/s19_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s19_bench.json 2> $OUT/s19_bench.err
echo "bench rc=$?"; head -c 1500 $OUT/s19_bench.json; tail -n 5 $OUT/s19_bench.err
