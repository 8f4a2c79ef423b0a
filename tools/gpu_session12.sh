#!/bin/bash
# Round-2 GPU session 12: whole GPU suite, default bench line, reference arm, racecheck of the LDL^T route after the pivot fix
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s12_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s12_pytest.log
tail -n 15 $OUT/s12_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s12_bench.json 2> $OUT/s12_bench.err
echo "bench rc=$?"; tail -c 1500 $OUT/s12_bench.json; tail -n 5 $OUT/s12_bench.err
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_probe.py ldlt > $OUT/s12_sanitizer_racecheck_ldlt.log 2>&1
echo "racecheck rc=$?"; tail -n 3 $OUT/s12_sanitizer_racecheck_ldlt.log
