#!/bin/bash
# full GPU suite + default bench after the Gram changes (equal K chunks, coalesced lower-triangle epilogue + mirror)
set -u
OUT=gpurun_out
rm -f $OUT/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s27_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s27_pytest.log
tail -n 6 $OUT/s27_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s27_bench.json 2> $OUT/s27_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s27_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms','e2e')}); print(d['minsr_phases_ms']); print(d['roofline']['frac'], d['clocks']); print(d['config_B']['value'], d['config_B']['minsr_step_ms'], d['config_B']['minsr_phases_ms'])
PY
tail -n 3 $OUT/s27_bench.err
