#!/bin/bash
# end-of-round check of HEAD: whole GPU suite, smoke(), default bench
set -u
OUT=gpurun_out
rm -f $OUT/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s32_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s32_pytest.log
tail -n 4 $OUT/s32_pytest.log
cp $OUT/parity_report.jsonl $OUT/s32_parity_report.jsonl
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/s32_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 $OUT/s32_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s32_bench.json 2> $OUT/s32_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s32_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms','e2e')}); print(d['minsr_phases_ms']); print(d['roofline']['frac'], d['clocks'])
PY
