#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_resconv_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fullsize_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s42_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/s42_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s42_bench.json 2> $OUT/s42_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s42_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms')}, d['e2e']['value'], d['value_path'], d['clocks']['sm_mhz'])
PY
