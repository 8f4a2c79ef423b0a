#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_resconv_gpu.py tests/test_baseline_shapes_gpu.py tests/test_fullsize_gpu.py tests/test_complex_gpu.py tests/test_symmetry_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s40_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/s40_pytest.log
for k in 0 1 0 1; do
  QTX_TC_KEEP_PAD=$k timeout 600 python bench.py --workload E --steps 2 --warmup 2 --no-cpu --no-peaks > $OUT/s40_bench_k$k.json 2> $OUT/s40_bench_k$k.err
  python -c "
import json;d=json.load(open('$OUT/s40_bench_k$k.json'));print('keep_pad=$k', round(d['value'],1), round(d['sweep_oloc_ms'],1), d['clocks']['sm_mhz'])"
done
