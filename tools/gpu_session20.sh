#!/bin/bash
set -u
OUT=gpurun_out
QTX_TC_DEBUG=1 timeout 300 python tools/tc_bwd_probe.py time > $OUT/s20_probe_time_dbg.log 2>&1
echo "dbg rc=$?"; grep "tc dbg" $OUT/s20_probe_time_dbg.log | head -4
timeout 600 python -m pytest tests/test_gram_tc_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s20_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 $OUT/s20_pytest.log
timeout 600 python bench.py --workload E --steps 2 --warmup 1 --no-cpu --no-peaks > $OUT/s20_bench_E.json 2> $OUT/s20_bench_E.err
echo "bench rc=$?"; python -c "
import json;d=json.load(open('$OUT/s20_bench_E.json'));print(d['ms_per_step'],d['minsr_step_ms'],d['minsr_phases_ms'])"; tail -n 3 $OUT/s20_bench_E.err
