#!/bin/bash
set -u
OUT=gpurun_out
for k in 1 0 1 0; do
  QTX_TC_PAIRK=$k timeout 600 python bench.py --workload E --steps 2 --warmup 2 --no-cpu --no-peaks > $OUT/s37_bench_k$k.json 2> $OUT/s37_bench_k$k.err
  python -c "
import json;d=json.load(open('$OUT/s37_bench_k$k.json'));print('pairk=$k', round(d['value'],1), round(d['sweep_oloc_ms'],1), d['clocks']['sm_mhz'])"
done
