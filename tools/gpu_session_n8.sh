#!/bin/bash
# Round-2 multi-GPU session (8 GPUs = BASELINE configs[4], the north-star configuration): distributed step == single-GPU
# step on 8 ranks, then the bench line at N = 8
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/n8_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
  tests/run_dist_check.py > $OUT/n8_dist_check.log 2>&1
echo "dist check rc=$?"; grep -c DIST_CHECK_OK $OUT/n8_dist_check.log; tail -n 6 $OUT/n8_dist_check.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 \
  bench.py --gpus 8 --steps 2 --warmup 3 > $OUT/n8_bench.json 2> $OUT/n8_bench.err
echo "bench N=8 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/n8_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms','e2e')}); print(d['minsr_phases_ms']); print(d['clocks']); print(d['config_B']['value'], d['config_B']['minsr_step_ms'], d['config_B']['minsr_phases_ms'])
PY
tail -n 5 $OUT/n8_bench.err
