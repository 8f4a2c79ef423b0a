#!/bin/bash
# Round-2 multi-GPU session (8 GPUs), bench line only, after the Gram changes (equal K chunks, coalesced epilogue)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 \
  bench.py --gpus 8 --steps 2 --warmup 3 > $OUT/n8b_bench.json 2> $OUT/n8b_bench.err
echo "bench N=8 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/n8b_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms','e2e')}); print(d['minsr_phases_ms']); print(d['clocks']); print(d['config_B']['value'], d['config_B']['minsr_step_ms'], d['config_B']['minsr_phases_ms'])
PY
tail -n 3 $OUT/n8b_bench.err
