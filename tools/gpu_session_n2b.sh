#!/bin/bash
# Round-2 multi-GPU session (2 GPUs), after the tensor-core Jacobian: distributed step == single-GPU step, bench lines at N = 2
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/n2b_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/n2b_pytest.log
tail -n 12 $OUT/n2b_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/n2b_bench.json 2> $OUT/n2b_bench.err
echo "bench N=2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/n2b_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms')}); print(d['minsr_phases_ms']); print(d['config_B']['minsr_step_ms'], d['config_B']['minsr_phases_ms'])
PY
tail -n 5 $OUT/n2b_bench.err
