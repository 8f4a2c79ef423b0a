#!/bin/bash
# 2 GPUs: distributed step == single-GPU step with the adaptive Lanczos of the library solve; bench line
set -u
OUT=gpurun_out
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/n2c_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 $OUT/n2c_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 2 --warmup 3 > $OUT/n2c_bench.json 2> $OUT/n2c_bench.err
echo "bench N=2 rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/n2c_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms')}); print({k:v for k,v in d['minsr_phases_ms'].items() if 'lanczos' in k or 'solve' in k}); print(d['config_B']['minsr_step_ms'], {k:v for k,v in d['config_B']['minsr_phases_ms'].items() if 'lanczos' in k or 'shifted' in k})
PY
tail -n 3 $OUT/n2c_bench.err
