#!/bin/bash
# Round-2 multi-GPU session (2 GPUs): distributed step == single-GPU step through the library's own collectives
# (csrc/comm.cu) and through torch.distributed, peer-memory Gram, and the bench line at N = 2.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/n2_smi.txt 2>&1
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --tb=short -p no:cacheprovider > $OUT/n2_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/n2_pytest.log
tail -n 25 $OUT/n2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 2 --warmup 1 > $OUT/n2_bench.json 2> $OUT/n2_bench.err
echo "bench N=2 rc=$?"; tail -c 3000 $OUT/n2_bench.json; tail -n 8 $OUT/n2_bench.err
QTX_DIST_C=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
  bench.py --gpus 2 --steps 3 --warmup 3 --workload B > $OUT/n2_bench_B_torchdist.json 2> $OUT/n2_bench_B_torchdist.err
echo "bench B torch.distributed rc=$?"; tail -c 1500 $OUT/n2_bench_B_torchdist.json
timeout 300 python tools/pinv_probe.py 2048 4096 > $OUT/n2_pinv_probe.jsonl 2> $OUT/n2_pinv_probe.err
echo "probe rc=$?"; cat $OUT/n2_pinv_probe.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zldlt" -c 70 --csv --log-file $OUT/n2_launches_ldlt4096.csv \
  python tools/pinv_probe.py 4096 > $OUT/n2_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
