#!/usr/bin/env bash
# One gpurun call that runs everything queued in DESIGN.md section 9 (code written without GPU access):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/first_gpu_session.sh'            (1 GPU)
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 900 -- 'bash tools/first_gpu_session.sh 8' (N GPUs: adds the probes
#                                                                                              that need several ranks)
# Logs land in gpurun_out/ (merged back by gpurun).  Every step has its own time-out; none of them is a bench value.
set -u
N="${1:-1}"
OUT=gpurun_out
mkdir -p "$OUT"
export QTX_UNVERIFIED=1
timeout 600 python -m pytest tests/test_zz_solver_variants_gpu.py -m gpu -q --tb=short > "$OUT/unverified_tests.log" 2>&1
echo "unverified tests rc=$?" | tee -a "$OUT/unverified_tests.log"
timeout 300 python tools/unverified_probe.py > "$OUT/unverified_probe.json" 2> "$OUT/unverified_probe.err"
echo "probe rc=$?"
# the rational route as the solver of the whole bench step (compare minsr_step_ms with the default line)
QTX_PINV=rational timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-probe > "$OUT/bench_rational_n1.json" 2> "$OUT/bench_rational_n1.err"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-probe > "$OUT/bench_eigh_n1.json" 2> "$OUT/bench_eigh_n1.err"
if [ "$N" -gt 1 ]; then
  export QTX_P2P_TIMEOUT_S=20
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29611 \
      tools/multi_gpu_probe.py > "$OUT/multi_gpu_probe_n$N.json" 2> "$OUT/multi_gpu_probe_n$N.err"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29612 \
      tests/run_p2p_gram_check.py > "$OUT/p2p_gram_check_n$N.log" 2>&1
  # the distributed VMC steps (RBM, ResConv, complex ResConv) must equal the single-GPU step with the shifts split over ranks
  QTX_PINV=rational timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29614 tests/run_dist_check.py > "$OUT/dist_check_rational_n$N.log" 2>&1
  QTX_PINV=rational timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29613 bench.py --gpus "$N" --steps 5 --warmup 3 --no-probe > "$OUT/bench_rational_n$N.json" 2> "$OUT/bench_rational_n$N.err"
fi
tail -n 3 "$OUT"/unverified_tests.log "$OUT"/unverified_probe.json
