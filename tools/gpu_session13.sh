#!/bin/bash
# Round-2 GPU session 13: first run of the tensor-core Jacobian (probe against the CUDA-core backward pass)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/tc_bwd_probe.py small > $OUT/s13_probe_small.log 2>&1
echo "small rc=$?"; tail -n 20 $OUT/s13_probe_small.log
timeout 300 python tools/tc_bwd_probe.py E > $OUT/s13_probe_E.log 2>&1
echo "E rc=$?"; tail -n 8 $OUT/s13_probe_E.log
timeout 300 python tools/tc_bwd_probe.py time > $OUT/s13_probe_time.log 2>&1
echo "time rc=$?"; tail -n 8 $OUT/s13_probe_time.log
