#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python tools/tc_bwd_parity_probe.py 8 > $OUT/s16_parity_probe.log 2>&1
echo "parity rc=$?"; tail -n 8 $OUT/s16_parity_probe.log
timeout 300 python tools/tc_bwd_probe.py time > $OUT/s16_probe_time.log 2>&1
echo "time rc=$?"; tail -n 4 $OUT/s16_probe_time.log
