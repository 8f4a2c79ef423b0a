#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python tools/tc_bwd_parity_probe.py 8 > $OUT/s17_parity_probe.log 2>&1
echo "parity rc=$?"; tail -n 10 $OUT/s17_parity_probe.log
