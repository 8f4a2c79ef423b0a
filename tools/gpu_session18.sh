#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python tools/tc_bwd_parity_probe.py 8 8 88 1.0,1.25,1.5,1.75,2.0 > $OUT/s18_parity_probe_E.log 2>&1
echo "parity rc=$?"; head -n 5 $OUT/s18_parity_probe_E.log | cut -c1-120
timeout 900 python tools/tc_bwd_parity_probe.py 8 4 32 0,1.0,1.5,1.75,2.0,2.2 > $OUT/s18_parity_probe_c32.log 2>&1
echo "parity rc=$?"; head -n 6 $OUT/s18_parity_probe_c32.log | cut -c1-120
timeout 900 python tools/tc_bwd_parity_probe.py 8 3 128 0,1.0,1.5,1.75,2.0,2.2 > $OUT/s18_parity_probe_c128.log 2>&1
echo "parity rc=$?"; head -n 6 $OUT/s18_parity_probe_c128.log | cut -c1-120
