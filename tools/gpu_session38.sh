#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python tools/tc_bwd_parity_probe.py 8 8 88 1.6,1.8,2.0,2.2,2.4 > $OUT/s38_parity_E.log 2>&1; head -n 5 $OUT/s38_parity_E.log | cut -c1-130
timeout 900 python tools/tc_bwd_parity_probe.py 8 3 40 1.6,1.8,2.0,2.2,2.4 > $OUT/s38_parity_c40.log 2>&1; head -n 5 $OUT/s38_parity_c40.log | cut -c1-130
timeout 900 python tools/tc_bwd_parity_probe.py 8 4 24 1.6,1.8,2.0,2.2,2.4 > $OUT/s38_parity_c24.log 2>&1; head -n 5 $OUT/s38_parity_c24.log | cut -c1-130
