#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zldlt_step -c 1 -o $OUT/s7_prof_zldlt_head \
  python tools/pinv_probe.py 4096 > $OUT/s7_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/s7_ncu.log
