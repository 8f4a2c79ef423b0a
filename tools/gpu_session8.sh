#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"zldlt|ztrsv" -c 40 --csv --log-file $OUT/s8_launches_small.csv \
  python tools/pinv_probe.py 64 128 256 > $OUT/s8_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/s8_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:zldlt_step -c 1 -o $OUT/s8_prof_diag_only \
  python tools/pinv_probe.py 64 > $OUT/s8_ncu2.log 2>&1
echo "ncu rc=$?"
