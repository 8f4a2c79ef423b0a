#!/bin/bash
set -u
OUT=gpurun_out
# one Jacobian call at config E (ns=2048): tower launch 0 = forward (save), 1 = backward-data, then the wgrad kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resconv_tc2|wgrad_tc" -s 1 -c 2 -o $OUT/s21_prof_bwd \
  python tools/tc_bwd_probe.py time > $OUT/s21_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 4 $OUT/s21_ncu.log
ls -la $OUT/s21_prof_bwd.ncu-rep
