#!/bin/bash
set -u
OUT=gpurun_out
timeout 300 python tools/tc_bwd_probe.py raster > $OUT/s30_probe_raster.log 2>&1
echo "raster rc=$?"; tail -n 9 $OUT/s30_probe_raster.log
timeout 300 python tools/tc_bwd_probe.py small > $OUT/s30_probe_small.log 2>&1
echo "small rc=$?"; tail -n 7 $OUT/s30_probe_small.log
