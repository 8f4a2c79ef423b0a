#!/bin/bash
set -u
OUT=gpurun_out
QTX_TC_DEBUG=1 timeout 300 python tools/resconv_probe.py E > $OUT/s34_fwd_dbg.log 2>&1
echo "rc=$?"; grep "tc dbg" $OUT/s34_fwd_dbg.log | head -2 | cut -c1-420; grep "forward\|jacobian" $OUT/s34_fwd_dbg.log | head -2
timeout 300 python tools/resconv_probe.py E > $OUT/s34_fwd.log 2>&1; grep "forward\|jacobian" $OUT/s34_fwd.log | head -2
timeout 300 python tools/tc_bwd_probe.py small > $OUT/s34_probe_small.log 2>&1
echo "small rc=$?"; tail -n 3 $OUT/s34_probe_small.log
timeout 300 python tools/tc_bwd_probe.py time > $OUT/s34_probe_time.log 2>&1; tail -n 3 $OUT/s34_probe_time.log
