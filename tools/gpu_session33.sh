#!/bin/bash
set -u
OUT=gpurun_out
QTX_TC_DEBUG=1 timeout 300 python tools/resconv_probe.py E > $OUT/s33_fwd_dbg.log 2>&1
echo "rc=$?"; grep "tc dbg" $OUT/s33_fwd_dbg.log | head -3; grep "forward" $OUT/s33_fwd_dbg.log | head -2
