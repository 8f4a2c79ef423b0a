set -x
TC_F64REF=1 TC_MODES=all timeout 200 python tools/tc_debug.py 16 88 8 512 2>&1 | tail -5
TC_ONLY=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:resconv_tc_kernel -c 1 -f -o gpurun_out/prof_tc_fwd python tools/tc_debug.py 16 88 8 592 2>&1 | tail -5
