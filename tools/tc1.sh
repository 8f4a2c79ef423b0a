set -x
for c in "16 16 2 64" "10 32 3 37" "12 24 2 33" "4 8 2 16" "16 40 2 700 sinhp1" "16 128 2 64"; do
  TC_F64REF=1 TC_MODES=all timeout 120 python tools/tc_debug.py $c 2>&1 | tail -4
done
TC_F64REF=1 TC_MODES=all timeout 200 python tools/tc_debug.py 16 88 8 2048 2>&1 | tail -4
TC_ONLY=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:resconv_tc_kernel -c 1 -f -o gpurun_out/prof_tc_fwd2 python tools/tc_debug.py 16 88 8 592 2>&1 | tail -3
