QTX_TC_DEBUG=1 TC_ONLY=1 timeout 200 python tools/tc_debug.py 16 88 8 592 2>&1 | tail -1
TC_F64REF=1 TC_MODES=all timeout 200 python tools/tc_debug.py 16 88 8 2048 2>&1 | tail -4
TC_F64REF=1 TC_MODES=all timeout 200 python tools/tc_debug.py 10 32 3 37 2>&1 | tail -4
