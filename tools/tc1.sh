for i in 1 2 3; do python tools/qs_debug.py 2>&1 | tail -2; done
QTX_GRAM_NSLICES=-1 python tools/qs_debug.py 2>&1 | tail -2
