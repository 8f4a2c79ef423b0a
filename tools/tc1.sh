TC_F64REF=1 TC_MODES=all timeout 100 python tools/tc_debug.py 16 88 8 2048 2>&1 | tail -3
QTX_TC_2CTA=0 TC_MODES=all timeout 100 python tools/tc_debug.py 16 88 8 2048 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -3
