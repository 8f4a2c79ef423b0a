for c in "16 88 8 2048" "10 36 3 337" "12 20 2 33"; do
  TC_F64REF=1 TC_MODES=all timeout 200 python tools/tc_debug.py $c 2>&1 | tail -3
done
QTX_TC_DEBUG=1 TC_ONLY=1 timeout 200 python tools/tc_debug.py 16 88 8 592 2>&1 | tail -1
python -m pytest tests/test_resconv_gpu.py -x -q 2>&1 | tail -3
