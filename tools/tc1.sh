set -x
for c in "16 16 2 64" "16 88 2 300" "10 32 3 37" "12 24 2 33" "4 8 2 16" "16 88 8 2048"; do
  timeout 120 python tools/tc_debug.py $c 2>&1 | tail -8
done
