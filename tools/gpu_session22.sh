#!/bin/bash
set -u
OUT=gpurun_out
timeout 300 python tools/tc_bwd_probe.py small > $OUT/s22_probe_small.log 2>&1
echo "small rc=$?"; tail -n 8 $OUT/s22_probe_small.log
timeout 300 python tools/tc_bwd_probe.py time > $OUT/s22_probe_time.log 2>&1
echo "time rc=$?"; tail -n 4 $OUT/s22_probe_time.log
timeout 600 python tools/tc_bwd_parity_probe.py 8 8 88 1.8 > $OUT/s22_parity_probe_E.log 2>&1
echo "parity rc=$?"; head -n 3 $OUT/s22_parity_probe_E.log | cut -c1-140
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc2|wgrad_tc" -c 6 --csv --log-file $OUT/s22_launches_jac_E.csv \
  python tools/tc_bwd_probe.py time > $OUT/s22_ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/s22_launches_jac_E.csv', errors='ignore')))
hdr=None
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r
        continue
    if len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')=='gpu__time_duration.sum': print(d['ID'], d['Kernel Name'][:36], d['Metric Value'], d['Metric Unit'])
PY
