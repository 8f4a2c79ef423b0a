#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/s25_launches_gram_n8shape.csv \
  python tools/gram_order_probe.py "N=8" > $OUT/s25_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/s25_launches_gram_n8shape.csv', errors='ignore')))
hdr=None
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r
        continue
    if len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')=='gpu__time_duration.sum': print(d['ID'], d['Kernel Name'][:50], d['Metric Value'], d['Metric Unit'])
PY
