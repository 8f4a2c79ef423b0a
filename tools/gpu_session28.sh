#!/bin/bash
set -u
OUT=gpurun_out
timeout 300 python tools/oloc_probe.py > $OUT/s28_oloc_probe.log 2>&1
echo "probe rc=$?"; tail -n 4 $OUT/s28_oloc_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/s28_launches_oloc_E.csv \
  python tools/oloc_probe.py > $OUT/s28_ncu.log 2>&1
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/s28_launches_oloc_E.csv', errors='ignore')))
hdr=None; seq=[]
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr=r
        continue
    if len(r)<len(hdr): continue
    d=dict(zip(hdr,r))
    if d.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']
    v = v/1e6 if u=='ns' else (v/1e3 if u=='us' else v)
    seq.append((re.sub(r'\(.*','',d['Kernel Name'])[:60], v))
# second Oloc call: between 2nd and 3rd conn kernels... print aggregate of the whole capture
agg=collections.defaultdict(lambda:[0,0.0])
for n,v in seq: agg[n][0]+=1; agg[n][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]: print(f"{t:9.2f} ms {n:5d} {t/n*1000:9.1f} us  {k}")
PY
