#!/bin/bash
# end-of-round evidence with the final build: launch list of one whole default bench step (thermalisation + sweep +
# Oloc + Jacobian + Gram + solve) and one full ncu capture of the forward tower
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 16000 --csv --log-file $OUT/s43_launches_E.csv \
  python bench.py --workload E --steps 1 --warmup 0 --no-cpu --no-peaks > $OUT/s43_ncu_launch.log 2>&1
echo "ncu launches rc=$?"; tail -n 2 $OUT/s43_ncu_launch.log | cut -c1-300
python - <<'PY'
import csv, collections
rows = []
with open('gpurun_out/s43_launches_E.csv') as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
agg = collections.OrderedDict()
n = 0
for row in r:
    if len(row) <= iv: continue
    k = row[ik][:80]; t = float(row[iv].replace(',', '')) * scale.get(row[iu], 1e-6)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t; n += 1
tot = sum(a[1] for a in agg.values())
with open('gpurun_out/s43_launches_E_summary.csv', 'w') as f:
    f.write('kernel,launches,total_ms,share_pct\n')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.3f,%.2f\n' % (k, a[0], a[1], 100 * a[1] / tot))
print(n, 'launches', tot, 'ms'); print(open('gpurun_out/s43_launches_E_summary.csv').read()[:3000])
PY
gzip -f $OUT/s43_launches_E.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resconv_tc2 -s 2 -c 2 -f -o $OUT/s43_prof_resconv_tc \
  python tools/resconv_probe.py E > $OUT/s43_ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la $OUT | grep s43
