#!/bin/bash
# final single-GPU evidence of the round: whole GPU suite with the parity report, default bench, reference arm,
# ncu of the Jacobian kernels (launch list + full set), NVTX smoke
set -u
OUT=gpurun_out
rm -f $OUT/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > $OUT/s29_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s29_pytest.log
tail -n 5 $OUT/s29_pytest.log
cp $OUT/parity_report.jsonl $OUT/s29_parity_report.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/s29_bench.json 2> $OUT/s29_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/s29_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sweep_oloc_ms','minsr_step_ms','e2e')}); print(d['minsr_phases_ms']); print(d['jacobian_roofline']); print(d['roofline']['frac'], d['clocks'])
PY
QTX_REF_BUDGET_S=30 timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/s29_ref.json 2> $OUT/s29_ref.err
echo "ref rc=$?"; cut -c1-300 $OUT/s29_ref.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"resconv_tc2|wgrad_tc" -s 1 -c 2 -o $OUT/s29_prof_bwd \
  python tools/tc_bwd_probe.py time > $OUT/s29_ncu.log 2>&1
echo "ncu rc=$?"
QTX_NVTX=1 timeout 300 python tools/sanitize_probe.py resconv > $OUT/s29_nvtx_smoke.log 2>&1
echo "nvtx smoke rc=$?"; tail -n 2 $OUT/s29_nvtx_smoke.log
