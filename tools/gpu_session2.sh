#!/bin/bash
# Round-2 GPU session 2: the own LDL^T pseudo-inverse (tests, timing probe), fixed parity tests, sanitizer passes.
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/parity_report.jsonl
timeout 900 python -m pytest tests/test_zz_solver_variants_gpu.py tests/test_baseline_shapes_gpu.py tests/test_solver_gpu.py \
  "tests/test_rbm_gpu.py::test_rbm_conv_matches_oracle" -m gpu -q --tb=short -p no:cacheprovider -x > $OUT/s2_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/s2_pytest.log
tail -n 30 $OUT/s2_pytest.log
timeout 600 python tools/pinv_probe.py 2048 4096 8192 16384 > $OUT/s2_pinv_probe.jsonl 2> $OUT/s2_pinv_probe.err
echo "probe rc=$?"; cat $OUT/s2_pinv_probe.jsonl; tail -n 5 $OUT/s2_pinv_probe.err
timeout 300 python bench.py --workload B --steps 5 --warmup 3 --no-cpu --no-peaks > $OUT/s2_bench_B.json 2> $OUT/s2_bench_B.err
echo "bench B rc=$?"; cat $OUT/s2_bench_B.json; tail -n 5 $OUT/s2_bench_B.err
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py ldlt rbm gram > $OUT/s2_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; tail -n 6 $OUT/s2_sanitizer_$tool.log
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_probe.py resconv > $OUT/s2_sanitizer_memcheck_resconv.log 2>&1
echo "sanitizer memcheck resconv rc=$?"; tail -n 6 $OUT/s2_sanitizer_memcheck_resconv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zldlt_update -s 30 -c 2 -o $OUT/s2_prof_zldlt_update \
  python tools/pinv_probe.py 4096 > $OUT/s2_ncu_zldlt.log 2>&1
echo "ncu rc=$?"
ls -la $OUT | tail -n 20
