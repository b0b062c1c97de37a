mkdir -p gpurun_out
rm -f gpurun_out/ab_quick.jsonl /tmp/ab_ref_tb.npy
for v in ring0 ring3x16 ring4x16c5 ring4x8 ring3x32c4 ring3x32c5; do
  RB_LIB_PATH=radiobear_b200/lib/librb_$v.so timeout 120 python tools/ab_quick.py $v f64 8 2>&1 | tail -1 | cut -c1-420
done
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12 -x; echo "pytest exit $?" ) > gpurun_out/r2h_pytest.log 2>&1
tail -5 gpurun_out/r2h_pytest.log
