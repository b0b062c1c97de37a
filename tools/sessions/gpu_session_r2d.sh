# Round 2, session D: one geometry kernel (compaction bit-identical?), 4-warp CTAs as default, variants, e2e host profile
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2d_pytest.log 2>&1
tail -4 gpurun_out/r2d_pytest.log
timeout 300 python tools/diag_compact.py 2>&1 | grep -v "^  ray" | tail -8
rm -f gpurun_out/ab_quick.jsonl /tmp/ab_ref_tb.npy
for v in w4 w8 w4c5 w4s10 w4s12 w2; do
  RB_LIB_PATH=radiobear_b200/lib/librb_$v.so timeout 120 python tools/ab_quick.py $v f64 8 2>&1 | tail -1
done
timeout 300 python tools/e2e_profile.py > gpurun_out/r2d_e2e_profile.log 2>&1; head -60 gpurun_out/r2d_e2e_profile.log
