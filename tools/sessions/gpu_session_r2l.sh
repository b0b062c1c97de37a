mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2l_pytest.log 2>&1
tail -12 gpurun_out/r2l_pytest.log
rm -f /tmp/ab_ref_tb.npy
RB_RT_COMPACT=0 timeout 120 python tools/ab_quick.py plain f64 8 2>&1 | tail -1 | cut -c1-330
timeout 120 python tools/ab_quick.py sorted_compaction f64 8 2>&1 | tail -1 | cut -c1-420
timeout 200 python tools/e2e_ab.py 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
