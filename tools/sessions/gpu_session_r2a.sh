# Round 2, session A: correctness of the new geometry / compaction / pair kernels + first A/B timings
mkdir -p gpurun_out
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?" ) > gpurun_out/r2a_smoke.log 2>&1
tail -3 gpurun_out/r2a_smoke.log
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
rm -f gpurun_out/ab_quick.jsonl /tmp/ab_ref_tb.npy
RB_RT_PAIRS=0 RB_RT_COMPACT=0 timeout 120 python tools/ab_quick.py legacy f64 8 2>&1 | tail -1
RB_RT_PAIRS=0 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py compact f64 8 2>&1 | tail -1
RB_RT_PAIRS=1 RB_RT_COMPACT=0 timeout 120 python tools/ab_quick.py pairs f64 8 2>&1 | tail -1
RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py pairs_compact f64 8 2>&1 | tail -1
RB_LIB_PATH=radiobear_b200/lib/librb_ctas3.so RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py pairs_compact_ctas3 f64 8 2>&1 | tail -1
RB_LIB_PATH=radiobear_b200/lib/librb_ctas3.so RB_RT_PAIRS=0 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py compact_ctas3 f64 8 2>&1 | tail -1
