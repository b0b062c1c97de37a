# Round 2, session B: geometry single-wave fix, A/B, ncu of the three hot kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2b_pytest.log 2>&1
tail -4 gpurun_out/r2b_pytest.log
rm -f gpurun_out/ab_quick.jsonl /tmp/ab_ref_tb.npy
RB_RT_PAIRS=0 RB_RT_COMPACT=0 timeout 120 python tools/ab_quick.py legacy f64 8 2>&1 | tail -1
RB_RT_PAIRS=0 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py compact f64 8 2>&1 | tail -1
RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py pairs_compact f64 8 2>&1 | tail -1
RB_LIB_PATH=radiobear_b200/lib/librb_ctas3.so RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py pairs_compact_ctas3 f64 8 2>&1 | tail -1
RB_LIB_PATH=radiobear_b200/lib/librb_ctas3.so RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:"rt_integrate_pairs|ray_geometry" -s 6 -c 2 -f -o gpurun_out/prof_r2b_pairs python tools/ab_quick.py ncu f64 1 > gpurun_out/r2b_ncu1.log 2>&1
RB_RT_PAIRS=0 RB_RT_COMPACT=1 timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:"rt_integrate_rays_kernel" -s 3 -c 1 -f -o gpurun_out/prof_r2b_rays python tools/ab_quick.py ncu f64 1 > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
