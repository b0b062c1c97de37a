# Round 2, session C: state of the GPU tests, compaction diagnosis, A/B of table / CTA-shape variants, convoy experiment
mkdir -p gpurun_out
cp radiobear_b200/lib/librb_base.so radiobear_b200/lib/libradiobear_b200.so
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2c_pytest.log 2>&1
tail -4 gpurun_out/r2c_pytest.log
timeout 300 python tools/diag_compact.py > gpurun_out/r2c_diag_compact.log 2>&1; tail -40 gpurun_out/r2c_diag_compact.log
rm -f gpurun_out/ab_quick.jsonl /tmp/ab_ref_tb.npy
for v in base rep16n128 rep16n256 w4 w4rep; do
  RB_LIB_PATH=radiobear_b200/lib/librb_$v.so RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 120 python tools/ab_quick.py $v f64 8 2>&1 | tail -1
done
M=gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max
for v in base rep16n128 w4; do
  for sf in "" 10 40; do
    RB_AB_SAMEFREQ=$sf RB_LIB_PATH=radiobear_b200/lib/librb_$v.so RB_RT_PAIRS=1 RB_RT_COMPACT=1 timeout 300 ncu --metrics $M --clock-control none \
      -k regex:"rt_integrate_pairs" -s 4 -c 1 --csv --log-file gpurun_out/r2c_ncu_${v}_sf${sf}.csv python tools/ab_quick.py ncu_${v}_sf${sf} f64 1 > gpurun_out/r2c_ncu_${v}_sf${sf}.log 2>&1
    tail -1 gpurun_out/r2c_ncu_${v}_sf${sf}.log | cut -c1-400
  done
done
