# One GPU-box session, most important evidence first (every step writes into gpurun_out/ as it finishes):
#   gpurun --timeout 900 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 540 python -m pytest tests -m gpu -x -q --tb=short --durations=8; echo "pytest exit $?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?" ) > gpurun_out/smoke.log 2>&1
tail -2 gpurun_out/smoke.log
timeout 120 python tools/ab_quick.py base f64 8 >> gpurun_out/ab.log 2>&1
timeout 120 python tools/ab_quick.py base mixed 8 >> gpurun_out/ab.log 2>&1
tail -2 gpurun_out/ab.log
RB_RT_PRECISION=mixed timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
tail -c 600 gpurun_out/bench_mixed.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rt_integrate_rays_mixed -c 1 -f \
  -o gpurun_out/prof_mixed python tools/ab_quick.py ncu_full mixed 1 > gpurun_out/ncu_full.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_mixed.csv \
  python tools/ab_quick.py ncu_list mixed 2 > gpurun_out/ncu_list.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
RB_LIB_PATH=radiobear_b200/lib/librb_m5.so timeout 120 python tools/ab_quick.py m5 mixed 8 >> gpurun_out/ab.log 2>&1
RB_LIB_PATH=radiobear_b200/lib/librb_t6d3.so timeout 120 python tools/ab_quick.py t6d3 f64 8 >> gpurun_out/ab.log 2>&1
RB_LIB_PATH=radiobear_b200/lib/librb_t8d3.so timeout 120 python tools/ab_quick.py t8d3 f64 8 >> gpurun_out/ab.log 2>&1
cat gpurun_out/ab_quick.jsonl
( RB_RT_PRECISION=mixed timeout 400 python -m pytest tests -m gpu -x -q --tb=short; echo "pytest exit $?" ) > gpurun_out/pytest_gpu_mixed_default.log 2>&1; tail -2 gpurun_out/pytest_gpu_mixed_default.log
