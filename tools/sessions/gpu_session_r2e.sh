# Round 2, session E: resident absorption path, bench line with the new keys, ncu evidence of the 4-warp pair kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2e_pytest.log 2>&1
tail -15 gpurun_out/r2e_pytest.log
timeout 300 python tools/e2e_profile.py > gpurun_out/r2e_e2e_profile.log 2>&1; head -14 gpurun_out/r2e_e2e_profile.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; tail -3 gpurun_out/r2e_bench_n1.err; cut -c1-1500 gpurun_out/r2e_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2e_launches_bench.log 2>&1
RB_BENCH_SKIP_MIXED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rt_integrate_pairs|ray_geometry" -s 6 -c 2 -f -o gpurun_out/prof_r2e_pairs python tools/ab_quick.py ncu f64 1 > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
