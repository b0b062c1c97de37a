# device-sized parts of the integration launch (default) against one part; ncu --set full of the integration; ray-trace tests
rm -f /tmp/ab_ref_tb.npy
for n in 1 default 1 default; do
  if [ $n = default ]; then unset RB_RT_PARTS; else export RB_RT_PARTS=$n; fi
  timeout 120 python tools/ab_quick.py parts_$n f64 10 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('parts', '$n', 'step', round(d['step_ms'], 4), 'rt', round(d['rt_ms'], 4), 'geo', round(d['geometry_ms'], 4), 'dTb', d.get('max_abs_dTb_K'))"
done
unset RB_RT_PARTS
RB_BENCH_SKIP_MIXED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rt_integrate_pairs" -s 3 -c 1 -f -o gpurun_out/prof_r2h_rt python tools/ab_quick.py ncu f64 1 > gpurun_out/r2h_ncu.log 2>&1
ls -la gpurun_out/prof_r2h_rt.ncu-rep
( timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rt.py tests/test_gpu_planet.py -m gpu -q --tb=short --maxfail=5; echo "pytest exit $?" ) 2>&1 | tail -4
