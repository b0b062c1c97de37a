# One GPU-box session (gpurun --timeout 900 -- "bash tools/sessions/gpu_session.sh"): GPU test suite, smoke, both bench arms,
# ncu --set full of the mixed integration kernel and the ncu launch list of bench.py; everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 540 python -m pytest tests -m gpu -x -q --tb=short; echo "pytest exit $?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -2 gpurun_out/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke exit $?" ) > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
RB_RT_PRECISION=mixed timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rt_integrate_rays_mixed -c 1 -f \
  -o gpurun_out/prof_mixed_v2 python tools/ab_quick.py ncu_full mixed 1 > gpurun_out/ncu_full.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list.log 2>&1
python -c "
import json
for f in ('bench_f64','bench_mixed'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['issue']['frac'], d['roofline']['smem']['frac'], (d.get('rt_mixed') or {}).get('ms_per_step'))"
