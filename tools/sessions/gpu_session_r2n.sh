mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; tail -2 gpurun_out/r2w_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err; cut -c1-400 gpurun_out/r2w_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2w_launches_bench.log 2>&1
RB_BENCH_SKIP_MIXED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rt_integrate_pairs|ray_geometry|ray_compact|ray_edge" -s 12 -c 4 -f -o gpurun_out/prof_r2w_rt python tools/ab_quick.py ncu f64 1 > gpurun_out/r2w_ncu.log 2>&1
ls -la gpurun_out/prof_r2w_rt.ncu-rep
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2w_bench_n1.json'))
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['kernels_ms'])
print('retrieval', {k:v for k,v in d['retrieval_loop'].items() if 'ms' in k})
r=d['roofline']; print({k:r[k] for k in ('bound','achieved','peak','frac','pipe_frac')}, r['issue']['frac'], r['smem']['frac'])
print('alpha_c5', d['alpha_c5']['ms'], d['alpha_c5']['fp64_frac'], 'gpu_launches', d['gpu_launches'])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"alpha_lines" -s 3 -c 1 -f -o gpurun_out/prof_r2w_alpha python tools/alpha_c5_probe.py > gpurun_out/r2w_ncu_alpha.log 2>&1
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
