# final code: full GPU suite, smoke, the bench line on one B200
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=8; echo "pytest exit $?" ) 2>&1 | tail -5
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err; tail -2 gpurun_out/r2final_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2final_bench_n1.json'))
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['kernels_ms'])
r=d['roofline']; print({k:r[k] for k in ('bound','achieved','peak','frac','pipe_frac','traffic')})
print('alpha_c5', d['alpha_c5']['ms'], d['alpha_c5']['fp64_frac'], 'gpu_launches', d['gpu_launches'])
PY
