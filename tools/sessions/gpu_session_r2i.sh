mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) > gpurun_out/r2i_pytest.log 2>&1
tail -25 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; tail -3 gpurun_out/r2i_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench_n1.json'))
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['kernels_ms'])
print('retrieval', d['retrieval_loop'])
r=d['roofline']; print({k:r[k] for k in ('bound','achieved','peak','frac','pipe_frac')})
print('alpha_c5', d['alpha_c5']['ms'], d['alpha_c5']['fp64_frac'])
PY
