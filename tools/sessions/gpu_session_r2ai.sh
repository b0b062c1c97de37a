# final bench line of round 2 on one B200 (both arms) + launch list + host stage times of Planet.run
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err; tail -2 gpurun_out/r2final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2final_bench_ref.json 2> gpurun_out/r2final_bench_ref.err; cut -c1-200 gpurun_out/r2final_bench_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2final_bench_n1.json'))
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['kernels_ms'])
print('retrieval', {k:v for k,v in d['retrieval_loop'].items() if 'ms' in k})
r=d['roofline']; print({k:r[k] for k in ('bound','achieved','peak','frac','pipe_frac','traffic')}, r['issue']['frac'], r['smem']['frac'])
print('alpha_c5', d['alpha_c5']['ms'], d['alpha_c5']['fp64_frac'], 'gpu_launches', d['gpu_launches'])
PY
RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2final_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2final_launches_bench.log 2>&1
RB_TRACE=1 timeout 200 python tools/e2e_ab.py 2>&1 | grep -v "^\[rb_trace\]" | tail -4
RB_TRACE=1 timeout 200 python tools/e2e_ab.py 2>&1 | grep "rb_trace" | tail -8
