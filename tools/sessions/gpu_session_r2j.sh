mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "bench exit $?"
grep -c "NCCL INFO" gpurun_out/r2j_bench_n2.err; grep -m3 "nranks\|NVLS\|Connected all" gpurun_out/r2j_bench_n2.err | cut -c1-200; tail -3 gpurun_out/r2j_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench_n2.json'))
print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['kernels_ms'])
for k in ('alpha_c5','alpha_c5_full'):
    a=d[k]; print(k, a['n_gpus'], a['ms'], a['kernel_ms'], a['value'], a['max_rel_err_vs_oracle'], a['fp64_frac'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py 2>&1 | tail -5
