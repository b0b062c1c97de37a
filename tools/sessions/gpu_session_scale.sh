# bench.py at N ranks (torchrun), NCCL_DEBUG=INFO on stderr:  bash tools/sessions/gpu_session_scale.sh N
N=$1
mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench exit $? lines $(wc -l < gpurun_out/r2_bench_n$N.json)"
grep -m1 "nranks" gpurun_out/r2_bench_n$N.err | cut -c1-160
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n$N.json'))
print('N', d['n_gpus'], 'step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['kernels_ms'])
for k in ('alpha_c5','alpha_c5_full'):
    a=d[k]; print(k, a['n_gpus'], 'ms', a['ms'], 'kernel', a['kernel_ms'], 'value', a['value'], a['max_rel_err_vs_oracle'], a['fp64_frac'])
PY
