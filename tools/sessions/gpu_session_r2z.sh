# N=8: the full bench line (sharded absorption in the step), then the same step with the absorption replicated
bash tools/sessions/gpu_session_scale.sh 8
RB_BENCH_STEP_ALPHA=replicated RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2z_n8_replicated.json 2> gpurun_out/r2z_n8_replicated.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2z_n8_replicated.json'))
print('replicated N', d['n_gpus'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if k != 'note'})
PY
