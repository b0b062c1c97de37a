# N=2: full GPU suite on one GPU, then the bench with the absorption sharded / replicated in the step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for mode in sharded replicated; do
  RB_BENCH_STEP_ALPHA=$mode RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y_n2_$mode.json 2> gpurun_out/r2y_n2_$mode.err; echo "bench exit $?"
  python - <<PY
import json
d=json.load(open('gpurun_out/r2y_n2_$mode.json'))
print('$mode', 'N', d['n_gpus'], 'step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if k != 'note'})
PY
done
RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2y_n1.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2y_n1.json'))
print('N1', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k: v for k, v in d['kernels_ms'].items() if k != 'note'})
PY
