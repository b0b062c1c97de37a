mkdir -p gpurun_out
( timeout 540 python -m pytest tests -m gpu -x -q --tb=short; echo "pytest exit $?" ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
python -c "
import json; d=json.load(open('gpurun_out/bench_f64.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('rt_mixed'), d['roofline'].get('issue'), d['roofline'].get('smem'), d['cpu_baseline'])"
timeout 120 python tools/ab_quick.py base f64 8 >> gpurun_out/ab.log 2>&1
RB_LIB_PATH=radiobear_b200/lib/librb_s6.so timeout 120 python tools/ab_quick.py s6 mixed 8 >> gpurun_out/ab.log 2>&1
RB_LIB_PATH=radiobear_b200/lib/librb_s10.so timeout 120 python tools/ab_quick.py s10 mixed 8 >> gpurun_out/ab.log 2>&1
tail -3 gpurun_out/ab.log
