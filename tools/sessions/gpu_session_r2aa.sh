# streamed trace: frequency groups last-to-first vs first-to-last, emulated rank shares
for w in 8 4 2; do
  for o in 0 1; do
    RB_RT_STREAM_ORDER=$o RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 RB_BENCH_EMULATE_WORLD=$w timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('world', $w, 'reverse', $o, 'ms', round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernels_ms'].items() if k != 'note'})"
  done
done
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_rt.py -x -q -m gpu 2>&1 | tail -2
