# streamed geometry: parity first, then A/B at emulated world sizes
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
for w in 8 4 2; do
  for s in 0 1; do
    RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 RB_BENCH_EMULATE_WORLD=$w RB_RT_STREAM_GEOMETRY=$s timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('world', $w, 'stream', $s, 'ms', d['ms_per_step'], 'e2e', d['e2e'].get('ms_per_step'), {k: d.get('kernels_ms', {}).get(k) for k in d.get('kernels_ms', {})})"
  done
done
for s in 0 1; do
  RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 RB_RT_STREAM_GEOMETRY=$s timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2x_n1_s$s.json
  python -c "
import json
d = json.load(open('gpurun_out/r2x_n1_s$s.json'))
print('N1 stream', $s, d['ms_per_step'], d['e2e'])"
done
