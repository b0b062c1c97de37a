# ray trace: chain shortened by one / two latencies per segment (RB_GEO_FUSED=2 / 1) x 8 / 7 CTAs per SM
rm -f /tmp/ab_ref_tb.npy
for v in cur f2c8 f1c8 p7 f2c7 f1c7; do
  lib=radiobear_b200/lib/librb_$v.so; [ $v = cur ] && lib=radiobear_b200/lib/libradiobear_b200.so
  RB_LIB_PATH=$lib timeout 120 python tools/ab_quick.py $v f64 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$v', 'N1 step', round(d['step_ms'], 4), 'rt', round(d['rt_ms'], 4), 'geo', round(d['geometry_ms'], 4), 'dTb', d.get('max_abs_dTb_K'), d.get('nan_pattern_equal'))"
  RB_LIB_PATH=$lib RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 RB_BENCH_EMULATE_WORLD=8 timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   world 8 share: ms', round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernels_ms'].items() if k != 'note'})"
done
