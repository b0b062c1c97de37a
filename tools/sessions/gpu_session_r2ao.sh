# warps per CTA of the pair kernel under the new launch order; then the new GPU tests
rm -f /tmp/ab_ref_tb.npy
timeout 120 python tools/ab_quick.py cur f64 10 2>&1 | tail -1 | cut -c1-330
for v in w2 w8 w4c7 ; do
  RB_LIB_PATH=radiobear_b200/lib/librb_$v.so timeout 120 python tools/ab_quick.py $v f64 10 2>&1 | tail -1 | cut -c1-330
done
for w in 8 ; do
  for v in cur w2; do
    lib=radiobear_b200/lib/librb_$v.so; [ $v = cur ] && lib=radiobear_b200/lib/libradiobear_b200.so
    RB_LIB_PATH=$lib RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 RB_BENCH_CPU_PIXELS=8 RB_BENCH_EMULATE_WORLD=$w timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('world', $w, '$v', 'ms', round(d['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernels_ms'].items() if k != 'note'})"
  done
done
( timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -k "fill or launch_order"; echo "pytest exit $?" ) 2>&1 | tail -4
