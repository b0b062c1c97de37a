for w in 2 1; do for fm in 0 1; do
RB_RT_FG_MAJOR=$fm RB_BENCH_EMULATE_WORLD=$w RB_BENCH_SKIP_MIXED=1 RB_BENCH_SKIP_C5_FULL=1 timeout 300 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); print('world', $w, 'fg_major', $fm, 'step', round(d['ms_per_step'],4), d['kernels_ms']['rt_integrate'], d['kernels_ms']['ray_geometry'], 'e2e', round(d['e2e']['ms_per_step'],3))"
done; done
