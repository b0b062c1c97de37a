mkdir -p gpurun_out
RB_ALPHA_RESIDENT=0 timeout 200 python tools/e2e_ab.py 2>&1 | tail -3
RB_ALPHA_RESIDENT=1 timeout 200 python tools/e2e_ab.py 2>&1 | tail -3
RB_ALPHA_RESIDENT=0 timeout 200 python tools/e2e_ab.py 2>&1 | tail -3
RB_ALPHA_RESIDENT=1 timeout 200 python tools/e2e_ab.py 2>&1 | tail -3
RB_TRACE=1 RB_ALPHA_RESIDENT=1 timeout 200 python tools/e2e_ab.py 2>&1 | grep rb_trace | tail -6
RB_TRACE=1 RB_ALPHA_RESIDENT=0 timeout 200 python tools/e2e_ab.py 2>&1 | grep rb_trace | tail -6
( timeout 600 python -m pytest tests/test_gpu_planet.py tests/test_gpu_alpha.py -m gpu -q --tb=short -x; echo "pytest exit $?" ) 2>&1 | tail -5
