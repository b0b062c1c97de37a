mkdir -p gpurun_out
( timeout 540 python -m pytest tests/test_gpu_rt.py -m gpu -x -q --tb=short; echo "pytest exit $?" ) > gpurun_out/pytest_gpu_rt.log 2>&1
tail -3 gpurun_out/pytest_gpu_rt.log
timeout 120 python tools/ab_quick.py base f64 8 >> gpurun_out/ab.log 2>&1
timeout 120 python tools/ab_quick.py v2c32 mixed 8 >> gpurun_out/ab.log 2>&1
tail -1 gpurun_out/ab.log
RB_LIB_PATH=radiobear_b200/lib/librb_c64.so timeout 120 python tools/ab_quick.py v2c64 mixed 8 >> gpurun_out/ab.log 2>&1; tail -2 gpurun_out/ab.log
