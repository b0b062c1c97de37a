mkdir -p gpurun_out
( timeout 540 python -m pytest tests/test_gpu_rt.py -m gpu -x -q --tb=short; echo "pytest exit $?" ) > gpurun_out/pytest_gpu_rt.log 2>&1
tail -3 gpurun_out/pytest_gpu_rt.log
timeout 120 python tools/ab_quick.py base f64 8 >> gpurun_out/ab.log 2>&1
timeout 120 python tools/ab_quick.py cut25 mixed 8 >> gpurun_out/ab.log 2>&1
tail -1 gpurun_out/ab.log
