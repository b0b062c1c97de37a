#!/bin/bash
# Last GPU session of round 2 (one gpurun call per line; 25-50 s of box time each).
# 1. GPU suite on the final code
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2zz_pytest.log
# 2. bench line of the final code (-> profiles/r2_final2_bench_n1.json)
python bench.py > gpurun_out/r2zz_bench_n1.json 2> gpurun_out/r2zz_bench_n1.err
# 3. memory checker: smoke(), then every -m gpu test (-> profiles/r2_memcheck_*.txt)
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_memcheck_smoke.log 2>&1
timeout 80 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_planet.py tests/test_gpu_alpha.py -m gpu -x -q > gpurun_out/r2zz_memcheck_tests.log 2>&1
timeout 70 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rt.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/r2zz_memcheck_tests_rt.log 2>&1
# 4. shared-memory race checker over smoke() (-> profiles/r2_racecheck_smoke.txt)
timeout 42 compute-sanitizer --tool racecheck --racecheck-report analysis python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zz_racecheck_smoke.log 2>&1
