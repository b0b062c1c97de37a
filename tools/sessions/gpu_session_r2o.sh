( timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12; echo "pytest exit $?" ) 2>&1 | tail -6
python - <<'PY'
import sys, time; sys.path.insert(0,'.')
import bench, numpy as np
from radiobear_b200.planet import Planet
atm, freqs, grid = bench.workload()
p = Planet('jupiter', atmosphere=atm, verbose=False)
for req in ('1:100:5', '1:10:1'):
    for _ in range(3): p.run(req, b='disc', reuse_override='false')
    t=time.perf_counter()
    for _ in range(20): p.run(req, b='disc', reuse_override='false')
    print(req, 'disc: %.3f ms per Planet.run' % ((time.perf_counter()-t)/20*1e3))
print(bench.retrieval_loop(atm))
PY
