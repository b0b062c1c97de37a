import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from radiobear_b200.planet import Planet
atm, freqs, grid = bench.workload()
p = Planet('jupiter', atmosphere=atm, verbose=False)
fl = list(freqs)
for i in range(3):
    p.run(fl, b=bench.BSTEP, reuse_override='false')
pts = np.asarray(p.b, dtype=np.float64)
for i in range(3):
    t0 = time.perf_counter()
    r = p.bright.batch(pts, p.freqs, p.atmos[0], p.alpha[0], p.config.orientation, out_f32=True)
    print('batch total %.3f ms' % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
