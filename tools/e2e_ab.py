"""A/B of Planet.run(freqs, b=0.005): absorption slab resident on the device (default) against through host memory
(RB_ALPHA_RESIDENT=0); host-clock stage times inside Planet.run and the library's own trace (RB_TRACE=1)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import bench
    from radiobear_b200.planet import Planet
    from radiobear_b200 import alpha as rbalpha, brightness as rbbright
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs)
    for _ in range(3):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    torch.cuda.synchronize()
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    torch.cuda.synchronize()
    print('RB_ALPHA_RESIDENT={}: Planet.run {:.3f} ms per call'.format(os.environ.get('RB_ALPHA_RESIDENT', '1'),
                                                                      1e3 * (time.perf_counter() - t0) / n))
    # stage times: wrap the three calls Planet.run makes
    marks = {}

    def wrap(obj, name):
        fn = getattr(obj, name)

        def inner(*a, **k):
            t = time.perf_counter()
            r = fn(*a, **k)
            marks.setdefault(name, []).append(1e3 * (time.perf_counter() - t))
            return r
        setattr(obj, name, inner)
    wrap(p.bright, 'prefetch')
    wrap(p, 'alpha_layers')
    wrap(p.bright, 'batch')
    wrap(p, 'set_b')
    wrap(p, 'set_header')
    for _ in range(n):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    print('   ' + '  '.join('{} {:.3f}'.format(k, float(np.mean(v))) for k, v in marks.items()))


if __name__ == '__main__':
    main()
