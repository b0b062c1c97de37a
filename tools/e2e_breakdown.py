#!/usr/bin/env python
"""Where the end-to-end time of Planet.run(freqs, b=0.005) goes (host clock around each stage; GPU box only)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                      # noqa: E402
from radiobear_b200 import _lib  # noqa: E402
from radiobear_b200.planet import Planet  # noqa: E402


def timeit(fn, n=8):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs)
    print('Planet.run                    {:.3f} ms'.format(timeit(lambda: p.run(fl, b=bench.BSTEP, reuse_override='false'))))
    print('Planet.run (alpha reused)     {:.3f} ms'.format(timeit(lambda: p.run(fl, b=bench.BSTEP, reuse_override='true'))))
    print('  set_freqs                   {:.3f} ms'.format(timeit(lambda: p.set_freqs(freqs=fl, freqUnit='GHz'))))
    print('  set_b                       {:.3f} ms'.format(timeit(lambda: p.set_b(b=bench.BSTEP, block=(1, 1)))))
    print('  alpha_layers                {:.3f} ms'.format(timeit(lambda: p.alpha_layers(freqs=p.freqs, atmos=p.atmos, scale=False, get_alpha='calc', save_alpha='none'))))
    pts = np.asarray(p.b, dtype=np.float64)
    print('  asarray(b)                  {:.3f} ms'.format(timeit(lambda: np.asarray(p.b, dtype=np.float64))))
    print('  bright.batch (f32)          {:.3f} ms'.format(timeit(lambda: p.bright.batch(pts, p.freqs, p.atmos[0], p.alpha[0], p.config.orientation, out_f32=True))))
    for nch in (1, 2, 4, 6, 8, 12):
        _lib.get_context().set_rt_chunks(nch)
        print('  bright.batch, {:2d} ray chunks  {:.3f} ms'.format(nch, timeit(lambda: p.bright.batch(pts, p.freqs, p.atmos[0], p.alpha[0], p.config.orientation, out_f32=True))))
    _lib.get_context().set_rt_chunks(0)
    print('  set_header                  {:.3f} ms'.format(timeit(lambda: p.set_header(__import__('datetime').datetime.now(), __import__('datetime').datetime.now()))))


if __name__ == '__main__':
    main()
