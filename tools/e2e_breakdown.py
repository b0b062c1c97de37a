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
    ctx = _lib.get_context()
    ctx.set_rt_chunks(0)
    ctx.enable_timing(True)
    for _ in range(3):
        p.bright.batch(pts, p.freqs, p.atmos[0], p.alpha[0], p.config.orientation, out_f32=True)
    print('  kernels inside bright.batch: geometry {:.3f} ms, integrate {} ms'.format(
        ctx.kernel_ms_history('geometry', 1)[-1], np.round(ctx.kernel_ms_history('rt', 8), 3)))
    import torch
    dev = torch.device('cuda', 0)
    cfg = p.config
    t64 = dict(dtype=torch.float64, device=dev)
    radius_t = torch.tensor(p.atmos[0].property[cfg.LP['R']], **t64)
    nidx = p.atmos[0].property[cfg.LP['N']]
    T_t = torch.tensor(p.atmos[0].gas[cfg.C['T']], **t64)
    b_t = torch.tensor(pts, **t64).contiguous()
    slab_t = torch.tensor(p.alpha[0].slab, **t64).contiguous()
    out_t = torch.empty((len(pts), len(p.freqs)), dtype=torch.float32, device=dev)
    from radiobear_b200 import engine
    orient = [float(cfg.orientation[0]), float(cfg.orientation[1])]

    def dev_only():
        engine.rt_batch_dev(radius_t, nidx[0], nidx[1], b_t, slab_t, T_t, cfg.Req, cfg.Rpol, orient, cfg.gtype, cfg.limb,
                            out_f32=True, out=out_t)
        torch.cuda.synchronize()
    print('  rt_batch_dev + sync           {:.3f} ms'.format(timeit(dev_only)))
    host = torch.empty(out_t.shape, dtype=torch.float32).pin_memory()

    def d2h():
        host.copy_(out_t, non_blocking=True)
        torch.cuda.synchronize()
    print('  D2H of the f32 cube (pinned)  {:.3f} ms'.format(timeit(d2h)))
    hb = torch.from_numpy(pts)

    def h2d():
        b_t.copy_(hb, non_blocking=True)
        torch.cuda.synchronize()
    print('  H2D of b                      {:.3f} ms  pinned={}'.format(timeit(h2d), hb.is_pinned()))
    print('  set_header                  {:.3f} ms'.format(timeit(lambda: p.set_header(__import__('datetime').datetime.now(), __import__('datetime').datetime.now()))))


if __name__ == '__main__':
    main()
