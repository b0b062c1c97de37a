"""Planet.run(freqs, b=0.005) with host buffers, one process per RB_FILL_STREAM mode (0 context stream, 1 side stream,
2 highest-priority side stream), next to the box's own D2H rate for the 93 MB cube (the link differs between boxes)."""
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
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs)
    from radiobear_b200 import _lib
    ctx = _lib.get_context(0)
    if os.environ.get('RB_AB_TIMING'):
        ctx.enable_timing(True)
    for _ in range(4):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        p.run(fl, b=bench.BSTEP, reuse_override='false')
        ts.append(1e3 * (time.perf_counter() - t0))
    if os.environ.get('RB_AB_TIMING'):
        print('   device families inside Planet.run (ms):', {w: round(float(np.median(ctx.kernel_ms_history(w, 16))), 4) for w in ('alpha', 'geometry', 'rt')})
    # the link: one 93 MB device -> pinned host copy
    n = 361201 * 64
    d = torch.empty(n, dtype=torch.float32, device='cuda')
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h.copy_(d); torch.cuda.synchronize()
    cs = []
    for _ in range(5):
        t0 = time.perf_counter()
        h.copy_(d, non_blocking=True); torch.cuda.synchronize()
        cs.append(time.perf_counter() - t0)
    print('RB_FILL_STREAM={}: Planet.run median {:.3f} ms  min {:.3f}  max {:.3f}   D2H {:.1f} GB/s ({:.2f} ms for the cube)'.format(
        os.environ.get('RB_FILL_STREAM', 'default'), float(np.median(ts)), min(ts), max(ts), 4e-9 * n / min(cs), 1e3 * min(cs)))


if __name__ == '__main__':
    main()
