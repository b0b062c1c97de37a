"""Where does the host time of Planet.run(freqs, b=0.005) go?  cProfile over a few calls (GPU box)."""
import cProfile
import os
import pstats
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from radiobear_b200.planet import Planet
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs)
    for _ in range(3):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    torch.cuda.synchronize()
    print('Planet.run: {:.3f} ms per call'.format(1e3 * (time.perf_counter() - t0) / n))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('cumulative').print_stats(28)


if __name__ == '__main__':
    main()
