"""Planet.run under torchrun: a timeline of one call on every rank on the host's common monotonic clock (start / end of
prefetch, alpha_layers, SharedHostExchange.begin, engine.rt_batch, SharedHostExchange.finish), relative to rank 0's
start of the call (development aid)."""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    import bench
    from radiobear_b200.planet import Planet
    from radiobear_b200 import engine, parallel
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs)
    for _ in range(4):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    ex = parallel.host_exchange()
    marks = []

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def inner(*a, **k):
            t = time.monotonic()
            r = fn(*a, **k)
            marks.append((label, t, time.monotonic()))
            return r
        setattr(obj, name, inner)
    wrap(p.bright, 'prefetch', 'prefetch')
    wrap(p, 'alpha_layers', 'alpha')
    wrap(ex, 'begin', 'begin')
    wrap(engine, 'rt_batch', 'rt_batch')
    wrap(ex, 'finish', 'finish')
    dist.barrier(); torch.cuda.synchronize()
    n = 12
    calls = []
    for _ in range(n):
        del marks[:]
        t0 = time.monotonic()
        p.run(fl, b=bench.BSTEP, reuse_override='false')
        calls.append((t0, time.monotonic(), list(marks)))
    dist.barrier(); torch.cuda.synchronize()
    # rank 0's start of every call
    starts = torch.tensor([c[0] for c in calls], dtype=torch.float64, device='cuda')
    dist.broadcast(starts, src=0)
    starts = starts.cpu().numpy()
    k = n - 3                                            # one call in steady state
    t0, t1, mk = calls[k]
    ref = starts[k]
    line = 'rank {}: call {:+.3f} .. {:+.3f}'.format(rank, 1e3 * (t0 - ref), 1e3 * (t1 - ref))
    for label, a, b in mk:
        line += '  {} {:+.3f}..{:+.3f}'.format(label, 1e3 * (a - ref), 1e3 * (b - ref))
    per = sorted(1e3 * (c[1] - c[0]) for c in calls)
    print(line + '   | median {:.3f} ms'.format(per[n // 2]), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
