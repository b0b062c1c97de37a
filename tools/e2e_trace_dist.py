"""Planet.run under torchrun: per-call wall time, device time of the kernel families inside the call, and the speed of
the concurrent device -> shared-host copies alone (development aid)."""
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
    from radiobear_b200 import _lib, parallel
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    ctx = _lib.get_context(local)
    ctx.enable_timing(True)
    fl = list(freqs)
    for _ in range(4):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    dist.barrier(); torch.cuda.synchronize()
    n = 20
    per = []
    for _ in range(n):
        t0 = time.perf_counter()
        p.run(fl, b=bench.BSTEP, reuse_override='false')
        per.append(1e3 * (time.perf_counter() - t0))
    dist.barrier(); torch.cuda.synchronize()
    fam = {w: float(np.median(ctx.kernel_ms_history(w, 16))) for w in ('alpha', 'geometry', 'rt')}
    print('rank {} world {} cw {}: Planet.run median {:.3f} ms  min {:.3f}  device families {}'.format(
        rank, world, os.environ.get('RB_ROW_COPY_WEIGHT', 'default'), sorted(per)[n // 2], min(per), {k: round(v, 3) for k, v in fam.items()}), flush=True)
    if os.environ.get('RB_TRACE_SKIP_COPIES'):
        dist.barrier()
        dist.destroy_process_group()
        return
    # the copies alone: every rank moves its share of a 92 MB image into (a) the shared page-locked segment,
    # (b) a private page-locked buffer, all ranks at once
    ex = parallel.host_exchange()
    rows = 361201
    s, e = rank * rows // world, (rank + 1) * rows // world
    src = torch.zeros((e - s, 64), dtype=torch.float32, device='cuda')
    priv = torch.empty((e - s, 64), dtype=torch.float32).pin_memory()
    for name in ('shared', 'private'):
        ts = []
        for i in range(8):
            if name == 'shared':
                full = ex.begin(rows, (64,), np.float32, pin=True)
                dst = torch.from_numpy(full[s:e])
            else:
                dst = priv
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
            if name == 'shared':
                ex.finish()
        print('rank {} copy to {} host memory: {:.3f} ms for {:.1f} MB ({:.1f} GB/s)'.format(
            rank, name, min(ts[2:]), src.numel() * 4 / 1e6, src.numel() * 4 / 1e6 / min(ts[2:])), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
