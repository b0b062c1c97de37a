"""Host profile of the row-sharded Planet.run under torchrun (rank 0 prints): where the per-call time beyond the device
step goes."""
import cProfile
import os
import pstats
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from radiobear_b200.planet import Planet
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = dist.get_rank()
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs)
    for _ in range(3):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    dist.barrier(); torch.cuda.synchronize()
    n = 20
    t0 = time.perf_counter()
    for _ in range(n):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    dist.barrier(); torch.cuda.synchronize()
    if rank == 0:
        print('world {}: Planet.run {:.3f} ms per call'.format(dist.get_world_size(), 1e3 * (time.perf_counter() - t0) / n))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(n):
        p.run(fl, b=bench.BSTEP, reuse_override='false')
    pr.disable()
    dist.barrier()
    if rank in (0, 1):
        print('---- rank', rank)
        pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
