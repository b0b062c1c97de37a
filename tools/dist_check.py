#!/usr/bin/env python
"""torchrun check (GPU box): the row-sharded Planet.run image equals the single-GPU image bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from radiobear_b200.planet import Planet  # noqa: E402


def main():
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = dist.get_rank()
    atm, freqs, grid = bench.workload()
    p = Planet('jupiter', atmosphere=atm, verbose=False)
    fl = list(freqs[::8])
    ok = True
    for bstep in (0.02, 0.01):
        rv = p.run(fl, b=bstep, reuse_override='false')
        if rank == 0:
            pts = np.asarray(p.b, dtype=np.float64)
            ref = p.bright.batch(pts, p.freqs, p.atmos[0], p.alpha[0], p.config.orientation, out_f32=True)['Tb']
            got = np.asarray(rv.Tb).reshape(ref.shape)
            same = np.array_equal(got, ref, equal_nan=True)
            print('b step', bstep, 'image', rv.Tb.shape, 'identical to the single-GPU result:', same,
                  'on-disc', int(np.sum(got[:, 0] != np.float32(2.725))))
            ok = ok and same
        else:
            ok = ok and rv.Tb is None
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('DIST CHECK', 'PASS' if int(flag.item()) == 1 else 'FAIL')
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
