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
    # absorption sharded by layer blocks (the all_gather inside the kernel over symmetric memory, else NCCL) and by
    # frequency blocks against the replicated computation, with a per-constituent scale
    from radiobear_b200.alpha import Alpha
    from radiobear_b200 import parallel
    L = atm.gas.shape[1]
    sc = {'nh3': list(np.linspace(0.5, 1.5, L)), 'h2o': [2.0] * L}
    ref = Alpha(config=atm.config, verbose=False, shard=False)
    ref.get_layers(fl, atm, scale=sc)
    for axis in ('layers', 'freqs'):
        a = Alpha(config=atm.config, verbose=False, shard=True, shard_axis=axis)
        a.get_layers(fl, atm, scale=sc)
        same = np.array_equal(a.layers, ref.layers)
        if rank == 0:
            used = parallel.symmetric_slab(L, len(fl), torch.device('cuda', local)).usable if axis == 'layers' else False
            print('alpha sharded by', axis, 'identical to the replicated slab:', same, '(in-kernel gather: {})'.format(used))
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print('DIST CHECK', 'PASS' if int(flag.item()) == 1 else 'FAIL')
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == '__main__':
    main()
