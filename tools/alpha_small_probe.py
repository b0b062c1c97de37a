"""Device time of the C4 absorption request (Jupiter fixture, 64 freqs) for blocks of its layers: what a rank of an
N-GPU run would compute if the slab were sharded by layers (development aid)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from radiobear_b200 import engine, _lib

atm, freqs, grid = bench.workload()
cfg = atm.config
dev = torch.device('cuda:0')
ctx = _lib.get_context()
t64 = dict(dtype=torch.float64, device=dev)
L = atm.gas.shape[1]
forms = [(c, f) for c, f in sorted(cfg.constituent_alpha.items()) if f is not None]
other = {'h2': {'h2state': cfg.h2state}}
freqs_t = torch.tensor(freqs, **t64)
for world in (1, 2, 4, 8):
    for rank in sorted({0, world // 2, world - 1}):
        lo, hi = rank * L // world, (rank + 1) * L // world
        T_t = torch.tensor(atm.gas[cfg.C['T']][lo:hi], **t64)
        P_t = torch.tensor(atm.gas[cfg.C['P']][lo:hi], **t64)
        gas_t = torch.tensor(atm.gas[:, lo:hi], **t64).contiguous()
        out = torch.empty((hi - lo, len(freqs)), **t64)
        ms = []
        for i in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine.alpha_layers_dev(freqs_t, T_t, P_t, gas_t, cfg.C, formalisms=forms, other_dicts=other,
                                    truncate_strength=cfg.truncate_strength, out=out, freqs_host=freqs, ctx=ctx)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(e0.elapsed_time(e1))
        print('world', world, 'rank', rank, 'layers', hi - lo, 'ms %.4f (min %.4f)' % (np.mean(ms), np.min(ms)), flush=True)
