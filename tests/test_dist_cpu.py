"""CPU, world_size 2, gloo: the N > 1 host path -- row partition, per-rank blocks, gather to rank 0, and
the frequency-block all_gather (no kernels; the compute is stubbed by a deterministic function)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radiobear_b200 import parallel, set_utils


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stub_tb(rows, cols, F):
    r = torch.arange(rows[0], rows[1], dtype=torch.float32)[:, None, None]
    c = torch.arange(cols, dtype=torch.float32)[None, :, None]
    f = torch.arange(F, dtype=torch.float32)[None, None, :]
    return r * 1000.0 + c + f / 100.0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        grid = set_utils.image_grid(0.05)
        q = 66854.0 / 71492.0
        parts = parallel.partition_rows(grid, q, world)
        local = _stub_tb(parts[rank], len(grid), 3)
        full = parallel.gather_blocks(local, parts, dst=0)
        fparts = parallel.partition_even(7, world)
        slab = torch.arange(5 * 7, dtype=torch.float64).reshape(5, 7)
        got = parallel.all_gather_freq_blocks(slab[:, fparts[rank][0]:fparts[rank][1]].contiguous(), fparts)
        ok_slab = bool(torch.equal(got, slab))
        if rank == 0:
            ref = _stub_tb((0, len(grid)), len(grid), 3)
            out.put((bool(torch.equal(full, ref)), ok_slab, parts))
        else:
            assert full is None
            out.put((True, ok_slab, parts))
    finally:
        dist.destroy_process_group()


def test_row_sharding_and_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] and r[1] for r in res)
    parts = res[0][2]
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0]


def test_partition_even():
    assert parallel.partition_even(7, 2) == [(0, 4), (4, 7)]
    assert parallel.partition_even(64, 8)[-1] == (56, 64)
    assert parallel.partition_even(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]


def _worker_exchange(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), LOCAL_WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        grid = set_utils.image_grid(0.05)
        q = 66854.0 / 71492.0
        rparts = parallel.partition_rows(grid, q, world)
        n = len(grid)
        parts = [(a * n, b * n) for a, b in rparts]
        ex = parallel.host_exchange()
        ok = ex.usable
        held = []
        for call in range(6):
            # flat pixel blocks [rows_i * n, F] like run_points_sharded delivers them
            local = (_stub_tb(rparts[rank], n, 3) + call).reshape(-1, 3).contiguous()
            got = ex.deliver(local, parts, (3,), np.float32)
            if rank == 0:
                ref = (_stub_tb((0, n), n, 3) + call).reshape(-1, 3).numpy()
                ok = ok and got.shape == ref.shape and bool(np.array_equal(got, ref))
                if call == 0:
                    held.append((got, ref))       # still referenced: the segment must not be recycled
                elif call == 1:
                    # only derived views survive (what Planet.run keeps: Tb.reshape(rows, cols, F)); numpy points
                    # their .base past `got`, at the array that owns the segment
                    held.append((got.reshape(n, n, 3)[:, :, :], ref.reshape(n, n, 3)))
                    held.append((got[:, 0], ref[:, 0]))
                del got
            else:
                ok = ok and got is None
        for got, ref in held:
            ok = ok and bool(np.array_equal(got, ref))
        nseg = len(ex.segments)
        dist.barrier()
        del held
        ex.close()
        out.put((ok, nseg))
    finally:
        dist.destroy_process_group()


def test_shared_host_exchange_world2():
    """Row-sharded delivery through the shared host segment: contents, None on rank 1, and no recycling
    of a segment whose array the caller still holds."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] for r in res)
    # calls 0, 1 are held (call 1 only through derived views) -> own segments; the later calls drop their result
    # before the next one and share a third
    assert max(r[1] for r in res) == 3


def _worker_alpha(rank, world, port, out):
    """Alpha.get_layers with the frequencies split over two ranks.  The kernel call (engine.alpha_layers) is stood in
    for by the oracle so that the host path -- block bounds, all_gather, slab / layers layout -- runs on CPU."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sys.path.insert(0, root)
        from oracle import alpha_oracle as ao
        from radiobear_b200 import alpha as rbalpha, engine
        from radiobear_b200.atmosphere import Atmosphere
        atm = Atmosphere.from_npz(os.path.join(root, 'tests', 'golden', 'atm_jupiter_benchmark.npz'), 'jupiter')
        cfg = atm.config
        keep = slice(0, atm.gas.shape[1], 40)                  # a thin atmosphere: the oracle is the slow part here
        atm.gas = np.ascontiguousarray(atm.gas[:, keep])
        atm.cloud = np.ascontiguousarray(atm.cloud[:, keep])
        calls = []

        def stub(freqs, T, P, gas, gas_dict, cloud=None, cloud_dict=None, formalisms=(), other_dicts=None, **kw):
            calls.append(np.array(freqs))
            lay = ao.get_layers(freqs, gas, cloud, gas_dict, cloud_dict, dict(formalisms), other_dicts=other_dicts,
                                truncate_strength=kw.get('truncate_strength'))
            return np.ascontiguousarray(lay.T)                 # [L][F] like the kernel
        engine.alpha_layers = stub
        freqs = np.linspace(2.0, 40.0, 7)
        a = rbalpha.Alpha(config=cfg, verbose=False, shard=True, shard_axis='freqs')
        a.get_layers(freqs, atm)
        sharded_calls = [c.copy() for c in calls]
        full = a.layers.copy()
        # the default axis: blocks of layers, here with a per-constituent scale that has to be cut the same way
        L = atm.gas.shape[1]
        sc = {'nh3': list(np.linspace(0.5, 1.5, L))}
        seen = []

        def stub_l(freqs, T, P, gas, gas_dict, cloud=None, cloud_dict=None, formalisms=(), other_dicts=None, **kw):
            seen.append((gas.shape[1], None if cloud is None else cloud.shape[1], len(kw['scale']['nh3'])))
            lay = ao.get_layers(freqs, gas, cloud, gas_dict, cloud_dict, dict(formalisms), other_dicts=other_dicts,
                                truncate_strength=kw.get('truncate_strength'))
            j = [c for c, _ in formalisms].index('nh3')
            cube = ao.get_layers(freqs, gas, cloud, gas_dict, cloud_dict, {'nh3': dict(formalisms)['nh3']},
                                 other_dicts=other_dicts, truncate_strength=kw.get('truncate_strength'))
            return np.ascontiguousarray((lay + cube * (np.array(kw['scale']['nh3'])[None, :] - 1.0)).T)
        engine.alpha_layers = stub_l
        al = rbalpha.Alpha(config=cfg, verbose=False, shard=True)
        al.get_layers(freqs, atm, scale=sc)
        llo, lhi = parallel.partition_even(L, world)[rank]
        by_layers = al.layers.copy()
        engine.alpha_layers = stub
        whole = stub(freqs, None, None, atm.gas, cfg.C, cloud=atm.cloud, cloud_dict=cfg.Cl, formalisms=a.formalisms(),
                     other_dicts=a.other_dict, truncate_strength=a.truncate_strength).T      # all frequencies in one call
        lo, hi = parallel.partition_even(7, world)[rank]
        whole_l = stub_l(freqs, None, None, atm.gas, cfg.C, cloud=atm.cloud, cloud_dict=cfg.Cl, formalisms=a.formalisms(),
                         other_dicts=a.other_dict, truncate_strength=a.truncate_strength, scale=sc).T
        checks = [seen[0] == (lhi - llo, lhi - llo, lhi - llo), by_layers.shape == (7, L),
                  bool(np.allclose(by_layers, whole_l, rtol=1e-12, atol=0.0)),
                  len(sharded_calls) == 1, np.array_equal(sharded_calls[0], freqs[lo:hi]),
                  full.shape == (7, atm.gas.shape[1]), bool(np.allclose(full, whole, rtol=1e-12, atol=0.0)),   # numpy SIMD tails: last-bit differences per batch shape
                  a.slab.shape == (atm.gas.shape[1], 7)]
        ok = all(checks)
        # 'auto' leaves a request of this size replicated; the size rule and the environment override
        ok = ok and not parallel.shard_alpha(1000, 64, 2) and parallel.shard_alpha(4096, 4096, 8)
        ok = ok and not parallel.shard_alpha(4096, 4096, 1) and parallel.shard_alpha(10, 10, 2, True)
        out.put(bool(ok) or checks)
    finally:
        dist.destroy_process_group()


def test_alpha_frequency_sharding_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_alpha, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r is True for r in res), res
