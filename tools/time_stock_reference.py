#!/usr/bin/env python
"""Time the UNMODIFIED reference (david-deboer/radiobear, /root/reference) and the oracle port on the same bounded sample
of bench.py's workload (config C4), one host core each, in the build container -- the GPU box has no /root/reference,
so bench.py cannot do this at run time; it reports the record this script writes next to its own same-run port timing.

    python tools/time_stock_reference.py [n_pixels] [n_layers_for_alpha]

Writes profiles/r2_stock_reference_timing.json.
"""
import json
import os
import platform
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))


def main():
    n_pix = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    import make_golden as mg
    cwd = os.getcwd()
    mg.bootstrap()
    p = mg.planet('jupiter')            # the reference first: its planet directory must be the first 'Jupiter' on sys.path
    os.chdir(cwd)
    import bench
    atm, freqs, grid = bench.workload()
    cfg = atm.config
    q = cfg.Rpol / cfg.Req
    mask = bench.on_disc_mask(grid, q)
    n_on = int(mask.sum())
    iy, ix = np.nonzero(mask)
    pick = np.random.default_rng(0).choice(len(iy), n_pix, replace=False)
    pix = [[float(grid[ix[k]]), float(grid[iy[k]])] for k in pick]
    L, F = atm.gas.shape[1], len(freqs)
    # ---- the stock reference: Planet.run on a list of impact points (alpha on all layers, then one ray at a time)
    assert p.atmos[0].gas.shape[1] == L
    t0 = time.perf_counter()
    p.alpha_layers(freqs=list(freqs), atmos=p.atmos)
    t_alpha_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    tb_ref = []
    for b in pix:
        tb_ref.append(np.array(p.bright.single(b, list(freqs), p.atmos[0], p.alpha[0], p.config.orientation)))
    t_pix_ref = time.perf_counter() - t0
    tb_ref = np.array(tb_ref)
    # ---- the oracle port on the same sample, same core
    t0 = time.perf_counter()
    lay = bench._cpu_alpha((list(range(L)),))
    t_alpha_port = time.perf_counter() - t0
    t0 = time.perf_counter()
    tb_port = bench._cpu_pixels((pix, lay, 0))
    t_pix_port = time.perf_counter() - t0
    full_ref = t_alpha_ref + t_pix_ref * (n_on / n_pix)
    full_port = t_alpha_port + t_pix_port * (n_on / n_pix)
    rec = {
        'what': 'stock reference (kind "reference") and oracle port (kind "port") on the same sample of config C4, one core each',
        'host': {'cpu': platform.processor() or platform.machine(), 'cores_used': 1, 'where': 'build container (no GPU)'},
        'sample': 'alpha on all {} layers x {} freqs; geometry + RT on {} of {} on-disc pixels x {} freqs; extrapolated '
                  'linearly in the pixel count'.format(L, F, n_pix, n_on, F),
        'reference': {'alpha_s': t_alpha_ref, 'pixels_s': t_pix_ref, 's_per_pixel': t_pix_ref / n_pix,
                      'full_cube_s': full_ref, 'value': n_on * F / full_ref, 'unit': bench.UNIT},
        'port': {'alpha_s': t_alpha_port, 'pixels_s': t_pix_port, 's_per_pixel': t_pix_port / n_pix,
                 'full_cube_s': full_port, 'value': n_on * F / full_port, 'unit': bench.UNIT},
        'port_over_reference': full_ref / full_port,
        'max_abs_dTb_K_port_vs_reference': float(np.nanmax(np.abs(tb_port - tb_ref))),
    }
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    with open(os.path.join(ROOT, 'profiles', 'r2_stock_reference_timing.json'), 'w') as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == '__main__':
    main()
