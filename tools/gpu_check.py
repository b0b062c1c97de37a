#!/usr/bin/env python
"""Quick GPU parity probe (development aid; the real checks live in tests/ -m gpu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from radiobear_b200 import engine, _lib
from oracle import alpha_oracle as ao, ray_oracle as ro, rt_oracle as rto

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
ctx = _lib.get_context(0)
ctx.enable_timing(True)

def relerr(a, r):
    with np.errstate(divide='ignore', invalid='ignore'):
        e = np.abs(a - r) / np.abs(r)
    e[(a == r)] = 0
    e[np.isnan(a) & np.isnan(r)] = 0
    return e

# --- plugins
g = np.load(os.path.join(G, 'plugins_trunc.npz'))
pts, freqs = g['points'], g['freqs']
C = {k: i for i, k in enumerate(g['C_keys'])}
od = {'h2state': 'e', 'coshape': 'voigt'}
fam = {'nh3_hs': 'nh3', 'nh3_dbs': 'nh3', 'nh3_sjs': 'nh3', 'nh3_hs_sjs': 'nh3', 'nh3_dbs_sjs': 'nh3', 'h2s_ddb': 'h2s',
       'ph3_jh': 'ph3', 'h2o_bk': 'h2o', 'co_ddb': 'co', 'h2_jj_ddb': 'h2', 'h2_jj': 'h2'}
gas = np.ascontiguousarray(pts.T)
for name, c in fam.items():
    for units in ['invcm', 'dBperkm']:
        ref = g[name + '__' + units]
        out = engine.alpha_layers(freqs, gas[C['T']], gas[C['P']], gas, C, formalisms=[(c, name)], other_dicts={c: od},
                                  units=units, truncate_strength={'h2s': 1e-22, 'ph3': 1e-22})
        e = relerr(out, ref)
        print('{:12s} {:8s} max rel {:.3e}  (nan mismatch {})'.format(name, units, np.nanmax(e), int(np.isnan(e).sum())))

# --- jupiter alpha cube
a = np.load(os.path.join(G, 'atm_jupiter.npz'))
al = np.load(os.path.join(G, 'alpha_jupiter.npz'))
Cj = {k: i for i, k in enumerate(a['C_keys'])}
forms = [(c, f) for c, f in zip(a['alpha_constituents'], a['alpha_formalisms']) if f != 'none']
t0 = time.time()
tot, cube = engine.alpha_layers(al['freqs'], a['gas'][Cj['T']], a['gas'][Cj['P']], a['gas'], Cj, formalisms=forms,
                                other_dicts={'h2': {'h2state': str(a['h2state'])}}, want_cube=True,
                                truncate_strength={'h2s': 1e-22, 'ph3': 1e-22})
print('jupiter alpha wall', time.time() - t0, 'kernel ms', ctx.last_kernel_ms('alpha'))
print('total  max rel', np.max(relerr(tot.T, al['layers'])))
print('cube   max rel', np.nanmax(relerr(cube, al['cube'])), [str(x) for x in al['ordered_constituents']], forms)

# --- rays
gr = np.load(os.path.join(G, 'rays.npz'))
for planet in ['jupiter', 'neptune']:
    a = np.load(os.path.join(G, 'atm_{}.npz'.format(planet)))
    LP = {k: i for i, k in enumerate(a['LP_keys'])}
    req, nr = a['property'][LP['R']], a['property'][LP['N']]
    ds, nseg, asp = engine.compute_ds(req, nr, gr['b'], float(a['Req']), float(a['Rpol']), a['orientation'],
                                      str(a['gtype']), str(a['limb']))
    for i, b in enumerate(gr['b']):
        ns = gr['nseg_' + planet][i]
        if ns == 0:
            print(planet, b, 'miss', nseg[i]); continue
        ref = gr['ds_' + planet][i][:ns]
        d = ds[i][:ns]
        m = ~np.isnan(ref)
        print(planet, b, 'nseg', nseg[i], ns, 'nan same', np.array_equal(np.isnan(d), np.isnan(ref)), 'max rel',
              np.max(np.abs(d[m] - ref[m]) / ref[m]) if np.array_equal(np.isnan(d), np.isnan(ref)) else 'n/a')

# --- Tb benchmark
a = np.load(os.path.join(G, 'atm_jupiter_benchmark.npz'))
tb = np.load(os.path.join(G, 'tb.npz'))
Cj = {k: i for i, k in enumerate(a['C_keys'])}
LP = {k: i for i, k in enumerate(a['LP_keys'])}
forms = [(c, f) for c, f in zip(a['alpha_constituents'], a['alpha_formalisms']) if f != 'none']
slab = engine.alpha_layers(tb['bench_freqs'], a['gas'][Cj['T']], a['gas'][Cj['P']], a['gas'], Cj, formalisms=forms,
                           other_dicts={'h2': {'h2state': str(a['h2state'])}}, truncate_strength={'h2s': 1e-22, 'ph3': 1e-22})
res = engine.rt_batch(a['property'][LP['R']], a['property'][LP['N']], tb['bench_b'], slab, a['gas'][Cj['T']],
                      float(a['Req']), float(a['Rpol']), a['orientation'], str(a['gtype']), str(a['limb']))
print('benchmark Tb max abs err K', np.max(np.abs(res['Tb'] - tb['bench_tb'])))
print(res['Tb'])
