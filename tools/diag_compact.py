"""Where do the plain and the compacted launch of the one-frequency integration kernel differ on the C4 cube?
Prints the differing pixels (radius, frequency, values) and checks the geometry of a few of them through compute_ds."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    from conftest import golden, keymap, formalisms_of, TRUNC
    from radiobear_b200 import engine as eng
    a = golden('atm_jupiter.npz')
    im = golden('image_c4_full.npz')
    C = keymap(a['C_keys'])
    LP = keymap(a['LP_keys'])
    g = dict(radius=a['property'][LP['R']], refr_index=a['property'][LP['N']], Req=float(a['Req']),
             Rpol=float(a['Rpol']), orientation=a['orientation'], gtype=str(a['gtype']), limb=str(a['limb']))
    slab = eng.alpha_layers(im['freqs'], a['gas'][C['T']], a['gas'][C['P']], a['gas'], C, formalisms=formalisms_of(a),
                            other_dicts={'h2': {'h2state': 'e'}, 'co': {'coshape': 'voigt'}}, truncate_strength=TRUNC)
    grid = im['grid']
    n = len(grid)
    b = np.ascontiguousarray(np.stack([np.tile(grid, n), np.repeat(grid, n)], axis=1))
    res = {}
    for pairs, compact in ((0, False), (0, True), (1, False), (1, True)):
        eng.set_rt_tuning(pairs, compact)
        got = eng.rt_batch(b=b, alpha_slab=slab, T=a['gas'][C['T']], want_intW=True, **g)
        res[(pairs, compact)] = (got['Tb'].copy(), got['integrated_W'].copy())
    eng.set_rt_tuning(-1, True)
    for pairs in (0, 1):
        x, y = res[(pairs, False)][0], res[(pairs, True)][0]
        d = ~((x == y) | (np.isnan(x) & np.isnan(y)))
        rr, ff = np.nonzero(d)
        print('pairs', pairs, 'differing values', d.sum(), 'rays', len(np.unique(rr)), 'nan-pattern equal',
              np.array_equal(np.isnan(x), np.isnan(y)))
        for r in np.unique(rr)[:12]:
            fs = ff[rr == r]
            print('  ray', r, 'b', b[r], 'r', np.hypot(*b[r]), 'nfreq', len(fs), 'f0', fs[0], 'plain', x[r, fs[0]],
                  'compact', y[r, fs[0]], 'diff', x[r, fs[0]] - y[r, fs[0]], 'lane', r % 32)
        w0, w1 = res[(pairs, False)][1], res[(pairs, True)][1]
        dw = ~((w0 == w1) | (np.isnan(w0) & np.isnan(w1)))
        print('   integrated_W differing', dw.sum())


if __name__ == '__main__':
    main()
