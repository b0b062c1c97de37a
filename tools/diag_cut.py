"""Diagnostic (GPU box): tau_cut = 5 through the FP64 and the mixed rays-major kernels against a numpy restatement
of the intended rule (integrate through the step that crosses tau_cut, then stop)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import golden, keymap, formalisms_of, TRUNC
from radiobear_b200 import engine as eng
from oracle import ray_oracle as ro

a = golden('atm_jupiter.npz'); im = golden('image_c4.npz')
C = keymap(a['C_keys']); LP = keymap(a['LP_keys'])
T = a['gas'][C['T']]
g = dict(radius=a['property'][LP['R']], refr_index=a['property'][LP['N']], Req=float(a['Req']), Rpol=float(a['Rpol']),
         orientation=a['orientation'], gtype=str(a['gtype']), limb=str(a['limb']))
freqs = im['freqs'][:8]
slab = eng.alpha_layers(freqs, T, a['gas'][C['P']], a['gas'], C, formalisms=formalisms_of(a),
                        other_dicts={'h2': {'h2state': 'e'}, 'co': {'coshape': 'voigt'}}, truncate_strength=TRUNC)
rng = np.random.default_rng(17)
b = np.ascontiguousarray(rng.uniform(-0.6, 0.6, (600, 2)))
CUT = 5.0
eng.set_rt_precision('f64')
r64 = eng.rt_batch(b=b, alpha_slab=slab, T=T, tau_cut=CUT, want_intW=True, **g)
r64 = {k: v.copy() for k, v in r64.items()}
small = eng.rt_batch(b=b[:8], alpha_slab=slab, T=T, tau_cut=CUT, want_intW=True, **g)
small = {k: v.copy() for k, v in small.items()}
eng.set_rt_precision('mixed')
rmx = eng.rt_batch(b=b, alpha_slab=slab, T=T, tau_cut=CUT, want_intW=True, **g)
rmx = {k: v.copy() for k, v in rmx.items()}
eng.set_rt_precision('f64')
ds_all, nseg, _ = eng.compute_ds(b=b[:8], **{k: v for k, v in g.items()})
for r in range(4):
    ds = ds_all[r][:nseg[r]] * 1e5
    n = nseg[r]
    for f in range(len(freqs)):
        al = slab[:, f]
        tau = 0.0; iW = 0.0; Tb = 0.0; stop_at = -1
        for i in range(n - 1):
            tau += 0.5 * (al[i] + al[i + 1]) * ds[i]
            dd = ds[i] + (ds[i + 1] if i + 1 < n - 1 else 0.0)
            w = np.exp(-tau) * dd * 0.5
            iW += al[i + 1] * w; Tb += T[i + 1] * al[i + 1] * w
            if tau > CUT:
                stop_at = i; break
        print('ray %d f %d  intended Tb %.6f iW %.8f stop %d | f64 %.6f %.8f | mixed %.6f %.8f | small %.6f %.8f' % (
            r, f, Tb / iW, iW, stop_at, r64['Tb'][r, f], r64['integrated_W'][r, f], rmx['Tb'][r, f], rmx['integrated_W'][r, f],
            small['Tb'][r, f], small['integrated_W'][r, f]))
d = np.abs(r64['Tb'] - rmx['Tb'])
print('max |f64 - mixed| with tau_cut 5:', np.nanmax(d), 'fraction of entries above 1e-3:', float(np.mean(d > 1e-3)))
