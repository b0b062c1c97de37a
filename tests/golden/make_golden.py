#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference; the GPU box does not have it).
The reference (david-deboer/radiobear v2.0.1, pure Python) is imported read-only with a
matplotlib stub ahead on sys.path (atm_modify.py:8 imports pyplot at module level) from a
scratch work directory bootstrapped the way scripts/initial_planet_setup.py:12-30 does.

    python tests/golden/make_golden.py [section ...]

sections: atm fileio rtm_tables plugins plugins_nh3_extra plugins_nh3_full plugins_notrunc alpha rays ray_fields gravity tb neptune uranus image
          image_full ring c3_full c5_saturn doppler   (default: all; image_full .. c5_saturn take ~10 min each on 8 cores)

Every array is float64 exactly as the reference produced it; nothing is post-processed.
Reference defects driven around: log-sweep strings and float image
requests crash in set_utils, so frequencies are passed as lists and image pixels are run
through Brightness.single one by one.
"""
import os
import sys
import shutil
import subprocess
import tempfile
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('RADIOBEAR_REFERENCE', '/root/reference')


def bootstrap():
    work = tempfile.mkdtemp(prefix='rbgolden_')
    stub = os.path.join(work, 'stub', 'matplotlib')
    os.makedirs(stub)
    open(os.path.join(stub, '__init__.py'), 'w').close()
    with open(os.path.join(stub, 'pyplot.py'), 'w') as fp:
        fp.write("def __getattr__(name):\n    raise AttributeError('matplotlib stub: ' + name)\n")
    for planet in ['Jupiter', 'Saturn', 'Uranus', 'Neptune']:
        os.mkdir(os.path.join(work, planet))
        src = os.path.join(REF, 'radiobear', planet)
        for pf in os.listdir(src):
            if pf[0] in ['.', '_']:
                continue
            shutil.copy(os.path.join(src, pf), os.path.join(work, planet, pf))
    for other in ['Logs', 'Output', 'Scratch']:
        os.mkdir(os.path.join(work, other))
    os.chdir(work)
    sys.path.insert(0, os.path.join(work, 'stub'))
    sys.path.insert(0, REF)
    return work


def save(name, **arrays):
    fn = os.path.join(HERE, name)
    np.savez_compressed(fn, **arrays)
    print('  wrote {} ({:.1f} kB)'.format(name, os.path.getsize(fn) / 1e3))


def planet(name, **kw):
    import radiobear as rb
    return rb.planet.Planet(name, plot_atm=False, plot_bright=False, verbose=False, **kw)


def cfg_arrays(p):
    c = p.config
    a = p.atmos[0]
    keysC = sorted(a.config.C, key=lambda k: a.config.C[k])
    keysCl = sorted(a.config.Cl, key=lambda k: a.config.Cl[k])
    keysLP = sorted(a.config.LP, key=lambda k: a.config.LP[k])
    ca = {k: ('none' if v is None else v) for k, v in c.constituent_alpha.items()}
    return dict(gas=a.gas, cloud=a.cloud, property=a.property,
                C_keys=np.array(keysC), Cl_keys=np.array(keysCl), LP_keys=np.array(keysLP),
                Req=c.Req, Rpol=c.Rpol, orientation=np.array(c.orientation[:2], dtype=float),
                gtype=c.gtype, limb=c.limb, h2state=c.h2state, coshape=c.coshape,
                alpha_constituents=np.array(sorted(ca.keys())),
                alpha_formalisms=np.array([ca[k] for k in sorted(ca.keys())]),
                truncate_strength_h2s=np.nan if c.truncate_strength['h2s'] is None else c.truncate_strength['h2s'],
                truncate_strength_ph3=np.nan if c.truncate_strength['ph3'] is None else c.truncate_strength['ph3'],
                cloud_p=np.array([c.water_p, c.ice_p, c.nh4sh_p, c.nh3ice_p, c.h2sice_p, c.ch4_p]),
                distance=c.distance, GM_ref=c.GM_ref, p_ref=c.p_ref)


# ---------------------------------------------------------------------------- sections
def sec_atm():
    """Regridded atmospheres (gas[16,L], cloud[12,L], property[11,L]) + config scalars."""
    save('atm_jupiter.npz', **cfg_arrays(planet('jupiter')))
    save('atm_jupiter_benchmark.npz', **cfg_arrays(planet('jupiter', config_file='config_benchmark.par')))
    save('atm_neptune.npz', **cfg_arrays(planet('neptune')))
    save('atm_saturn.npz', **cfg_arrays(planet('saturn')))
    save('atm_uranus.npz', **cfg_arrays(planet('uranus')))


PLUGIN_FREQS = [0.6, 1.0, 5.2, 10.0, 21.9, 23.7, 25.99, 26.0, 26.5, 29.0, 30.0, 30.000001, 31.5, 33.9,
                34.0, 35.0, 50.0, 100.0, 183.31, 250.0, 500.0, 1000.0]


def _points():
    """(T, P, mixing ratios) test points: the hand-picked sets of the reference's ta.py scripts
    (h2s/ta.py:12-20, ph3/ta.py:5-14, nh3/ta.py:23-36 style) plus Jupiter layers spanning the
    0.01 bar .. 10 kbar grid (incl. the 400..2000 bar NH3 blend region)."""
    j = planet('jupiter')
    a = j.atmos[0]
    C = a.config.C
    pts = []
    for lyr in [0, 50, 200, 333, 400, 500, 600, 700, 760, 767, 800, 850, 883, 884, 900, 950, 999]:
        pts.append(a.gas[:, lyr].copy())
    # hand-picked: deep/warm, cold/low-P, PH3/CO/H2S/H2O rich
    base = a.gas[:, 500].copy()
    for T, P, nh3, h2s, h2o, ph3, co, ch4 in [
            (216.4, 1.009, 1.0e-4, 0.0095, 1.0e-6, 6.0e-7, 1.0e-6, 2.0e-3),
            (150.0, 0.5, 2.0e-4, 3.0e-5, 1.0e-7, 8.2e-4, 1.0e-3, 1.8e-3),
            (300.0, 8.0, 4.0e-4, 9.0e-5, 3.0e-3, 1.0e-6, 0.0, 2.0e-3),
            (80.0, 0.02, 1.0e-9, 1.0e-12, 1.0e-15, 1.0e-7, 1.0e-6, 2.0e-2),
            (52.0, 0.05, 1.0e-8, 1.0e-10, 1.0e-12, 1.0e-9, 1.0e-7, 2.0e-2),
            (1200.0, 900.0, 3.0e-4, 8.0e-5, 2.5e-3, 5.0e-7, 1.0e-6, 2.0e-3),
            (1800.0, 3000.0, 3.0e-4, 8.0e-5, 2.5e-3, 5.0e-7, 1.0e-6, 2.0e-3),
            (500.0, 0.0005, 1.0e-6, 1.0e-6, 1.0e-6, 1.0e-6, 1.0e-4, 1.0e-3),
            (120.0, 0.0999, 1.0e-6, 1.0e-6, 1.0e-6, 1.0e-6, 1.0e-4, 1.0e-3)]:
        g = base.copy()
        g[C['T']], g[C['P']] = T, P
        g[C['NH3']], g[C['H2S']], g[C['H2O']], g[C['PH3']], g[C['CO']], g[C['CH4']] = nh3, h2s, h2o, ph3, co, ch4
        g[C['H2']], g[C['HE']] = 0.862, 0.136
        pts.append(g)
    return np.array(pts), dict(C), dict(a.config.Cl), a


def _plugin_table(truncate):
    """Call every plugin exactly like Alpha.get_alpha_from_calc does (alpha.py:202-213)."""
    import importlib
    pts, C, Cl, atm = _points()
    cpath = os.path.join(REF, 'radiobear', 'constituents')
    mods = {}
    for gas, names in {'nh3': ['nh3_hs', 'nh3_dbs', 'nh3_sjs', 'nh3_hs_sjs', 'nh3_dbs_sjs'],
                       'h2s': ['h2s_ddb'], 'ph3': ['ph3_jh'], 'h2o': ['h2o_bk'], 'co': ['co_ddb'],
                       'h2': ['h2_jj_ddb', 'h2_jj'], 'clouds': ['clouds_idp']}.items():
        sys.path.append(os.path.join(cpath, gas))
        for n in names:
            mods[n] = (gas, importlib.import_module(n))
    out = {'points': pts, 'freqs': np.array(PLUGIN_FREQS), 'C_keys': np.array(sorted(C, key=lambda k: C[k])),
           'Cl_keys': np.array(sorted(Cl, key=lambda k: Cl[k]))}
    tstr = {'h2s': 1e-22, 'ph3': 1e-22} if truncate else {}
    # cloud columns for the clouds plugin: use Jupiter layers that carry clouds + a synthetic column
    cl_pts = []
    for lyr in [300, 350, 400, 450, 500]:
        cl_pts.append(atm.cloud[:, lyr].copy())
    syn = atm.cloud[:, 400].copy()
    syn[3:11] = [1e-6, 2e-6, 3e-6, 4e-6, 5e-6, 6e-6, 0.0, 0.0]
    cl_pts.append(syn)
    cl_pts = np.array(cl_pts)
    out['cloud_points'] = cl_pts
    out['cloud_T'] = np.array([260.0, 280.0, 150.0, 272.9, 273.0, 300.0])
    for name, (gas, mod) in mods.items():
        for units in ['invcm', 'dBperkm']:
            res = []
            if gas == 'clouds':
                od = {'water_p': 1e-4, 'ice_p': 1e-4, 'nh4sh_p': 1e-4, 'nh3ice_p': 1e-4, 'h2sice_p': 1e-4,
                      'ch4_p': 1e-4}
                for x, T in zip(cl_pts, out['cloud_T']):
                    res.append(np.asarray(mod.alpha(PLUGIN_FREQS, T, 1.0, x, Cl, od, units=units,
                                                    truncate_freq=None, truncate_strength=None,
                                                    path=os.path.join(cpath, gas), verbose=False), dtype=float))
            else:
                od = {'h2state': 'e', 'h2newset': True, 'coshape': 'voigt'}
                for g in pts:
                    T, P = g[C['T']], g[C['P']]
                    r = mod.alpha(PLUGIN_FREQS, T, P, g, C, od, units=units, truncate_freq=None,
                                  truncate_strength=tstr.get(gas), path=os.path.join(cpath, gas), verbose=False)
                    res.append(np.asarray(r, dtype=float))
            out['{}__{}'.format(name, units)] = np.array(res)
    # extra option variants
    res_n, res_vvw = [], []
    for g in pts:
        T, P = g[C['T']], g[C['P']]
        res_n.append(np.asarray(mods['h2_jj_ddb'][1].alpha(PLUGIN_FREQS, T, P, g, C, {'h2state': 'n'}, units='invcm',
                                                            truncate_freq=None, truncate_strength=None,
                                                            path='', verbose=False), dtype=float))
        res_vvw.append(np.asarray(mods['co_ddb'][1].alpha(PLUGIN_FREQS, T, P, g, C, {'coshape': 'vvw'}, units='invcm',
                                                          truncate_freq=None, truncate_strength=None,
                                                          path=os.path.join(cpath, 'co'), verbose=False), dtype=float))
    out['h2_jj_ddb_n__invcm'] = np.array(res_n)
    out['co_ddb_vvw__invcm'] = np.array(res_vvw)
    return out


def sec_plugins():
    save('plugins_trunc.npz', **_plugin_table(True))


def sec_plugins_nh3_extra():
    """The remaining NH3 formalisms behind the same plugin API: nh3_kd (prints P on every call, nh3_kd.py:146),
    nh3_sjsd, nh3_bg; extra points cover the 12..20 bar constant switch and the 10..100 bar blend."""
    import importlib
    pts, C, Cl, atm = _points()
    extra = []
    base = atm.gas[:, 500].copy()
    for T, P in [(330.0, 9.99), (335.0, 10.0), (340.0, 11.9), (345.0, 12.0), (350.0, 15.0), (360.0, 19.99),
                 (362.0, 20.0), (365.0, 20.01), (380.0, 35.0), (420.0, 60.0), (500.0, 100.0), (505.0, 100.5)]:
        g = base.copy()
        g[C['T']], g[C['P']] = T, P
        extra.append(g)
    pts = np.concatenate([pts, np.array(extra)])
    cpath = os.path.join(REF, 'radiobear', 'constituents')
    sys.path.append(os.path.join(cpath, 'nh3'))
    out = {'points': pts, 'freqs': np.array(PLUGIN_FREQS), 'C_keys': np.array(sorted(C, key=lambda k: C[k]))}
    for name in ['nh3_kd', 'nh3_sjsd', 'nh3_bg']:
        mod = importlib.import_module(name)
        for units in ['invcm', 'dBperkm']:
            res = []
            for g in pts:
                r = mod.alpha(PLUGIN_FREQS, g[C['T']], g[C['P']], g, C, {}, units=units, truncate_freq=None,
                              truncate_strength=None, path=os.path.join(cpath, 'nh3'), verbose=False)
                res.append(np.asarray(r, dtype=float))
            out['{}__{}'.format(name, units)] = np.array(res)
    save('plugins_nh3_extra.npz', **out)


def sec_plugins_nh3_full():
    """SURVEY 8d "full catalog" variant of config C5: the reference's own NH3 plugins reading an ammonia.npz rebuilt from
    the untrimmed line lists (constituents/txt2npz.py:7-57 with rm = vm = 0: 415 + 1301 + 4198 = 5914 lines) through
    their `path` keyword.  The plugins cache the file in module globals, so this runs in a fresh interpreter."""
    if os.environ.get('RB_GOLDEN_CHILD') != '1':
        subprocess.check_call([sys.executable, os.path.abspath(__file__), 'plugins_nh3_full'],
                              env=dict(os.environ, RB_GOLDEN_CHILD='1'))
        return
    import importlib
    import shutil
    import tempfile
    assert 'nh3_dbs' not in sys.modules and 'nh3_hs' not in sys.modules
    pts, C, Cl, atm = _points()
    cpath = os.path.join(REF, 'radiobear', 'constituents', 'nh3')
    tmp = tempfile.mkdtemp(prefix='rb_fullcat_')
    inv = np.loadtxt(os.path.join(cpath, 'ammonia_inversion.dat'), skiprows=1, unpack=True)
    rot = np.loadtxt(os.path.join(cpath, 'ammonia_rotational.dat'), skiprows=1, unpack=True)
    v2 = np.loadtxt(os.path.join(cpath, 'ammonia_rotovibrational.dat'), skiprows=1, unpack=True)
    np.savez(os.path.join(tmp, 'ammonia.npz'), fo=inv[0], Io=inv[1], Eo=inv[2], gammaNH3o=inv[3], H2HeBroad=inv[4],
             fo_rot=rot[0], Io_rot=rot[1], Eo_rot=rot[2], gNH3_rot=rot[3], gH2_rot=rot[4], gHe_rot=rot[5],
             fo_v2=v2[0], Io_v2=v2[1], Eo_v2=v2[2])
    shutil.copy(os.path.join(cpath, 'nh3.npz'), tmp)
    sys.path.append(cpath)
    out = {'points': pts, 'freqs': np.array(PLUGIN_FREQS), 'C_keys': np.array(sorted(C, key=lambda k: C[k])),
           'nlines': np.array([inv.shape[1], rot.shape[1], v2.shape[1]])}
    for name in ['nh3_hs', 'nh3_dbs', 'nh3_kd', 'nh3_dbs_sjs']:
        mod = importlib.import_module(name)
        for units in ['invcm', 'dBperkm']:
            res = []
            for g in pts:
                r = mod.alpha(PLUGIN_FREQS, g[C['T']], g[C['P']], g, C, {}, units=units, truncate_freq=None,
                              truncate_strength=None, path=tmp, verbose=False)
                res.append(np.asarray(r, dtype=float))
            out['{}__{}'.format(name, units)] = np.array(res)
    shutil.rmtree(tmp)
    save('plugins_nh3_full.npz', **out)


def sec_plugins_h2_orton():
    """h2_orton.alpha called directly (the reference's Alpha cannot: h2_orton.py:120 rejects the truncate_* kwargs
    alpha.py:210-213 passes).  Points cover the three temperature branches (below 40 K: quartic extrapolation,
    40..400 K: cubic spline through the 10 tabulated temperatures, above: scaled h2_jj) for both h2 states;
    the first frequency lies below the lowest tabulated wavenumber (extrapolated, h2_orton.py:84-85)."""
    import importlib
    pts, C, Cl, atm = _points()
    extra = []
    base = atm.gas[:, 500].copy()
    for T, P in [(20.0, 0.01), (35.0, 0.02), (39.999, 0.03), (40.5, 0.05), (61.0, 0.1), (100.0, 0.3), (178.0, 1.0),
                 (251.0, 3.0), (399.0, 9.0), (401.0, 9.5), (650.0, 60.0)]:
        g = base.copy()
        g[C['T']], g[C['P']] = T, P
        g[C['CH4']] = 1.9e-3
        extra.append(g)
    pts = np.concatenate([pts, np.array(extra)])
    cpath = os.path.join(REF, 'radiobear', 'constituents')
    sys.path.append(os.path.join(cpath, 'h2'))
    mod = importlib.import_module('h2_orton')
    freqs = [0.5] + PLUGIN_FREQS
    out = {'points': pts, 'freqs': np.array(freqs), 'C_keys': np.array(sorted(C, key=lambda k: C[k]))}
    for state in ['e', 'n']:
        for units in ['invcm', 'dBperkm']:
            res = []
            for i, g in enumerate(pts):
                r = mod.alpha(freqs, g[C['T']], g[C['P']], g, C, {'h2state': state, 'h2newset': i == 0}, units=units,
                              path=os.path.join(cpath, 'h2'), verbose=False)
                res.append(np.asarray(r, dtype=float))
            out['h2_orton_{}__{}'.format(state, units)] = np.array(res)
    save('plugins_h2_orton.npz', **out)


def sec_plugins_notrunc():
    """No-truncation variant: the reference caches catalogs per process (h2s_ddb.py:11, ph3_jh.py:14),
    so this runs in a fresh interpreter."""
    if os.environ.get('RB_GOLDEN_CHILD') == '1':
        save('plugins_notrunc.npz', **{k: v for k, v in _plugin_table(False).items()
                                       if k.startswith(('h2s', 'ph3', 'points', 'freqs', 'C_keys'))})
        return
    env = dict(os.environ, RB_GOLDEN_CHILD='1')
    subprocess.check_call([sys.executable, os.path.abspath(__file__), 'plugins_notrunc'], env=env)


ALPHA_FREQS = [1.0, 5.2, 10.0, 22.0, 29.9, 30.1, 31.5, 100.0]


def sec_alpha():
    """Alpha.get_layers on the Jupiter default atmosphere: total [F,L] and the per-constituent cube
    (save_alpha='memory', alpha.py:110-131)."""
    j = planet('jupiter')
    al = j.alpha[0]
    al.get_layers(ALPHA_FREQS, j.atmos[0], save_alpha='memory')
    total = np.array(al.layers)
    cube = np.array(al.memory.alpha_data)
    al.get_layers(ALPHA_FREQS, j.atmos[0], scale={'nh3': list(np.linspace(0.5, 1.5, 1000)), 'h2o': [2.0] * 1000})
    scaled_dict = np.array(al.layers)
    al.get_layers(ALPHA_FREQS, j.atmos[0], scale=list(np.linspace(2.0, 0.1, 1000)))
    scaled_list = np.array(al.layers)
    save('alpha_jupiter.npz', freqs=np.array(ALPHA_FREQS), layers=total, cube=cube,
         ordered_constituents=np.array(al.ordered_constituents), layers_scaled_dict=scaled_dict,
         layers_scaled_list=scaled_list)


RAY_B = [[0.0, 0.0], [0.5, 0.3], [0.0, 0.9], [0.97, 0.0], [0.2588190451025207, 0.0], [-0.4, -0.6],
         [0.0, 0.93], [0.98, 0.0], [0.985, 0.0], [0.99, 0.0], [0.0, 0.934], [0.0, 0.936],
         [0.7, 0.7], [0.6, 0.75], [1.0, 0.2], [0.999, 0.0], [0.0, 1.0e-3], [0.3, -0.0]]


def _rays(p, blist):
    from radiobear import raypath
    out = []
    hit = []
    for b in blist:
        ray = raypath.compute_ds(p.atmos[0], b, p.config.orientation, gtype=None, verbose=False)
        if ray.ds is None:
            out.append(np.full(len(p.atmos[0].gas[0]) - 1, -1.0))
            hit.append(0)
        else:
            assert list(ray.layer4ds) == list(range(len(ray.ds))), 'layer4ds not 0..S-1'
            d = np.full(len(p.atmos[0].gas[0]) - 1, -2.0)
            d[:len(ray.ds)] = ray.ds
            out.append(d)
            hit.append(len(ray.ds))
    return np.array(out), np.array(hit)


def sec_rays():
    """raypath.compute_ds: ds[ray, S] (-1 rows: off planet; -2 padding: ray ended early)."""
    j = planet('jupiter')
    ds, nseg = _rays(j, RAY_B)
    n = planet('neptune')       # tilted: orientation 347.67, -29.08
    dsn, nsegn = _rays(n, RAY_B)
    j2 = planet('jupiter', limb='sec')
    dss, nsegs = _rays(j2, RAY_B[:6])
    save('rays.npz', b=np.array(RAY_B), ds_jupiter=ds, nseg_jupiter=nseg, ds_neptune=dsn, nseg_neptune=nsegn,
         ds_jupiter_sec=dss, nseg_jupiter_sec=nsegs)


def sec_ray_fields():
    """raypath.compute_ds: the descriptive fields of a ray beside ds -- r4ds and doppler (raypath.py:186-187, 224) --
    for four Jupiter rays and four rays of tilted Neptune (doppler is one entry longer than ds when the loop ends in a
    break; the first len(ds) entries are kept)."""
    import radiobear as rb
    out = {}
    blist = [[0.0, 0.0], [0.3, 0.2], [0.6, -0.4], [-0.85, 0.3]]
    for name in ('jupiter', 'neptune'):
        p = planet(name)
        r4, dop, nn = [], [], []
        for b in blist:
            ray = rb.raypath.compute_ds(p.atmos[0], b, p.config.orientation)
            n = len(ray.ds)
            S = len(p.atmos[0].gas[0]) - 1
            a, d = np.zeros(S), np.zeros(S)
            a[:n] = ray.r4ds
            d[:n] = ray.doppler[:n]
            r4.append(a)
            dop.append(d)
            nn.append(n)
        out['r4ds_' + name] = np.array(r4)
        out['doppler_' + name] = np.array(dop)
        out['nseg_' + name] = np.array(nn)
        out['omega_m_' + name] = p.config.omega_m
        out['vwlat_' + name] = np.array(p.config.vwlat, dtype=float)     # the zonal wind table (config.zonal), an input
        out['vwdat_' + name] = np.array(p.config.vwdat, dtype=float)
    save('ray_fields.npz', b=np.array(blist), **out)


def _w_gravity_ray(b):
    import radiobear as rb
    j = _W.get('jg') or planet('jupiter')
    _W['jg'] = j
    ray = _quiet(rb.raypath.compute_ds, j.atmos[0], list(b), j.config.orientation, gtype='gravity')
    return np.array(ray.ds), np.array(ray.r4ds)


def sec_gravity():
    """gtype='gravity' (Shape._calcGeoid / _gravity, shape.py:141-221).  The reference builds a scipy orthopoly1d object
    for every Legendre evaluation and marches 0.01 deg at a time, so a mid-latitude ray costs hours; pinned here are
    (a) the shape itself -- rmag, gamma, r, n of calcShape at latitudes up to 0.3 deg (30 march steps) on three layers
    -- and (b) two complete rays close to the equatorial plane (a few march steps per shape)."""
    import radiobear as rb
    j = planet('jupiter')
    atm = j.atmos[0]
    req = atm.property[atm.config.LP['R']]
    geoid = rb.shape.Shape('gravity')
    lats = [0.0, 0.004, 0.01, 0.0100001, 0.05, -0.03, 0.2, -0.25, 0.3]
    layers = [0, 300, 999]
    rows = []
    for l in layers:
        for lat in lats:
            rmag = _quiet(geoid.calcShape, atm, req[l], lat, 17.0)
            rows.append([l, lat, rmag, geoid.gamma] + list(geoid.r) + list(geoid.n))
    blist = [[0.3, 0.001], [-0.6, 0.002]]
    with _pool(len(blist), _w_gravity_init, None) as pool:
        res = pool.map(_w_gravity_ray, blist)
    S = len(req) - 1
    ds = np.full((len(blist), S), -2.0)
    r4 = np.zeros((len(blist), S))
    nn = []
    for k, (d, r) in enumerate(res):
        ds[k, :len(d)] = d
        r4[k, :len(r)] = r
        nn.append(len(d))
    save('gravity.npz', shape_rows=np.array(rows), shape_dlng=17.0, b=np.array(blist), ds=ds, r4ds=r4, nseg=np.array(nn),
         Jn=np.array(j.config.Jn, dtype=float), RJ=j.config.RJ, omega_m=j.config.omega_m,
         vwlat=np.array(j.config.vwlat, dtype=float), vwdat=np.array(j.config.vwdat, dtype=float))


def _w_gravity_init(_):
    _W.clear()


def sec_tb():
    """End-to-end Tb: the reference's own known-answer case (scripts/benchmark.py:10-24), the Jupiter
    default disc spectrum '1:100:5' (config C1), point rays and a limb profile (config C3 subset)."""
    jb = planet('jupiter', config_file='config_benchmark.par')
    freqs = [0.6, 1.25, 2.6, 5.2, 10, 21.9]
    b = [[0.0, 0.0], [np.sin(np.radians(15)), 0.0], [np.sin(np.radians(30)), 0.0], [np.sin(np.radians(45)), 0.0]]
    rv = jb.run(freqs, b=b, reuse_override='False')
    tb_bench = np.array(rv.Tb, dtype=np.float64)
    tb_bench64 = np.array(jb.Tb, dtype=np.float64)
    j = planet('jupiter')
    rv = j.run('1:100:5', b='disc')
    c1_f = np.array(j.freqs, dtype=float)
    c1_tb = np.array(j.Tb, dtype=np.float64)
    c1_alpha = np.array(j.alpha[0].layers)
    prof = dict(tau=j.bright.tau, W=j.bright.W, Tb_lyr=j.bright.Tb_lyr, integrated_W=j.bright.integrated_W)
    fpt = [1.0, 10.0, 22.0, 30.0, 31.0, 100.0]
    bpt = [[0.0, 0.0], [0.5, 0.3], [0.0, 0.9], [0.97, 0.0], [0.985, 0.0], [1.0, 0.2]]
    j.run(fpt, b=bpt)
    pt_tb = np.array(j.Tb, dtype=np.float64)
    j.run(fpt, b='disc')
    disc_tb = np.array(j.Tb, dtype=np.float64)
    # C3 subset: limb profile b='0.0:1.0:0.01<0' -> every 9th ray + the last four; 50 freqs 1..50
    f3 = list(np.linspace(1, 50, 50))
    ball = [[v, 0.0] for v in np.arange(0.0, 1.0 + 0.005, 0.01) if v < 0.995 * (j.config.Rpol / j.config.Req) /
            np.sqrt(0.0 + (j.config.Rpol / j.config.Req)**2)]
    sel = sorted(set(list(range(0, len(ball), 9)) + list(range(len(ball) - 4, len(ball)))))
    b3 = [ball[i] for i in sel]
    j.run(f3, b=b3)
    c3_tb = np.array(j.Tb, dtype=np.float64)
    save('tb.npz', bench_freqs=np.array(freqs, dtype=float), bench_b=np.array(b), bench_tb_f32=tb_bench,
         bench_tb=tb_bench64, c1_freqs=c1_f, c1_tb=c1_tb, c1_alpha=c1_alpha, c1_tau=prof['tau'], c1_W=prof['W'],
         c1_Tb_lyr=prof['Tb_lyr'], c1_integrated_W=prof['integrated_W'], pt_freqs=np.array(fpt), pt_b=np.array(bpt),
         pt_tb=pt_tb, disc_tb=disc_tb, c3_freqs=np.array(f3), c3_b=np.array(b3), c3_nb_total=len(ball), c3_tb=c3_tb)


def sec_neptune():
    """Config C2: Neptune disc-averaged, 200 log-spaced freqs 1..100 GHz (passed as a list: the
    '1;100;200' string crashes in set_utils.py:141)."""
    n = planet('neptune')
    freqs = list(np.logspace(0, 2, 200))
    t0 = time.time()
    n.run(freqs, b='disc')
    print('  neptune run {:.1f} s'.format(time.time() - t0))
    lay = np.array(n.alpha[0].layers)
    save('neptune_c2.npz', freqs=np.array(freqs), tb=np.array(n.Tb, dtype=np.float64),
         alpha_every8=lay[:, ::8], ordered_constituents=np.array(n.alpha[0].ordered_constituents))


def sec_uranus():
    """Uranus (10x solar wet atmosphere, its own cloud file, no tilt): disc-averaged Tb and two points at 8 frequencies
    through the unmodified reference -- a planet none of the BASELINE configs names."""
    u = planet('uranus')
    freqs = [1.0, 3.0, 10.0, 22.0, 30.5, 45.0, 100.0, 200.0]
    t0 = time.time()
    u.run(freqs, b='disc')
    disc = np.array(u.Tb, dtype=np.float64)
    lay = np.array(u.alpha[0].layers)
    pts = [[0.0, 0.0], [0.6, 0.3]]
    u.run(freqs, b=pts)
    print('  uranus runs {:.1f} s'.format(time.time() - t0))
    save('uranus.npz', freqs=np.array(freqs), disc_tb=disc, pts=np.array(pts), pt_tb=np.array(u.Tb, dtype=np.float64),
         alpha=lay, ordered_constituents=np.array(u.alpha[0].ordered_constituents))


def sec_image():
    """Config C4 subset: Jupiter image grid b=0.005 (601x601, set_utils.py:65-77), 64 freqs 1..100 GHz;
    a seeded random subset of on-disc pixels + limb-ring pixels + off-disc pixels through
    Brightness.single (float image requests crash in set_utils.py:78)."""
    j = planet('jupiter')
    freqs = list(np.linspace(1, 100, 64))
    j.alpha_layers(freqs=freqs, atmos=j.atmos)
    bstep = 0.005
    grid = -1.0 * np.flipud(np.arange(bstep, 1.5 + bstep, bstep))
    grid = np.concatenate((grid, np.arange(0.0, 1.5 + bstep, bstep)))
    n = len(grid)
    rng = np.random.default_rng(20261017)
    q = j.config.Rpol / j.config.Req
    xx, yy = np.meshgrid(grid, grid)          # pixel (row=y index, col=x index)
    rr = np.sqrt(xx**2 + (yy / q)**2)
    on = np.argwhere(rr < 0.97)
    ring = np.argwhere((rr >= 0.97) & (rr < 1.01))
    off = np.argwhere(rr >= 1.01)
    pick = np.concatenate([on[rng.choice(len(on), 96, replace=False)],
                           ring[rng.choice(len(ring), 40, replace=False)],
                           off[rng.choice(len(off), 8, replace=False)]])
    tbs = []
    t0 = time.time()
    import contextlib
    import io
    for (iy, ix) in pick:
        with contextlib.redirect_stdout(io.StringIO()):
            tb = j.bright.single([grid[ix], grid[iy]], freqs, j.atmos[0], j.alpha[0], j.config.orientation)
        tbs.append(np.array(tb, dtype=np.float64))
    print('  image subset {:.1f} s'.format(time.time() - t0))
    save('image_c4.npz', freqs=np.array(freqs), grid=grid, pick_iy_ix=pick, tb=np.array(tbs), imsize=n)



# ---------------------------------------------------------------------------- round-2 sections (SURVEY 8d sizes)
# These run the reference over thousands of rays; they are spread over a fork pool (one Planet per worker, the
# reference is single-threaded).  RB_GOLDEN_PROCS sets the pool size (default: all cores).
_W = {}


def _pool(n_items, init, *initargs):
    import multiprocessing as mp
    procs = int(os.environ.get('RB_GOLDEN_PROCS', os.cpu_count() or 1))
    return mp.get_context('fork').Pool(max(1, min(procs, n_items)), initializer=init, initargs=initargs)


def _quiet(fn, *a, **kw):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


def _image_grid(bstep=0.005):
    grid = -1.0 * np.flipud(np.arange(bstep, 1.5 + bstep, bstep))      # set_utils.py:65-77
    return np.concatenate((grid, np.arange(0.0, 1.5 + bstep, bstep)))


def _w_tb_init(freqs):
    j = planet('jupiter')
    _quiet(j.alpha_layers, freqs=freqs, atmos=j.atmos)
    _W['j'], _W['freqs'] = j, freqs


def _w_tb(b):
    j = _W['j']
    tb = _quiet(j.bright.single, list(b), _W['freqs'], j.atmos[0], j.alpha[0], j.config.orientation)
    return np.array(tb, dtype=np.float64)


def sec_image_full():
    """Config C4 at the size SURVEY 8d states: 288 seeded on-disc pixels (r < 0.97), 64 limb-ring pixels
    (0.97 <= r < 1.01: finite / NaN mix) and 16 off-disc pixels of the 601 x 601 grid, 64 freqs 1..100 GHz, each
    through the unmodified Brightness.single (brightness.py:30-126)."""
    freqs = list(np.linspace(1, 100, 64))
    grid = _image_grid()
    rng = np.random.default_rng(2026101702)
    jq = planet('jupiter')
    q = jq.config.Rpol / jq.config.Req
    xx, yy = np.meshgrid(grid, grid)
    rr = np.sqrt(xx**2 + (yy / q)**2)
    on = np.argwhere(rr < 0.97)
    ring = np.argwhere((rr >= 0.97) & (rr < 1.01))
    off = np.argwhere(rr >= 1.01)
    pick = np.concatenate([on[rng.choice(len(on), 288, replace=False)],
                           ring[rng.choice(len(ring), 64, replace=False)],
                           off[rng.choice(len(off), 16, replace=False)]])
    blist = [(grid[ix], grid[iy]) for (iy, ix) in pick]
    t0 = time.time()
    with _pool(len(blist), _w_tb_init, freqs) as pool:
        tbs = pool.map(_w_tb, blist, chunksize=4)
    print('  image_full {} pixels {:.1f} s'.format(len(blist), time.time() - t0))
    save('image_c4_full.npz', freqs=np.array(freqs), grid=grid, pick_iy_ix=pick, tb=np.array(tbs), imsize=len(grid))


def _w_geo_init(_):
    _W['j'] = planet('jupiter')


def _w_geo(b):
    from radiobear import raypath
    j = _W['j']
    ray = _quiet(raypath.compute_ds, j.atmos[0], list(b), j.config.orientation, gtype=None, verbose=False)
    if ray.ds is None:
        return (-1, 0, -1, 0.0)
    ds = np.asarray(ray.ds, dtype=np.float64)
    n = len(ds)
    bad = np.nonzero(np.isnan(ds))[0]
    first_nan = int(bad[0]) if len(bad) else -1
    used_nan = int(first_nan >= 0 and first_nan <= n - 2)          # Brightness.single uses ds[0 .. n-2]
    return (n, used_nan, first_nan, float(np.nansum(ds)))


def sec_ring():
    """Geometry-only classification of EVERY pixel of the limb ring of the C4 grid in one quadrant (x >= 0, y >= 0;
    the orientation is (0, 0), so the other three are mirror images): 0.93 < r < 1.02 with r = sqrt(x^2 + (y/q)^2).
    Per pixel: number of segments raypath.compute_ds returns (-1: ray misses), whether a segment Brightness.single
    uses is NaN (=> Tb is NaN), index of the first NaN segment, nansum(ds).  raypath.py:108-273."""
    grid = _image_grid()
    jq = planet('jupiter')
    q = jq.config.Rpol / jq.config.Req
    xx, yy = np.meshgrid(grid, grid)
    rr = np.sqrt(xx**2 + (yy / q)**2)
    sel = np.argwhere((rr > 0.93) & (rr < 1.02) & (xx >= 0.0) & (yy >= 0.0))
    blist = [(grid[ix], grid[iy]) for (iy, ix) in sel]
    t0 = time.time()
    with _pool(len(blist), _w_geo_init, 0) as pool:
        res = pool.map(_w_geo, blist, chunksize=16)
    print('  ring {} pixels {:.1f} s'.format(len(blist), time.time() - t0))
    res = np.array(res)
    save('ring_quadrant.npz', grid=grid, iy_ix=sel, nseg=res[:, 0].astype(np.int32), used_nan=res[:, 1].astype(np.int8),
         first_nan=res[:, 2].astype(np.int32), nansum_ds=res[:, 3], q=q)


def sec_c3_full():
    """Config C3 in full: Planet.run(freqs 1..50 GHz (50), b='0.0:1.0:0.01<0') -- all 100 limb-profile rays the
    reference derives from the string (set_utils.py:52-63), each through Brightness.single."""
    freqs = list(np.linspace(1, 50, 50))
    jq = planet('jupiter')
    from radiobear import set_utils
    bq = set_utils.set_b('0.0:1.0:0.01<0', [1, 1], Rpol=jq.config.Rpol, Req=jq.config.Req)
    blist = [tuple(b) for b in bq.b]
    t0 = time.time()
    with _pool(len(blist), _w_tb_init, freqs) as pool:
        tbs = pool.map(_w_tb, blist, chunksize=2)
    print('  c3_full {} rays {:.1f} s'.format(len(blist), time.time() - t0))
    save('c3_full.npz', freqs=np.array(freqs), b=np.array(blist), tb=np.array(tbs), data_type=np.array(bq.data_type))


def _w_c5_init(_):
    import importlib
    s = planet('saturn', regridType=4096)
    cpath = os.path.join(REF, 'radiobear', 'constituents', 'nh3')
    sys.path.append(cpath)
    _W['s'], _W['mod'], _W['path'] = s, importlib.import_module('nh3_dbs_sjs'), cpath


def _w_c5(args):
    lyr, freqs = args
    a = _W['s'].atmos[0]
    C = a.config.C
    g = a.gas[:, lyr]
    r = _quiet(_W['mod'].alpha, freqs, g[C['T']], g[C['P']], g, C, {}, units='invcm', truncate_freq=None,
               truncate_strength=None, path=_W['path'], verbose=False)
    return np.asarray(r, dtype=np.float64)


def sec_c5_saturn():
    """Config C5 on its concrete input: Planet('saturn', regridType=4096) x np.linspace(1, 100, 4096) x NH3 only
    (nh3_dbs_sjs, nh3_dbs_sjs.py:6-26 called like alpha.py:210-213): 96 layers -- every 64th of the 4096 plus 16 around
    each of the 400 bar and 2000 bar switches of the pressure blend."""
    s = planet('saturn', regridType=4096)
    a = s.atmos[0]
    P = a.gas[a.config.C['P']]
    L = len(P)
    freqs = list(np.linspace(1, 100, 4096))
    i400, i2000 = int(np.argmin(np.abs(P - 400.0))), int(np.argmin(np.abs(P - 2000.0)))
    lyrs = sorted(set(list(range(0, L, 64)) + list(range(max(0, i400 - 8), min(L, i400 + 8))) +
                      list(range(max(0, i2000 - 8), min(L, i2000 + 8)))))
    t0 = time.time()
    with _pool(len(lyrs), _w_c5_init, 0) as pool:
        res = pool.map(_w_c5, [(l, freqs) for l in lyrs], chunksize=1)
    print('  c5_saturn {} layers x {} freqs {:.1f} s'.format(len(lyrs), len(freqs), time.time() - t0))
    save('c5_saturn.npz', layers=np.array(lyrs, dtype=np.int32), freqs=np.array(freqs), alpha=np.array(res),
         gas=a.gas, C_keys=np.array(sorted(a.config.C, key=lambda k: a.config.C[k])))


FILEIO_CASES = [('spectrum', [[0.0, 0.0], [0.5, 0.25]], [[100.123, 200.5, 300.25], [90.1, 80.2, 70.3]]),
                ('spectrum', ['disc'], [[100.123, 200.5, 300.25]]),
                ('profile', [[0.1 * i, 0.0] for i in range(6)], [[100.0 + i, 200.5 + i, 300.25 + i] for i in range(6)]),
                ('image', [[0, 0]] * 25, [[float(i * j) + 0.123 for i in range(4)] for j in range(3)])]


def sec_doppler():
    """Doppler-shifted absorption (brightness.py:80-96, config key `doppler`).  The branch calls `alpha.get_alpha`, the
    name `Alpha.get_alpha_from_calc` (alpha.py:194) had in earlier versions; with today's class it raises AttributeError
    (the reference prints "Doppler currently broken since the get_alpha call is different").  ONE change is made to run
    it, here and not in the reference tree: the old name is restored as a method that returns the total absorption of
    its single frequency -- the sum over constituents of get_alpha_from_calc's row, which is what the arithmetic under
    the call expects (dtau = (a0 + a1) * ds / 2 with numbers).  Everything else is the reference's own loop.
    Jupiter's rotation moves the frequencies by 4e-5 at most; omega_m is multiplied by 100 for the fixture so that the
    branch changes Tb by far more than the comparison tolerance (the factor is stored)."""
    from radiobear import alpha as ralpha

    def get_alpha(self, freqs, T, P, gas, gas_dict, cloud, cloud_dict, units='invcm'):
        return float(self.get_alpha_from_calc(freqs, T, P, gas, gas_dict, cloud, cloud_dict, units)[0].sum())
    ralpha.Alpha.get_alpha = get_alpha
    factor = 100.0
    freqs = [4.0, 22.0, 23.9]
    bpts = [[0.5, 0.0], [-0.8, 0.3], [0.0, 0.6], [1.2, 1.2]]
    try:
        j = planet('jupiter')
        j.config.omega_m = j.config.omega_m * factor
        out = {}
        for flag in (False, True):
            j.config.Doppler = flag
            rv = j.run(freqs, b=bpts, reuse_override='False')
            out[flag] = np.array(j.Tb, dtype=np.float64)
        dop = [np.array(ray_doppler(j, b)) for b in bpts[:3]]
        j.config.Doppler = True
        j.run(freqs, b='disc', reuse_override='False')                 # b = (0, 0): delta_lng = 0, no shift
        disc = np.array(j.Tb, dtype=np.float64)
    finally:
        del ralpha.Alpha.get_alpha
    print('  Doppler - plain (K):', np.abs(out[True] - out[False]).max(axis=1))
    save('doppler.npz', freqs=np.array(freqs), b=np.array(bpts), omega_factor=factor, tb_doppler=out[True],
         tb_plain=out[False], tb_disc=disc, doppler0=dop[0], doppler1=dop[1], doppler2=dop[2])


def ray_doppler(p, b):
    from radiobear import raypath as ray
    return ray.compute_ds(p.atmos[0], b, p.config.orientation, gtype=None, verbose=False).doppler


def data_show(d):
    """What Data.show / show_header / show_log print (data_handling.py:51-85) for the last FILEIO case with a log file."""
    import contextlib
    import io
    with open('show_case.log', 'w') as fp:
        fp.write('  first line  \nsecond\n')
    d.set('logfile', 'show_case.log')
    d.set('start', 'START')
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        d.show()
        d.show(include=['f', 'nonsense', 'type'], indent=0)
        d.show_header(indent=2)
    return buf.getvalue()


def sec_fileio():
    """fileIO.FileIO.write (fileIO.py:19-107): the text a reference run writes for each output type."""
    from radiobear import fileIO, data_handling
    out = {}
    for n, (typ, b, Tb) in enumerate(FILEIO_CASES):
        d = data_handling.Data()
        d.set('f', [1.0, 10.5, 22.0])
        d.set('freqUnit', 'GHz')
        d.set('b', b)
        d.set('Tb', Tb)
        d.set('type', typ)
        d.set('header', {'z': '# z line', 'a': '# a line', 'data-type': '#* type:  ' + typ})
        fn = fileIO.FileIO().write('fileio_case.dat', d)
        out['case{}'.format(n)] = np.array(open(fn).read())
        out['type{}'.format(n)] = np.array(typ)
    out['show'] = np.array(data_show(d))
    save('fileio.npz', **out)


def rtm_state(B):
    """Side attributes of a finished Brightness.single call (brightness.py:114-123), seeded, 3 freqs x 7 segments,
    set on a Brightness instance of either implementation."""
    rng = np.random.default_rng(11)
    nF, n = 3, 7
    B.freqs = [1.0, 22.235, 100.5]
    B.P = np.geomspace(0.01, 3000.0, n) * (1.0 + 0.1 * rng.random(n))
    B.z = np.linspace(250.123, -800.987, n)
    B.tau = np.cumsum(rng.random((nF, n)) * 3.0, axis=1)
    B.W = rng.random((nF, n)) * np.exp(-B.tau)
    B.Tb_lyr = np.cumsum(rng.random((nF, n)) * 60.0, axis=1)
    B.alpha = type('A', (), {})()
    B.alpha.layers = rng.random((nF, n + 1)) * 1e-5
    return B


RTM_CALLS = [('savertm', (), {}), ('savertm', (None, '.'), {}), ('savertm', ('t1', '.'), {}), ('savertm', ('t2', 'wgt_path.out'), {}),
             ('saveWeight', (True,), {}), ('saveWeight', (False, 'w.out', 'nowhere'), {}), ('saveTau', ('tau2.out', 'nowhere'), {}),
             ('saveit', (), {})]


def rtm_run(B, outdir):
    """Run RTM_CALLS in an empty working directory -> {'<call index>/<relative file name>': text, 'ret<i>': repr}."""
    out = {}
    here = os.getcwd()
    for n, (name, a, kw) in enumerate(RTM_CALLS):
        d = tempfile.mkdtemp(prefix='rtm_', dir=outdir)
        os.mkdir(os.path.join(d, 'Output'))
        B.config = type('C', (), {'output_directory': 'Output'})()
        os.chdir(d)
        try:
            ret = repr(getattr(B, name)(*a, **kw))
        except Exception as e:                                  # savertm's default path=None fails in saveTau
            ret = 'raised ' + type(e).__name__
        finally:
            os.chdir(here)
        out['ret{}'.format(n)] = np.array(ret)
        for root, _, files in os.walk(d):
            for fn in files:
                rel = os.path.relpath(os.path.join(root, fn), d)
                out['{}/{}'.format(n, rel)] = np.array(open(os.path.join(root, fn)).read())
    return out


def sec_rtm_tables():
    """Brightness.savertm / saveAlpha / saveWeight / saveTau / saveTblayer / saveit (brightness.py:128-250): the files
    (names, locations, bytes) and return values the reference leaves for a fixed state."""
    import gc
    from radiobear import brightness
    B = rtm_state(object.__new__(brightness.Brightness))
    out = rtm_run(B, os.getcwd())
    gc.collect()
    save('rtm_tables.npz', **out)


SECTIONS = {'atm': sec_atm, 'rtm_tables': sec_rtm_tables, 'fileio': sec_fileio, 'plugins_nh3_extra': sec_plugins_nh3_extra, 'plugins_nh3_full': sec_plugins_nh3_full, 'plugins_h2_orton': sec_plugins_h2_orton, 'plugins': sec_plugins, 'plugins_notrunc': sec_plugins_notrunc, 'alpha': sec_alpha,
            'rays': sec_rays, 'ray_fields': sec_ray_fields, 'gravity': sec_gravity, 'tb': sec_tb, 'neptune': sec_neptune, 'uranus': sec_uranus, 'image': sec_image,
            'image_full': sec_image_full, 'ring': sec_ring, 'c3_full': sec_c3_full, 'c5_saturn': sec_c5_saturn,
            'doppler': sec_doppler}

if __name__ == '__main__':
    import warnings
    warnings.filterwarnings('ignore')
    todo = sys.argv[1:] or list(SECTIONS)
    work = bootstrap()
    import contextlib
    import io
    for s in todo:
        t0 = time.time()
        print('[{}]'.format(s))
        buf = io.StringIO()
        if s in ('plugins_notrunc', 'plugins_nh3_full') and os.environ.get('RB_GOLDEN_CHILD') != '1':
            SECTIONS[s]()
        else:
            with contextlib.redirect_stdout(buf):
                try:
                    SECTIONS[s]()
                finally:
                    sys.stderr.write(''.join(l + '\n' for l in buf.getvalue().splitlines() if l.startswith('  ')))
        print('  {:.1f} s'.format(time.time() - t0))
    shutil.rmtree(work, ignore_errors=True)
