"""CPU: pin the oracle (oracle/) to the reference.

Golden vectors come from the unmodified reference run in the build container
(tests/golden/make_golden.py) plus the reference's own known-answer table
(scripts/benchmark.py:20-24).  The oracle is float64 numpy in the reference's operation order, so the
bars here are tight: 1e-12 relative for absorption, bit-exact NaN pattern and 1e-12 for ray segments,
1e-6 K for brightness temperatures.
"""
import numpy as np
import pytest

from conftest import golden, keymap, relerr, formalisms_of, TRUNC
from oracle import alpha_oracle as ao
from oracle import ray_oracle as ro
from oracle import rt_oracle as rto

# scripts/benchmark.py:20-24 (float32 values printed by the reference authors)
T_RB = np.array([[805.38385, 424.58337, 298.74915, 227.0338, 179.9148, 147.3627],
                 [790.66736, 420.24728, 296.62033, 225.69223, 179.22531, 147.24104],
                 [745.55975, 407.28714, 290.07492, 221.56863, 177.13437, 146.8625],
                 [668.4528, 385.59015, 278.51672, 214.30847, 173.56506, 146.17966]])

GAS_PLUGINS = ['nh3_hs', 'nh3_dbs', 'nh3_sjs', 'nh3_hs_sjs', 'nh3_dbs_sjs', 'h2s_ddb', 'ph3_jh', 'h2o_bk', 'co_ddb',
               'h2_jj_ddb', 'h2_jj']


@pytest.mark.parametrize('name', GAS_PLUGINS)
@pytest.mark.parametrize('units', ['invcm', 'dBperkm'])
def test_plugin_matches_reference(name, units):
    g = golden('plugins_trunc.npz')
    C = keymap(g['C_keys'])
    od = {'h2state': 'e', 'coshape': 'voigt'}
    ref = g['{}__{}'.format(name, units)]
    for p, r in zip(g['points'], ref):
        kw = dict(units=units)
        if name.startswith(('h2s', 'ph3')):
            kw['truncate_strength'] = 1e-22
        a = ao.FORMALISMS[name](g['freqs'], p[C['T']], p[C['P']], p, C, od, **kw)
        assert np.array_equal(np.isnan(a), np.isnan(r))
        assert np.nanmax(relerr(a, r)) < 1e-12


def test_plugin_option_variants():
    g = golden('plugins_trunc.npz')
    C = keymap(g['C_keys'])
    for p, rn, rv in zip(g['points'], g['h2_jj_ddb_n__invcm'], g['co_ddb_vvw__invcm']):
        a = ao.h2_jj_ddb(g['freqs'], p[C['T']], p[C['P']], p, C, {'h2state': 'n'}, units='invcm')
        assert np.max(relerr(a, rn)) < 1e-13
        a = ao.co_ddb(g['freqs'], p[C['T']], p[C['P']], p, C, {'coshape': 'vvw'}, units='invcm')
        assert np.max(relerr(a, rv)) < 1e-12


def test_plugins_without_truncation():
    g = golden('plugins_notrunc.npz')
    C = keymap(g['C_keys'])
    for name in ['h2s_ddb', 'ph3_jh']:
        for p, r in zip(g['points'], g[name + '__invcm']):
            a = ao.FORMALISMS[name](g['freqs'], p[C['T']], p[C['P']], p, C, {}, units='invcm')
            assert np.max(relerr(a, r)) < 1e-12
    cat = ao.default_catalog()
    assert cat.get('h2s').shape[1] == 200 and cat.get('h2s', 1e-22).shape[1] == 121
    assert cat.get('ph3').shape[1] == 320 and cat.get('ph3', 1e-22).shape[1] == 33


def test_clouds_plugin():
    g = golden('plugins_trunc.npz')
    Cl = keymap(g['Cl_keys'])
    od = {'water_p': 1e-4, 'ice_p': 1e-4, 'nh4sh_p': 1e-4, 'nh3ice_p': 1e-4, 'h2sice_p': 1e-4, 'ch4_p': 1e-4}
    for units in ['invcm', 'dBperkm']:
        for x, T, r in zip(g['cloud_points'], g['cloud_T'], g['clouds_idp__' + units]):
            a = ao.clouds_idp(g['freqs'], T, 1.0, x, Cl, od, units=units)
            assert np.max(relerr(a, r)) < 1e-13


def test_get_layers_jupiter_cube_and_scaling():
    a = golden('atm_jupiter.npz')
    al = golden('alpha_jupiter.npz')
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    ca = dict(formalisms_of(a))
    lay = list(range(0, 1000, 7)) + [997, 998, 999]
    kw = dict(other_dicts={'h2': {'h2state': str(a['h2state'])}}, truncate_strength=TRUNC, layers=lay)
    tot, cube, ordered = ao.get_layers(al['freqs'], a['gas'], a['cloud'], C, Cl, ca, return_per_constituent=True, **kw)
    assert ordered == [str(x) for x in al['ordered_constituents']]
    assert np.max(relerr(tot, al['layers'][:, lay])) < 1e-12
    assert np.nanmax(relerr(cube, al['cube'][lay])) < 1e-12
    sc = {'nh3': list(np.linspace(0.5, 1.5, 1000)), 'h2o': [2.0] * 1000}
    tot = ao.get_layers(al['freqs'], a['gas'], a['cloud'], C, Cl, ca, scale=sc, **kw)
    assert np.max(relerr(tot, al['layers_scaled_dict'][:, lay])) < 1e-12
    tot = ao.get_layers(al['freqs'], a['gas'], a['cloud'], C, Cl, ca, scale=list(np.linspace(2.0, 0.1, 1000)), **kw)
    assert np.max(relerr(tot, al['layers_scaled_list'][:, lay])) < 1e-12


@pytest.mark.parametrize('planet', ['jupiter', 'neptune'])
def test_compute_ds_matches_reference(planet):
    g = golden('rays.npz')
    a = golden('atm_{}.npz'.format(planet))
    LP = keymap(a['LP_keys'])
    req, nr = a['property'][LP['R']], a['property'][LP['N']]
    for b, dsr, ns in zip(g['b'], g['ds_' + planet], g['nseg_' + planet]):
        out = ro.compute_ds(req, nr, b, float(a['Req']), float(a['Rpol']), a['orientation'], str(a['gtype']), 'shape')
        if ns == 0:
            assert out['ds'] is None
            continue
        ref = dsr[:ns]
        assert len(out['ds']) == ns
        assert np.array_equal(np.isnan(out['ds']), np.isnan(ref))          # NaN from the tangent depth on
        assert np.nanmax(relerr(out['ds'], ref)) < 1e-12
        assert list(out['layer4ds']) == list(range(ns))


def test_compute_ds_secant_limb():
    g = golden('rays.npz')
    a = golden('atm_jupiter.npz')
    LP = keymap(a['LP_keys'])
    for b, dsr, ns in zip(g['b'][:6], g['ds_jupiter_sec'], g['nseg_jupiter_sec']):
        out = ro.compute_ds(a['property'][LP['R']], a['property'][LP['N']], b, float(a['Req']), float(a['Rpol']),
                            a['orientation'], 'ellipse', 'sec')
        assert np.nanmax(relerr(out['ds'], dsr[:ns])) < 1e-12


def _tb_oracle(atm, slab_FL, blist, disc=False):
    C, LP = keymap(atm['C_keys']), keymap(atm['LP_keys'])
    T = atm['gas'][C['T']]
    out = []
    for b in blist:
        ray = ro.compute_ds(atm['property'][LP['R']], atm['property'][LP['N']], b, float(atm['Req']),
                            float(atm['Rpol']), atm['orientation'], str(atm['gtype']), str(atm['limb']))
        out.append(rto.integrate_ray(ray['ds'], ray['layer4ds'], slab_FL, T, disc_average=disc))
    return np.array(out)


def test_reference_known_answer_table():
    """scripts/benchmark.py: Jupiter benchmark config, 6 freqs x 4 emission angles (A+B end to end)."""
    a = golden('atm_jupiter_benchmark.npz')
    tb = golden('tb.npz')
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    lay = ao.get_layers(tb['bench_freqs'], a['gas'], a['cloud'], C, Cl, dict(formalisms_of(a)),
                        other_dicts={'h2': {'h2state': 'e'}}, truncate_strength=TRUNC)
    Tb = _tb_oracle(a, lay, tb['bench_b'])
    assert np.max(np.abs(Tb - tb['bench_tb'])) < 1e-6          # vs the reference run here (float64)
    assert np.max(np.abs(Tb - T_RB)) < 2e-4                    # vs the table printed in the reference (float32)
    assert np.max(np.abs(tb['bench_tb_f32'] - T_RB)) < 1e-4


def test_disc_and_point_tb_with_reference_alpha():
    a = golden('atm_jupiter.npz')
    tb = golden('tb.npz')
    # C1: disc-averaged (E2 weighting) using the reference's own alpha.layers -> isolates path B
    Tb = _tb_oracle(a, tb['c1_alpha'], [[0.0, 0.0]], disc=True)
    assert np.max(np.abs(Tb - tb['c1_tb'])) < 1e-9
    C, LP = keymap(a['C_keys']), keymap(a['LP_keys'])
    ray = ro.compute_ds(a['property'][LP['R']], a['property'][LP['N']], [0.0, 0.0], float(a['Req']), float(a['Rpol']),
                        a['orientation'], 'ellipse', 'shape')
    _, prof = rto.integrate_ray(ray['ds'], ray['layer4ds'], tb['c1_alpha'], a['gas'][C['T']], disc_average=True,
                                return_profiles=True)
    for k in ['tau', 'W', 'Tb_lyr']:
        assert np.max(relerr(prof[k], tb['c1_' + k])) < 1e-12
    assert np.max(relerr(prof['integrated_W'], tb['c1_integrated_W'])) < 1e-12


def test_uranus_end_to_end():
    """A planet outside the BASELINE configs (Uranus, 10x solar wet, own cloud file): oracle alpha on a layer subset,
    then disc-averaged and point Tb with the reference's alpha, against the unmodified reference."""
    a = golden('atm_uranus.npz')
    u = golden('uranus.npz')
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    lay = list(range(0, 1000, 9)) + [998, 999]
    tot, _, ordered = ao.get_layers(u['freqs'], a['gas'], a['cloud'], C, Cl, dict(formalisms_of(a)), return_per_constituent=True,
                                    other_dicts={'h2': {'h2state': str(a['h2state'])}}, truncate_strength=TRUNC, layers=lay)
    assert ordered == [str(x) for x in u['ordered_constituents']]
    assert np.max(relerr(tot, u['alpha'][:, lay])) < 1e-12
    assert np.max(np.abs(_tb_oracle(a, u['alpha'], [[0.0, 0.0]], disc=True) - u['disc_tb'])) < 1e-9
    assert np.max(np.abs(_tb_oracle(a, u['alpha'], u['pts']) - u['pt_tb'])) < 1e-9


def test_image_subset_c4():
    """C4 subset: on-disc pixels, the NaN limb ring and off-disc pixels (= T_cmb)."""
    a = golden('atm_jupiter.npz')
    im = golden('image_c4.npz')
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    sel = list(range(0, 96, 8)) + list(range(96, 136, 4)) + [136, 140]
    lay = ao.get_layers(im['freqs'][::8], a['gas'], a['cloud'], C, Cl, dict(formalisms_of(a)),
                        other_dicts={'h2': {'h2state': 'e'}}, truncate_strength=TRUNC)
    grid = im['grid']
    assert len(grid) == 601 and np.allclose(grid, ro.image_grid(0.005))
    bl = [[grid[ix], grid[iy]] for iy, ix in im['pick_iy_ix'][sel]]
    Tb = _tb_oracle(a, lay, bl)
    ref = im['tb'][sel][:, ::8]
    assert np.array_equal(np.isnan(Tb), np.isnan(ref))
    assert np.nanmax(np.abs(Tb - ref)) < 1e-6
    assert np.all(im['tb'][136:] == 2.725)


@pytest.mark.parametrize('name', ['nh3_kd', 'nh3_sjsd', 'nh3_bg'])
def test_remaining_nh3_formalisms(name):
    g = golden('plugins_nh3_extra.npz')
    C = keymap(g['C_keys'])
    for units in ['invcm', 'dBperkm']:
        for p, r in zip(g['points'], g['{}__{}'.format(name, units)]):
            a = ao.FORMALISMS[name](g['freqs'], p[C['T']], p[C['P']], p, C, {}, units=units)
            assert np.array_equal(np.isnan(a), np.isnan(r))
            assert np.nanmax(relerr(a, r)) < 1e-12


@pytest.mark.parametrize('name', ['nh3_hs', 'nh3_dbs', 'nh3_kd', 'nh3_dbs_sjs'])
def test_nh3_full_catalog(name):
    """SURVEY 8d / BASELINE config C5 "full NH3 catalog": the reference's plugins on an ammonia.npz rebuilt from the
    untrimmed line lists (415 + 1301 + 4198 = 5914 lines) against the oracle with LineCatalog(full_nh3=True)."""
    g = golden('plugins_nh3_full.npz')
    assert list(g['nlines']) == [415, 1301, 4198]
    cat = ao.LineCatalog(full_nh3=True)
    assert cat.get('nh3_rot').shape == (6, 1301) and cat.get('nh3_v2').shape == (3, 4198)
    C = keymap(g['C_keys'])
    trimmed = golden('plugins_trunc.npz')
    for units in ['invcm', 'dBperkm']:
        for p, r in zip(g['points'], g['{}__{}'.format(name, units)]):
            a = ao.FORMALISMS[name](g['freqs'], p[C['T']], p[C['P']], p, C, {}, units=units, cat=cat)
            assert np.array_equal(np.isnan(a), np.isnan(r))
            assert np.nanmax(relerr(a, r)) < 1e-12
    if name in ('nh3_hs', 'nh3_dbs'):          # the extra lines matter: not the trimmed-catalog answer
        assert np.nanmax(relerr(g[name + '__invcm'], trimmed[name + '__invcm'])) > 1e-6


@pytest.mark.parametrize('state', ['e', 'n'])
def test_h2_orton_oracle(state):
    """Orton's H2 CIA tables; 37 points over the three temperature branches."""
    g = golden('plugins_h2_orton.npz')
    C = keymap(g['C_keys'])
    for units in ['invcm', 'dBperkm']:
        for p, r in zip(g['points'], g['h2_orton_{}__{}'.format(state, units)]):
            a = ao.h2_orton(g['freqs'], p[C['T']], p[C['P']], p, C, {'h2state': state}, units=units)
            assert np.max(relerr(a, r)) < 1e-13


def test_h2_orton_host_table():
    """The product's host-side table preparation (frequency quadratic + not-a-knot spline coefficients)
    against the oracle's restatement of readInputFiles and scipy's interp1d(kind='cubic')."""
    from scipy.interpolate import interp1d
    from radiobear_b200 import catalogs
    f = np.array([0.5, 0.6, 1.0, 22.0, 100.0, 500.0, 1000.0])
    Ttab, h2vab = ao.orton_tables(list(f))
    for st, tabs in (('e', (0, 2, 4)), ('n', (1, 3, 5))):
        tab = catalogs.orton_table(f, st)
        assert tab.shape == (121, len(f)) and np.array_equal(tab[:10, 0], Ttab)
        for t, ii in enumerate(tabs):
            base = 10 + 37 * t
            assert np.max(relerr(tab[base:base + 10].T, h2vab[ii])) < 1e-14
            for T in (40.0, 41.0, 77.7, 150.0, 250.0, 399.9):
                k = int(np.clip(np.searchsorted(Ttab, T, side='right') - 1, 0, 8))
                dt = T - Ttab[k]
                c = tab[base + 10 + 3 * k:base + 13 + 3 * k]
                got = tab[base + k] + dt * (c[0] + dt * (c[1] + dt * c[2]))
                ref = np.array([interp1d(Ttab, h2vab[ii, j], kind='cubic')(T) for j in range(len(f))])
                assert np.max(relerr(got, ref)) < 1e-12
    with pytest.raises(ValueError):
        catalogs.orton_table([80000.0], 'e')


# ---- round 2: the SURVEY 8d sized fixtures (make_golden.py sections ring, c3_full, image_full, c5_saturn) ----------
def _geom_args(a):
    LP = keymap(a['LP_keys'])
    return (a['property'][LP['R']], a['property'][LP['N']]), (float(a['Req']), float(a['Rpol']), a['orientation'],
                                                             str(a['gtype']), str(a['limb']))


def test_ring_quadrant_classification_subset():
    """Every 19th pixel of the limb-ring quadrant (hit / miss, segment count, NaN in the segments Brightness.single
    uses, first NaN layer): the oracle reproduces the reference's classification exactly.  (The GPU test covers all
    5177 pixels.)"""
    ring = golden('ring_quadrant.npz')
    a = golden('atm_jupiter.npz')
    (req, nr), rest = _geom_args(a)
    grid = ring['grid']
    sel = np.arange(0, len(ring['nseg']), 19)
    assert (ring['nseg'][sel] < 0).any() and ring['used_nan'][sel].any() and (ring['used_nan'][sel] == 0).any()
    for k in sel:
        iy, ix = ring['iy_ix'][k]
        out = ro.compute_ds(req, nr, [grid[ix], grid[iy]], *rest)
        if ring['nseg'][k] < 0:
            assert out['ds'] is None
            continue
        ds = np.asarray(out['ds'])
        assert len(ds) == ring['nseg'][k]
        bad = np.nonzero(np.isnan(ds))[0]
        assert (int(bad[0]) if len(bad) else -1) == ring['first_nan'][k]
        assert abs(np.nansum(ds) / ring['nansum_ds'][k] - 1.0) < 1e-12


def test_c3_full_profile_rays():
    """Config C3 in full: the 100 rays of b = '0.0:1.0:0.01<0'.  NaN rays at the same b; Tb of every 7th ray at
    every 7th frequency within 1e-6 K."""
    from radiobear_b200 import set_utils
    a = golden('atm_jupiter.npz')
    c3 = golden('c3_full.npz')
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    rv = set_utils.set_b('0.0:1.0:0.01<0', [1, 1], Rpol=float(a['Rpol']), Req=float(a['Req']))
    assert np.array_equal(np.array(rv.b), c3['b'])
    sel = sorted(set(list(range(0, 100, 7)) + [96, 97, 98, 99]))
    lay = ao.get_layers(c3['freqs'][::7], a['gas'], a['cloud'], C, Cl, dict(formalisms_of(a)),
                        other_dicts={'h2': {'h2state': 'e'}}, truncate_strength=TRUNC)
    Tb = _tb_oracle(a, lay, [c3['b'][i] for i in sel])
    ref = c3['tb'][sel][:, ::7]
    assert np.isnan(ref).any()
    assert np.array_equal(np.isnan(Tb), np.isnan(ref))
    assert np.nanmax(np.abs(Tb - ref)) < 1e-6


def test_image_full_golden_is_what_8d_asks_for():
    """The C4 fixture holds >= 256 finite on-disc pixels, NaN-ring pixels and off-disc pixels; a slice of it through
    the oracle."""
    a = golden('atm_jupiter.npz')
    im = golden('image_c4_full.npz')
    ref = im['tb']
    nan = np.isnan(ref).any(axis=1)
    sky = (ref == 2.725).all(axis=1)
    assert (~nan & ~sky).sum() >= 256 and nan.sum() >= 8 and sky.sum() >= 16 and ref.shape[1] == 64
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    sel = list(range(0, 288, 24)) + list(range(288, 352, 8)) + [352, 360]
    lay = ao.get_layers(im['freqs'][::16], a['gas'], a['cloud'], C, Cl, dict(formalisms_of(a)),
                        other_dicts={'h2': {'h2state': 'e'}}, truncate_strength=TRUNC)
    grid = im['grid']
    Tb = _tb_oracle(a, lay, [[grid[ix], grid[iy]] for iy, ix in im['pick_iy_ix'][sel]])
    assert np.array_equal(np.isnan(Tb), np.isnan(ref[sel][:, ::16]))
    assert np.nanmax(np.abs(Tb - ref[sel][:, ::16])) < 1e-6


def test_c5_saturn_oracle_rows():
    """Config C5's concrete input (Saturn regridType=4096 x 4096 freqs x nh3_dbs_sjs): oracle against the reference's
    plugin on layers from all three branches of the pressure blend."""
    g = golden('c5_saturn.npz')
    C = keymap(g['C_keys'])
    P = g['gas'][C['P']][g['layers']]
    pick = [0, 20, 40, int(np.argmax(P >= 400)) - 1, int(np.argmax(P >= 400)), int(np.argmax(P >= 400)) + 5,
            int(np.argmax(P > 2000)) - 1, int(np.argmax(P > 2000)), len(P) - 1]
    for k in pick:
        col = g['gas'][:, g['layers'][k]]
        out = ao.FORMALISMS['nh3_dbs_sjs'](g['freqs'], col[C['T']], col[C['P']], col, C, {}, units='invcm')
        assert np.max(relerr(out, g['alpha'][k])) < 1e-12


def test_gravity_geoid_oracle_vs_reference():
    """gtype='gravity' (shape.py:141-221): the oracle's scalar restatement reproduces the reference's calcShape (27
    samples up to 0.3 deg on three layers) and two complete near-equatorial rays (the only ones the reference can
    afford: it builds a scipy polynomial object per Legendre evaluation); the all-layers table the GPU tests use agrees
    with the scalar march."""
    from conftest import golden, keymap
    from oracle import ray_oracle as ro
    a, g = golden('atm_jupiter.npz'), golden('gravity.npz')
    LP = keymap(a['LP_keys'])
    req, GM, nr = a['property'][LP['R']], a['property'][LP['GM']], a['property'][LP['N']]
    model = dict(GM=GM, Jn=g['Jn'], RJ=float(g['RJ']), omega_m=float(g['omega_m']), vwlat=g['vwlat'], vwdat=g['vwdat'])
    S = ro.Geoid(req, GM, g['Jn'], float(g['RJ']), float(g['omega_m']), g['vwlat'], g['vwdat'])
    T = ro.GeoidTable(req, GM, g['Jn'], float(g['RJ']), float(g['omega_m']), g['vwlat'], g['vwdat'], max_abs_lat=1.0)
    for row in g['shape_rows']:
        l, lat = int(row[0]), row[1]
        for G in (S, T):
            rm = G.calc(req[l], lat, float(g['shape_dlng']))
            assert abs(rm / row[2] - 1.0) < 1e-13 and abs(G.gamma - row[3]) < 1e-15
            assert np.max(np.abs(G.r - row[4:7])) < 1e-8 and np.max(np.abs(G.n - row[7:10])) < 1e-14
        nsp, k = T.kindex(lat)
        assert k == len(np.arange(0.0, lat + nsp * 0.01, nsp * 0.01)) - 1
    b, n = list(g['b'][0]), int(g['nseg'][0])
    with np.errstate(invalid='ignore'):
        ray = ro.compute_ds(req, nr, b, float(a['Req']), float(a['Rpol']), a['orientation'], 'gravity', 'shape', gravity=model)
        fast = ro.compute_ds(req, nr, b, float(a['Req']), float(a['Rpol']), a['orientation'], 'gravity', 'shape',
                             gravity=dict(table=T))
    assert len(ray['ds']) == n == len(fast['ds'])
    assert np.max(np.abs(ray['ds'] / g['ds'][0, :n] - 1.0)) < 1e-12 and np.max(np.abs(ray['r4ds'] / g['r4ds'][0, :n] - 1.0)) < 1e-13
    assert np.max(np.abs(fast['ds'] / g['ds'][0, :n] - 1.0)) < 1e-8


def test_doppler_branch_oracle_vs_reference():
    """Brightness.single with config Doppler (brightness.py:80-96), the reference's branch run with its renamed
    `alpha.get_alpha` call restored (tests/golden/make_golden.py section `doppler`; omega_m x 100 so that the branch
    moves Tb by 0.02-0.3 K).  The oracle walks the same steps with the oracle absorption at the shifted frequencies;
    one ray completely (998 steps x 2 absorption evaluations), and the no-shift cases (central meridian, disc average,
    sky) against the plain run."""
    a = golden('atm_jupiter.npz')
    d = golden('doppler.npz')
    C, Cl, LP = keymap(a['C_keys']), keymap(a['Cl_keys']), keymap(a['LP_keys'])
    freqs = d['freqs']
    kw = dict(other_dicts={'h2': {'h2state': str(a['h2state'])}}, truncate_strength=TRUNC)
    fa = dict(formalisms_of(a))

    def alpha_at(layer, fv):
        return ao.get_layers(fv, a['gas'], a['cloud'], C, Cl, fa, layers=[layer], **kw)[:, 0]
    assert np.max(np.abs(d['tb_doppler'] - d['tb_plain'])[:2]) > 0.1 and float(d['omega_factor']) == 100.0
    assert np.array_equal(d['tb_doppler'][2], d['tb_plain'][2]) or np.max(np.abs(d['tb_doppler'][2] - d['tb_plain'][2])) < 1e-9
    assert (d['tb_doppler'][3] == 2.725).all()
    k = 1                                                      # b = (-0.8, 0.3)
    ray = ro.compute_ds(a['property'][LP['R']], a['property'][LP['N']], d['b'][k], float(a['Req']), float(a['Rpol']),
                        a['orientation'], 'ellipse', 'shape')
    dop = d['doppler%d' % k]
    assert len(dop) == len(ray['ds'])
    Tb = rto.integrate_ray_doppler(ray['ds'], ray['layer4ds'], dop, freqs, alpha_at, a['gas'][C['T']])
    assert np.max(np.abs(Tb - d['tb_doppler'][k])) < 1e-6
    # doppler == 1 everywhere reduces to the plain loop
    slab = ao.get_layers(freqs, a['gas'], a['cloud'], C, Cl, fa, **kw)
    plain = rto.integrate_ray(ray['ds'], ray['layer4ds'], slab, a['gas'][C['T']])
    same = rto.integrate_ray_doppler(ray['ds'], ray['layer4ds'], np.ones_like(dop), freqs, lambda l, fv: slab[:, l], a['gas'][C['T']])
    assert np.array_equal(plain, same)
    assert np.max(np.abs(plain - d['tb_plain'][k])) < 1e-6
