"""GPU: absorption parity -- CUDA alpha_lines kernel (through the C ABI) vs the reference's golden
vectors and vs the oracle on seeded inputs.  Bar (north_star): alpha within 1e-6 relative; the
kernel actually sits near 1e-13, asserted at 1e-9 so regressions show."""
import numpy as np
import pytest

from conftest import golden, keymap, relerr, formalisms_of, TRUNC

pytestmark = pytest.mark.gpu
TOL = 1e-6        # north_star tolerance
TIGHT = 1e-9      # what the FP64 kernel is expected to hold

FAMILY = {'nh3_hs': 'nh3', 'nh3_dbs': 'nh3', 'nh3_sjs': 'nh3', 'nh3_hs_sjs': 'nh3', 'nh3_dbs_sjs': 'nh3',
          'h2s_ddb': 'h2s', 'ph3_jh': 'ph3', 'h2o_bk': 'h2o', 'co_ddb': 'co', 'h2_jj_ddb': 'h2', 'h2_jj': 'h2'}


@pytest.fixture(scope='module')
def eng():
    from radiobear_b200 import engine
    return engine


@pytest.mark.parametrize('name', sorted(FAMILY))
@pytest.mark.parametrize('units', ['invcm', 'dBperkm'])
def test_plugin_golden(eng, name, units):
    g = golden('plugins_trunc.npz')
    C = keymap(g['C_keys'])
    gas = np.ascontiguousarray(g['points'].T)
    c = FAMILY[name]
    out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[(c, name)],
                           other_dicts={c: {'h2state': 'e', 'coshape': 'voigt'}}, units=units, truncate_strength=TRUNC)
    ref = g['{}__{}'.format(name, units)]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.nanmax(relerr(out, ref)) < TIGHT


def test_plugin_shims_keep_reference_signature():
    """constituents/<gas>/<formalism>.alpha(freq, T, P, X, P_dict, other_dict, **kwargs) -- called the way
    Alpha.get_alpha_from_calc does (alpha.py:210-213), one layer at a time."""
    import importlib
    g = golden('plugins_trunc.npz')
    C = keymap(g['C_keys'])
    for name, c in FAMILY.items():
        mod = importlib.import_module('radiobear_b200.constituents.{}.{}'.format(c, name))
        for i in (3, 10, 17, 22):
            p = g['points'][i]
            a = mod.alpha(list(g['freqs']), p[C['T']], p[C['P']], p, C, {'h2state': 'e', 'coshape': 'voigt'},
                          truncate_freq=None, truncate_strength=TRUNC.get(c), units='invcm', path='ignored',
                          verbose=False)
            assert a.shape == (len(g['freqs']),)
            assert np.nanmax(relerr(a, g[name + '__invcm'][i])) < TIGHT
        # default units are dB/km (parameters.py:4-8)
        a = mod.alpha(list(g['freqs']), p[C['T']], p[C['P']], p, C, {'h2state': 'e'}, truncate_strength=TRUNC.get(c))
        assert np.nanmax(relerr(a, g[name + '__dBperkm'][22])) < TIGHT


def test_nh3_plugin_reorders_lo_then_hi():
    """nh3_hs.alpha concatenates the f<=30 block before the f>30 block (nh3_hs.py:76-88)."""
    from radiobear_b200.constituents.nh3 import nh3_hs
    g = golden('plugins_trunc.npz')
    C = keymap(g['C_keys'])
    p = g['points'][5]
    f = np.array([40.0, 10.0, 35.0, 20.0])
    a = nh3_hs.alpha(f, p[C['T']], p[C['P']], p, C, {}, units='invcm')
    b = nh3_hs.alpha(np.array([10.0, 20.0, 40.0, 35.0]), p[C['T']], p[C['P']], p, C, {}, units='invcm')
    assert np.array_equal(a, b)


def test_option_variants_and_clouds(eng):
    g = golden('plugins_trunc.npz')
    C, Cl = keymap(g['C_keys']), keymap(g['Cl_keys'])
    gas = np.ascontiguousarray(g['points'].T)
    out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('h2', 'h2_jj_ddb')],
                           other_dicts={'h2': {'h2state': 'n'}})
    assert np.max(relerr(out, g['h2_jj_ddb_n__invcm'])) < TIGHT
    out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('co', 'co_ddb')],
                           other_dicts={'co': {'coshape': 'vvw'}})
    assert np.max(relerr(out, g['co_ddb_vvw__invcm'])) < TIGHT
    od = {'water_p': 1e-4, 'ice_p': 1e-4, 'nh4sh_p': 1e-4, 'nh3ice_p': 1e-4, 'h2sice_p': 1e-4, 'ch4_p': 1e-4}
    cl = np.ascontiguousarray(g['cloud_points'].T)
    for units in ['invcm', 'dBperkm']:
        out = eng.alpha_layers(g['freqs'], g['cloud_T'], np.ones(len(g['cloud_T'])), np.zeros((1, cl.shape[1])), {},
                               cloud=cl, cloud_dict=Cl, formalisms=[('clouds', 'clouds_idp')], other_dicts={'clouds': od},
                               units=units)
        assert np.max(relerr(out, g['clouds_idp__' + units])) < TIGHT
    with pytest.raises(ValueError):
        eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('h2', 'h2_jj_ddb')],
                         other_dicts={'h2': {'h2state': 'x'}})
    with pytest.raises(NotImplementedError):
        eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('h2o', 'h2o_ddb')])


def test_no_truncation(eng):
    g = golden('plugins_notrunc.npz')
    C = keymap(g['C_keys'])
    gas = np.ascontiguousarray(g['points'].T)
    for name, c in [('h2s_ddb', 'h2s'), ('ph3_jh', 'ph3')]:
        out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[(c, name)])
        assert np.max(relerr(out, g[name + '__invcm'])) < TIGHT
        out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[(c, name)],
                               truncate_strength={c: 1e-22})
        assert np.max(relerr(out, golden('plugins_trunc.npz')[name + '__invcm'])) < TIGHT


def test_jupiter_full_cube_and_scaling(eng):
    """Alpha.get_layers on all 1000 Jupiter layers: total, per-constituent cube, dict / list scales."""
    a = golden('atm_jupiter.npz')
    al = golden('alpha_jupiter.npz')
    C = keymap(a['C_keys'])
    kw = dict(formalisms=formalisms_of(a), other_dicts={'h2': {'h2state': 'e'}}, truncate_strength=TRUNC)
    tot, cube = eng.alpha_layers(al['freqs'], a['gas'][C['T']], a['gas'][C['P']], a['gas'], C, want_cube=True, **kw)
    assert np.max(relerr(tot.T, al['layers'])) < TIGHT
    assert np.nanmax(relerr(cube, al['cube'])) < TIGHT
    sc = {'nh3': list(np.linspace(0.5, 1.5, 1000)), 'h2o': [2.0] * 1000}
    tot = eng.alpha_layers(al['freqs'], a['gas'][C['T']], a['gas'][C['P']], a['gas'], C, scale=sc, **kw)
    assert np.max(relerr(tot.T, al['layers_scaled_dict'])) < TIGHT
    tot = eng.alpha_layers(al['freqs'], a['gas'][C['T']], a['gas'][C['P']], a['gas'], C,
                           scale=list(np.linspace(2.0, 0.1, 1000)), **kw)
    assert np.max(relerr(tot.T, al['layers_scaled_list'])) < TIGHT


def test_neptune_c2(eng):
    """Config C2: Neptune, 1500 layers, 200 log-spaced freqs, 6 constituents (co contributes exactly 0)."""
    a = golden('atm_neptune.npz')
    n = golden('neptune_c2.npz')
    C = keymap(a['C_keys'])
    forms = formalisms_of(a)
    assert [c for c, _ in forms] == [str(x) for x in n['ordered_constituents']]
    tot, cube = eng.alpha_layers(n['freqs'], a['gas'][C['T']], a['gas'][C['P']], a['gas'], C, formalisms=forms,
                                 other_dicts={'h2': {'h2state': 'e'}, 'co': {'coshape': 'voigt'}},
                                 truncate_strength=TRUNC, want_cube=True)
    assert np.max(relerr(tot.T[:, ::8], n['alpha_every8'])) < TIGHT
    assert np.all(cube[:, :, [c for c, _ in forms].index('co')] == 0.0)


@pytest.mark.parametrize('F', [1, 7, 33, 64, 257, 513, 1100])
def test_ragged_frequency_counts_vs_oracle(eng, F):
    """Every tiling path (1/2 freqs per thread, 1..8 line slices, partial warps) against the oracle on
    seeded synthetic layers of the C5 kind (T, P, X drawn log-uniformly)."""
    from oracle import alpha_oracle as ao
    rng = np.random.default_rng(F)
    L = 9
    C = {'Z': 0, 'T': 1, 'P': 2, 'H2': 3, 'HE': 4, 'CH4': 5, 'NH3': 6, 'H2O': 7, 'H2S': 8, 'PH3': 9, 'CO': 10}
    gas = np.zeros((11, L))
    gas[C['T']] = rng.uniform(80, 1800, L)
    gas[C['P']] = 10**rng.uniform(-2, np.log10(5e3), L)
    gas[C['P']][:3] = [399.9, 400.0, 2000.5]
    gas[C['H2']], gas[C['HE']], gas[C['CH4']] = 0.86, 0.135, 2e-3
    for k in ('NH3', 'H2O', 'H2S', 'PH3', 'CO'):
        gas[C[k]] = 10**rng.uniform(-7, -3, L)
    freqs = np.sort(rng.uniform(0.5, 300.0, F))
    if F >= 7:
        freqs[:7] = [25.5, 26.0, 27.0, 30.0, 30.5, 34.0, 36.0]
        freqs = np.sort(freqs)
    ca = {'nh3': 'nh3_dbs_sjs', 'h2s': 'h2s_ddb', 'ph3': 'ph3_jh', 'h2o': 'h2o_bk', 'h2': 'h2_jj_ddb', 'co': 'co_ddb'}
    od = {'h2': {'h2state': 'e'}, 'co': {'coshape': 'voigt'}}
    ref, rcube, ordered = ao.get_layers(freqs, gas, np.zeros((1, L)), C, {}, ca, other_dicts=od, truncate_strength=TRUNC,
                                        return_per_constituent=True)
    tot, cube = eng.alpha_layers(freqs, gas[C['T']], gas[C['P']], gas, C, formalisms=[(c, ca[c]) for c in ordered],
                                 other_dicts=od, truncate_strength=TRUNC, want_cube=True)
    assert np.array_equal(np.isnan(cube), np.isnan(rcube))
    assert np.nanmax(relerr(cube, rcube)) < TIGHT
    assert np.nanmax(relerr(tot.T, ref)) < TIGHT


def test_alpha_object_api(tmp_path):
    """Alpha.get_layers / save_alpha / get_alpha='memory' + scale (the MCMC reuse path, scripts/demo_batch.py)."""
    import os
    from conftest import GOLDEN
    from radiobear_b200.atmosphere import Atmosphere
    from radiobear_b200.alpha import Alpha
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, 'atm_jupiter.npz'), 'jupiter')
    al = golden('alpha_jupiter.npz')
    atm.config.scratch_directory = str(tmp_path)
    A = Alpha(config=atm.config, verbose=False)
    assert A.ordered_constituents == [str(x) for x in al['ordered_constituents']]
    A.get_layers(list(al['freqs']), atm, save_alpha='memory')
    assert A.layers.shape == (8, 1000) and np.max(relerr(A.layers, al['layers'])) < TIGHT
    assert np.nanmax(relerr(A.memory.alpha_data, al['cube'])) < TIGHT
    assert A.layers[3][500] == A.slab[500][3]
    A.get_layers(list(al['freqs']), atm, scale={'nh3': list(np.linspace(0.5, 1.5, 1000)), 'h2o': [2.0] * 1000},
                 get_alpha='memory')
    assert np.max(relerr(A.layers, al['layers_scaled_dict'])) < TIGHT
    A.get_layers(list(al['freqs']), atm, save_alpha='file')
    assert os.path.exists(A.alphafile)
    A.get_layers(list(al['freqs']), atm, scale=list(np.linspace(2.0, 0.1, 1000)), get_alpha='file')
    assert np.max(relerr(A.layers, al['layers_scaled_list'])) < TIGHT
    one = A.get_single_layer(list(al['freqs']), 500, atm)
    assert np.max(relerr(one, al['layers'][:, 500])) < TIGHT
    with pytest.raises(ValueError):
        A.get_layers(list(al['freqs']), atm, scale={'bogus': [1.0] * 1000})
    # one layer the way alpha.py:218-233 does it: per-constituent absorption, then the scale-sum of that layer
    C = atm.config.C
    freqs = list(al['freqs'])
    absorb = A.get_alpha_from_calc(freqs, atm.gas[C['T']][500], atm.gas[C['P']][500], atm.gas[:, 500], C,
                                   atm.cloud[:, 500] if np.size(atm.cloud) else None, atm.config.Cl)
    assert absorb.shape == (8, len(A.ordered_constituents)) and np.nanmax(relerr(absorb, al['cube'][500])) < TIGHT
    A.freqs = freqs
    assert np.max(relerr(A.total_layer_alpha(absorb, 1.0), al['layers'][:, 500])) < TIGHT
    assert np.max(relerr(A.total_layer_alpha(absorb, 2.5), 2.5 * absorb.sum(axis=1))) < 1e-14
    j_nh3, j_h2o = A.ordered_constituents.index('nh3'), A.ordered_constituents.index('h2o')
    w = np.ones(absorb.shape[1])
    w[j_nh3], w[j_h2o] = 0.5, 2.0
    lscale = {'nh3': 0.5, 'h2o': 2.0, 'not_a_constituent': 7.0}
    assert np.max(relerr(A.total_layer_alpha(absorb, lscale), (absorb * w).sum(axis=1))) < 1e-14
    assert np.max(relerr(A.get_single_layer(freqs, 500, atm, lscale), (absorb * w).sum(axis=1))) < TIGHT
    A._save_alpha_memfil, A.tosave = True, []                       # the reference's per-layer cache list
    A.total_layer_alpha(absorb, lscale)
    assert len(A.tosave) == 1 and np.max(relerr(A.tosave[0], absorb * w)) < 1e-15


@pytest.mark.parametrize('name', ['nh3_kd', 'nh3_sjsd', 'nh3_bg'])
@pytest.mark.parametrize('units', ['invcm', 'dBperkm'])
def test_remaining_nh3_formalisms(eng, name, units):
    """nh3_kd (pressure-switched constants), nh3_sjsd (10..100 bar blend), nh3_bg."""
    import importlib
    g = golden('plugins_nh3_extra.npz')
    C = keymap(g['C_keys'])
    gas = np.ascontiguousarray(g['points'].T)
    out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('nh3', name)], units=units)
    ref = g['{}__{}'.format(name, units)]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    assert np.nanmax(relerr(out, ref)) < TIGHT
    mod = importlib.import_module('radiobear_b200.constituents.nh3.' + name)
    p = g['points'][-4]
    a = mod.alpha(list(g['freqs']), p[C['T']], p[C['P']], p, C, {}, units=units)
    assert np.nanmax(relerr(a, ref[-4])) < TIGHT


@pytest.mark.parametrize('name', ['nh3_hs', 'nh3_dbs', 'nh3_kd', 'nh3_dbs_sjs'])
def test_nh3_full_catalog(eng, name):
    """SURVEY 8d / BASELINE config C5 "full NH3 catalog" (415 + 1301 + 4198 = 5914 lines, 181 KB of line tables in
    shared memory with the sjs blend): the kernel against the reference's plugins run on the untrimmed line lists;
    switching back restores the shipped catalog."""
    from radiobear_b200 import catalogs
    g = golden('plugins_nh3_full.npz')
    t = golden('plugins_trunc.npz')
    C = keymap(g['C_keys'])
    gas = np.ascontiguousarray(g['points'].T)
    catalogs.use_full_nh3_catalog(True)
    try:
        for units in ('invcm', 'dBperkm'):
            out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('nh3', name)], units=units)
            ref = g['{}__{}'.format(name, units)]
            assert np.array_equal(np.isnan(out), np.isnan(ref))
            assert np.nanmax(relerr(out, ref)) < TIGHT
    finally:
        catalogs.use_full_nh3_catalog(False)
    if name in t.files or (name + '__invcm') in t.files:
        out = eng.alpha_layers(t['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('nh3', name)], units='invcm')
        assert np.nanmax(relerr(out, t[name + '__invcm'])) < TIGHT


@pytest.mark.parametrize('state', ['e', 'n'])
@pytest.mark.parametrize('units', ['invcm', 'dBperkm'])
def test_h2_orton(eng, state, units):
    """h2_orton behind the common plugin signature (the reference module cannot be driven by
    Alpha); T^4 extrapolation below 40 K, cubic spline inside the table, scaled h2_jj above 400 K."""
    import importlib
    g = golden('plugins_h2_orton.npz')
    C = keymap(g['C_keys'])
    gas = np.ascontiguousarray(g['points'].T)
    out = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('h2', 'h2_orton')], units=units,
                           other_dicts={'h2': {'h2state': state}})
    ref = g['h2_orton_{}__{}'.format(state, units)]
    assert np.max(relerr(out, ref)) < TIGHT
    mod = importlib.import_module('radiobear_b200.constituents.h2.h2_orton')
    for i in (0, 20, 26, 30, 36):
        p = g['points'][i]
        a = mod.alpha(list(g['freqs']), p[C['T']], p[C['P']], p, C, {'h2state': state, 'h2newset': True}, units=units)
        assert np.max(relerr(a, ref[i])) < TIGHT
    # a different frequency vector re-prepares the table; a wrong state is refused
    out2 = eng.alpha_layers(g['freqs'][3:9], gas[C['T']], gas[C['P']], gas, C, formalisms=[('h2', 'h2_orton')], units=units,
                            other_dicts={'h2': {'h2state': state}})
    assert np.max(relerr(out2, ref[:, 3:9])) < TIGHT
    with pytest.raises(ValueError):
        eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('h2', 'h2_orton')],
                         other_dicts={'h2': {'h2state': 'x'}})


def test_resident_slab_and_retrieval_loop(tmp_path):
    """The absorption stays on the device between Alpha.get_layers and the integration (no copy to the host unless
    `.layers` is read), and the retrieval loop of scripts/demo_batch.py -- save_alpha='memory' once, then
    get_alpha='memory' with a new `scale` per iteration -- re-runs only the scale-sum on the resident cube.
    Values against the reference's own scaled layers (alpha.py:151-192) and against the host path."""
    import os
    from conftest import GOLDEN
    from radiobear_b200 import engine
    from radiobear_b200.atmosphere import Atmosphere
    from radiobear_b200.alpha import Alpha
    from radiobear_b200.brightness import Brightness
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, 'atm_jupiter.npz'), 'jupiter')
    al = golden('alpha_jupiter.npz')
    freqs = list(al['freqs'])
    atm.config.scratch_directory = str(tmp_path)
    A = Alpha(config=atm.config, verbose=False)
    A.get_layers(freqs, atm)
    assert A._slab is None and A._res is not None and A._res.valid() and A.has_layers()      # nothing copied back yet
    h = A.rt_slab()
    assert isinstance(h, engine.ResidentSlab) and h.shape == (1000, 8)
    bright = Brightness(config=atm.config, verbose=False)
    pts = np.array([[0.0, 0.0], [0.3, 0.2], [0.6, -0.4], [0.9, 0.1]])
    tb_res = bright.batch(pts, freqs, atm, A, atm.config.orientation)['Tb'].copy()
    assert A._slab is None                                            # the integration read the device copy
    assert np.max(relerr(A.layers, al['layers'])) < TIGHT             # first host access copies it back
    tb_host = bright.batch(pts, freqs, atm, A, atm.config.orientation)['Tb']
    assert np.array_equal(tb_res, tb_host)
    # a second Alpha takes the resident buffer: the first one is handed its slab before it is overwritten
    A.get_layers(freqs, atm)
    B = Alpha(config=atm.config, verbose=False)
    B.get_layers(freqs[:3], atm)
    assert A._slab is not None and not (A._res is not None and A._res.valid())
    assert np.max(relerr(A.layers, al['layers'])) < TIGHT and B.layers.shape == (3, 1000)
    # retrieval loop: the cube stays on the device
    A.get_layers(freqs, atm, save_alpha='memory')
    assert np.nanmax(relerr(A.memory.alpha_data, al['cube'])) < TIGHT
    assert A._dev_cube is not None and A._dev_cube[0].cube_valid()
    sc = {'nh3': list(np.linspace(0.5, 1.5, 1000)), 'h2o': [2.0] * 1000}
    A.get_layers(freqs, atm, scale=sc, get_alpha='memory')
    assert A._slab is None and A._res.valid() and A._res.cube_gen == A._dev_cube[0].cube_gen   # device scale-sum, no copy
    tb_scaled = bright.batch(pts, freqs, atm, A, atm.config.orientation)['Tb'].copy()
    assert np.max(relerr(A.layers, al['layers_scaled_dict'])) < TIGHT
    A.get_layers(freqs, atm, scale=list(np.linspace(2.0, 0.1, 1000)), get_alpha='memory')
    assert np.max(relerr(A.layers, al['layers_scaled_list'])) < TIGHT
    # the same iteration through the host path (another Alpha whose memory cache is a host array only)
    H = Alpha(config=atm.config, verbose=False)
    H.memory = A.memory
    H.get_layers(freqs, atm, scale=sc, get_alpha='memory')
    assert H._res is None and np.max(relerr(H.layers, al["layers_scaled_dict"])) < TIGHT
    tb_scaled_host = bright.batch(pts, freqs, atm, H, atm.config.orientation)['Tb']
    assert np.allclose(tb_scaled, tb_scaled_host, rtol=0, atol=1e-9) and np.abs(tb_scaled - tb_res).max() > 0.1
    # a replaced cache array is not mistaken for the device copy
    A.memory.alpha_data = np.array(A.memory.alpha_data) * 2.0
    A.get_layers(freqs, atm, get_alpha='memory')
    assert A._res is None and np.max(relerr(A.layers, 2.0 * al['layers'])) < TIGHT
    # error handling of the C ABI: no resident slab of the right shape
    with pytest.raises((ValueError, RuntimeError)):
        engine.rt_batch(b=pts, alpha_slab=h, T=atm.gas[atm.config.C['T']], radius=atm.property[atm.config.LP['R']],
                        refr_index=atm.property[atm.config.LP['N']], Req=atm.config.Req, Rpol=atm.config.Rpol)


def test_per_layer_frequencies(eng):
    """rb_alpha_desc::freqs_per_layer: every layer at its own frequency list (what the Doppler branch of
    Brightness.single needs, brightness.py:83-92).  Row l of the result equals a one-layer call with that row's
    frequencies -- bit for bit, it is the same kernel -- for lists that straddle the 30 GHz band switch and the 26 / 34 GHz
    interpolation band differently from layer to layer, for both frequencies-per-lane paths; a constant matrix equals the
    shared-list call; the oracle agrees; h2_orton (table prepared per list) refuses."""
    from oracle import alpha_oracle as ao
    a = golden('atm_jupiter.npz')
    C, Cl = keymap(a['C_keys']), keymap(a['Cl_keys'])
    kw = dict(formalisms=formalisms_of(a), other_dicts={'h2': {'h2state': 'e'}}, truncate_strength=TRUNC)
    rng = np.random.default_rng(11)
    lay = np.arange(0, 1000, 37)
    gas, cloud = np.ascontiguousarray(a['gas'][:, lay]), np.ascontiguousarray(a['cloud'][:, lay])
    T, P = gas[C['T']], gas[C['P']]
    for F in (5, 70, 600):
        base = np.sort(rng.uniform(1.0, 100.0, F))
        fm = base[None, :] * (1.0 + 0.2 * rng.uniform(-1.0, 1.0, (len(lay), 1)))      # a different shift per layer
        got = eng.alpha_layers(fm, T, P, gas, C, cloud=cloud, cloud_dict=Cl, **kw)
        assert got.shape == (len(lay), F)
        for i in (0, 7, len(lay) - 1):
            one = eng.alpha_layers(fm[i], T[i:i + 1], P[i:i + 1], np.ascontiguousarray(gas[:, i:i + 1]), C,
                                   cloud=np.ascontiguousarray(cloud[:, i:i + 1]), cloud_dict=Cl, **kw)
            assert np.array_equal(got[i], one[0]), (F, i)
        same = eng.alpha_layers(np.tile(base, (len(lay), 1)), T, P, gas, C, cloud=cloud, cloud_dict=Cl, **kw)
        assert np.array_equal(same, eng.alpha_layers(base, T, P, gas, C, cloud=cloud, cloud_dict=Cl, **kw))
    i = 9
    ref = ao.get_layers(fm[i], gas, cloud, C, Cl, dict(formalisms_of(a)), other_dicts={'h2': {'h2state': 'e'}},
                        truncate_strength=TRUNC, layers=[i])[:, 0]
    assert np.max(relerr(got[i], ref)) < TIGHT
    with pytest.raises(NotImplementedError):
        eng.alpha_layers(fm, T, P, gas, C, formalisms=[('h2', 'h2_orton')], other_dicts={'h2': {'h2state': 'e'}})
    with pytest.raises(ValueError):
        eng.alpha_layers(fm[:3], T, P, gas, C, **kw)
