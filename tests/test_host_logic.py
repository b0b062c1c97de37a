"""CPU: host-side request parsing, configuration, scale handling, catalog truncation, the atmosphere
provider and the multi-GPU row partition (no kernels)."""
import os

import numpy as np
import pytest

from conftest import ROOT, golden, relerr
from radiobear_b200 import set_utils, utils, catalogs, engine, data_handling
from radiobear_b200.atmosphere import Atmosphere
from radiobear_b200 import parallel


def test_set_freq_forms():
    assert set_utils.set_freq(5)[0] == [5.0]
    assert set_utils.set_freq([1.0, 2.0])[0] == [1.0, 2.0]
    f, u = set_utils.set_freq('1:100:5')
    assert len(f) == 21 and f[-1] == 101.0                       # set_utils.py:138
    f, _ = set_utils.set_freq('1;100;200')                       # crashes in the reference
    assert len(f) == 200 and np.allclose(f, np.logspace(0, 2, 200))
    assert set_utils.set_freq('1,2,3')[0] == [1.0, 2.0, 3.0]
    assert set_utils.set_freq([1000.0], 'MHz')[0] == [1.0]
    with pytest.raises(ValueError):
        set_utils.set_freq({'a': 1})


def test_set_b_forms():
    r = set_utils.set_b('disc')
    assert r.b == ['disc'] and r.data_type == 'spectrum'
    r = set_utils.set_b([0.1, 0.2])
    assert r.b == [[0.1, 0.2]] and r.data_type == 'spectrum'
    r = set_utils.set_b([[0.0, 0.0]] * 6)
    assert r.data_type == 'profile'
    r = set_utils.set_b('0.0:1.0:0.01<0', Rpol=66854.0, Req=71492.0)
    assert len(r.b) == 100 and r.data_type == 'profile' and abs(r.b[-1][0] - 0.99) < 1e-12   # config C3
    r = set_utils.set_b('0.1,0.2<90', Rpol=66854.0, Req=71492.0)
    assert abs(r.b[0][1] - 0.1) < 1e-15 and abs(r.b[0][0]) < 1e-16
    r = set_utils.set_b(0.005)                                   # float image request crashes in the reference
    assert r.data_type == 'image' and r.imSize == [601, 601] and len(r.b) == 601 * 601
    assert list(r.b[0]) == [-1.5, -1.5] and r.b[1][1] == -1.5 and abs(r.b[1][0] + 1.495) < 1e-12   # rows of constant y, x fastest
    r2 = set_utils.set_b(0.005, block=[2, 4])
    assert r2.imSize[0] == 601 and len(r2.b) == 601 * r2.imSize[1]
    r = set_utils.set_b('stamp:0.1:-0.2,0.2,-0.1,0.1')
    assert r.data_type == 'image' and len(r.b) == 5 * 3


def test_image_grid_on_disc_count():
    grid = set_utils.image_grid(0.005)
    q = 66854.0 / 71492.0
    xx, yy = np.meshgrid(grid, grid)
    assert len(grid) == 601
    # 117 493 on-disc pixels of the 601 x 601 Jupiter grid (ellipse x^2 + (y/q)^2 < 1)
    assert abs(int(np.sum(xx**2 + (yy / q)**2 < 1.0)) - 117493) < 600


def test_units_and_helpers():
    assert utils.proc_unit('MHz') == 'GHz' and utils.convert_unit(2.0, 'AU') == 2.0 * 149597870.691
    assert utils.isanynum(3) and utils.isanynum('4.5') and not utils.isanynum(True) and not utils.isanynum([1])
    assert utils.b_type('DISC') == 'disc' and utils.b_type([[0, 0]] * 25) == 'image'
    assert utils.T_cmb == 2.725


def test_data_return_dtypes():
    d = data_handling.Data()
    d.set('Tb', [[1.0, 2.0]])
    assert d.Tb.dtype == np.float32                              # data_handling.py:46-47
    d.set('b', ['disc'])
    assert d.b == ['disc']
    d.set('bogus', 1)
    assert not hasattr(d, 'bogus')


def test_scale_matrix_rules():
    ordered = ['h2', 'h2o', 'nh3']
    assert engine.scale_matrix(False, ordered, 4) is None
    assert engine.scale_matrix(1.0, ordered, 4) is None
    m = engine.scale_matrix(2, ordered, 4)
    assert m.shape == (3, 4) and np.all(m == 2.0)
    m = engine.scale_matrix([1, 2, 3, 4], ordered, 4)
    assert np.all(m[1] == [1, 2, 3, 4])
    m = engine.scale_matrix({'nh3': [0.5] * 4}, ordered, 4)
    assert np.all(m[2] == 0.5) and np.all(m[:2] == 1.0)
    with pytest.raises(ValueError):
        engine.scale_matrix({'xx': [1] * 4}, ordered, 4)         # alpha.py:241-243
    with pytest.raises(ValueError):
        engine.scale_matrix({'nh3': [1] * 3}, ordered, 4)        # alpha.py:244-245
    with pytest.raises(ValueError):
        engine.scale_matrix([1, 2], ordered, 4)                  # alpha.py:251-252


def test_catalog_truncation():
    assert catalogs.table('h2s').shape == (4, 200) and catalogs.table('h2s', 1e-22).shape == (4, 121)
    assert catalogs.table('ph3').shape == (6, 320) and catalogs.table('ph3', 1e-22).shape == (6, 33)
    assert catalogs.table('h2o').shape == (9, 15) and catalogs.table('h2o', 0).shape == (9, 15)
    assert catalogs.table('h2o', truncate_freq=400.0).shape == (9, 5)
    assert catalogs.table('nh3_inv').shape == (4, 415) and catalogs.table('nh3_rot').shape == (6, 201)
    assert catalogs.table('nh3_v2').shape == (3, 198) and catalogs.table('nh3_sjs').shape == (4, 200)
    assert catalogs.table('co').shape == (3, 26)


def test_config_parsing(tmp_path):
    from radiobear_b200 import config as pcfg
    os.makedirs(tmp_path / 'Jupiter')
    cf = tmp_path / 'Jupiter' / 'config.par'
    cf.write_text('# comment\n'
                  'gasfile my.gas\nconstituents Z T P H2 HE CH4 NH3 H2O H2S SOLN OTHER PH3 CO CO13 HCN DZ\n'
                  'alpha nh3:nh3_dbs_sjs h2s:h2s_ddb cloud:none co:co_ddb\n'
                  'regridtype 1000\npmin 0.01\npmax 10000.0\ndistance 5.2 AU\norientation 10.0 -3.0  # deg\n'
                  'gtype sphere\nh2state n\n')
    c = pcfg.planetConfig('jupiter', configFile=str(cf))
    assert c.gasFile == ['my.gas'] and c.C['NH3'] == 6 and c.C['DZ'] == 15
    assert c.constituent_alpha['nh3'] == 'nh3_dbs_sjs' and c.constituent_alpha['cloud'] is None
    assert c.constituent_alpha['co'] == 'co_ddb' and c.constituent_alpha['h2'] == 'h2_jj_ddb'
    assert c.regridType == 1000 and c.pmin == 0.01 and c.gtype == 'sphere' and c.h2state == 'n'
    assert abs(c.distance - 5.2 * 149597870.691) < 1e-3 and c.orientation == [10.0, -3.0]
    assert c.Req == 71492.0 and c.Rpol == 66854.0 and c.truncate_strength['h2s'] == 1e-22
    c.update_config(limb='sec', Req='70000.0')
    assert c.limb == 'sec' and c.Req == 70000.0


def test_atmosphere_snapshot_roundtrip(tmp_path):
    a = Atmosphere.from_npz(os.path.join(ROOT, 'tests', 'golden', 'atm_neptune.npz'), 'neptune')
    assert a.gas.shape == (16, 1500) and a.config.constituent_alpha['co'] == 'co_ddb'
    assert a.config.orientation == [347.67, -29.08]
    fn = str(tmp_path / 'snap.npz')
    a.to_npz(fn)
    b = Atmosphere.from_npz(fn, 'neptune')
    assert np.array_equal(a.gas, b.gas) and b.config.constituent_alpha == a.config.constituent_alpha


REF_PLANETS = '/root/reference/radiobear'


@pytest.mark.skipif(not os.path.isdir(REF_PLANETS), reason='reference planet files only exist in the build container')
@pytest.mark.parametrize('planet,cfile,gold', [('jupiter', 'config.par', 'atm_jupiter.npz'),
                                               ('neptune', 'config.par', 'atm_neptune.npz'),
                                               ('saturn', 'config.par', 'atm_saturn.npz'),
                                               ('uranus', 'config.par', 'atm_uranus.npz')])
def test_atmosphere_pipeline_matches_reference(planet, cfile, gold, tmp_path, monkeypatch):
    """readGas/readCloud/regrid/tweak/computeProp on the reference's own input files."""
    import shutil
    P = planet.capitalize()
    shutil.copytree(os.path.join(REF_PLANETS, P), str(tmp_path / P))
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(str(tmp_path / P))
    a = Atmosphere(planet, config=cfile)
    assert a.std() == golden(gold)['gas'].shape[1]
    g = golden(gold)
    assert np.max(relerr(a.gas, g['gas'])) < 1e-11
    assert np.max(relerr(a.cloud, g['cloud'])) < 1e-11
    assert np.max(relerr(a.property, g['property'])) < 1e-11


def test_row_partition_balances_on_disc_pixels():
    grid = set_utils.image_grid(0.005)
    q = 66854.0 / 71492.0
    for n in (1, 2, 4, 8):
        parts = parallel.partition_rows(grid, q, n)
        assert parts[0][0] == 0 and parts[-1][1] == len(grid)
        assert all(parts[i][1] == parts[i + 1][0] for i in range(n - 1))
        w = [parallel.row_weights(grid, q)[a:b].sum() for a, b in parts]
        assert max(w) <= 1.25 * (sum(w) / n) + 700


def test_fileio_text_formats_match_reference(tmp_path):
    """fileIO.write: byte-identical files for spectrum / disc spectrum / profile / image (fileIO.py:19-107)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from make_golden import FILEIO_CASES
    from radiobear_b200 import fileIO
    g = golden('fileio.npz')
    for n, (typ, b, Tb) in enumerate(FILEIO_CASES):
        d = data_handling.Data()
        d.set('f', [1.0, 10.5, 22.0])
        d.set('freqUnit', 'GHz')
        d.set('b', b)
        d.set('Tb', Tb)
        d.set('type', typ)
        d.set('header', {'z': '# z line', 'a': '# a line', 'data-type': '#* type:  ' + typ})
        fn = fileIO.FileIO(directory=str(tmp_path)).write(str(tmp_path / 'out{}.dat'.format(n)), d)
        assert open(fn).read() == str(g['case{}'.format(n)])
    from make_golden import data_show
    here = os.getcwd()
    os.chdir(tmp_path)
    try:
        assert data_show(d) == str(g['show'])                   # Data.show / show_header / show_log
    finally:
        os.chdir(here)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's
    keys, the oracle port timed on every host core, zero host<->device bytes."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'pixel*freq/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'C4' in d['config']['workload']


def test_bench_clock_sampler_without_gpu():
    import bench
    s = bench.ClockSampler(0)
    s.start()
    got = s.stop()
    assert set(got) >= {'sm_mhz', 'sm_max_mhz', 'reasons', 'samples'}


@pytest.mark.skipif(not os.path.isdir(REF_PLANETS), reason='imports the reference: build container only')
def test_saturn_regrid_4096_against_live_reference(tmp_path, monkeypatch):
    """SURVEY 8d, config C5's concrete input: Planet('saturn', regridType=4096) -- the reference regrids
    saturn.paulSolar onto 4096 log-spaced pressures (0.01 .. 5000 bar); our Atmosphere pipeline on the same files."""
    import shutil
    import sys
    import types
    shutil.copytree(os.path.join(REF_PLANETS, 'Saturn'), str(tmp_path / 'Saturn'))
    for d in ('Logs', 'Output', 'Scratch'):
        os.makedirs(tmp_path / d)
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(str(tmp_path / 'Saturn'))
    monkeypatch.syspath_prepend(os.path.dirname(REF_PLANETS))
    if 'matplotlib' not in sys.modules:            # atm_modify.py:8 imports it at module level; never used here
        stub = types.ModuleType('matplotlib')
        stub.pyplot = types.ModuleType('matplotlib.pyplot')
        monkeypatch.setitem(sys.modules, 'matplotlib', stub)
        monkeypatch.setitem(sys.modules, 'matplotlib.pyplot', stub.pyplot)
    import radiobear as rb
    ref = rb.planet.Planet('saturn', plot_atm=False, plot_bright=False, verbose=False, regridType=4096).atmos[0]
    from radiobear_b200 import config as pcfg
    c = pcfg.planetConfig('Saturn', configFile='Saturn/config.par')
    c.update_config(regridType=4096)
    a = Atmosphere('saturn', config=c)
    assert a.std() == 4096 and a.gas.shape == ref.gas.shape == (16, 4096)
    assert np.max(relerr(a.gas, ref.gas)) < 1e-10
    assert np.max(relerr(a.cloud, ref.cloud)) < 1e-10
    assert np.max(relerr(a.property, ref.property)) < 1e-10


# ---- round 2: advisor findings ------------------------------------------------------------------------------------
def test_scale_matrix_accepts_constituents_read_back_from_a_file_cache():
    """get_alpha='file' hands Alpha.ordered_constituents back as a numpy string array (np.load); a dict scale must
    still resolve its keys (the reference looks constituents up by key, alpha.py:151-192)."""
    import numpy as np
    from radiobear_b200 import engine
    ordered = np.array(['h2', 'h2o', 'nh3'])
    m = engine.scale_matrix({'nh3': [2.0, 3.0], 'h2': [0.5, 0.25]}, ordered, 2)
    assert m.shape == (3, 2) and np.array_equal(m, [[0.5, 0.25], [1.0, 1.0], [2.0, 3.0]])
    with pytest.raises(ValueError):
        engine.scale_matrix({'co': [1.0, 1.0]}, ordered, 2)


def test_cloud_layer_count_is_validated_before_any_device_work():
    import numpy as np
    from radiobear_b200 import engine
    gas = np.ones((16, 5))
    C = {'Z': 0, 'T': 1, 'P': 2, 'H2': 3, 'HE': 4, 'NH3': 6}
    with pytest.raises(ValueError, match='cloud has 3 layers'):
        engine.alpha_layers([1.0, 2.0], gas[1], gas[2], gas, C, cloud=np.zeros((12, 3)), cloud_dict={'NH3': 6},
                            formalisms=[('nh3', 'nh3_hs')])


def test_fileio_read_back_round_trip(tmp_path):
    """FileIO.read / flist (fileIO.py:109-330): spectrum, disc spectrum, profile and image files written by
    FileIO.write come back as Data with the same f, b and Tb (to the precision of the text format)."""
    import numpy as np
    from radiobear_b200 import fileIO, data_handling
    io = fileIO.FileIO(directory=str(tmp_path))
    cases = [('spectrum', [[0.0, 0.0], [0.5, 0.25]], [[100.123, 200.5, 300.25], [90.1, 80.2, 70.3]]),
             ('spectrum', ['disc'], [[100.123, 200.5, 300.25]]),
             ('profile', [[0.1 * i, 0.0] for i in range(6)], [[100.0 + i, 200.5 + i, 300.25 + i] for i in range(6)]),
             ('image', [[0, 0]] * 12, np.arange(12.0).reshape(3, 4) + 0.5)]
    for n, (typ, b, Tb) in enumerate(cases):
        d = data_handling.Data()
        d.set('f', [1.0, 10.5, 22.0] if typ != 'image' else [22.0])
        d.set('freqUnit', 'GHz')
        d.set('b', b)
        d.set('Tb', Tb)
        d.set('type', typ)
        d.set('header', {'z': '# z line', 'data-type': '#* type:  ' + typ, 'start': '#* start: 2026-01-01 00:00:00'})
        fn = io.write(str(tmp_path / 'jupiter_{}_{}.dat'.format(typ, n)), d)
        out = io.read(fn, file_type=typ)[fn]
        assert out.type == typ and out.start == '2026-01-01 00:00:00'
        assert np.allclose(np.asarray(out.Tb), np.asarray(d.Tb), atol=6e-3)
        if typ != 'image':
            assert np.allclose(out.f, d.f)
        if typ == 'profile':
            assert np.allclose(out.b, np.asarray(b), atol=1e-3)
    assert len(io.flist(None, 'dat')) == 4 and io.flist(1, 'dat') == [sorted(io.flist(None, 'dat'))[1]]
    assert len(io.read(None, tag='dat', file_type='spectrum')) == 2


def test_doppler_host_logic_builds_the_steps_the_reference_takes(monkeypatch):
    """Brightness._doppler_ray without a GPU: trace, absorption and integration are replaced by recorders, the host
    logic in between must ask for exactly what brightness.py:83-92 evaluates -- step i: the lower node (layer i + 1) at
    f / doppler[i], the upper node (layer i) at f / doppler[i + 1] -- and hand the integration the slab pair in the
    roles rb_rt_desc::alpha0 defines.  Checked by integrating the recorded slabs with the oracle's plain loop formula
    against oracle.rt_oracle.integrate_ray_doppler on a synthetic absorption law."""
    import os
    from conftest import GOLDEN
    from oracle import rt_oracle as rto
    from radiobear_b200 import brightness, raypath

    atm = Atmosphere.from_npz(os.path.join(GOLDEN, 'atm_jupiter.npz'), 'jupiter')
    L = atm.gas.shape[1]
    n = 37
    rng = np.random.default_rng(5)
    ray = raypath.Ray()
    ray.update(ds=list(rng.uniform(5.0, 60.0, n)), layer4ds=list(range(n)), doppler=list(1.0 + 0.01 * rng.uniform(-1, 1, n)))
    monkeypatch.setattr(raypath, 'compute_ds', lambda a, b, o=None, gtype=None, verbose=False: ray)
    T = atm.gas[atm.config.C['T']]

    def law(layer, fv):                                       # any smooth function of (layer, frequency)
        return 1e-7 * (1.0 + layer) ** 1.5 * (np.asarray(fv) / 10.0) ** 2

    class FakeAlpha:
        config = atm.config

        def layers_at(self, fm, a):
            assert fm.shape == (L, 3)
            return np.array([law(l, fm[l]) for l in range(L)])
    seen = {}

    def fake_integrate(ds, nseg, alpha_slab, Tl, disc_average=False, tau_cut=0.0, want_intW=False, alpha0_slab=None,
                       profile_ray=-1, **kw):
        seen.update(ds=np.array(ds), nseg=list(nseg), a1=alpha_slab, a0=alpha0_slab)
        return np.zeros((1, 3)), np.zeros((1, 3))
    monkeypatch.setattr(brightness.engine, 'rt_integrate', fake_integrate)
    freqs = [4.0, 22.0, 23.9]
    B = brightness.Brightness(config=atm.config, verbose=False)
    B._doppler_ray([0.3, 0.1], freqs, atm, FakeAlpha(), None, False, False)
    assert seen['nseg'] == [n] and seen['ds'].shape == (1, L - 1) and np.array_equal(seen['ds'][0, :n], ray.ds)
    # the kernel's recurrence on the recorded slabs (rb_rt_desc::alpha0): dtau_i = (a0[i] + a1[i+1]) ds_i / 2, W from a1[i+1]
    tau, W, Tb, iW = np.zeros(3), np.zeros(3), np.zeros(3), np.zeros(3)
    for i in range(n - 1):
        h = ray.ds[i] * 1e5 / 2.0
        tau = tau + (seen['a0'][i] + seen['a1'][i + 1]) * h
        Wn = seen['a1'][i + 1] * np.exp(-tau)
        iW = iW + (Wn + W) * h
        Tb = Tb + (T[i + 1] * Wn + T[i] * W) * h
        W = Wn
    ref = rto.integrate_ray_doppler(ray.ds, ray.layer4ds, ray.doppler, freqs, law, T)
    assert np.max(np.abs(Tb / iW - ref)) < 1e-9


def test_point_blocks_weigh_the_pixels_to_copy_for_host_output():
    """parallel.point_blocks: whole image rows per rank; results that go to the host weigh every pixel of a row (all of
    them are copied) next to its on-disc pixels (device work), so the ranks holding the sky rows get fewer rows than with
    the on-disc weight alone; the answer is cached and handed out as a fresh list."""
    grid = set_utils.image_grid(0.005)
    ncol = len(grid)
    cfg = type('Cfg', (), dict(Rpol=66854.0, Req=71492.0, gtype='ellipse'))()
    for world in (2, 4, 8):
        host = parallel.point_blocks(ncol * ncol, cfg, (grid, ncol), world)
        dev = parallel.point_blocks(ncol * ncol, cfg, (grid, ncol), world, host_output=False)
        for parts in (host, dev):
            assert parts[0][0] == 0 and parts[-1][1] == ncol * ncol
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert all(a % ncol == 0 and b % ncol == 0 for a, b in parts)
        rows_host = [(b - a) // ncol for a, b in host]
        rows_dev = [(b - a) // ncol for a, b in dev]
        assert max(rows_host) < max(rows_dev) or world == 2       # the sky-row blocks at both ends shrink
        assert rows_host[0] <= rows_dev[0] and rows_host[-1] <= rows_dev[-1]
        again = parallel.point_blocks(ncol * ncol, cfg, (grid, ncol), world)
        assert again == host and again is not host
        again.append('x')                                          # a caller's list: the cache must not see this
        assert parallel.point_blocks(ncol * ncol, cfg, (grid, ncol), world) == host
    assert parallel.point_blocks(10, cfg, None, 4) == parallel.partition_even(10, 4)


def test_brightness_profile_tables_match_reference(tmp_path):
    """Brightness.savertm / saveAlpha / saveWeight / saveTau / saveTblayer / saveit (brightness.py:128-250): same file
    names, locations, bytes, return values and exceptions as the reference for the same state of the last ray."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
    from make_golden import rtm_state, rtm_run, RTM_CALLS
    from radiobear_b200 import brightness
    g = golden('rtm_tables.npz')
    here = os.getcwd()
    try:
        got = rtm_run(rtm_state(object.__new__(brightness.Brightness)), str(tmp_path))
    finally:
        os.chdir(here)
    assert sorted(got) == sorted(g.files) and len(RTM_CALLS) == 8
    for k in g.files:
        assert str(got[k]) == str(g[k]), k


def test_frame_rotations_of_the_ray_module():
    """raypath.rotate2planet / rotate2obs (raypath.py:47-57) against the oracle's rotation matrices; one undoes the other
    with the angles negated."""
    from radiobear_b200 import raypath
    from oracle import ray_oracle as ro
    rng = np.random.default_rng(3)
    for _ in range(20):
        tip, rot = rng.uniform(-1.5, 1.5, 2)
        v = rng.normal(size=3)
        p = raypath.rotate2planet(rot, tip, v)
        assert np.array_equal(p, ro.rotX(rot, ro.rotZ(tip, v)))
        assert np.array_equal(raypath.rotate2obs(rot, tip, v), ro.rotZ(tip, ro.rotX(rot, v)))
        assert np.max(np.abs(raypath.rotate2obs(-rot, -tip, p) - v)) < 1e-14


def test_planet_run_writes_log_and_output_file(tmp_path, monkeypatch, capsys):
    """Host side of Planet.run with write_log_file / write_output_files (planet.py:45-52, 103-146): absorption and
    integration replaced by recorders; the log holds what the reference logs, the spectrum file reads back."""
    from radiobear_b200.planet import Planet
    from conftest import GOLDEN
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, 'atm_jupiter.npz'), 'jupiter')
    j = Planet('jupiter', atmosphere=atm, verbose=False, write_log_file=True, write_output_files=True,
               log_directory=str(tmp_path / 'Logs'), output_directory=str(tmp_path / 'Output'))
    want = np.array([[500.25, 140.5], [480.125, 139.75]])
    monkeypatch.setattr(j, 'alpha_layers', lambda **kw: None)
    monkeypatch.setattr(j.bright, 'prefetch', lambda *a, **kw: None)
    monkeypatch.setattr(j.bright, 'batch', lambda pts, *a, **kw: {'Tb': want.copy()})
    rv = j.run([2.0, 22.0], b=[[0.0, 0.0], [0.4, 0.3]])
    outs = os.listdir(tmp_path / 'Output')
    assert len(outs) == 1 and outs[0].startswith('Jupiter_spectrum_') and outs[0].endswith('.dat')
    d = list(j.fIO.read(file_type='spectrum').values())[0]
    assert np.allclose(d.f, [2.0, 22.0]) and np.max(np.abs(d.Tb - want)) < 1e-3
    assert rv.header['log-file:'] == '#* logfile: ' + j.log.logfile
    capsys.readouterr()
    rv.show(include=['log'])
    text = capsys.readouterr().out
    for piece in ('<<<Log>>>', 'Jupiter start ', 'config.par', 'Run parameters:', 'Jupiter at 2.0 GHz', 'Run start ', 'Run stop '):
        assert piece in text, piece
    j.fIO.show()                                                     # header + values of the file read back; no log key
    assert outs[0] in capsys.readouterr().out


def test_small_utilities_against_the_live_reference(tmp_path):
    """utils.getRFband / invertDictionary / ls / get_expected_number_of_entries (utils.py:105-182) give what the
    reference's give (build container only: the reference is imported)."""
    import sys
    import io
    if not os.path.isdir('/root/reference/radiobear'):
        pytest.skip('reference not present')
    sys.path.insert(0, '/root/reference')
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_utils', '/root/reference/radiobear/utils.py')
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for f, u in [(0.001, 'GHz'), (1.0, 'GHz'), (1.9999, 'GHz'), (22.0, 'GHz'), (26.5, 'GHz'), (109.0, 'GHz'), (110.0, 'GHz'), (1400.0, 'MHz')]:
        assert utils.getRFband(f, u) == ref.getRFband(f, u)
    d = {'a': 3, 'b': 1, 'c': 2}
    assert utils.invertDictionary(d) == ref.invertDictionary(d) and utils.invertDictionary(d, True) == ref.invertDictionary(d, True)
    for name in ('x.dat', 'y.dat', '.hidden.dat', 'z.txt'):
        (tmp_path / name).write_text('')
    for tag in ('dat', None, 'txt'):
        assert sorted(utils.ls(str(tmp_path), tag, show=False, returnList=True)) == \
            sorted(ref.ls(str(tmp_path), tag, show=False, returnList=True))
    table = '# comment\n1 2 3 4\n5 6 7 8\n1 2\nbad line here x\n9 9 9 9\n'
    assert utils.get_expected_number_of_entries(io.StringIO(table)) == ref.get_expected_number_of_entries(io.StringIO(table)) == 4
    bad = '10 20 30\n' * 3 + '10 20\n' * 2 + '10 20 30 40\n'
    for mod in (utils, ref):
        with pytest.raises(ValueError):
            mod.get_expected_number_of_entries(io.StringIO(bad))


def _reference_module(name):
    """A module of the reference imported read-only (build container only)."""
    import sys
    if not os.path.isdir('/root/reference/radiobear'):
        pytest.skip('reference not present')
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
    import importlib
    return importlib.import_module('radiobear.' + name)


def test_request_parsers_against_the_live_reference(tmp_path):
    """set_utils.set_freq / set_b (set_utils.py:12-150) on every request form the reference can parse: same frequency
    lists (units converted), same impact points, data type and image size, same exception for a bad request."""
    ref = _reference_module('set_utils')
    ffile = tmp_path / 'freqs.txt'
    ffile.write_text('1.5\n2.5\n10.0\n')
    for req, unit in [([1.0, 2.0, 3.5], 'GHz'), (np.array([4.0, 8.0]), 'GHz'), ('1:100:5', 'GHz'), ('1,2.5,8', 'GHz'),
                      ('22.2', 'GHz'), (43, 'GHz'), (1400.0, 'MHz'), ('100:1000:100', 'MHz'), (list(np.linspace(1, 100, 64)), 'GHz'),
                      (str(ffile), 'GHz'), ([1.0e9, 2.0e9], 'Hz')]:
        mine, mu = set_utils.set_freq(req.copy() if hasattr(req, 'copy') else req, unit)
        want, wu = ref.set_freq(req.copy() if hasattr(req, 'copy') else req, unit)
        assert mu == wu and len(mine) == len(want) and all(float(a) == float(b) for a, b in zip(mine, want)), (req, unit)
    for mod in (set_utils, ref):
        with pytest.raises(ValueError):
            mod.set_freq({'a': 1})
    with pytest.raises(ValueError):                              # the reference lets float(None)'s TypeError through
        set_utils.set_freq(None)
    kw = dict(Rpol=66854.0, Req=71492.0)
    for req in ['disc', 'DISK', [0.1, 0.2], [[0.0, 0.0], [0.3, 0.1]], [[0.01 * i, 0.0] for i in range(7)], '0.0:1.0:0.01<0',
                '0.0:0.9:0.1<45', '0.1,0.2,0.95<90', '0.2,0.4', 'stamp:0.1:-0.2,0.2,-0.1,0.1', 'stamp:0.05:0.0,0.2,0.3,0.4']:
        mine, want = set_utils.set_b(req, **kw), ref.set_b(req, **kw)
        assert mine.data_type == want.data_type, req
        assert np.array_equal(np.asarray(mine.b), np.asarray(want.b)), req
        if isinstance(req, str) and req.startswith('stamp'):
            # the reference's imSize is [pixels per row, len(request string) / pixels per row] (set_utils.py:51): the row
            # length agrees, the second entry is meaningless there and the row count here
            assert mine.imSize[0] == want.imSize[0] and mine.imSize[0] * mine.imSize[1] == len(mine.b)
        else:
            assert mine.imSize == want.imSize
    # a float request builds the full grid in both; the reference then fails on len(float) (set_utils.py:78): compare
    # with the points it had built by then
    grid = -1.0 * np.flipud(np.arange(0.25, 1.5 + 0.25, 0.25))
    grid = np.concatenate((grid, np.arange(0.0, 1.5 + 0.25, 0.25)))
    with pytest.raises(TypeError):
        ref.set_b(0.25, **kw)
    mine = set_utils.set_b(0.25, **kw)
    assert np.array_equal(np.asarray(mine.b), np.array([[x, y] for y in grid for x in grid])) and mine.imSize == [13, 13]


def test_check_reuse_against_the_live_reference():
    """Planet.check_reuse (planet_base.py:303-346) decides whether a run recomputes the absorption: same answer as the
    reference's method for the previous / new request pairs a retrieval loop produces (one-sided tolerances included)."""
    ref = _reference_module('planet_base').PlanetBase.check_reuse
    from radiobear_b200.planet import Planet
    mine = Planet.check_reuse
    rng = np.random.default_rng(5)
    f0 = [1.0, 2.0, 4.0, 8.0]
    scales = [False, 1.0, 1.00001, 0.5, 2.0, [1.0, 2.0, 3.0], [1.0, 2.0, 3.0003], [1.0, 2.0, 2.9], [1.0, 2.0],
              {'nh3': [1.0, 2.0], 'h2o': [0.5, 0.5]}, {'nh3': [1.0, 2.0], 'h2o': [0.5, 0.4]},
              {'nh3': [1.0, 2.0]}, {'nh3': [1.0, 2.0], 'h2s': [0.5, 0.5]}, {'nh3': [1.0], 'h2o': [0.5, 0.5]}]
    freq_sets = [f0, list(reversed(f0)), [1.0, 2.0, 4.0, 8.001], [1.0, 2.0, 4.0, 7.9], [0.9, 2.0, 4.0, 8.0], [10.0, 20.0, 40.0, 80.0],
                 [1.0, 2.0, 4.0], f0 + [16.0]]
    n = 0
    for prev_scale in scales:
        for prev_f in (f0, [10.0, 20.0, 40.0, 80.0]):
            for prev_ga, prev_sa in (('none', 'none'), ('memory', 'none'), ('none', 'memory')):
                state = type('S', (), dict(freqs=prev_f, scale=prev_scale, get_alpha=prev_ga, save_alpha=prev_sa))()
                for _ in range(12):
                    sc = scales[rng.integers(len(scales))]
                    fr = freq_sets[rng.integers(len(freq_sets))]
                    ga, sa = [('none', 'none'), ('memory', 'none'), ('none', 'memory')][rng.integers(3)]
                    ov = ['check', 'check', 'check', 'true', 'false'][rng.integers(5)]
                    try:
                        want = ref(state, fr, sc, ga, sa, ov)
                    except Exception as e:                       # e.g. bool - dict: the reference raises, so do we
                        with pytest.raises(type(e)):
                            mine(state, fr, sc, ga, sa, ov)
                        continue
                    assert mine(state, fr, sc, ga, sa, ov) == want, (prev_scale, prev_f, sc, fr, ga, sa, ov)
                    n += 1
    assert n > 800
    # a numpy array as per-layer scale: the reference's isanynum lets float(array)'s TypeError through; here it compares
    arr = type('S', (), dict(freqs=f0, scale=np.array([1.0, 2.0, 3.0]), get_alpha='none', save_alpha='none'))()
    assert mine(arr, f0, np.array([1.0, 2.0, 3.0]), 'none', 'none') and not mine(arr, f0, np.array([1.0, 2.0, 3.5]), 'none', 'none')


def test_total_layer_alpha_builds_the_layer_scale_the_reference_applies(monkeypatch):
    """Alpha.total_layer_alpha / the {constituent: number} scale of get_single_layer (alpha.py:151-192, 218-233): the
    host side (scale column per constituent, unknown names ignored, per-layer cache list) with the device scale-sum
    replaced by its definition, against the reference's method on the same array."""
    from radiobear_b200 import alpha as rbalpha
    ref_alpha = _reference_module('alpha').Alpha

    def scale_sum(cube, scale_mat=None, want_cube=False, ctx=None):           # rb_alpha_scale_sum, stated in numpy
        cube = np.asarray(cube, dtype=np.float64)
        scaled = cube * (1.0 if scale_mat is None else np.asarray(scale_mat).T[:, None, :])
        return (scaled.sum(axis=2), scaled) if want_cube else scaled.sum(axis=2)
    monkeypatch.setattr(engine, 'alpha_scale_sum', scale_sum)
    rng = np.random.default_rng(9)
    names = ['h2', 'h2o', 'h2s', 'nh3', 'ph3']
    absorb = rng.random((6, 5)) * 1e-5
    mine = object.__new__(rbalpha.Alpha)
    ref = object.__new__(ref_alpha)
    for obj in (mine, ref):
        obj.ordered_constituents, obj.freqs, obj._save_alpha_memfil, obj.tosave = names, np.arange(6.0), True, []
    for lscale in (1.0, 2.5, 3, {'nh3': 0.5, 'h2o': 2.0}, {'ph3': 0.0, 'bogus': 9.0}, {}):
        got, want = mine.total_layer_alpha(absorb.copy(), lscale), ref.total_layer_alpha(absorb.copy(), lscale)
        assert got.shape == want.shape and np.max(relerr(got, want)) < 1e-15, lscale
        assert np.max(relerr(mine.tosave[-1], ref.tosave[-1])) == 0.0
    assert mine._one_layer_scale({'nh3': 0.5, 'bogus': 2.0}) == {'nh3': [0.5]} and mine._one_layer_scale(2.0) == 2.0
    with pytest.raises(ValueError):
        mine.total_layer_alpha(absorb[0], 1.0)


def test_committed_bench_lines_carry_the_contract():
    """The bench lines measured on B200 and committed under profiles/ (the round's final N = 1 / 2 / 4 / 8 runs): every
    key the measurement contract names, the roofline on the binding resource with hbm beside it, sane relations between
    the numbers (kernel time <= step time <= Planet.run time, achieved <= peak, bytes moved <= HBM peak x time)."""
    import json
    seen = 0
    for name in ('r2_final2_bench_n1.json', 'r2_final_bench_n1.json', 'r2_final_bench_n2.json', 'r2_final_bench_n4.json',
                 'r2_final_bench_n8.json'):
        path = os.path.join(ROOT, 'profiles', name)
        lines = [ln for ln in open(path) if ln.startswith('{')]
        d = json.loads(lines[-1])
        seen += 1
        for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                  'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'gpu_launches', 'roofline'):
            assert k in d, (name, k)
        assert d['unit'] == 'pixel*freq/s' and d['dtype'] == 'f64' and d['vs_baseline'] is None and d['warmup'] >= 3
        assert 'C4' in d['config']['workload'] and 'l2' in d['config'] and d['gpu_launches'] > 0
        assert d['clocks']['sm_mhz'] >= 0.9 * d['clocks']['sm_max_mhz']
        assert not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
        e = d['e2e']
        assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and e['value'] < d['value']
        assert abs(d['value'] - d['config']['on_disc_pixels'] * d['config']['freqs'] / (d['ms_per_step'] * 1e-3)) < 1e-6 * d['value']
        r = d['roofline']
        assert r['bound'] == 'fp64' and r['unit'] == 'TFLOP/s' and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
        assert 0.3 < r['frac'] < 1.0 and r['ms_per_launch'] * r['launches_per_step'] <= d['ms_per_step'] * 1.001
        h = r['hbm']
        assert h['peak'] == 6453.4 and abs(h['frac'] - h['achieved'] / h['peak']) < 1e-9 and h['frac'] < 1.0
        if d['n_gpus'] == 1:
            c = d['cpu_baseline']
            assert c['kind'] == 'port' and c['cores'] >= 1 and c['value'] > 0 and 'sample' in c
            assert r['traffic'] >= h['algorithmic_bytes'] and r['traffic'] < 1.2 * h['algorithmic_bytes']
            assert r['traffic'] / (r['ms_per_launch'] * 1e-3) < h['peak'] * 1e9
    assert seen == 5
