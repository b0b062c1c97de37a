"""GPU: the executive API -- Planet.run(freqs, b) -> DataReturn, Brightness.single side attributes."""
import os

import numpy as np
import pytest

from conftest import golden, relerr, GOLDEN

pytestmark = pytest.mark.gpu


def make_planet(name, snap):
    from radiobear_b200.atmosphere import Atmosphere
    from radiobear_b200.planet import Planet
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, snap), name)
    return Planet(name, atmosphere=atm, verbose=False)


def test_benchmark_script_flow():
    """The flow of scripts/benchmark.py:10-24."""
    tb = golden('tb.npz')
    j = make_planet('jupiter', 'atm_jupiter_benchmark.npz')
    freqs = [0.6, 1.25, 2.6, 5.2, 10, 21.9]
    b = [[0.0, 0.0], [np.sin(np.radians(15)), 0.0], [np.sin(np.radians(30)), 0.0], [np.sin(np.radians(45)), 0.0]]
    rv = j.run(freqs, b=b, reuse_override='False')
    assert rv.Tb.dtype == np.float32 and rv.Tb.shape == (4, 6) and rv.type == 'spectrum'
    assert np.max(np.abs(rv.Tb - tb['bench_tb_f32'])) < 1e-3
    assert np.allclose(rv.f, freqs) and rv.header['gtype'] == '# gtype: ellipse'


def test_disc_profile_image_requests():
    tb = golden('tb.npz')
    j = make_planet('jupiter', 'atm_jupiter.npz')
    rv = j.run('1:100:5', b='disc')
    assert rv.b == ['disc'] and rv.type == 'spectrum' and len(rv.f) == 21
    assert np.max(np.abs(np.asarray(j.Tb) - tb['c1_tb'])) < 1e-4
    a0 = j.alpha[0].layers
    j.run('1:100:5', b=[0.1, 0.2])
    assert j.alpha[0].layers is a0                                   # check_reuse skipped the absorption step
    rv = j.run(list(np.linspace(1, 50, 50)), b='0.0:1.0:0.01<0')     # config C3
    assert rv.type == 'profile' and rv.Tb.shape == (100, 50)
    assert np.isnan(rv.Tb[-1]).all() and np.isfinite(rv.Tb[:90]).all()
    rv = j.run([10.0, 22.0], b='stamp:0.1:-0.2,0.2,-0.1,0.1')
    assert rv.type == 'image' and rv.Tb.shape == (5, 3, 2)
    rv = j.run(22.0, b=0.05)                                          # float b: full image (crashes in the reference)
    assert rv.type == 'image' and rv.Tb.shape == (61, 61)
    assert rv.Tb[0, 0] == np.float32(2.725) and rv.Tb[30, 30] > 100.0
    rv = j.run('1;100;8', b='disc', scale=2.0)                        # log sweep string + scalar scale
    assert len(rv.f) == 8


def test_brightness_single_attributes():
    tb = golden('tb.npz')
    j = make_planet('jupiter', 'atm_jupiter.npz')
    freqs = list(tb['c1_freqs'])
    j.alpha_layers(freqs, j.atmos)
    Tb = j.bright.single('disc', freqs, j.atmos[0], j.alpha[0], j.config.orientation)
    assert np.max(np.abs(np.array(Tb) - tb['c1_tb'][0])) < 1e-4
    B = j.bright
    assert B.tau.shape == tb['c1_tau'].shape and np.max(relerr(B.tau, tb['c1_tau'])) < 1e-8
    assert B.Tb_lyr.shape == tb['c1_Tb_lyr'].shape and len(B.P) == B.tau.shape[1] == len(B.z)
    assert np.max(relerr(B.integrated_W, tb['c1_integrated_W'])) < 1e-8
    assert len(B.travel.ds) == 999
    assert j.bright.single([1.0, 0.2], freqs, j.atmos[0], j.alpha[0]) == [2.725] * len(freqs)


def test_neptune_c2_run():
    n = golden('neptune_c2.npz')
    p = make_planet('neptune', 'atm_neptune.npz')
    rv = p.run(list(n['freqs']), b='disc')
    assert np.max(np.abs(np.asarray(p.Tb) - n['tb'])) < 1e-4
    assert p.alpha[0].ordered_constituents == [str(x) for x in n['ordered_constituents']]


def test_planet_run_with_the_gravity_shape():
    """config gtype = 'gravity' through the executive: Planet.run and raypath.compute_ds take the geoid of
    shape.py:141-221 (its model read from the planet's config: Jn, RJ, omega_m, GM profile, zonal winds); an image
    request (>= 512 rays, prefetched geometry) gives the same pixels as a list of points; Brightness.single works."""
    from radiobear_b200 import raypath
    j = make_planet('jupiter', 'atm_jupiter.npz')
    e = make_planet('jupiter', 'atm_jupiter.npz')
    j.config.gtype = 'gravity'
    pts = [[0.0, 0.0], [0.3, 0.2], [0.6, -0.4]]
    rg = j.run([2.0, 10.0, 30.0], b=pts, reuse_override='false')
    re_ = e.run([2.0, 10.0, 30.0], b=pts, reuse_override='false')
    assert rg.header['gtype'] == '# gtype: gravity' and rg.Tb.shape == (3, 3) and np.isfinite(rg.Tb).all()
    tb_g = np.array(rg.Tb, dtype=float)                              # (Planet.run hands out the same Data object every call)
    d = np.abs(tb_g - np.asarray(re_.Tb, dtype=float))
    assert 1e-4 < d[1:].max() < 5.0                                  # another shape: close, not equal
    ray = raypath.compute_ds(j.atmos[0], pts[1], j.config.orientation)
    ell = raypath.compute_ds(e.atmos[0], pts[1], e.config.orientation)
    assert len(ray.ds) == len(ell.ds) == len(ray.r4ds) and abs(ray.r4ds[0] - ell.r4ds[0]) > 1.0
    img = np.array(j.run([10.0], b=0.05, reuse_override='false').Tb)  # 61 x 61 image through the batched path
    grid = np.array(j.b).reshape(61, 61, 2)
    pick = [(30, 30), (36, 34), (22, 42), (2, 2)]
    one = j.run([10.0], b=[list(grid[iy, ix]) for iy, ix in pick], reuse_override='false')
    assert np.allclose([img[iy, ix] for iy, ix in pick], np.asarray(one.Tb)[:, 0], rtol=0, atol=1e-4)
    assert img[2, 2] == np.float32(2.725)
    tb1 = j.bright.single(pts[1], [10.0], j.atmos[0], j.alpha[0], j.config.orientation)
    assert abs(tb1[0] - tb_g[1, 1]) < 1e-3 and j.bright.travel.r4ds is not None


def test_doppler_shifted_absorption():
    """config Doppler (brightness.py:80-96): the reference's branch, run with its renamed `alpha.get_alpha` call
    restored (tests/golden/make_golden.py section `doppler`, omega_m x 100).  Here: per-layer frequency lists in the
    absorption kernel (two launches per ray) + the integration with the slab pair (rb_rt_desc::alpha0)."""
    d = golden('doppler.npz')
    j = make_planet('jupiter', 'atm_jupiter.npz')
    g = golden('ray_fields.npz')
    j.config.vwlat, j.config.vwdat = list(g['vwlat_jupiter']), list(g['vwdat_jupiter'])      # config.zonal table
    j.config.omega_m = j.config.omega_m * float(d['omega_factor'])
    freqs, bpts = [float(x) for x in d['freqs']], [[float(x), float(y)] for x, y in d['b']]
    j.config.Doppler = False
    j.run(freqs, b=bpts, reuse_override='False')
    plain = np.asarray(j.Tb, dtype=np.float64)
    assert np.max(np.abs(plain - d['tb_plain'])) < 1e-4
    j.config.Doppler = True
    rv = j.run(freqs, b=bpts, reuse_override='False')
    got = np.asarray(j.Tb, dtype=np.float64)
    assert rv.Tb.shape == (4, 3)
    assert np.max(np.abs(got - d['tb_doppler'])) < 1e-4, np.abs(got - d['tb_doppler'])
    assert np.max(np.abs(got - plain)[:2]) > 0.1                     # the branch does something
    assert np.max(np.abs(got[2] - plain[2])) < 1e-9                  # central meridian: no line-of-sight velocity
    assert (got[3] == 2.725).all()
    j.run(freqs, b='disc', reuse_override='False')                   # b = (0, 0): doppler = 1, E2 weighting
    assert np.max(np.abs(np.asarray(j.Tb, dtype=np.float64) - d['tb_disc'])) < 1e-4
    # Brightness.single under Doppler: Tb and the profile attributes of the ray
    j.alpha_layers(freqs, j.atmos)
    Tb = j.bright.single(bpts[1], freqs, j.atmos[0], j.alpha[0], j.config.orientation)
    assert np.max(np.abs(np.array(Tb) - d['tb_doppler'][1])) < 1e-4
    B = j.bright
    n = len(B.travel.ds)
    assert B.tau.shape == (3, n) and B.W.shape == (3, n) and B.Tb_lyr.shape == (3, n)
    assert np.all(np.diff(B.tau, axis=1) >= 0) and np.max(np.abs(B.Tb_lyr[:, -1] / B.integrated_W - np.array(Tb))) < 1e-9
    assert np.max(np.abs(np.array(B.travel.doppler) - d['doppler1'])) < 1e-12


def test_log_output_files_and_profile_tables(tmp_path, capsys):
    """write_log_file / write_output_files through Planet.run (planet.py:45-52, 103-146) and the profile tables of the
    last ray (Brightness.saveAlpha / saveWeight / saveTau / saveTblayer, brightness.py:166-250): files land where the
    reference puts them, read back to the values the run returned."""
    from radiobear_b200.atmosphere import Atmosphere
    from radiobear_b200.planet import Planet
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, 'atm_jupiter.npz'), 'jupiter')
    j = Planet('jupiter', atmosphere=atm, verbose=False, write_log_file=True, write_output_files=True,
               log_directory=str(tmp_path / 'Logs'), output_directory=str(tmp_path / 'Output'))
    freqs = [2.0, 22.0]
    rv = j.run(freqs, b=[[0.0, 0.0], [0.4, 0.3]])
    tb = np.array(rv.Tb, dtype=np.float64)
    outs = os.listdir(tmp_path / 'Output')
    assert len(outs) == 1 and outs[0].startswith('Jupiter_spectrum_') and outs[0].endswith('.dat')
    back = j.fIO.read(file_type='spectrum')
    d = list(back.values())[0]
    assert np.allclose(d.f, freqs) and np.max(np.abs(d.Tb - tb)) < 1e-3
    rv.show(include=['log'])                                         # closes the log and prints it
    text = capsys.readouterr().out
    assert '<<<Log>>>' in text and 'Run start ' in text and 'Run stop ' in text and 'Run parameters:' in text
    assert 'Jupiter at 2.0 GHz' in text                              # verbose False: the short frequency line
    # the profile tables of one ray
    B = j.bright
    Tb1 = B.single([0.4, 0.3], freqs, j.atmos[0], j.alpha[0], j.config.orientation)
    assert np.max(np.abs(np.array(Tb1) - tb[1])) < 1e-3
    here = os.getcwd()
    os.chdir(tmp_path)
    try:
        B.saveAlpha('alpha.out', str(tmp_path))
        assert B.saveTau() == 'tau.out (2 x {})'.format(len(B.P))
        B.saveTblayer()
        B.saveWeight(norm=True)
        B.saveit()
        def table(fn):
            rows = [ln.replace('np.float64(', '').replace(')', '').split() for ln in open(fn) if not ln.startswith('#')]
            return np.array(rows, dtype=np.float64)
        n = len(B.P)
        t = table('tau.out')
        assert t.shape == (n, 4) and np.array_equal(t[:, 0], B.P) and np.array_equal(t[:, 2:], B.tau.T)
        assert np.array_equal(table('tblayer.out')[:, 2:], B.Tb_lyr.T)
        assert np.array_equal(table('alpha.out')[:, 2:], np.asarray(j.alpha[0].layers)[:, :n].T)
        w = table('wgt.out')[:, 2:]
        assert np.allclose(w.max(axis=0), 1.0) and np.allclose(w * B.W.max(axis=1), B.W.T, rtol=1e-14, atol=0)
        p = table('pawtt_22.000.out')
        assert p.shape == (n, 5) and np.array_equal(p[:, 3], B.tau[1]) and np.array_equal(p[:, 4], B.Tb_lyr[1])
    finally:
        os.chdir(here)
