"""GPU: ray geometry and brightness-temperature parity (through the C ABI).

Bars: ds within 1e-8 relative of the reference (the kernel is trig-free algebra, the reference goes
through asin/atan2/sin/cos: agreement is ~1e-10), NaN limb segments at the same depth, Tb within 0.01 K
(north_star), off-planet rays = 2.725 K."""
import os

import numpy as np
import pytest

from conftest import golden, keymap, relerr, formalisms_of, TRUNC, GOLDEN

pytestmark = pytest.mark.gpu
TB_TOL = 0.01


MIXED_TOL = 1e-3      # K; measured ~3e-5 K (one float32 ulp of Tb), asserted well inside the 0.01 K bar


@pytest.fixture(scope='module')
def eng():
    """The engine with the batched integration in FP64 (the exactness assertions below: 1e-7 K, bit identity);
    the mixed-precision integration has its own tests at the end of the file."""
    from radiobear_b200 import engine
    before = engine.rt_precision()
    engine.set_rt_precision('f64')
    yield engine
    engine.set_rt_precision(before)


@pytest.fixture()
def mixed(eng):
    eng.set_rt_precision('mixed')
    yield eng
    eng.set_rt_precision('f64')


def geom(a):
    LP = keymap(a['LP_keys'])
    return dict(radius=a['property'][LP['R']], refr_index=a['property'][LP['N']], Req=float(a['Req']),
                Rpol=float(a['Rpol']), orientation=a['orientation'], gtype=str(a['gtype']), limb=str(a['limb']))


@pytest.mark.parametrize('planet', ['jupiter', 'neptune'])
def test_compute_ds_golden(eng, planet):
    from oracle import ray_oracle as ro
    g = golden('rays.npz')
    a = golden('atm_{}.npz'.format(planet))
    ds, nseg, aspect = eng.compute_ds(b=g['b'], **geom(a))
    tip, rotate = ro.compute_aspect(a['orientation'], 1.0 - float(a['Rpol']) / float(a['Req']))
    assert abs(aspect[0] - tip) < 1e-15 and abs(aspect[1] - rotate) < 1e-15 and aspect[2] == geom(a)['radius'][0]
    for i, ns in enumerate(g['nseg_' + planet]):
        if ns == 0:
            assert nseg[i] == -1                                   # Ray.ds is None
            continue
        assert nseg[i] == ns
        ref, d = g['ds_' + planet][i][:ns], ds[i][:ns]
        # The recurrence is chaotic once a ray grazes a shell (the reference's asin/acos pair reflects the
        # direction when cos(t_inc) < 0): compare strictly up to the first NaN of either, minus a margin.
        first_nan = min(int(np.argmax(np.isnan(ref))) if np.isnan(ref).any() else ns,
                        int(np.argmax(np.isnan(d))) if np.isnan(d).any() else ns)
        assert np.isnan(ref).any() == np.isnan(d).any()
        smooth = max(1, first_nan - 300) if np.isnan(ref).any() else ns
        assert np.max(relerr(d[:smooth], ref[:smooth])) < 1e-8


def test_compute_ds_secant_and_sphere(eng):
    from oracle import ray_oracle as ro
    g = golden('rays.npz')
    a = golden('atm_jupiter.npz')
    ga = geom(a)
    ga['limb'] = 'sec'
    ds, nseg, _ = eng.compute_ds(b=g['b'][:6], **ga)
    assert np.max(relerr(ds, g['ds_jupiter_sec'][:, :999])) < 1e-8
    ga = geom(a)
    ga['gtype'] = 'sphere'
    ds, nseg, _ = eng.compute_ds(b=g['b'][:6], **ga)
    for i in range(6):
        o = ro.compute_ds(ga['radius'], ga['refr_index'], g['b'][i], ga['Req'], ga['Rpol'], ga['orientation'], 'sphere')
        assert nseg[i] == len(o['ds']) and np.max(relerr(ds[i], o['ds'])) < 1e-8
    with pytest.raises(NotImplementedError):                 # shape.py:107 calls a name-mangled method: broken upstream
        eng.compute_ds(b=g['b'][:1], **dict(geom(a), gtype='reference'))
    with pytest.raises(ValueError):                           # 'gravity' without the planet's gravity model
        eng.compute_ds(b=g['b'][:1], **dict(geom(a), gtype='gravity'))


def test_raypath_api(eng):
    from radiobear_b200.atmosphere import Atmosphere
    from radiobear_b200 import raypath
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, 'atm_jupiter.npz'), 'jupiter')
    g = golden('rays.npz')
    ray = raypath.compute_ds(atm, [0.5, 0.3])
    assert len(ray.ds) == 999 and ray.layer4ds == list(range(999)) and ray.rNorm == atm.property[1][0]
    assert np.max(relerr(np.array(ray.ds), g['ds_jupiter'][1])) < 1e-8
    assert raypath.compute_ds(atm, [1.0, 0.2]).ds is None
    with pytest.raises(ValueError):
        ray.update(bogus=1)


def _slab(eng, a, freqs):
    C = keymap(a['C_keys'])
    return eng.alpha_layers(freqs, a['gas'][C['T']], a['gas'][C['P']], a['gas'], C, formalisms=formalisms_of(a),
                            other_dicts={'h2': {'h2state': 'e'}, 'co': {'coshape': 'voigt'}}, truncate_strength=TRUNC)


def test_reference_known_answer_table(eng):
    """scripts/benchmark.py:10-24 end to end on the GPU path."""
    a = golden('atm_jupiter_benchmark.npz')
    tb = golden('tb.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, tb['bench_freqs'])
    res = eng.rt_batch(b=tb['bench_b'], alpha_slab=slab, T=a['gas'][C['T']], **geom(a))
    assert np.max(np.abs(res['Tb'] - tb['bench_tb'])) < 1e-4
    res32 = eng.rt_batch(b=tb['bench_b'], alpha_slab=slab, T=a['gas'][C['T']], out_f32=True, **geom(a))
    assert res32['Tb'].dtype == np.float32 and np.max(np.abs(res32['Tb'] - tb['bench_tb_f32'])) < 1e-3


def test_c1_disc_average_and_profiles(eng):
    """Config C1: Jupiter disc-averaged '1:100:5' -- E2 weighting, tau / W / Tb_lyr profiles."""
    a = golden('atm_jupiter.npz')
    tb = golden('tb.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, tb['c1_freqs'])
    assert np.max(relerr(slab.T, tb['c1_alpha'])) < 1e-9
    res = eng.rt_batch(b=[[0.0, 0.0]], alpha_slab=slab, T=a['gas'][C['T']], disc_average=True, tau_cut=0.0,
                       want_intW=True, profile_ray=0, **geom(a))
    assert np.max(np.abs(res['Tb'] - tb['c1_tb'])) < 1e-4
    assert np.max(relerr(res['integrated_W'][0], tb['c1_integrated_W'])) < 1e-8
    n = tb['c1_tau'].shape[1]
    assert np.max(relerr(res['tau'][:, :n], tb['c1_tau'])) < 1e-8
    big = tb['c1_W'] > 1e-300
    assert np.max(relerr(res['W'][:, :n][big], tb['c1_W'][big])) < 1e-7
    assert np.max(relerr(res['Tb_lyr'][:, :n], tb['c1_Tb_lyr'])) < 1e-8
    cut = eng.rt_batch(b=[[0.0, 0.0]], alpha_slab=slab, T=a['gas'][C['T']], disc_average=True, tau_cut=eng.TAU_CUT, **geom(a))
    assert np.array_equal(cut['Tb'], res['Tb'])                    # the tau cut does not change a single bit


def test_points_disc_and_limb_profile(eng):
    a = golden('atm_jupiter.npz')
    tb = golden('tb.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, tb['pt_freqs'])
    res = eng.rt_batch(b=tb['pt_b'], alpha_slab=slab, T=a['gas'][C['T']], **geom(a))
    assert np.array_equal(np.isnan(res['Tb']), np.isnan(tb['pt_tb']))
    assert np.nanmax(np.abs(res['Tb'] - tb['pt_tb'])) < 1e-4
    assert np.all(res['Tb'][-1] == 2.725)                          # b = (1.0, 0.2) misses the planet
    res = eng.rt_batch(b=[[0.0, 0.0]], alpha_slab=slab, T=a['gas'][C['T']], disc_average=True, **geom(a))
    assert np.max(np.abs(res['Tb'] - tb['disc_tb'])) < 1e-4
    # config C3: limb profile, 50 freqs; the last rays are NaN in the reference
    slab = _slab(eng, a, tb['c3_freqs'])
    res = eng.rt_batch(b=tb['c3_b'], alpha_slab=slab, T=a['gas'][C['T']], **geom(a))
    assert np.array_equal(np.isnan(res['Tb']), np.isnan(tb['c3_tb']))
    assert np.isnan(tb['c3_tb']).any()
    assert np.nanmax(np.abs(res['Tb'] - tb['c3_tb'])) < TB_TOL
    assert np.nanmax(np.abs(res['Tb'] - tb['c3_tb'])) < 1e-4


def test_neptune_c2_disc(eng):
    a = golden('atm_neptune.npz')
    n = golden('neptune_c2.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, n['freqs'])
    res = eng.rt_batch(b=[[0.0, 0.0]], alpha_slab=slab, T=a['gas'][C['T']], disc_average=True, **geom(a))
    assert np.max(np.abs(res['Tb'] - n['tb'])) < 1e-4


def test_uranus_disc_and_points(eng):
    """Uranus (not one of the BASELINE configs): alpha, disc-averaged Tb and two points against the reference."""
    a = golden('atm_uranus.npz')
    u = golden('uranus.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, u['freqs'])
    assert np.max(relerr(slab.T, u['alpha'])) < 1e-9
    T = a['gas'][C['T']]
    res = eng.rt_batch(b=[[0.0, 0.0]], alpha_slab=slab, T=T, disc_average=True, **geom(a))
    assert np.max(np.abs(res['Tb'] - u['disc_tb'])) < 1e-4
    res = eng.rt_batch(b=u['pts'], alpha_slab=slab, T=T, **geom(a))
    assert np.max(np.abs(res['Tb'] - u['pt_tb'])) < 1e-4


def test_image_c4_subset_and_full_size_properties(eng):
    """Config C4: 601 x 601 pixels x 64 freqs.  Parity on the reference-computed subset (on-disc, NaN ring,
    off-disc) and size-independent properties on the full cube: off-disc pixels are exactly 2.725 K, the
    image of an untilted planet is mirror-symmetric in x and y, the NaN ring hugs the limb."""
    from radiobear_b200 import set_utils
    a = golden('atm_jupiter.npz')
    im = golden('image_c4.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, im['freqs'])
    grid = im['grid']
    bsub = np.array([[grid[ix], grid[iy]] for iy, ix in im['pick_iy_ix']])
    res = eng.rt_batch(b=bsub, alpha_slab=slab, T=a['gas'][C['T']], **geom(a))
    ref = im['tb']
    finite = ~np.isnan(ref).any(axis=1)
    assert np.max(np.abs(res['Tb'][finite] - ref[finite])) < TB_TOL
    # NaN <-> NaN at identical pixels
    assert np.array_equal(np.isnan(res['Tb']).any(axis=1), np.isnan(ref).any(axis=1))
    # full cube through the float32 image path
    rv = set_utils.set_b(0.005)
    ball = np.array(rv.b)
    out = eng.rt_batch(b=ball, alpha_slab=slab, T=a['gas'][C['T']], out_f32=True, **geom(a))['Tb'].reshape(601, 601, 64)
    q = float(a['Rpol']) / float(a['Req'])
    xx, yy = np.meshgrid(grid, grid)
    rr = np.sqrt(xx**2 + (yy / q)**2)
    # (the reference's shell radius uses the geocentric latitude as the ellipse parameter, shape.py:240-244,
    #  so its limb sits up to ~0.3 % outside the true ellipse at mid-latitudes)
    assert np.all(out[rr >= 1.01] == np.float32(2.725))
    assert np.all(np.isfinite(out[rr < 0.95])) and out[rr < 0.95].min() > 100.0
    nanmask = np.isnan(out).any(axis=2)
    assert nanmask.any() and rr[nanmask].min() > 0.95
    assert np.array_equal(np.isnan(out), np.isnan(out[:, ::-1]))
    ok = ~np.isnan(out) & ~np.isnan(out[:, ::-1]) & ~np.isnan(out[::-1])
    assert np.max(np.abs(out - out[:, ::-1])[ok]) < 2e-3           # mirror in x
    assert np.max(np.abs(out - out[::-1])[ok]) < 2e-3              # mirror in y
    for k, (iy, ix) in enumerate(im['pick_iy_ix']):
        if finite[k]:
            assert np.max(np.abs(out[iy, ix] - ref[k])) < TB_TOL


def test_rt_integrate_with_supplied_segments(eng):
    """Integration alone on caller-supplied ds (ragged nseg, NaN segments) against the oracle."""
    from oracle import rt_oracle as rto
    a = golden('atm_jupiter.npz')
    tb = golden('tb.npz')
    g = golden('rays.npz')
    C = keymap(a['C_keys'])
    T = a['gas'][C['T']]
    slab = np.ascontiguousarray(tb['c1_alpha'].T)
    rows = [0, 1, 3, 8, 9]
    ds = g['ds_jupiter'][rows].copy()
    nseg = np.array([999, 700, 999, 999, 2], dtype=np.int32)
    out, iw = eng.rt_integrate(ds, nseg, slab, T, tau_cut=0.0, want_intW=True)
    for k in range(len(rows)):
        n = nseg[k]
        ref = rto.integrate_ray(ds[k][:n], np.arange(n), tb['c1_alpha'], T)
        assert np.array_equal(np.isnan(out[k]), np.isnan(ref))
        if not np.isnan(ref).any():
            assert np.max(np.abs(out[k] - ref)) < 1e-6
    d = eng.rt_integrate(ds[:1], nseg[:1], slab, T, disc_average=True)
    ref = rto.integrate_ray(ds[0], np.arange(999), tb['c1_alpha'], T, disc_average=True)
    assert np.max(np.abs(d[0] - ref)) < 1e-6


def test_rays_kernel_matches_small_batch_kernel_and_tau_cut_is_exact(eng):
    """The rays-major kernel (R >= 512: cp.async tiles, table exp, regrouped sums) against the simple
    lanes = frequency kernel (R < 512: libdevice exp, reference operation order) on the same pixels, and
    the default tau_cut (engine.TAU_CUT = 50) against no cut: bit-identical."""
    a = golden('atm_jupiter.npz')
    im = golden('image_c4.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, im['freqs'][::3])
    rng = np.random.default_rng(7)
    th = rng.uniform(0, 2 * np.pi, 1200)
    rad = np.sqrt(rng.uniform(0, 1.02, 1200))
    q = float(a['Rpol']) / float(a['Req'])
    b = np.stack([rad * np.cos(th), rad * np.sin(th) * q], axis=1)
    T = a['gas'][C['T']]
    big = eng.rt_batch(b=b, alpha_slab=slab, T=T, want_intW=True, **geom(a))
    parts = [eng.rt_batch(b=b[i:i + 400], alpha_slab=slab, T=T, want_intW=True, **geom(a)) for i in range(0, 1200, 400)]
    small = np.concatenate([p['Tb'] for p in parts])
    small_w = np.concatenate([p['integrated_W'] for p in parts])
    assert np.array_equal(np.isnan(big['Tb']), np.isnan(small))
    assert np.isnan(small).any() and (small == 2.725).any()
    assert np.nanmax(np.abs(big['Tb'] - small)) < 1e-7            # K
    ok = ~np.isnan(small_w) & (small_w > 0)
    assert np.max(np.abs(big['integrated_W'][ok] / small_w[ok] - 1.0)) < 1e-10
    nocut = eng.rt_batch(b=b, alpha_slab=slab, T=T, tau_cut=0.0, **geom(a))
    assert np.array_equal(nocut['Tb'][~np.isnan(small)], big['Tb'][~np.isnan(small)])
    f32 = eng.rt_batch(b=b, alpha_slab=slab, T=T, out_f32=True, **geom(a))['Tb']
    assert np.array_equal(f32[~np.isnan(small)], big['Tb'][~np.isnan(small)].astype(np.float32))


def test_rt_edge_cases(eng):
    a = golden('atm_jupiter.npz')
    C = keymap(a['C_keys'])
    T = a['gas'][C['T']]
    slab = _slab(eng, a, np.array([22.0]))
    # single frequency, ragged ray counts around the 32-ray tile and the 512-ray kernel switch
    for R in (1, 31, 33, 511, 512, 513, 1025):
        b = np.zeros((R, 2))
        b[:, 0] = np.linspace(-0.9, 0.9, R)
        out = eng.rt_batch(b=b, alpha_slab=slab, T=T, **geom(a))['Tb']
        assert out.shape == (R, 1) and np.all(np.isfinite(out)) and np.all(out > 100.0)
        ref = eng.rt_batch(b=b[:1], alpha_slab=slab, T=T, **geom(a))['Tb']
        assert abs(out[0, 0] - ref[0, 0]) < 1e-7
    with pytest.raises(ValueError):
        eng.rt_batch(b=np.zeros((0, 2)), alpha_slab=slab, T=T, **geom(a))
    with pytest.raises(ValueError):
        eng.rt_batch(b=np.zeros((4, 2)), alpha_slab=slab[:10], T=T, **geom(a))
    # NaN / infinite impact parameters miss the planet like |b| >= 1
    out = eng.rt_batch(b=np.array([[np.nan, 0.0], [np.inf, 0.0], [0.0, 0.0]]), alpha_slab=slab, T=T, **geom(a))['Tb']
    assert out[0, 0] == 2.725 and out[1, 0] == 2.725 and out[2, 0] > 100.0


def test_geometry_prefetch_is_only_an_ordering_optimisation(eng):
    """rb_geometry_prefetch: identical Tb with the geometry started ahead of the absorption; a ticket for
    other rays, or one followed by compute_ds (shared buffers), is dropped and the rays are recomputed."""
    a = golden('atm_jupiter.npz')
    C = keymap(a['C_keys'])
    T = a['gas'][C['T']]
    freqs = np.array([2.0, 22.0, 60.0])
    rng = np.random.default_rng(11)
    b = np.ascontiguousarray(rng.uniform(-1.02, 1.02, (6000, 2)))
    g = geom(a)
    radius = np.ascontiguousarray(g.pop('radius'), dtype=np.float64)
    ref = eng.rt_batch(radius=radius, b=b, alpha_slab=_slab(eng, a, freqs), T=T, **g)['Tb'].copy()
    # hit: prefetch, then the absorption call, then the rt call with the same arrays
    eng.geometry_prefetch(radius, g['refr_index'], b, g['Req'], g['Rpol'], g['orientation'], g['gtype'], g['limb'])
    slab = _slab(eng, a, freqs)
    got = eng.rt_batch(radius=radius, b=b, alpha_slab=slab, T=T, **g)['Tb']
    assert np.array_equal(got, ref, equal_nan=True)
    # miss: other rays than the prefetched ones
    eng.geometry_prefetch(radius, g['refr_index'], b, g['Req'], g['Rpol'], g['orientation'], g['gtype'], g['limb'])
    b2 = np.ascontiguousarray(b[::-1])
    got2 = eng.rt_batch(radius=radius, b=b2, alpha_slab=slab, T=T, **g)['Tb']
    assert np.array_equal(got2, ref[::-1], equal_nan=True)
    # dropped: compute_ds in between reuses the segment buffers
    eng.geometry_prefetch(radius, g['refr_index'], b, g['Req'], g['Rpol'], g['orientation'], g['gtype'], g['limb'])
    ds, nseg, _ = eng.compute_ds(radius=radius, b=b[:7], **g)
    assert ds.shape[0] == 7
    got3 = eng.rt_batch(radius=radius, b=b, alpha_slab=slab, T=T, **g)['Tb']
    assert np.array_equal(got3, ref, equal_nan=True)


def test_chunked_copy_out_equals_single_copy(eng):
    """Host-output pipeline: one launch whose CTAs report per-chunk progress while a copy stream moves finished
    chunks out (rotated launch order, wrapped chunk, ragged last tile) == the plain launch + one copy, bit for
    bit, for Tb (f32 and f64) and integrated_W."""
    from radiobear_b200 import _lib
    a = golden('atm_jupiter.npz')
    C = keymap(a['C_keys'])
    T = a['gas'][C['T']]
    slab = _slab(eng, a, np.array([1.5, 9.0, 22.0, 44.0, 95.0]))
    rng = np.random.default_rng(5)
    R = 16384 + 4321                                  # not a multiple of 32; above the pipelining threshold
    b = np.ascontiguousarray(rng.uniform(-1.05, 1.05, (R, 2)))
    ctx = _lib.get_context()
    try:
        ctx.set_rt_chunks(1)
        ref = eng.rt_batch(b=b, alpha_slab=slab, T=T, want_intW=True, **geom(a))
        ref = {k: v.copy() for k, v in ref.items()}
        ref32 = eng.rt_batch(b=b, alpha_slab=slab, T=T, out_f32=True, **geom(a))['Tb'].copy()
        for nch in (0, 3, 7, 16):
            ctx.set_rt_chunks(nch)
            got = eng.rt_batch(b=b, alpha_slab=slab, T=T, want_intW=True, **geom(a))
            assert np.array_equal(got['Tb'], ref['Tb'], equal_nan=True)
            assert np.array_equal(got['integrated_W'], ref['integrated_W'], equal_nan=True)
            got32 = eng.rt_batch(b=b, alpha_slab=slab, T=T, out_f32=True, **geom(a))['Tb']
            assert np.array_equal(got32, ref32, equal_nan=True)
    finally:
        ctx.set_rt_chunks(0)
    assert np.isnan(ref['Tb']).any() and (ref['Tb'] == 2.725).any() and (ref['Tb'] > 100.0).any()


# ---- mixed-precision integration (rb_set_rt_precision(ctx, RB_RT_MIXED)) -----------------------------------------
def _random_disc_points(a, n, seed, rmax=1.02):
    rng = np.random.default_rng(seed)
    th = rng.uniform(0, 2 * np.pi, n)
    rad = np.sqrt(rng.uniform(0, rmax, n))
    q = float(a['Rpol']) / float(a['Req'])
    return np.ascontiguousarray(np.stack([rad * np.cos(th), rad * np.sin(th) * q], axis=1))


def test_mixed_precision_matches_f64_kernel(eng):
    """FP64 optical depth + SFU exponential + FP32 weights against the all-FP64 kernel on the same rays: Tb within
    MIXED_TOL (0.01 K is the bar), integrated_W within 1e-5, NaN rays and off-planet rays at identical pixels,
    tau_cut variants, ragged ray / frequency counts, float32 output."""
    a = golden('atm_jupiter.npz')
    im = golden('image_c4.npz')
    C = keymap(a['C_keys'])
    T = a['gas'][C['T']]
    b = _random_disc_points(a, 3000, 17)
    assert eng.rt_precision() == 'f64'
    worst = 0.0
    for freqs in (im['freqs'], im['freqs'][::3], np.array([22.0])):        # 64 / 22 / 1 frequencies (ragged groups of 8)
        slab = _slab(eng, a, freqs)
        ref = eng.rt_batch(b=b, alpha_slab=slab, T=T, want_intW=True, **geom(a))
        ref = {k: v.copy() for k, v in ref.items()}
        ref_cut5 = eng.rt_batch(b=b, alpha_slab=slab, T=T, tau_cut=5.0, **geom(a))['Tb'].copy()
        eng.set_rt_precision('mixed')
        try:
            assert eng.rt_precision() == 'mixed'
            got = eng.rt_batch(b=b, alpha_slab=slab, T=T, want_intW=True, **geom(a))
            got = {k: v.copy() for k, v in got.items()}
            nocut = eng.rt_batch(b=b, alpha_slab=slab, T=T, tau_cut=0.0, **geom(a))['Tb'].copy()
            cut5 = eng.rt_batch(b=b, alpha_slab=slab, T=T, tau_cut=5.0, **geom(a))['Tb'].copy()
            f32 = eng.rt_batch(b=b, alpha_slab=slab, T=T, out_f32=True, **geom(a))['Tb'].copy()
            ragged = eng.rt_batch(b=b[:1025], alpha_slab=slab, T=T, **geom(a))['Tb'].copy()
        finally:
            eng.set_rt_precision('f64')
        nan = np.isnan(ref['Tb'])
        assert np.array_equal(np.isnan(got['Tb']), nan) and nan.any()
        assert np.array_equal(got['Tb'] == 2.725, ref['Tb'] == 2.725) and (ref['Tb'] == 2.725).any()
        err = np.abs(got['Tb'] - ref['Tb'])[~nan]
        worst = max(worst, float(err.max()))
        assert err.max() < MIXED_TOL
        ok = ~np.isnan(ref['integrated_W']) & (ref['integrated_W'] > 0)
        assert np.max(np.abs(got['integrated_W'][ok] / ref['integrated_W'][ok] - 1.0)) < 1e-5
        assert np.max(np.abs(nocut - got['Tb'])[~nan]) < 1e-6            # e^-50 terms are invisible here as well
        # one tau_cut rule in every kernel (high word of tau against the high word of the cut; the crossing step is
        # integrated, then the ray stops): both kernels stop at the same step also for a small cut
        dcut = np.abs(cut5 - ref_cut5)[~nan]
        assert dcut.max() < MIXED_TOL
        assert np.array_equal(f32[~nan], got['Tb'][~nan].astype(np.float32))
        assert np.array_equal(ragged, got['Tb'][:1025], equal_nan=True)   # results do not depend on the batch
    assert worst < MIXED_TOL


def test_mixed_precision_c4_subset_against_reference(mixed):
    """Config C4 pixels the reference computed (on-disc, NaN ring, off-disc), padded above the 512-ray switch so that
    they run through the mixed-precision rays-major kernel: the north-star bar (0.01 K) against the reference."""
    eng = mixed
    a = golden('atm_jupiter.npz')
    im = golden('image_c4.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, im['freqs'])
    grid = im['grid']
    bsub = np.array([[grid[ix], grid[iy]] for iy, ix in im['pick_iy_ix']])
    pad = _random_disc_points(a, 700, 3)
    res = eng.rt_batch(b=np.concatenate([bsub, pad]), alpha_slab=slab, T=a['gas'][C['T']], **geom(a))['Tb'][:len(bsub)]
    ref = im['tb']
    finite = ~np.isnan(ref).any(axis=1)
    assert np.array_equal(np.isnan(res).any(axis=1), ~finite)
    assert np.max(np.abs(res[finite] - ref[finite])) < MIXED_TOL < TB_TOL


def test_mixed_precision_pipelines_are_bit_identical(mixed):
    """Prefetched geometry, the chunked copy-out pipeline and the plain launch give the same bits in mixed mode."""
    from radiobear_b200 import _lib
    eng = mixed
    a = golden('atm_jupiter.npz')
    C = keymap(a['C_keys'])
    T = a['gas'][C['T']]
    slab = _slab(eng, a, np.array([1.5, 9.0, 22.0, 44.0, 95.0]))
    R = 16384 + 4321
    b = np.ascontiguousarray(np.random.default_rng(5).uniform(-1.05, 1.05, (R, 2)))
    g = geom(a)
    radius = np.ascontiguousarray(g.pop('radius'), dtype=np.float64)
    ctx = _lib.get_context()
    try:
        ctx.set_rt_chunks(1)
        ref = eng.rt_batch(radius=radius, b=b, alpha_slab=slab, T=T, out_f32=True, **g)['Tb'].copy()
        for nch in (0, 7):
            ctx.set_rt_chunks(nch)
            got = eng.rt_batch(radius=radius, b=b, alpha_slab=slab, T=T, out_f32=True, **g)['Tb']
            assert np.array_equal(got, ref, equal_nan=True)
        eng.geometry_prefetch(radius, g['refr_index'], b, g['Req'], g['Rpol'], g['orientation'], g['gtype'], g['limb'])
        got = eng.rt_batch(radius=radius, b=b, alpha_slab=slab, T=T, out_f32=True, **g)['Tb']
        assert np.array_equal(got, ref, equal_nan=True)
        # a geometry prefetched in the other precision has the other slab layout: the switch drops it
        eng.set_rt_precision('f64')
        eng.geometry_prefetch(radius, g['refr_index'], b, g['Req'], g['Rpol'], g['orientation'], g['gtype'], g['limb'])
        eng.set_rt_precision('mixed')
        got = eng.rt_batch(radius=radius, b=b, alpha_slab=slab, T=T, out_f32=True, **g)['Tb']
        assert np.array_equal(got, ref, equal_nan=True)
    finally:
        ctx.set_rt_chunks(0)
    assert np.isnan(ref).any() and (ref == np.float32(2.725)).any() and (ref > 100.0).any()
    with pytest.raises(ValueError):
        eng.set_rt_precision('f16')


def test_ray_fields_r4ds_and_doppler():
    """raypath.compute_ds returns the descriptive fields of the ray like the reference (raypath.py:186-187, 224):
    r4ds (shell radius per step) and doppler, for Jupiter and for tilted Neptune."""
    import os
    from conftest import GOLDEN
    from radiobear_b200 import raypath
    from radiobear_b200.atmosphere import Atmosphere
    g = golden('ray_fields.npz')
    for name, fn in (('jupiter', 'atm_jupiter.npz'), ('neptune', 'atm_neptune.npz')):
        atm = Atmosphere.from_npz(os.path.join(GOLDEN, fn), name)
        assert abs(atm.config.omega_m - float(g['omega_m_' + name])) < 1e-18
        atm.config.vwlat, atm.config.vwdat = list(g['vwlat_' + name]), list(g['vwdat_' + name])   # config.zonal table
        for k, b in enumerate(g['b']):
            ray = raypath.compute_ds(atm, list(b), atm.config.orientation)
            n = int(g['nseg_' + name][k])
            assert len(ray.ds) == n == len(ray.r4ds) == len(ray.doppler) == len(ray.layer4ds)
            assert np.max(np.abs(np.array(ray.r4ds) / g['r4ds_' + name][k, :n] - 1.0)) < 1e-10
            assert np.max(np.abs(np.array(ray.doppler) - g['doppler_' + name][k, :n])) < 1e-12


def _gravity_setup(name, fn):
    import os
    from conftest import GOLDEN
    from radiobear_b200.atmosphere import Atmosphere
    from oracle import ray_oracle as ro
    g = golden('ray_fields.npz')
    atm = Atmosphere.from_npz(os.path.join(GOLDEN, fn), name)
    atm.config.vwlat, atm.config.vwdat = list(g['vwlat_' + name]), list(g['vwdat_' + name])   # config.zonal table
    cfg = atm.config
    req, GM = atm.property[cfg.LP['R']], atm.property[cfg.LP['GM']]
    tab = ro.GeoidTable(req, GM, cfg.Jn, cfg.RJ, cfg.omega_m, cfg.vwlat, cfg.vwdat, max_abs_lat=60.0)
    return atm, tab


@pytest.mark.parametrize('name,fn', [('jupiter', 'atm_jupiter.npz'), ('neptune', 'atm_neptune.npz')])
def test_gravity_geoid_rays_vs_oracle(eng, name, fn):
    """gtype='gravity' (Shape._calcGeoid / _gravity, shape.py:141-221): the device's shape table + vector ray march
    against the oracle's restatement -- path lengths, segment counts, the NaN tail of a limb ray, Ray.r4ds -- and Tb
    through the integration.  (tests/test_oracle_golden.py pins the oracle to rays the reference itself computed.)"""
    from radiobear_b200 import raypath
    from oracle import ray_oracle as ro
    atm, tab = _gravity_setup(name, fn)
    cfg = atm.config
    req, nr = atm.property[cfg.LP['R']], atm.property[cfg.LP['N']]
    S = len(req) - 1
    pts = [[0.0, 0.0], [0.3, 0.2], [0.6, -0.4], [-0.85, 0.3], [0.2, 0.7], [0.99, 0.05], [1.2, 0.1]]
    args = raypath._geometry_args(atm, cfg.orientation, 'gravity')
    ds, nseg, aspect = eng.compute_ds(b=np.array(pts), **args)
    fields = eng.compute_ray_fields(b=np.array(pts), **args)
    n_nan = 0
    for k, b in enumerate(pts):
        ref = ro.compute_ds(req, nr, b, cfg.Req, cfg.Rpol, cfg.orientation, 'gravity', 'shape', gravity=dict(table=tab))
        if ref['ds'] is None:
            assert nseg[k] == -1
            continue
        n = len(ref['ds'])
        assert nseg[k] == n, (b, nseg[k], n)
        bad = np.isnan(ref['ds'])
        assert np.array_equal(np.isnan(ds[k, :n]), bad)
        n_nan += int(bad.any())
        assert np.max(np.abs(ds[k, :n][~bad] / ref['ds'][~bad] - 1.0)) < 1e-8, b
        ok = ~np.isnan(ref['r4ds'])
        assert np.max(np.abs(fields[k, 0, :n][ok] / ref['r4ds'][ok] - 1.0)) < 1e-11
    assert n_nan >= 1                                         # the grazing ray goes below its tangent shell
    if name == 'jupiter':
        # the two rays the reference itself computed (tests/golden/make_golden.py section gravity)
        gg = golden('gravity.npz')
        gds, gn, _ = eng.compute_ds(b=gg['b'], **args)
        gf = eng.compute_ray_fields(b=gg['b'], **args)
        for k in range(len(gg['b'])):
            n = int(gg['nseg'][k])
            assert gn[k] == n
            assert np.max(np.abs(gds[k, :n] / gg['ds'][k, :n] - 1.0)) < 1e-8
            assert np.max(np.abs(gf[k, 0, :n] / gg['r4ds'][k, :n] - 1.0)) < 1e-11
    # the shape differs from the ellipse by kilometres; the integration takes the gravity segments as they are
    e_ds, e_n, _ = eng.compute_ds(b=np.array(pts[:3]), **raypath._geometry_args(atm, cfg.orientation, 'ellipse'))
    assert np.max(np.abs(ds[1, :e_n[1]] / e_ds[1, :e_n[1]] - 1.0)) > 1e-4
    a = golden(fn)
    freqs = [2.0, 10.0, 30.0]
    slab = _slab(eng, a, freqs) if name == 'jupiter' else None
    if slab is not None:
        C = keymap(a['C_keys'])
        T = a['gas'][C['T']]
        got = eng.rt_batch(b=np.array(pts[:3]), alpha_slab=slab, T=T, **args)['Tb']
        want = eng.rt_integrate(ds[:3], nseg[:3], slab, T)
        assert np.max(np.abs(got - want)) < 1e-9
        ell = eng.rt_batch(b=np.array(pts[:3]), alpha_slab=slab, T=T, **raypath._geometry_args(atm, cfg.orientation, 'ellipse'))['Tb']
        assert 1e-4 < np.max(np.abs(got - ell)) < 5.0
