"""GPU: parity at the sizes SURVEY 8d states, against vectors the unmodified reference produced
(tests/golden/make_golden.py sections image_full, ring, c3_full, c5_saturn), and the work decompositions of the
FP64 ray integration (one / two frequencies per thread, plain / compacted ray order) against each other.

Bars: Tb within 0.01 K of the reference (asserted: 1e-4 K), NaN <-> NaN at identical pixels, off-planet pixels
exactly 2.725 K, alpha within 1e-6 relative (asserted: 1e-9), decompositions bit-identical."""
import numpy as np
import pytest

from conftest import golden, keymap, relerr, formalisms_of, TRUNC

pytestmark = pytest.mark.gpu
TB_TOL = 0.01


@pytest.fixture(scope='module')
def eng():
    from radiobear_b200 import engine
    before = engine.rt_precision()
    engine.set_rt_precision('f64')
    engine.set_rt_tuning(-1, True)
    yield engine
    engine.set_rt_tuning(-1, True)
    engine.set_rt_precision(before)


def geom(a):
    LP = keymap(a['LP_keys'])
    return dict(radius=a['property'][LP['R']], refr_index=a['property'][LP['N']], Req=float(a['Req']),
                Rpol=float(a['Rpol']), orientation=a['orientation'], gtype=str(a['gtype']), limb=str(a['limb']))


def _slab(eng, a, freqs):
    C = keymap(a['C_keys'])
    return eng.alpha_layers(freqs, a['gas'][C['T']], a['gas'][C['P']], a['gas'], C, formalisms=formalisms_of(a),
                            other_dicts={'h2': {'h2state': 'e'}, 'co': {'coshape': 'voigt'}}, truncate_strength=TRUNC)


def _image_points(grid):
    n = len(grid)
    return np.ascontiguousarray(np.stack([np.tile(grid, n), np.repeat(grid, n)], axis=1))   # x fastest (set_utils.py:65-77)


@pytest.fixture(scope='module')
def c4(eng):
    """The full C4 cube (601 x 601 pixels x 64 freqs, FP64 outputs) through the default path: compacted ray list,
    two frequencies per thread."""
    a = golden('atm_jupiter.npz')
    im = golden('image_c4_full.npz')
    C = keymap(a['C_keys'])
    slab = _slab(eng, a, im['freqs'])
    b = _image_points(im['grid'])
    out = eng.rt_batch(b=b, alpha_slab=slab, T=a['gas'][C['T']], **geom(a))['Tb'].copy()
    return dict(a=a, im=im, slab=slab, b=b, T=a['gas'][C['T']], cube=out.reshape(601, 601, 64))


def test_c4_full_size_golden_pixels(eng, c4):
    """288 on-disc + 64 limb-ring + 16 off-disc pixels the reference computed (Brightness.single), read out of the
    full cube (rays-major kernels) and recomputed as a small batch (lanes = frequency kernel)."""
    im, cube = c4['im'], c4['cube']
    ref = im['tb']
    pick = im['pick_iy_ix']
    got = np.array([cube[iy, ix] for iy, ix in pick])
    nan_ref = np.isnan(ref).any(axis=1)
    assert (~nan_ref).sum() >= 256 + 16 and nan_ref.sum() >= 8
    assert np.array_equal(np.isnan(got), np.isnan(ref))                       # NaN <-> NaN, every frequency
    err = np.abs(got - ref)[~np.isnan(ref)]
    assert err.max() < 1e-4 < TB_TOL
    off = (ref == 2.725).all(axis=1)
    assert off.sum() >= 16 and np.all(got[off] == 2.725)
    bsub = np.array([[im['grid'][ix], im['grid'][iy]] for iy, ix in pick])
    small = eng.rt_batch(b=bsub, alpha_slab=c4['slab'], T=c4['T'], **geom(c4['a']))['Tb']
    assert np.array_equal(np.isnan(small), np.isnan(ref))
    assert np.nanmax(np.abs(small - ref)) < 1e-4
    assert np.nanmax(np.abs(small - got)) < 1e-7                               # the two kernels agree far below the bar


def test_limb_ring_quadrant_mask(eng, c4):
    """Every pixel with 0.93 < r < 1.02 of one quadrant (5177 pixels; raypath.compute_ds of the reference): hit / miss,
    number of segments and NaN-ness of the segments Brightness.single uses -- through the compute_ds entry point and
    as the NaN / sky mask of the image cube, in all four mirror images of the quadrant."""
    ring = golden('ring_quadrant.npz')
    a = c4['a']
    grid = ring['grid']
    iy, ix = ring['iy_ix'][:, 0], ring['iy_ix'][:, 1]
    b = np.stack([grid[ix], grid[iy]], axis=1)
    ds, nseg, _ = eng.compute_ds(b=b, **geom(a))
    ref_n, ref_nan = ring['nseg'], ring['used_nan'].astype(bool)
    assert np.array_equal(nseg, ref_n), 'hit / miss or segment count differs at {}'.format(b[nseg != ref_n][:8])
    hit = ref_n > 0
    used_nan = np.zeros(len(b), dtype=bool)
    for k in np.nonzero(hit)[0]:
        used_nan[k] = np.isnan(ds[k, :ref_n[k] - 1]).any()
    diff = np.nonzero(used_nan != ref_nan)[0]
    assert len(diff) == 0, 'NaN classification differs at pixels (iy, ix) {}'.format(ring['iy_ix'][diff][:16])
    # the smooth part of the ring's geometry: total path length of the finite rays
    fin = hit & ~np.isnan(ring['nansum_ds']) & (ring['first_nan'] < 0)
    tot = np.array([ds[k, :ref_n[k]].sum() for k in np.nonzero(fin)[0]])
    assert np.max(np.abs(tot / ring['nansum_ds'][fin] - 1.0)) < 1e-9
    # the same classification as seen in the image cube; x -> -x, y -> -y mirror images (orientation 0, 0)
    cube = c4['cube']
    n = cube.shape[0]
    for my, mx in ((iy, ix), (n - 1 - iy, ix), (iy, n - 1 - ix), (n - 1 - iy, n - 1 - ix)):
        px = cube[my, mx]
        assert np.array_equal(np.isnan(px).all(axis=1), ref_nan)
        assert np.array_equal(np.isnan(px).any(axis=1), ref_nan)
        assert np.array_equal((px == 2.725).all(axis=1), ~hit)


def test_c3_full_limb_profile(eng):
    """Config C3 complete: the 100 rays of b = '0.0:1.0:0.01<0' x 50 frequencies, as the reference's Planet.run
    computes them; once as given (small-batch kernel) and once tiled above the 512-ray switch (rays-major kernels)."""
    from radiobear_b200 import set_utils
    a = golden('atm_jupiter.npz')
    c3 = golden('c3_full.npz')
    C = keymap(a['C_keys'])
    rv = set_utils.set_b('0.0:1.0:0.01<0', [1, 1], Rpol=float(a['Rpol']), Req=float(a['Req']))
    assert str(c3['data_type']) == rv.data_type == 'profile'
    assert np.array_equal(np.array(rv.b), c3['b']) and len(rv.b) == 100       # the request string parses to the same rays
    slab = _slab(eng, a, c3['freqs'])
    ref = c3['tb']
    assert np.isnan(ref).any(axis=1).sum() >= 1
    res = eng.rt_batch(b=c3['b'], alpha_slab=slab, T=a['gas'][C['T']], **geom(a))['Tb']
    assert np.array_equal(np.isnan(res), np.isnan(ref))
    assert np.nanmax(np.abs(res - ref)) < 1e-4 < TB_TOL
    tiled = np.ascontiguousarray(np.tile(c3['b'], (6, 1)))
    big = eng.rt_batch(b=tiled, alpha_slab=slab, T=a['gas'][C['T']], **geom(a))['Tb'].reshape(6, 100, -1)
    for k in range(6):
        assert np.array_equal(np.isnan(big[k]), np.isnan(ref))
        assert np.nanmax(np.abs(big[k] - ref)) < 1e-4


def test_c5_saturn_4096_layers_against_reference(eng):
    """Config C5 on its concrete input: Planet('saturn', regridType=4096) x np.linspace(1, 100, 4096) x nh3_dbs_sjs
    (the FPT = 2 absorption path at L = F = 4096) against 96 layers of the reference's own plugin."""
    g = golden('c5_saturn.npz')
    C = keymap(g['C_keys'])
    gas = np.ascontiguousarray(g['gas'])
    assert gas.shape[1] == 4096 and len(g['freqs']) == 4096
    slab = eng.alpha_layers(g['freqs'], gas[C['T']], gas[C['P']], gas, C, formalisms=[('nh3', 'nh3_dbs_sjs')])
    assert slab.shape == (4096, 4096)
    P = gas[C['P']][g['layers']]
    assert (P < 400).any() and ((P >= 400) & (P <= 2000)).any() and (P > 2000).any()   # all three branches of the blend
    err = relerr(slab[g['layers']], g['alpha'])
    assert err.max() < 1e-9
    assert np.all(np.isfinite(slab)) and slab.min() >= 0.0


def test_decompositions_agree(eng, c4):
    """One / two frequencies per thread x plain / compacted ray order on the full C4 cube.  Compaction changes no
    bit (same kernel, same arithmetic per ray; NaN ring and sky included).  The two kernels run the same operations
    per step but hand a ray from the small-tau polynomial to the table exponential at different segments (groups of
    4 vs 2 segments, and a pair leaves the polynomial together): they agree to 1e-9 K, five orders inside the bar."""
    a, b, slab, T = c4['a'], c4['b'], c4['slab'], c4['T']
    ref = c4['cube'].reshape(-1, 64)
    res = {}
    try:
        for pairs, compact in ((0, False), (1, False), (0, True), (1, True)):
            eng.set_rt_tuning(pairs, compact)
            got = eng.rt_batch(b=b, alpha_slab=slab, T=T, want_intW=True, **geom(a))
            res[(pairs, compact)] = (got['Tb'].copy(), got['integrated_W'].copy())
        for pairs in (0, 1):
            assert np.array_equal(res[(pairs, False)][0], res[(pairs, True)][0], equal_nan=True), pairs
            assert np.array_equal(res[(pairs, False)][1], res[(pairs, True)][1], equal_nan=True), pairs
        one, two = res[(0, True)], res[(1, True)]
        assert np.array_equal(np.isnan(one[0]), np.isnan(two[0])) and np.array_equal(one[0] == 2.725, two[0] == 2.725)
        assert np.nanmax(np.abs(one[0] - two[0])) < 1e-9
        ok = ~np.isnan(one[1]) & (one[1] > 0)
        assert np.max(np.abs(two[1][ok] / one[1][ok] - 1.0)) < 1e-12
        assert np.array_equal(res[(1, True)][0], ref, equal_nan=True)          # the default path is the pair kernel here
        # ragged frequency counts through the pair kernel (ghost partner, partial groups)
        for nf in (2, 3, 17, 33):
            eng.set_rt_tuning(0, True)
            o1 = eng.rt_batch(b=b[180000:184000], alpha_slab=np.ascontiguousarray(slab[:, :nf]), T=T, **geom(a))['Tb'].copy()
            eng.set_rt_tuning(1, True)
            o2 = eng.rt_batch(b=b[180000:184000], alpha_slab=np.ascontiguousarray(slab[:, :nf]), T=T, **geom(a))['Tb']
            assert o2.shape == (4000, nf) and np.array_equal(np.isnan(o1), np.isnan(o2))
            assert np.nanmax(np.abs(o1 - o2)) < 1e-9, nf
        # small tau_cut: both kernels stop at the same step (one rule: tau >= tau_cut by the high words)
        for cut in (5.0, 0.3):
            eng.set_rt_tuning(0, False)
            o1 = eng.rt_batch(b=b[180000:200000], alpha_slab=slab, T=T, tau_cut=cut, **geom(a))['Tb'].copy()
            eng.set_rt_tuning(1, True)
            o2 = eng.rt_batch(b=b[180000:200000], alpha_slab=slab, T=T, tau_cut=cut, **geom(a))['Tb']
            assert np.nanmax(np.abs(o1 - o2)) < 1e-9, cut
    finally:
        eng.set_rt_tuning(-1, True)
    f32 = eng.rt_batch(b=b, alpha_slab=slab, T=T, out_f32=True, **geom(a))['Tb']
    assert np.array_equal(f32, ref.astype(np.float32), equal_nan=True)


def test_widely_separated_frequency_pair(eng, c4):
    """A pair whose members stop hundreds of layers apart (1 GHz next to 100 GHz): the frequency left alone finishes
    on the single-segment path; same values as the one-frequency kernel (1e-9 K)."""
    a, b, T = c4['a'], c4['b'], c4['T']
    slab = _slab(eng, a, np.array([1.0, 100.0, 3.0, 60.0]))
    try:
        eng.set_rt_tuning(0, True)
        one = eng.rt_batch(b=b[170000:190000], alpha_slab=slab, T=T, **geom(a))['Tb'].copy()
        eng.set_rt_tuning(1, True)
        two = eng.rt_batch(b=b[170000:190000], alpha_slab=slab, T=T, **geom(a))['Tb']
    finally:
        eng.set_rt_tuning(-1, True)
    assert np.array_equal(np.isnan(one), np.isnan(two)) and np.isfinite(one).any()
    assert np.nanmax(np.abs(one - two)) < 1e-9


def test_degenerate_ray_lists(eng, c4):
    """The compacted, ordered ray list at its edges: a request with no hit at all, one hit among sky, exactly one and
    exactly two super-blocks of rays, 513 rays, and 40 001 random points through the host copy-out pipeline -- against
    the plain ray order / the one-frequency kernel."""
    a, slab, T = c4['a'], c4['slab'], c4['T']
    rng = np.random.default_rng(5)
    sky = np.stack([rng.uniform(1.05, 1.4, 1000), rng.uniform(-0.2, 0.2, 1000)], axis=1)
    out = eng.rt_batch(b=sky, alpha_slab=slab, T=T, **geom(a))['Tb']
    assert out.shape == (1000, 64) and np.all(out == 2.725)
    one = sky.copy()
    one[777] = [0.25, 0.5]
    out = eng.rt_batch(b=one, alpha_slab=slab, T=T, **geom(a))['Tb']
    assert np.all(np.delete(out, 777, axis=0) == 2.725) and np.all(out[777] > 50.0)
    th, rad = rng.uniform(0, 2 * np.pi, 40001), np.sqrt(rng.uniform(0, 1.1, 40001))
    pts = np.stack([rad * np.cos(th), rad * np.sin(th) * 0.935], axis=1)
    try:
        eng.set_rt_tuning(0, False)
        ref = eng.rt_batch(b=pts, alpha_slab=slab, T=T, **geom(a))['Tb'].copy()
    finally:
        eng.set_rt_tuning(-1, True)
    for n in (513, 8192, 16384, 40001):
        got = eng.rt_batch(b=np.ascontiguousarray(pts[:n]), alpha_slab=slab, T=T, **geom(a))['Tb']
        assert got.shape == (n, 64) and np.array_equal(np.isnan(got), np.isnan(ref[:n]))
        assert np.array_equal(got == 2.725, ref[:n] == 2.725)
        assert np.nanmax(np.abs(got - ref[:n])) < 1e-9, n
    f32 = eng.rt_batch(b=pts, alpha_slab=slab, T=T, out_f32=True, **geom(a))['Tb']
    big = eng.rt_batch(b=pts, alpha_slab=slab, T=T, **geom(a))['Tb']
    assert np.array_equal(f32, big.astype(np.float32), equal_nan=True)


def test_integration_follows_a_running_trace(eng, c4):
    """rb_set_rt_stream_geometry: the pair kernel starts behind a prefetched trace that still runs and waits chunk by
    chunk on the trace's progress counters.  Same bits as the trace-then-integrate order: image rows with sky, NaN limb
    ring and the disc centre; ragged frequency counts; every consumer that cannot follow (one frequency: lanes =
    frequency kernel, disc average, a request that misses the ticket) waits for the end of the trace."""
    a, b, slab, T = c4['a'], c4['b'], c4['slab'], c4['T']
    g = geom(a)
    radius = np.ascontiguousarray(g.pop('radius'), dtype=np.float64)
    pre = lambda pts: eng.geometry_prefetch(radius, g['refr_index'], pts, g['Req'], g['Rpol'], g['orientation'], g['gtype'], g['limb'])
    rows = np.ascontiguousarray(b[601 * 270:601 * 330])           # 60 image rows through the centre: 36 060 rays
    edge = np.ascontiguousarray(b[601 * 108:601 * 136])           # sky rows, then rows that graze the limb
    try:
        res = {}
        for mode in (0, 1):
            eng.set_rt_stream_geometry(mode)
            out = []
            for pts in (rows, edge):
                pre(pts)
                got = eng.rt_batch(radius=radius, b=pts, alpha_slab=slab, T=T, want_intW=True, **g)
                out += [got['Tb'].copy(), got['integrated_W'].copy()]
            for nf in (1, 2, 3, 17):
                pre(edge)
                out.append(eng.rt_batch(radius=radius, b=edge, alpha_slab=np.ascontiguousarray(slab[:, :nf]), T=T, **g)['Tb'].copy())
            pre(rows)
            out.append(eng.rt_batch(radius=radius, b=rows, alpha_slab=slab, T=T, out_f32=True, **g)['Tb'].copy())
            pre(rows)                                              # ticket missed: other points
            out.append(eng.rt_batch(radius=radius, b=edge, alpha_slab=slab, T=T, **g)['Tb'].copy())
            pre(rows)                                              # followed twice in a row (counters are reset per trace)
            out.append(eng.rt_batch(radius=radius, b=rows, alpha_slab=slab, T=T, tau_cut=0.3, **g)['Tb'].copy())
            res[mode] = out
        for x, y in zip(res[0], res[1]):
            assert x.shape == y.shape and np.array_equal(x, y, equal_nan=True)
        assert np.isnan(res[1][2]).any() and (res[1][2] == 2.725).any() and np.nanmax(res[1][0]) > 100.0
        ref = c4['cube'].reshape(-1, 64)
        assert np.array_equal(res[1][0], ref[601 * 270:601 * 330], equal_nan=True)
    finally:
        eng.set_rt_stream_geometry(-1)


def test_launch_order_changes_no_bit(eng, c4):
    """The integration's launch order (rb_launch_integrate: frequency group by frequency group inside parts of the
    tile list; device-resident outputs get L2-sized parts whose number is chosen on the device, RB_RT_PARTS overrides,
    0 = the frequency groups of a tile adjacent) decides which CTAs run together, not what they compute: the full C4
    cube through the device-pointer entry is bit-identical for every order, and equals the host-output cube (12 parts
    under the copy-out pipeline)."""
    import os
    import torch
    a, b, slab, T = c4['a'], c4['b'], c4['slab'], c4['T']
    g = geom(a)
    dev = torch.device('cuda', 0)
    t64 = dict(dtype=torch.float64, device=dev)
    radius_t = torch.tensor(np.ascontiguousarray(g['radius'], dtype=np.float64), **t64)
    b_t, slab_t, T_t = torch.tensor(b, **t64), torch.tensor(np.ascontiguousarray(slab), **t64), torch.tensor(np.ascontiguousarray(T), **t64)
    out_t = torch.empty((len(b), slab.shape[1]), **t64)
    n = g['refr_index']
    before = os.environ.pop('RB_RT_PARTS', None)
    try:
        cubes = {}
        for parts in (None, '0', '1', '5', '37'):
            if parts is None:
                os.environ.pop('RB_RT_PARTS', None)
            else:
                os.environ['RB_RT_PARTS'] = parts
            out_t.fill_(-1.0)
            eng.rt_batch_dev(radius_t, float(n[0]), float(n[1]), b_t, slab_t, T_t, g['Req'], g['Rpol'],
                             [float(g['orientation'][0]), float(g['orientation'][1])], g['gtype'], g['limb'],
                             out_f32=False, tau_cut=eng.TAU_CUT, out=out_t)
            torch.cuda.synchronize()
            cubes[parts] = out_t.cpu().numpy()
        ref = c4['cube'].reshape(-1, 64)
        for parts, cube in cubes.items():
            assert np.array_equal(cube, ref, equal_nan=True), parts
    finally:
        os.environ.pop('RB_RT_PARTS', None)
        if before is not None:
            os.environ['RB_RT_PARTS'] = before


def test_sky_fill_stream_changes_no_bit(eng, c4):
    """The sky pixels of a compacted launch are written by rt_fill_miss on the context stream (RB_FILL_STREAM=0), a side
    stream (1) or a side stream of the highest priority (2, the default: the fill must not queue behind the integration
    CTAs, the copy-out stream waits for it).  Host-output pipeline and single-launch path: same cube, sky exactly 2.725 K."""
    import os
    a, b, slab, T = c4['a'], c4['b'], c4['slab'], c4['T']
    ref = c4['cube'].reshape(-1, 64)
    before = os.environ.pop('RB_FILL_STREAM', None)
    try:
        for mode in ('0', '1', '2'):
            os.environ['RB_FILL_STREAM'] = mode
            got = eng.rt_batch(b=b, alpha_slab=slab, T=T, **geom(a))['Tb']
            assert np.array_equal(got, ref, equal_nan=True), mode
            part = eng.rt_batch(b=b[:8000], alpha_slab=slab, T=T, **geom(a))['Tb']      # below the chunked pipeline: one launch
            assert np.array_equal(part, ref[:8000], equal_nan=True) and (part == 2.725).all(), mode
    finally:
        os.environ.pop('RB_FILL_STREAM', None)
        if before is not None:
            os.environ['RB_FILL_STREAM'] = before
