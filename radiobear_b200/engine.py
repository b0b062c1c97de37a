"""Host-side driver of the sm_100a kernels: builds the C-ABI descriptors and calls the library.

This is the only module that talks to libradiobear_b200.so; the reference-shaped classes
(alpha.Alpha, brightness.Brightness, raypath.compute_ds, constituents/<gas>/<formalism>.alpha)
are thin layers over the three functions here.  numpy in / numpy out; `*_dev` variants take and
return torch CUDA tensors for device-resident pipelines (Planet.run, bench.py).
"""
import ctypes as C

import numpy as np

from . import _lib
from . import catalogs
from ._lib import (AlphaDesc, GeometryDesc, RtDesc, FORMALISM_IDS, GAS_ORDER, CLOUD_ORDER, CLOUD_FLAG_KEYS,
                   RB_MAX_CONSTITUENTS, f64, ptr)

T_CMB = 2.725


from .hostmem import PinnedPool, pinned_pool  # noqa: E402,F401



def set_rt_precision(precision, ctx=None):
    """Arithmetic of the batched ray integration: 'f64' (every operation in FP64) or 'mixed' (FP64 optical
    depth, SFU exponential, FP32 weights; |dTb| ~ one float32 ulp of Tb -- include/radiobear_b200.h).  The
    library starts in the mode RB_RT_PRECISION names (default 'f64')."""
    (ctx or _lib.get_context()).set_rt_precision(precision)


def rt_precision(ctx=None):
    return (ctx or _lib.get_context()).rt_precision()


def set_rt_tuning(pairs=-1, compact=True, ctx=None):
    """Work decomposition of the batched FP64 ray integration (rb_set_rt_tuning): two frequencies per thread
    (-1 automatic, 0, 1) and ray compaction.  Results do not depend on it."""
    (ctx or _lib.get_context()).set_rt_tuning(pairs, compact)


def set_rt_stream_geometry(mode=-1, ctx=None):
    """Let the integration of rt_batch follow a prefetched ray trace while it is still running
    (rb_set_rt_stream_geometry): 1 always, 0 never, -1 automatic (small requests).  Results do not depend on it."""
    (ctx or _lib.get_context()).set_rt_stream_geometry(mode)


UNITS = {'invcm': 0, 'dBperkm': 1}
COSHAPE = {'voigt': 0, 'vvw': 1, 'diff': 2}


def _cloud_flags(other):
    flags = 0
    for bit, key in enumerate(CLOUD_FLAG_KEYS):
        v = other.get(key, 0.0) if other else 0.0
        try:
            if v is not None and float(v) > 0.0:
                flags |= (1 << bit)
        except (TypeError, ValueError):
            pass
    return flags


def build_alpha_desc(formalisms, L, F, gas_rows, gas_dict, cloud_rows, cloud_dict, other_dicts, units):
    """formalisms: list of (constituent, formalism-name) in call order (sorted constituents)."""
    if len(formalisms) == 0 or len(formalisms) > RB_MAX_CONSTITUENTS:
        raise ValueError('need 1..{} constituents'.format(RB_MAX_CONSTITUENTS))
    d = AlphaDesc()
    d.n_layers, d.n_freqs, d.n_constituents = L, F, len(formalisms)
    other_dicts = other_dicts or {}
    h2state, coshape, cflags = 0, 0, 0
    for i, (c, name) in enumerate(formalisms):
        if name not in FORMALISM_IDS:
            raise NotImplementedError('formalism {} is not built in radiobear_b200'.format(name))
        d.formalism[i] = FORMALISM_IDS[name]
        od = other_dicts.get(c, {}) or {}
        if name in ('h2_jj_ddb', 'h2_orton'):
            st = od.get('h2state', 'e')
            if st not in ('e', 'n'):
                raise ValueError('INVALID H2STATE {!r}'.format(st))      # h2_jj_ddb.py:28-30 prints and returns 0
            h2state = 0 if st == 'e' else 1
        if name == 'co_ddb':
            coshape = COSHAPE.get(od.get('coshape', 'voigt'), 3)
        if name == 'clouds_idp':
            cflags = _cloud_flags(od)
    d.h2state, d.coshape, d.cloud_flags = h2state, coshape, cflags
    d.gas_rows = gas_rows
    for i, g in enumerate(GAS_ORDER):
        d.gas_col[i] = int(gas_dict[g]) if (gas_dict and g in gas_dict) else -1
    d.cloud_rows = cloud_rows
    for i, g in enumerate(CLOUD_ORDER):
        d.cloud_col[i] = int(cloud_dict[g]) if (cloud_dict and g in cloud_dict) else -1
    if units not in UNITS:
        raise ValueError("units must be 'invcm' or 'dBperkm'")
    d.units = UNITS[units]
    return d


def scale_matrix(scale, ordered, L):
    """alpha.py:235-259 + 151-192: turn the user's scale (number / per-layer list / dict by constituent)
    into a [C][L] matrix, or None for 'no scaling'."""
    ordered = [str(c) for c in ordered]
    C_ = len(ordered)
    if isinstance(scale, dict):
        for k, v in scale.items():
            if k not in ordered:
                raise ValueError("{} not found as constituent for alpha".format(k))
            if len(v) != L:
                raise ValueError("Incorrect scale for {}:  N {} vs {}".format(k, len(v), L))
        m = np.ones((C_, L))
        for k, v in scale.items():
            m[ordered.index(k)] = np.asarray(v, dtype=np.float64)
        return m
    if isinstance(scale, (list, np.ndarray)):
        if len(scale) != L:
            raise ValueError("Incorrect number of scale layers.")
        return np.tile(np.asarray(scale, dtype=np.float64), (C_, 1))
    if isinstance(scale, bool) or scale is None:
        return None
    try:
        s = float(scale)
    except (TypeError, ValueError):
        return None
    return None if s == 1.0 else np.full((C_, L), s)


def prepare_catalogs(ctx, formalisms, truncate_strength=None, truncate_freq=None, freqs=None, other_dicts=None):
    truncate_strength = truncate_strength or {}
    truncate_freq = truncate_freq or {}
    for c, name in formalisms:
        if name not in catalogs.FORMALISM_CATALOGS:
            raise NotImplementedError('formalism {} is not built in radiobear_b200'.format(name))
        catalogs.upload(ctx, name, truncate_strength.get(c), truncate_freq.get(c), freqs=freqs,
                        other=(other_dicts or {}).get(c))


class ResidentSlab:
    """Handle of an absorption slab [L][F] (and possibly the per-constituent cube [L][F][C]) living in the context's
    device buffers (rb_alpha_layers_resident / rb_alpha_rescale_resident).  `rt_batch(alpha_slab=handle)` integrates
    from it without any copy; `fetch()` brings it to the host on demand.  A later resident computation overwrites the
    buffers: the generation numbers tell whether this handle's data is still there."""

    def __init__(self, ctx, L, F, slab_gen, cube_shape=None, cube_gen=0):
        self.ctx, self.shape, self.slab_gen = ctx, (int(L), int(F)), int(slab_gen)
        self.cube_shape, self.cube_gen = cube_shape, int(cube_gen)

    def _gens(self):
        sg, cg = C.c_uint64(0), C.c_uint64(0)
        self.ctx.check(self.ctx.lib.rb_alpha_resident_info(self.ctx.h, None, C.byref(sg), None, C.byref(cg)))
        return int(sg.value), int(cg.value)

    def valid(self):
        return self.slab_gen != 0 and self._gens()[0] == self.slab_gen

    def cube_valid(self):
        return self.cube_gen != 0 and self._gens()[1] == self.cube_gen

    def fetch(self):
        if not self.valid():
            raise RuntimeError('the resident absorption slab has been overwritten')
        out = np.empty(self.shape)
        self.ctx.check(self.ctx.lib.rb_alpha_fetch(self.ctx.h, ptr(out), None))
        return out

    def fetch_cube(self):
        if not self.cube_valid():
            raise RuntimeError('the resident absorption cube has been overwritten')
        out = np.empty(self.cube_shape)
        self.ctx.check(self.ctx.lib.rb_alpha_fetch(self.ctx.h, None, ptr(out)))
        return out


_resident_owner = None      # weak reference to the object whose (not yet fetched) slab sits in the resident buffer


def _claim_resident(owner):
    """Before the resident slab is overwritten: let its current owner copy it to the host if it still needs it."""
    import weakref
    global _resident_owner
    prev = _resident_owner() if _resident_owner is not None else None
    if prev is not None and prev is not owner and hasattr(prev, 'materialize'):
        prev.materialize()
    _resident_owner = weakref.ref(owner) if owner is not None else None


def alpha_layers_resident(freqs, T, P, gas, gas_dict, cloud=None, cloud_dict=None, formalisms=(), other_dicts=None,
                          units='invcm', scale=None, keep_cube=False, truncate_strength=None, truncate_freq=None,
                          owner=None, ctx=None):
    """alpha_layers whose results stay on the device: returns a ResidentSlab handle, does not synchronise."""
    return alpha_layers(freqs, T, P, gas, gas_dict, cloud, cloud_dict, formalisms, other_dicts, units, scale, keep_cube,
                        truncate_strength, truncate_freq, ctx, _resident=(owner,))


def alpha_rescale_resident(handle, scale_mat=None, owner=None):
    """Scale-sum of the resident cube of `handle` into the resident slab (the retrieval inner loop: get_alpha='memory'
    with a new `scale`, alpha.py:151-192) -> new ResidentSlab handle sharing the cube."""
    if not handle.cube_valid():
        raise RuntimeError('the resident absorption cube has been overwritten')
    ctx = handle.ctx
    _claim_resident(owner)
    sm = None if scale_mat is None else f64(scale_mat)
    if sm is not None and sm.shape != (handle.cube_shape[2], handle.cube_shape[0]):
        raise ValueError('scale matrix must be [C][L]')
    gen = C.c_uint64(0)
    ctx.check(ctx.lib.rb_alpha_rescale_resident(ctx.h, ptr(sm), C.byref(gen)))
    return ResidentSlab(ctx, handle.cube_shape[0], handle.cube_shape[1], gen.value, handle.cube_shape, handle.cube_gen)


def alpha_layers(freqs, T, P, gas, gas_dict, cloud=None, cloud_dict=None, formalisms=(), other_dicts=None,
                 units='invcm', scale=None, want_cube=False, truncate_strength=None, truncate_freq=None, ctx=None,
                 _resident=None):
    """Total absorption for every (layer, freq) -> slab[L][F] (+ cube[L][F][C]).  Host arrays.

    Replaces the layer loop of Alpha.get_layers (alpha.py:298-300) and the plugin calls under it.
    """
    freqs, T, P = f64(np.atleast_1d(freqs)), f64(np.atleast_1d(T)), f64(np.atleast_1d(P))
    gas = f64(gas)
    if gas.ndim == 1:
        gas = np.ascontiguousarray(gas[:, None])
    # freqs[L][F]: every layer at its own frequencies (Doppler-shifted absorption, brightness.py:80-92)
    per_layer = freqs.ndim == 2
    L, F = T.shape[0], freqs.shape[-1]
    if per_layer and freqs.shape[0] != L:
        raise ValueError('per-layer frequencies: freqs must be [L][F]')
    if gas.shape[1] != L or P.shape[0] != L:
        raise ValueError('T, P and gas disagree on the number of layers')
    if cloud is not None:
        cloud = f64(cloud)
        if cloud.ndim == 1:
            cloud = np.ascontiguousarray(cloud[:, None])
        if cloud.shape[1] != L:
            # regridType none: the cloud file may hold fewer layers than the gas file (the reference fails with an
            # IndexError in the clouds plugin); the kernel indexes cloud[c][layer] for every layer
            raise ValueError('cloud has {} layers, the gas profile {}'.format(cloud.shape[1], L))
    ctx = ctx or _lib.get_context()
    ctx.use_own_stream()
    formalisms = list(formalisms)
    if per_layer and any(name == 'h2_orton' for _, name in formalisms):
        raise NotImplementedError('h2_orton prepares its table per frequency list: no per-layer frequencies')
    prepare_catalogs(ctx, formalisms, truncate_strength, truncate_freq, freqs=freqs[0] if per_layer else freqs,
                     other_dicts=other_dicts)
    d = build_alpha_desc(formalisms, L, F, gas.shape[0], gas_dict, 0 if cloud is None else cloud.shape[0], cloud_dict,
                         other_dicts, units)
    sm = scale_matrix(scale, [c for c, _ in formalisms], L)
    sm = None if sm is None else f64(sm)
    d.freqs, d.T, d.P, d.gas = ptr(freqs), ptr(T), ptr(P), ptr(gas)
    d.freqs_per_layer = 1 if per_layer else 0
    d.cloud = ptr(cloud)
    d.scale = ptr(sm)
    if _resident is not None:
        _claim_resident(_resident[0])
        sg, cg = C.c_uint64(0), C.c_uint64(0)
        ctx.check(ctx.lib.rb_alpha_layers_resident(ctx.h, C.byref(d), 1 if want_cube else 0, C.byref(sg), C.byref(cg)))
        return ResidentSlab(ctx, L, F, sg.value, (L, F, len(formalisms)) if want_cube else None, cg.value)
    total = np.empty((L, F))
    cube = np.empty((L, F, len(formalisms))) if want_cube else None
    ctx.check(ctx.lib.rb_alpha_layers(ctx.h, C.byref(d), ptr(total), ptr(cube)))
    return (total, cube) if want_cube else total


def alpha_layers_dev(freqs_t, T_t, P_t, gas_t, gas_dict, cloud_t=None, cloud_dict=None, formalisms=(), other_dicts=None,
                     units='invcm', scale_t=None, want_cube=False, truncate_strength=None, truncate_freq=None, ctx=None,
                     out=None, freqs_host=None, scatter=None):
    """Device-resident variant: torch float64 CUDA tensors in, slab[L][F] CUDA tensor out (async).

    scatter = (peer_ptrs, first_row): a layer-sharded run -- the inputs describe this rank's block of layers and the
    kernel stores its values into rows first_row.. of the full slab on every GPU (device addresses `peer_ptrs`, e.g.
    `torch.distributed._symmetric_memory` buffer_ptrs) instead of into `out`; returns None."""
    import torch
    ctx = ctx or _lib.get_context()
    L, F = T_t.shape[0], freqs_t.shape[0]
    formalisms = list(formalisms)
    if any(name == 'h2_orton' for _, name in formalisms) and freqs_host is None:
        freqs_host = freqs_t.cpu().numpy()          # the Orton table is prepared on the host per frequency vector
    prepare_catalogs(ctx, formalisms, truncate_strength, truncate_freq, freqs=freqs_host, other_dicts=other_dicts)
    if gas_t.shape[1] != L or P_t.shape[0] != L or (cloud_t is not None and cloud_t.shape[1] != L):
        raise ValueError('T, P, gas and cloud disagree on the number of layers')
    d = build_alpha_desc(formalisms, L, F, gas_t.shape[0], gas_dict, 0 if cloud_t is None else cloud_t.shape[0],
                         cloud_dict, other_dicts, units)
    for t in (freqs_t, T_t, P_t, gas_t):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
    d.freqs, d.T, d.P, d.gas = freqs_t.data_ptr(), T_t.data_ptr(), P_t.data_ptr(), gas_t.data_ptr()
    d.cloud = cloud_t.data_ptr() if cloud_t is not None else None
    d.scale = scale_t.data_ptr() if scale_t is not None else None
    if freqs_host is not None:
        freqs_host = f64(freqs_host)
        d.freqs_host = ptr(freqs_host)
    ctx.set_stream(torch.cuda.current_stream(T_t.device).cuda_stream)
    if scatter is not None:
        ptrs, first_row = scatter
        arr = (C.c_uint64 * len(ptrs))(*[int(x) for x in ptrs])
        ctx.check(ctx.lib.rb_alpha_layers_dev_scatter(ctx.h, C.byref(d), len(ptrs), arr, int(first_row)))
        return None
    total = out if out is not None else torch.empty((L, F), dtype=torch.float64, device=T_t.device)
    cube = torch.empty((L, F, len(formalisms)), dtype=torch.float64, device=T_t.device) if want_cube else None
    ctx.check(ctx.lib.rb_alpha_layers_dev(ctx.h, C.byref(d), total.data_ptr(), cube.data_ptr() if want_cube else None))
    return (total, cube) if want_cube else total


def alpha_scale_sum(cube, scale_mat=None, want_cube=False, ctx=None):
    """Scale-sum of a cached per-constituent cube [L][F][C] -> slab[L][F] (+ the scaled cube)."""
    ctx = ctx or _lib.get_context()
    ctx.use_own_stream()
    cube = f64(cube)
    L, F, C_ = cube.shape
    sm = None if scale_mat is None else f64(scale_mat)
    total = np.empty((L, F))
    scaled = np.empty((L, F, C_)) if want_cube else None
    ctx.check(ctx.lib.rb_alpha_scale_sum(ctx.h, L, F, C_, ptr(cube), ptr(sm), ptr(total), ptr(scaled)))
    return (total, scaled) if want_cube else total


GTYPE = {'ellipse': 0, 'circle': 1, 'sphere': 1, 'gravity': 2}
_gravity_key = {}      # context handle -> digest of the gravity model whose shape table is on the device


def set_gravity_model(radius, GM_profile, Jn, RJ, omega_m, vwlat, vwdat, latstep=0.01, max_lat=90.0, ctx=None):
    """The 'gravity' shape (Shape._calcGeoid, shape.py:141-221) for a radius / GM profile: builds the table of every
    shape the geoid march can return on the device (once per model: a digest of the inputs is remembered)."""
    import hashlib
    ctx = ctx or _lib.get_context()
    radius, GMp = f64(radius), f64(GM_profile)
    Jn, vwlat, vwdat = f64(np.atleast_1d(Jn)), f64(np.atleast_1d(vwlat)), f64(np.atleast_1d(vwdat))
    h = hashlib.sha1()
    for a in (radius, GMp, Jn, vwlat, vwdat, np.array([RJ, omega_m, latstep, max_lat], dtype=np.float64)):
        h.update(a.tobytes())
    if _gravity_key.get(ctx.h.value) == h.digest():
        return
    # GM of the march that starts at radius[l]: np.interp exactly as shape.py:156-157 calls it (the radius profile
    # decreases with the index; the reference passes it to np.interp as it is)
    GM_layer = f64(np.array([np.interp(r, radius, GMp) for r in radius]))
    m = _lib.GravityModel()
    m.n_layers, m.radius, m.GM_layer = len(radius), ptr(radius), ptr(GM_layer)
    m.n_J, m.Jn, m.RJ, m.omega_m = len(Jn), ptr(Jn), float(RJ), float(omega_m)
    m.n_vw, m.vwlat, m.vwdat = len(vwlat), ptr(vwlat), ptr(vwdat)
    m.latstep, m.max_lat = float(latstep), float(max_lat)
    ctx.use_own_stream()
    ctx.check(ctx.lib.rb_set_gravity_model(ctx.h, C.byref(m)))
    _gravity_key[ctx.h.value] = h.digest()
LIMB = {'shape': 0, 'sec': 1}


def build_geometry_desc(L, n0, n1, Req, Rpol, orientation, gtype, limb, gravity_model=None, radius=None):
    if gtype == 'gravity':
        if gravity_model is None:
            raise ValueError("gtype 'gravity' needs the gravity model of the planet (GM profile, Jn, RJ, omega_m, zonal winds)")
        set_gravity_model(radius, **gravity_model)
    if gtype not in GTYPE:
        raise NotImplementedError("gtype '{}' is not built (ellipse / circle / sphere / gravity; 'reference' raises in the reference too, shape.py:107)".format(gtype))
    g = GeometryDesc()
    g.n_layers = L
    g.n0, g.n1 = float(n0), float(n1)
    g.Req, g.Rpol = float(Req), float(Rpol)
    g.orientation[0], g.orientation[1] = float(orientation[0]), float(orientation[1])
    g.gtype = GTYPE[gtype]
    g.limb = LIMB.get(limb, 0)
    return g


def compute_ds(radius, refr_index, b, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape', ctx=None, gravity_model=None):
    """raypath.compute_ds for a batch of impact points.  Returns ds[R][L-1], nseg[R], (tip, rotate, rNorm)."""
    ctx = ctx or _lib.get_context()
    ctx.use_own_stream()
    radius = f64(radius)
    b = f64(np.atleast_2d(b))
    R, L = b.shape[0], radius.shape[0]
    g = build_geometry_desc(L, refr_index[0], refr_index[1], Req, Rpol, orientation, gtype, limb, gravity_model, radius)
    g.radius = ptr(radius)
    ds = np.empty((R, L - 1))
    nseg = np.empty(R, dtype=np.int32)
    aspect = np.empty(3)
    ctx.check(ctx.lib.rb_compute_ds(ctx.h, C.byref(g), R, ptr(b), ptr(ds), ptr(nseg), ptr(aspect)))
    return ds, nseg, aspect


def compute_ray_fields(radius, refr_index, b, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape', ctx=None, gravity_model=None):
    """Per step of every ray: Ray.r4ds (km) and the planetocentric latitude / longitude (deg) of the point the step starts
    at (raypath.py:186-187, 224).  Returns fields[R][3][L-1]."""
    ctx = ctx or _lib.get_context()
    ctx.use_own_stream()
    radius = f64(radius)
    b = f64(np.atleast_2d(b))
    R, L = b.shape[0], radius.shape[0]
    g = build_geometry_desc(L, refr_index[0], refr_index[1], Req, Rpol, orientation, gtype, limb, gravity_model, radius)
    g.radius = ptr(radius)
    out = np.empty((R, 3, L - 1))
    ctx.check(ctx.lib.rb_compute_ray_fields(ctx.h, C.byref(g), R, ptr(b), ptr(out)))
    return out


# A ray stops once tau > TAU_CUT.  Every later term of the two sums is below e^-50 (2e-22) x T (<= 2000 K) x dtau,
# i.e. < 1e-16: less than half an ulp of the accumulated sums (integrated_W ~ 1, Tb ~ 100 K), so adding it would not
# change either accumulator -- the cut result is bit-identical to integrating every layer like the reference
# (asserted in tests/test_gpu_rt.py).  tau_cut=0 disables the cut.
TAU_CUT = 50.0


def geometry_prefetch(radius, refr_index, b, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape', ctx=None, gravity_model=None):
    """Start the ray geometry of the next rt_batch(b=<the same array>, same geometry) now so that it overlaps
    the absorption kernel (rb_geometry_prefetch).  `radius` and `b` must be the very arrays (same memory) later
    given to rt_batch and must stay unchanged until then; returns the (radius, b) pair to pass on."""
    ctx = ctx or _lib.get_context()
    ctx.use_own_stream()
    radius = f64(radius)
    b = f64(np.atleast_2d(b))
    g = build_geometry_desc(radius.shape[0], refr_index[0], refr_index[1], Req, Rpol, orientation, gtype, limb, gravity_model,
                            radius)
    g.radius = ptr(radius)
    ctx.check(ctx.lib.rb_geometry_prefetch(ctx.h, C.byref(g), b.shape[0], ptr(b)))
    return radius, b


def geometry_prefetch_dev(radius_t, n0, n1, b_t, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape', ctx=None, gravity_model=None):
    """Device-resident variant of geometry_prefetch (torch CUDA tensors; pairs with rt_batch_dev)."""
    import torch
    ctx = ctx or _lib.get_context()
    g = build_geometry_desc(radius_t.shape[0], n0, n1, Req, Rpol, orientation, gtype, limb, gravity_model,
                            radius_t.cpu().numpy() if gtype == 'gravity' else None)
    g.radius = radius_t.data_ptr()
    ctx.set_stream(torch.cuda.current_stream(b_t.device).cuda_stream)
    ctx.check(ctx.lib.rb_geometry_prefetch_dev(ctx.h, C.byref(g), b_t.shape[0], b_t.data_ptr()))


def rt_batch(radius, refr_index, b, alpha_slab, T, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape',
             disc_average=False, out_f32=False, tau_cut=TAU_CUT, want_intW=False, profile_ray=-1, ctx=None, out=None,
             gravity_model=None):
    """Brightness.single over a batch of rays: Tb[R][F] (+ integrated_W, + profiles of one ray)."""
    resident = isinstance(alpha_slab, ResidentSlab)
    ctx = ctx or (alpha_slab.ctx if resident else _lib.get_context())
    ctx.use_own_stream()
    radius, T = f64(radius), f64(T)
    if resident:
        if not alpha_slab.valid():
            raise RuntimeError('the resident absorption slab has been overwritten')
    else:
        alpha_slab = f64(alpha_slab)
    b = f64(np.atleast_2d(b))
    R, L, F = b.shape[0], radius.shape[0], alpha_slab.shape[1]
    if alpha_slab.shape[0] != L or T.shape[0] != L:
        raise ValueError('alpha slab must be [L][F] with L = number of layers')
    g = build_geometry_desc(L, refr_index[0], refr_index[1], Req, Rpol, orientation, gtype, limb, gravity_model, radius)
    g.radius = ptr(radius)
    rt = RtDesc()
    rt.n_freqs, rt.alpha, rt.T = F, (None if resident else ptr(alpha_slab)), ptr(T)
    rt.disc_average, rt.out_f32, rt.tau_cut = int(bool(disc_average)), int(bool(out_f32)), float(tau_cut or 0.0)
    if out is None:
        out = pinned_pool.get((R, F), np.float32 if out_f32 else np.float64)
    elif out.shape != (R, F) or out.dtype != (np.float32 if out_f32 else np.float64) or not out.flags.c_contiguous:
        raise ValueError('rt_batch: out must be a C-contiguous [R][F] array of the output dtype')
    intW = np.empty((R, F)) if want_intW else None
    prof = None
    if profile_ray >= 0:
        prof = [np.zeros((F, L - 1)) for _ in range(3)]
    call = ctx.lib.rb_rt_batch_resident if resident else ctx.lib.rb_rt_batch
    ctx.check(call(ctx.h, C.byref(g), C.byref(rt), R, ptr(b), ptr(out), ptr(intW), int(profile_ray),
                   ptr(prof[0]) if prof else None, ptr(prof[1]) if prof else None, ptr(prof[2]) if prof else None))
    res = {'Tb': out}
    if want_intW:
        res['integrated_W'] = intW
    if prof:
        res['tau'], res['W'], res['Tb_lyr'] = prof
    return res


def rt_batch_dev(radius_t, n0, n1, b_t, alpha_t, T_t, Req, Rpol, orientation=(0.0, 0.0), gtype='ellipse', limb='shape',
                 disc_average=False, out_f32=True, tau_cut=TAU_CUT, ctx=None, out=None, gravity_model=None):
    """Device-resident variant (torch CUDA tensors, async on the current stream)."""
    import torch
    ctx = ctx or _lib.get_context()
    R, L, F = b_t.shape[0], radius_t.shape[0], alpha_t.shape[1]
    g = build_geometry_desc(L, n0, n1, Req, Rpol, orientation, gtype, limb, gravity_model,
                            radius_t.cpu().numpy() if gtype == 'gravity' else None)
    g.radius = radius_t.data_ptr()
    rt = RtDesc()
    rt.n_freqs, rt.alpha, rt.T = F, alpha_t.data_ptr(), T_t.data_ptr()
    rt.disc_average, rt.out_f32, rt.tau_cut = int(bool(disc_average)), int(bool(out_f32)), float(tau_cut or 0.0)
    if out is None:
        out = torch.empty((R, F), dtype=torch.float32 if out_f32 else torch.float64, device=b_t.device)
    ctx.set_stream(torch.cuda.current_stream(b_t.device).cuda_stream)
    ctx.check(ctx.lib.rb_rt_batch_dev(ctx.h, C.byref(g), C.byref(rt), R, b_t.data_ptr(), out.data_ptr(), None))
    return out


def rt_integrate(ds, nseg, alpha_slab, T, disc_average=False, out_f32=False, tau_cut=TAU_CUT, want_intW=False, ctx=None,
                 alpha0_slab=None, profile_ray=-1):
    """Integration only, for caller-supplied segments ds[R][L-1] (km).

    alpha0_slab: the Doppler form -- dtau of step i is (alpha0[i] + alpha[i+1]) ds_i / 2 (rb_rt_desc::alpha0).
    profile_ray >= 0: returns a dict with Tb, integrated_W and the profiles tau / W / Tb_lyr [F][L-1] of that ray."""
    ctx = ctx or _lib.get_context()
    ctx.use_own_stream()
    ds = f64(np.atleast_2d(ds))
    alpha_slab, T = f64(alpha_slab), f64(T)
    R, S = ds.shape
    L, F = alpha_slab.shape
    nseg = np.ascontiguousarray(nseg, dtype=np.int32)
    rt = RtDesc()
    rt.n_freqs, rt.alpha, rt.T = F, ptr(alpha_slab), ptr(T)
    rt.disc_average, rt.out_f32, rt.tau_cut = int(bool(disc_average)), int(bool(out_f32)), float(tau_cut or 0.0)
    if alpha0_slab is not None:
        alpha0_slab = f64(alpha0_slab)
        if alpha0_slab.shape != alpha_slab.shape:
            raise ValueError('alpha0_slab must have the shape of alpha_slab')
        rt.alpha0 = ptr(alpha0_slab)
    out = np.empty((R, F), dtype=np.float32 if out_f32 else np.float64)
    intW = np.empty((R, F)) if (want_intW or profile_ray >= 0) else None
    if profile_ray >= 0:
        tau, W, Tbl = np.zeros((F, S)), np.zeros((F, S)), np.zeros((F, S))
        ctx.check(ctx.lib.rb_rt_integrate_profile(ctx.h, C.byref(rt), L, R, S, ptr(ds), ptr(nseg), ptr(out), ptr(intW),
                                                  int(profile_ray), ptr(tau), ptr(W), ptr(Tbl)))
        return {'Tb': out, 'integrated_W': intW, 'tau': tau, 'W': W, 'Tb_lyr': Tbl}
    ctx.check(ctx.lib.rb_rt_integrate(ctx.h, C.byref(rt), L, R, S, ptr(ds), ptr(nseg), ptr(out), ptr(intW)))
    return (out, intW) if want_intW else out
