"""ctypes binding of the C ABI in include/radiobear_b200.h (libradiobear_b200.so).

There is NO CPU fallback: a missing library or a missing sm_100 GPU raises.  Build the library
with ``python -c "import __graft_entry__ as g; g.build()"`` (or ``python -m radiobear_b200.build``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RB_LIB_PATH: another build of the same library (tools/ab_quick.py compares build variants on the GPU box)
LIB_PATH = os.environ.get('RB_LIB_PATH') or os.path.join(_HERE, 'lib', 'libradiobear_b200.so')

RB_OK, RB_ERR_INVALID, RB_ERR_CUDA, RB_ERR_NOMEM, RB_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
RB_MAX_CONSTITUENTS = 8
RB_NUM_GAS = 8
RB_NUM_CLD = 6

CATALOG_IDS = {'nh3_inv': 0, 'nh3_rot': 1, 'nh3_v2': 2, 'nh3_sjs': 3, 'h2s': 4, 'ph3': 5, 'co': 6, 'h2o': 7,
               'h2_orton': 8}
FORMALISM_IDS = {'nh3_hs': 1, 'nh3_dbs': 2, 'nh3_sjs': 3, 'nh3_hs_sjs': 4, 'nh3_dbs_sjs': 5, 'h2s_ddb': 6,
                 'ph3_jh': 7, 'h2o_bk': 8, 'h2_jj_ddb': 9, 'h2_jj': 10, 'clouds_idp': 11, 'co_ddb': 12,
                 'nh3_kd': 13, 'nh3_sjsd': 14, 'nh3_bg': 15, 'h2_orton': 16}
GAS_ORDER = ['H2', 'HE', 'CH4', 'NH3', 'H2O', 'H2S', 'PH3', 'CO']           # RB_GAS_*
CLOUD_ORDER = ['H2O', 'SOLN', 'NH4SH', 'NH3', 'H2S', 'CH4']                 # RB_CLD_*
CLOUD_FLAG_KEYS = ['ice_p', 'water_p', 'nh4sh_p', 'nh3ice_p', 'h2sice_p', 'ch4']   # clouds_idp.py:17-45

RT_PRECISIONS = {'f64': 0, 'mixed': 1}                                      # RB_RT_*

_dp = C.POINTER(C.c_double)


class AlphaDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('n_freqs', C.c_int32), ('n_constituents', C.c_int32),
                ('formalism', C.c_int32 * RB_MAX_CONSTITUENTS),
                ('freqs', C.c_void_p), ('T', C.c_void_p), ('P', C.c_void_p),
                ('gas', C.c_void_p), ('gas_rows', C.c_int32), ('gas_col', C.c_int32 * RB_NUM_GAS),
                ('cloud', C.c_void_p), ('cloud_rows', C.c_int32), ('cloud_col', C.c_int32 * RB_NUM_CLD),
                ('cloud_flags', C.c_uint32), ('h2state', C.c_int32), ('coshape', C.c_int32),
                ('units', C.c_int32), ('scale', C.c_void_p), ('freqs_host', C.c_void_p), ('freqs_per_layer', C.c_int32)]


class GeometryDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('radius', C.c_void_p), ('n0', C.c_double), ('n1', C.c_double),
                ('Req', C.c_double), ('Rpol', C.c_double), ('orientation', C.c_double * 2),
                ('gtype', C.c_int32), ('limb', C.c_int32)]


class GravityModel(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('radius', C.c_void_p), ('GM_layer', C.c_void_p), ('n_J', C.c_int32),
                ('Jn', C.c_void_p), ('RJ', C.c_double), ('omega_m', C.c_double), ('n_vw', C.c_int32),
                ('vwlat', C.c_void_p), ('vwdat', C.c_void_p), ('latstep', C.c_double), ('max_lat', C.c_double)]


class RtDesc(C.Structure):
    _fields_ = [('n_freqs', C.c_int32), ('alpha', C.c_void_p), ('T', C.c_void_p),
                ('disc_average', C.c_int32), ('out_f32', C.c_int32), ('tau_cut', C.c_double), ('alpha0', C.c_void_p)]


class RadiobearB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RadiobearB200Error(
            'radiobear_b200: CUDA library {} is missing and there is no CPU fallback. '
            'Build it with: python -c "import __graft_entry__ as g; g.build()"'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        'rb_abi_version': (C.c_int, []),
        'rb_create': (C.c_int, [C.c_int, C.POINTER(vp)]),
        'rb_destroy': (None, [vp]),
        'rb_last_error': (C.c_char_p, [vp]),
        'rb_set_stream': (C.c_int, [vp, vp]),
        'rb_use_own_stream': (C.c_int, [vp]),
        'rb_synchronize': (C.c_int, [vp]),
        'rb_launch_count': (i64, [vp]),
        'rb_enable_timing': (C.c_int, [vp, C.c_int]),
        'rb_last_kernel_ms': (dbl, [vp, C.c_int]),
        'rb_kernel_ms_history': (C.c_int, [vp, C.c_int, vp, C.c_int]),
        'rb_kernel_timed_count': (i64, [vp, C.c_int]),
        'rb_set_rt_chunks': (C.c_int, [vp, C.c_int]),
        'rb_set_rt_precision': (C.c_int, [vp, C.c_int]),
        'rb_get_rt_precision': (C.c_int, [vp]),
        'rb_set_rt_tuning': (C.c_int, [vp, C.c_int, C.c_int]),
        'rb_set_rt_stream_geometry': (C.c_int, [vp, C.c_int]),
        'rb_count_steps': (i64, [vp, C.c_int]),
        'rb_count_small_steps': (i64, [vp]),
        'rb_set_catalog': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
        'rb_alpha_layers': (C.c_int, [vp, C.POINTER(AlphaDesc), vp, vp]),
        'rb_alpha_layers_dev': (C.c_int, [vp, C.POINTER(AlphaDesc), vp, vp]),
        'rb_alpha_scale_sum': (C.c_int, [vp, i32, i32, i32, vp, vp, vp, vp]),
        'rb_alpha_layers_dev_scatter': (C.c_int, [vp, C.POINTER(AlphaDesc), i32, C.POINTER(C.c_uint64), i64]),
        'rb_alpha_layers_resident': (C.c_int, [vp, C.POINTER(AlphaDesc), i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
        'rb_alpha_rescale_resident': (C.c_int, [vp, vp, C.POINTER(C.c_uint64)]),
        'rb_alpha_resident_info': (C.c_int, [vp, vp, C.POINTER(C.c_uint64), vp, C.POINTER(C.c_uint64)]),
        'rb_alpha_fetch': (C.c_int, [vp, vp, vp]),
        'rb_rt_batch_resident': (C.c_int, [vp, C.POINTER(GeometryDesc), C.POINTER(RtDesc), i64, vp, vp, vp, i64, vp, vp, vp]),
        'rb_compute_ds': (C.c_int, [vp, C.POINTER(GeometryDesc), i64, vp, vp, vp, vp]),
        'rb_set_gravity_model': (C.c_int, [vp, C.POINTER(GravityModel)]),
        'rb_compute_ray_fields': (C.c_int, [vp, C.POINTER(GeometryDesc), i64, vp, vp]),
        'rb_rt_batch': (C.c_int, [vp, C.POINTER(GeometryDesc), C.POINTER(RtDesc), i64, vp, vp, vp, i64, vp, vp, vp]),
        'rb_rt_batch_dev': (C.c_int, [vp, C.POINTER(GeometryDesc), C.POINTER(RtDesc), i64, vp, vp, vp]),
        'rb_geometry_prefetch': (C.c_int, [vp, C.POINTER(GeometryDesc), i64, vp]),
        'rb_geometry_prefetch_dev': (C.c_int, [vp, C.POINTER(GeometryDesc), i64, vp]),
        'rb_rt_integrate': (C.c_int, [vp, C.POINTER(RtDesc), i32, i64, i32, vp, vp, vp, vp]),
        'rb_rt_integrate_profile': (C.c_int, [vp, C.POINTER(RtDesc), i32, i64, i32, vp, vp, vp, vp, i64, vp, vp, vp]),
        'rb_probe_fp64_peak': (C.c_int, [vp, C.c_int, C.POINTER(dbl)]),
        'rb_probe_rcp': (C.c_int, [vp, C.c_int, C.c_int, vp, vp]),
        'rb_probe_fp64_mix': (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(dbl)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = ['rb_abi_version', 'rb_create', 'rb_destroy', 'rb_last_error', 'rb_set_stream', 'rb_use_own_stream',
                    'rb_synchronize',
                    'rb_launch_count', 'rb_enable_timing', 'rb_last_kernel_ms', 'rb_kernel_ms_history', 'rb_kernel_timed_count',
                    'rb_set_rt_chunks', 'rb_set_rt_precision', 'rb_get_rt_precision', 'rb_set_rt_tuning', 'rb_set_rt_stream_geometry', 'rb_count_steps', 'rb_count_small_steps', 'rb_set_catalog', 'rb_alpha_layers',
                    'rb_alpha_layers_dev', 'rb_alpha_scale_sum', 'rb_alpha_layers_dev_scatter', 'rb_alpha_layers_resident', 'rb_alpha_rescale_resident',
                    'rb_alpha_resident_info', 'rb_alpha_fetch', 'rb_rt_batch_resident', 'rb_compute_ds', 'rb_compute_ray_fields', 'rb_set_gravity_model', 'rb_geometry_prefetch', 'rb_geometry_prefetch_dev', 'rb_rt_batch', 'rb_rt_batch_dev', 'rb_rt_integrate', 'rb_rt_integrate_profile',
                    'rb_probe_fp64_peak', 'rb_probe_rcp', 'rb_probe_fp64_mix']

_EXC = {RB_ERR_INVALID: ValueError, RB_ERR_CUDA: RadiobearB200Error, RB_ERR_NOMEM: MemoryError,
        RB_ERR_UNSUPPORTED: NotImplementedError}


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One rb_context: a GPU, a stream, device-resident line catalogs and scratch buffers."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        st = self.lib.rb_create(int(device), C.byref(self.h))
        if st != RB_OK:
            msg = self.lib.rb_last_error(self.h).decode() if self.h else 'rb_create failed'
            if self.h:
                self.lib.rb_destroy(self.h)
                self.h = C.c_void_p()
            raise _EXC.get(st, RadiobearB200Error)(msg)
        self.device = int(device)
        self.catalog_key = {}

    def check(self, st):
        if st != RB_OK:
            raise _EXC.get(st, RadiobearB200Error)(self.lib.rb_last_error(self.h).decode())

    def close(self):
        if getattr(self, 'h', None):
            self.lib.rb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def set_stream(self, cuda_stream):
        """Launch on the given cudaStream_t handle (0 = the legacy default stream, e.g. torch's default)."""
        self.check(self.lib.rb_set_stream(self.h, C.c_void_p(int(cuda_stream))))

    def use_own_stream(self):
        self.check(self.lib.rb_use_own_stream(self.h))

    def synchronize(self):
        self.check(self.lib.rb_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.rb_launch_count(self.h))

    def enable_timing(self, on=True):
        self.check(self.lib.rb_enable_timing(self.h, 1 if on else 0))

    def last_kernel_ms(self, which):
        return float(self.lib.rb_last_kernel_ms(self.h, {'alpha': 0, 'geometry': 1, 'rt': 2}.get(which, which)))

    def kernel_timed_count(self, which):
        return int(self.lib.rb_kernel_timed_count(self.h, {'alpha': 0, 'geometry': 1, 'rt': 2}.get(which, which)))

    def count_steps(self, enable):
        """Start (True) / stop (False) counting integrated segment-steps; returns the count so far."""
        return int(self.lib.rb_count_steps(self.h, 1 if enable else 0))

    def count_small_steps(self):
        """Of the last count_steps() result: the steps of the small-tau phase."""
        return int(self.lib.rb_count_small_steps(self.h))

    def set_rt_chunks(self, n):
        self.check(self.lib.rb_set_rt_chunks(self.h, int(n)))

    def set_rt_precision(self, precision):
        """'f64' or 'mixed' (include/radiobear_b200.h: RB_RT_F64 / RB_RT_MIXED) for the batched ray integration."""
        if precision not in RT_PRECISIONS:
            raise ValueError("rt precision must be one of {}".format(sorted(RT_PRECISIONS)))
        self.check(self.lib.rb_set_rt_precision(self.h, RT_PRECISIONS[precision]))

    def set_rt_stream_geometry(self, mode=-1):
        """Integration follows a prefetched trace that is still running: 1 always, 0 never, -1 automatic
        (include/radiobear_b200.h: rb_set_rt_stream_geometry)."""
        self.check(self.lib.rb_set_rt_stream_geometry(self.h, int(mode)))

    def set_rt_tuning(self, pairs=-1, compact=True):
        """Work decomposition of the FP64 ray integration (results are bit-identical): pairs -1 auto / 0 / 1,
        compact: integrate only the rays that hit the planet (include/radiobear_b200.h: rb_set_rt_tuning)."""
        self.check(self.lib.rb_set_rt_tuning(self.h, int(pairs), 1 if compact else 0))

    def rt_precision(self):
        code = int(self.lib.rb_get_rt_precision(self.h))
        return {v: k for k, v in RT_PRECISIONS.items()}[code]

    def kernel_ms_history(self, which, n=256):
        buf = np.zeros(n)
        got = self.lib.rb_kernel_ms_history(self.h, {'alpha': 0, 'geometry': 1, 'rt': 2}.get(which, which), ptr(buf), n)
        return buf[:got]

    def fp64_peak_tflops(self, iters=20000):
        out = C.c_double(0.0)
        self.check(self.lib.rb_probe_fp64_peak(self.h, int(iters), C.byref(out)))
        return out.value

    def probe_rcp(self, x, newton):
        x = f64(x)
        y = np.empty_like(x)
        self.check(self.lib.rb_probe_rcp(self.h, int(newton), x.size, ptr(x), ptr(y)))
        return y

    def set_catalog(self, name, cols, key=None):
        """cols: [ncols][nlines] float64.  `key` lets callers skip re-uploads of an identical table."""
        if key is not None and self.catalog_key.get(name) == key:
            return
        a = f64(cols)
        if a.ndim != 2:
            raise ValueError('catalog must be 2-D [ncols][nlines]')
        self.check(self.lib.rb_set_catalog(self.h, CATALOG_IDS[name], a.shape[1], a.shape[0], ptr(a)))
        self.catalog_key[name] = key


_contexts = {}


def get_context(device=None):
    """Process-wide context per device (created on first use)."""
    if device is None:
        device = int(os.environ.get('RB_DEVICE', os.environ.get('LOCAL_RANK', '0')))
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
